// conv_gemm_tc.cu - the embedding / quality networks of the key-frame aggregation on Blackwell tensor cores.
//
// `get_embednet` (SYM:118-130: em_conv1 1x1 1024->512 +ReLU, em_conv2 3x3 512->512
// +ReLU, em_conv3 1x1 512->2048), `compute_weight` (SYM:111-116: channel L2 norm + dot) and the convolutions of
// `Nq_net` (SYM:94-101: Nq_conv1 3x3 1024->256 +ReLU, Nq_conv2 1x1 256->16 +ReLU, Nq_conv3 1x1 16->1).
// SYM = dff_rfcn/symbols/resnet_v1_101_flownet_rfcn.py.
//
// One warp-specialised, persistent implicit-GEMM kernel (bf16 operands, fp32 accumulation in tensor memory):
//   * activations are channels-last bf16 (NB,H,W,Cin) read through a 4-D TMA tensor map (C,W,H,NB): a tile of
//     128 output pixels is a BW x BH window of one image, and tap (dy,dx) of a 3x3 convolution is the SAME box
//     shifted by (dx-1,dy-1) - the copy engine zero-fills what falls outside the image, which IS the
//     convolution's zero padding: no im2col buffer, no halo code.  A 1x1 convolution is the 1-tap case.
//   * weights are (Cout, taps*Cin) bf16, K-major, read through a 2-D tensor map.
//   * every CTA works on a PAIR of 128-pixel tiles that share each weight tile: the same window of image b and
//     of image b + NB/2.  The reference always convolves Concat_0(cur, warp) (SYM:95,133), so the pair is
//     "this pixel of the current feature and of the warped feature".  Two M=128 accumulators of N=256 fp32
//     columns fill the SM's 512 tensor-memory columns; per 64-deep k-step the CTA loads 2 x 16 KB of
//     activations + 32 KB of weights for 2 x 128 x 256 x 64 MACs (128 flop per byte of L2 traffic).
//   * warp 0 = TMA producer (one lane), warp 1 = tcgen05.mma issuer (one lane), warp 2 = tensor-memory
//     allocator, warps 4-11 = epilogue (two per tensor-memory lane quarter): thread r owns accumulator row r
//     (tcgen05.ld 32x32b), so every per-pixel reduction over channels is a private register sum - no shuffles.
//   * conv_gemm_tc2_kernel (further down) is the CTA-pair form of the same work item for the STORE / NQ epilogues:
//     tcgen05.mma.cta_group::2, half of the weight tile per CTA, double-buffered accumulators.
//   * epilogues: bias(+ReLU) -> bf16 store (em_conv1/2);  COSINE (em_conv3): sum e_cur^2, sum e_warp^2,
//     sum e_warp*e_cur over this CTA's 256 output channels -> 3 floats per pixel; the 2048-channel embeddings
//     never leave the SM (the reference round-trips 2 x 19.6 MB per frame through HBM for them);
//     NQ (Nq_conv1): bias+ReLU, then the 256->16 (+ReLU) ->1 tail per pixel in registers -> the logit itself.
//
// SASS of this file shows UTCHMMA / UTCHMMA.2CTA (tcgen05.mma), UTCBAR.2CTA.MULTICAST (tcgen05.commit), LDTM (tcgen05.ld),
// UTMALDG / UTMALDG.2CTA (cp.async.bulk.tensor), SYNCS (mbarrier).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "lsfa_device.cuh"
#include "conv_gemm_tc.h"

namespace lsfa {
namespace tc {

constexpr int BM = 128;                 // pixels per accumulator = tensor-memory lanes
constexpr int BN = 256;                 // output channels per work item = fp32 columns per accumulator
constexpr int BK = 64;                  // bf16 per k-step = 128 B = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int STAGES = 3;
constexpr int A_BYTES = BM * BK * 2;    // 16 KB
constexpr int B_BYTES = BN * BK * 2;    // 32 KB
constexpr int STAGE_BYTES = 2 * A_BYTES + B_BYTES;   // 64 KB
constexpr int NQ_MID = 16;              // Nq_conv2 width (SYM:99)
constexpr int EPI_WARPS = 8;            // two epilogue warps per tensor-memory lane quarter (see the epilogue)
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int EPI_TILE_FLOATS = EPI_WARPS * 1024;    // STORE: a 4 KB transpose tile per epilogue warp (NQ: the 16 x 256 tail weights)
constexpr int EPI_SMEM_FLOATS = BN + EPI_TILE_FLOATS + 3 * NQ_MID + 16;
constexpr int SMEM_BYTES = 1024 /*alignment slack*/ + STAGES * STAGE_BYTES + EPI_SMEM_FLOATS * 4 + 128 /*barriers*/;
constexpr int NUM_THREADS = 128 + EPI_THREADS;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
static_assert(NQ_MID * BN <= EPI_TILE_FLOATS, "Nq tail weights share the transpose tiles' space");
constexpr int TMEM_COLS = 512;

// ---------------------------------------------------------------------------------------------------------
// PTX wrappers (tcgen05 / TMA tensor copies).  mbarrier helpers come from lsfa_device.cuh.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32, issued by ONE thread for the CTA
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive once every tcgen05.mma issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp receives row (lane base + t)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
      "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// the same wait, but naming the 32 registers a preceding tmem_ld32 filled as read-write operands: no instruction that reads
// them can be scheduled above the wait (tcgen05.ld is asynchronous: its destinations are valid only after wait::ld)
__device__ __forceinline__ void tmem_ld_wait_for(float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                 "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]),
                 "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]),
                 "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory"); }

// shared-memory matrix descriptor, K-major operand whose rows are 128 B (64 bf16) in the SWIZZLE_128B pattern the TMA
// wrote: 8-row atoms of 1024 B, stride between atoms (SBO) 1024 B, LBO unused for swizzled K-major; version 1 (sm_100).
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);          // start address, bits [0,14)
  d |= (uint64_t)(1024 >> 4) << 32;                    // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                              // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                              // layout type SWIZZLE_128B
  return d;
}
// instruction descriptor of kind::f16: D fp32, A/B bf16, both K-major, M = 128, N = BN
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n) {
  return (1u << 4)                 // c_format  = F32
         | (1u << 7)               // a_format  = BF16
         | (1u << 10)              // b_format  = BF16
         | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// A work item is a PAIR of pixel tiles (image b and image b + NB/2, one weight tile for both).  The pair items that would
// form a last, partly filled wave of the persistent grid are cut into two SINGLE-tile items each (half the time): with
// T pair items on G CTAs the step costs floor(T/G) + 0.5 waves instead of ceil(T/G) when 2*(T mod G) <= G.
struct Item {
  int chunk, pair, img0, img1, x0, y0;
  int single;     // 1: only accumulator 0 is used, on image img0
  int which;      // single: 0 = the pair's first image, 1 = its second
};
__device__ __forceinline__ Item decode_item(const ConvParams& P, int item) {
  Item it;
  it.single = item >= P.pair_items;
  it.which = 0;
  if (it.single) {
    const int j = item - P.pair_items;
    it.which = j & 1;
    item = P.pair_items + (j >> 1);
  }
  it.chunk = item % P.n_chunks;
  int t = item / P.n_chunks;
  const int tile = t % (P.tiles_x * P.tiles_y);
  it.pair = t / (P.tiles_x * P.tiles_y);
  // image b is paired with image b + ceil(NB/2); with an odd image count the last pair's partner is clamped to the last
  // image (computed twice, stored twice with identical values) - only the plain STORE epilogues accept odd counts
  const int half = (P.NB + 1) / 2;
  it.img1 = min(it.pair + half, P.NB - 1);
  it.img0 = it.which ? it.img1 : it.pair;
  it.x0 = (tile % P.tiles_x) * P.BW;
  it.y0 = (tile / P.tiles_x) * P.BH;
  return it;
}

template <int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvParams P) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for SWIZZLE_128B; plain pointer arithmetic on the __shared__ array keeps the address space known
  // to the compiler (an integer round trip made every bias / weight read a generic LD.E instead of LDS)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* s_bias = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);      // [BN]
  float* s_w2t = s_bias + BN;                                                 // NQ: [BN][16] (transposed Nq_conv2 weight)
  float* s_nq = s_w2t + EPI_TILE_FLOATS;                                          // NQ: b2[16], w3[16], b3
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + EPI_SMEM_FLOATS * 4);
  uint64_t* full = bars;                    // [STAGES] TMA -> MMA
  uint64_t* empty = bars + STAGES;          // [STAGES] MMA -> TMA
  uint64_t* tmem_full = bars + 2 * STAGES;  // MMA -> epilogue
  uint64_t* tmem_empty = tmem_full + 1;     // epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_items = P.num_items;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, EPI_THREADS);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
  if (EPI == EPI_NQ && warp >= 4) {          // the per-pixel tail's parameters: once per CTA
    const int t = threadIdx.x - 128;
    for (int i = t; i < NQ_MID * BN; i += EPI_THREADS) {
      const int j = i / BN, c = i % BN;       // global (16, 256) row-major -> smem [c][j]
      s_w2t[c * NQ_MID + j] = P.nq_w2[i];
    }
    if (t < NQ_MID) {
      s_nq[t] = P.nq_b2[t];
      s_nq[NQ_MID + t] = P.nq_w3[t];
    }
    if (t == 0) s_nq[2 * NQ_MID] = P.nq_b3[0];
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const Item it = decode_item(P, item);
        for (int ks = 0; ks < P.k_steps; ++ks) {
          const int tap = ks / P.kc_per_tap, kc = ks - tap * P.kc_per_tap;
          const int dy = P.taps == 9 ? tap / 3 - 1 : 0, dx = P.taps == 9 ? tap % 3 - 1 : 0;
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* st = smem + stage * STAGE_BYTES;
          mbar_expect_tx(&full[stage], it.single ? A_BYTES + B_BYTES : STAGE_BYTES);
          tma_load_4d(st, &tmA, kc * BK, it.x0 + dx, it.y0 + dy, it.img0, &full[stage]);
          if (!it.single) tma_load_4d(st + A_BYTES, &tmA, kc * BK, it.x0 + dx, it.y0 + dy, it.img1, &full[stage]);
          tma_load_2d(st + 2 * A_BYTES, &tmB, tap * P.Cin + kc * BK, it.chunk * BN, &full[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
      uint32_t stage = 0, phase = 0, tphase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const bool single = item >= P.pair_items;
        mbar_wait(tmem_empty, tphase ^ 1);        // the epilogue has drained both accumulators
        tcgen05_fence_after();
        for (int ks = 0; ks < P.k_steps; ++ks) {
          mbar_wait(&full[stage], phase);
          tcgen05_fence_after();
          const uint32_t a0 = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t da0 = make_kmajor_sw128_desc(a0);
          const uint64_t da1 = make_kmajor_sw128_desc(a0 + A_BYTES);
          const uint64_t db = make_kmajor_sw128_desc(a0 + 2 * A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advancing 16 bf16 = 32 B along K inside the 128 B swizzle row: +2 in the (>>4) start-address field
            umma_bf16(tmem_base, da0 + 2 * k, db + 2 * k, idesc, (ks | k) != 0);
            if (!single) umma_bf16(tmem_base + BN, da1 + 2 * k, db + 2 * k, idesc, (ks | k) != 0);
          }
          umma_commit(&empty[stage]);             // frees the stage once these MMAs have read it
          if (ks == P.k_steps - 1) umma_commit(tmem_full);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tphase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: thread r owns accumulator row r =====================
    // EIGHT epilogue warps: warp w may read tensor-memory lanes 32*(w % 4)..+31, so warps 4-7 and 8-11 cover the four
    // quarters twice.  The epilogue is latency bound (clock64 trace: 1,365 cycles per 64-column step for ~300 instructions
    // with ONE warp per scheduler), so the second set halves its exposed time: group 0 takes accumulator 0 and group 1
    // accumulator 1 (STORE, NQ), or each group half of the columns of both (COSINE).
    const int ew = warp & 3;                       // the tensor-memory lane quarter this warp may read
    const int grp = (warp - 4) >> 2;               // 0: warps 4-7, 1: warps 8-11
    const int row = ew * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16);
    uint32_t tphase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const Item it = decode_item(P, item);
      epi_bar_sync();                              // previous item's readers of s_bias are done
      for (int i = threadIdx.x - 128; i < BN; i += EPI_THREADS) s_bias[i] = P.bias[it.chunk * BN + i];
      epi_bar_sync();
      const int x = it.x0 + (row & (P.BW - 1)), y = it.y0 + (row >> P.bw_shift);
      const bool valid = x < P.W && y < P.H;
      const size_t pix = (size_t)y * P.W + x;
      mbar_wait(tmem_full, tphase);
      tcgen05_fence_after();
      if (EPI == EPI_STORE_RELU || EPI == EPI_STORE) {
        const int n_acc = it.single ? 1 : 2;
#pragma unroll 1
        for (int acc = grp; acc < n_acc; acc += 2) {
          const int img = acc ? it.img1 : it.img0;
          // 64 output channels (128 B of bf16) per step.  A thread owns a ROW of the accumulator, but rows are Cout*2 bytes
          // apart in the output: storing from the row owner is 32 scattered 16-byte writes per instruction (ncu: the epilogue
          // warps sat on the store queue and the tensor pipe idled 48 % of em_conv1).  So each warp transposes its 32 rows x
          // 128 B through a 4 KB shared-memory tile (16-byte chunks XOR-swizzled by row: conflict-free both ways) and stores
          // with 8 lanes per row: four full 128-byte lines per instruction.
          uint4* tile = reinterpret_cast<uint4*>(s_w2t) + (warp - 4) * 256;  // [32 rows][8 chunks of 16 B]
          // the 8 rows this lane stores (4k + lane/8) and its 16-byte chunk are the same for every column step of the item:
          // their element offsets are computed once (BW is a power of two: shifts, no divisions in the store loop)
          const int ch = lane & 7;
          int roff[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int rr = ew * 32 + 4 * k + (lane >> 3);
            const int xx = it.x0 + (rr & (P.BW - 1)), yy = it.y0 + (rr >> P.bw_shift);
            roff[k] = (xx < P.W && yy < P.H) ? (yy * P.W + xx) * P.Cout + ch * 8 : -1;
          }
          __nv_bfloat16* obase = P.out + (size_t)img * P.H * P.W * P.Cout + (size_t)it.chunk * BN;
#pragma unroll 1
          for (int c0 = 0; c0 < BN; c0 += 64) {
            float v[64];
            tmem_ld32(taddr + acc * BN + c0, v);
            tmem_ld32(taddr + acc * BN + c0 + 32, v + 32);
            tmem_ld_wait_for(v);
            tmem_ld_wait_for(v + 32);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              uint32_t pk[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int c = q * 8 + 2 * j;
                float a = v[c] + s_bias[c0 + c], b = v[c + 1] + s_bias[c0 + c + 1];
                if (EPI == EPI_STORE_RELU) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
                __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
                pk[j] = *reinterpret_cast<uint32_t*>(&h);
              }
              tile[lane * 8 + (q ^ (lane & 7))] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int r = 4 * k + (lane >> 3);
              const uint4 val = tile[r * 8 + (ch ^ (r & 7))];
              if (roff[k] >= 0) *reinterpret_cast<uint4*>(obase + roff[k] + c0) = val;
            }
            __syncwarp();
          }
        }
      } else if (EPI == EPI_COSINE) {
        // acc0 = image b of Concat_0(conv_feat, warp_feat) = current-frame embedding, acc1 = warped one (SYM:133-138)
        float scc = 0.f, sww = 0.f, swc = 0.f;
        float bc[2][32], bw[2][32];
        const int cbase = grp * (BN / 2);          // this group's half of the columns
        tmem_ld32(taddr + cbase, bc[0]);
        tmem_ld32(taddr + BN + cbase, bw[0]);
#pragma unroll 2
        for (int i = 0; i < BN / 64; ++i) {
          float* ec = bc[i & 1];
          float* ewp = bw[i & 1];
          tmem_ld_wait_for(ec);
          tmem_ld_wait_for(ewp);
          if (i + 1 < BN / 64) {
            tmem_ld32(taddr + cbase + (i + 1) * 32, bc[(i + 1) & 1]);
            tmem_ld32(taddr + BN + cbase + (i + 1) * 32, bw[(i + 1) & 1]);
          }
          const int c0 = cbase + i * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float b = s_bias[c0 + j];
            const float c = ec[j] + b, w = ewp[j] + b;
            scc = fmaf(c, c, scc);
            sww = fmaf(w, w, sww);
            swc = fmaf(w, c, swc);
          }
        }
        if (valid) {
          const size_t NP = (size_t)(P.NB / 2) * P.H * P.W;
          float* dst = P.partial + (size_t)(it.chunk * 2 + grp) * 3 * NP + (size_t)it.pair * P.H * P.W + pix;
          dst[0] = scc;
          dst[NP] = sww;
          dst[2 * NP] = swc;
        }
      } else {   // EPI_NQ
        const int n_acc = it.single ? 1 : 2;
#pragma unroll 1
        for (int acc = grp; acc < n_acc; acc += 2) {
          float q[NQ_MID];
#pragma unroll
          for (int j = 0; j < NQ_MID; ++j) q[j] = s_nq[j];                 // Nq_conv2 bias
          float buf[2][32];
          tmem_ld32(taddr + acc * BN, buf[0]);
#pragma unroll 2
          for (int i = 0; i < BN / 32; ++i) {
            float* v = buf[i & 1];
            tmem_ld_wait_for(v);
            if (i + 1 < BN / 32) tmem_ld32(taddr + acc * BN + (i + 1) * 32, buf[(i + 1) & 1]);
            const int c0 = i * 32;
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const float a = fmaxf(v[c] + s_bias[c0 + c], 0.f);           // Nq_conv1 bias + ReLU (SYM:97-98)
              const float4* wr = reinterpret_cast<const float4*>(s_w2t + (c0 + c) * NQ_MID);
#pragma unroll
              for (int j4 = 0; j4 < NQ_MID / 4; ++j4) {
                const float4 w = wr[j4];
                q[4 * j4 + 0] = fmaf(w.x, a, q[4 * j4 + 0]);
                q[4 * j4 + 1] = fmaf(w.y, a, q[4 * j4 + 1]);
                q[4 * j4 + 2] = fmaf(w.z, a, q[4 * j4 + 2]);
                q[4 * j4 + 3] = fmaf(w.w, a, q[4 * j4 + 3]);
              }
            }
          }
          float o = s_nq[2 * NQ_MID];                                      // Nq_conv3 bias
#pragma unroll
          for (int j = 0; j < NQ_MID; ++j) o = fmaf(s_nq[NQ_MID + j], fmaxf(q[j], 0.f), o);   // ReLU (SYM:100), Nq_conv3
          // image b of Concat_0(warp, conv) is the warped feature -> logits[:,0]; image b + N the current one -> [:,1]
          if (valid) P.logits[((size_t)it.pair * 2 + (it.single ? it.which : acc)) * P.H * P.W + pix] = o;
        }
      }
      tcgen05_fence_before();
      mbar_arrive(tmem_empty);
      tphase ^= 1;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// =========================================================================================================
// 2-CTA form (cta_group::2): a CTA PAIR (thread-block cluster of 2, one TPC) works on one pair item.
//   CTA r holds the activation tile of image (r ? img1 : img0) and HALF of the weight tile (128 of the 256 filters);
//   one tcgen05.mma.cta_group::2 (M = 256, N = 256), issued by the leader CTA, multiplies both: each SM reads 4 KB of A
//   and 4 KB of B per 128 cycles from its own shared memory (64 B/cycle instead of 96) and TMA writes 32 KB per k-step
//   per SM (instead of 64 KB), so operand traffic fits the 128 B/cycle a shared memory delivers; and each SM's
//   accumulator is ONE 128 x 256 tile, so two fit in tensor memory: the epilogue of item i overlaps the MMAs of item
//   i + 1 (the single-CTA kernel above exposes it: em_conv1 spends 10 k of its 27 k cycles per item there).
//   Barriers: full[s] lives in the leader (its producer arms it for the bytes of BOTH CTAs; the peer's copies signal it
//   through cp.async.bulk.tensor...cta_group::2), empty[s] / tmem_full[b] are armed in both CTAs by multicast
//   tcgen05.commit, tmem_empty[b] lives in the leader and collects one arrival per epilogue warp of both CTAs.
//   Epilogue group g (warps 4-7 | 8-11) drains accumulator buffer g, i.e. every other item.
// Serves the STORE and NQ epilogues (COSINE needs both images of a pixel in one thread: single-CTA kernel).
// =========================================================================================================
constexpr int STAGES2 = 5;
constexpr int B2_BYTES = (BN / 2) * BK * 2;           // 16 KB: this CTA's half of the weight tile
constexpr int STAGE2_BYTES = A_BYTES + B2_BYTES;      // 32 KB
constexpr int EPI2_SMEM_FLOATS = 2 * BN + EPI_TILE_FLOATS + 3 * NQ_MID + 16;
constexpr int SMEM2_BYTES = 1024 + STAGES2 * STAGE2_BYTES + EPI2_SMEM_FLOATS * 4 + 256;
static_assert(SMEM2_BYTES <= 227 * 1024, "shared memory budget (2-CTA form)");

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// tensor copies whose completion is signalled on a barrier of the LEADER CTA (bar_cluster_addr: shared::cluster address)
__device__ __forceinline__ void tma2_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar_cluster_addr) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint32_t bar_cluster_addr) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once the MMAs issued so far have completed) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma2_commit_both(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
conv_gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB2, const ConvParams P) {
  static_assert(EPI == EPI_STORE_RELU || EPI == EPI_STORE || EPI == EPI_NQ, "the 2-CTA form serves the STORE and NQ epilogues");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* s_bias = reinterpret_cast<float*>(smem + STAGES2 * STAGE2_BYTES);    // [2 groups][BN]
  float* s_w2t = s_bias + 2 * BN;                                             // STORE: transpose tiles; NQ: [BN][16]
  float* s_nq = s_w2t + EPI_TILE_FLOATS;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES2 * STAGE2_BYTES + EPI2_SMEM_FLOATS * 4);
  uint64_t* full = bars;                      // [STAGES2] (leader's are used)
  uint64_t* empty = bars + STAGES2;           // [STAGES2] in each CTA
  uint64_t* tmem_full = bars + 2 * STAGES2;   // [2] in each CTA
  uint64_t* tmem_empty = tmem_full + 2;       // [2] (leader's are used)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int num_items = P.num_items;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB2);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES2; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], 2 * 4);       // one arrival per epilogue warp of the group, both CTAs
    }
    fence_barrier_init();
  }
  if (EPI == EPI_NQ && warp >= 4) {
    const int t = threadIdx.x - 128;
    for (int i = t; i < NQ_MID * BN; i += EPI_THREADS) {
      const int j = i / BN, c = i % BN;
      s_w2t[c * NQ_MID + j] = P.nq_w2[i];
    }
    if (t < NQ_MID) {
      s_nq[t] = P.nq_b2[t];
      s_nq[NQ_MID + t] = P.nq_w3[t];
    }
    if (t == 0) s_nq[2 * NQ_MID] = P.nq_b3[0];
  }
  __syncthreads();
  cluster_sync_all();                          // both CTAs' barriers exist before anything remote can arrive on them
  if (warp == 2) tmem_alloc2(tmem_slot, TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs: own activation tile + own half of the weight tile) =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int item = cluster_id; item < num_items; item += num_clusters) {
        const Item it = decode_item(P, item);
        const int img = rank ? it.img1 : it.img0;
        for (int ks = 0; ks < P.k_steps; ++ks) {
          const int tap = ks / P.kc_per_tap, kc = ks - tap * P.kc_per_tap;
          const int dy = P.taps == 9 ? tap / 3 - 1 : 0, dx = P.taps == 9 ? tap % 3 - 1 : 0;
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* st = smem + stage * STAGE2_BYTES;
          if (rank == 0) mbar_expect_tx(&full[stage], 2 * STAGE2_BYTES);
          const uint32_t fb = mapa_shared(smem_u32(&full[stage]), 0);
          tma2_load_4d(st, &tmA, kc * BK, it.x0 + dx, it.y0 + dy, img, fb);
          tma2_load_2d(st + A_BYTES, &tmB2, tap * P.Cin + kc * BK, it.chunk * BN + (int)rank * (BN / 2), fb);
          if (++stage == STAGES2) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN);
      uint32_t stage = 0, phase = 0;
      int n = 0;
      for (int item = cluster_id; item < num_items; item += num_clusters, ++n) {
        const int b = n & 1;
        mbar_wait(&tmem_empty[b], (((unsigned)n >> 1) & 1u) ^ 1u);   // both CTAs' epilogues have drained this buffer
        tcgen05_fence_after();
        for (int ks = 0; ks < P.k_steps; ++ks) {
          mbar_wait(&full[stage], phase);
          tcgen05_fence_after();
          const uint32_t a0 = smem_u32(smem + stage * STAGE2_BYTES);
          const uint64_t da = make_kmajor_sw128_desc(a0);
          const uint64_t db = make_kmajor_sw128_desc(a0 + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) umma2_bf16(tmem_base + b * BN, da + 2 * k, db + 2 * k, idesc, (ks | k) != 0);
          umma2_commit_both(&empty[stage]);
          if (ks == P.k_steps - 1) umma2_commit_both(&tmem_full[b]);
          if (++stage == STAGES2) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: group g drains accumulator buffer g (every other item) =====================
    const int ew = warp & 3;
    const int grp = (warp - 4) >> 2;
    const int row = ew * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)grp * BN;
    float* bias_g = s_bias + grp * BN;
    const uint32_t te_addr = mapa_shared(smem_u32(&tmem_empty[grp]), 0);
    unsigned use = 0;                            // uses of this group's buffer so far
    int n = 0;
    for (int item = cluster_id; item < num_items; item += num_clusters, ++n) {
      if ((n & 1) != grp) continue;
      const Item it = decode_item(P, item);
      const int img = rank ? it.img1 : it.img0;
      asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");      // previous item's readers of bias_g are done
      for (int i = (threadIdx.x - 128) & 127; i < BN; i += 128) bias_g[i] = P.bias[it.chunk * BN + i];
      asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
      const int x = it.x0 + (row & (P.BW - 1)), y = it.y0 + (row >> P.bw_shift);
      const bool valid = x < P.W && y < P.H;
      const size_t pix = (size_t)y * P.W + x;
      mbar_wait(&tmem_full[grp], use & 1u);
      tcgen05_fence_after();
      if (EPI == EPI_STORE_RELU || EPI == EPI_STORE) {
        uint4* tile = reinterpret_cast<uint4*>(s_w2t) + (warp - 4) * 256;
        const int ch = lane & 7;
        int roff[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int rr = ew * 32 + 4 * k + (lane >> 3);
          const int xx = it.x0 + (rr & (P.BW - 1)), yy = it.y0 + (rr >> P.bw_shift);
          roff[k] = (xx < P.W && yy < P.H) ? (yy * P.W + xx) * P.Cout + ch * 8 : -1;
        }
        __nv_bfloat16* obase = P.out + (size_t)img * P.H * P.W * P.Cout + (size_t)it.chunk * BN;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 64) {
          float v[64];
          tmem_ld32(taddr + c0, v);
          tmem_ld32(taddr + c0 + 32, v + 32);
          tmem_ld_wait_for(v);
          tmem_ld_wait_for(v + 32);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            uint32_t pk[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int c = q * 8 + 2 * j;
              float a = v[c] + bias_g[c0 + c], b = v[c + 1] + bias_g[c0 + c + 1];
              if (EPI == EPI_STORE_RELU) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
              __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
              pk[j] = *reinterpret_cast<uint32_t*>(&h);
            }
            tile[lane * 8 + (q ^ (lane & 7))] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
          __syncwarp();
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int r = 4 * k + (lane >> 3);
            const uint4 val = tile[r * 8 + (ch ^ (r & 7))];
            if (roff[k] >= 0) *reinterpret_cast<uint4*>(obase + roff[k] + c0) = val;
          }
          __syncwarp();
        }
      } else {   // EPI_NQ
        float q[NQ_MID];
#pragma unroll
        for (int j = 0; j < NQ_MID; ++j) q[j] = s_nq[j];
        float buf[2][32];
        tmem_ld32(taddr, buf[0]);
#pragma unroll 2
        for (int i = 0; i < BN / 32; ++i) {
          float* v = buf[i & 1];
          tmem_ld_wait_for(v);
          if (i + 1 < BN / 32) tmem_ld32(taddr + (i + 1) * 32, buf[(i + 1) & 1]);
          const int c0 = i * 32;
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const float a = fmaxf(v[c] + bias_g[c0 + c], 0.f);
            const float4* wr = reinterpret_cast<const float4*>(s_w2t + (c0 + c) * NQ_MID);
#pragma unroll
            for (int j4 = 0; j4 < NQ_MID / 4; ++j4) {
              const float4 w = wr[j4];
              q[4 * j4 + 0] = fmaf(w.x, a, q[4 * j4 + 0]);
              q[4 * j4 + 1] = fmaf(w.y, a, q[4 * j4 + 1]);
              q[4 * j4 + 2] = fmaf(w.z, a, q[4 * j4 + 2]);
              q[4 * j4 + 3] = fmaf(w.w, a, q[4 * j4 + 3]);
            }
          }
        }
        float o = s_nq[2 * NQ_MID];
#pragma unroll
        for (int j = 0; j < NQ_MID; ++j) o = fmaf(s_nq[NQ_MID + j], fmaxf(q[j], 0.f), o);
        if (valid) P.logits[((size_t)it.pair * 2 + rank) * P.H * P.W + pix] = o;
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(te_addr);
      ++use;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                          // the peer's tensor memory reads and remote arrivals are done
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc2(tmem_base, TMEM_COLS);
  }
}

// cosine logits from the per-chunk partial sums (fixed summation order: deterministic).  compute_weight SYM:111-116 with
// MXNet's L2Normalization(mode='channel'): e / sqrt(sum e^2 + 1e-10);  logits[n,0] = <e_warp^, e_cur^>, [n,1] = <e_cur^, e_cur^>
__global__ void cosine_from_partials_kernel(const float* __restrict__ partial, float* __restrict__ logits, int n_chunks, int N,
                                            int HW) {
  const size_t NP = (size_t)N * HW;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NP) return;
  float scc = 0.f, sww = 0.f, swc = 0.f;
  for (int c = 0; c < n_chunks; ++c) {
    const float* p = partial + (size_t)c * 3 * NP + i;
    scc += p[0];
    sww += p[NP];
    swc += p[2 * NP];
  }
  const float nc = sqrtf(scc + 1e-10f), nw = sqrtf(sww + 1e-10f);
  const size_t n = i / HW, p = i % HW;
  logits[(n * 2 + 0) * HW + p] = swc / (nw * nc);
  logits[(n * 2 + 1) * HW + p] = scc / (nc * nc);
}

// (Cout,Cin,kh,kw) fp32 (MXNet's layout) -> (Cout, kh*kw*Cin) bf16: K index = tap * Cin + cin, tap = ky*kw + kx
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int Cout, int Cin, int kk) {
  const size_t total = (size_t)Cout * Cin * kk;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int cin = (int)(i % Cin);
    const size_t r = i / Cin;
    const int tap = (int)(r % kk);
    const size_t co = r / kk;
    out[i] = __float2bfloat16_rn(w[(co * Cin + cin) * kk + tap]);
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }();
  return fn;
}

void pick_tile(int H, int W, int* BW, int* BH) {
  long best = -1;
  for (int bw = 128; bw >= 8; bw >>= 1) {
    const int bh = BM / bw;
    const long cost = (long)((W + bw - 1) / bw) * bw * ((H + bh - 1) / bh) * bh;
    if (best < 0 || cost < best) { best = cost; *BW = bw; *BH = bh; }
  }
}

static const char* make_maps(const void* x, const void* w, const ConvParams& P, CUtensorMap* tmA, CUtensorMap* tmB) {
  auto enc = encode_fn();
  if (!enc) return "cuTensorMapEncodeTiled not available from the driver";
  {
    cuuint64_t dims[4] = {(cuuint64_t)P.Cin, (cuuint64_t)P.W, (cuuint64_t)P.H, (cuuint64_t)P.NB};
    cuuint64_t strides[3] = {(cuuint64_t)P.Cin * 2, (cuuint64_t)P.W * P.Cin * 2, (cuuint64_t)P.H * P.W * P.Cin * 2};
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)P.BW, (cuuint32_t)P.BH, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return "cuTensorMapEncodeTiled failed for the activation tensor";
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)P.taps * P.Cin, (cuuint64_t)P.Cout};
    cuuint64_t strides[1] = {(cuuint64_t)P.taps * P.Cin * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BN};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return "cuTensorMapEncodeTiled failed for the weight tensor";
  }
  return nullptr;
}

// 2-CTA form: grid = 2 x (clusters that can be co-resident), cluster dims are a compile-time attribute of the kernel
template <int EPI>
static const char* launch_epi2(const CUtensorMap& tmA, const CUtensorMap& tmB2, const ConvParams& P, cudaStream_t stream, bool* launched) {
  // per device: the shared-memory opt-in is a per-device attribute and the cluster occupancy a per-device fact
  // (benign race between host threads: both compute the same values)
  static int max_clusters_dev[64];
  static bool known[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  int& max_clusters = max_clusters_dev[dev];
  if (!known[dev]) {
    max_clusters = -1;
    known[dev] = true;
  }
  auto kfn = conv_gemm_tc2_kernel<EPI>;
  if (max_clusters < 0) {
    if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES) != cudaSuccess) {
      cudaGetLastError();
      max_clusters = 0;
    } else {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2 * 74);
      cfg.blockDim = dim3(NUM_THREADS);
      cfg.dynamicSmemBytes = SMEM2_BYTES;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kfn, &cfg) != cudaSuccess) {
        cudaGetLastError();
        n = 0;
      }
      max_clusters = n;
    }
  }
  *launched = false;
  if (max_clusters < 8) return nullptr;          // not worth it / not possible: the caller uses the single-CTA kernel
  const int clusters = P.num_items < max_clusters ? P.num_items : max_clusters;
  kfn<<<2 * clusters, NUM_THREADS, SMEM2_BYTES, stream>>>(tmA, tmB2, P);
  if (cudaPeekAtLastError() != cudaSuccess) {    // e.g. a partition that cannot co-schedule CTA pairs: single-CTA kernel instead
    cudaGetLastError();
    max_clusters = 0;
    return nullptr;
  }
  *launched = true;
  return nullptr;
}

template <int EPI>
static const char* launch_epi(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvParams& P, int sms, cudaStream_t stream) {
  static bool attr_set[64];        // per device; benign race: the attribute is idempotent
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!attr_set[dev]) {
    if (cudaFuncSetAttribute(conv_gemm_tc_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess)
      return "cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed";
    attr_set[dev] = true;
  }
  const int grid = P.num_items < sms ? P.num_items : sms;
  conv_gemm_tc_kernel<EPI><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tmA, tmB, P);
  return nullptr;
}

const char* launch_conv(const void* x, const void* w, ConvParams P, int epi, int sms, cudaStream_t stream) {
  if (P.NB <= 0) return "no images";
  if ((P.NB & 1) && (epi == EPI_COSINE || epi == EPI_NQ)) return "the cosine / Nq epilogues pair image b with image b + N (Concat_0 of two N-batches): NB must be even";
  if (P.Cin % BK) return "Cin must be a multiple of 64";
  if (P.Cout % BN) return "Cout must be a multiple of 256";
  if (P.taps != 1 && P.taps != 9) return "kernel must be 1x1 or 3x3";
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w)) & 15) return "activation / weight pointers must be 16-byte aligned";
  pick_tile(P.H, P.W, &P.BW, &P.BH);
  P.bw_shift = 0;
  while ((1 << P.bw_shift) < P.BW) ++P.bw_shift;
  if ((long long)P.H * P.W * P.Cout >= (1LL << 31)) return "one image's output must stay below 2^31 elements";
  P.tiles_x = (P.W + P.BW - 1) / P.BW;
  P.tiles_y = (P.H + P.BH - 1) / P.BH;
  P.n_chunks = P.Cout / BN;
  P.kc_per_tap = P.Cin / BK;
  P.k_steps = P.taps * P.kc_per_tap;
  P.pair_items = ((P.NB + 1) / 2) * P.tiles_x * P.tiles_y * P.n_chunks;
  P.num_items = P.pair_items;
  if (epi != EPI_COSINE && 2 * P.pair_items <= sms) {   // fewer items than half the SMs: every tile on its own CTA
    P.num_items = 2 * P.pair_items;
    P.pair_items = 0;
  } else if (epi != EPI_COSINE && P.pair_items > sms) { // cut the pair items of a partly filled last wave in two
    const int rem = P.pair_items % sms;
    if (rem > 0 && 2 * rem <= sms) {
      P.pair_items -= rem;
      P.num_items = P.pair_items + 2 * rem;
    }
  }
  CUtensorMap tmA, tmB;
  if (const char* e = make_maps(x, w, P, &tmA, &tmB)) return e;
  // CTA-pair form (cta_group::2) for the STORE / NQ epilogues on even image counts: every pair item on a cluster
  if (epi != EPI_COSINE && !(P.NB & 1) && !(epi == EPI_NQ && P.Cout != BN) && knob("LSFA_TC_NO_PAIR") == nullptr) {
    ConvParams P2 = P;
    P2.pair_items = (P.NB / 2) * P.tiles_x * P.tiles_y * P.n_chunks;
    P2.num_items = P2.pair_items;
    CUtensorMap tmB2;
    auto enc = encode_fn();
    cuuint64_t dims[2] = {(cuuint64_t)P.taps * P.Cin, (cuuint64_t)P.Cout};
    cuuint64_t strides[1] = {(cuuint64_t)P.taps * P.Cin * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)(BN / 2)};
    cuuint32_t es[2] = {1, 1};
    if (enc(&tmB2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS) {
      bool launched = false;
      const char* e = nullptr;
      if (epi == EPI_STORE_RELU) e = launch_epi2<EPI_STORE_RELU>(tmA, tmB2, P2, stream, &launched);
      else if (epi == EPI_STORE) e = launch_epi2<EPI_STORE>(tmA, tmB2, P2, stream, &launched);
      else e = launch_epi2<EPI_NQ>(tmA, tmB2, P2, stream, &launched);
      if (e) return e;
      if (launched) return nullptr;
    }
  }
  switch (epi) {
    case EPI_STORE_RELU: return launch_epi<EPI_STORE_RELU>(tmA, tmB, P, sms, stream);
    case EPI_STORE: return launch_epi<EPI_STORE>(tmA, tmB, P, sms, stream);
    case EPI_COSINE: return launch_epi<EPI_COSINE>(tmA, tmB, P, sms, stream);
    case EPI_NQ:
      if (P.Cout != BN) return "Nq_conv1 must have exactly 256 filters (SYM:97)";
      return launch_epi<EPI_NQ>(tmA, tmB, P, sms, stream);
  }
  return "unknown epilogue";
}

void launch_cosine_finalize(const float* partial, float* logits, int n_chunks, int N, int HW, cudaStream_t stream) {
  const size_t NP = (size_t)N * HW;
  cosine_from_partials_kernel<<<(unsigned)((NP + 255) / 256), 256, 0, stream>>>(partial, logits, n_chunks, N, HW);
}

void launch_pack_weight(const float* w, void* out, int Cout, int Cin, int kk, cudaStream_t stream) {
  const size_t total = (size_t)Cout * Cin * kk;
  const unsigned grid = (unsigned)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  pack_conv_weight_kernel<<<grid, 256, 0, stream>>>(w, reinterpret_cast<__nv_bfloat16*>(out), Cout, Cin, kk);
}

}  // namespace tc
}  // namespace lsfa
