// agg_nchw_plane_kernel - "plane-resident gather": the headline kernel of the NCHW fp32 path.
//
// Replaces, in ONE pass over HBM, the operator chain of
//   SYM:571-572 GridGenerator(warp) + BilinearSampler      (a7, a8)
//   SYM:308/470/680  * scale_map                            (a9)
//   SYM:66,576  + rnet_conv0(res_diff)                      (a10)
//   SYM:236 / 104-108 / 144-147 / 315  aggregation          (a11-a14)
//   operator_py/choose_feat.py:23-31  per-frame select      (a15)
// and, when raw motion vectors are given, lib/utils/image.py:207-228 (a3,a5,a6).
//
// NCHW planes are small (38x63 fp32 = 9.6 KB), so K whole key planes are streamed into shared
// memory with one TMA bulk copy (cp.async.bulk -> mbarrier, SASS UBLKCP) per stage of a ring
// and the 4-tap gather runs against shared memory: HBM sees every key byte exactly once
// whatever the motion magnitude.  scale_map / cur / out are pure coalesced streams.
// The per-pixel sampling record (4 weights, 4 packed 16-bit tap offsets, blend weight) lives in
// REGISTERS: each thread owns the same <= 8 pixels for every channel of a frame, so the MV
// pooling, the exact fp32 grid round trip and the softmax run once per frame per CTA.
// Persistent grid: one 512-thread CTA per SM walks a contiguous range of
// (frame, channel-chunk, pixel-part) items.
//
// The per-element work is specialised at compile time (VAR) so the inner loop carries no
// mode branches: ~22 instructions per element (4 LDS, 2 LDG, 1 STG, 6 FP, address adds).
#pragma once
#include "lsfa_device.cuh"

namespace lsfa {

constexpr int kPlaneThreads = 512;
constexpr int kPlaneWarps = kPlaneThreads / 32;
constexpr int kMaxStages = 8;
constexpr int kBarrierBytes = 128;  // 2 * kMaxStages * 8

// compile-time variants of the element arithmetic
enum PlaneVariant : int {
  kVarRuntime = 0,    // every flag read from AggParams (rare combinations, req = add)
  kVarWarpOnly = 1,   // out = warp                       (BilinearSampler alone, SYM:572)
  kVarScale = 2,      // out = warp * scale               (batch path, SYM:678-680)
  kVarScaleCur = 3,   // out = wc*cur + ww*warp*scale     (key frame Nq/Fgfa/mean, SYM:468-476)
  kVarResCur = 4,     // out = cur + warp + rnet(res)     (non-key frame as shipped, SYM:571-586)
  kNumPlaneVariants = 5
};

template <int K, int PPT, int VAR>
__global__ void __launch_bounds__(kPlaneThreads, 1)
agg_nchw_plane_kernel(const __grid_constant__ AggParams P) {
  constexpr bool RT = VAR == kVarRuntime;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* empty = full + kMaxStages;
  float* ring = reinterpret_cast<float*>(smem_raw + kBarrierBytes);
  const unsigned stage_floats = P.stage_bytes / 4;
  float* res_s = ring + (size_t)P.stages * stage_floats;  // [3][part_pix], thread-private slots

  const int tid = threadIdx.x;
  const long long i0 = P.items * (long long)blockIdx.x / gridDim.x;
  const long long i1 = P.items * (long long)(blockIdx.x + 1) / gridDim.x;
  const long long items_per_frame = (long long)P.parts * P.chunks;
  const float* __restrict__ scale = static_cast<const float*>(P.scale);
  const float* __restrict__ cur = static_cast<const float*>(P.cur);
  float* __restrict__ out = static_cast<float*>(P.out);
  const bool has_scale = RT ? (scale != nullptr) : (VAR == kVarScale || VAR == kVarScaleCur);
  const bool has_cur = RT ? (P.mode != LSFA_W_NONE) : (VAR == kVarScaleCur || VAR == kVarResCur);
  const bool has_res = RT ? (P.res != nullptr) : (VAR == kVarResCur);
  const bool req_add = RT ? (P.req_add != 0) : false;
  const bool has_bypass = P.bypass != nullptr;

  if (tid == 0) {
    for (int s = 0; s < P.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kPlaneWarps);
    }
    fence_barrier_init();
  }
  __syncthreads();

  // Item = (frame n, channel chunk, pixel part), walked incrementally: no divisions in the loop.
  struct Cursor {
    int n, chunk, part;
    __device__ __forceinline__ void advance(int parts, int chunks) {
      if (++part == parts) {
        part = 0;
        if (++chunk == chunks) {
          chunk = 0;
          ++n;
        }
      }
    }
  };
  auto cursor_at = [&](long long it) {
    Cursor c;
    c.part = (int)(it % P.parts);
    c.chunk = (int)((it / P.parts) % P.chunks);
    c.n = (int)(it / items_per_frame);
    return c;
  };

  // ---- producer state (meaningful in thread 0 only) ----
  Cursor pc = cursor_at(i0);   // next item whose key planes have not been requested yet
  long long pit = i0;
  int ps = 0, pround = 0;      // ring slot / how many times the ring has wrapped
  auto try_issue_one = [&]() {
    while (pit < i1 && has_bypass && __ldg(P.bypass + pc.n)) {  // bypass frames never touch the key feature
      pit += items_per_frame - ((long long)pc.chunk * P.parts + pc.part);
      pc.part = 0;
      pc.chunk = 0;
      ++pc.n;
    }
    if (pit >= i1) return;
    const int kn = key_slot(P, pc.n);
    const float* src = static_cast<const float*>(P.key) + ((size_t)kn * P.C + (size_t)pc.chunk * K) * P.HWk;
    if (pround >= 1) mbar_wait(&empty[ps], (unsigned)(pround - 1) & 1u);
    mbar_expect_tx(&full[ps], P.stage_bytes);
    bulk_g2s(ring + (size_t)ps * stage_floats, src, P.stage_bytes, &full[ps]);
    if (++ps == P.stages) {
      ps = 0;
      ++pround;
    }
    ++pit;
    pc.advance(P.parts, P.chunks);
  };
  if (tid == 0)
    for (int s = 0; s < P.stages - 1; ++s) try_issue_one();

  // ---- consumer state: the register-resident sampling records of this thread's pixels ----
  float w00[PPT], w01[PPT], w10[PPT], w11[PPT], wc[PPT], ww[PPT];
  unsigned o_top[PPT], o_bot[PPT];   // byte offsets inside a plane: (i00 | i01 << 16), (i10 | i11 << 16)
  unsigned valid = 0;                // bit j: this thread's pixel slot j is inside the part
  int cur_n = -1, cur_part = -1;
  int cs = 0;                        // ring slot of the next use
  unsigned cphase = 0;               // its phase parity
  const unsigned plane_bytes = (unsigned)P.HWk * 4u;
  const size_t chan_stride = (size_t)P.HW;
  Cursor cc = cursor_at(i0);

  for (long long it = i0; it < i1; ++it, cc.advance(P.parts, P.chunks)) {
    const int n = cc.n, part = cc.part;
    const int pix0 = part * P.part_pix;
    const bool byp = has_bypass && (__ldg(P.bypass + n) != 0);

    if (n != cur_n || part != cur_part) {  // new frame (or pixel part): rebuild the records
      cur_n = n;
      cur_part = part;
      const int pend = min(P.HW, pix0 + P.part_pix);
      valid = 0;
      // the old records are dead from here on: clearing them first frees their registers for the
      // load batch below (otherwise ptxas spills the batch and every spill store waits on its load)
#pragma unroll
      for (int j = 0; j < PPT; ++j) {
        w00[j] = w01[j] = w10[j] = w11[j] = wc[j] = ww[j] = 0.0f;
        o_top[j] = o_bot[j] = 0u;      // slot outside the part: taps read offset 0, store is predicated off
      }
      // phase A0: L2 prefetch of everything the records need, all pixel slots back to back
      if (!byp) {
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
          const int p = pix0 + tid + j * kPlaneThreads;
          if (p < pend) prefetch_pixel_loads(P, n, p / P.W, p % P.W);
        }
      }
      // phase A: the loads proper, a few pixel slots at a time (they now hit in L2)
      constexpr int RG = PPT >= 3 ? 3 : PPT;              // slots per load batch (register budget)
#pragma unroll
      for (int j0 = 0; j0 < PPT; j0 += RG) {
        PixelLoads ld[RG];
#pragma unroll
        for (int g = 0; g < RG; ++g) {
          const int j = j0 + g;
          const int p = pix0 + tid + j * kPlaneThreads;
          if (j < PPT && p < pend && !byp) ld[g] = issue_pixel_loads(P, n, p / P.W, p % P.W);
        }
        // phase B: the arithmetic (float64 pooling, exact fp32 grid round trip, softmax, fold)
#pragma unroll
        for (int g = 0; g < RG; ++g) {
          const int j = j0 + g;
          if (j >= PPT) continue;
          const int p = pix0 + tid + j * kPlaneThreads;
          if (p < pend) {
            valid |= 1u << j;
            if (!byp) {
              const PixelRec t = finish_pixel(P, ld[g], n, p / P.W, p % P.W);
              w00[j] = t.w00; w01[j] = t.w01; w10[j] = t.w10; w11[j] = t.w11;
              wc[j] = t.wc; ww[j] = t.ww;
              o_top[j] = (unsigned)(t.i00 * 4) | ((unsigned)(t.i01 * 4) << 16);
              o_bot[j] = (unsigned)(t.i10 * 4) | ((unsigned)(t.i11 * 4) << 16);
              if (has_res) {
#pragma unroll
                for (int k = 0; k < 3; ++k)
                  res_s[k * P.part_pix + (p - pix0)] = __ldg(P.res + ((size_t)n * 3 + k) * P.HW + p);
              }
            }
          }
        }
      }
    }

    const int c0 = cc.chunk * K;
    const size_t e0 = ((size_t)n * P.C + c0) * chan_stride + pix0 + tid;   // element of (k=0, j=0)

    // once-touched streams first: they are in flight while we wait for the key planes.
    // Predicated (not branched) so the unrolled loop stays straight-line code.
    float sc[K][PPT], cu[K][PPT];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float* sp = scale + e0 + (size_t)k * chan_stride;
      const float* cp = cur + e0 + (size_t)k * chan_stride;
#pragma unroll
      for (int j = 0; j < PPT; ++j) {
        const unsigned ok = (valid >> j) & 1u;
        sc[k][j] = (has_scale && !byp) ? ldg_stream_if(sp + j * kPlaneThreads, ok) : 1.0f;
        cu[k][j] = has_cur ? ldg_stream_if(cp + j * kPlaneThreads, ok) : 0.0f;
      }
    }

    if (byp) {  // ChooseFeat: keep the current frame's own feature
#pragma unroll
      for (int k = 0; k < K; ++k) {
        float* op = out + e0 + (size_t)k * chan_stride;
#pragma unroll
        for (int j = 0; j < PPT; ++j)
          if ((valid >> j) & 1u) __stcs(op + j * kPlaneThreads, req_add ? cu[k][j] + op[j * kPlaneThreads] : cu[k][j]);
      }
      continue;
    }

    if (tid == 0) try_issue_one();  // refills the stage consumed one item ago

    mbar_wait(&full[cs], cphase);
    const unsigned char* stage_s = smem_raw + kBarrierBytes + (size_t)cs * P.stage_bytes;

#pragma unroll
    for (int k = 0; k < K; ++k) {
      float rw0 = 0.f, rw1 = 0.f, rw2 = 0.f, rb = 0.f;
      if (has_res) {
        rw0 = __ldg(P.rnet_w + (size_t)(c0 + k) * 3 + 0);
        rw1 = __ldg(P.rnet_w + (size_t)(c0 + k) * 3 + 1);
        rw2 = __ldg(P.rnet_w + (size_t)(c0 + k) * 3 + 2);
        rb = __ldg(P.rnet_b + c0 + k);
      }
      const unsigned char* plane_s = stage_s + (size_t)k * plane_bytes;
      float* op = out + e0 + (size_t)k * chan_stride;
#pragma unroll
      for (int j = 0; j < PPT; ++j) {
        const float v00 = *reinterpret_cast<const float*>(plane_s + (o_top[j] & 0xffffu));
        const float v01 = *reinterpret_cast<const float*>(plane_s + (o_top[j] >> 16));
        const float v10 = *reinterpret_cast<const float*>(plane_s + (o_bot[j] & 0xffffu));
        const float v11 = *reinterpret_cast<const float*>(plane_s + (o_bot[j] >> 16));
        float v = w00[j] * v00;
        v = fmaf(w01[j], v01, v);
        v = fmaf(w10[j], v10, v);
        v = fmaf(w11[j], v11, v);
        if (has_scale) v *= sc[k][j];
        if (has_res) {
          const int q = tid + j * kPlaneThreads;   // slots outside the part read slot-0-initialised garbage: never stored
          v = fmaf(ww[j], rnet_term(rw0, rw1, rw2, rb, res_s[q], res_s[P.part_pix + q], res_s[2 * P.part_pix + q]), v);
        }
        float o = has_cur ? fmaf(wc[j], cu[k][j], v) : v;
        const unsigned ok = (valid >> j) & 1u;
        if (req_add) {
          if (ok) __stcs(op + j * kPlaneThreads, o + op[j * kPlaneThreads]);   // kAddTo: the rare path
        } else {
          stg_stream_if(op + j * kPlaneThreads, o, ok);
        }
      }
    }
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&empty[cs]);
    if (++cs == P.stages) {
      cs = 0;
      cphase ^= 1u;
    }
  }
}

// One translation unit per variant instantiates every (K, PPT) it serves; see plane_var*.cu.
template <int VAR>
cudaError_t launch_plane_variant(const AggParams& P, size_t smem, int grid, cudaStream_t st);

#define LSFA_PLANE_FOREACH_KP(X) X(2, 1) X(2, 3) X(2, 5) X(2, 8) X(4, 1) X(4, 3) X(4, 5)

#define LSFA_PLANE_LAUNCH(VAR, KK, PP)                                                            \
  if (P.K == KK && ppt == PP) {                                                                   \
    auto kfn = agg_nchw_plane_kernel<KK, PP, VAR>;                                                \
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return e;                                                               \
    kfn<<<grid, kPlaneThreads, smem, st>>>(P);                                                    \
    return cudaPeekAtLastError();                                                                 \
  }

}  // namespace lsfa
