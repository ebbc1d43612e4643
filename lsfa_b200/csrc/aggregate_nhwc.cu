// Fused warp x scale + aggregation, NHWC (channels-last) float32 / bfloat16.
//
// Same operator chain as aggregate_nchw.cu (SYM:571-576, 308/470/680, 104-108, 132-148, 236;
// operator_py/choose_feat.py:23-31), laid out for channel-vectorised access: a pixel's C
// channels are contiguous (C=1024 bf16 = 2 KB), so every one of the 4 bilinear taps, the
// scale map, the current feature and the output are 16-byte-per-lane, fully coalesced
// vector accesses (bf16x8 / f32x4).  Persistent CTAs walk 4x8 tiles of output pixels; a
// dedicated warp builds the next tile's sampling records into shared memory while 8 warps
// stream the channels of the current tile; neighbouring taps (x and y) hit in L1, the rest of
// the key feature's reuse is served by the 126 MB L2.  fp32 arithmetic throughout.
// The cosine-embedding weights (Fgfa_net) are computed in the same pass with warp-shuffle
// reductions, so this layout needs no workspace and no second kernel.
#include <cstdlib>

#include "aggregate_nchw_plane.cuh"   // PlaneVariant: the same compile-time variants as the NCHW kernels
#include "aggregate_nhwc_tma.cuh"     // the all-TMA form (default where it applies)

namespace lsfa {

constexpr int kNhwcThreads = 256;
constexpr int kNhwcWarps = kNhwcThreads / 32;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Tile geometry: a persistent CTA walks kTileH x kTileW tiles of output pixels.  Warp w (0..7)
// owns column w of the tile and streams its rows two at a time (12 independent 16-byte loads in
// flight per lane); a ninth warp builds the sampling records of the NEXT tile meanwhile, so the
// long dependent chain of the index math (MV taps -> float64 pooling -> fp32 round trip ->
// softmax) never sits on the streaming warps' critical path.  Channels are walked in chunks of
// kChunkVecs 16-byte vectors per pixel so a tile pass touches ~45 key pixels x 512 B: the x- and
// y-neighbour taps of the 4-tap stencil are re-read while still in L1.
#ifndef LSFA_NHWC_TILE_H
#define LSFA_NHWC_TILE_H 4
#endif
constexpr int kTileH = LSFA_NHWC_TILE_H;   // rows per tile (multiple of 4: the record warp builds 32 pixels per pass)
constexpr int kTileW = 8;
constexpr int kTilePix = kTileH * kTileW;
constexpr int kStreamWarps = kTileW;                       // one warp per tile column
constexpr int kNhwcTileThreads = (kStreamWarps + 1) * 32;   // + the record warp
#ifndef LSFA_NHWC_PREFETCH
#define LSFA_NHWC_PREFETCH 1
#endif
constexpr bool kNhwcPrefetch = LSFA_NHWC_PREFETCH != 0;

struct __align__(16) TileRec {
  float w00, w01, w10, w11;   // tap weights, blend weight folded in
  float ww, wc;
  int i00, i01, i10, i11;     // key pixel indices of the taps (in-bounds)
  int use;                    // bit 8: pixel exists, bit 9: bypass frame
  int n;                      // frame of the tile
};

template <typename T, int VAR>
__global__ void __launch_bounds__(kNhwcTileThreads, 2)
agg_nhwc_kernel(const __grid_constant__ AggParams P) {
  constexpr bool RT = VAR == kVarRuntime;
  constexpr bool HAS_RES = RT || VAR == kVarResCur;   // RT: decided at run time below
  using V = Vec16<T>;
  constexpr int L = V::kLanes;        // channels per 16-byte vector
  constexpr int CSTEP = 32 * L;       // channels per warp pass
  __shared__ TileRec recs[2][kTilePix];
  // residual conv weights (a10, SYM:66) as (w0,w1,w2,b) per channel, transposed so that the lanes of a warp read
  // consecutive float4: channel cb + lane*L + k sits at cb + k*32 + lane (cb = multiple of 32*L).  Per-element
  // global loads of the (C,3) array touched ~24 cache lines per warp instruction: 0.21 of the HBM peak measured.
  extern __shared__ float4 rnet_s[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const T* __restrict__ key = static_cast<const T*>(P.key);
  const T* __restrict__ scale = static_cast<const T*>(P.scale);
  const T* __restrict__ cur = static_cast<const T*>(P.cur);
  T* __restrict__ out = static_cast<T*>(P.out);
  const bool has_cur = RT ? (P.mode != LSFA_W_NONE) : (VAR == kVarScaleCur || VAR == kVarResCur);
  const bool has_scale = RT ? (scale != nullptr) : (VAR == kVarScale || VAR == kVarScaleCur);
  const bool has_res = RT ? (P.res != nullptr) : (VAR == kVarResCur);
  const bool req_add = RT ? (P.req_add != 0) : false;
  const int tiles_x = (P.W + kTileW - 1) / kTileW, tiles_y = (P.H + kTileH - 1) / kTileH;
  const long long tiles = (long long)P.N * tiles_x * tiles_y;

  // ---- record builder: one lane per pixel of the tile (a3,a5,a6,a7,a8 index math, a13 softmax) ----
  auto build_records = [&](long long tile, TileRec* dst) {
    const int tx = (int)(tile % tiles_x);
    const int ty = (int)((tile / tiles_x) % tiles_y);
    const int n = (int)(tile / ((long long)tiles_x * tiles_y));
   for (int pass = 0; pass < kTilePix / 32; ++pass) {   // one lane per pixel, 32 pixels per pass
    const int r = pass * (32 / kTileW) + lane / kTileW, cx = lane % kTileW;
    const int y = ty * kTileH + r, x = tx * kTileW + cx;
    TileRec rec;
    rec.w00 = rec.w01 = rec.w10 = rec.w11 = rec.ww = rec.wc = 0.f;
    rec.i00 = rec.i01 = rec.i10 = rec.i11 = 0;
    rec.use = 0;
    rec.n = n;
    if (y < P.H && x < P.W) {
      rec.use = 1 << 8;
      if (P.bypass != nullptr && __ldg(P.bypass + n) != 0) {
        rec.use |= 1 << 9;
      } else {
        const PixelLoads ld = issue_pixel_loads(P, n, y, x);
        // cosine weights come later (phase 1b); every other mode folds its blend weight here
        const PixelRec t = finish_pixel(P, ld, n, y, x, /*fold=*/P.mode != LSFA_W_COSINE);
        // taps outside the key plane keep weight 0 and a clamped in-bounds index (make_taps)
        rec.w00 = t.w00; rec.w01 = t.w01; rec.w10 = t.w10; rec.w11 = t.w11;
        rec.ww = t.ww; rec.wc = t.wc;
        rec.i00 = t.i00; rec.i01 = t.i01; rec.i10 = t.i10; rec.i11 = t.i11;
      }
    }
    dst[pass * 32 + lane] = rec;
   }
  };

  long long tile = blockIdx.x;
  if (tile >= tiles) return;
  if (HAS_RES && has_res && P.rnet_smem) {
    const int cpad = (P.C + CSTEP - 1) / CSTEP * CSTEP;
    for (int j = threadIdx.x; j < cpad; j += blockDim.x) {
      const int cb = j / CSTEP * CSTEP, k = (j - cb) / 32, ln = j & 31;
      const int ch = cb + ln * L + k;
      rnet_s[j] = ch < P.C ? make_float4(__ldg(P.rnet_w + (size_t)ch * 3), __ldg(P.rnet_w + (size_t)ch * 3 + 1),
                                         __ldg(P.rnet_w + (size_t)ch * 3 + 2), __ldg(P.rnet_b + ch))
                           : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  if (warp == kStreamWarps) build_records(tile, recs[0]);
  __syncthreads();

  for (int it = 0; tile < tiles; tile += gridDim.x, ++it) {
    TileRec* cur_recs = recs[it & 1];
    if (warp == kStreamWarps) {
      // ============ record warp: the next tile's records, overlapped with the streaming ============
      const long long next = tile + gridDim.x;
      if (next < tiles) build_records(next, recs[(it + 1) & 1]);
    } else {
      // ================================== streaming warps =========================================
      const int tx = (int)(tile % tiles_x);
      const int ty = (int)((tile / tiles_x) % tiles_y);
      const int n = (int)(tile / ((long long)tiles_x * tiles_y));
      const int x = tx * kTileW + warp;

      // Fgfa_net: cosine of the two embeddings of each pixel of this warp's column, shuffle reduce
      if (P.mode == LSFA_W_COSINE) {
        for (int r = 0; r < kTileH; ++r) {
          TileRec& rec = cur_recs[r * kTileW + warp];
          if ((rec.use & (3 << 8)) != (1 << 8)) continue;
          const size_t q = (size_t)n * P.HW + (size_t)(ty * kTileH + r) * P.W + x;
          const T* __restrict__ ew = static_cast<const T*>(P.emb_warp) + q * P.E;
          const T* __restrict__ ec = static_cast<const T*>(P.emb_cur) + q * P.E;
          float sww = 0.f, scc = 0.f, swc = 0.f;
          for (int e = lane * L; e < P.E; e += 2 * CSTEP) {
            const uint4 va0 = ldg_stream_v4(ew + e), vb0 = ldg_stream_v4(ec + e);
            const bool two = e + CSTEP < P.E;
            uint4 va1 = make_uint4(0, 0, 0, 0), vb1 = va1;
            if (two) {
              va1 = ldg_stream_v4(ew + e + CSTEP);
              vb1 = ldg_stream_v4(ec + e + CSTEP);
            }
            float a[L], b[L];
            V::unpack(va0, a);
            V::unpack(vb0, b);
#pragma unroll
            for (int i = 0; i < L; ++i) {
              sww = fmaf(a[i], a[i], sww);
              scc = fmaf(b[i], b[i], scc);
              swc = fmaf(a[i], b[i], swc);
            }
            V::unpack(va1, a);
            V::unpack(vb1, b);
#pragma unroll
            for (int i = 0; i < L; ++i) {
              sww = fmaf(a[i], a[i], sww);
              scc = fmaf(b[i], b[i], scc);
              swc = fmaf(a[i], b[i], swc);
            }
          }
          sww = warp_sum(sww);
          scc = warp_sum(scc);
          swc = warp_sum(swc);
          __syncwarp();
          if (lane == 0) {
            const float nw = sqrtf(sww + 1e-10f), nc = sqrtf(scc + 1e-10f);
            float bw, bc;
            softmax2(swc / (nw * nc), scc / (nc * nc), bw, bc);
            rec.ww = bw; rec.wc = bc;
            rec.w00 *= bw; rec.w01 *= bw; rec.w10 *= bw; rec.w11 *= bw;
          }
          __syncwarp();
        }
      }

      if (x < P.W) {
        const int kn = key_slot(P, n);
        const T* kbase = key + (size_t)kn * P.HWk * P.C + lane * L;
#pragma unroll 1
        for (int r = 0; r < kTileH; ++r) {
          const TileRec& rec = cur_recs[r * kTileW + warp];
          const int use = rec.use;
          if (!(use & (1 << 8))) continue;          // row outside the frame (warp-uniform)
          PixelRec t;
          t.w00 = rec.w00; t.w01 = rec.w01; t.w10 = rec.w10; t.w11 = rec.w11;
          t.ww = rec.ww; t.wc = rec.wc;
          const bool bp = (use & (1 << 9)) != 0;     // ChooseFeat: keep the current feature
          const f32x2 w00p = pair2(t.w00, t.w00), w01p = pair2(t.w01, t.w01), w10p = pair2(t.w10, t.w10),
                      w11p = pair2(t.w11, t.w11), wcp = pair2(t.wc, t.wc), wwp = pair2(t.ww, t.ww);
          (void)wwp;
          const int p = (ty * kTileH + r) * P.W + x;
          const size_t obase = ((size_t)n * P.HW + p) * P.C + lane * L;
          // invalid taps carry weight 0 and an in-bounds (clamped) index: every load is unconditional
          const T* k00 = kbase + (size_t)rec.i00 * P.C;
          const T* k01 = kbase + (size_t)rec.i01 * P.C;
          const T* k10 = kbase + (size_t)rec.i10 * P.C;
          const T* k11 = kbase + (size_t)rec.i11 * P.C;
          const T* ps = scale + obase;     // only dereferenced when has_scale / has_cur
          const T* pc = cur + obase;
          T* po = out + obase;
          float r0 = 0.f, r1 = 0.f, r2 = 0.f;
          if (HAS_RES && has_res) {
            r0 = __ldg(P.res + ((size_t)n * 3 + 0) * P.HW + p);
            r1 = __ldg(P.res + ((size_t)n * 3 + 1) * P.HW + p);
            r2 = __ldg(P.res + ((size_t)n * 3 + 2) * P.HW + p);
          }
          // two channel chunks (2 x 32 lanes x 16 B per stream) in flight: 12 independent loads
#pragma unroll 1
          for (int c = 0; c + lane * L < P.C; c += 2 * CSTEP, k00 += 2 * CSTEP, k01 += 2 * CSTEP, k10 += 2 * CSTEP,
                   k11 += 2 * CSTEP, ps += 2 * CSTEP, pc += 2 * CSTEP, po += 2 * CSTEP) {
            const bool two = c + CSTEP + lane * L < P.C;
            const uint4 z = make_uint4(0, 0, 0, 0);
            uint4 v[2][6];
            if (kNhwcPrefetch) {
              // L2 prefetch of the NEXT batch (same pixel, next channel chunks; or the first chunks of the next
              // row's pixel): the loads of that batch then wait an L2 round trip instead of a DRAM one, with
              // no registers held meanwhile.  The streaming warps were latency bound (long_scoreboard 4.3/issue).
              if (c + 2 * CSTEP + lane * L < P.C) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  const int co = (2 + h) * CSTEP;
                  if (h == 1 && !(c + 3 * CSTEP + lane * L < P.C)) break;
                  if (!bp) {
                    prefetch_l2(k00 + co); prefetch_l2(k01 + co); prefetch_l2(k10 + co); prefetch_l2(k11 + co);
                    if (has_scale) prefetch_l2(ps + co);
                  }
                  if (has_cur) prefetch_l2(pc + co);
                }
              } else if (r + 1 < kTileH) {
                const TileRec& nx = cur_recs[(r + 1) * kTileW + warp];
                if (nx.use & (1 << 8)) {
                  const bool nbp = (nx.use & (1 << 9)) != 0;
                  const size_t nob = obase + (size_t)P.W * P.C;
#pragma unroll
                  for (int h = 0; h < 2; ++h) {
                    const int co = h * CSTEP;
                    if (!(co + lane * L < P.C)) break;
                    if (!nbp) {
                      prefetch_l2(kbase + (size_t)nx.i00 * P.C + co); prefetch_l2(kbase + (size_t)nx.i01 * P.C + co);
                      prefetch_l2(kbase + (size_t)nx.i10 * P.C + co); prefetch_l2(kbase + (size_t)nx.i11 * P.C + co);
                      if (has_scale) prefetch_l2(scale + nob + co);
                    }
                    if (has_cur) prefetch_l2(cur + nob + co);
                  }
                }
              }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int co = h * CSTEP;            // compile-time offset from the running pointers
              const bool on = h == 0 || two;
              v[h][0] = (on && !bp) ? ldg_cached_v4(k00 + co) : z;
              v[h][1] = (on && !bp) ? ldg_cached_v4(k01 + co) : z;
              v[h][2] = (on && !bp) ? ldg_cached_v4(k10 + co) : z;
              v[h][3] = (on && !bp) ? ldg_cached_v4(k11 + co) : z;
              v[h][4] = (on && !bp && has_scale) ? ldg_stream_v4(ps + co) : z;
              v[h][5] = (on && has_cur) ? ldg_stream_v4(pc + co) : z;
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              if (h == 1 && !two) break;
              const int co = h * CSTEP;
              // packed dual-fp32 math (FFMA2/FMUL2): half the FP instructions, same bits as scalar fmaf
              constexpr int H2 = L / 2;
              f32x2 f00[H2], f01[H2], f10[H2], f11[H2], fs[H2], fc[H2], o[H2];
              V::unpack2(v[h][0], f00);
              V::unpack2(v[h][1], f01);
              V::unpack2(v[h][2], f10);
              V::unpack2(v[h][3], f11);
              V::unpack2(v[h][4], fs);
              V::unpack2(v[h][5], fc);
              if (bp) {
#pragma unroll
                for (int i = 0; i < H2; ++i) o[i] = fc[i];
              } else {
#pragma unroll
                for (int i = 0; i < H2; ++i) {
                  f32x2 val = mul2(w00p, f00[i]);
                  val = fma2(w01p, f01[i], val);
                  val = fma2(w10p, f10[i], val);
                  val = fma2(w11p, f11[i], val);
                  if (has_scale) val = mul2(val, fs[i]);
                  if (HAS_RES && has_res) {
                    float ra, rb;
                    if (P.rnet_smem) {
                      const float4 qa = rnet_s[c + co + (2 * i) * 32 + lane], qb = rnet_s[c + co + (2 * i + 1) * 32 + lane];
                      ra = rnet_term(qa.x, qa.y, qa.z, qa.w, r0, r1, r2);
                      rb = rnet_term(qb.x, qb.y, qb.z, qb.w, r0, r1, r2);
                    } else {
                      const int ch = c + co + lane * L + 2 * i;
                      const float* rw = P.rnet_w + (size_t)ch * 3;
                      ra = rnet_term(__ldg(rw), __ldg(rw + 1), __ldg(rw + 2), __ldg(P.rnet_b + ch), r0, r1, r2);
                      rb = rnet_term(__ldg(rw + 3), __ldg(rw + 4), __ldg(rw + 5), __ldg(P.rnet_b + ch + 1), r0, r1, r2);
                    }
                    val = fma2(wwp, pair2(ra, rb), val);
                  }
                  o[i] = has_cur ? fma2(wcp, fc[i], val) : val;
                }
              }
              if (req_add) {
                f32x2 bq[H2];
                V::unpack2(*reinterpret_cast<const uint4*>(po + co), bq);
                const f32x2 one = pair2(1.0f, 1.0f);
#pragma unroll
                for (int i = 0; i < H2; ++i) o[i] = fma2(one, bq[i], o[i]);
              }
              stg_stream_v4(po + co, V::pack2(o));
            }
          }
        }
      }
    }
    __syncthreads();   // next tile's records are complete; this tile's records may be overwritten
  }
}

// cosine logits on their own (NHWC embeddings) - same reduction as inside the fused kernel
template <typename T>
__global__ void __launch_bounds__(kNhwcThreads)
cosine_logits_nhwc_kernel(const T* __restrict__ ew_all, const T* __restrict__ ec_all,
                          float* __restrict__ logits, int N, int E, int HW) {
  using V = Vec16<T>;
  constexpr int L = V::kLanes;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long total = (long long)N * HW;
  for (long long q = (long long)blockIdx.x * kNhwcWarps + warp; q < total;
       q += (long long)gridDim.x * kNhwcWarps) {
    const T* ew = ew_all + (size_t)q * E;
    const T* ec = ec_all + (size_t)q * E;
    float sww = 0.f, scc = 0.f, swc = 0.f;
    for (int e = lane * L; e < E; e += 32 * L) {
      float a[L], b[L];
      V::unpack(ldg_stream_v4(ew + e), a);
      V::unpack(ldg_stream_v4(ec + e), b);
#pragma unroll
      for (int i = 0; i < L; ++i) {
        sww = fmaf(a[i], a[i], sww);
        scc = fmaf(b[i], b[i], scc);
        swc = fmaf(a[i], b[i], swc);
      }
    }
    sww = warp_sum(sww);
    scc = warp_sum(scc);
    swc = warp_sum(swc);
    if (lane == 0) {
      const int n = (int)(q / HW);
      const int p = (int)(q - (long long)n * HW);
      const float nw = sqrtf(sww + 1e-10f), nc = sqrtf(scc + 1e-10f);
      logits[((size_t)n * 2 + 0) * HW + p] = swc / (nw * nc);
      logits[((size_t)n * 2 + 1) * HW + p] = scc / (nc * nc);
    }
  }
}

static int nhwc_grid(long long work_groups) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long g = (long long)sms * 8;   // 8 resident 256-thread CTAs per SM
  if (g > work_groups) g = work_groups;
  if (g < 1) g = 1;
  return (int)g;
}

cudaError_t launch_agg_nhwc_win(const AggParams& P_in, bool bf16, int var, cudaStream_t st);   // aggregate_nhwc_win.cu

// kernel: 0 = auto, 1 = LDG/STG tile kernel, 3 = all-TMA gather-by-bulk-copy or cudaErrorNotSupported,
// 5 = window-resident all-TMA (tensor maps) or cudaErrorNotSupported
cudaError_t launch_agg_nhwc(const AggParams& P_in, bool bf16, int kernel, cudaStream_t st) {
  AggParams P = P_in;
  const int tiles_x = (P.W + kTileW - 1) / kTileW, tiles_y = (P.H + kTileH - 1) / kTileH;
  const long long tiles = (long long)P.N * tiles_x * tiles_y;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long grid = (long long)sms * 2;    // persistent: 2 resident CTAs per SM
  if (grid > tiles) grid = tiles;
  const bool has_scale = P.scale != nullptr, has_cur = P.mode != LSFA_W_NONE, has_res = P.res != nullptr;
  int var = kVarRuntime;
  if (!P.req_add) {
    if (!has_scale && !has_cur && !has_res) var = kVarWarpOnly;
    else if (has_scale && !has_cur && !has_res) var = kVarScale;
    else if (has_scale && has_cur && !has_res) var = kVarScaleCur;
    else if (!has_scale && has_cur && has_res) var = kVarResCur;
  }
  if (kernel == 5) return launch_agg_nhwc_win(P, bf16, var, st);
  NtPlan Q;
  const char* env = knob("LSFA_NHWC_TMA");                    // ablation knob: LSFA_NHWC_TMA=0 keeps the LDG/STG kernel
  bool want_tma = kernel == 3 || (kernel == 0 && !(env && env[0] == '0'));
  // bf16 blend variants on small batches: the two kernels are within 1-2 % at large batch, and the all-TMA kernel
  // pays ~12 us of ramp-up and tail per launch (measured: 64 frames of 1024x38x63 bf16 0.887 vs 0.919 of the peak for
  // the tile kernel; 512 frames 0.939 vs 0.929).  Below ~64 record batches per SM the tile kernel is chosen.
  if (kernel == 0 && bf16 && (var == kVarScaleCur || var == kVarScale) &&
      (long long)P.N * ((P.HW + 31) / 32) < 64LL * sms)
    want_tma = false;
  if (want_tma && plan_nhwc_tma(P, bf16, var, &Q)) {
    if (knob("LSFA_TMA_STATIC")) P.sched = nullptr;
    if (P.sched) {                                              // one claim counter, zeroed per launch
      cudaError_t e = cudaMemsetAsync(P.sched, 0, sizeof(unsigned), st);
      if (e != cudaSuccess) return e;
    }
    const long long nb = (long long)P.N * ((P.HW + 31) / 32);
    const int g = (int)(nb < sms ? nb : sms);                   // persistent: one CTA per SM
#define LSFA_NT_CASE(V)                                                                         \
    if (var == V) return bf16 ? launch_nhwc_tma_variant<__nv_bfloat16, V>(P, Q, g, st)          \
                              : launch_nhwc_tma_variant<float, V>(P, Q, g, st);
    LSFA_NT_CASE(kVarWarpOnly) LSFA_NT_CASE(kVarScale) LSFA_NT_CASE(kVarScaleCur) LSFA_NT_CASE(kVarResCur)
#undef LSFA_NT_CASE
  }
  if (kernel == 3) return cudaErrorNotSupported;
  // residual variant: the conv weights go to (dynamic) shared memory when they fit next to a second CTA
  size_t dsm = 0;
  P.rnet_smem = 0;
  if (has_res) {
    const int cstep = 32 * (bf16 ? 8 : 4);
    const size_t need = (size_t)((P.C + cstep - 1) / cstep * cstep) * sizeof(float4);
    if (need <= 40 * 1024 && knob("LSFA_NHWC_RNET_GLOBAL") == nullptr) {
      dsm = need;
      P.rnet_smem = 1;
    }
  }
#define LSFA_NHWC_CASE(V)                                                                              \
  if (var == V) {                                                                                      \
    if (bf16) agg_nhwc_kernel<__nv_bfloat16, V><<<(unsigned)grid, kNhwcTileThreads, dsm, st>>>(P);     \
    else agg_nhwc_kernel<float, V><<<(unsigned)grid, kNhwcTileThreads, dsm, st>>>(P);                  \
  }
  LSFA_NHWC_CASE(kVarRuntime) LSFA_NHWC_CASE(kVarWarpOnly) LSFA_NHWC_CASE(kVarScale)
  LSFA_NHWC_CASE(kVarScaleCur) LSFA_NHWC_CASE(kVarResCur)
#undef LSFA_NHWC_CASE
  return cudaPeekAtLastError();
}

cudaError_t launch_cosine_logits_nhwc(const void* ew, const void* ec, float* logits, int N, int E,
                                      int HW, bool bf16, cudaStream_t st) {
  const long long groups = ((long long)N * HW + kNhwcWarps - 1) / kNhwcWarps;
  const int grid = nhwc_grid(groups);
  if (bf16)
    cosine_logits_nhwc_kernel<__nv_bfloat16><<<grid, kNhwcThreads, 0, st>>>(
        static_cast<const __nv_bfloat16*>(ew), static_cast<const __nv_bfloat16*>(ec), logits, N, E, HW);
  else
    cosine_logits_nhwc_kernel<float><<<grid, kNhwcThreads, 0, st>>>(
        static_cast<const float*>(ew), static_cast<const float*>(ec), logits, N, E, HW);
  return cudaPeekAtLastError();
}

}  // namespace lsfa
