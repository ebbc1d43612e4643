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
#include "lsfa_device.cuh"

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
constexpr int kTileH = 4;
constexpr int kTileW = 8;
constexpr int kTilePix = kTileH * kTileW;
constexpr int kStreamWarps = kTileW;                       // one warp per tile column
constexpr int kNhwcTileThreads = (kStreamWarps + 1) * 32;   // + the record warp

struct __align__(16) TileRec {
  float w00, w01, w10, w11;   // tap weights, blend weight folded in
  float ww, wc;
  int i00, i01, i10, i11;     // key pixel indices of the taps (in-bounds)
  int use;                    // bit t: tap t is inside the key plane (read it), bit 8: pixel exists, bit 9: bypass
  int n;                      // frame of the tile
};

template <typename T, bool HAS_RES>
__global__ void __launch_bounds__(kNhwcTileThreads, 2)
agg_nhwc_kernel(const __grid_constant__ AggParams P) {
  using V = Vec16<T>;
  constexpr int L = V::kLanes;        // channels per 16-byte vector
  constexpr int CSTEP = 32 * L;       // channels per warp pass
  __shared__ TileRec recs[2][kTilePix];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const T* __restrict__ key = static_cast<const T*>(P.key);
  const T* __restrict__ scale = static_cast<const T*>(P.scale);
  const T* __restrict__ cur = static_cast<const T*>(P.cur);
  T* __restrict__ out = static_cast<T*>(P.out);
  const bool has_cur = P.mode != LSFA_W_NONE;
  const int tiles_x = (P.W + kTileW - 1) / kTileW, tiles_y = (P.H + kTileH - 1) / kTileH;
  const long long tiles = (long long)P.N * tiles_x * tiles_y;

  // ---- record builder: one lane per pixel of the tile (a3,a5,a6,a7,a8 index math, a13 softmax) ----
  auto build_records = [&](long long tile, TileRec* dst) {
    const int tx = (int)(tile % tiles_x);
    const int ty = (int)((tile / tiles_x) % tiles_y);
    const int n = (int)(tile / ((long long)tiles_x * tiles_y));
    const int r = lane / kTileW, cx = lane % kTileW;
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
        // taps outside the key plane are not read at all (the reference does not read them);
        // softmax weights are never 0, so a zero weight here still means "tap invalid or weightless"
        rec.use |= (t.w00 != 0.f ? 1 : 0) | (t.w01 != 0.f ? 2 : 0) | (t.w10 != 0.f ? 4 : 0) | (t.w11 != 0.f ? 8 : 0);
        rec.w00 = t.w00; rec.w01 = t.w01; rec.w10 = t.w10; rec.w11 = t.w11;
        rec.ww = t.ww; rec.wc = t.wc;
        rec.i00 = t.i00; rec.i01 = t.i01; rec.i10 = t.i10; rec.i11 = t.i11;
      }
    }
    dst[lane] = rec;
  };

  long long tile = blockIdx.x;
  if (tile >= tiles) return;
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
        const int kn = P.key_index ? __ldg(P.key_index + n) : n;
        const T* kbase = key + (size_t)kn * P.HWk * P.C;
#pragma unroll 1
        for (int r0 = 0; r0 < kTileH; r0 += 2) {
          PixelRec t[2];
          int use[2];
          size_t obase[2];
          const T* kp[2][4];
          float r3[2][3];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const TileRec& rec = cur_recs[(r0 + u) * kTileW + warp];
            t[u].w00 = rec.w00; t[u].w01 = rec.w01; t[u].w10 = rec.w10; t[u].w11 = rec.w11;
            t[u].ww = rec.ww; t[u].wc = rec.wc;
            use[u] = rec.use;
            const int p = (ty * kTileH + r0 + u) * P.W + x;
            obase[u] = ((size_t)n * P.HW + p) * P.C;
            kp[u][0] = kbase + (size_t)rec.i00 * P.C;
            kp[u][1] = kbase + (size_t)rec.i01 * P.C;
            kp[u][2] = kbase + (size_t)rec.i10 * P.C;
            kp[u][3] = kbase + (size_t)rec.i11 * P.C;
            r3[u][0] = r3[u][1] = r3[u][2] = 0.f;
            if (HAS_RES && (use[u] & (1 << 8))) {
              r3[u][0] = __ldg(P.res + ((size_t)n * 3 + 0) * P.HW + p);
              r3[u][1] = __ldg(P.res + ((size_t)n * 3 + 1) * P.HW + p);
              r3[u][2] = __ldg(P.res + ((size_t)n * 3 + 2) * P.HW + p);
            }
          }
#pragma unroll 1
          for (int c = lane * L; c < P.C; c += CSTEP) {
            const uint4 z = make_uint4(0, 0, 0, 0);
            uint4 v[2][4], vs[2], vc[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {      // every load of both rows is issued before any is used
              const bool live = (use[u] & (1 << 8)) != 0, bp = (use[u] & (1 << 9)) != 0;
#pragma unroll
              for (int k = 0; k < 4; ++k) v[u][k] = (live && !bp && (use[u] & (1 << k))) ? ldg_cached_v4(kp[u][k] + c) : z;
              vs[u] = (live && !bp && scale) ? ldg_stream_v4(scale + obase[u] + c) : z;
              vc[u] = (live && has_cur) ? ldg_stream_v4(cur + obase[u] + c) : z;
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              if (!(use[u] & (1 << 8))) continue;
              float f00[L], f01[L], f10[L], f11[L], fs[L], fc[L], o[L];
              V::unpack(v[u][0], f00);
              V::unpack(v[u][1], f01);
              V::unpack(v[u][2], f10);
              V::unpack(v[u][3], f11);
              V::unpack(vs[u], fs);
              V::unpack(vc[u], fc);
              if (use[u] & (1 << 9)) {         // ChooseFeat: keep the current feature
#pragma unroll
                for (int i = 0; i < L; ++i) o[i] = fc[i];
              } else {
#pragma unroll
                for (int i = 0; i < L; ++i) {
                  float val = tap_chain(t[u], f00[i], f01[i], f10[i], f11[i]);
                  if (scale) val *= fs[i];
                  if (HAS_RES) {
                    const float* rw = P.rnet_w + (size_t)(c + i) * 3;
                    val = fmaf(t[u].ww, rnet_term(__ldg(rw), __ldg(rw + 1), __ldg(rw + 2), __ldg(P.rnet_b + c + i),
                                                  r3[u][0], r3[u][1], r3[u][2]), val);
                  }
                  o[i] = has_cur ? fmaf(t[u].wc, fc[i], val) : val;
                }
              }
              if (P.req_add) {
                float b[L];
                V::unpack(*reinterpret_cast<const uint4*>(out + obase[u] + c), b);
#pragma unroll
                for (int i = 0; i < L; ++i) o[i] += b[i];
              }
              stg_stream_v4(out + obase[u] + c, V::pack(o));
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

cudaError_t launch_agg_nhwc(const AggParams& P, bool bf16, cudaStream_t st) {
  const int tiles_x = (P.W + kTileW - 1) / kTileW, tiles_y = (P.H + kTileH - 1) / kTileH;
  const long long tiles = (long long)P.N * tiles_x * tiles_y;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long grid = (long long)sms * 2;    // persistent: 2 resident CTAs per SM
  if (grid > tiles) grid = tiles;
  const bool res = P.res != nullptr;
  if (bf16 && res) agg_nhwc_kernel<__nv_bfloat16, true><<<(unsigned)grid, kNhwcTileThreads, 0, st>>>(P);
  else if (bf16) agg_nhwc_kernel<__nv_bfloat16, false><<<(unsigned)grid, kNhwcTileThreads, 0, st>>>(P);
  else if (res) agg_nhwc_kernel<float, true><<<(unsigned)grid, kNhwcTileThreads, 0, st>>>(P);
  else agg_nhwc_kernel<float, false><<<(unsigned)grid, kNhwcTileThreads, 0, st>>>(P);
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
