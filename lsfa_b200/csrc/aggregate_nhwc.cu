// Fused warp x scale + aggregation, NHWC (channels-last) float32 / bfloat16.
//
// Same operator chain as aggregate_nchw.cu (SYM:571-576, 308/470/680, 104-108, 132-148, 236;
// operator_py/choose_feat.py:23-31), laid out for channel-vectorised access: a pixel's C
// channels are contiguous (C=1024 bf16 = 2 KB), so every one of the 4 bilinear taps, the
// scale map, the current feature and the output are 16-byte-per-lane, fully coalesced
// vector accesses (bf16x8 / f32x4).  One warp owns one output pixel; a CTA owns a run of
// 8 horizontally adjacent pixels so neighbouring taps hit in L1; the key feature's 4-tap
// reuse across rows is served by the 126 MB L2.  Arithmetic is fp32 regardless of storage.
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

template <typename T>
__global__ void __launch_bounds__(kNhwcThreads)
agg_nhwc_kernel(const __grid_constant__ AggParams P) {
  using V = Vec16<T>;
  constexpr int L = V::kLanes;        // channels per 16-byte vector
  constexpr int CSTEP = 32 * L;       // channels per warp iteration
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const T* __restrict__ key = static_cast<const T*>(P.key);
  const T* __restrict__ scale = static_cast<const T*>(P.scale);
  const T* __restrict__ cur = static_cast<const T*>(P.cur);
  T* __restrict__ out = static_cast<T*>(P.out);
  const bool has_cur = P.mode != LSFA_W_NONE;
  const long long total = (long long)P.N * P.HW;
  const long long groups = (total + kNhwcWarps - 1) / kNhwcWarps;

  for (long long g = blockIdx.x; g < groups; g += gridDim.x) {
    const long long q = g * kNhwcWarps + warp;   // global output pixel
    if (q >= total) continue;
    const int n = (int)(q / P.HW);
    const int p = (int)(q - (long long)n * P.HW);
    const size_t obase = (size_t)q * P.C;
    const bool byp = (P.bypass != nullptr) && (__ldg(P.bypass + n) != 0);

    if (byp) {
      for (int c = lane * L; c < P.C; c += CSTEP) {
        uint4 v = ldg_stream_v4(cur + obase + c);
        if (P.req_add) {
          float a[L], b[L];
          V::unpack(v, a);
          V::unpack(*reinterpret_cast<const uint4*>(out + obase + c), b);
#pragma unroll
          for (int i = 0; i < L; ++i) a[i] += b[i];
          v = V::pack(a);
        }
        stg_stream_v4(out + obase + c, v);
      }
      continue;
    }

    const int y = p / P.W, x = p - y * P.W;
    float gx, gy;
    pixel_grid(P, n, y, x, gx, gy);   // every lane computes the same record (warp-uniform)
    PixelRec t = make_taps(gx, gy, P.Hk, P.Wk, P.wk_m1, P.hk_m1);
    float ww, wc;
    if (P.mode == LSFA_W_COSINE) {
      const T* __restrict__ ew = static_cast<const T*>(P.emb_warp) + (size_t)q * P.E;
      const T* __restrict__ ec = static_cast<const T*>(P.emb_cur) + (size_t)q * P.E;
      float sww = 0.f, scc = 0.f, swc = 0.f;
      for (int e = lane * L; e < P.E; e += 2 * CSTEP) {
        uint4 va0 = ldg_stream_v4(ew + e), vb0 = ldg_stream_v4(ec + e);
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
      const float nw = sqrtf(sww + 1e-10f), nc = sqrtf(scc + 1e-10f);
      softmax2(swc / (nw * nc), scc / (nc * nc), ww, wc);
    } else {
      pixel_weights(P, n, p, ww, wc);
    }
    float r0 = 0.f, r1 = 0.f, r2 = 0.f;
    if (P.res) {
      r0 = __ldg(P.res + ((size_t)n * 3 + 0) * P.HW + p);
      r1 = __ldg(P.res + ((size_t)n * 3 + 1) * P.HW + p);
      r2 = __ldg(P.res + ((size_t)n * 3 + 2) * P.HW + p);
    }

    // taps outside the key plane are not read at all (the reference does not read them)
    const bool u00 = t.w00 != 0.f, u01 = t.w01 != 0.f, u10 = t.w10 != 0.f, u11 = t.w11 != 0.f;
    fold_blend(t, ww, wc);
    const int kn = P.key_index ? __ldg(P.key_index + n) : n;
    const T* __restrict__ kbase = key + (size_t)kn * P.HWk * P.C;
    const T* __restrict__ k00 = kbase + (size_t)t.i00 * P.C;
    const T* __restrict__ k01 = kbase + (size_t)t.i01 * P.C;
    const T* __restrict__ k10 = kbase + (size_t)t.i10 * P.C;
    const T* __restrict__ k11 = kbase + (size_t)t.i11 * P.C;

    for (int c = lane * L; c < P.C; c += CSTEP) {
      const uint4 z = make_uint4(0, 0, 0, 0);
      const uint4 v00 = u00 ? ldg_cached_v4(k00 + c) : z;
      const uint4 v01 = u01 ? ldg_cached_v4(k01 + c) : z;
      const uint4 v10 = u10 ? ldg_cached_v4(k10 + c) : z;
      const uint4 v11 = u11 ? ldg_cached_v4(k11 + c) : z;
      const uint4 vs = scale ? ldg_stream_v4(scale + obase + c) : z;
      const uint4 vc = has_cur ? ldg_stream_v4(cur + obase + c) : z;
      float f00[L], f01[L], f10[L], f11[L], fs[L], fc[L], o[L];
      V::unpack(v00, f00);
      V::unpack(v01, f01);
      V::unpack(v10, f10);
      V::unpack(v11, f11);
      V::unpack(vs, fs);
      V::unpack(vc, fc);
#pragma unroll
      for (int i = 0; i < L; ++i) {
        float v = tap_chain(t, f00[i], f01[i], f10[i], f11[i]);
        if (scale) v *= fs[i];
        if (P.res) {
          const float* rw = P.rnet_w + (size_t)(c + i) * 3;
          v = fmaf(t.ww, rnet_term(__ldg(rw), __ldg(rw + 1), __ldg(rw + 2), __ldg(P.rnet_b + c + i), r0, r1, r2), v);
        }
        o[i] = has_cur ? fmaf(t.wc, fc[i], v) : v;
      }
      if (P.req_add) {
        float b[L];
        V::unpack(*reinterpret_cast<const uint4*>(out + obase + c), b);
#pragma unroll
        for (int i = 0; i < L; ++i) o[i] += b[i];
      }
      stg_stream_v4(out + obase + c, V::pack(o));
    }
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
  const long long groups = ((long long)P.N * P.HW + kNhwcWarps - 1) / kNhwcWarps;
  const int grid = nhwc_grid(groups);
  if (bf16) agg_nhwc_kernel<__nv_bfloat16><<<grid, kNhwcThreads, 0, st>>>(P);
  else agg_nhwc_kernel<float><<<grid, kNhwcThreads, 0, st>>>(P);
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
