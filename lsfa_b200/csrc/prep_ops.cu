// The small operators either side of the fused kernel: MV / residual preparation
// (lib/utils/image.py:52-60, 202-228), GridGenerator(warp), the sampler-coordinate probe,
// layout converters, and the op-by-op "unfused" chain used as the ablation baseline.
#include "lsfa_device.cuh"

namespace lsfa {

// ---------------------------------------------------------------------------------------
// a3+a5+a6: raw MV (N,h,w,2) -> flow (N,2,H,W) f32     image.py:207-215,220-228
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void mv_pool_kernel(const T* __restrict__ mv, float* __restrict__ flow, int N, int h,
                               int w, int H, int W, double scale, int mode) {
  const long long total = (long long)N * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / (H * W));
    const int p = (int)(i - (long long)n * H * W);
    const int y = p / W, x = p - y * W;
    const T* img = mv + (size_t)n * h * w * 2;
    const double fx = __dmul_rn(pool_cell(img, h, w, 2, y, x, 0, mode), scale);
    const double fy = __dmul_rn(pool_cell(img, h, w, 2, y, x, 1, mode), scale);
    flow[((size_t)n * 2 + 0) * H * W + p] = (float)fx;
    flow[((size_t)n * 2 + 1) * H * W + p] = (float)fy;
  }
}

// a3+a4+a5 for the residual, with the in-place aliasing of image.py:217-218:
//   i=0: ch0 = (ch2 - m2)*s ; i=1: ch1 = (ch1 - m1)*s ; i=2: ch2 = (ch0_new - m0)*s
// applied to the zero-padded float64 image BEFORE the stride-16 reduction.
struct ResMeans {
  double m0, m1, m2, s;
};

template <typename T>
__device__ __forceinline__ double res_value(const T* img, int h, int w, int y, int x, int ch,
                                            const ResMeans& M) {
  const double c2 = raw_at(img, h, w, 3, y, x, 2);
  const double n0 = __dmul_rn(__dsub_rn(c2, M.m2), M.s);
  if (ch == 0) return n0;
  if (ch == 1) return __dmul_rn(__dsub_rn(raw_at(img, h, w, 3, y, x, 1), M.m1), M.s);
  return __dmul_rn(__dsub_rn(n0, M.m0), M.s);
}

template <typename T>
__global__ void res_pool_kernel(const T* __restrict__ res, float* __restrict__ out, int N, int h,
                                int w, int H, int W, ResMeans M, int mode) {
  const long long total = (long long)N * 3 * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % (H * W));
    const int ch = (int)((i / (H * W)) % 3);
    const int n = (int)(i / ((long long)3 * H * W));
    const int y = p / W, x = p - y * W;
    const T* img = res + (size_t)n * h * w * 3;
    double v;
    if (mode == LSFA_POOL_CENTRE2X2) {
      const double a = res_value(img, h, w, 16 * y + 7, 16 * x + 7, ch, M);
      const double b = res_value(img, h, w, 16 * y + 7, 16 * x + 8, ch, M);
      const double c = res_value(img, h, w, 16 * y + 8, 16 * x + 7, ch, M);
      const double d = res_value(img, h, w, 16 * y + 8, 16 * x + 8, ch, M);
      v = __dmul_rn(__dadd_rn(__dadd_rn(a, b), __dadd_rn(c, d)), 0.25);
    } else {
      double acc = 0.0;
      for (int r = 0; r < 16; ++r)
        for (int q = 0; q < 16; ++q)
          acc = __dadd_rn(acc, res_value(img, h, w, 16 * y + r, 16 * x + q, ch, M));
      v = __dmul_rn(acc, 1.0 / 256.0);
    }
    out[i] = (float)v;
  }
}

// ---------------------------------------------------------------------------------------
// a1+a2: sign, h-flip (image.py:53-60) and cv2.resize(fx=fy=im_scale, INTER_LINEAR) on
// float32 (image.py:204): horizontal pass then vertical pass, float32, no FMA; coefficient
// fx = (float)((d+0.5)/im_scale - 0.5), floor/frac split, edge rules of cv::resize.
// ---------------------------------------------------------------------------------------
__global__ void mv_prepare_kernel(const int* __restrict__ in, float* __restrict__ out, int N,
                                  int h, int w, int oh, int ow, double inv_scale, int identity,
                                  int negate, int hflip) {
  const long long total = (long long)N * oh * ow;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % ow);
    const int oy = (int)((i / ow) % oh);
    const int n = (int)(i / ((long long)ow * oh));
    const int* img = in + (size_t)n * h * w * 2;
    float r[2];
    r[0] = coviar_resized(img, h, w, oy, ox, 0, inv_scale, identity, negate, hflip);
    r[1] = coviar_resized(img, h, w, oy, ox, 1, inv_scale, identity, negate, hflip);
    reinterpret_cast<float2*>(out)[i] = make_float2(r[0], r[1]);
  }
}

// ---------------------------------------------------------------------------------------
// a1+a2+a3+a4+a5 for the residual in ONE pass over coviar's int32 (N,h,w,3) array: h-flip
// (image.py:59), cv2.resize(fx=fy=im_scale, INTER_LINEAR) on float32 (image.py:205; same routine as the
// MV's: horizontal pass, vertical pass, no FMA), zero pad to stride 16, the aliased colour/mean loop
// (image.py:217-218) and the stride-16 reduction (image.py:221).  Only the resized samples the reduction
// reads are ever computed (4 per cell in parity mode): no (oh,ow,3) float image, no float64 temporaries.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float res_src(const int* __restrict__ img, int w, int y, int x, int ch, int hflip) {
  const int xs = hflip ? (w - 1 - x) : x;
  return (float)__ldg(img + ((size_t)y * w + xs) * 3 + ch);
}

__device__ __forceinline__ double res_resized(const int* __restrict__ img, int h, int w, int oh, int ow, int oy, int ox,
                                              int ch, double inv_scale, int identity, int hflip) {
  if (oy >= oh || ox >= ow) return 0.0;                              // the pad of image.py:207-215
  if (identity) return (double)res_src(img, w, oy, ox, ch, hflip);
  int sx, sy;
  float fx, fy;
  linear_coeff(ox, w, inv_scale, sx, fx);
  linear_coeff(oy, h, inv_scale, sy, fy);
  const int sx1 = min(sx + 1, w - 1), sy1 = min(sy + 1, h - 1);
  const float a0 = __fsub_rn(1.0f, fx), b0 = __fsub_rn(1.0f, fy);
  const float t0 = __fadd_rn(__fmul_rn(res_src(img, w, sy, sx, ch, hflip), a0), __fmul_rn(res_src(img, w, sy, sx1, ch, hflip), fx));
  const float t1 = __fadd_rn(__fmul_rn(res_src(img, w, sy1, sx, ch, hflip), a0), __fmul_rn(res_src(img, w, sy1, sx1, ch, hflip), fx));
  return (double)__fadd_rn(__fmul_rn(t0, b0), __fmul_rn(t1, fy));
}

__global__ void res_coviar_pool_kernel(const int* __restrict__ res, float* __restrict__ out, int N, int h, int w, int oh,
                                       int ow, int H, int W, double inv_scale, int identity, int hflip, ResMeans M,
                                       int mode) {
  const long long total = (long long)N * 3 * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % (H * W));
    const int ch = (int)((i / (H * W)) % 3);
    const int n = (int)(i / ((long long)3 * H * W));
    const int y = p / W, x = p - y * W;
    const int* img = res + (size_t)n * h * w * 3;
    auto at = [&](int ry, int rx) -> double {                        // padded image after the aliased loop
      const double c2 = res_resized(img, h, w, oh, ow, ry, rx, 2, inv_scale, identity, hflip);
      const double n0 = __dmul_rn(__dsub_rn(c2, M.m2), M.s);
      if (ch == 0) return n0;
      if (ch == 1) return __dmul_rn(__dsub_rn(res_resized(img, h, w, oh, ow, ry, rx, 1, inv_scale, identity, hflip), M.m1), M.s);
      return __dmul_rn(__dsub_rn(n0, M.m0), M.s);
    };
    double v;
    if (mode == LSFA_POOL_CENTRE2X2) {
      const double a = at(16 * y + 7, 16 * x + 7), b = at(16 * y + 7, 16 * x + 8);
      const double c = at(16 * y + 8, 16 * x + 7), d = at(16 * y + 8, 16 * x + 8);
      v = __dmul_rn(__dadd_rn(__dadd_rn(a, b), __dadd_rn(c, d)), 0.25);
    } else {
      double acc = 0.0;
#pragma unroll 1
      for (int r = 0; r < 16; ++r)
#pragma unroll 1
        for (int q = 0; q < 16; ++q) acc = __dadd_rn(acc, at(16 * y + r, 16 * x + q));
      v = __dmul_rn(acc, 1.0 / 256.0);
    }
    out[i] = (float)v;
  }
}

// ---------------------------------------------------------------------------------------
// a7: GridGenerator(warp)
// ---------------------------------------------------------------------------------------
__global__ void grid_generator_warp_kernel(const float* __restrict__ flow, float* __restrict__ grid,
                                           int N, int H, int W, float half_w, float half_h) {
  const long long total = (long long)N * 2 * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % (H * W));
    const int ch = (int)((i / (H * W)) & 1);
    const int y = p / W, x = p - y * W;
    grid[i] = ch == 0 ? exact_grid(flow[i], (float)x, half_w) : exact_grid(flow[i], (float)y, half_h);
  }
}

// parity probe of a7+a8 index math
__global__ void sampler_coords_kernel(const float* __restrict__ fg, int is_grid, int* __restrict__ x0,
                                      int* __restrict__ y0, float* __restrict__ wx,
                                      float* __restrict__ wy, int N, int H, int W, float half_w,
                                      float half_h, float wk_m1, float hk_m1) {
  const long long total = (long long)N * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / (H * W));
    const int p = (int)(i - (long long)n * H * W);
    const int y = p / W, x = p - y * W;
    float gx = fg[((size_t)n * 2) * H * W + p], gy = fg[((size_t)n * 2 + 1) * H * W + p];
    if (!is_grid) {
      gx = exact_grid(gx, (float)x, half_w);
      gy = exact_grid(gy, (float)y, half_h);
    }
    const float xr = exact_denorm(gx, wk_m1), yr = exact_denorm(gy, hk_m1);
    const int ix = exact_floor_index(xr), iy = exact_floor_index(yr);
    x0[i] = ix;
    y0[i] = iy;
    wx[i] = exact_tl_weight(xr, ix);
    wy[i] = exact_tl_weight(yr, iy);
  }
}

// ---------------------------------------------------------------------------------------
// layout converters (harness only): NCHW f32 <-> NHWC {f32,bf16}; 32x32 smem transpose
// ---------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, T* __restrict__ dst, int C, int HW) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r, p = p0 + threadIdx.x;
    tile[r][threadIdx.x] = (c < C && p < HW) ? src[((size_t)n * C + c) * HW + p] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int p = p0 + r, c = c0 + threadIdx.x;
    if (c < C && p < HW) dst[((size_t)n * HW + p) * C + c] = from_f32<T>(tile[threadIdx.x][r]);
  }
}

template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ src, float* __restrict__ dst, int C, int HW) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int p = p0 + r, c = c0 + threadIdx.x;
    tile[r][threadIdx.x] = (c < C && p < HW) ? to_f32(src[((size_t)n * HW + p) * C + c]) : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r, p = p0 + threadIdx.x;
    if (c < C && p < HW) dst[((size_t)n * C + c) * HW + p] = tile[threadIdx.x][r];
  }
}

// ---------------------------------------------------------------------------------------
// The reference graph op by op (ablation baseline): each kernel is one MXNet operator and
// one full pass over its operands, exactly the traffic SURVEY.md section 8d charges.
// ---------------------------------------------------------------------------------------
// BilinearSampler as MXNet's GPU kernel shapes it: one thread per output element.
__global__ void unfused_sampler_kernel(const float* __restrict__ data, const float* __restrict__ grid,
                                       float* __restrict__ out, int N, int C, int H, int W) {
  const long long total = (long long)N * C * H * W;
  const int HW = H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW);
    const long long nc = i / HW;
    const int n = (int)(nc / C);
    const float gx = grid[((size_t)n * 2) * HW + p], gy = grid[((size_t)n * 2 + 1) * HW + p];
    const PixelRec t = make_taps(gx, gy, H, W, (float)(W - 1), (float)(H - 1));
    const float* plane = data + (size_t)nc * HW;
    out[i] = tap_chain(t, plane[t.i00], plane[t.i01], plane[t.i10], plane[t.i11]);
  }
}
__global__ void ew_mul_kernel(const float* __restrict__ a, const float* __restrict__ b,
                              float* __restrict__ o, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    o[i] = a[i] * b[i];
}
__global__ void ew_add_kernel(const float* __restrict__ a, const float* __restrict__ b,
                              float* __restrict__ o, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    o[i] = a[i] + b[i];
}
// softmax(axis=0) over the two logits maps, in place into w (N,2,HW)
__global__ void softmax_pair_kernel(const float* __restrict__ logits, float* __restrict__ w, int N, int HW) {
  const long long total = (long long)N * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / HW), p = (int)(i % HW);
    float a, b;
    softmax2(logits[((size_t)n * 2) * HW + p], logits[((size_t)n * 2 + 1) * HW + p], a, b);
    w[((size_t)n * 2) * HW + p] = a;
    w[((size_t)n * 2 + 1) * HW + p] = b;
  }
}
// mx.symbol.tile(weights[k], reps=(1,C,1,1)): materialise the broadcast (N,C,HW)
__global__ void tile_channels_kernel(const float* __restrict__ w, int which, float* __restrict__ o,
                                     int N, int C, int HW) {
  const long long total = (long long)N * C * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW);
    const int n = (int)(i / ((long long)C * HW));
    o[i] = w[((size_t)n * 2 + which) * HW + p];
  }
}

// a15: ChooseFeat (operator_py/choose_feat.py:23-31) with the flag left on the device:
// out[n] = flag[n] ? conv_feat[n] : conv_feat_prop[n]; only the selected source is read.
__global__ void choose_feat_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                   const unsigned char* __restrict__ flag, float* __restrict__ o,
                                   long long per_frame, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / per_frame;
    o[i] = flag[n] ? a[i] : b[i];
  }
}

// K2 of the exact two-phase key-frame graph: the Nq / Fgfa tail on two materialised features
// (SYM:104-108, 141-147): out = w1 * src0 + w2 * cur with (w1,w2) = softmax over the two logits,
// bypass frames keep cur (choose_feat.py:23-31).  Thread = pixel (weights computed once), loop
// over a slice of channels: every access is coalesced along the pixel axis.  3F of traffic.
constexpr int kBlendCG = 32;
__global__ void __launch_bounds__(256)
blend_logits_kernel(const float* __restrict__ src0, const float* __restrict__ cur, const float* __restrict__ logits,
                    const unsigned char* __restrict__ bypass, float* __restrict__ out, int C, int HW) {
  const int n = blockIdx.z;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  const int c0 = blockIdx.y * kBlendCG, c1 = min(C, c0 + kBlendCG);
  const bool byp = bypass != nullptr && __ldg(bypass + n) != 0;
  float ww = 0.f, wc = 1.f;
  if (!byp) softmax2(__ldg(logits + ((size_t)n * 2) * HW + p), __ldg(logits + ((size_t)n * 2 + 1) * HW + p), ww, wc);
  size_t e = ((size_t)n * C + c0) * HW + p;
  for (int c = c0; c < c1; ++c, e += HW) {
    const float b = ldg_stream(cur + e);
    out[e] = byp ? b : fmaf(wc, b, ww * ldg_stream(src0 + e));
  }
}

// ---------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------
static inline int ew_grid(long long total, int threads) {
  long long g = (total + threads - 1) / threads;
  const long long cap = 148LL * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

cudaError_t launch_choose_feat(const float* a, const float* b, const unsigned char* flag, float* o,
                               long long per_frame, long long total, cudaStream_t st);

cudaError_t launch_mv_pool(const void* mv, bool is_i32, float* flow, int N, int h, int w, int H, int W,
                           double scale, int mode, cudaStream_t st) {
  const long long total = (long long)N * H * W;
  if (is_i32)
    mv_pool_kernel<int><<<ew_grid(total, 256), 256, 0, st>>>(static_cast<const int*>(mv), flow, N, h, w, H, W, scale, mode);
  else
    mv_pool_kernel<float><<<ew_grid(total, 256), 256, 0, st>>>(static_cast<const float*>(mv), flow, N, h, w, H, W, scale, mode);
  return cudaPeekAtLastError();
}

cudaError_t launch_res_pool(const void* res, bool is_i32, float* out, int N, int h, int w, int H, int W,
                            const double* means, double pixel_scale, int mode, cudaStream_t st) {
  ResMeans M{means ? means[0] : 0.0, means ? means[1] : 0.0, means ? means[2] : 0.0, pixel_scale};
  const long long total = (long long)N * 3 * H * W;
  if (is_i32)
    res_pool_kernel<int><<<ew_grid(total, 256), 256, 0, st>>>(static_cast<const int*>(res), out, N, h, w, H, W, M, mode);
  else
    res_pool_kernel<float><<<ew_grid(total, 256), 256, 0, st>>>(static_cast<const float*>(res), out, N, h, w, H, W, M, mode);
  return cudaPeekAtLastError();
}

cudaError_t launch_res_coviar_pool(const int* res, float* out, int N, int h, int w, int oh, int ow, int H, int W,
                                   double im_scale, int hflip, const double* means, double pixel_scale, int mode,
                                   cudaStream_t st) {
  ResMeans M{means ? means[0] : 0.0, means ? means[1] : 0.0, means ? means[2] : 0.0, pixel_scale};
  const long long total = (long long)N * 3 * H * W;
  res_coviar_pool_kernel<<<ew_grid(total, 256), 256, 0, st>>>(res, out, N, h, w, oh, ow, H, W, 1.0 / im_scale,
                                                              im_scale == 1.0 ? 1 : 0, hflip, M, mode);
  return cudaPeekAtLastError();
}

cudaError_t launch_mv_prepare(const int* in, float* out, int N, int h, int w, int oh, int ow,
                              double im_scale, int negate, int hflip, cudaStream_t st) {
  const long long total = (long long)N * oh * ow;
  const int identity = (im_scale == 1.0) ? 1 : 0;
  mv_prepare_kernel<<<ew_grid(total, 256), 256, 0, st>>>(in, out, N, h, w, oh, ow, 1.0 / im_scale,
                                                         identity, negate, hflip);
  return cudaPeekAtLastError();
}

cudaError_t launch_grid_generator(const float* flow, float* grid, int N, int H, int W, float half_w,
                                  float half_h, cudaStream_t st) {
  const long long total = (long long)N * 2 * H * W;
  grid_generator_warp_kernel<<<ew_grid(total, 256), 256, 0, st>>>(flow, grid, N, H, W, half_w, half_h);
  return cudaPeekAtLastError();
}

cudaError_t launch_sampler_coords(const float* fg, int is_grid, int* x0, int* y0, float* wx, float* wy,
                                  int N, int H, int W, float half_w, float half_h, float wk_m1,
                                  float hk_m1, cudaStream_t st) {
  const long long total = (long long)N * H * W;
  sampler_coords_kernel<<<ew_grid(total, 256), 256, 0, st>>>(fg, is_grid, x0, y0, wx, wy, N, H, W,
                                                             half_w, half_h, wk_m1, hk_m1);
  return cudaPeekAtLastError();
}

cudaError_t launch_nchw_to_nhwc(const float* src, void* dst, int N, int C, int HW, bool bf16, cudaStream_t st) {
  dim3 grid((HW + 31) / 32, (C + 31) / 32, N), block(32, 8);
  if (bf16) nchw_to_nhwc_kernel<__nv_bfloat16><<<grid, block, 0, st>>>(src, static_cast<__nv_bfloat16*>(dst), C, HW);
  else nchw_to_nhwc_kernel<float><<<grid, block, 0, st>>>(src, static_cast<float*>(dst), C, HW);
  return cudaPeekAtLastError();
}

cudaError_t launch_nhwc_to_nchw(const void* src, float* dst, int N, int C, int HW, bool bf16, cudaStream_t st) {
  dim3 grid((HW + 31) / 32, (C + 31) / 32, N), block(32, 8);
  if (bf16) nhwc_to_nchw_kernel<__nv_bfloat16><<<grid, block, 0, st>>>(static_cast<const __nv_bfloat16*>(src), dst, C, HW);
  else nhwc_to_nchw_kernel<float><<<grid, block, 0, st>>>(static_cast<const float*>(src), dst, C, HW);
  return cudaPeekAtLastError();
}

// Nq_net tail of the key-frame graph, operator by operator (SYM:468-470, 104-108):
//   grid = GridGenerator(flow); t0 = BilinearSampler(key, grid); t1 = t0 * scale_map;
//   w = softmax(logits, axis=0); t2 = tile(w[0]); t3 = tile(w[1]);
//   t0 = t2 * t1; t4 = t3 * cur; out = t0 + t4
constexpr int kUnfusedLaunches = 9;
cudaError_t launch_unfused_chain(const float* key, const float* flow, const float* scale_map,
                                 const float* cur, const float* logits, float* out, float* tmp, int N,
                                 int C, int H, int W, float half_w, float half_h, cudaStream_t st) {
  const int HW = H * W;
  const long long F = (long long)N * C * HW;
  float* t0 = tmp;
  float* t1 = tmp + F;
  float* t2 = tmp + 2 * F;
  float* t3 = tmp + 3 * F;
  float* t4 = tmp + 4 * F;
  float* grid = t4;                        // (N,2,HW) scratch, consumed before t4 is written
  float* wts = t4 + (size_t)N * 2 * HW;    // (N,2,HW)
  const int g = ew_grid(F, 256);
  grid_generator_warp_kernel<<<ew_grid((long long)N * 2 * HW, 256), 256, 0, st>>>(flow, grid, N, H, W, half_w, half_h);
  unfused_sampler_kernel<<<g, 256, 0, st>>>(key, grid, t0, N, C, H, W);
  ew_mul_kernel<<<g, 256, 0, st>>>(t0, scale_map, t1, F);
  softmax_pair_kernel<<<ew_grid((long long)N * HW, 256), 256, 0, st>>>(logits, wts, N, HW);
  tile_channels_kernel<<<g, 256, 0, st>>>(wts, 0, t2, N, C, HW);
  tile_channels_kernel<<<g, 256, 0, st>>>(wts, 1, t3, N, C, HW);
  ew_mul_kernel<<<g, 256, 0, st>>>(t2, t1, t0, F);
  ew_mul_kernel<<<g, 256, 0, st>>>(t3, cur, t4, F);
  ew_add_kernel<<<g, 256, 0, st>>>(t0, t4, out, F);
  return cudaPeekAtLastError();
}
int unfused_chain_launches() { return kUnfusedLaunches; }

cudaError_t launch_blend_logits(const float* src0, const float* cur, const float* logits, const unsigned char* bypass,
                                float* out, int N, int C, int HW, cudaStream_t st) {
  dim3 grid((HW + 255) / 256, (C + kBlendCG - 1) / kBlendCG, N);
  if (grid.y > 65535 || grid.z > 65535) return cudaErrorInvalidValue;
  blend_logits_kernel<<<grid, 256, 0, st>>>(src0, cur, logits, bypass, out, C, HW);
  return cudaPeekAtLastError();
}

cudaError_t launch_choose_feat(const float* a, const float* b, const unsigned char* flag, float* o,
                               long long per_frame, long long total, cudaStream_t st) {
  choose_feat_kernel<<<ew_grid(total, 256), 256, 0, st>>>(a, b, flag, o, per_frame, total);
  return cudaPeekAtLastError();
}

}  // namespace lsfa
