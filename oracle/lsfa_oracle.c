/*
 * lsfa_oracle.c - plain-C restatement of the reference's CPU path for the non-key-frame
 * propagation + aggregation step, operator by operator.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE: linked/loaded only by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.
 *
 * Why a port and not the reference itself: the arithmetic of this path lives in Apache
 * MXNet @75a9e187d (GridGenerator, BilinearSampler, elementwise, softmax, tile) which is not
 * vendored under /root/reference and cannot be built offline.  Each
 * function below restates one MXNet CPU operator in the op order and precision of its
 * published source, and is checked bit-for-bit against oracle/lsfa_oracle.py in
 * tests/test_oracle_cport_cpu.py.  Reference call-sites: SYM =
 * dff_rfcn/symbols/resnet_v1_101_flownet_rfcn.py.
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -ffp-contract=off; no -ffast-math).
 * MXNet's CPU BilinearSampler is a single-threaded 4-deep loop; here the two outer loops are
 * OpenMP-parallel so the baseline can use every host core it is given.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define EXPORT __attribute__((visibility("default")))

/* torchrun exports OMP_NUM_THREADS=1; the baseline is entitled to every host core it can use */
EXPORT void lsfa_ref_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

EXPORT int lsfa_ref_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* lib/utils/image.py:207-215,220-228 - pad to 16, stride-16 reduction (what cv2.resize
 * fx=1/16 INTER_LINEAR computes: mean of the 2x2 centre pixels, horizontal pass first) in
 * float64, times im_scale/16, cast to float32.  mv (N,h,w,2) int32 -> flow (N,2,H,W). */
EXPORT void lsfa_ref_mv_pool_i32(const int32_t* mv, float* flow, int N, int h, int w, double im_scale,
                                 int mode) {
  const int H = (h + 15) / 16, W = (w + 15) / 16;
  const double scale = im_scale * (1.0 / 16.0);
#pragma omp parallel for collapse(2) schedule(static)
  for (int n = 0; n < N; ++n)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x)
        for (int ch = 0; ch < 2; ++ch) {
          const int32_t* img = mv + (size_t)n * h * w * 2;
          double v;
#define AT(yy, xx) (((yy) < h && (xx) < w) ? (double)img[((size_t)(yy) * w + (xx)) * 2 + ch] : 0.0)
          if (mode == 0) {
            const double a = AT(16 * y + 7, 16 * x + 7), b = AT(16 * y + 7, 16 * x + 8);
            const double c = AT(16 * y + 8, 16 * x + 7), d = AT(16 * y + 8, 16 * x + 8);
            v = ((a + b) + (c + d)) * 0.25;
          } else {
            double acc = 0.0;
            for (int r = 0; r < 16; ++r)
              for (int q = 0; q < 16; ++q) acc = acc + AT(16 * y + r, 16 * x + q);
            v = acc * (1.0 / 256.0);
          }
#undef AT
          flow[(((size_t)n * 2 + ch) * H + y) * W + x] = (float)(v * scale);
        }
}

/* mx.sym.GridGenerator(transform_type='warp') - MXNet src/operator/grid_generator-inl.h:
 * grid_dst x = i - int(i/W)*W, y = int(i/W) (float32); out = (data + grid_dst) /
 * ((dim-1)/2) - 1.  SYM:306,320,468,571,678. */
EXPORT void lsfa_ref_grid_generator_warp(const float* flow, float* grid, int N, int H, int W) {
  const float half_w = (float)(((double)(float)W - 1.0) / 2.0);
  const float half_h = (float)(((double)(float)H - 1.0) / 2.0);
#pragma omp parallel for schedule(static)
  for (int n = 0; n < N; ++n)
    for (int i = 0; i < H * W; ++i) {
      const float fi = (float)i;
      const float q = (float)(int)(fi / (float)W);
      const float x = fi - q * (float)W;
      const float y = q;
      const size_t o = (size_t)n * 2 * H * W;
      grid[o + i] = (flow[o + i] + x) / half_w - 1.0f;
      grid[o + (size_t)H * W + i] = (flow[o + (size_t)H * W + i] + y) / half_h - 1.0f;
    }
}

static inline int between(int v, int lo, int hi) { return v >= lo && v <= hi; }

/* mx.sym.BilinearSampler - MXNet src/operator/bilinear_sampler.cc BilinearSamplerForward,
 * DType = float: the `1.0` literals promote parts of the expression to double exactly as
 * written there.  SYM:307,321,469,572,679. */
EXPORT void lsfa_ref_bilinear_sampler(const float* data, const float* grid, float* out, int o_n, int o_c,
                                      int i_h, int i_w, int o_h, int o_w) {
  const int i_c = o_c;
#pragma omp parallel for collapse(2) schedule(static)
  for (int n = 0; n < o_n; ++n) {
    for (int c = 0; c < o_c; ++c) {
      for (int h = 0; h < o_h; ++h) {
        for (int w = 0; w < o_w; ++w) {
          const size_t out_index = (((size_t)n * o_c + c) * o_h + h) * o_w + w;
          const size_t grid_index = (size_t)n * o_h * o_w * 2 + (size_t)h * o_w + w;
          float y_real = (*(grid + grid_index + (size_t)o_h * o_w) + 1) * (i_h - 1) / 2;
          float x_real = (*(grid + grid_index) + 1) * (i_w - 1) / 2;
          float fy = floorf(y_real), fx = floorf(x_real);
          /* static_cast<int> of an out-of-range float is UB; clamp like the NumPy oracle */
          if (!(fy >= -16777216.0f)) fy = -16777216.0f;
          if (fy > 16777216.0f) fy = 16777216.0f;
          if (!(fx >= -16777216.0f)) fx = -16777216.0f;
          if (fx > 16777216.0f) fx = 16777216.0f;
          int top_left_y = (int)fy;
          int top_left_x = (int)fx;
          float top_left_y_w = 1.0 - (y_real - top_left_y);
          float top_left_x_w = 1.0 - (x_real - top_left_x);
          const float* plane = data + ((size_t)n * i_c + c) * i_h * i_w;
          float top_left_v = 0, top_right_v = 0, bottom_left_v = 0, bottom_right_v = 0;
          if (between(top_left_x, 0, i_w - 1) && between(top_left_y, 0, i_h - 1))
            top_left_v = plane[(size_t)top_left_y * i_w + top_left_x];
          if (between(top_left_x + 1, 0, i_w - 1) && between(top_left_y, 0, i_h - 1))
            top_right_v = plane[(size_t)top_left_y * i_w + top_left_x + 1];
          if (between(top_left_x, 0, i_w - 1) && between(top_left_y + 1, 0, i_h - 1))
            bottom_left_v = plane[(size_t)(top_left_y + 1) * i_w + top_left_x];
          if (between(top_left_x + 1, 0, i_w - 1) && between(top_left_y + 1, 0, i_h - 1))
            bottom_right_v = plane[(size_t)(top_left_y + 1) * i_w + top_left_x + 1];
          *(out + out_index) = top_left_v * top_left_y_w * top_left_x_w +
                               top_right_v * top_left_y_w * (1.0 - top_left_x_w) +
                               bottom_left_v * (1.0 - top_left_y_w) * top_left_x_w +
                               bottom_right_v * (1.0 - top_left_y_w) * (1.0 - top_left_x_w);
        }
      }
    }
  }
}

/* elementwise operators of the graph: `*` (SYM:308,470,680,108,147), `+` (SYM:108,147,236,576) */
EXPORT void lsfa_ref_mul(const float* a, const float* b, float* o, size_t n) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; ++i) o[i] = a[i] * b[i];
}
EXPORT void lsfa_ref_add(const float* a, const float* b, float* o, size_t n) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; ++i) o[i] = a[i] + b[i];
}

/* mx.sym.softmax(axis=0) over the two stacked maps (SYM:104,141): exp(x-max)/sum, float32.
 * logits (N,2,HW) -> w (N,2,HW) */
EXPORT void lsfa_ref_softmax_pair(const float* logits, float* w, int N, int HW) {
#pragma omp parallel for schedule(static)
  for (int n = 0; n < N; ++n)
    for (int p = 0; p < HW; ++p) {
      const float a = logits[((size_t)n * 2) * HW + p], b = logits[((size_t)n * 2 + 1) * HW + p];
      const float m = a > b ? a : b;
      const float ea = expf(a - m), eb = expf(b - m);
      const float s = ea + eb;
      w[((size_t)n * 2) * HW + p] = ea / s;
      w[((size_t)n * 2 + 1) * HW + p] = eb / s;
    }
}

/* mx.symbol.tile(weights[k], reps=(1,C,1,1)) (SYM:105-106,145-146) */
EXPORT void lsfa_ref_tile_channels(const float* w, int which, float* o, int N, int C, int HW) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < C; ++c)
      memcpy(o + ((size_t)n * C + c) * HW, w + ((size_t)n * 2 + which) * HW, (size_t)HW * sizeof(float));
}

/* The key-frame Nq tail exactly as the graph runs it (SYM:468-470,104-108): one pass per
 * operator.  tmp: 5 feature-sized buffers.  This is what bench.py times as the CPU baseline. */
EXPORT void lsfa_ref_chain_nq(const int32_t* mv, int mv_h, int mv_w, double im_scale, const float* key,
                              const float* scale_map, const float* cur, const float* logits, float* out,
                              float* tmp, int N, int C, int H, int W) {
  const size_t HW = (size_t)H * W, F = (size_t)N * C * HW;
  float *t0 = tmp, *t1 = tmp + F, *t2 = tmp + 2 * F, *t3 = tmp + 3 * F, *t4 = tmp + 4 * F;
  float* flow = t4;                  /* (N,2,HW) scratch inside t4, consumed before t4 is written */
  float* grid = t4 + (size_t)N * 2 * HW;
  float* wts = t4 + (size_t)N * 4 * HW;
  lsfa_ref_mv_pool_i32(mv, flow, N, mv_h, mv_w, im_scale, 0);
  lsfa_ref_grid_generator_warp(flow, grid, N, H, W);
  lsfa_ref_bilinear_sampler(key, grid, t0, N, C, H, W, H, W);
  lsfa_ref_mul(t0, scale_map, t1, F);
  lsfa_ref_softmax_pair(logits, wts, N, (int)HW);
  lsfa_ref_tile_channels(wts, 0, t2, N, C, (int)HW);
  lsfa_ref_tile_channels(wts, 1, t3, N, C, (int)HW);
  lsfa_ref_mul(t2, t1, t0, F);
  lsfa_ref_mul(t3, cur, t4, F);
  lsfa_ref_add(t0, t4, out, F);
}

/* ------------------------------------------------------------------------------------------
 * coviar's accumulated MV field and residual: a literal restatement of
 * external/data_loader_py2/coviar_data_loader.c:71-177 (create_and_load_mv_residual with
 * accumulate = 1) and the identity initialisation of :318-328, minus the FFmpeg / NumPy-C-API
 * plumbing (those headers are not available offline, so the function cannot be compiled from
 * the reference file itself).  Kept in the reference's own [x][y][2] layout and loop order.
 * mvs: (T,M,6) int32 {w,h,src_x,src_y,dst_x,dst_y}, counts (T,), mv_out (height,width,2).
 * ------------------------------------------------------------------------------------------ */
EXPORT void lsfa_ref_mv_accumulate(const int32_t* mvs, const int32_t* counts, int T, int M, int height, int width,
                                   int* accu_src, int* accu_src_old, int32_t* mv_out) {
  for (int x = 0; x < width; ++x)
    for (int y = 0; y < height; ++y) {
      accu_src_old[x * height * 2 + y * 2] = x;
      accu_src_old[x * height * 2 + y * 2 + 1] = y;
    }
  memcpy(accu_src, accu_src_old, (size_t)height * width * 2 * sizeof(int));
  for (int t = 0; t < T; ++t) {
    const int n = counts[t] < M ? counts[t] : M;
    for (int i = 0; i < n; i++) {
      const int32_t* mv = mvs + ((size_t)t * M + i) * 6;
      const int w = mv[0], h = mv[1], src_x = mv[2], src_y = mv[3], dst_x = mv[4], dst_y = mv[5];
      if (dst_x - src_x != 0 || dst_y - src_y != 0) {
        for (int x_start = (-1 * w / 2); x_start < w / 2; ++x_start) {
          for (int y_start = (-1 * h / 2); y_start < h / 2; ++y_start) {
            const int p_dst_x = dst_x + x_start, p_dst_y = dst_y + y_start;
            const int p_src_x = src_x + x_start, p_src_y = src_y + y_start;
            if (p_dst_y >= 0 && p_dst_y < height && p_dst_x >= 0 && p_dst_x < width && p_src_y >= 0 &&
                p_src_y < height && p_src_x >= 0 && p_src_x < width) {
              for (int c = 0; c < 2; ++c)
                accu_src[p_dst_x * height * 2 + p_dst_y * 2 + c] = accu_src_old[p_src_x * height * 2 + p_src_y * 2 + c];
            }
          }
        }
      }
    }
    memcpy(accu_src_old, accu_src, (size_t)width * height * 2 * sizeof(int));
  }
  for (int x = 0; x < width; ++x)
    for (int y = 0; y < height; ++y) {
      mv_out[((size_t)y * width + x) * 2] = x - accu_src[x * height * 2 + y * 2];
      mv_out[((size_t)y * width + x) * 2 + 1] = y - accu_src[x * height * 2 + y * 2 + 1];
    }
}

/* coviar_data_loader.c:141-175, accumulate case: res = cur - iframe[src], src = accu = (x,y) - mv */
EXPORT void lsfa_ref_coviar_residual(const uint8_t* iframe, const uint8_t* cur, const int32_t* mv, int32_t* res,
                                     int height, int width) {
  for (int y = 0; y < height; ++y)
    for (int x = 0; x < width; ++x) {
      const int src_x = x - mv[((size_t)y * width + x) * 2], src_y = y - mv[((size_t)y * width + x) * 2 + 1];
      for (int c = 0; c < 3; ++c)
        res[((size_t)y * width + x) * 3 + c] =
            (int32_t)cur[((size_t)y * width + x) * 3 + c] - (int32_t)iframe[((size_t)src_y * width + src_x) * 3 + c];
    }
}

/* ------------------------------------------------------------------------------------------
 * Backward of the two operators (get_train_symbol, SYM:305-307,319-321).
 * MXNet src/operator/bilinear_sampler.cc BilinearSamplerBackward, DType = float, restated in
 * its loop order (n, h, w, c), its sequential float32 `+=` and its float/double promotions
 * (the `1.0` literals).  g_input and grad_grid are accumulated into: the operator zeroes them
 * first for kWriteTo (bilinear_sampler-inl.h Backward), which the caller does here.
 * Single-threaded like the original: the scatter `+=` is order-dependent.
 * ------------------------------------------------------------------------------------------ */
EXPORT void lsfa_ref_bilinear_sampler_backward(const float* data, const float* grid, const float* grad,
                                               float* g_input, float* grad_grid, int o_n, int o_c, int i_h,
                                               int i_w, int o_h, int o_w) {
  const int i_c = o_c;
  for (int n = 0; n < o_n; ++n) {
    for (int h = 0; h < o_h; ++h) {
      for (int w = 0; w < o_w; ++w) {
        float top_left_y_gw = 0.0;
        float top_left_x_gw = 0.0;
        const size_t grid_index = (size_t)n * o_h * o_w * 2 + (size_t)h * o_w + w;
        float y_real = (*(grid + grid_index + (size_t)o_h * o_w) + 1) * (i_h - 1) / 2;
        float x_real = (*(grid + grid_index) + 1) * (i_w - 1) / 2;
        float fy = floorf(y_real), fx = floorf(x_real);
        if (!(fy >= -16777216.0f)) fy = -16777216.0f;
        if (fy > 16777216.0f) fy = 16777216.0f;
        if (!(fx >= -16777216.0f)) fx = -16777216.0f;
        if (fx > 16777216.0f) fx = 16777216.0f;
        int top_left_y = (int)fy;
        int top_left_x = (int)fx;
        float top_left_y_w = 1.0 - (y_real - top_left_y);
        float top_left_x_w = 1.0 - (x_real - top_left_x);
        for (int c = 0; c < o_c; ++c) {
          const size_t grad_index = (((size_t)n * o_c + c) * o_h + h) * o_w + w;
          const float* plane = data + ((size_t)n * i_c + c) * i_h * i_w;
          float* gplane = g_input + ((size_t)n * i_c + c) * i_h * i_w;
          const long data_index = (long)top_left_y * i_w + top_left_x;
          float top_left_v = 0, top_right_v = 0, bottom_left_v = 0, bottom_right_v = 0;
          if (between(top_left_x, 0, i_w - 1) && between(top_left_y, 0, i_h - 1)) {
            *(gplane + data_index) += *(grad + grad_index) * top_left_y_w * top_left_x_w;
            top_left_v = *(plane + data_index);
          }
          if (between(top_left_x + 1, 0, i_w - 1) && between(top_left_y, 0, i_h - 1)) {
            *(gplane + data_index + 1) += *(grad + grad_index) * top_left_y_w * (1.0 - top_left_x_w);
            top_right_v = *(plane + data_index + 1);
          }
          if (between(top_left_x, 0, i_w - 1) && between(top_left_y + 1, 0, i_h - 1)) {
            *(gplane + data_index + i_w) += *(grad + grad_index) * (1.0 - top_left_y_w) * top_left_x_w;
            bottom_left_v = *(plane + data_index + i_w);
          }
          if (between(top_left_x + 1, 0, i_w - 1) && between(top_left_y + 1, 0, i_h - 1)) {
            *(gplane + data_index + i_w + 1) += *(grad + grad_index) * (1.0 - top_left_y_w) * (1.0 - top_left_x_w);
            bottom_right_v = *(plane + data_index + i_w + 1);
          }
          /* grad of the top-left weights; times -1 it is the grad of grid_src */
          top_left_y_gw -= *(grad + grad_index) * (top_right_v - bottom_right_v +
                           (top_left_v - top_right_v - bottom_left_v + bottom_right_v) * top_left_x_w);
          top_left_x_gw -= *(grad + grad_index) * (bottom_left_v - bottom_right_v +
                           (top_left_v - top_right_v - bottom_left_v + bottom_right_v) * top_left_y_w);
        }
        *(grad_grid + grid_index + (size_t)o_h * o_w) += top_left_y_gw * (i_h - 1) / 2;
        *(grad_grid + grid_index) += top_left_x_gw * (i_w - 1) / 2;
      }
    }
  }
}

/* GridGenerator backward, kWarp branch (grid_generator-inl.h): gdata = grad / ((dim-1)/2) */
EXPORT void lsfa_ref_grid_generator_warp_backward(const float* grad, float* gdata, int N, int H, int W) {
  const float half_w = (float)(((double)(float)W - 1.0) / 2.0);
  const float half_h = (float)(((double)(float)H - 1.0) / 2.0);
  for (int n = 0; n < N; ++n)
    for (int i = 0; i < H * W; ++i) {
      const size_t o = (size_t)n * 2 * H * W;
      gdata[o + i] = grad[o + i] / half_w;
      gdata[o + (size_t)H * W + i] = grad[o + (size_t)H * W + i] / half_h;
    }
}
