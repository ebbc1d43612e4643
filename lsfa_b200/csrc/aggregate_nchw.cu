// Fused warp x scale + aggregation, NCHW float32 (MXNet's layout: the drop-in variant).
//
// Replaces, in ONE pass over HBM, the operator chain of
//   SYM:571-572 GridGenerator(warp) + BilinearSampler      (a7, a8)
//   SYM:308/470/680  * scale_map                            (a9)
//   SYM:66,576  + rnet_conv0(res_diff)                      (a10)
//   SYM:236 / 104-108 / 144-147 / 315  aggregation          (a11-a14)
//   operator_py/choose_feat.py:23-31  per-frame select      (a15)
// and, when raw motion vectors are given, lib/utils/image.py:207-228 (a3,a5,a6).
//
// Kernel 1: agg_nchw_plane_kernel  ("plane-resident gather")
//   NCHW planes are small (38x63 fp32 = 9.6 KB), so K whole key planes are streamed into
//   shared memory with one TMA bulk copy (cp.async.bulk -> mbarrier) per stage of a ring,
//   and the 4-tap gather runs against shared memory: HBM sees every key byte exactly once
//   whatever the motion magnitude.  scale_map / cur / out are pure coalesced streams.
//   The per-pixel sampling record (4 weights + packed address + blend weights) lives in
//   REGISTERS: each thread owns the same <=8 pixels for every channel of a frame, so the
//   MV pooling, the exact fp32 grid round trip and the softmax run once per frame per CTA.
//   Persistent grid: one 512-thread CTA per SM walks a contiguous range of
//   (frame, channel-chunk, pixel-part) items.
// Kernel 2: agg_nchw_generic_kernel - any shape / alignment, gathers straight from global.
#include "lsfa_device.cuh"

namespace lsfa {

constexpr int kPlaneThreads = 512;
constexpr int kPlaneWarps = kPlaneThreads / 32;
constexpr int kMaxStages = 8;
constexpr int kBarrierBytes = 128;  // 2 * kMaxStages * 8

template <int K, int PPT>
__global__ void __launch_bounds__(kPlaneThreads, 1)
agg_nchw_plane_kernel(const __grid_constant__ AggParams P) {
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
  const bool has_scale = scale != nullptr;
  const bool has_cur = P.mode != LSFA_W_NONE;
  const bool has_res = P.res != nullptr;

  if (tid == 0) {
    for (int s = 0; s < P.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kPlaneWarps);
    }
    fence_barrier_init();
  }
  __syncthreads();

  // ---- producer state (meaningful in thread 0 only) ----
  long long pit = i0;  // next item whose key planes have not been requested yet
  int issued = 0;      // ring uses issued so far
  auto try_issue_one = [&]() {
    while (pit < i1 && P.bypass != nullptr) {  // bypass frames never touch the key feature
      const long long f = pit / items_per_frame;
      if (!__ldg(P.bypass + f)) break;
      pit = (f + 1) * items_per_frame;
    }
    if (pit >= i1) return;
    const int n = (int)(pit / items_per_frame);
    const int chunk = (int)((pit / P.parts) % P.chunks);
    const int kn = P.key_index ? __ldg(P.key_index + n) : n;
    const float* src = static_cast<const float*>(P.key) + ((size_t)kn * P.C + (size_t)chunk * K) * P.HWk;
    const int s = issued % P.stages;
    const int j = issued / P.stages;
    if (j >= 1) mbar_wait(&empty[s], (unsigned)(j - 1) & 1u);
    mbar_expect_tx(&full[s], P.stage_bytes);
    bulk_g2s(ring + (size_t)s * stage_floats, src, P.stage_bytes, &full[s]);
    ++issued;
    ++pit;
  };
  if (tid == 0)
    for (int s = 0; s < P.stages - 1; ++s) try_issue_one();

  // ---- consumer state ----
  Taps tp[PPT];
  float ww[PPT], wc[PPT];
  int cur_n = -1, cur_part = -1;
  int used = 0;

  for (long long it = i0; it < i1; ++it) {
    const int part = (int)(it % P.parts);
    const int chunk = (int)((it / P.parts) % P.chunks);
    const int n = (int)(it / items_per_frame);
    const int pix0 = part * P.part_pix;
    const int pend = min(P.HW, pix0 + P.part_pix);
    const bool byp = (P.bypass != nullptr) && (__ldg(P.bypass + n) != 0);

    if (n != cur_n || part != cur_part) {  // new frame (or pixel part): rebuild the records
      cur_n = n;
      cur_part = part;
#pragma unroll
      for (int j = 0; j < PPT; ++j) {
        const int p = pix0 + tid + j * kPlaneThreads;
        tp[j].w00 = tp[j].w01 = tp[j].w10 = tp[j].w11 = 0.0f;
        tp[j].packed = 0u;
        ww[j] = wc[j] = 0.0f;
        if (p < pend && !byp) {
          const int y = p / P.W, x = p - y * P.W;
          float gx, gy;
          pixel_grid(P, n, y, x, gx, gy);
          tp[j] = make_taps(gx, gy, P.Hk, P.Wk, P.wk_m1, P.hk_m1);
          pixel_weights(P, n, p, ww[j], wc[j]);
          if (has_res) {
#pragma unroll
            for (int k = 0; k < 3; ++k)
              res_s[k * P.part_pix + (p - pix0)] = __ldg(P.res + ((size_t)n * 3 + k) * P.HW + p);
          }
        }
      }
    }

    const int c0 = chunk * K;
    const size_t fbase = ((size_t)n * P.C + c0) * P.HW;

    // once-touched streams first: they are in flight while we wait for the key planes
    float sc[K][PPT], cu[K][PPT];
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
      for (int j = 0; j < PPT; ++j) {
        const int p = pix0 + tid + j * kPlaneThreads;
        sc[k][j] = 1.0f;
        cu[k][j] = 0.0f;
        if (p < pend) {
          const size_t e = fbase + (size_t)k * P.HW + p;
          if (has_scale && !byp) sc[k][j] = ldg_stream(scale + e);
          if (has_cur) cu[k][j] = ldg_stream(cur + e);
        }
      }
    }

    if (byp) {  // ChooseFeat: keep the current frame's own feature
#pragma unroll
      for (int k = 0; k < K; ++k) {
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
          const int p = pix0 + tid + j * kPlaneThreads;
          if (p < pend) {
            float* dst = out + fbase + (size_t)k * P.HW + p;
            stg_stream(dst, P.req_add ? cu[k][j] + *dst : cu[k][j]);
          }
        }
      }
      continue;
    }

    if (tid == 0) try_issue_one();  // refills the stage consumed one item ago

    const int s = used % P.stages;
    mbar_wait(&full[s], (unsigned)(used / P.stages) & 1u);
    const float* __restrict__ buf = ring + (size_t)s * stage_floats;

#pragma unroll
    for (int k = 0; k < K; ++k) {
      float rw0 = 0.f, rw1 = 0.f, rw2 = 0.f, rb = 0.f;
      if (has_res) {
        rw0 = __ldg(P.rnet_w + (size_t)(c0 + k) * 3 + 0);
        rw1 = __ldg(P.rnet_w + (size_t)(c0 + k) * 3 + 1);
        rw2 = __ldg(P.rnet_w + (size_t)(c0 + k) * 3 + 2);
        rb = __ldg(P.rnet_b + c0 + k);
      }
      const float* __restrict__ plane = buf + (size_t)k * P.HWk;
#pragma unroll
      for (int j = 0; j < PPT; ++j) {
        const int p = pix0 + tid + j * kPlaneThreads;
        if (p < pend) {
          const unsigned a = tp[j].packed & 0xffffffu;
          const unsigned dx = (tp[j].packed >> 24) & 1u;
          const unsigned dy = ((tp[j].packed >> 25) & 1u) ? (unsigned)P.Wk : 0u;
          float v = tp[j].w00 * plane[a];
          v = fmaf(tp[j].w01, plane[a + dx], v);
          v = fmaf(tp[j].w10, plane[a + dy], v);
          v = fmaf(tp[j].w11, plane[a + dy + dx], v);
          if (has_scale) v *= sc[k][j];
          if (has_res) {
            const int q = p - pix0;
            float r = rw0 * res_s[q];
            r = fmaf(rw1, res_s[P.part_pix + q], r);
            r = fmaf(rw2, res_s[2 * P.part_pix + q], r);
            v += r + rb;
          }
          float o;
          if (P.mode == LSFA_W_NONE) o = v;
          else if (P.mode == LSFA_W_ADD) o = cu[k][j] + v;
          else if (P.mode == LSFA_W_MEAN) o = 0.5f * (v + cu[k][j]);
          else o = fmaf(wc[j], cu[k][j], ww[j] * v);
          float* dst = out + fbase + (size_t)k * P.HW + p;
          if (P.req_add) o += *dst;   // kAddTo is the rare path: read late, no register buffer
          stg_stream(dst, o);
        }
      }
    }
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&empty[s]);
    ++used;
  }
}

// ---------------------------------------------------------------------------------------
// Generic kernel: one thread per output pixel, CG channels per thread, taps gathered from
// global memory (L1/L2 serve the 4-tap reuse).  No alignment or size restriction.
// ---------------------------------------------------------------------------------------
constexpr int kGenericThreads = 256;
constexpr int kGenericCG = 16;

__global__ void __launch_bounds__(kGenericThreads)
agg_nchw_generic_kernel(const __grid_constant__ AggParams P) {
  const int n = blockIdx.z;
  const int p = blockIdx.x * kGenericThreads + threadIdx.x;
  if (p >= P.HW) return;
  const int c_begin = blockIdx.y * kGenericCG;
  const int c_end = min(P.C, c_begin + kGenericCG);
  const float* __restrict__ scale = static_cast<const float*>(P.scale);
  const float* __restrict__ cur = static_cast<const float*>(P.cur);
  float* __restrict__ out = static_cast<float*>(P.out);
  const bool has_cur = P.mode != LSFA_W_NONE;
  const bool byp = (P.bypass != nullptr) && (__ldg(P.bypass + n) != 0);

  if (byp) {
    for (int c = c_begin; c < c_end; ++c) {
      const size_t e = ((size_t)n * P.C + c) * P.HW + p;
      float o = cur[e];
      if (P.req_add) o += out[e];
      out[e] = o;
    }
    return;
  }
  const int y = p / P.W, x = p - y * P.W;
  float gx, gy;
  pixel_grid(P, n, y, x, gx, gy);
  const Taps t = make_taps(gx, gy, P.Hk, P.Wk, P.wk_m1, P.hk_m1);
  float ww, wc;
  pixel_weights(P, n, p, ww, wc);
  float r0 = 0.f, r1 = 0.f, r2 = 0.f;
  if (P.res) {
    r0 = __ldg(P.res + ((size_t)n * 3 + 0) * P.HW + p);
    r1 = __ldg(P.res + ((size_t)n * 3 + 1) * P.HW + p);
    r2 = __ldg(P.res + ((size_t)n * 3 + 2) * P.HW + p);
  }
  const int kn = P.key_index ? __ldg(P.key_index + n) : n;
  const unsigned a = t.packed & 0xffffffu;
  const unsigned dx = (t.packed >> 24) & 1u;
  const unsigned dy = ((t.packed >> 25) & 1u) ? (unsigned)P.Wk : 0u;
  for (int c = c_begin; c < c_end; ++c) {
    const float* __restrict__ plane = static_cast<const float*>(P.key) + ((size_t)kn * P.C + c) * P.HWk;
    const size_t e = ((size_t)n * P.C + c) * P.HW + p;
    float v = t.w00 * __ldg(plane + a);
    v = fmaf(t.w01, __ldg(plane + a + dx), v);
    v = fmaf(t.w10, __ldg(plane + a + dy), v);
    v = fmaf(t.w11, __ldg(plane + a + dy + dx), v);
    if (scale) v *= __ldg(scale + e);
    if (P.res) {
      float r = __ldg(P.rnet_w + (size_t)c * 3) * r0;
      r = fmaf(__ldg(P.rnet_w + (size_t)c * 3 + 1), r1, r);
      r = fmaf(__ldg(P.rnet_w + (size_t)c * 3 + 2), r2, r);
      v += r + __ldg(P.rnet_b + c);
    }
    const float cv = has_cur ? __ldg(cur + e) : 0.0f;
    float o;
    if (P.mode == LSFA_W_NONE) o = v;
    else if (P.mode == LSFA_W_ADD) o = cv + v;
    else if (P.mode == LSFA_W_MEAN) o = 0.5f * (v + cv);
    else o = fmaf(wc, cv, ww * v);
    if (P.req_add) o += out[e];
    out[e] = o;
  }
}

// ---------------------------------------------------------------------------------------
// Cosine logits from NCHW embeddings (Fgfa_net compute_weight, SYM:111-116,137-139).
// Block = 32 pixels x 8 channel groups; each group walks E/8 channels with coalesced
// 128-byte rows, partial sums are combined in a fixed order in shared memory.
// ---------------------------------------------------------------------------------------
constexpr int kCosGroups = 8;

__global__ void __launch_bounds__(32 * kCosGroups)
cosine_logits_nchw_kernel(const float* __restrict__ ew, const float* __restrict__ ec,
                          float* __restrict__ logits, int E, int HW) {
  __shared__ float part[3][kCosGroups][32];
  const int n = blockIdx.y;
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int p = blockIdx.x * 32 + lane;
  float sww = 0.f, scc = 0.f, swc = 0.f;
  if (p < HW) {
    const size_t base = (size_t)n * E * HW + p;
    const int per = (E + kCosGroups - 1) / kCosGroups;
    const int e0 = grp * per, e1 = min(E, e0 + per);
    int e = e0;
    for (; e + 4 <= e1; e += 4) {
      float a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a[u] = ldg_stream(ew + base + (size_t)(e + u) * HW);
        b[u] = ldg_stream(ec + base + (size_t)(e + u) * HW);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        sww = fmaf(a[u], a[u], sww);
        scc = fmaf(b[u], b[u], scc);
        swc = fmaf(a[u], b[u], swc);
      }
    }
    for (; e < e1; ++e) {
      const float a = ldg_stream(ew + base + (size_t)e * HW);
      const float b = ldg_stream(ec + base + (size_t)e * HW);
      sww = fmaf(a, a, sww);
      scc = fmaf(b, b, scc);
      swc = fmaf(a, b, swc);
    }
  }
  part[0][grp][lane] = sww;
  part[1][grp][lane] = scc;
  part[2][grp][lane] = swc;
  __syncthreads();
  if (grp == 0 && p < HW) {
    float tww = 0.f, tcc = 0.f, twc = 0.f;
#pragma unroll
    for (int g = 0; g < kCosGroups; ++g) {
      tww += part[0][g][lane];
      tcc += part[1][g][lane];
      twc += part[2][g][lane];
    }
    const float nw = sqrtf(tww + 1e-10f), nc = sqrtf(tcc + 1e-10f);
    logits[((size_t)n * 2 + 0) * HW + p] = twc / (nw * nc);
    logits[((size_t)n * 2 + 1) * HW + p] = tcc / (nc * nc);
  }
}

// ---------------------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------------------
static int g_sm_count_cache[64];  // written once per device with the same value: benign

static int sm_count() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev >= 0 && dev < 64 && g_sm_count_cache[dev] > 0) return g_sm_count_cache[dev];
  int n = 148;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (dev >= 0 && dev < 64) g_sm_count_cache[dev] = n;
  return n;
}

template <int K, int PPT>
static cudaError_t launch_plane(const AggParams& P, size_t smem, cudaStream_t st) {
  auto kfn = agg_nchw_plane_kernel<K, PPT>;
  cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  long long grid = sm_count();
  if (grid > P.items) grid = P.items;
  kfn<<<(unsigned)grid, kPlaneThreads, smem, st>>>(P);
  return cudaPeekAtLastError();
}

// Decide whether the plane-resident kernel can serve these args; fill the tiling fields.
bool plan_plane_kernel(AggParams& P, size_t* smem_out) {
  const size_t kSmemBudget = 220 * 1024;
  if (P.HW > 8 * kPlaneThreads * 64) return false;
  if (P.HWk >= (1 << 24)) return false;
  int K = 0;
  const int prefer[3] = {2, 4, 1};
  for (int i = 0; i < 3 && K == 0; ++i) {
    const int k = prefer[i];
    if (P.C % k) continue;
    if (((long long)k * P.HWk) % 4) continue;          // bulk copy size multiple of 16 B
    if ((size_t)k * P.HWk * 4 > 96 * 1024) continue;   // keep >= 2 stages
    K = k;
  }
  if (K == 0) return false;
  if (reinterpret_cast<uintptr_t>(P.key) % 16) return false;
  if (((long long)P.C * P.HWk) % 4) return false;      // every frame's planes stay 16 B aligned
  int ppt = 8;
  const int ppt_options[6] = {1, 2, 4, 5, 6, 8};
  for (int i = 5; i >= 0; --i)
    if ((long long)ppt_options[i] * kPlaneThreads >= P.HW) ppt = ppt_options[i];
  P.K = K;
  P.chunks = P.C / K;
  P.part_pix = ppt * kPlaneThreads;
  P.parts = (P.HW + P.part_pix - 1) / P.part_pix;
  P.stage_bytes = (unsigned)((size_t)K * P.HWk * 4);
  const size_t res_bytes = P.res ? (size_t)3 * P.part_pix * 4 : 0;
  long long stages = ((long long)kSmemBudget - kBarrierBytes - (long long)res_bytes) / P.stage_bytes;
  if (stages < 2) return false;
  if (stages > kMaxStages) stages = kMaxStages;
  P.stages = (int)stages;
  P.items = (long long)P.N * P.chunks * P.parts;
  *smem_out = kBarrierBytes + (size_t)P.stages * P.stage_bytes + res_bytes;
  return true;
}

cudaError_t launch_agg_nchw_plane(const AggParams& P, size_t smem, cudaStream_t st) {
  const int ppt = P.part_pix / kPlaneThreads;
#define LSFA_PLANE_CASE(KK, PP) \
  if (P.K == KK && ppt == PP) return launch_plane<KK, PP>(P, smem, st);
  LSFA_PLANE_CASE(1, 1) LSFA_PLANE_CASE(1, 2) LSFA_PLANE_CASE(1, 4) LSFA_PLANE_CASE(1, 5) LSFA_PLANE_CASE(1, 6) LSFA_PLANE_CASE(1, 8)
  LSFA_PLANE_CASE(2, 1) LSFA_PLANE_CASE(2, 2) LSFA_PLANE_CASE(2, 4) LSFA_PLANE_CASE(2, 5) LSFA_PLANE_CASE(2, 6) LSFA_PLANE_CASE(2, 8)
  LSFA_PLANE_CASE(4, 1) LSFA_PLANE_CASE(4, 2) LSFA_PLANE_CASE(4, 4) LSFA_PLANE_CASE(4, 5) LSFA_PLANE_CASE(4, 6) LSFA_PLANE_CASE(4, 8)
#undef LSFA_PLANE_CASE
  return cudaErrorInvalidValue;
}

cudaError_t launch_agg_nchw_generic(const AggParams& P, cudaStream_t st) {
  dim3 grid((P.HW + kGenericThreads - 1) / kGenericThreads, (P.C + kGenericCG - 1) / kGenericCG, P.N);
  if (grid.y > 65535 || grid.z > 65535) return cudaErrorInvalidValue;
  agg_nchw_generic_kernel<<<grid, kGenericThreads, 0, st>>>(P);
  return cudaPeekAtLastError();
}

cudaError_t launch_cosine_logits_nchw(const float* ew, const float* ec, float* logits, int N,
                                      int E, int HW, cudaStream_t st) {
  dim3 grid((HW + 31) / 32, N);
  if (grid.y > 65535) return cudaErrorInvalidValue;
  cosine_logits_nchw_kernel<<<grid, 32 * kCosGroups, 0, st>>>(ew, ec, logits, E, HW);
  return cudaPeekAtLastError();
}

}  // namespace lsfa
