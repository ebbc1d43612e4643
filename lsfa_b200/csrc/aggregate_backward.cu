// aggregate_backward.cu - backward of the fused operator's tails (get_train_symbol SYM:306-338):
//
//   out = wc * cur + ww * ( BilinearSampler(key, GridGenerator(flow)) * scale_map + rnet_conv0(res) )
//   (ww, wc) = (1,0) | (1,1) | (.5,.5) | softmax(logit_warp, logit_cur);   bypass frames: out = cur
//
// Given d/d(out) this file produces, in ONE pass over the feature streams,
//   d/d(cur)        = wc * g                                      (SYM:104-108 / 236 / 315)
//   d/d(scale_map)  = ww * g * warp                               (SYM:308: flow_warp * scale_map)
//   gw              = ww * g * scale_map = d/d(warped feature)    -> fed to the a7+a8 backward (sampler_backward.cu),
//                                                                    which yields d/d(key) and d/d(flow)
//   per-pixel sums over channels  T1 = sum_c g * ww*src0,  T2 = sum_c g * cur   -> d/d(logits) through the 2-way softmax
//   d/d(res)[j]     = sum_c ww * g * W[c][j]                      (SYM:66 rnet_conv0, 1x1 conv 3 -> C with bias)
//   d/d(W)[c][j]    = sum_{n,p} ww * g * res[j],   d/d(b)[c] = sum_{n,p} ww * g
// Every reduction is written as ordered partial sums and added by a finalise kernel in a fixed order: deterministic.
// The sampling weights / tap offsets are the forward's own records (agg_records_kernel), so the backward is the exact
// transpose of this library's forward.  NCHW float32.
#include "lsfa_device.cuh"
#include "aggregate_backward.h"

namespace lsfa {

cudaError_t launch_agg_records(const AggParams& P, uint4* rec, cudaStream_t st);   // aggregate_nchw.cu
bool plan_tma_kernel(AggParams& P, size_t* smem_out);
cudaError_t launch_agg_nchw_tma(const AggParams& Pin, size_t smem, cudaStream_t st);

constexpr int kTailThreads = 128;   // pixels per block
constexpr int kTailCH = 32;         // channels per block

__device__ __forceinline__ void put(float* dst, float v, int add) { *dst = add ? *dst + v : v; }

__global__ void __launch_bounds__(kTailThreads) agg_tail_backward_kernel(const __grid_constant__ TailBwdParams Q) {
  const AggParams& P = Q.P;
  const int n = blockIdx.z, chunk = blockIdx.y, tile = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int p = tile * kTailThreads + tid;
  const bool active = p < P.HW;
  const bool byp = P.bypass != nullptr && __ldg(P.bypass + n) != 0;
  const bool has_scale = P.scale != nullptr, has_res = P.res != nullptr, has_cur = P.mode != LSFA_W_NONE;
  // rnet parameter gradients: d/d(W)[c][j] = sum_p ww*g[p,c] * res[j][p], d/d(b)[c] = sum_p ww*g[p,c] - a (32 x 128) by
  // (128 x 4) product per block.  ww*g goes to a padded shared tile and the block multiplies once at the end (a warp
  // shuffle tree per channel and quantity was 20 SHFL + 20 FADD per channel and thread: half of this kernel's instructions)
  // (dynamic shared memory, allocated only when these gradients are wanted: 18 KB per block would cost the other variants
  // a fifth of their resident warps)
  extern __shared__ float tail_smem[];
  float (*wgT)[kTailThreads + 1] = reinterpret_cast<float (*)[kTailThreads + 1]>(tail_smem);
  float (*rS)[kTailThreads] = reinterpret_cast<float (*)[kTailThreads]>(tail_smem + kTailCH * (kTailThreads + 1));

  float w00 = 0.f, w01 = 0.f, w10 = 0.f, w11 = 0.f, wc = 0.f, ww = 0.f;
  unsigned o_top = 0u, o_bot = 0u;
  if (active && !byp) {
    const uint4* rp = P.records + 2 * ((size_t)n * P.HW + p);
    const uint4 ra = __ldg(rp), rb = __ldg(rp + 1);
    w00 = __uint_as_float(ra.x); w01 = __uint_as_float(ra.y); w10 = __uint_as_float(ra.z); w11 = __uint_as_float(ra.w);
    wc = __uint_as_float(rb.x); ww = __uint_as_float(rb.y);
    o_top = rb.z; o_bot = rb.w;
  }
  float r0 = 0.f, r1 = 0.f, r2 = 0.f;
  if (has_res && active && !byp) {
    r0 = __ldg(P.res + ((size_t)n * 3 + 0) * P.HW + p);
    r1 = __ldg(P.res + ((size_t)n * 3 + 1) * P.HW + p);
    r2 = __ldg(P.res + ((size_t)n * 3 + 2) * P.HW + p);
  }
  const int kn = key_slot(P, n);
  if (Q.partRnet) { rS[0][tid] = r0; rS[1][tid] = r1; rS[2][tid] = r2; }
  float T1 = 0.f, T2 = 0.f, gr0 = 0.f, gr1 = 0.f, gr2 = 0.f;
  const int c_begin = chunk * kTailCH, c_end = min(P.C, c_begin + kTailCH);
  // four channels per step: all streaming loads of the step are issued before its stores (the outputs may alias the
  // ww*warp buffer, so the compiler will not hoist them itself): 16 independent loads in flight per thread
  constexpr int U = 4;
  for (int c0 = c_begin; c0 < c_end; c0 += U) {
    float gv[U], vv[U], sv[U], cv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int c = c0 + u;
      const bool ok = active && c < c_end;
      const size_t e = ((size_t)n * P.C + min(c, P.C - 1)) * P.HW + (active ? p : 0);
      gv[u] = ok ? __ldg(Q.og + e) : 0.f;
      vv[u] = (ok && !byp && Q.vf != nullptr) ? Q.vf[e] : 0.f;
      sv[u] = (ok && !byp && has_scale) ? __ldg(static_cast<const float*>(P.scale) + e) : 1.0f;
      cv[u] = (ok && !byp && Q.partT && has_cur) ? __ldg(static_cast<const float*>(P.cur) + e) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int c = c0 + u;
      if (c >= c_end) break;                                 // uniform
      const size_t e = ((size_t)n * P.C + c) * P.HW + p;
      const float g = gv[u];
      float a3 = 0.f;                                      // ww * g: this thread's share of the rnet parameter gradients
      if (active) {
        if (byp) {          // ChooseFeat kept conv_feat: the whole gradient goes to cur
          if (Q.gcur) put(Q.gcur + e, g, Q.add_cur);
          if (Q.gw) Q.gw[e] = 0.f;
          if (Q.gscale) put(Q.gscale + e, 0.f, Q.add_scale);
        } else {
          float vf = vv[u];                                   // = ww * warp (the forward's folded chain), from the all-TMA pass
          if (Q.vf == nullptr) {                              // shapes that pass does not serve: gather here
            const unsigned char* plane = reinterpret_cast<const unsigned char*>(static_cast<const float*>(P.key) + ((size_t)kn * P.C + c) * P.HWk);
            const float v00 = __ldg(reinterpret_cast<const float*>(plane + (o_top & 0xffffu)));
            const float v01 = __ldg(reinterpret_cast<const float*>(plane + (o_top >> 16)));
            const float v10 = __ldg(reinterpret_cast<const float*>(plane + (o_bot & 0xffffu)));
            const float v11 = __ldg(reinterpret_cast<const float*>(plane + (o_bot >> 16)));
            vf = w00 * v00;
            vf = fmaf(w01, v01, vf);
            vf = fmaf(w10, v10, vf);
            vf = fmaf(w11, v11, vf);
          }
          const float sc = sv[u];
          float src0f = vf * sc;
          const float wg = ww * g;
          if (has_res) {
            const float rw0 = __ldg(P.rnet_w + (size_t)c * 3), rw1 = __ldg(P.rnet_w + (size_t)c * 3 + 1), rw2 = __ldg(P.rnet_w + (size_t)c * 3 + 2);
            src0f = fmaf(ww, rnet_term(rw0, rw1, rw2, __ldg(P.rnet_b + c), r0, r1, r2), src0f);
            gr0 = fmaf(wg, rw0, gr0);
            gr1 = fmaf(wg, rw1, gr1);
            gr2 = fmaf(wg, rw2, gr2);
            a3 = wg;
          }
          if (Q.gscale) put(Q.gscale + e, g * vf, Q.add_scale);
          if (Q.gw) Q.gw[e] = wg * sc;
          if (Q.gcur) put(Q.gcur + e, wc * g, Q.add_cur);
          T1 = fmaf(g, src0f, T1);
          if (Q.partT && has_cur) T2 = fmaf(g, cv[u], T2);
        }
      }
      if (Q.partRnet) wgT[c - c_begin][tid] = a3;          // ww * g of this pixel and channel (0 for bypass / inactive)
    }
  }
  if (active) {
    if (Q.partT) {
      float* t = Q.partT + (((size_t)chunk * P.N + n) * 2) * P.HW + p;
      t[0] = T1;
      t[P.HW] = T2;
    }
    if (Q.partRes) {
      float* t = Q.partRes + (((size_t)chunk * P.N + n) * 3) * P.HW + p;
      t[0] = gr0; t[P.HW] = gr1; t[2 * (size_t)P.HW] = gr2;
    }
  }
  if (Q.partRnet) {
    __syncthreads();
    for (int i = tid; i < 4 * (c_end - c_begin); i += kTailThreads) {
      const int cc = i >> 2, q = i & 3;
      float s = 0.f;
      if (q < 3) {
        for (int pp = 0; pp < kTailThreads; ++pp) s = fmaf(wgT[cc][pp], rS[q][pp], s);
      } else {
        for (int pp = 0; pp < kTailThreads; ++pp) s += wgT[cc][pp];
      }
      Q.partRnet[(((size_t)n * Q.tiles + tile) * P.C + (c_begin + cc)) * 4 + q] = s;
    }
  }
}

// d/d(logits) from the channel-chunk partial sums (fixed order), through the 2-way softmax:
//   d l_warp = wc * sum_c g*(ww*src0) - ww*wc * sum_c g*cur,   d l_cur = - d l_warp
__global__ void tail_logits_finalize_kernel(const float* __restrict__ partT, const float* __restrict__ logits,
                                            const unsigned char* __restrict__ bypass, float* __restrict__ glogits, int chunks, int N,
                                            int HW, int add) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)N * HW) return;
  const int n = (int)(i / HW), p = (int)(i % HW);
  float T1 = 0.f, T2 = 0.f;
  for (int c = 0; c < chunks; ++c) {
    const float* t = partT + (((size_t)c * N + n) * 2) * HW + p;
    T1 += t[0];
    T2 += t[HW];
  }
  float ww, wc;
  softmax2(logits[((size_t)n * 2) * HW + p], logits[((size_t)n * 2 + 1) * HW + p], ww, wc);
  float d = wc * T1 - ww * wc * T2;
  if (bypass != nullptr && bypass[n] != 0) d = 0.f;
  put(glogits + ((size_t)n * 2) * HW + p, d, add);
  put(glogits + ((size_t)n * 2 + 1) * HW + p, -d, add);
}

__global__ void tail_res_finalize_kernel(const float* __restrict__ partRes, float* __restrict__ gres, int chunks, int N, int HW, int add) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;       // over N*3*HW
  if (i >= (size_t)N * 3 * HW) return;
  float s = 0.f;
  for (int c = 0; c < chunks; ++c) s += partRes[(size_t)c * N * 3 * HW + i];
  put(gres + i, s, add);
}

__global__ void tail_rnet_finalize_kernel(const float* __restrict__ partRnet, float* __restrict__ gw, float* __restrict__ gb, int slots,
                                          int C, int add) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;                  // over C*4
  if (i >= C * 4) return;
  float s = 0.f;
  for (int k = 0; k < slots; ++k) s += partRnet[(size_t)k * C * 4 + i];
  const int c = i / 4, q = i % 4;
  if (q < 3) put(gw + c * 3 + q, s, add);
  else put(gb + c, s, add);
}

size_t tail_backward_workspace_bytes(int N, int C, int HW, bool want_gw, bool want_logits, bool want_res) {
  const size_t chunks = (size_t)(C + kTailCH - 1) / kTailCH, tiles = (size_t)(HW + kTailThreads - 1) / kTailThreads;
  auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
  size_t b = up((size_t)N * HW * 32);                                   // records
  b += up((size_t)N * C * HW * 4);                                      // ww * warp of the forward pass, then d/d(warp) in place
  b += up((size_t)N * 16 * 4 + 64);                                     // work-claim counters of that pass
  (void)want_gw;
  if (want_logits) b += up(chunks * N * 2 * HW * 4);
  if (want_res) b += up(chunks * N * 3 * HW * 4) + up((size_t)N * tiles * C * 16);
  return b;
}

cudaError_t launch_tail_backward(AggParams P, const TailBwdRequest& R, void* workspace, cudaStream_t st) {
  const size_t chunks = (size_t)(P.C + kTailCH - 1) / kTailCH, tiles = (size_t)(P.HW + kTailThreads - 1) / kTailThreads;
  auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
  char* ws = static_cast<char*>(workspace);
  uint4* rec = reinterpret_cast<uint4*>(ws);
  ws += up((size_t)P.N * P.HW * 32);
  TailBwdParams Q{};
  Q.og = R.out_grad;
  Q.gscale = R.grad_scale; Q.add_scale = R.add_scale;
  Q.gcur = R.grad_cur; Q.add_cur = R.add_cur;
  Q.tiles = (int)tiles;
  float* vfbuf = reinterpret_cast<float*>(ws);
  ws += up((size_t)P.N * P.C * P.HW * 4);
  unsigned* sched = reinterpret_cast<unsigned*>(ws);
  ws += up((size_t)P.N * 16 * 4 + 64);
  if (R.want_gw) Q.gw = vfbuf;
  if (R.grad_logits) { Q.partT = reinterpret_cast<float*>(ws); ws += up(chunks * P.N * 2 * P.HW * 4); }
  if (R.grad_res || R.grad_rnet_w) {
    Q.partRes = reinterpret_cast<float*>(ws); ws += up(chunks * P.N * 3 * P.HW * 4);
    Q.partRnet = reinterpret_cast<float*>(ws); ws += up((size_t)P.N * tiles * P.C * 16);
  }
  // the forward's sampling records (index math of a3-a8 + blend weights), once per output pixel
  P.records = nullptr; P.rowrange = nullptr; P.sched = nullptr;
  P.parts = 1; P.part_pix = P.HW;
  // ww * warp by a warp-only pass of the all-TMA forward kernel over the SAME (folded) records, written where d/d(warp)
  // will go: the tail kernel is then a pure streaming kernel (its own gather from global memory was 4 scattered loads per
  // element: 1.7 ms per 64 frames; 0.24 ms + a streaming pass this way).  Shapes that kernel does not serve gather here.
  AggParams F = P;
  F.scale = nullptr; F.cur = nullptr; F.res = nullptr; F.rnet_w = nullptr; F.rnet_b = nullptr;
  // F.bypass stays: bypass frames have no records (the pre-pass skips them) and the tail ignores their ww * warp
  F.mode = LSFA_W_NONE; F.req_add = 0; F.out = vfbuf; F.logits = nullptr; F.emb_warp = nullptr; F.emb_cur = nullptr;
  size_t fsmem = 0;
  const bool staged = plan_tma_kernel(F, &fsmem) && F.parts == 1;
  P.canon = staged ? 1 : 0;
  cudaError_t e = launch_agg_records(P, rec, st);
  if (e != cudaSuccess) return e;
  P.records = rec;
  if (staged) {
    F.records = rec;
    F.records_ready = 1;
    F.sched = sched;
    if ((e = launch_agg_nchw_tma(F, fsmem, st)) != cudaSuccess) return e;
    Q.vf = vfbuf;
  }
  Q.P = P;
  dim3 grid((unsigned)tiles, (unsigned)chunks, (unsigned)P.N);
  const size_t tail_smem_bytes = Q.partRnet ? (size_t)(kTailCH * (kTailThreads + 1) + 3 * kTailThreads) * sizeof(float) : 0;
  agg_tail_backward_kernel<<<grid, kTailThreads, tail_smem_bytes, st>>>(Q);
  if ((e = cudaPeekAtLastError()) != cudaSuccess) return e;
  const size_t NP = (size_t)P.N * P.HW;
  if (R.grad_logits)
    tail_logits_finalize_kernel<<<(unsigned)((NP + 255) / 256), 256, 0, st>>>(Q.partT, P.logits, P.bypass, R.grad_logits, (int)chunks, P.N,
                                                                             P.HW, R.add_logits);
  if (R.grad_res)
    tail_res_finalize_kernel<<<(unsigned)((NP * 3 + 255) / 256), 256, 0, st>>>(Q.partRes, R.grad_res, (int)chunks, P.N, P.HW, R.add_res);
  if (R.grad_rnet_w)
    tail_rnet_finalize_kernel<<<(unsigned)((P.C * 4 + 255) / 256), 256, 0, st>>>(Q.partRnet, R.grad_rnet_w, R.grad_rnet_b,
                                                                                (int)((size_t)P.N * tiles), P.C, R.add_rnet);
  *R.gw_out = Q.gw;
  return cudaPeekAtLastError();
}

}  // namespace lsfa
