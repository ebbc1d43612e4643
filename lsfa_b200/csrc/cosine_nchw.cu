// Cosine logits of Fgfa_net from NCHW (channel-planar) embeddings, all-TMA form.
//   compute_weight SYM:111-116: l = sum_c l2n(a)_c * l2n(b)_c, L2Normalization(mode='channel', eps 1e-10);
//   Fgfa_net SYM:137-139: logits[n,0] = cos(emb_warp, emb_cur), logits[n,1] = cos(emb_cur, emb_cur).
// This is the largest stream of the cosine variant (2 x 2048 x HW floats per frame, twice the four feature
// tensors together), so it gets the same treatment as the fused kernel: a producer lane moves whole channel
// planes HBM -> smem with cp.async.bulk (each byte once, no sector over-fetch on the 8-byte-phase planes), 15
// consumer warps keep three running sums (a.a, b.b, a.b) per owned pixel in registers.
// Work = (frame, pair of channels), split statically and contiguously over the CTAs; a CTA writes ONE partial
// per frame it touches into a fixed slot (slot = its rank among the CTAs touching that frame), and a second tiny
// kernel adds the slots in order: the sums are deterministic (no float atomics).
#include "aggregate_nchw_tma.cuh"

namespace lsfa {

struct CosParams {
  const float* ew;
  const float* ec;
  float* part;        // [N][slots][3][HW]
  float* logits;      // [N][2][HW]
  int N, E, HW, chunks, slots, stages;
  unsigned pair_bytes, stage_bytes;
  long long items;
};

// CTA whose static range [items*b/grid, items*(b+1)/grid) contains item i
__host__ __device__ inline int cos_owner(long long i, long long items, int grid) {
  int g = (int)((i * grid) / items);
  while (g > 0 && items * g / grid > i) --g;
  while (g + 1 < grid && items * (g + 1) / grid <= i) ++g;
  return g;
}

template <int PPT>
__global__ void __launch_bounds__(kTmaThreads, 1) cosine_partials_tma_kernel(const __grid_constant__ CosParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* empty = full + kMaxStages;
  unsigned char* ring = smem_raw + kTmaHeaderBytes;
  const int tid = threadIdx.x, warp = tid >> 5;
  const long long i0 = P.items * (long long)blockIdx.x / gridDim.x, i1 = P.items * (long long)(blockIdx.x + 1) / gridDim.x;
  if (tid == 0) {
    for (int s = 0; s < P.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kTmaConsumerWarps);
    }
    fence_barrier_init();
  }
  __syncthreads();

  if (warp == kTmaConsumerWarps) {
    if ((tid & 31) == 0) {
      int s = 0;
      unsigned ph = 0;
      for (long long i = i0; i < i1; ++i) {
        if (i - i0 >= P.stages) mbar_wait(&empty[s], ph ^ 1u);     // the stage's previous occupant has been consumed
        unsigned char* st = ring + (size_t)s * P.stage_bytes;
        const size_t off = (size_t)i * 2 * P.HW;                   // item i = (n, pair): planes are contiguous over (n, e)
        mbar_expect_tx(&full[s], 2u * P.pair_bytes);
        bulk_g2s(st, P.ew + off, P.pair_bytes, &full[s]);
        bulk_g2s(st + P.stage_bytes / 2, P.ec + off, P.pair_bytes, &full[s]);
        if (++s == P.stages) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
    return;
  }

  float sww[PPT], scc[PPT], swc[PPT];
#pragma unroll
  for (int j = 0; j < PPT; ++j) sww[j] = scc[j] = swc[j] = 0.f;
  auto flush = [&](int n) {
    const int slot = (int)blockIdx.x - cos_owner((long long)n * P.chunks, P.items, (int)gridDim.x);
    float* dst = P.part + ((size_t)n * P.slots + slot) * 3 * P.HW;
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      const int p = tid + j * kTmaConsumers;
      if (p < P.HW) {
        dst[p] = sww[j];
        dst[P.HW + p] = scc[j];
        dst[2 * P.HW + p] = swc[j];
      }
      sww[j] = scc[j] = swc[j] = 0.f;
    }
  };
  int s = 0, cur_n = -1;
  unsigned ph = 0;
  for (long long i = i0; i < i1; ++i) {
    const int n = (int)(i / P.chunks);
    if (n != cur_n) {
      if (cur_n >= 0) flush(cur_n);
      cur_n = n;
    }
    mbar_wait(&full[s], ph);
    const float* a_s = reinterpret_cast<const float*>(ring + (size_t)s * P.stage_bytes);
    const float* b_s = reinterpret_cast<const float*>(ring + (size_t)s * P.stage_bytes + P.stage_bytes / 2);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      float a[PPT], b[PPT];
#pragma unroll
      for (int j = 0; j < PPT; ++j) {
        const int p = tid + j * kTmaConsumers;
        const bool ok = p < P.HW;
        a[j] = ok ? a_s[k * P.HW + p] : 0.f;
        b[j] = ok ? b_s[k * P.HW + p] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < PPT; ++j) {
        sww[j] = fmaf(a[j], a[j], sww[j]);
        scc[j] = fmaf(b[j], b[j], scc[j]);
        swc[j] = fmaf(a[j], b[j], swc[j]);
      }
    }
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&empty[s]);
    if (++s == P.stages) {
      s = 0;
      ph ^= 1u;
    }
  }
  if (cur_n >= 0) flush(cur_n);
}

// add the per-CTA partials of every frame in slot order, then the two cosines
__global__ void __launch_bounds__(256) cosine_finalize_kernel(const __grid_constant__ CosParams P, int grid_main) {
  const long long total = (long long)P.N * P.HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / P.HW), p = (int)(i - (long long)n * P.HW);
    const int first = cos_owner((long long)n * P.chunks, P.items, grid_main);
    const int last = cos_owner((long long)n * P.chunks + P.chunks - 1, P.items, grid_main);
    float tww = 0.f, tcc = 0.f, twc = 0.f;
    for (int j = 0; j <= last - first; ++j) {
      const float* src = P.part + ((size_t)n * P.slots + j) * 3 * P.HW;
      tww += src[p];
      tcc += src[P.HW + p];
      twc += src[2 * P.HW + p];
    }
    const float nw = sqrtf(tww + 1e-10f), nc = sqrtf(tcc + 1e-10f);
    P.logits[((size_t)n * 2 + 0) * P.HW + p] = twc / (nw * nc);
    P.logits[((size_t)n * 2 + 1) * P.HW + p] = tcc / (nc * nc);
  }
}

static int cos_sm_count() {
  int dev = 0, n = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n > 0 ? n : 148;
}

static bool plan_cosine_tma(const float* ew, const float* ec, int N, int E, int HW, CosParams& P, int* grid, size_t* smem,
                            int* ppt) {
  if (E % 2 || (2LL * HW) % 4 || HW > 9 * kTmaConsumers) return false;
  if ((reinterpret_cast<uintptr_t>(ew) % 16) || (reinterpret_cast<uintptr_t>(ec) % 16)) return false;
  P.ew = ew; P.ec = ec; P.N = N; P.E = E; P.HW = HW;
  P.chunks = E / 2;
  P.items = (long long)N * P.chunks;
  P.pair_bytes = (unsigned)(2u * HW * 4u);
  P.stage_bytes = 2u * ((P.pair_bytes + 127u) / 128u * 128u);
  long long st = (227 * 1024 - kTmaHeaderBytes) / P.stage_bytes;
  if (st > kMaxStages) st = kMaxStages;
  if (st < 2) return false;
  P.stages = (int)st;
  long long g = cos_sm_count();
  if (g > P.items) g = P.items;
  *grid = (int)g;
  int slots = 1;
  for (int n = 0; n < N; ++n) {
    const int c = cos_owner((long long)n * P.chunks + P.chunks - 1, P.items, (int)g) - cos_owner((long long)n * P.chunks, P.items, (int)g) + 1;
    if (c > slots) slots = c;
  }
  P.slots = slots;
  *smem = kTmaHeaderBytes + (size_t)P.stages * P.stage_bytes;
  *ppt = HW <= 5 * kTmaConsumers ? 5 : 9;
  return true;
}

// scratch the TMA form needs for its partial sums (0 = it cannot serve these arguments)
size_t cosine_tma_workspace_bytes(int N, int E, int HW) {
  CosParams P;
  int grid = 0, ppt = 0;
  size_t smem = 0;
  if (!plan_cosine_tma(reinterpret_cast<const float*>(16), reinterpret_cast<const float*>(16), N, E, HW, P, &grid, &smem, &ppt))
    return 0;
  // the grid depends on the SM count of the current device; size for the worst case (one slot per CTA of a 148+ SM part)
  const size_t slots = (size_t)P.slots > 4 ? (size_t)P.slots : 4;
  return (size_t)N * slots * 3 * HW * sizeof(float);
}

// returns cudaErrorNotSupported when the arguments / scratch do not fit: the caller then runs the LDG kernel
cudaError_t launch_cosine_logits_nchw_tma(const float* ew, const float* ec, float* logits, int N, int E, int HW,
                                          void* scratch, size_t scratch_bytes, cudaStream_t st) {
  CosParams P;
  int grid = 0, ppt = 0;
  size_t smem = 0;
  if (!scratch || (reinterpret_cast<uintptr_t>(scratch) % 16)) return cudaErrorNotSupported;
  if (!plan_cosine_tma(ew, ec, N, E, HW, P, &grid, &smem, &ppt)) return cudaErrorNotSupported;
  if (scratch_bytes < (size_t)N * P.slots * 3 * HW * sizeof(float)) return cudaErrorNotSupported;
  P.part = static_cast<float*>(scratch);
  P.logits = logits;
  cudaError_t e;
  if (ppt == 5) {
    e = cudaFuncSetAttribute(cosine_partials_tma_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cosine_partials_tma_kernel<5><<<grid, kTmaThreads, smem, st>>>(P);
  } else {
    e = cudaFuncSetAttribute(cosine_partials_tma_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cosine_partials_tma_kernel<9><<<grid, kTmaThreads, smem, st>>>(P);
  }
  e = cudaPeekAtLastError();
  if (e != cudaSuccess) return e;
  const long long total = (long long)N * HW;
  long long fg = (total + 255) / 256;
  if (fg > 148 * 8) fg = 148 * 8;
  cosine_finalize_kernel<<<(unsigned)fg, 256, 0, st>>>(P, grid);
  return cudaPeekAtLastError();
}

}  // namespace lsfa
