// Window-resident channels-last kernel (aggregate_nhwc_win.cuh): tensor maps, planning and launch.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "aggregate_nhwc_win.cuh"

namespace lsfa {

static PFN_cuTensorMapEncodeTiled_v12000 win_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }();
  return fn;
}

// (C, W, H, frames) view of a channels-last tensor with a (chunk, bw, bh, 1) box; elements outside the tensor read as zero
static bool win_map(CUtensorMap* m, const void* base, bool bf16, int C, int W, int H, long long frames, int bw, int bh) {
  auto enc = win_encode_fn();
  if (!enc || base == nullptr) return false;
  const cuuint64_t es = bf16 ? 2 : 4;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)frames};
  cuuint64_t strides[3] = {(cuuint64_t)C * es, (cuuint64_t)W * C * es, (cuuint64_t)H * W * C * es};
  cuuint32_t box[4] = {(cuuint32_t)(kWinChunk / es), (cuuint32_t)bw, (cuuint32_t)bh, 1};
  cuuint32_t one[4] = {1, 1, 1, 1};
  return enc(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), dims, strides,
             box, one, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <typename T, int VAR>
static cudaError_t win_launch(const AggParams& P, const WinPlan& Q, int grid, const CUtensorMap& mk, const CUtensorMap& mk2,
                              const CUtensorMap& ms, const CUtensorMap& mc, cudaStream_t st) {
  auto kfn = agg_nhwc_win_kernel<T, VAR>;
  static bool attr_done[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  kfn<<<(unsigned)grid, kWinThreads, Q.smem, st>>>(P, Q, mk, mk2, ms, mc);
  return cudaPeekAtLastError();
}

// cudaErrorNotSupported: the arguments are outside what this kernel serves (the caller falls back)
cudaError_t launch_agg_nhwc_win(const AggParams& P_in, bool bf16, int var, cudaStream_t st) {
  AggParams P = P_in;
  WinPlan Q;
  if (!plan_nhwc_win(P, bf16, var, &Q)) return cudaErrorNotSupported;
  const void* ptrs[4] = {P.key, P.scale, P.cur, P.out};
  for (const void* q : ptrs)
    if (q && (reinterpret_cast<uintptr_t>(q) % 16)) return cudaErrorNotSupported;
  const bool has_scale = var == kVarScale || var == kVarScaleCur;
  const bool has_cur = var == kVarScaleCur || var == kVarResCur;
  const long long keys = P.key_index ? P.num_keys : P.N;
  CUtensorMap mk, mk2, ms, mc;
  if (!win_map(&mk, P.key, bf16, P.C, P.Wk, P.Hk, keys, kWinB, kWinB)) return cudaErrorNotSupported;
  if (!win_map(&mk2, P.key, bf16, P.C, P.Wk, P.Hk, keys, 2, 2)) return cudaErrorNotSupported;
  ms = mk;
  mc = mk;
  if (has_scale && !win_map(&ms, P.scale, bf16, P.C, P.W, P.H, P.N, kWinT, kWinT)) return cudaErrorNotSupported;
  if (has_cur && !win_map(&mc, P.cur, bf16, P.C, P.W, P.H, P.N, kWinT, kWinT)) return cudaErrorNotSupported;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long ntiles = (long long)P.N * ((P.W + kWinT - 1) / kWinT) * ((P.H + kWinT - 1) / kWinT);
  const int grid = (int)(ntiles < sms ? ntiles : sms);
  if (P.sched) {                                                // one claim counter, zeroed per launch
    cudaError_t e = cudaMemsetAsync(P.sched, 0, sizeof(unsigned), st);
    if (e != cudaSuccess) return e;
  }
#define LSFA_WIN_CASE(VV)                                                                                  \
  if (var == VV) return bf16 ? win_launch<__nv_bfloat16, VV>(P, Q, grid, mk, mk2, ms, mc, st)              \
                             : win_launch<float, VV>(P, Q, grid, mk, mk2, ms, mc, st);
  LSFA_WIN_CASE(kVarWarpOnly) LSFA_WIN_CASE(kVarScale) LSFA_WIN_CASE(kVarScaleCur) LSFA_WIN_CASE(kVarResCur)
#undef LSFA_WIN_CASE
  return cudaErrorNotSupported;
}

}  // namespace lsfa
