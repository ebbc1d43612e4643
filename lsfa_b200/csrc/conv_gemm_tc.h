// conv_gemm_tc.h - host interface of the tensor-core implicit-GEMM convolutions (conv_gemm_tc.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stddef.h>

namespace lsfa {
namespace tc {

enum Epilogue { EPI_STORE_RELU = 0, EPI_STORE = 1, EPI_COSINE = 2, EPI_NQ = 3 };

struct ConvParams {
  // problem (filled by the caller)
  int NB, H, W, Cin, Cout, taps;        // NB images (even: pairs b, b + NB/2), stride-1 'same' convolution, taps = 1 | 9
  const float* bias;                    // (Cout)
  __nv_bfloat16* out;                   // STORE: (NB,H,W,Cout) bf16
  float* partial;                       // COSINE: (2 * Cout/256, 3, NB/2 * H*W) f32 (two column halves per chunk)
  float* logits;                        // NQ: (NB/2, 2, H, W) f32
  const float *nq_w2, *nq_b2, *nq_w3, *nq_b3;   // NQ tail: (16,256), (16), (16), (1)
  // tiling (filled by launch_conv)
  int BW, BH, bw_shift, tiles_x, tiles_y, n_chunks, kc_per_tap, k_steps, num_items, pair_items;
};

// nullptr = enqueued; otherwise a static message (nothing was launched)
const char* launch_conv(const void* x, const void* w, ConvParams P, int epi, int sms, cudaStream_t stream);
void launch_cosine_finalize(const float* partial, float* logits, int n_chunks, int N, int HW, cudaStream_t stream);
void launch_pack_weight(const float* w, void* out, int Cout, int Cin, int kk, cudaStream_t stream);
void pick_tile(int H, int W, int* BW, int* BH);

}  // namespace tc
}  // namespace lsfa
