// aggregate_backward.h - host interface of the fused operator's tail backward (aggregate_backward.cu)
#pragma once
#include "lsfa_device.cuh"

namespace lsfa {

struct TailBwdParams {
  AggParams P;            // the forward's parameters, records filled in
  const float* og;        // d/d(out)
  const float* vf;        // ww * warp of every element from a warp-only pass of the all-TMA forward kernel (may alias gw:
                          // an element is read before it is overwritten by the same thread), or NULL = gather here
  float* gw;              // d/d(warped feature) = ww * g * scale, or NULL
  float* gscale;          // d/d(scale_map) or NULL
  float* gcur;            // d/d(cur) or NULL
  int add_scale, add_cur;
  float* partT;           // [chunks][N][2][HW] or NULL
  float* partRes;         // [chunks][N][3][HW] or NULL
  float* partRnet;        // [N][tiles][C][4] or NULL
  int tiles;
};

struct TailBwdRequest {
  const float* out_grad;
  float* grad_scale; int add_scale;
  float* grad_cur; int add_cur;
  float* grad_logits; int add_logits;
  float* grad_res; int add_res;
  float* grad_rnet_w; float* grad_rnet_b; int add_rnet;
  bool want_gw;
  float** gw_out;         // where the tail kernel left d/d(warped feature) (inside the workspace)
};

size_t tail_backward_workspace_bytes(int N, int C, int HW, bool want_gw, bool want_logits, bool want_res);
cudaError_t launch_tail_backward(AggParams P, const TailBwdRequest& R, void* workspace, cudaStream_t st);

}  // namespace lsfa
