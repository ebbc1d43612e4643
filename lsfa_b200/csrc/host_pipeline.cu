// host_pipeline.cu - the reference-facing host path behind the C ABI (include/lsfa_ops.h: lsfa_host_aggregate_*).
//
// Replaces, for the fused operator, the reference's host<->device contract: `_load_data`
// (dff_rfcn/core/DataParallelExecutorGroup.py:24-39) + executor forward + `asnumpy()` (core/tester.py:138-145), with
// the key feature kept on the device across a GOP (core/tester.py:246-252).  Pure enqueue code: copies, the C-ABI
// fused kernel, events; no kernel of its own, no allocation, no synchronisation.
#include <cstdint>
#include <cstring>

#include <cuda_runtime.h>

#include "../../include/lsfa_ops.h"

namespace lsfa {
int set_error(int code, const char* fmt, ...);   // cabi.cu: thread-local message behind lsfa_last_error()
}

namespace {

inline size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }

struct SlotLayout {
  size_t key, scale, cur, out, mv, logits, kidx, ws, total;   // byte offsets inside one slot
  size_t ws_bytes;
};

// LsfaAggArgs of one chunk of `m` frames on staging slot `base`
LsfaAggArgs chunk_args(const LsfaHostAggArgs& a, const SlotLayout& L, char* base, int m) {
  LsfaAggArgs g;
  std::memset(&g, 0, sizeof(g));
  g.struct_bytes = (int32_t)sizeof(LsfaAggArgs);
  g.layout = LSFA_LAYOUT_NCHW_F32;
  g.N = m; g.C = a.C; g.H = a.H; g.W = a.W;
  if (a.key_index) {
    g.key = a.key_table;
    g.num_keys = a.num_slots;
    g.key_index = reinterpret_cast<const int32_t*>(base + L.kidx);
  } else {
    g.key = base + L.key;
    g.num_keys = m;
  }
  g.flow_kind = LSFA_FLOW_RAW_I32;
  g.flow = base + L.mv;
  g.mv_h = a.mv_h; g.mv_w = a.mv_w;
  g.im_scale = a.im_scale;
  g.pool_mode = LSFA_POOL_CENTRE2X2;
  g.scale_map = a.scale_map ? base + L.scale : nullptr;
  g.cur = a.cur ? base + L.cur : nullptr;
  g.weight_mode = a.weight_mode;
  g.logits = a.weight_mode == LSFA_W_LOGITS ? reinterpret_cast<const float*>(base + L.logits) : nullptr;
  g.out = base + L.out;
  g.req = LSFA_REQ_WRITE;
  return g;
}

int validate(const LsfaHostAggArgs* a) {
  if (!a) return lsfa::set_error(LSFA_E_BADARG, "args is NULL");
  if (a->struct_bytes != (int32_t)sizeof(LsfaHostAggArgs))
    return lsfa::set_error(LSFA_E_BADARG, "args->struct_bytes=%d, this library expects %zu", a->struct_bytes, sizeof(LsfaHostAggArgs));
  if (a->N <= 0 || a->C <= 0 || a->H <= 0 || a->W <= 0 || a->mv_h <= 0 || a->mv_w <= 0) return lsfa::set_error(LSFA_E_SHAPE, "non-positive dims");
  if ((a->mv_h + 15) / 16 != a->H || (a->mv_w + 15) / 16 != a->W)
    return lsfa::set_error(LSFA_E_SHAPE, "H,W must be ceil(mv_h/16), ceil(mv_w/16): got %dx%d for a %dx%d MV image", a->H, a->W, a->mv_h, a->mv_w);
  if (!(a->im_scale > 0.0)) return lsfa::set_error(LSFA_E_BADARG, "im_scale must be > 0");
  if (a->weight_mode < LSFA_W_NONE || a->weight_mode > LSFA_W_LOGITS)
    return lsfa::set_error(LSFA_E_BADARG, "weight_mode %d: the host path serves NONE, ADD, MEAN and LOGITS", a->weight_mode);
  if (!a->mv || !a->out) return lsfa::set_error(LSFA_E_BADARG, "mv and out are required");
  if (a->weight_mode != LSFA_W_NONE && !a->cur) return lsfa::set_error(LSFA_E_BADARG, "cur is required unless weight_mode is NONE");
  if (a->weight_mode == LSFA_W_LOGITS && !a->logits) return lsfa::set_error(LSFA_E_BADARG, "logits are required for LSFA_W_LOGITS");
  if (a->chunk <= 0 || a->depth < 2 || a->depth > 8) return lsfa::set_error(LSFA_E_BADARG, "chunk must be >= 1 and depth in 2..8");
  if (a->key_index) {
    if (!a->key_table || a->num_slots <= 0) return lsfa::set_error(LSFA_E_BADARG, "GOP mode needs key_table and num_slots > 0");
    if (a->num_new_keys < 0 || (a->num_new_keys > 0 && (!a->key || !a->key_slot)))
      return lsfa::set_error(LSFA_E_BADARG, "GOP mode: num_new_keys > 0 needs key and key_slot");
    for (int i = 0; i < a->num_new_keys; ++i)
      if (a->key_slot[i] < 0 || a->key_slot[i] >= a->num_slots) return lsfa::set_error(LSFA_E_BADARG, "key_slot[%d]=%d outside the table of %d slots", i, a->key_slot[i], a->num_slots);
    for (int n = 0; n < a->N; ++n)
      if (a->key_index[n] < 0 || a->key_index[n] >= a->num_slots) return lsfa::set_error(LSFA_E_BADARG, "key_index[%d]=%d outside the table of %d slots", n, a->key_index[n], a->num_slots);
  } else if (!a->key) {
    return lsfa::set_error(LSFA_E_BADARG, "key is required (one key feature per frame) when key_index is NULL");
  }
  return LSFA_OK;
}

SlotLayout slot_layout(const LsfaHostAggArgs& a) {
  const int m = a.chunk < a.N ? a.chunk : a.N;
  const size_t F = (size_t)m * a.C * a.H * a.W * 4;
  SlotLayout L{};
  size_t at = 0;
  L.key = at;    at += a.key_index ? 0 : up256(F);
  L.scale = at;  at += a.scale_map ? up256(F) : 0;
  L.cur = at;    at += a.cur ? up256(F) : 0;
  L.out = at;    at += up256(F);
  L.mv = at;     at += up256((size_t)m * a.mv_h * a.mv_w * 8);
  L.logits = at; at += a.weight_mode == LSFA_W_LOGITS ? up256((size_t)m * 2 * a.H * a.W * 4) : 0;
  L.kidx = at;   at += a.key_index ? up256((size_t)m * 4) : 0;
  L.ws = at;
  LsfaAggArgs g = chunk_args(a, L, nullptr, m);
  g.key = reinterpret_cast<const void*>(256);      // non-NULL placeholders: the size query does not dereference
  L.ws_bytes = lsfa_warp_scale_aggregate_workspace_bytes(&g);
  at += up256(L.ws_bytes);
  L.total = at;
  return L;
}

size_t mv_rows_bytes(int h, int w) {
  size_t rows = 0;
  for (int r = 0; r < h; ++r) rows += (r % 16 == 7 || r % 16 == 8);
  return rows * (size_t)w * 8;
}

}  // namespace

extern "C" {

size_t lsfa_host_aggregate_staging_bytes(const LsfaHostAggArgs* a) {
  if (validate(a) != LSFA_OK) return 0;
  return slot_layout(*a).total * (size_t)a->depth;
}

int lsfa_host_aggregate_bytes(const LsfaHostAggArgs* a, size_t* h2d, size_t* d2h) {
  if (int r = validate(a)) return r;
  const size_t F = (size_t)a->C * a->H * a->W * 4;
  size_t in = (size_t)a->N * ((a->scale_map ? F : 0) + (a->cur ? F : 0) + mv_rows_bytes(a->mv_h, a->mv_w) +
                              (a->weight_mode == LSFA_W_LOGITS ? (size_t)2 * a->H * a->W * 4 : 0));
  in += a->key_index ? (size_t)a->num_new_keys * F + (size_t)a->N * 4 : (size_t)a->N * F;
  if (h2d) *h2d = in;
  if (d2h) *d2h = (size_t)a->N * F;
  return LSFA_OK;
}

int lsfa_host_aggregate_f32_nchw(const LsfaHostAggArgs* a) {
  if (int r = validate(a)) return r;
  const SlotLayout L = slot_layout(*a);
  if (!a->staging || a->staging_bytes < L.total * (size_t)a->depth)
    return lsfa::set_error(LSFA_E_BADARG, "staging too small: need %zu bytes", L.total * (size_t)a->depth);
  if (reinterpret_cast<uintptr_t>(a->staging) & 255) return lsfa::set_error(LSFA_E_ALIGN, "staging must be 256-byte aligned");
  cudaStream_t s_in = static_cast<cudaStream_t>(a->stream_in), s_run = static_cast<cudaStream_t>(a->stream_run),
               s_out = static_cast<cudaStream_t>(a->stream_out);
  const int chunk = a->chunk < a->N ? a->chunk : a->N;
  const int n_chunks = (a->N + chunk - 1) / chunk;
  const size_t F1 = (size_t)a->C * a->H * a->W * 4;                 // one frame's feature
  constexpr int kMaxEv = 3 * 8 + 3;
  cudaEvent_t ev_in[8], ev_run[8], ev_out[8], ev_a = nullptr, ev_b = nullptr, ev_keys = nullptr;   // rings of `depth` events
  bool used[8] = {false, false, false, false, false, false, false, false};
  int n_ev = 0;
  cudaEvent_t all[kMaxEv];
  auto make = [&](cudaEvent_t* e) {
    cudaError_t r = cudaEventCreateWithFlags(e, cudaEventDisableTiming);
    if (r == cudaSuccess) all[n_ev++] = *e;
    return r;
  };
  cudaError_t err = cudaSuccess;
  int rc = LSFA_OK;
#define HP_CUDA(x) do { err = (x); if (err != cudaSuccess) goto done; } while (0)
  for (int i = 0; i < a->depth; ++i) {
    HP_CUDA(make(&ev_in[i]));
    HP_CUDA(make(&ev_run[i]));
    HP_CUDA(make(&ev_out[i]));
  }
  HP_CUDA(make(&ev_a));
  HP_CUDA(make(&ev_b));
  HP_CUDA(make(&ev_keys));
  // the previous call on these streams may still be reading / draining the staging slots (and the key table)
  HP_CUDA(cudaEventRecord(ev_a, s_run));
  HP_CUDA(cudaStreamWaitEvent(s_in, ev_a, 0));
  HP_CUDA(cudaEventRecord(ev_b, s_out));
  HP_CUDA(cudaStreamWaitEvent(s_run, ev_b, 0));
  if (a->key_index && a->num_new_keys > 0) {       // the key frames of this batch enter the device table (tester.py:251-252)
    for (int i = 0; i < a->num_new_keys; ++i)
      HP_CUDA(cudaMemcpyAsync(reinterpret_cast<char*>(a->key_table) + (size_t)a->key_slot[i] * F1,
                              reinterpret_cast<const char*>(a->key) + (size_t)i * F1, F1, cudaMemcpyHostToDevice, s_in));
  }
  for (int i = 0; i < n_chunks; ++i) {
    const int slot = i % a->depth;
    const int lo = i * chunk, m = (a->N - lo) < chunk ? (a->N - lo) : chunk;
    char* base = static_cast<char*>(a->staging) + (size_t)slot * L.total;
    const size_t Fm = (size_t)m * F1;
    // ---- H2D ----
    if (used[slot]) HP_CUDA(cudaStreamWaitEvent(s_in, ev_run[slot], 0));            // the slot's previous kernel has read its inputs
    if (!a->key_index) HP_CUDA(cudaMemcpyAsync(base + L.key, reinterpret_cast<const char*>(a->key) + (size_t)lo * F1, Fm, cudaMemcpyHostToDevice, s_in));
    if (a->scale_map) HP_CUDA(cudaMemcpyAsync(base + L.scale, reinterpret_cast<const char*>(a->scale_map) + (size_t)lo * F1, Fm, cudaMemcpyHostToDevice, s_in));
    if (a->cur) HP_CUDA(cudaMemcpyAsync(base + L.cur, reinterpret_cast<const char*>(a->cur) + (size_t)lo * F1, Fm, cudaMemcpyHostToDevice, s_in));
    if (a->weight_mode == LSFA_W_LOGITS)
      HP_CUDA(cudaMemcpyAsync(base + L.logits, a->logits + (size_t)lo * 2 * a->H * a->W, (size_t)m * 2 * a->H * a->W * 4, cudaMemcpyHostToDevice, s_in));
    if (a->key_index) HP_CUDA(cudaMemcpyAsync(base + L.kidx, a->key_index + lo, (size_t)m * 4, cudaMemcpyHostToDevice, s_in));
    rc = lsfa_mv_centre_rows_h2d(a->mv + (size_t)lo * a->mv_h * a->mv_w * 2, base + L.mv, m, a->mv_h, a->mv_w, nullptr, s_in);
    if (rc != LSFA_OK) goto done;
    HP_CUDA(cudaEventRecord(ev_in[slot], s_in));
    // ---- fused kernel ----
    HP_CUDA(cudaStreamWaitEvent(s_run, ev_in[slot], 0));
    if (used[slot]) HP_CUDA(cudaStreamWaitEvent(s_run, ev_out[slot], 0));           // the slot's previous output has left
    {
      LsfaAggArgs g = chunk_args(*a, L, base, m);
      if (L.ws_bytes) { g.workspace = base + L.ws; g.workspace_bytes = L.ws_bytes; }
      rc = lsfa_warp_scale_aggregate(&g, s_run);
      if (rc != LSFA_OK) goto done;
    }
    HP_CUDA(cudaEventRecord(ev_run[slot], s_run));
    // ---- D2H ----
    HP_CUDA(cudaStreamWaitEvent(s_out, ev_run[slot], 0));
    HP_CUDA(cudaMemcpyAsync(reinterpret_cast<char*>(a->out) + (size_t)lo * F1, base + L.out, Fm, cudaMemcpyDeviceToHost, s_out));
    HP_CUDA(cudaEventRecord(ev_out[slot], s_out));
    used[slot] = true;
  }
#undef HP_CUDA
done:
  for (int i = 0; i < n_ev; ++i) cudaEventDestroy(all[i]);      // released by the runtime once the recorded work has completed
  if (err != cudaSuccess) return lsfa::set_error(LSFA_E_CUDA, "host_aggregate: %s", cudaGetErrorString(err));
  return rc;
}

}  // extern "C"
