// Minimal, known-correct mbarrier hand-offs for compute-sanitizer --tool racecheck.
//
// Two textbook producer/consumer rings, each ordered ONLY by mbarriers (no bar.sync between the roles), re-using
// their slots many times - the two protocols the all-TMA kernels of this repo are built on:
//   A  generic-proxy ring : warp 0 writes a 32-record slot with ordinary st.shared, __syncwarp, one lane does
//                           mbarrier.arrive (release) on full[s]; warp 1 does mbarrier.try_wait (acquire) on
//                           full[s], reads the slot, __syncwarp, one lane arrives on free[s]; warp 0 waits free[s]
//                           before overwriting the slot.          (= the record ring of agg_nhwc_tma_kernel)
//   B  async-proxy ring   : one lane issues cp.async.bulk (global -> shared, mbarrier complete_tx) into stage s;
//                           a consumer warp waits full[s], reads the stage, __syncwarp, arrives on done[s]; the
//                           producer waits done[s] before refilling.   (= the stage ring of every all-TMA kernel)
// Each kernel checks its own result (sum of everything consumed) so that a REAL race would show as a wrong answer.
// Build and run:  nvcc -arch=sm_100a -O2 -lineinfo -o /tmp/repro mbarrier_handoff_repro.cu && /tmp/repro
//                 compute-sanitizer --tool racecheck /tmp/repro
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}" ::"r"(s32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(bytes),
               "r"(s32(bar))
               : "memory");
}

constexpr int RING = 2, ITERS = 64;

__global__ void ring_generic(unsigned long long* out) {
  __shared__ uint64_t full[RING], freeb[RING];
  __shared__ int slot[RING][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < RING; ++s) { mbar_init(&full[s], 1); mbar_init(&freeb[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == 0) {
    for (int it = 0; it < ITERS; ++it) {
      const int s = it % RING;
      if (it >= RING) mbar_wait(&freeb[s], ((it / RING) - 1) & 1);
      slot[s][lane] = it * 32 + lane;                 // generic-proxy write
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[s]);           // release
    }
  } else if (warp == 1) {
    unsigned long long acc = 0;
    for (int it = 0; it < ITERS; ++it) {
      const int s = it % RING;
      mbar_wait(&full[s], (it / RING) & 1);           // acquire
      acc += (unsigned long long)slot[s][(lane + it) & 31];
      __syncwarp();
      if (lane == 0) mbar_arrive(&freeb[s]);
    }
    atomicAdd(out, acc);
  }
}

__global__ void ring_bulk(const int* __restrict__ src, unsigned long long* out) {
  __shared__ uint64_t full[RING], done[RING];
  __shared__ __align__(128) int stage[RING][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < RING; ++s) { mbar_init(&full[s], 1); mbar_init(&done[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < ITERS; ++it) {
        const int s = it % RING;
        if (it >= RING) mbar_wait(&done[s], ((it / RING) - 1) & 1);
        mbar_expect_tx(&full[s], 1024);
        bulk_g2s(stage[s], src + (size_t)it * 256, 1024, &full[s]);     // async-proxy write
      }
    }
  } else if (warp == 1) {
    unsigned long long acc = 0;
    for (int it = 0; it < ITERS; ++it) {
      const int s = it % RING;
      mbar_wait(&full[s], (it / RING) & 1);
      for (int i = lane; i < 256; i += 32) acc += (unsigned long long)stage[s][i];
      __syncwarp();
      if (lane == 0) mbar_arrive(&done[s]);
    }
    atomicAdd(out + 1, acc);
  }
}

int main() {
  int* src;
  unsigned long long* out;
  cudaMalloc(&src, sizeof(int) * 256 * ITERS);
  cudaMallocManaged(&out, 16);
  int* h = new int[256 * ITERS];
  unsigned long long want_b = 0, want_a = 0;
  for (int i = 0; i < 256 * ITERS; ++i) { h[i] = i % 1000; want_b += h[i]; }
  for (int i = 0; i < 32 * ITERS; ++i) want_a += i;
  cudaMemcpy(src, h, sizeof(int) * 256 * ITERS, cudaMemcpyHostToDevice);
  out[0] = out[1] = 0;
  ring_generic<<<1, 64>>>(out);
  ring_bulk<<<1, 64>>>(src, out);
  cudaError_t e = cudaDeviceSynchronize();
  printf("cuda: %s\nA generic ring: got %llu want %llu %s\nB bulk ring:    got %llu want %llu %s\n", cudaGetErrorString(e), out[0], want_a,
         out[0] == want_a ? "OK" : "WRONG", out[1], want_b, out[1] == want_b ? "OK" : "WRONG");
  return (e == cudaSuccess && out[0] == want_a && out[1] == want_b) ? 0 : 1;
}
