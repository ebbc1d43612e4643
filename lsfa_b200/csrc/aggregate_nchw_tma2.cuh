// agg_nchw_tma2_kernel - the all-TMA kernel for planes that need TWO pixel parts (e.g. 68x120),
// as a 2-CTA thread-block cluster with a MULTICAST key load.
//
// In the single-CTA kernel each pixel part TMA-loads the whole key plane, so a two-part plane moves
// 20 B per element between L2 and the SMs instead of 16 B and the kernel stops at the SM<-L2 ceiling
// (measured 0.81 of HBM peak at 1024x68x120).  Here CTA rank r of a cluster owns pixel part r of the
// SAME (frame, channel chunk) item; the leader issues ONE
//     cp.async.bulk...multicast::cluster   (key planes -> the same smem offset in both CTAs,
//                                           complete_tx on both CTAs' full[s])
// and each CTA loads only its own slice of scale / cur and stores its own slice of out.
//
// Cross-CTA protocol per ring stage s (use u of the stage = round r):
//   leader producer : [r>0: own stage free] [r>0: wait peer_free[s]] claim item ->
//                     write descriptor locally and into the peer (st.shared::cluster) ->
//                     arrive.release.cluster on the peer's desc_ready[s] -> arm full[s] -> multicast key,
//                     own scale/cur slices
//   peer producer   : [r>0: own stage free -> arrive.release.cluster on the leader's peer_free[s]]
//                     wait desc_ready[s] -> arm full[s] (key bytes come from the leader) -> own slices
// Consumers are the shared tma_consumer_loop with fixed_part = cluster rank.
#pragma once
#include "aggregate_nchw_tma.cuh"

namespace lsfa {

constexpr int kTma2HeaderBytes = 512;   // full[8] done[8] desc_ready[8] peer_free[8] + descriptors

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_smem_ptr` in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(const void* local_smem_ptr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(local_smem_ptr)), "r"(rank));
  return r;
}
__device__ __forceinline__ void remote_arrive(uint32_t remote_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar) : "memory");
}
__device__ __forceinline__ void remote_store_v2(uint32_t remote_addr, int a, int b) {
  asm volatile("st.shared::cluster.v2.u32 [%0], {%1, %2};" ::"r"(remote_addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {   // acquire at cluster scope
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1, %2;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(0x989680u)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s_multicast(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                                   uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}

template <int K, int PPT, int VAR>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTmaThreads, 1)
agg_nchw_tma2_kernel(const __grid_constant__ AggParams P) {
  static_assert(VAR != kVarRuntime, "the TMA kernels are only built for the compile-time variants");
  constexpr bool has_scale = VAR == kVarScale || VAR == kVarScaleCur;
  constexpr bool has_cur = VAR == kVarScaleCur || VAR == kVarResCur;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* done = full + kMaxStages;
  uint64_t* desc_ready = done + kMaxStages;      // in the PEER: the leader's descriptor for stage s has landed
  uint64_t* peer_free = desc_ready + kMaxStages;  // in the LEADER: the peer's stage s may be overwritten
  volatile int2* desc = reinterpret_cast<volatile int2*>(smem_raw + 256);
  unsigned char* ring = smem_raw + kTma2HeaderBytes;
  float* res_s = reinterpret_cast<float*>(ring + (size_t)P.stages * P.stage_bytes);
  float4* rnet_s = reinterpret_cast<float4*>(res_s + (PPT <= 5 ? 0 : 3 * PPT * kTmaConsumers));

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();        // = pixel part of this CTA
  const bool has_bypass = P.bypass != nullptr;

  if (tid == 0) {
    for (int s = 0; s < P.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&done[s], kTmaConsumerWarps);
      mbar_init(&desc_ready[s], 1);
      mbar_init(&peer_free[s], 1);
    }
    fence_barrier_init();
  }
  cluster_sync_all();   // both CTAs' barriers exist before any remote arrive / multicast

  if (warp == kTmaConsumerWarps) {
    if ((tid & 31) == 0) {
      const int part = (int)rank;
      const int pix0 = part * P.part_pix;
      const uint32_t len_bytes = (uint32_t)(min(P.part_pix, P.HW - pix0)) * 4u;   // this CTA's slice of one plane
      const uint32_t io_tx = (uint32_t)K * len_bytes;
      const int n_clusters = (int)(gridDim.x >> 1), cluster_id = (int)(blockIdx.x >> 1);

      // ---- work source (leader only): per-frame queues or a static range over (frame, chunk) ----
      unsigned* sched = P.sched;
      const long long cluster_items = (long long)P.N * P.chunks;
      int f = (int)(((long long)cluster_id * P.N) / n_clusters), c = 0, cend = 0, hops = 0;
      if (sched == nullptr) {
        const long long i0 = cluster_items * cluster_id / n_clusters, i1 = cluster_items * (cluster_id + 1) / n_clusters;
        f = (int)(i0 / P.chunks);
        c = (int)(i0 - (long long)f * P.chunks);
        hops = (int)(i1 - i0);
        cend = P.chunks;
      }
      auto next_item = [&](int& n, int& chunk) -> bool {
        if (sched == nullptr) {
          if (hops <= 0) return false;
          --hops;
          n = f;
          chunk = c;
          if (++c == P.chunks) {
            c = 0;
            ++f;
          }
          return true;
        }
        while (true) {
          if (c < cend) {
            n = f;
            chunk = c++;
            return true;
          }
          if (hops >= P.N) return false;
          const int got = (int)atomicAdd(sched + f, (unsigned)kTmaClaim);
          if (got < P.chunks) {
            c = got;
            cend = min(got + kTmaClaim, P.chunks);
          } else {
            f = (f + 1 == P.N) ? 0 : f + 1;
            ++hops;
          }
        }
      };
      // this CTA's slices of the streams (one bulk copy per plane: plane starts are 16-byte aligned, HW % 4 == 0)
      auto issue_own_loads = [&](int s, int n, int chunk, bool byp) {
        unsigned char* st = ring + (size_t)s * P.stage_bytes;
        const size_t e0 = ((size_t)n * P.C + (size_t)chunk * K) * P.HW + pix0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const size_t ek = e0 + (size_t)k * P.HW;
          const uint32_t dk = (uint32_t)k * (uint32_t)P.part_pix * 4u;
          if (!byp && has_scale) bulk_g2s(st + P.off_scale + dk, static_cast<const float*>(P.scale) + ek, len_bytes, &full[s]);
          if (has_cur) bulk_g2s(st + P.off_io + dk, static_cast<const float*>(P.cur) + ek, len_bytes, &full[s]);
        }
      };
      auto issue_store = [&](int s, int n, int chunk) {
        unsigned char* st = ring + (size_t)s * P.stage_bytes;
        float* dst = static_cast<float*>(P.out) + ((size_t)n * P.C + (size_t)chunk * K) * P.HW + pix0;
#pragma unroll
        for (int k = 0; k < K; ++k)
          bulk_s2g(dst + (size_t)k * P.HW, st + P.off_io + (uint32_t)k * (uint32_t)P.part_pix * 4u, len_bytes);
        bulk_commit();
      };
      // Refill of stage s for its round r, in two steps so the key multicast overlaps the drain of the out store:
      //   begin_refill : hand-shake (the consumers of BOTH CTAs are done with the stage: its key region is free),
      //                  descriptor exchange, arm full[s], multicast key.        -> false: work exhausted
      //   finish_refill: after this CTA's store has drained the io region: this CTA's own scale / cur slices.
      int nn = -1, nchunk = 0;                       // leader: the item claimed ahead of time (hides the atomic)
      bool have_next = false;
      if (rank == 0) have_next = next_item(nn, nchunk);
      int cur_n = -1, cur_chunk = 0;
      bool cur_byp = false;
      auto begin_refill = [&](int s, int r) -> bool {
        if (rank == 0) {
          if (r > 0) mbar_wait_cluster(&peer_free[s], (unsigned)(r - 1) & 1u);   // the peer's consumers left the stage
          cur_n = have_next ? nn : -1;
          cur_chunk = nchunk;
          desc[s].x = cur_n;
          desc[s].y = cur_chunk;
          remote_store_v2(map_to_cta((const void*)&desc[s], 1), cur_n, cur_chunk);
          remote_arrive(map_to_cta(&desc_ready[s], 1));
        } else {
          mbar_wait_cluster(&desc_ready[s], (unsigned)r & 1u);
          cur_n = desc[s].x;
          cur_chunk = desc[s].y;
        }
        if (cur_n < 0) {
          mbar_arrive(&full[s]);                      // stop sentinel for the local consumers
          return false;
        }
        cur_byp = has_bypass && __ldg(P.bypass + cur_n) != 0;
        uint32_t bytes = has_cur ? io_tx : 0u;
        if (!cur_byp) bytes += P.key_bytes + (has_scale ? io_tx : 0u);
        mbar_expect_tx(&full[s], bytes);
        if (rank == 0) {
          if (!cur_byp) {
            const int kn = key_slot(P, cur_n);
            const float* ksrc = static_cast<const float*>(P.key) + ((size_t)kn * P.C + (size_t)cur_chunk * K) * P.HWk;
            bulk_g2s_multicast(ring + (size_t)s * P.stage_bytes, ksrc, P.key_bytes, &full[s], (uint16_t)0x3);
          }
          have_next = next_item(nn, nchunk);          // claim the following item while this one is in flight
        }
        return true;
      };
      auto finish_refill = [&](int s) { issue_own_loads(s, cur_n, cur_chunk, cur_byp); };

      int live = 0;
      bool stopped = false;
      for (int s = 0; s < P.stages; ++s) {
        if (begin_refill(s, 0)) {
          finish_refill(s);
          ++live;
        } else {
          stopped = true;
          break;
        }
      }
      int s = 0, round = 0;
      unsigned ph = 0;
      while (live > 0) {
        mbar_wait(&done[s], ph);                       // own consumers finished this stage; out slice is in smem
        const int sn = desc[s].x, schunk = desc[s].y;  // read BEFORE the leader may overwrite this descriptor remotely
        if (rank == 1 && !stopped) remote_arrive(map_to_cta(&peer_free[s], 0));   // key region free: leader may multicast
        issue_store(s, sn, schunk);
        --live;
        if (!stopped) {
          if (begin_refill(s, round + 1)) {
            bulk_wait_read_all();                      // the store has drained the io region of this stage
            finish_refill(s);
            ++live;
          } else {
            stopped = true;
          }
        }
        if (++s == P.stages) {
          s = 0;
          ph ^= 1u;
          ++round;
        }
      }
      bulk_wait_all();
    }
  } else {
    tma_consumer_loop<K, PPT, VAR, false>(P, full, done, desc, ring, res_s, rnet_s, tid, (int)rank);
  }
  cluster_sync_all();   // neither CTA leaves while the other may still signal or multicast into it
}

template <int VAR>
cudaError_t launch_tma2_variant(const AggParams& P, size_t smem, int grid, cudaStream_t st);

#define LSFA_TMA2_LAUNCH(VAR, KK, PP)                                                             \
  if (P.K == KK && ppt == PP) {                                                                   \
    auto kfn = agg_nchw_tma2_kernel<KK, PP, VAR>;                                                 \
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return e;                                                               \
    kfn<<<grid, kTmaThreads, smem, st>>>(P);                                                      \
    return cudaPeekAtLastError();                                                                 \
  }

}  // namespace lsfa
