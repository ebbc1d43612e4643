// agg_nhwc_tma_kernel - all-TMA, warp-specialised channels-last kernel ("gather by bulk copy").
//
// Same operator chain as agg_nhwc_kernel (SYM:571-576, 308/470/680, 104-108, 236, 315;
// operator_py/choose_feat.py:23-31).  In NHWC a pixel's C channels are one contiguous run
// (C=1024: 2 KB bf16 / 4 KB fp32), so each of the 4 bilinear taps of an output pixel IS a bulk
// copy: no thread touches global memory for the feature streams.
//
//   warp 0   record warp : claims batches of 32 consecutive output pixels (one atomic per batch, or a
//                          static stride without scratch), one lane per pixel does the whole index
//                          chain (a3/a5/a6 pooling in float64, a7 grid, a8 floor/weights, a13 softmax,
//                          blend fold) into a ring of record batches in shared memory
//   warps 1-2 producers  : per group of G pixels: 2G tap-row copies (one lane each; the two taps of a row are
//                          neighbours in memory) + one copy of the scale run + one of the cur run
//                          HBM/L2 --cp.async.bulk--> stage; a stage is refilled the moment its consumers are done;
//                          producer k issues items k, k+2, ...
//   warps 3+ consumers   : 16 warps as two groups of 8 (group k computes items k, k+2, ...: two stages at once) or, for
//                          the residual variant, one group of 16:
//                          6 x LDS.128 + packed fp32 math + one 16-byte streaming store per 16 bytes of output
//                          (a pixel's channel run is 16-byte aligned and contiguous: every store instruction
//                          writes 512 contiguous bytes).  A bulk store from shared memory was measured first:
//                          the producer then waits ~0.8 us per group for the store to leave the copy engine's
//                          queue behind the loads already issued (0.56-0.62 of the HBM peak).
//
// Barriers: rec_full/rec_free (record warp <-> producer), full (tx-count) / done (consumers ->
// producer) per stage.  The LDG form (agg_nhwc_kernel) is latency bound: 18 warps per SM, each
// waiting a DRAM round trip per batch of 12 loads (ncu long_scoreboard 3.9 per issue); here the
// copy engine keeps 3 stages (144 KB) in flight per SM whatever the warps do.
// Every value is computed by the same expression chain as agg_nhwc_kernel: bit-identical results.
#pragma once
#include <cstdlib>
#include <type_traits>

#include "aggregate_nchw_tma.cuh"   // bulk_s2g / commit / wait / fence helpers

namespace lsfa {

constexpr int kNtAllConsumerWarps = 16;   // consumer warps of the CTA
// They work as NtPlan::groups groups (1 or 2): a group computes one stage together, group k takes items k, k+groups, ...
// Same-box A/B (fraction of the copy peak; bf16 batch 512 / 68x120 bf16 / fp32 / shipped non-key bf16 / warp alone bf16):
//   2 groups of 8: 0.944 / 0.94 / 0.995 / 0.74 / 0.71        1 group of 16: 0.926 / 0.925 / 0.984 / 0.775 / 0.635
// -> two groups, except for the residual variant (one pass in flight per warp there: 16 warps on a stage hide more).
constexpr int kNtMaxGroups = 2;
#ifndef LSFA_NT_PRODUCERS
#define LSFA_NT_PRODUCERS 2
#endif
constexpr int kNtProducers = LSFA_NT_PRODUCERS;   // producer warps: warp k issues the copies of items k, k + NP, ... (one warp alone
                                                  // needs ~0.9 us of dependent instructions per pixel group: it was the bottleneck)
constexpr int kNtFirstConsumer = 1 + kNtProducers;
// Stage s is always filled by the same producer warp and always drained by the same consumer group only if the ring
// depth is a multiple of both counts.  That is what makes the parity waits safe: a warp that waits for "the previous
// use of stage s" created that use's predecessor itself, so the barrier is never two phases behind (a parity wait
// on a barrier two phases behind returns at once).
constexpr int kNtStageMultiple = (kNtProducers % kNtMaxGroups == 0) ? kNtProducers
                               : (kNtMaxGroups % kNtProducers == 0) ? kNtMaxGroups : kNtProducers * kNtMaxGroups;
static_assert(kNtStageMultiple <= 8, "producer / consumer-group counts need too deep a ring");
constexpr int kNtThreads = (kNtFirstConsumer + kNtAllConsumerWarps) * 32;
constexpr int kNtMaxStages = 8;     // stages actually used: NtPlan::stages (ring budget / stage size)
constexpr int kNtRecRing = 4;       // record batches in flight ahead of the producer
constexpr int kNtMaxG = 8;          // pixels per stage, at most
constexpr unsigned kNtStageBudget = 48u * 1024u;
constexpr unsigned kNtRingBudget = 4u * kNtStageBudget;

struct __align__(16) NtRec {        // one output pixel's sampling record (64 B)
  float w00, w01, w10, w11;         // tap weights, blend weight folded in
  float wc, ww, r0, r1;             // blend weights; pooled residual (a10)
  float r2;
  int i00, i01, i10, i11;           // key pixel indices of the taps (clamped in-bounds)
  int pad0, pad1, pad2;
};
static_assert(sizeof(NtRec) == 64, "NtRec must be 64 bytes");

struct __align__(16) NtPixW {       // what the consumers need of a record (48 B)
  float w00, w01, w10, w11, wc, ww, r0, r1, r2, p0, p1, p2;
};

struct __align__(16) NtDesc {       // per-stage descriptor (512 B)
  int npix;                         // pixels in the stage; < 0 = stop
  int bypass;                       // ChooseFeat: the frame keeps its current feature
  int pad0, pad1;
  unsigned long long out_elem;      // element offset of the first output pixel
  unsigned long long pad2;
  NtPixW pw[kNtMaxG];
  unsigned char fill[512 - 32 - sizeof(NtPixW) * kNtMaxG];
};
static_assert(sizeof(NtDesc) == 512, "NtDesc must be 512 bytes");

// shared-memory layout (bytes)
constexpr unsigned kNtOffHdr = 192;                               // int4 batch header per record slot
constexpr unsigned kNtOffDesc = 256;
constexpr unsigned kNtOffRecs = kNtOffDesc + kNtMaxStages * 512;
constexpr unsigned kNtOffRing = kNtOffRecs + kNtRecRing * 32 * 64; // 128-byte aligned: 256 + 4096 + 8192 = 12544

struct NtPlan {
  int G;                    // pixels per stage (power of two, <= 8)
  int stages;               // ring depth (<= kNtMaxStages)
  int merge_pairs;          // the two taps of a row travel as one copy when they are neighbours in memory
  int groups;               // consumer groups (1 or 2)
  unsigned slot;            // bytes of one pixel's channel run
  unsigned stage_bytes;
  unsigned off_scale, off_io;
  size_t smem;
};

template <typename T, int VAR, bool FQ>
__global__ void __launch_bounds__(kNtThreads, 1)
agg_nhwc_tma_kernel(const __grid_constant__ AggParams P, const NtPlan Q) {
  static_assert(VAR != kVarRuntime, "only the compile-time variants");
  constexpr bool has_scale = VAR == kVarScale || VAR == kVarScaleCur;
  constexpr bool has_cur = VAR == kVarScaleCur || VAR == kVarResCur;
  constexpr bool has_res = VAR == kVarResCur;
  using V = Vec16<T>;
  constexpr int L = V::kLanes;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* done = full + kNtMaxStages;
  uint64_t* rec_full = done + kNtMaxStages;
  uint64_t* rec_free = rec_full + kNtRecRing;
  volatile int4* hdr = reinterpret_cast<volatile int4*>(smem_raw + kNtOffHdr);   // (n, p0, npix | -1, bypass)
  NtDesc* desc = reinterpret_cast<NtDesc*>(smem_raw + kNtOffDesc);
  NtRec* recs = reinterpret_cast<NtRec*>(smem_raw + kNtOffRecs);
  unsigned char* ring = smem_raw + kNtOffRing;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = Q.G;
  const int S = Q.stages;
  const unsigned slot = Q.slot;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&done[s], kNtAllConsumerWarps / Q.groups);
    }
    for (int s = 0; s < kNtRecRing; ++s) {
      mbar_init(&rec_full[s], 1);
      mbar_init(&rec_free[s], kNtProducers);
    }
    fence_barrier_init();
  }
  __syncthreads();

  const int bpf = (P.HW + 31) / 32;                    // record batches per frame
  const long long nbatches = (long long)P.N * bpf;

  if (warp == 0) {
    // ===================================== record warp =====================================
    int rs = 0;
    for (long long it = 0;; ++it) {
      long long b;
      if (P.sched != nullptr) {
        unsigned got = 0;
        if (lane == 0) got = atomicAdd(P.sched, 1u);
        b = (long long)__shfl_sync(0xffffffffu, got, 0);
      } else {
        b = blockIdx.x + it * (long long)gridDim.x;
      }
      if (it >= kNtRecRing) mbar_wait(&rec_free[rs], (unsigned)((it / kNtRecRing) - 1) & 1u);
      if (b >= nbatches) {
        if (lane == 0) {
          hdr[rs].z = -1;
          mbar_arrive(&rec_full[rs]);
        }
        break;
      }
      const int n = (int)(b / bpf);
      const int p0 = (int)(b - (long long)n * bpf) * 32;
      const int npix = min(32, P.HW - p0);
      const bool byp = P.bypass != nullptr && __ldg(P.bypass + n) != 0;
      NtRec rec;
      rec.w00 = rec.w01 = rec.w10 = rec.w11 = rec.wc = rec.ww = rec.r0 = rec.r1 = rec.r2 = 0.f;
      rec.i00 = rec.i01 = rec.i10 = rec.i11 = 0;
      rec.pad0 = rec.pad1 = rec.pad2 = 0;
      if (lane < npix && !byp) {
        const int p = p0 + lane;
        const int y = p / P.W, x = p - y * P.W;
        const PixelLoads ld = issue_pixel_loads(P, n, y, x);
        if (has_res) {                                 // issued with the MV taps: one DRAM round trip, not two
          rec.r0 = __ldg(P.res + ((size_t)n * 3 + 0) * P.HW + p);
          rec.r1 = __ldg(P.res + ((size_t)n * 3 + 1) * P.HW + p);
          rec.r2 = __ldg(P.res + ((size_t)n * 3 + 2) * P.HW + p);
        }
        const PixelRec t = finish_pixel(P, ld, n, y, x, /*fold=*/true);
        rec.w00 = t.w00; rec.w01 = t.w01; rec.w10 = t.w10; rec.w11 = t.w11;
        rec.wc = t.wc; rec.ww = t.ww;
        rec.i00 = t.i00; rec.i01 = t.i01; rec.i10 = t.i10; rec.i11 = t.i11;
      }
      recs[rs * 32 + lane] = rec;
      if (lane == 0) {
        hdr[rs].x = n;
        hdr[rs].y = p0;
        hdr[rs].w = byp ? 1 : 0;
        hdr[rs].z = npix;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&rec_full[rs]);       // release: records + header visible to the producer
      if (++rs == kNtRecRing) rs = 0;
    }
    return;
  }

  if (warp < kNtFirstConsumer) {
    // ================================ producer warps (32 lanes each) ================================
    // every producer warp walks the same sequence of items (record batches -> pixel groups) and keeps the same
    // stage / wrap counters; it issues the copies of the items whose number is congruent to its rank
    const int pk = warp - 1;
    int turn = 0;                                      // item number modulo kNtProducers
    const T* __restrict__ key = static_cast<const T*>(P.key);
    const T* __restrict__ scale = static_cast<const T*>(P.scale);
    const T* __restrict__ cur = static_cast<const T*>(P.cur);
    unsigned wraps = 0;                                // times the stage ring has wrapped
    int s = 0;
    int rs = 0;
    for (long long rb = 0;; ++rb) {
      mbar_wait(&rec_full[rs], (unsigned)(rb / kNtRecRing) & 1u);
      const int npix_b = hdr[rs].z;
      if (npix_b < 0) break;
      const int n = hdr[rs].x, p0 = hdr[rs].y;
      const bool byp = hdr[rs].w != 0;
      const int kn = key_slot(P, n);
      const T* kbase = key + (size_t)kn * P.HWk * P.C;
      for (int g0 = 0; g0 < npix_b; g0 += G) {
        const int np = min(G, npix_b - g0);
        const bool mine = turn == pk;
        if (++turn == kNtProducers) turn = 0;
        if (!mine) {
          if (++s == S) {
            s = 0;
            ++wraps;
          }
          continue;
        }
        if (wraps >= 1) mbar_wait(&done[s], (wraps - 1u) & 1u);   // its consumers are done with the stage's previous item
        const size_t oe = ((size_t)n * P.HW + (size_t)(p0 + g0)) * P.C;
        const NtRec* rr = recs + rs * 32 + g0;
        if (lane < np) {
          const NtRec r = rr[lane];
          NtPixW w;
          w.w00 = r.w00; w.w01 = r.w01; w.w10 = r.w10; w.w11 = r.w11;
          w.wc = r.wc; w.ww = r.ww; w.r0 = r.r0; w.r1 = r.r1; w.r2 = r.r2;
          w.p0 = w.p1 = w.p2 = 0.f;
          desc[s].pw[lane] = w;
        }
        const unsigned run = (unsigned)np * slot;
        if (lane == 0) {
          desc[s].npix = np;
          desc[s].bypass = byp ? 1 : 0;
          desc[s].out_elem = (unsigned long long)oe;
        }
        __syncwarp();
        if (lane == 0) {
          unsigned bytes = has_cur ? run : 0u;
          if (!byp) bytes += 4u * run + (has_scale ? run : 0u);
          mbar_expect_tx(&full[s], bytes);             // release: the descriptor is visible with the data
        }
        __syncwarp();
        unsigned char* st = ring + (size_t)s * Q.stage_bytes;
        if (!byp) {
          // stage layout [pixel][tap]; the two taps of a row are neighbours in the key feature unless the
          // sampling position was clamped at the plane's edge: one copy of two channel runs per tap row
          for (int i = lane; i < 2 * np; i += 32) {
            const int g = i >> 1, row = i & 1;
            const NtRec& r = rr[g];
            const int ia = row ? r.i10 : r.i00, ib = row ? r.i11 : r.i01;
            unsigned char* dst = st + (size_t)(g * 4 + row * 2) * slot;
            if (ib == ia + 1 && Q.merge_pairs) {
              bulk_g2s(dst, kbase + (size_t)ia * P.C, 2u * slot, &full[s]);
            } else {
              bulk_g2s(dst, kbase + (size_t)ia * P.C, slot, &full[s]);
              bulk_g2s(dst + slot, kbase + (size_t)ib * P.C, slot, &full[s]);
            }
          }
          if (has_scale && lane == 30) bulk_g2s(st + Q.off_scale, scale + oe, run, &full[s]);
        }
        if (has_cur && lane == 31) bulk_g2s(st + Q.off_io, cur + oe, run, &full[s]);
        if (++s == S) {
          s = 0;
          ++wraps;                                     // fills of stage s so far = wraps (+1 for stages < s)
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&rec_free[rs]);
      if (++rs == kNtRecRing) rs = 0;
    }
    // one stop marker per consumer group: in the slots of items j and j+1 (once their previous items are consumed)
    for (int k = 0; k < Q.groups; ++k) {
      const bool mine = turn == pk;
      if (++turn == kNtProducers) turn = 0;
      if (mine) {
        if (wraps >= 1) mbar_wait(&done[s], (wraps - 1u) & 1u);
        if (lane == 0) {
          desc[s].npix = -1;
          mbar_arrive(&full[s]);
        }
      }
      if (++s == S) {
        s = 0;
        ++wraps;
      }
    }
    return;
  }

  // ===================================== consumer warps =====================================
  const int gw = kNtAllConsumerWarps / Q.groups;       // warps per consumer group
  const int cgrp = (warp - kNtFirstConsumer) / gw;     // consumer group: items cgrp, cgrp + groups, ...
  const int cw = (warp - kNtFirstConsumer) % gw;
  const int VP = (int)(slot / 16u);                    // 16-byte vectors per pixel
  const int PPX = (VP + 31) / 32;                      // warp passes per pixel
  // FQ (residual variant only): the passes of a pixel divide the consumer warps, so a warp always works on the same
  // 16-byte vector of every pixel: its channels never change and the residual conv's weights (a10) live in
  // registers, packed as channel pairs for the dual-fp32 instructions.
  constexpr bool fixed_q = has_res && FQ;
  constexpr int NW = fixed_q ? L / 2 : 1;
  f32x2 rw0p[NW], rw1p[NW], rw2p[NW], rbp[NW];
  if (fixed_q) {
    const int v = (cw % PPX) * 32 + lane;
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
      const int ch = v * L + 2 * i;
      if (ch + 1 < P.C) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          a[k] = __ldg(P.rnet_w + (size_t)ch * 3 + k);
          b[k] = __ldg(P.rnet_w + (size_t)(ch + 1) * 3 + k);
        }
        a[3] = __ldg(P.rnet_b + ch);
        b[3] = __ldg(P.rnet_b + ch + 1);
      }
      rw0p[i] = pair2(a[0], b[0]);
      rw1p[i] = pair2(a[1], b[1]);
      rw2p[i] = pair2(a[2], b[2]);
      rbp[i] = pair2(a[3], b[3]);
    }
  }
  T* __restrict__ out = static_cast<T*>(P.out);
  constexpr int NF = has_res ? 1 : 2;                  // passes in flight (the residual variant holds 4*L weights in registers)
  const int ppx_shift = (PPX & (PPX - 1)) == 0 ? __ffs(PPX) - 1 : -1;   // passes per pixel is normally a power of two
  // FAST: every lane of every pass is live (a pixel is a whole number of 512-byte warp passes: C = 1024 in either type)
  // and passes per pixel is a power of two - no per-lane predicates, no zero fill, no division: the bookkeeping was
  // half of the instructions of a pass (ncu: 17 instructions per element on the warp-only bf16 variant, 7 of them math).
  auto consume = [&](auto fast_tag) {
  constexpr bool FAST = decltype(fast_tag)::value;
  int s = cgrp % S;
  unsigned ph = (unsigned)(cgrp / S) & 1u;
  while (true) {
    mbar_wait(&full[s], ph);
    const int np = desc[s].npix;
    if (np < 0) break;
    const bool byp = desc[s].bypass != 0;
    const unsigned char* st = ring + (size_t)s * Q.stage_bytes;
    T* obase = out + desc[s].out_elem;
    const int passes = np * PPX;
#pragma unroll 1
    for (int pass0 = cw; pass0 < passes; pass0 += NF * gw) {   // pass = pixel * PPX + q: q stays cw % PPX when FQ
      // NF passes of this warp in flight: all shared-memory reads first, then the arithmetic and the stores
      uint4 d[NF][6];
      NtPixW w[NF];
      bool on[NF];
      int vv[NF], gg[NF];
#pragma unroll
      for (int h = 0; h < NF; ++h) {
        const int pass = pass0 + h * gw;
        const int g = (FAST || ppx_shift >= 0) ? (pass >> ppx_shift) : pass / PPX;
        const int v = FAST ? (((pass & (PPX - 1)) << 5) + lane) : ((pass - g * PPX) * 32 + lane);
        gg[h] = g;
        vv[h] = v;
        on[h] = pass < passes && (FAST || v < VP);
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        if (!FAST) {
#pragma unroll
          for (int k = 0; k < 6; ++k) d[h][k] = z;
        } else {
          if (!has_scale) d[h][4] = z;
          if (!has_cur) d[h][5] = z;
        }
        if (on[h]) {
          w[h] = desc[s].pw[g];                        // broadcast read
          const unsigned off = (unsigned)g * slot + (unsigned)v * 16u;
          const unsigned toff = 4u * (unsigned)g * slot + (unsigned)v * 16u;
          if (!byp) {
            d[h][0] = *reinterpret_cast<const uint4*>(st + toff);
            d[h][1] = *reinterpret_cast<const uint4*>(st + toff + slot);
            d[h][2] = *reinterpret_cast<const uint4*>(st + toff + 2u * slot);
            d[h][3] = *reinterpret_cast<const uint4*>(st + toff + 3u * slot);
            if (has_scale) d[h][4] = *reinterpret_cast<const uint4*>(st + Q.off_scale + off);
          }
          if (has_cur) d[h][5] = *reinterpret_cast<const uint4*>(st + Q.off_io + off);
        }
      }
#pragma unroll
      for (int h = 0; h < NF; ++h) {
        if (!on[h]) continue;
        uint4 res4 = d[h][5];                          // ChooseFeat bypass: the current feature as is
        if (!byp) {
          const f32x2 w00p = pair2(w[h].w00, w[h].w00), w01p = pair2(w[h].w01, w[h].w01), w10p = pair2(w[h].w10, w[h].w10),
                      w11p = pair2(w[h].w11, w[h].w11), wcp = pair2(w[h].wc, w[h].wc), wwp = pair2(w[h].ww, w[h].ww);
          (void)wwp;
          f32x2 r0p[NF], r1p[NF], r2p[NF];
          r0p[h] = pair2(w[h].r0, w[h].r0);
          r1p[h] = pair2(w[h].r1, w[h].r1);
          r2p[h] = pair2(w[h].r2, w[h].r2);
          constexpr int H2 = L / 2;
          f32x2 f00[H2], f01[H2], f10[H2], f11[H2], fs[H2], fc[H2], o[H2];
          V::unpack2(d[h][0], f00);
          V::unpack2(d[h][1], f01);
          V::unpack2(d[h][2], f10);
          V::unpack2(d[h][3], f11);
          V::unpack2(d[h][4], fs);
          V::unpack2(d[h][5], fc);
#pragma unroll
          for (int i = 0; i < H2; ++i) {
            f32x2 val = mul2(w00p, f00[i]);
            val = fma2(w01p, f01[i], val);
            val = fma2(w10p, f10[i], val);
            val = fma2(w11p, f11[i], val);
            if (has_scale) val = mul2(val, fs[i]);
            if (has_res) {
              f32x2 term;
              if (fixed_q) {
                // rnet_term() on a channel pair: r = w0*r0; r = fma(w1,r1,r); r = fma(w2,r2,r); r + b  (1*r + b is exact)
                f32x2 r = mul2(rw0p[fixed_q ? i : 0], r0p[h]);
                r = fma2(rw1p[fixed_q ? i : 0], r1p[h], r);
                r = fma2(rw2p[fixed_q ? i : 0], r2p[h], r);
                term = fma2(pair2(1.0f, 1.0f), r, rbp[fixed_q ? i : 0]);
              } else {
                const int ch = vv[h] * L + 2 * i;
                const float* rw = P.rnet_w + (size_t)ch * 3;
                const float ra = rnet_term(__ldg(rw), __ldg(rw + 1), __ldg(rw + 2), __ldg(P.rnet_b + ch), w[h].r0, w[h].r1, w[h].r2);
                const float rb = rnet_term(__ldg(rw + 3), __ldg(rw + 4), __ldg(rw + 5), __ldg(P.rnet_b + ch + 1), w[h].r0, w[h].r1, w[h].r2);
                term = pair2(ra, rb);
              }
              val = fma2(wwp, term, val);
            }
            o[i] = has_cur ? fma2(wcp, fc[i], val) : val;
          }
          res4 = V::pack2(o);
        }
        stg_stream_v4(obase + (size_t)gg[h] * P.C + (size_t)vv[h] * L, res4);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&done[s]);              // every shared-memory read of the stage is complete
    s += Q.groups;
    if (s >= S) {                                      // S >= groups: at most one wrap per step
      s -= S;
      ph ^= 1u;
    }
  }
  };
  if (ppx_shift >= 0 && VP == PPX * 32) consume(std::true_type{});
  else consume(std::false_type{});
}

// Can the all-TMA channels-last kernel serve these arguments?  (compile-time variant, no cosine, req = write,
// a pixel's channel run of at most 8 KB.)
inline bool plan_nhwc_tma(const AggParams& P, bool bf16, int var, NtPlan* Q) {
  if (var == kVarRuntime || P.mode == LSFA_W_COSINE || P.req_add) return false;
  const size_t slot = (size_t)P.C * (bf16 ? 2 : 4);
  if (slot % 16 || slot > 8192) return false;
  const bool has_scale = var == kVarScale || var == kVarScaleCur;
  const bool has_cur = var == kVarScaleCur || var == kVarResCur;
  const unsigned ns = 4u + (has_scale ? 1u : 0u) + (has_cur ? 1u : 0u);   // taps + scale + cur
  unsigned g = kNtStageBudget / (ns * (unsigned)slot);
  if (g < 1) return false;
  int G = 1;
  while (G * 2 <= (int)g && G * 2 <= kNtMaxG) G *= 2;
  if (const char* e = knob("LSFA_NT_G")) {                   // experiment knob: smaller pixel groups, more stages
    const int want = atoi(e);
    while (G > 1 && G > want) G /= 2;
  }
  Q->G = G;
  Q->merge_pairs = knob("LSFA_NT_NO_MERGE") ? 0 : 1;
  Q->groups = var == kVarResCur ? 1 : 2;
  if (const char* e = knob("LSFA_NT_GROUPS")) Q->groups = atoi(e) == 1 ? 1 : 2;
  Q->slot = (unsigned)slot;
  Q->off_scale = 4u * (unsigned)G * (unsigned)slot;
  Q->off_io = Q->off_scale + (has_scale ? (unsigned)G * (unsigned)slot : 0u);
  Q->stage_bytes = Q->off_io + (has_cur ? (unsigned)G * (unsigned)slot : 0u);
  int stages = (int)(kNtRingBudget / Q->stage_bytes);
  if (stages > kNtMaxStages) stages = kNtMaxStages;
  if (const char* e = knob("LSFA_NT_STAGES")) stages = max(2, min(stages, atoi(e)));
  stages -= stages % kNtStageMultiple;
  if (stages < kNtStageMultiple) return false;
  Q->stages = stages;
  Q->smem = (size_t)kNtOffRing + (size_t)stages * Q->stage_bytes;
  return Q->smem <= 227u * 1024u;
}

template <typename T, int VAR, bool FQ>
cudaError_t launch_nhwc_tma_fq(const AggParams& P, const NtPlan& Q, int grid, cudaStream_t st) {
  auto kfn = agg_nhwc_tma_kernel<T, VAR, FQ>;
  cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Q.smem);
  if (e != cudaSuccess) return e;
  kfn<<<(unsigned)grid, kNtThreads, Q.smem, st>>>(P, Q);
  return cudaPeekAtLastError();
}

template <typename T, int VAR>
cudaError_t launch_nhwc_tma_variant(const AggParams& P, const NtPlan& Q, int grid, cudaStream_t st) {
  if constexpr (VAR == kVarResCur) {
    const int ppx = ((int)(Q.slot / 16u) + 31) / 32;            // warp passes per pixel
    const int L = (int)(16 / sizeof(T));
    const int gw = kNtAllConsumerWarps / Q.groups;
    if (ppx <= gw && gw % ppx == 0 && P.C % (2 * L) == 0)
      return launch_nhwc_tma_fq<T, VAR, true>(P, Q, grid, st);
  }
  return launch_nhwc_tma_fq<T, VAR, false>(P, Q, grid, st);
}

}  // namespace lsfa
