// agg_nhwc_win_kernel - channels-last, window-resident form of the fused operator (tensor-map TMA).
//
// Why a third channels-last kernel.  The gather-by-bulk-copy kernel (aggregate_nhwc_tma.cuh) fetches the four taps of
// every output pixel from L2: 4 x the key bytes per frame, 10.7-11.6 TB/s of L2 -> SM traffic for every channels-last
// variant (profiles/r2_ncu_nhwc_variants_l2.txt), close to what the L2 slices deliver.  This kernel was built to test
// whether that is what holds the bf16 warp-only / shipped-path variants at 0.73 / 0.76 of the copy peak: neighbouring
// output pixels sample neighbouring key pixels, so it loads ONE window of key pixels per 8x8 output tile and gathers
// from shared memory.  Outcome: L2 -> SM traffic 1.27 -> 0.81 GB per 64 frames, bit-identical results,
// the SAME time - L2 was not the limiter (the consumers' instruction stream is) - so the launcher never picks this
// kernel by itself; it is reachable with force_generic = 5 and kept as the lower-traffic form.
//
//   work item  = (frame, 8x8 output tile), channel chunks of 256 bytes per pixel streamed through a stage ring
//   window     = 12 x 12 key pixels x 256 B, ONE cp.async.bulk.tensor.4d per stage from a (C, Wk, Hk, keys) tensor map:
//                pixels outside the plane are zero-filled by the copy engine (the sampler's zero padding for free), and
//                the window origin is the tile's tap bounding box (found by the record warp).  12 = 8 + 4: motion of up to
//                ~2 cells inside a tile.  L2 traffic per output pixel: 144/64 = 2.25 key pixels instead of 4.
//   scale, cur = 8 x 8 x 256 B boxes of their own tensor maps (tiles that overhang the plane are zero-filled)
//   fallback   = a tile whose taps do not fit the window becomes two GATHER items of 32 pixels, each pixel with its own
//                2x2 box (a second tensor map): the same consumer code, tap rows 2 slots apart instead of 12.
//
//   warp 0   record warp : per tile the index chain of its 64 pixels (two per lane; same functions as every other kernel:
//                          issue_pixel_loads / finish_pixel), the bounding box, the item header and 64 records into a ring
//   warp 1   producer    : per item and channel chunk: expect_tx + the tensor copies into the next free stage
//   warps 2+ consumers   : 16 warps; lane = (pixel parity, 16-byte vector of the chunk); a lane keeps its two pixels'
//                          records in registers for all chunks of the item; 4 (+2) LDS.128, packed fp32 math, one
//                          16-byte streaming store (a pixel's 256-byte chunk is written by 16 consecutive lanes)
// Arithmetic per element is the expression chain of agg_nhwc_kernel: bit-identical results.
#pragma once
#include <cuda.h>

#include "aggregate_nhwc_tma.cuh"

namespace lsfa {

constexpr int kWinT = 8;                 // output tile edge
constexpr int kWinB = 12;                // window edge (key pixels)
constexpr int kWinChunk = 256;           // bytes of one pixel's channel chunk in a stage
constexpr int kWinVecs = kWinChunk / 16; // 16-byte vectors per pixel chunk
constexpr int kWinConsumerWarps = 16;
constexpr int kWinThreads = (2 + kWinConsumerWarps) * 32;
constexpr int kWinRecRing = 4;
constexpr int kWinMaxStages = 6;
constexpr unsigned kWinWindowBytes = kWinB * kWinB * kWinChunk;   // 36864
constexpr unsigned kWinTileBytes = kWinT * kWinT * kWinChunk;     // 16384
constexpr unsigned kWinRecBytes = 48;
constexpr unsigned kWinSlotBytes = 64 + 64 * kWinRecBytes;        // header + 64 records
constexpr unsigned kWinOffDesc = 192;                             // after 24 mbarriers
constexpr unsigned kWinOffSlots = 320;
constexpr unsigned kWinOffRing = (kWinOffSlots + kWinRecRing * kWinSlotBytes + 127u) / 128u * 128u;

struct WinPlan {
  int stages;
  int nchunks;               // channel chunks per pixel (C * sizeof(T) / 256)
  unsigned stage_bytes, off_scale, off_cur;
  unsigned off_rnet;         // residual variant: [C/2] x 32 B pair table after the ring
  size_t smem;
};

struct WinHdr {              // 64 B
  int n, kn, ox, oy;
  int kind;                  // 0 window, 1 gather, < 0 stop
  int wx0, wy0;              // window origin (key pixels; may be -1)
  int bypass;
  int pad[8];
};

__device__ __forceinline__ void win_tma_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

template <typename T, int VAR>
__global__ void __launch_bounds__(kWinThreads, 1)
agg_nhwc_win_kernel(const __grid_constant__ AggParams P, const WinPlan Q, const __grid_constant__ CUtensorMap m_key,
                    const __grid_constant__ CUtensorMap m_key2, const __grid_constant__ CUtensorMap m_scale,
                    const __grid_constant__ CUtensorMap m_cur) {
  static_assert(VAR != kVarRuntime, "only the compile-time variants");
  constexpr bool has_scale = VAR == kVarScale || VAR == kVarScaleCur;
  constexpr bool has_cur = VAR == kVarScaleCur || VAR == kVarResCur;
  constexpr bool has_res = VAR == kVarResCur;
  using V = Vec16<T>;
  constexpr int L = V::kLanes;                       // elements per 16-byte vector
  constexpr int CCE = kWinChunk / (int)sizeof(T);    // elements per pixel chunk
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* done = full + 8;
  uint64_t* rec_full = done + 8;
  uint64_t* rec_free = rec_full + kWinRecRing;
  volatile int4* desc = reinterpret_cast<volatile int4*>(smem_raw + kWinOffDesc);   // (slot | -1, chunk, last, 0)
  unsigned char* slots = smem_raw + kWinOffSlots;
  unsigned char* ring = smem_raw + kWinOffRing;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int S = Q.stages;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&done[s], kWinConsumerWarps);
    }
    for (int s = 0; s < kWinRecRing; ++s) {
      mbar_init(&rec_full[s], 1);
      mbar_init(&rec_free[s], kWinConsumerWarps);
    }
    fence_barrier_init();
  }
  if (has_res) {
    // rnet_conv0 (SYM:66) as channel pairs {w0a,w0b, w1a,w1b} {w2a,w2b, ba,bb}, laid out [chunk][pair i of a vector][part][vector]:
    // the 16 lanes of a half warp read 16 consecutive 16-byte words (the natural [channel pair] order puts them 128 bytes
    // apart: a 16-way bank conflict, measured 3.5x on the whole kernel)
    uint4* tab = reinterpret_cast<uint4*>(smem_raw + Q.off_rnet);
    for (int j = tid; j < P.C / 2; j += kWinThreads) {
      const int ch = 2 * j, c = ch / CCE, rem = ch - c * CCE, vv = rem / L, i = (rem - vv * L) / 2;
      const float* wa = P.rnet_w + (size_t)ch * 3;
      const size_t at = (((size_t)c * (L / 2) + i) * 2) * kWinVecs + vv;
      tab[at] = make_uint4(__float_as_uint(__ldg(wa + 0)), __float_as_uint(__ldg(wa + 3)), __float_as_uint(__ldg(wa + 1)),
                           __float_as_uint(__ldg(wa + 4)));
      tab[at + kWinVecs] = make_uint4(__float_as_uint(__ldg(wa + 2)), __float_as_uint(__ldg(wa + 5)),
                                      __float_as_uint(__ldg(P.rnet_b + ch)), __float_as_uint(__ldg(P.rnet_b + ch + 1)));
    }
  }
  __syncthreads();

  const int tiles_x = (P.W + kWinT - 1) / kWinT, tiles_y = (P.H + kWinT - 1) / kWinT;
  const int tpf = tiles_x * tiles_y;
  const long long ntiles = (long long)P.N * tpf;

  if (warp == 0) {
    // ===================================== record warp =====================================
    int rs = 0;
    long long emitted = 0;
    auto acquire = [&]() {
      if (emitted >= kWinRecRing) mbar_wait(&rec_free[rs], (unsigned)((emitted / kWinRecRing) - 1) & 1u);
    };
    auto publish = [&]() {
      __syncwarp();
      if (lane == 0) mbar_arrive(&rec_full[rs]);       // release: header + records visible
      if (++rs == kWinRecRing) rs = 0;
      ++emitted;
    };
    for (long long it = 0;; ++it) {
      long long tl;
      if (P.sched != nullptr) {
        unsigned got = 0;
        if (lane == 0) got = atomicAdd(P.sched, 1u);
        tl = (long long)__shfl_sync(0xffffffffu, got, 0);
      } else {
        tl = blockIdx.x + it * (long long)gridDim.x;
      }
      if (tl >= ntiles) {
        acquire();
        if (lane == 0) reinterpret_cast<WinHdr*>(slots + rs * kWinSlotBytes)->kind = -1;
        publish();
        break;
      }
      const int n = (int)(tl / tpf);
      const int r = (int)(tl - (long long)n * tpf);
      const int ty = r / tiles_x, tx = r - ty * tiles_x;
      const int ox = tx * kWinT, oy = ty * kWinT;
      const bool byp = P.bypass != nullptr && __ldg(P.bypass + n) != 0;
      // two pixels per lane: idx = h * 32 + lane -> (py, px) = (idx / 8, idx % 8)
      PixelLoads ld[2];
      bool val[2];
      float r0[2] = {0.f, 0.f}, r1[2] = {0.f, 0.f}, r2[2] = {0.f, 0.f};
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int idx = h * 32 + lane;
        const int x = ox + (idx & 7), y = oy + (idx >> 3);
        val[h] = x < P.W && y < P.H;
        if (val[h] && !byp) {
          ld[h] = issue_pixel_loads(P, n, y, x);
          if (has_res) {
            const int p = y * P.W + x;
            r0[h] = __ldg(P.res + ((size_t)n * 3 + 0) * P.HW + p);
            r1[h] = __ldg(P.res + ((size_t)n * 3 + 1) * P.HW + p);
            r2[h] = __ldg(P.res + ((size_t)n * 3 + 2) * P.HW + p);
          }
        }
      }
      PixelRec t[2];
      int x0c[2] = {0, 0}, y0c[2] = {0, 0};
      int mnx = 1 << 20, mny = 1 << 20, mxx = -(1 << 20), mxy = -(1 << 20);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        t[h].w00 = t[h].w01 = t[h].w10 = t[h].w11 = t[h].wc = t[h].ww = 0.f;
        if (val[h] && !byp) {
          const int idx = h * 32 + lane;
          t[h] = finish_pixel(P, ld[h], n, oy + (idx >> 3), ox + (idx & 7), /*fold=*/true);
          // top-left tap, clamped to one pixel outside the plane: beyond that every weight is zero and any readable
          // (zero-filled) position will do
          const int ya = t[h].i00 / P.Wk, xa = t[h].i00 - ya * P.Wk;
          x0c[h] = (t[h].edge & 1) ? -1 : xa;
          y0c[h] = (t[h].edge & 4) ? -1 : ya;
          mnx = min(mnx, x0c[h]); mxx = max(mxx, x0c[h]);
          mny = min(mny, y0c[h]); mxy = max(mxy, y0c[h]);
        }
      }
      mnx = __reduce_min_sync(0xffffffffu, mnx); mny = __reduce_min_sync(0xffffffffu, mny);
      mxx = __reduce_max_sync(0xffffffffu, mxx); mxy = __reduce_max_sync(0xffffffffu, mxy);
      const bool any = mxx >= mnx;
      const bool fits = !any || (mxx - mnx + 2 <= kWinB && mxy - mny + 2 <= kWinB);
      const int kn = key_slot(P, n);
      auto write_rec = [&](unsigned char* slot, int pos, int h, unsigned tapbase, unsigned rstride, bool ok) {
        uint4* q = reinterpret_cast<uint4*>(slot + 64 + (size_t)pos * kWinRecBytes);
        const int idx = h * 32 + lane;
        q[0] = make_uint4(__float_as_uint(t[h].w00), __float_as_uint(t[h].w01), __float_as_uint(t[h].w10), __float_as_uint(t[h].w11));
        q[1] = make_uint4(__float_as_uint(t[h].wc), __float_as_uint(t[h].ww), __float_as_uint(r0[h]), __float_as_uint(r1[h]));
        q[2] = make_uint4(__float_as_uint(r2[h]), tapbase | (rstride << 16),
                          (unsigned)(idx & 7) | ((unsigned)(idx >> 3) << 8) | (ok ? 0x10000u : 0u),
                          ((unsigned)x0c[h] & 0xffffu) | ((unsigned)y0c[h] << 16));
      };
      auto write_hdr = [&](unsigned char* slot, int kind, int wx0, int wy0) {
        if (lane == 0) {
          WinHdr* hd = reinterpret_cast<WinHdr*>(slot);
          hd->n = n; hd->kn = kn; hd->ox = ox; hd->oy = oy;
          hd->kind = kind; hd->wx0 = wx0; hd->wy0 = wy0; hd->bypass = byp ? 1 : 0;
        }
      };
      if (byp || fits) {
        acquire();
        unsigned char* slot = slots + rs * kWinSlotBytes;
        const int wx0 = any ? mnx : 0, wy0 = any ? mny : 0;
        write_hdr(slot, 0, wx0, wy0);
#pragma unroll
        for (int h = 0; h < 2; ++h)
          write_rec(slot, h * 32 + lane, h, (unsigned)((y0c[h] - wy0) * kWinB + (x0c[h] - wx0)), (unsigned)kWinB, val[h]);
        publish();
      } else {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (!__any_sync(0xffffffffu, val[h])) continue;
          acquire();
          unsigned char* slot = slots + rs * kWinSlotBytes;
          write_hdr(slot, 1, 0, 0);
          write_rec(slot, lane, h, (unsigned)(lane * 4), 2u, val[h]);
          reinterpret_cast<uint4*>(slot + 64 + (size_t)(32 + lane) * kWinRecBytes)[2] = make_uint4(0u, 0u, 0u, 0u);   // not valid
          publish();
        }
      }
    }
    return;
  }

  if (warp == 1) {
    // ===================================== producer warp =====================================
    long long f = 0;                                    // stages filled so far
    int rs = 0;
    for (long long item = 0;; ++item) {
      mbar_wait(&rec_full[rs], (unsigned)(item / kWinRecRing) & 1u);
      const unsigned char* slot = slots + rs * kWinSlotBytes;
      const WinHdr* hd = reinterpret_cast<const WinHdr*>(slot);
      const int kind = hd->kind;
      if (kind < 0) break;
      const int n = hd->n, kn = hd->kn, ox = hd->ox, oy = hd->oy, wx0 = hd->wx0, wy0 = hd->wy0;
      const bool byp = hd->bypass != 0;
      // gather items: this lane's pixel (list position = lane)
      const uint4 rc = reinterpret_cast<const uint4*>(slot + 64 + (size_t)lane * kWinRecBytes)[2];
      const bool gok = kind == 1 && (rc.z & 0x10000u) != 0;
      const int gx = (int)(short)(rc.w & 0xffffu), gy = (int)(short)(rc.w >> 16);
      const unsigned ng = __popc(__ballot_sync(0xffffffffu, gok));
      for (int c = 0; c < Q.nchunks; ++c, ++f) {
        const int s = (int)(f % S);
        if (f >= S) mbar_wait(&done[s], (unsigned)((f / S) - 1) & 1u);   // the consumers are done with the stage's previous use
        unsigned char* st = ring + (size_t)s * Q.stage_bytes;
        if (lane == 0) {
          desc[s].x = rs;
          desc[s].y = c;
          desc[s].z = (c == Q.nchunks - 1) ? 1 : 0;
          unsigned bytes = has_cur ? kWinTileBytes : 0u;
          if (!byp) bytes += (kind == 0 ? kWinWindowBytes : ng * 4u * kWinChunk) + (has_scale ? kWinTileBytes : 0u);
          mbar_expect_tx(&full[s], bytes);               // release: the descriptor is visible with the data
        }
        __syncwarp();
        if (!byp) {
          if (kind == 0) {
            if (lane == 0) win_tma_4d(st, &m_key, c * CCE, wx0, wy0, kn, &full[s]);
          } else if (gok) {
            win_tma_4d(st + (size_t)lane * 4 * kWinChunk, &m_key2, c * CCE, gx, gy, kn, &full[s]);
          }
          if (has_scale && lane == 1) win_tma_4d(st + Q.off_scale, &m_scale, c * CCE, ox, oy, n, &full[s]);
        }
        if (has_cur && lane == 2) win_tma_4d(st + Q.off_cur, &m_cur, c * CCE, ox, oy, n, &full[s]);
      }
      if (++rs == kWinRecRing) rs = 0;
    }
    {   // stop marker in the next stage
      const int s = (int)(f % S);
      if (f >= S) mbar_wait(&done[s], (unsigned)((f / S) - 1) & 1u);
      if (lane == 0) {
        desc[s].x = -1;
        mbar_arrive(&full[s]);
      }
    }
    return;
  }

  // ===================================== consumer warps =====================================
  const int cw = warp - 2;
  const int v = lane & 15, hsel = lane >> 4;
  T* __restrict__ out = static_cast<T*>(P.out);
  float w00[2], w01[2], w10[2], w11[2], wc[2], ww[2], r0[2], r1[2], r2[2];
  unsigned tap0[2], tap1[2], pixoff[2];                // byte offsets inside a stage: top tap row, bottom tap row, pixel's tile slot
  bool ok[2] = {false, false};
  size_t oelem[2] = {0, 0};
  bool byp = false;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    w00[t] = w01[t] = w10[t] = w11[t] = wc[t] = ww[t] = r0[t] = r1[t] = r2[t] = 0.f;
    tap0[t] = tap1[t] = pixoff[t] = 0u;
  }
  (void)ww; (void)r0; (void)r1; (void)r2;
  for (long long f = 0;; ++f) {
    const int s = (int)(f % S);
    mbar_wait(&full[s], (unsigned)(f / S) & 1u);
    const int rs = desc[s].x;
    if (rs < 0) break;
    const int c = desc[s].y;
    const bool last = desc[s].z != 0;
    if (c == 0) {   // new item: this lane's two pixels
      const unsigned char* slot = slots + rs * kWinSlotBytes;
      const WinHdr* hd = reinterpret_cast<const WinHdr*>(slot);
      const int n = hd->n, ox = hd->ox, oy = hd->oy;
      byp = hd->bypass != 0;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int idx = t * 32 + cw * 2 + hsel;
        const uint4* q = reinterpret_cast<const uint4*>(slot + 64 + (size_t)idx * kWinRecBytes);
        const uint4 a = q[0], b = q[1], cc = q[2];
        w00[t] = __uint_as_float(a.x); w01[t] = __uint_as_float(a.y); w10[t] = __uint_as_float(a.z); w11[t] = __uint_as_float(a.w);
        wc[t] = __uint_as_float(b.x); ww[t] = __uint_as_float(b.y); r0[t] = __uint_as_float(b.z); r1[t] = __uint_as_float(b.w);
        r2[t] = __uint_as_float(cc.x);
        const unsigned tapbase = cc.y & 0xffffu, rstride = cc.y >> 16;
        const int px = (int)(cc.z & 0xffu), py = (int)((cc.z >> 8) & 0xffu);
        ok[t] = (cc.z & 0x10000u) != 0;
        tap0[t] = tapbase * kWinChunk + (unsigned)v * 16u;
        tap1[t] = (tapbase + rstride) * kWinChunk + (unsigned)v * 16u;
        pixoff[t] = (unsigned)(py * kWinT + px) * kWinChunk + (unsigned)v * 16u;
        oelem[t] = ((size_t)n * P.HW + (size_t)(oy + py) * P.W + (size_t)(ox + px)) * P.C + (size_t)v * L;
      }
    }
    const unsigned char* st = ring + (size_t)s * Q.stage_bytes;
    uint4 d[2][6];
#pragma unroll
    for (int t = 0; t < 2; ++t) {        // all shared-memory reads of both pixels first
      const uint4 z = make_uint4(0u, 0u, 0u, 0u);
      d[t][4] = z;
      d[t][5] = z;
      if (ok[t]) {
        if (!byp) {
          d[t][0] = *reinterpret_cast<const uint4*>(st + tap0[t]);
          d[t][1] = *reinterpret_cast<const uint4*>(st + tap0[t] + kWinChunk);
          d[t][2] = *reinterpret_cast<const uint4*>(st + tap1[t]);
          d[t][3] = *reinterpret_cast<const uint4*>(st + tap1[t] + kWinChunk);
          if (has_scale) d[t][4] = *reinterpret_cast<const uint4*>(st + Q.off_scale + pixoff[t]);
        }
        if (has_cur) d[t][5] = *reinterpret_cast<const uint4*>(st + Q.off_cur + pixoff[t]);
      }
    }
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      if (!ok[t]) continue;
      uint4 res4 = d[t][5];                              // ChooseFeat bypass: the current feature as is
      if (!byp) {
        const f32x2 w00p = pair2(w00[t], w00[t]), w01p = pair2(w01[t], w01[t]), w10p = pair2(w10[t], w10[t]),
                    w11p = pair2(w11[t], w11[t]), wcp = pair2(wc[t], wc[t]), wwp = pair2(ww[t], ww[t]);
        (void)wwp;
        constexpr int H2 = L / 2;
        f32x2 f00[H2], f01[H2], f10[H2], f11[H2], fs[H2], fc[H2], o[H2];
        V::unpack2(d[t][0], f00);
        V::unpack2(d[t][1], f01);
        V::unpack2(d[t][2], f10);
        V::unpack2(d[t][3], f11);
        V::unpack2(d[t][4], fs);
        V::unpack2(d[t][5], fc);
#pragma unroll
        for (int i = 0; i < H2; ++i) {
          f32x2 val = mul2(w00p, f00[i]);
          val = fma2(w01p, f01[i], val);
          val = fma2(w10p, f10[i], val);
          val = fma2(w11p, f11[i], val);
          if (has_scale) val = mul2(val, fs[i]);
          if (has_res) {
            // rnet_term() on a channel pair: r = w0*r0; r = fma(w1,r1,r); r = fma(w2,r2,r); r + b  (1*r + b is exact)
            const uint4* tp = reinterpret_cast<const uint4*>(smem_raw + Q.off_rnet) + ((size_t)c * H2 + i) * 2 * kWinVecs + v;
            const uint4 ta = tp[0], tb = tp[kWinVecs];
            f32x2 r = mul2(pair2(__uint_as_float(ta.x), __uint_as_float(ta.y)), pair2(r0[t], r0[t]));
            r = fma2(pair2(__uint_as_float(ta.z), __uint_as_float(ta.w)), pair2(r1[t], r1[t]), r);
            r = fma2(pair2(__uint_as_float(tb.x), __uint_as_float(tb.y)), pair2(r2[t], r2[t]), r);
            const f32x2 term = fma2(pair2(1.0f, 1.0f), r, pair2(__uint_as_float(tb.z), __uint_as_float(tb.w)));
            val = fma2(wwp, term, val);
          }
          o[i] = has_cur ? fma2(wcp, fc[i], val) : val;
        }
        res4 = V::pack2(o);
      }
      stg_stream_v4(out + oelem[t] + (size_t)c * CCE, res4);
    }
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(&done[s]);                             // every shared-memory read of the stage is complete
      if (last) mbar_arrive(&rec_free[rs]);              // ... and of the item's records
    }
  }
}

// ---- host side ------------------------------------------------------------------------------------------
inline bool plan_nhwc_win(const AggParams& P, bool bf16, int var, WinPlan* Q) {
  if (var == kVarRuntime || P.mode == LSFA_W_COSINE || P.req_add) return false;
  const size_t es = bf16 ? 2 : 4;
  if (((size_t)P.C * es) % kWinChunk) return false;
  if (P.Hk < 1 || P.Wk < 1 || P.Wk > 32767 || P.Hk > 32767) return false;
  const bool has_scale = var == kVarScale || var == kVarScaleCur;
  const bool has_cur = var == kVarScaleCur || var == kVarResCur;
  Q->nchunks = (int)((size_t)P.C * es / kWinChunk);
  Q->off_scale = kWinWindowBytes;
  Q->off_cur = Q->off_scale + (has_scale ? kWinTileBytes : 0u);
  Q->stage_bytes = Q->off_cur + (has_cur ? kWinTileBytes : 0u);
  const size_t rnet = var == kVarResCur ? (size_t)(P.C / 2) * 32 : 0;
  long long stages = ((long long)227 * 1024 - kWinOffRing - (long long)rnet) / Q->stage_bytes;
  if (stages > kWinMaxStages) stages = kWinMaxStages;
  if (stages < 2) return false;
  Q->stages = (int)stages;
  Q->off_rnet = kWinOffRing + (unsigned)stages * Q->stage_bytes;
  Q->smem = (size_t)Q->off_rnet + rnet;
  return true;
}

}  // namespace lsfa
