// Upstream of the path: the accumulated motion-vector field and
// the residual that coviar hands to get_image().  GPU restatement of
// external/data_loader_py2/coviar_data_loader.c:71-177 (create_and_load_mv_residual) with the
// initialisation of :318-328, for N independent GOPs at once.  Integer work, bit-exact.
//
// Reference semantics per decoded P-frame t = 1..T (sequential - frame t reads frame t-1):
//     accu_new = accu_old                                  (accu starts as the identity (x,y))
//     for every motion vector i IN LIST ORDER with dst != src:
//         for every pixel of its w x h block, if the dst AND the src pixel are inside the frame:
//             accu_new[dst] = accu_old[src]                 (a later vector overwrites an earlier one)
// and at the target frame  mv[y][x] = (x,y) - accu[y][x],  res = cur[y][x] - iframe[accu[y][x]].
//
// Parallel form: (A) every (vector, block pixel) thread does atomicMax(owner[dst], i+1): the
// winner is exactly the vector the sequential loop would have left there; (B) every pixel
// gathers accu_old at the source its owner dictates (or keeps its value) and clears its owner.
#include "lsfa_device.cuh"

namespace lsfa {

struct MvRec {  // one AVMotionVector, the six fields the reference reads
  int w, h, src_x, src_y, dst_x, dst_y;
};

// The accumulated field holds pixel coordinates: 16 bits per component are enough for frames up to 32767 x 32767
// (A = short2: half the bytes of the reference's int pairs through HBM); larger frames use A = int2.
template <typename A>
__device__ __forceinline__ A make_xy(int x, int y);
template <>
__device__ __forceinline__ int2 make_xy<int2>(int x, int y) { return make_int2(x, y); }
template <>
__device__ __forceinline__ short2 make_xy<short2>(int x, int y) { return make_short2((short)x, (short)y); }

// The per-pixel kernels run on a (width/256, height, N) grid: x, y and the GOP come from the block and thread indices
// (a flat 64-bit index cost a 64-bit division and two 32-bit ones per pixel, which made the gather instruction bound).
template <typename A>
__global__ void mvacc_init_kernel(A* __restrict__ accu, int* __restrict__ owner, int height, int width) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= width) return;
  const size_t i = ((size_t)blockIdx.z * height + y) * width + x;
  accu[i] = make_xy<A>(x, y);     // coviar_data_loader.c:322-326
  owner[i] = 0;
}

// tag = (t+1) << 24 when the owner map is never cleared (a later frame's entries are larger than any stale one: one
// write and one pass over the map less per P-frame), 0 when the gather clears it (T > 126 or M >= 2^24 vectors)
__global__ void mvacc_owner_kernel(const MvRec* __restrict__ mvs, const int* __restrict__ counts, int* __restrict__ owner,
                                   int t, int T, int M, int height, int width, int tag) {
  const int n = blockIdx.z;
  const int cnt = min(__ldg(counts + (size_t)n * T + t), M);
  const int i = blockIdx.y * blockDim.y + threadIdx.y;          // vector index
  if (i >= cnt) return;
  const MvRec mv = mvs[((size_t)n * T + t) * M + i];
  if (mv.dst_x - mv.src_x == 0 && mv.dst_y - mv.src_y == 0) return;      // :92
  const int x_lo = (-1 * mv.w) / 2, x_hi = mv.w / 2, y_lo = (-1 * mv.h) / 2, y_hi = mv.h / 2;   // :97-98 (C division)
  const int bw = x_hi - x_lo, bh = y_hi - y_lo;
  if (bw <= 0 || bh <= 0) return;
  int* own = owner + (size_t)n * height * width;
  const int val = tag + i + 1;
  // lanes run along x: 16 neighbouring pixels per 64-byte segment (the reference walks x outer, y inner; the order is
  // irrelevant here).  Block widths are powers of two in practice (16, 8, 4): shift and mask instead of a division.
  const int sh = (bw & (bw - 1)) == 0 ? __ffs(bw) - 1 : -1;
  for (int q = threadIdx.x; q < bw * bh; q += blockDim.x) {
    const int qy = sh >= 0 ? (q >> sh) : q / bw;
    const int ys = y_lo + qy, xs = x_lo + (q - qy * bw);
    const int dx = mv.dst_x + xs, dy = mv.dst_y + ys, sx = mv.src_x + xs, sy = mv.src_y + ys;
    // one unsigned compare per bound: negative coordinates wrap to large values
    if ((unsigned)dy < (unsigned)height && (unsigned)dx < (unsigned)width && (unsigned)sy < (unsigned)height &&
        (unsigned)sx < (unsigned)width)
      atomicMax(own + dy * width + dx, val);
  }
}

// (B) gather (+ clear owner when the entries are not tagged).  Every pixel is a chain of three dependent loads
// (owner -> that vector's src/dst -> accu_old at the source pixel) and the kernel is latency bound (one pixel per
// thread: 2.5 TB/s of traffic at full occupancy), so a thread walks kGatherRows rows of its column at once, each round
// of loads issued for all of them before any is used (measured per P-frame, 64 GOPs at 720p: 1 row 303 us, 4 rows
// 182 us, 8 rows 229 us).
#ifndef LSFA_GATHER_ROWS
#define LSFA_GATHER_ROWS 4
#endif
constexpr int kGatherRows = LSFA_GATHER_ROWS;

template <typename A>
__global__ void __launch_bounds__(256)
mvacc_gather_kernel(const MvRec* __restrict__ mvs, const A* __restrict__ accu_old, A* __restrict__ accu_new,
                    int* __restrict__ owner, int t, int T, int M, int height, int width, int tag) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y0 = blockIdx.y * kGatherRows, n = blockIdx.z;
  if (x >= width) return;
  // per-frame base pointers + 32-bit offsets inside the frame (height * width < 2^31 is checked by the entry point):
  // the 64-bit index arithmetic of a flat layout made this kernel issue bound (ncu: issue_active 79 %)
  const size_t frame = (size_t)n * height * width;
  const A* __restrict__ ao = accu_old + frame;
  A* __restrict__ an = accu_new + frame;
  int* __restrict__ ow = owner + frame;
  const int2* __restrict__ list = reinterpret_cast<const int2*>(mvs + ((size_t)n * T + t) * M);   // 3 x int2 per vector
  const int rows = min(kGatherRows, height - y0);
  const int i0 = y0 * width + x;
  const int tagv = tag >> 24;
  int o[kGatherRows];
  A v[kGatherRows];
#pragma unroll
  for (int u = 0; u < kGatherRows; ++u) {                // round 1: owner + own value
    o[u] = 0;
    if (u < rows) {
      o[u] = ow[i0 + u * width];
      v[u] = ao[i0 + u * width];
    }
  }
  int2 ms[kGatherRows], md[kGatherRows];
#pragma unroll
  for (int u = 0; u < kGatherRows; ++u) {                // round 2: the owning vector's src and dst
    if (tag) o[u] = (o[u] >> 24) == tagv ? (o[u] & 0xffffff) : 0;       // entries of earlier frames are stale
    ms[u] = md[u] = make_int2(0, 0);
    if (o[u] > 0) {
      const int2* rec = list + 3 * (o[u] - 1);
      ms[u] = __ldg(rec + 1);                            // (src_x, src_y)
      md[u] = __ldg(rec + 2);                            // (dst_x, dst_y)
    }
  }
#pragma unroll
  for (int u = 0; u < kGatherRows; ++u)                  // round 3: the gather proper
    if (o[u] > 0) {
      const int sx = ms[u].x + (x - md[u].x), sy = ms[u].y + (y0 + u - md[u].y);
      v[u] = ao[sy * width + sx];
    }
#pragma unroll
  for (int u = 0; u < kGatherRows; ++u)
    if (u < rows) {
      an[i0 + u * width] = v[u];
      if (!tag && o[u] > 0) ow[i0 + u * width] = 0;
    }
}

// target frame: mv = (x,y) - accu   (:130-139), layout (N,height,width,2) like the numpy array coviar returns
template <typename A>
__global__ void mvacc_finish_kernel(const A* __restrict__ accu, int2* __restrict__ mv_out, int height, int width) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= width) return;
  const size_t i = ((size_t)blockIdx.z * height + y) * width + x;
  const A a = accu[i];
  mv_out[i] = make_int2(x - (int)a.x, y - (int)a.y);
}

// residual (:141-175, accumulate case): res[y][x][c] = cur[y][x][c] - iframe[src_y][src_x][c], src = (x,y) - mv
__global__ void coviar_residual_kernel(const unsigned char* __restrict__ iframe, const unsigned char* __restrict__ cur,
                                       const int2* __restrict__ mv, int* __restrict__ res, int height, int width) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= width) return;
  const size_t frame = (size_t)blockIdx.z * height * width;
  const size_t i = frame + (size_t)y * width + x;
  const int2 m = mv[i];
  const int sx = x - m.x, sy = y - m.y;                  // always in bounds for an accumulated field
  const size_t src = (frame + (size_t)sy * width + sx) * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) res[i * 3 + c] = (int)cur[i * 3 + c] - (int)iframe[src + c];
}


// ---------------------------------------------------------------------------------------------------------
// Round 2: the back-trace without a per-frame field.  accu_t = accu_{t-1} o s_t with accu_0 = identity, so
//     accu_T(p) = s_1(s_2(... s_T(p) ...)),   s_t(p) = p + (src - dst) of the LAST vector of frame t (list order) whose
//                                              block covers p with p and its source inside the frame, else p:
// a chain of T look-ups per pixel, walked from the target frame back to the I-frame.  "Which vector owns pixel p in frame
// t" is answered from a CELL index instead of a per-pixel owner map: per 8x8 cell the largest index among the vectors
// whose destination block touches the cell (one atomicMax per vector and touched cell: 57 KB per 720p frame, L2-resident).
// Any vector covering p touches p's cell, so the owner's index is <= that maximum: the pixel tests the cell's top vector
// first and walks down the list only if that one does not cover it (block-aligned streams - every real MPEG-4 stream -
// never do: a 16x16 or 8x8 block on the 8-pixel grid covers its cells completely; frame borders and unaligned vectors do).
// The accumulated field never crosses HBM: per GOP the traffic is the vector lists + the cell index + ONE write of the
// result, instead of 2 reads + 1 write of the field and an owner-map round trip for every P-frame.  Bit-exact by
// construction (the same three predicates as the reference's inner loop, coviar_data_loader.c:92-110).
// ---------------------------------------------------------------------------------------------------------
constexpr int kCellShift = 3;   // 8x8 pixel cells

// PASS 0: celltop[cell] = max index + 1 over the vectors touching the cell;  PASS 1: celltop2[cell] = the same over the
// vectors below the cell's top one.  A pixel the top vector does not cover (the strip of a border block whose source
// leaves the frame; a cell an unaligned block covers in part) then tries the second candidate - in block-aligned streams
// there is none and the pixel is done - and only walks further down the list if that one fails too.
template <int PASS>
__global__ void mvacc_celltop_kernel(const MvRec* __restrict__ mvs, const int* __restrict__ counts, int* __restrict__ celltop,
                                     int* __restrict__ celltop2, int T, int M, int height, int width, int cw, int ch) {
  const int nt = blockIdx.y;                                     // n * T + t
  const int cnt = min(__ldg(counts + nt), M);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cnt) return;
  const MvRec mv = mvs[(size_t)nt * M + i];
  if (mv.dst_x - mv.src_x == 0 && mv.dst_y - mv.src_y == 0) return;      // :92
  const int x_lo = (-1 * mv.w) / 2, x_hi = mv.w / 2, y_lo = (-1 * mv.h) / 2, y_hi = mv.h / 2;   // :97-98 (C division)
  if (x_hi <= x_lo || y_hi <= y_lo) return;
  const int x0 = max(mv.dst_x + x_lo, 0), x1 = min(mv.dst_x + x_hi - 1, width - 1);
  const int y0 = max(mv.dst_y + y_lo, 0), y1 = min(mv.dst_y + y_hi - 1, height - 1);
  if (x1 < x0 || y1 < y0) return;
  int* ct = celltop + (size_t)nt * cw * ch;
  int* ct2 = celltop2 + (size_t)nt * cw * ch;
  for (int cy = y0 >> kCellShift; cy <= (y1 >> kCellShift); ++cy)
    for (int cx = x0 >> kCellShift; cx <= (x1 >> kCellShift); ++cx) {
      if (PASS == 0) atomicMax(ct + cy * cw + cx, i + 1);
      else if (i + 1 < ct[cy * cw + cx]) atomicMax(ct2 + cy * cw + cx, i + 1);
    }
}

template <int ROWS>
__global__ void __launch_bounds__(256)
mvacc_trace_kernel(const MvRec* __restrict__ mvs, const int* __restrict__ counts, const int* __restrict__ celltop,
                   const int* __restrict__ celltop2, int2* __restrict__ mv_out, int T, int M, int height, int width, int cw, int ch) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y0 = blockIdx.y * ROWS, n = blockIdx.z;
  if (x >= width) return;
  int px[ROWS], py[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) { px[r] = x; py[r] = y0 + r; }
  for (int t = T - 1; t >= 0; --t) {                             // from the target frame back to the I-frame
    const int nt = n * T + t;
    const int cnt = min(__ldg(counts + nt), M);
    if (cnt == 0) continue;
    const int* __restrict__ ct = celltop + (size_t)nt * cw * ch;
    const int* __restrict__ ct2 = celltop2 + (size_t)nt * cw * ch;
    const int2* __restrict__ list = reinterpret_cast<const int2*>(mvs + (size_t)nt * M);    // 3 x int2 per vector
    int o[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r)                               // the independent chains of this thread issue together
      o[r] = (y0 + r < height) ? __ldg(ct + (py[r] >> kCellShift) * cw + (px[r] >> kCellShift)) : 0;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const int cell = (py[r] >> kCellShift) * cw + (px[r] >> kCellShift);
      int i = o[r] - 1;
      bool first = true;
      while (i >= 0) {                                           // almost always one iteration (see the header)
        const int2 wh = __ldg(list + 3 * i), sr = __ldg(list + 3 * i + 1), ds = __ldg(list + 3 * i + 2);
        const int xs = px[r] - ds.x, ys = py[r] - ds.y;          // offset inside the block: [-w/2, w/2) x [-h/2, h/2)
        const int sx = sr.x + xs, sy = sr.y + ys;
        if (!(ds.x - sr.x == 0 && ds.y - sr.y == 0) && xs >= (-1 * wh.x) / 2 && xs < wh.x / 2 && ys >= (-1 * wh.y) / 2 &&
            ys < wh.y / 2 && (unsigned)sx < (unsigned)width && (unsigned)sy < (unsigned)height) {   // :92, :97-98, :105-108
          px[r] = sx;
          py[r] = sy;
          break;
        }
        if (first) {                                             // the cell's second candidate, then down the list
          first = false;
          i = __ldg(ct2 + cell) - 1;
        } else {
          --i;
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < ROWS; ++r)
    if (y0 + r < height)                                         // :130-139: mv = (x,y) - accu
      mv_out[((size_t)n * height + y0 + r) * width + x] = make_int2(x - px[r], y0 + r - py[r]);
}

size_t mvacc_trace_workspace_bytes(int N, int T, int height, int width) {
  const size_t cw = (size_t)(width + 7) >> kCellShift, ch = (size_t)(height + 7) >> kCellShift;
  return 2 * (size_t)N * (T > 0 ? T : 1) * cw * ch * sizeof(int);      // top + second candidate per cell
}

static cudaError_t run_mv_trace(const int* mvs, const int* counts, int N, int T, int M, int height, int width, int* mv_out,
                                void* workspace, cudaStream_t st) {
  const int cw = (width + 7) >> kCellShift, ch = (height + 7) >> kCellShift;
  int* celltop = static_cast<int*>(workspace);
  int* celltop2 = celltop + (size_t)N * (T > 0 ? T : 1) * cw * ch;
  cudaError_t e = cudaMemsetAsync(celltop, 0, mvacc_trace_workspace_bytes(N, T, height, width), st);
  if (e != cudaSuccess) return e;
  if (T > 0) {
    const dim3 grid((M + 127) / 128, N * T);
    mvacc_celltop_kernel<0><<<grid, 128, 0, st>>>(reinterpret_cast<const MvRec*>(mvs), counts, celltop, celltop2, T, M, height, width, cw, ch);
    mvacc_celltop_kernel<1><<<grid, 128, 0, st>>>(reinterpret_cast<const MvRec*>(mvs), counts, celltop, celltop2, T, M, height, width, cw, ch);
  }
  // independent chains per thread: 1 row 1.88 ms, 2 rows 1.55, 4 rows 1.44, 8 rows 1.41 (64 GOPs x 11 P-frames at 720p; the
  // per-frame field form: 2.86 ms) - the walk is instruction bound (~35 integer instructions per pixel and P-frame)
  constexpr int ROWS = 4;
  mvacc_trace_kernel<ROWS><<<dim3((width + 255) / 256, (height + ROWS - 1) / ROWS, N), 256, 0, st>>>(
      reinterpret_cast<const MvRec*>(mvs), counts, celltop, celltop2, reinterpret_cast<int2*>(mv_out), T, M, height, width, cw, ch);
  return cudaPeekAtLastError();
}

template <typename A>
static cudaError_t run_mv_accumulate(const int* mvs, const int* counts, int N, int T, int M, int height, int width,
                                     int* mv_out, void* workspace, cudaStream_t st) {
  const long long hw = (long long)height * width;
  A* accu0 = static_cast<A*>(workspace);
  A* accu1 = accu0 + (size_t)N * hw;
  int* owner = reinterpret_cast<int*>(accu1 + (size_t)N * hw);
  const dim3 pg((width + 255) / 256, height, N);         // one thread per pixel: (x block, y, GOP)
  const dim3 gg((width + 255) / 256, (height + kGatherRows - 1) / kGatherRows, N);
  const bool tagged = T <= 126 && M < (1 << 24);
  mvacc_init_kernel<A><<<pg, 256, 0, st>>>(accu0, owner, height, width);
  A *old_b = accu0, *new_b = accu1;
  for (int t = 0; t < T; ++t) {
    const int tag = tagged ? (t + 1) << 24 : 0;
    dim3 blk(32, 8), grid(1, (M + 7) / 8, N);
    mvacc_owner_kernel<<<grid, blk, 0, st>>>(reinterpret_cast<const MvRec*>(mvs), counts, owner, t, T, M, height, width, tag);
    mvacc_gather_kernel<A><<<gg, 256, 0, st>>>(reinterpret_cast<const MvRec*>(mvs), old_b, new_b, owner, t, T, M, height, width, tag);
    A* tmp = old_b; old_b = new_b; new_b = tmp;       // :126 memcpy(accu_src_old, accu_src)
  }
  mvacc_finish_kernel<A><<<pg, 256, 0, st>>>(old_b, reinterpret_cast<int2*>(mv_out), height, width);
  return cudaPeekAtLastError();
}

cudaError_t launch_mv_accumulate(const int* mvs, const int* counts, int N, int T, int M, int height, int width,
                                 int* mv_out, void* workspace, size_t workspace_bytes, int algo, cudaStream_t st) {
  // algo: 0 auto, 1 the per-frame field form (round 1), 2 the cell-index back-trace (needs N*T*ceil(w/8)*ceil(h/8)*4 bytes)
  const bool trace_fits = (long long)N * T <= 65535 && workspace_bytes >= mvacc_trace_workspace_bytes(N, T, height, width);
  if (algo == 2 && !trace_fits) return cudaErrorNotSupported;
  if (algo != 1 && trace_fits) return run_mv_trace(mvs, counts, N, T, M, height, width, mv_out, workspace, st);
  if (workspace_bytes < (size_t)N * height * width * 20) return cudaErrorNotSupported;
  if (height <= 32767 && width <= 32767)
    return run_mv_accumulate<short2>(mvs, counts, N, T, M, height, width, mv_out, workspace, st);
  return run_mv_accumulate<int2>(mvs, counts, N, T, M, height, width, mv_out, workspace, st);
}

cudaError_t launch_coviar_residual(const unsigned char* iframe, const unsigned char* cur, const int* mv, int* res, int N,
                                   int height, int width, cudaStream_t st) {
  const dim3 pg((width + 255) / 256, height, N);
  coviar_residual_kernel<<<pg, 256, 0, st>>>(iframe, cur, reinterpret_cast<const int2*>(mv), res, height, width);
  return cudaPeekAtLastError();
}

}  // namespace lsfa
