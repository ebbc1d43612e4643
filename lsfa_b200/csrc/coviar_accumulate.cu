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

__global__ void mvacc_init_kernel(int2* __restrict__ accu, int* __restrict__ owner, int N, int height, int width) {
  const long long total = (long long)N * height * width;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % ((long long)height * width));
    accu[i] = make_int2(p % width, p / width);     // coviar_data_loader.c:322-326
    owner[i] = 0;
  }
}

// (A) owner[dst] = max over covering vectors of (list index + 1)
__global__ void mvacc_owner_kernel(const MvRec* __restrict__ mvs, const int* __restrict__ counts, int* __restrict__ owner,
                                   int t, int T, int M, int height, int width, int max_block) {
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
  for (int q = threadIdx.x; q < bw * bh; q += blockDim.x) {
    const int xs = x_lo + q / bh, ys = y_lo + q % bh;           // x outer, y inner like the reference (order is irrelevant here)
    const int dx = mv.dst_x + xs, dy = mv.dst_y + ys, sx = mv.src_x + xs, sy = mv.src_y + ys;
    if (dy >= 0 && dy < height && dx >= 0 && dx < width && sy >= 0 && sy < height && sx >= 0 && sx < width)
      atomicMax(own + (size_t)dy * width + dx, i + 1);
  }
  (void)max_block;
}

// (B) gather + clear owner
__global__ void mvacc_gather_kernel(const MvRec* __restrict__ mvs, const int2* __restrict__ accu_old, int2* __restrict__ accu_new,
                                    int* __restrict__ owner, int t, int T, int M, int N, int height, int width) {
  const long long hw = (long long)height * width, total = (long long)N * hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / hw);
    const int p = (int)(i - (long long)n * hw);
    const int o = owner[i];
    int2 v = accu_old[i];
    if (o > 0) {
      const MvRec mv = mvs[((size_t)n * T + t) * M + (o - 1)];
      const int x = p % width, y = p / width;
      const int sx = mv.src_x + (x - mv.dst_x), sy = mv.src_y + (y - mv.dst_y);
      v = accu_old[(size_t)n * hw + (size_t)sy * width + sx];
      owner[i] = 0;
    }
    accu_new[i] = v;
  }
}

// target frame: mv = (x,y) - accu   (:130-139), layout (N,height,width,2) like the numpy array coviar returns
__global__ void mvacc_finish_kernel(const int2* __restrict__ accu, int2* __restrict__ mv_out, int N, int height, int width) {
  const long long total = (long long)N * height * width;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % ((long long)height * width));
    const int2 a = accu[i];
    mv_out[i] = make_int2(p % width - a.x, p / width - a.y);
  }
}

// residual (:141-175, accumulate case): res[y][x][c] = cur[y][x][c] - iframe[src_y][src_x][c], src = (x,y) - mv
__global__ void coviar_residual_kernel(const unsigned char* __restrict__ iframe, const unsigned char* __restrict__ cur,
                                       const int2* __restrict__ mv, int* __restrict__ res, int N, int height, int width) {
  const long long hw = (long long)height * width, total = (long long)N * hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / hw);
    const int p = (int)(i - (long long)n * hw);
    const int2 m = mv[i];
    const int sx = p % width - m.x, sy = p / width - m.y;      // always in bounds for an accumulated field
    const size_t src = ((size_t)n * hw + (size_t)sy * width + sx) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) res[i * 3 + c] = (int)cur[i * 3 + c] - (int)iframe[src + c];
  }
}

static inline int ew_grid2(long long total, int threads) {
  long long g = (total + threads - 1) / threads;
  if (g > 148LL * 16) g = 148LL * 16;
  if (g < 1) g = 1;
  return (int)g;
}

cudaError_t launch_mv_accumulate(const int* mvs, const int* counts, int N, int T, int M, int height, int width,
                                 int* mv_out, void* workspace, cudaStream_t st) {
  const long long hw = (long long)height * width;
  int2* accu0 = static_cast<int2*>(workspace);
  int2* accu1 = accu0 + (size_t)N * hw;
  int* owner = reinterpret_cast<int*>(accu1 + (size_t)N * hw);
  const int g = ew_grid2((long long)N * hw, 256);
  mvacc_init_kernel<<<g, 256, 0, st>>>(accu0, owner, N, height, width);
  int2 *old_b = accu0, *new_b = accu1;
  for (int t = 0; t < T; ++t) {
    dim3 blk(32, 8), grid(1, (M + 7) / 8, N);
    mvacc_owner_kernel<<<grid, blk, 0, st>>>(reinterpret_cast<const MvRec*>(mvs), counts, owner, t, T, M, height, width, 0);
    mvacc_gather_kernel<<<g, 256, 0, st>>>(reinterpret_cast<const MvRec*>(mvs), old_b, new_b, owner, t, T, M, N, height, width);
    int2* tmp = old_b; old_b = new_b; new_b = tmp;       // :126 memcpy(accu_src_old, accu_src)
  }
  mvacc_finish_kernel<<<g, 256, 0, st>>>(old_b, reinterpret_cast<int2*>(mv_out), N, height, width);
  return cudaPeekAtLastError();
}

cudaError_t launch_coviar_residual(const unsigned char* iframe, const unsigned char* cur, const int* mv, int* res, int N,
                                   int height, int width, cudaStream_t st) {
  coviar_residual_kernel<<<ew_grid2((long long)N * height * width, 256), 256, 0, st>>>(
      iframe, cur, reinterpret_cast<const int2*>(mv), res, N, height, width);
  return cudaPeekAtLastError();
}

}  // namespace lsfa
