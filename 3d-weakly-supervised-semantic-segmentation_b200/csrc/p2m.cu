// p2m.cu -- ops/point2mask kernels.
//
// Same results as ops/point2mask/_ext_src/src/ball_query_gpu.cu:9-45 and group_points_gpu.cu:8-28,43-64,
// re-shaped for a 148-SM part: the reference launches ONE block per batch element and scans the points
// serially from global memory; here queries are spread over (query tile, batch) CTAs, candidate points are
// staged through shared memory in index order (so "first nsample hits in ascending index" is preserved
// exactly, including the reference's `k < n - ptnum` scan bound), and a CTA stops as soon as all of its
// queries are full.
#include "common.cuh"

namespace b200scn {

constexpr int kBqTile = 2048;  // candidate points per shared-memory stage (16 KB)

__global__ void __launch_bounds__(256)
ball_query_kernel(int n, int m, float radius2, int nsample, const float *__restrict__ new_xy,
                  const float *__restrict__ xy, const int32_t *__restrict__ pointnums,
                  int32_t *__restrict__ idx) {
  __shared__ float2 pts[kBqTile];
  const int bi = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const float2 *p = reinterpret_cast<const float2 *>(xy) + (int64_t)bi * n;
  const int limit = n - pointnums[bi];  // reference scan bound (ball_query_gpu.cu:28)
  float qx = 0.f, qy = 0.f;
  int32_t *o = nullptr;
  const bool live = j < m;
  if (live) {
    float2 q = reinterpret_cast<const float2 *>(new_xy)[(int64_t)bi * m + j];
    qx = q.x; qy = q.y;
    o = idx + ((int64_t)bi * m + j) * nsample;
  }
  int cnt = 0;
  for (int base = 0; base < limit; base += kBqTile) {
    const int len = min(kBqTile, limit - base);
    __syncthreads();
    for (int e = threadIdx.x; e < len; e += blockDim.x) pts[e] = __ldg(p + base + e);
    __syncthreads();
    if (live && cnt < nsample) {
      for (int e = 0; e < len; ++e) {
        float dx = qx - pts[e].x, dy = qy - pts[e].y;
        // same operation order as the reference: (dx*dx) + (dy*dy), no fma contraction
        float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        if (d2 < radius2) {
          o[cnt] = base + e;
          if (++cnt >= nsample) break;
        }
      }
    }
    if (__syncthreads_and(!live || cnt >= nsample)) break;
  }
  if (live)
    for (int c = cnt; c < nsample; ++c) o[c] = -1;
}

// ------------------------------------------------------------------------------------------------------------------
// Cell-bucketed ball query for the production shape (B = 64, 65 536 pixel-centre queries, ~200 k points,
// radius 1, nsample 20: ops/pseudo_dataset_generator/configs.py:11-12, preprocess_mask.py:31-32), where the scan above is
// O(m n) per instance.  Candidate points (index < n - ptnum, the reference's scan bound) are counting-sorted into square
// cells of side r over the queries' bounding box; a query visits the cells that cover [q - r, q + r]^2 and MERGES their
// index-sorted lists, so hits still come out in ascending original index and "the first nsample in index order" is
// exactly the reference's selection.  Same distance expression, same strict `<`.
struct BqGrid { float x0, y0, inv; int nx, ny; };

__device__ __forceinline__ unsigned ordf(float v) {
  unsigned u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unordf(unsigned o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o);
}

// bounding box of ALL queries (one grid geometry for the whole batch keeps the cell arithmetic uniform)
__global__ void bq_bbox_kernel(const float2 *__restrict__ q, int64_t total, unsigned *__restrict__ box /*4: min x,y max x,y*/) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float x = 0.f, y = 0.f;
  const bool live = i < total;
  if (live) { const float2 v = __ldg(q + i); x = v.x; y = v.y; }
  unsigned mnx = live ? ordf(x) : 0xFFFFFFFFu, mny = live ? ordf(y) : 0xFFFFFFFFu;
  unsigned mxx = live ? ordf(x) : 0u, mxy = live ? ordf(y) : 0u;
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, d)); mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, d));
    mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, d)); mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, d));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(box + 0, mnx); atomicMin(box + 1, mny); atomicMax(box + 2, mxx); atomicMax(box + 3, mxy);
  }
}

__device__ __forceinline__ BqGrid bq_grid(const unsigned *__restrict__ box, float radius, int max_side) {
  BqGrid g;
  const float pad = radius * 1.001f;
  g.x0 = unordf(box[0]) - pad; g.y0 = unordf(box[1]) - pad;
  g.inv = 1.f / radius;
  const float ex = unordf(box[2]) + pad - g.x0, ey = unordf(box[3]) + pad - g.y0;
  g.nx = min(max_side, max(1, (int)floorf(ex * g.inv) + 1));
  g.ny = min(max_side, max(1, (int)floorf(ey * g.inv) + 1));
  return g;
}
// cell of a coordinate, or -1 outside the grid (monotone in v: subtraction, multiplication by a positive constant, floor)
__device__ __forceinline__ int bq_cell1(float v, float v0, float inv, int n) {
  const float c = floorf((v - v0) * inv);
  return (c >= 0.f && c < (float)n) ? (int)c : -1;
}

// MODE 0: count candidates per cell.  MODE 1: place them (cursor = running count; order inside a cell fixed afterwards).
template <int MODE>
__global__ void bq_bin_kernel(int n, const float2 *__restrict__ xy, const int32_t *__restrict__ pointnums,
                              const unsigned *__restrict__ box, float radius, int max_side, int cells_cap,
                              int32_t *__restrict__ counts, const int32_t *__restrict__ starts, int32_t *__restrict__ items) {
  const int bi = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n - __ldg(pointnums + bi)) return;          // reference scan bound (ball_query_gpu.cu:28)
  const BqGrid g = bq_grid(box, radius, max_side);
  const float2 p = __ldg(xy + (int64_t)bi * n + k);
  const int cx = bq_cell1(p.x, g.x0, g.inv, g.nx), cy = bq_cell1(p.y, g.y0, g.inv, g.ny);
  if (cx < 0 || cy < 0) return;
  const int64_t cell = (int64_t)bi * cells_cap + (int64_t)cy * g.nx + cx;
  if (MODE == 0) {
    atomicAdd(counts + cell, 1);
  } else {
    const int pos = atomicAdd(counts + cell, 1);
    items[__ldg(starts + cell) + pos] = k;
  }
}

struct BqCountLoader {
  const int32_t *counts;
  __device__ int live(int n) const { return n; }
  __device__ int operator()(int64_t i) const { return counts[i]; }
};
struct BqStartWriter {
  int32_t *starts; int32_t *counts;
  __device__ void operator()(int64_t i, int flag, int pos) const { starts[i] = pos; counts[i] = 0; (void)flag; }
};

// ascending index order inside every cell (lists are a handful of entries: insertion sort by one thread per cell)
__global__ void bq_sort_cells_kernel(int64_t total_cells, const int32_t *__restrict__ starts,
                                     const int32_t *__restrict__ counts, int32_t *__restrict__ items) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= total_cells) return;
  const int s = starts[c], e = s + counts[c];
  for (int i = s + 1; i < e; ++i) {
    const int v = items[i];
    int j = i - 1;
    while (j >= s && items[j] > v) { items[j + 1] = items[j]; --j; }
    items[j + 1] = v;
  }
}

constexpr int kBqMaxLists = 16;   // cells covering [q - r, q + r]^2: 3 x 3, occasionally 4 along an axis

__global__ void __launch_bounds__(128)
bq_query_kernel(int n, int m, float radius, float radius2, int nsample, const float2 *__restrict__ new_xy,
                const float2 *__restrict__ xy, const unsigned *__restrict__ box, int max_side, int cells_cap,
                const int32_t *__restrict__ starts, const int32_t *__restrict__ counts,
                const int32_t *__restrict__ items, int32_t *__restrict__ idx) {
  const int bi = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  const BqGrid g = bq_grid(box, radius, max_side);
  const float2 q = __ldg(new_xy + (int64_t)bi * m + j);
  const float2 *p = xy + (int64_t)bi * n;
  int32_t *o = idx + ((int64_t)bi * m + j) * nsample;
  const float pad = radius * 1.0001f;   // the cell range is padded; membership is decided by the distance test alone
  int x_lo = (int)floorf((q.x - pad - g.x0) * g.inv), x_hi = (int)floorf((q.x + pad - g.x0) * g.inv);
  int y_lo = (int)floorf((q.y - pad - g.y0) * g.inv), y_hi = (int)floorf((q.y + pad - g.y0) * g.inv);
  x_lo = max(x_lo, 0); y_lo = max(y_lo, 0); x_hi = min(x_hi, g.nx - 1); y_hi = min(y_hi, g.ny - 1);
  int cur[kBqMaxLists], end[kBqMaxLists];
  int nl = 0;
  for (int cy = y_lo; cy <= y_hi; ++cy)
    for (int cx = x_lo; cx <= x_hi; ++cx) {
      const int64_t cell = (int64_t)bi * cells_cap + (int64_t)cy * g.nx + cx;
      const int c = __ldg(counts + cell);
      if (c > 0 && nl < kBqMaxLists) {
        cur[nl] = __ldg(starts + cell);
        end[nl] = cur[nl] + c;
        ++nl;
      }
    }
  int cnt = 0;
  while (cnt < nsample) {
    // smallest head among the lists = next candidate in ascending point index
    int best = -1, bk = 0x7FFFFFFF;
#pragma unroll
    for (int l = 0; l < kBqMaxLists; ++l) {
      if (l < nl && cur[l] < end[l]) {
        const int k = __ldg(items + cur[l]);
        if (k < bk) { bk = k; best = l; }
      }
    }
    if (best < 0) break;
#pragma unroll
    for (int l = 0; l < kBqMaxLists; ++l)
      if (l == best) ++cur[l];
    const float2 c = __ldg(p + bk);
    const float dx = q.x - c.x, dy = q.y - c.y;
    const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    if (d2 < radius2) o[cnt++] = bk;
  }
  for (int c = cnt; c < nsample; ++c) o[c] = -1;
}

__global__ void group_points_kernel(int c, int n, int npoints, int nsample,
                                    const float *__restrict__ points, const int32_t *__restrict__ idx,
                                    float *__restrict__ out, int64_t total) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int64_t ps = (int64_t)npoints * nsample;
  int64_t bl = t / ps;            // bi*c + l
  int64_t jk = t - bl * ps;       // j*nsample + k
  int64_t bi = bl / c;
  int ii = __ldg(idx + bi * ps + jk);
  out[t] = ii >= 0 ? __ldg(points + bl * n + ii) : 0.f;
}

__global__ void group_points_grad_kernel(int c, int n, int npoints, int nsample,
                                         const float *__restrict__ grad_out,
                                         const int32_t *__restrict__ idx, float *grad_points,
                                         int64_t total) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int64_t ps = (int64_t)npoints * nsample;
  int64_t bl = t / ps;
  int64_t jk = t - bl * ps;
  int64_t bi = bl / c;
  int ii = __ldg(idx + bi * ps + jk);
  if (ii >= 0) atomicAdd(grad_points + bl * n + ii, __ldg(grad_out + t));
}

}  // namespace b200scn

using namespace b200scn;

extern "C" {

/* bytes of scratch the bucketed ball query needs, or 0 when the shape is served by the scan kernel (tiny inputs) */
size_t b200scn_p2m_ball_query_scratch_bytes(int b, int n, int m, int max_side) {
  if (b <= 0 || n <= 0 || m <= 0 || max_side <= 0) return 0;
  const size_t cells = (size_t)b * max_side * max_side;
  return 16 + sizeof(int32_t) * (2 * cells + (size_t)b * n + scan_scratch_ints((int64_t)cells)) + 256;
}

/* Cell-bucketed variant: identical results to b200scn_p2m_ball_query.  max_side bounds the cells per axis (the grid covers
 * the queries' bounding box padded by the radius with cells of side `radius`; if the box needs more cells the outer ones
 * are clamped away and their points are simply never candidates -- so max_side must be >= extent / radius + 3; the
 * Python binding derives it from the known pixel-grid resolution). */
int b200scn_p2m_ball_query_bucketed(int b, int n, int m, float radius, int nsample, const float *new_xy,
                                    const float *xy, const int32_t *pointnums, int32_t *idx, int max_side,
                                    void *scratch, size_t scratch_bytes, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (b <= 0 || m <= 0 || nsample <= 0) return 0;
  if (!(radius > 0.f)) return set_error("p2m_ball_query_bucketed: radius must be positive");
  if (scratch_bytes < b200scn_p2m_ball_query_scratch_bytes(b, n, m, max_side) || scratch_bytes == 0)
    return set_error("p2m_ball_query_bucketed: scratch too small");
  const int cells_cap = max_side * max_side;
  const int64_t total_cells = (int64_t)b * cells_cap;
  if (total_cells >= ((int64_t)1 << 31)) return set_error("p2m_ball_query_bucketed: too many cells");
  unsigned *box = reinterpret_cast<unsigned *>(scratch);
  int32_t *counts = reinterpret_cast<int32_t *>(reinterpret_cast<uint8_t *>(scratch) + 16);
  int32_t *starts = counts + total_cells;
  int32_t *items = starts + total_cells;
  int32_t *sums = items + (size_t)b * n;
  const unsigned init[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u};
  SCN_CUDA(cudaMemcpyAsync(box, init, sizeof(init), cudaMemcpyHostToDevice, st));
  SCN_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * (size_t)total_cells, st));
  const float2 *q2 = reinterpret_cast<const float2 *>(new_xy), *p2 = reinterpret_cast<const float2 *>(xy);
  bq_bbox_kernel<<<(unsigned)ceil_div((int64_t)b * m, 256), 256, 0, st>>>(q2, (int64_t)b * m, box);
  dim3 pg((unsigned)ceil_div(n, 256), (unsigned)b);
  bq_bin_kernel<0><<<pg, 256, 0, st>>>(n, p2, pointnums, box, radius, max_side, cells_cap, counts, nullptr, nullptr);
  SCN_CHECK_LAUNCH("p2m_bin_count");
  BqCountLoader ld{counts};
  BqStartWriter wr{starts, counts};     // starts = exclusive scan of the counts; counts zeroed to serve as cursors
  if (scan_flags(ld, wr, total_cells, nullptr, sums, nullptr, st)) return 1;
  bq_bin_kernel<1><<<pg, 256, 0, st>>>(n, p2, pointnums, box, radius, max_side, cells_cap, counts, starts, items);
  bq_sort_cells_kernel<<<(unsigned)ceil_div(total_cells, 256), 256, 0, st>>>(total_cells, starts, counts, items);
  dim3 qg((unsigned)ceil_div(m, 128), (unsigned)b);
  bq_query_kernel<<<qg, 128, 0, st>>>(n, m, radius, radius * radius, nsample, q2, p2, box, max_side, cells_cap, starts,
                                      counts, items, idx);
  SCN_CHECK_LAUNCH("p2m_ball_query_bucketed");
  count_launch(5);
  return 0;
}

int b200scn_p2m_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xy,
                           const float *xy, const int32_t *pointnums, int32_t *idx, void *stream) {
  if (b <= 0 || m <= 0 || nsample <= 0) return 0;
  dim3 grid((unsigned)ceil_div(m, 256), (unsigned)b);
  ball_query_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, m, radius * radius, nsample, new_xy, xy, pointnums, idx);
  SCN_CHECK_LAUNCH("p2m_ball_query");
  count_launch(1);
  return 0;
}

int b200scn_p2m_group_points(int b, int c, int n, int npoints, int nsample, const float *points,
                             const int32_t *idx, float *out, void *stream) {
  int64_t total = (int64_t)b * c * npoints * nsample;
  if (total <= 0) return 0;
  group_points_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(c, n, npoints, nsample, points, idx, out, total);
  SCN_CHECK_LAUNCH("p2m_group_points");
  count_launch(1);
  return 0;
}

int b200scn_p2m_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out,
                                  const int32_t *idx, float *grad_points, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if ((int64_t)b * c * n > 0) SCN_CUDA(cudaMemsetAsync(grad_points, 0, sizeof(float) * (size_t)b * c * n, st));
  int64_t total = (int64_t)b * c * npoints * nsample;
  if (total <= 0) return 0;
  group_points_grad_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(c, n, npoints, nsample, grad_out, idx, grad_points, total);
  SCN_CHECK_LAUNCH("p2m_group_points_grad");
  count_launch(1);
  return 0;
}

}  // extern "C"
