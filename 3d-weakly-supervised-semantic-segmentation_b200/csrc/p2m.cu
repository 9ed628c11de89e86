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
