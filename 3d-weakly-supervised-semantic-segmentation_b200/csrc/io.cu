// io.cu -- InputLayer / OutputLayer feature movement (App. B.2, B.3).
// Replaces upstream scn's InputLayer_updateOutput / OutputLayer_updateOutput (+GradInput) in IOLayers.cu.
#include "common.cuh"

namespace b200scn {

__device__ __forceinline__ float input_mult(int mode, int64_t r, int v, const int32_t *count,
                                            const int32_t *first_row, const int32_t *last_row) {
  switch (mode) {
    case 4: return 1.f / (float)__ldg(count + v);
    case 3: return 1.f;
    case 2: return __ldg(first_row + v) == (int32_t)r ? 1.f : 0.f;
    default: return __ldg(last_row + v) == (int32_t)r ? 1.f : 0.f;
  }
}
__device__ __forceinline__ float output_sel(int mode, int64_t r, int v, const int32_t *first_row,
                                            const int32_t *last_row) {
  if (mode == 2) return __ldg(first_row + v) == (int32_t)r ? 1.f : 0.f;
  if (mode == 1) return __ldg(last_row + v) == (int32_t)r ? 1.f : 0.f;
  return 1.f;
}

__global__ void input_features_kernel(const float *__restrict__ feats, int64_t P, int C,
                                      const int32_t *__restrict__ pv, const int32_t *count,
                                      const int32_t *first_row, const int32_t *last_row, int mode,
                                      float *out) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P * C) return;
  int64_t r = t / C;
  int c = (int)(t - r * C);
  int v = __ldg(pv + r);
  float m = input_mult(mode, r, v, count, first_row, last_row);
  if (m != 0.f) atomicAdd(out + (int64_t)v * C + c, m * __ldg(feats + t));
}

__global__ void input_features_bwd_kernel(const float *__restrict__ d_out, int64_t P, int C,
                                          const int32_t *__restrict__ pv, const int32_t *count,
                                          const int32_t *first_row, const int32_t *last_row, int mode,
                                          float *__restrict__ d_in) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P * C) return;
  int64_t r = t / C;
  int c = (int)(t - r * C);
  int v = __ldg(pv + r);
  d_in[t] = input_mult(mode, r, v, count, first_row, last_row) * __ldg(d_out + (int64_t)v * C + c);
}

template <int VEC>
__global__ void output_features_kernel(const float *__restrict__ feats, int64_t ldf, int64_t P, int C,
                                       const int32_t *__restrict__ pv, const int32_t *first_row,
                                       const int32_t *last_row, int mode, float *__restrict__ out) {
  const int cv = C / VEC;
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P * cv) return;
  int64_t r = t / cv;
  int c0 = (int)(t - r * cv) * VEC;
  int v = __ldg(pv + r);
  float s = output_sel(mode, r, v, first_row, last_row);
  if (VEC == 4) {
    float4 a = __ldg(reinterpret_cast<const float4 *>(feats + (int64_t)v * ldf + c0));
    a.x *= s; a.y *= s; a.z *= s; a.w *= s;
    *reinterpret_cast<float4 *>(out + r * C + c0) = a;
  } else {
    out[r * C + c0] = s * __ldg(feats + (int64_t)v * ldf + c0);
  }
}

// points grouped by site (counting sort): start[v] = first slot of site v, rows[start[v] .. start[v]+count[v]) = its rows.
// The order inside a site is the arrival order of the atomics (any order gives the same set).
__global__ void site_rows_kernel(const int32_t *__restrict__ pv, int64_t P, const int32_t *__restrict__ start,
                                 int32_t *__restrict__ cursor, int32_t *__restrict__ rows) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= P) return;
  const int v = __ldg(pv + r);
  rows[__ldg(start + v) + atomicAdd(cursor + v, 1)] = (int32_t)r;
}

struct CountLoader {
  const int32_t *count;
  __device__ int live(int n) const { return n; }
  __device__ int operator()(int64_t i) const { return count[i]; }
};
struct StartWriter {
  int32_t *start;
  __device__ void operator()(int64_t i, int c, int pos) const { start[i] = pos; }
};

// d_feats[v,:] = sum over the rows of site v of sel(row) * d_out[row,:]  -- gather, no atomics, every d_out row read once
template <int VEC>
__global__ void __launch_bounds__(256)
output_features_bwd_csr_kernel(const float *__restrict__ d_out, int64_t N, int C, const int32_t *__restrict__ start,
                               const int32_t *__restrict__ count, const int32_t *__restrict__ rows,
                               const int32_t *first_row, const int32_t *last_row, int mode,
                               float *__restrict__ d_feats, int64_t ldf) {
  const int cv = C / VEC;
  const int64_t total = N * cv;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = t / cv;
    const int c0 = (int)(t - v * cv) * VEC;
    const int s0 = __ldg(start + v), n = __ldg(count + v);
    float acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
    for (int q = 0; q < n; ++q) {
      const int r = __ldg(rows + s0 + q);
      if (output_sel(mode, r, (int)v, first_row, last_row) == 0.f) continue;
      if (VEC == 4) {
        float4 a = __ldg(reinterpret_cast<const float4 *>(d_out + (int64_t)r * C + c0));
        acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
      } else {
        acc[0] += __ldg(d_out + (int64_t)r * C + c0);
      }
    }
    if (VEC == 4) *reinterpret_cast<float4 *>(d_feats + v * ldf + c0) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    else d_feats[v * ldf + c0] = acc[0];
  }
}

__global__ void output_features_bwd_kernel(const float *__restrict__ d_out, int64_t P, int C,
                                           const int32_t *__restrict__ pv, const int32_t *first_row,
                                           const int32_t *last_row, int mode, float *d_feats,
                                           int64_t ldf) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P * C) return;
  int64_t r = t / C;
  int c = (int)(t - r * C);
  int v = __ldg(pv + r);
  float s = output_sel(mode, r, v, first_row, last_row);
  if (s != 0.f) atomicAdd(d_feats + (int64_t)v * ldf + c, __ldg(d_out + t));
}

}  // namespace b200scn

using namespace b200scn;

extern "C" {

static int check_mode(int mode) {
  if (mode < 1 || mode > 4) return set_error("InputLayer mode %d unsupported (1..4)", mode);
  return 0;
}

int b200scn_input_features(const float *feats, int64_t P, int C, const int32_t *pv,
                           const int32_t *count, const int32_t *first_row, const int32_t *last_row,
                           int mode, float *out, void *stream) {
  if (check_mode(mode)) return 1;
  if (P * C <= 0) return 0;
  input_features_kernel<<<(unsigned)ceil_div(P * C, 256), 256, 0, (cudaStream_t)stream>>>(feats, P, C, pv, count, first_row, last_row, mode, out);
  SCN_CHECK_LAUNCH("input_features");
  count_launch(1);
  return 0;
}

int b200scn_input_features_bwd(const float *d_out, int64_t P, int C, const int32_t *pv,
                               const int32_t *count, const int32_t *first_row,
                               const int32_t *last_row, int mode, float *d_in, void *stream) {
  if (check_mode(mode)) return 1;
  if (P * C <= 0) return 0;
  input_features_bwd_kernel<<<(unsigned)ceil_div(P * C, 256), 256, 0, (cudaStream_t)stream>>>(d_out, P, C, pv, count, first_row, last_row, mode, d_in);
  SCN_CHECK_LAUNCH("input_features_bwd");
  count_launch(1);
  return 0;
}

int b200scn_output_features(const float *feats, int64_t ldf, int64_t P, int C, const int32_t *pv,
                            const int32_t *first_row, const int32_t *last_row, int mode, float *out,
                            void *stream) {
  if (check_mode(mode)) return 1;
  if (P * C <= 0) return 0;
  bool vec = C % 4 == 0 && ldf % 4 == 0 && ((reinterpret_cast<uintptr_t>(feats) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  if (vec)
    output_features_kernel<4><<<(unsigned)ceil_div(P * (C / 4), 256), 256, 0, (cudaStream_t)stream>>>(feats, ldf, P, C, pv, first_row, last_row, mode, out);
  else
    output_features_kernel<1><<<(unsigned)ceil_div(P * C, 256), 256, 0, (cudaStream_t)stream>>>(feats, ldf, P, C, pv, first_row, last_row, mode, out);
  SCN_CHECK_LAUNCH("output_features");
  count_launch(1);
  return 0;
}

int b200scn_output_features_bwd(const float *d_out, int64_t P, int C, const int32_t *pv,
                                const int32_t *first_row, const int32_t *last_row, int mode,
                                float *d_feats, int64_t ldf, void *stream) {
  if (check_mode(mode)) return 1;
  if (P * C <= 0) return 0;
  output_features_bwd_kernel<<<(unsigned)ceil_div(P * C, 256), 256, 0, (cudaStream_t)stream>>>(d_out, P, C, pv, first_row, last_row, mode, d_feats, ldf);
  SCN_CHECK_LAUNCH("output_features_bwd");
  count_launch(1);
  return 0;
}

}  // extern "C"

extern "C" {

size_t b200scn_site_rows_scratch_bytes(int64_t n_sites) {
  return sizeof(int32_t) * ((size_t)(n_sites > 0 ? n_sites : 1) + scan_scratch_ints(n_sites) + 64);
}

int b200scn_site_rows(const int32_t *pv, int64_t P, const int32_t *count, int64_t n_sites, int32_t *start,
                      int32_t *rows, void *scratch, size_t scratch_bytes, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (scratch_bytes < b200scn_site_rows_scratch_bytes(n_sites)) return set_error("site_rows: scratch too small");
  if (n_sites <= 0 || P <= 0) return 0;
  int32_t *cursor = (int32_t *)scratch;
  int32_t *block_sums = cursor + n_sites;
  SCN_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int32_t) * (size_t)n_sites, st));
  CountLoader ld{count};
  StartWriter wr{start};
  if (scan_flags(ld, wr, n_sites, nullptr, block_sums, nullptr, st)) return 1;
  site_rows_kernel<<<(unsigned)ceil_div(P, 256), 256, 0, st>>>(pv, P, start, cursor, rows);
  SCN_CHECK_LAUNCH("site_rows");
  count_launch(1);
  return 0;
}

int b200scn_output_features_bwd_csr(const float *d_out, int64_t n_sites, int C, const int32_t *start,
                                    const int32_t *count, const int32_t *rows, const int32_t *first_row,
                                    const int32_t *last_row, int mode, float *d_feats, int64_t ldf, void *stream) {
  if (check_mode(mode)) return 1;
  if (n_sites * C <= 0) return 0;
  const bool vec = C % 4 == 0 && ldf % 4 == 0 &&
                   ((reinterpret_cast<uintptr_t>(d_out) | reinterpret_cast<uintptr_t>(d_feats)) & 15) == 0;
  const int64_t total = n_sites * (vec ? C / 4 : C);
  const unsigned blocks = (unsigned)min((int64_t)1 << 30, ceil_div(total, 256));   // one element group per thread
  if (vec)
    output_features_bwd_csr_kernel<4><<<blocks, 256, 0, (cudaStream_t)stream>>>(d_out, n_sites, C, start, count, rows, first_row, last_row, mode, d_feats, ldf);
  else
    output_features_bwd_csr_kernel<1><<<blocks, 256, 0, (cudaStream_t)stream>>>(d_out, n_sites, C, start, count, rows, first_row, last_row, mode, d_feats, ldf);
  SCN_CHECK_LAUNCH("output_features_bwd_csr");
  count_launch(1);
  return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// Head pooling (SparseConvBase_.postProcessing, models/SparseConvNet.py:20-26, and
// models/MultiLabelContrastive.py:35-40 -- a Python loop of torch.mean over out_feats[batch_offsets[i]:batch_offsets[i+1]]).
// The per-point tensor only exists to be averaged per scene, and OutputLayer copies a voxel's row to each of its points, so
//   pooled[b][c] = (1 / P_b) * sum over voxels v of scene b of  w(v) * feats[v][c],   w(v) = points in v (modes 3, 4) or 1
// never needs the (sum P, C) tensor.  Level-0 ids are contiguous per scene (first occurrence in input row order, scenes
// concatenated), so a block walking consecutive rows flushes its partial sums once per scene change.
namespace b200scn {

constexpr int kPoolRows = 256;

__global__ void __launch_bounds__(128)
scene_pool_kernel(const float *__restrict__ feats, int64_t ldf, const uint64_t *__restrict__ ukeys,
                  const int32_t *__restrict__ count, int mode, int64_t n, int C, float *__restrict__ sums,
                  float *__restrict__ npts) {
  const int64_t r0 = (int64_t)blockIdx.x * kPoolRows, r1 = min(n, r0 + kPoolRows);
  for (int c0 = 0; c0 < C; c0 += 128) {
    const int c = c0 + threadIdx.x;
    float acc = 0.f, pts = 0.f;
    int cur = r0 < r1 ? (int)(__ldg(ukeys + r0) >> 48) : 0;
    for (int64_t r = r0; r < r1; ++r) {
      const int b = (int)(__ldg(ukeys + r) >> 48);
      if (b != cur) {   // block-uniform
        if (c < C) atomicAdd(sums + (int64_t)cur * C + c, acc);
        if (c == 0) atomicAdd(npts + cur, pts);
        acc = 0.f; pts = 0.f; cur = b;
      }
      const float cnt = (float)__ldg(count + r);
      const float w = mode >= 3 ? cnt : 1.f;
      if (c < C) acc = fmaf(w, __ldg(feats + r * ldf + c), acc);
      pts += cnt;
    }
    if (r0 < r1) {
      if (c < C) atomicAdd(sums + (int64_t)cur * C + c, acc);
      if (c == 0) atomicAdd(npts + cur, pts);
    }
  }
}

__global__ void scene_pool_finish_kernel(float *__restrict__ sums, const float *__restrict__ npts, int B, int C) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * C) return;
  const float p = npts[t / C];
  sums[t] = p > 0.f ? sums[t] / p : 0.f;
}

__global__ void scene_pool_bwd_kernel(const float *__restrict__ g, const uint64_t *__restrict__ ukeys,
                                      const int32_t *__restrict__ count, int mode, const float *__restrict__ npts,
                                      int64_t n, int C, float *__restrict__ d_feats, int64_t ldd) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * C) return;
  const int64_t r = t / C;
  const int c = (int)(t - r * C);
  const int b = (int)(__ldg(ukeys + r) >> 48);
  const float w = mode >= 3 ? (float)__ldg(count + r) : 1.f;
  d_feats[r * ldd + c] = w * __ldg(g + (int64_t)b * C + c) / __ldg(npts + b);
}

}  // namespace b200scn

extern "C" {

int b200scn_scene_mean(const float *feats, int64_t ldf, const uint64_t *ukeys, const int32_t *count, int mode,
                       int64_t n, int C, int B, float *out, float *npts, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (B <= 0 || C <= 0) return set_error("scene_mean: bad sizes B=%d C=%d", B, C);
  SCN_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)B * C, st));
  SCN_CUDA(cudaMemsetAsync(npts, 0, sizeof(float) * (size_t)B, st));
  if (n > 0) {
    // npts is accumulated once per channel chunk by thread 0 of the first chunk only (c == 0 exists in chunk c0 == 0)
    scene_pool_kernel<<<(unsigned)ceil_div(n, kPoolRows), 128, 0, st>>>(feats, ldf, ukeys, count, mode, n, C, out, npts);
    scene_pool_finish_kernel<<<(unsigned)ceil_div((int64_t)B * C, 256), 256, 0, st>>>(out, npts, B, C);
    SCN_CHECK_LAUNCH("scene_mean");
    count_launch(2);
  }
  return 0;
}

int b200scn_scene_mean_bwd(const float *g, const uint64_t *ukeys, const int32_t *count, int mode, const float *npts,
                           int64_t n, int C, float *d_feats, int64_t ldd, void *stream) {
  if (n <= 0) return 0;
  scene_pool_bwd_kernel<<<(unsigned)ceil_div(n * C, 256), 256, 0, (cudaStream_t)stream>>>(g, ukeys, count, mode, npts, n, C,
                                                                                         d_feats, ldd);
  SCN_CHECK_LAUNCH("scene_mean_bwd");
  count_launch(1);
  return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// scn.MaxPooling (size == stride) and scn.SparseToDense (models/projector/components.py:78-100,
// Function_test.py:46,87,203): adjacent scn ops on the rulebooks that already exist.
namespace b200scn {

// out[j][c] = max(0, max_k in[child[j][k]][c])   (upstream zero-initialises the output and takes the running maximum)
__global__ void maxpool_kernel(const float *__restrict__ in, int64_t ldi, const int32_t *__restrict__ child,
                               int64_t n_coarse, int K, int C, float *__restrict__ out, int64_t ldo) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_coarse * C) return;
  const int64_t j = t / C;
  const int c = (int)(t - j * C);
  float m = 0.f;
  for (int k = 0; k < K; ++k) {
    const int i = __ldg(child + j * K + k);
    if (i >= 0) m = fmaxf(m, __ldg(in + (int64_t)i * ldi + c));
  }
  out[j * ldo + c] = m;
}

// d_in[i][c] = g[parent[i]][c] where in[i][c] equals the pooled value (ties all receive it), else 0
__global__ void maxpool_bwd_kernel(const float *__restrict__ g, int64_t ldg, const float *__restrict__ in, int64_t ldi,
                                   const float *__restrict__ out, int64_t ldo, const int32_t *__restrict__ parent,
                                   int64_t n_fine, int C, float *__restrict__ d_in, int64_t ldd) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_fine * C) return;
  const int64_t i = t / C;
  const int c = (int)(t - i * C);
  const int64_t j = __ldg(parent + i);
  d_in[i * ldd + c] = __ldg(in + i * ldi + c) == __ldg(out + j * ldo + c) ? __ldg(g + j * ldg + c) : 0.f;
}

// dense (B, C, S, S, S): scatter == 1: dense[b][c][x][y][z] = feats[v][c]; scatter == 0: feats[v][c] = dense[...] (backward)
__global__ void sparse_dense_kernel(float *__restrict__ feats, int64_t ldf, const uint64_t *__restrict__ ukeys, int64_t n,
                                    int C, int64_t S, float *__restrict__ dense, int scatter) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * C) return;
  const int64_t v = t / C;
  const int c = (int)(t - v * C);
  int x, y, z, b;
  split_key(__ldg(ukeys + v), x, y, z, b);
  const int64_t at = ((((int64_t)b * C + c) * S + x) * S + y) * S + z;
  if (scatter) dense[at] = feats[v * ldf + c];
  else feats[v * ldf + c] = dense[at];
}

}  // namespace b200scn

extern "C" {

int b200scn_maxpool(const float *in, int64_t ldi, const int32_t *child, int64_t n_coarse, int K, int C, float *out,
                    int64_t ldo, void *stream) {
  if (n_coarse <= 0) return 0;
  maxpool_kernel<<<(unsigned)ceil_div(n_coarse * C, 256), 256, 0, (cudaStream_t)stream>>>(in, ldi, child, n_coarse, K, C,
                                                                                          out, ldo);
  SCN_CHECK_LAUNCH("maxpool");
  count_launch(1);
  return 0;
}

int b200scn_maxpool_bwd(const float *g, int64_t ldg, const float *in, int64_t ldi, const float *out, int64_t ldo,
                        const int32_t *parent, int64_t n_fine, int C, float *d_in, int64_t ldd, void *stream) {
  if (n_fine <= 0) return 0;
  maxpool_bwd_kernel<<<(unsigned)ceil_div(n_fine * C, 256), 256, 0, (cudaStream_t)stream>>>(g, ldg, in, ldi, out, ldo,
                                                                                           parent, n_fine, C, d_in, ldd);
  SCN_CHECK_LAUNCH("maxpool_bwd");
  count_launch(1);
  return 0;
}

int b200scn_sparse_to_dense(const float *feats, int64_t ldf, const uint64_t *ukeys, int64_t n, int C, int64_t spatial_size,
                            float *dense, void *stream) {
  if (n <= 0) return 0;
  sparse_dense_kernel<<<(unsigned)ceil_div(n * C, 256), 256, 0, (cudaStream_t)stream>>>(const_cast<float *>(feats), ldf,
                                                                                       ukeys, n, C, spatial_size, dense, 1);
  SCN_CHECK_LAUNCH("sparse_to_dense");
  count_launch(1);
  return 0;
}

int b200scn_sparse_to_dense_bwd(const float *d_dense, const uint64_t *ukeys, int64_t n, int C, int64_t spatial_size,
                                float *d_feats, int64_t ldf, void *stream) {
  if (n <= 0) return 0;
  sparse_dense_kernel<<<(unsigned)ceil_div(n * C, 256), 256, 0, (cudaStream_t)stream>>>(d_feats, ldf, ukeys, n, C, spatial_size,
                                                                                       const_cast<float *>(d_dense), 0);
  SCN_CHECK_LAUNCH("sparse_to_dense_bwd");
  count_launch(1);
  return 0;
}

}  // extern "C"
