// augment.cu -- the reference's per-step data path on the GPU: dataset/data.py:165-200 (trainMerge) and
// :266-290 (valMerge) as device kernels that end in the PACKED KEYS the voxeliser consumes, so the (sum P, 4) int64
// coordinate tensor is never built and only fp32 xyz + rgb cross PCIe.
//
// Per scene b (points scene_start[b] .. scene_start[b+1]), with M[b] (3x3), pre[b], r1[b], r2[b] drawn on the host exactly
// as the reference draws them (the random draws are a few numbers per scene and stay there):
//   a      = xyz . M[b]  [ + pre0 + pre[b] ]             (float64, like numpy's promotion of float32 xyz by a float64 matrix;
//                                                         valMerge adds full_scale/2 and then U(-2,2)^3, data.py:273)
//   lo, hi = a.min(0), a.max(0)                          (data.py:174-175 / 274-275)
//   offset = -lo + clip(S - (hi - lo) - 0.001, 0, None) * r1 + clip(S - (hi - lo) + 0.001, None, 0) * r2     (:177, form 0)
//          = -lo + clip(S - hi + lo - 0.001, 0, None) * r1 + clip(S - hi + lo + 0.001, None, 0) * r2         (:277, form 1)
//            (the two forms round differently in float64; each is evaluated in the reference's own order)
//   a     += offset;  keep rows with 0 <= a < S on all axes (:180 / :280);  coords = trunc(a) (:186 / :284, a >= 0)
// Output: keys[j] (b << 48 | x << 32 | y << 16 | z) and kept_rows[j] (index of the point in the input) for the kept rows
// in input order, their number on the device, per-scene kept counts (-> batch_offsets), and the offsets (the reference
// returns them, data.py:205).  Three passes over the points: min/max, (flag, scan), write -- all bandwidth-trivial.
#include "common.cuh"

namespace b200scn {

// order-preserving map double -> uint64 (so that atomicMin/Max on the integer orders the doubles)
__device__ __forceinline__ unsigned long long ord_of(double v) {
  unsigned long long u = (unsigned long long)__double_as_longlong(v);
  return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double dbl_of(unsigned long long o) {
  unsigned long long u = (o & 0x8000000000000000ull) ? (o & 0x7FFFFFFFFFFFFFFFull) : ~o;
  return __longlong_as_double((long long)u);
}

__device__ __forceinline__ int scene_of(const int32_t *__restrict__ start, int B, int64_t i) {
  int lo = 0, hi = B - 1;
  while (lo < hi) {   // last b with start[b] <= i
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(start + mid) <= i) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// a[j] = x0 m0j + x1 m1j + x2 m2j + pre_j in float64, accumulated in the order k = 0, 1, 2 with fused multiply-adds (the
// order and contraction of a BLAS dgemm micro-kernel, which is what numpy.matmul runs for (n,3) @ (3,3))
__device__ __forceinline__ void transform(const float *__restrict__ xyz, int64_t i, const double *__restrict__ M,
                                          double pre0, const double *__restrict__ pre, double a[3]) {
  const double x0 = (double)__ldg(xyz + 3 * i), x1 = (double)__ldg(xyz + 3 * i + 1), x2 = (double)__ldg(xyz + 3 * i + 2);
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    double acc = x0 * M[j];
    acc = fma(x1, M[3 + j], acc);
    acc = fma(x2, M[6 + j], acc);
    a[j] = pre ? __dadd_rn(__dadd_rn(acc, pre0), pre[j]) : acc;
  }
}

__global__ void augment_minmax_kernel(const float *__restrict__ xyz, int64_t P, const int32_t *__restrict__ start, int B,
                                      const double *__restrict__ mats, double pre0, const double *__restrict__ pre,
                                      unsigned long long *__restrict__ mm /*B x 6: min xyz, max xyz (ordered ints)*/) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const int b = scene_of(start, B, i);
  double a[3];
  transform(xyz, i, mats + 9 * b, pre0, pre ? pre + 3 * b : nullptr, a);
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    atomicMin(mm + 6 * b + j, ord_of(a[j]));
    atomicMax(mm + 6 * b + 3 + j, ord_of(a[j]));
  }
}

__global__ void augment_offset_kernel(const unsigned long long *__restrict__ mm, int B, const double *__restrict__ r1,
                                      const double *__restrict__ r2, double S, int form, double *__restrict__ offset) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 3 * B) return;
  const int b = t / 3, j = t - 3 * b;
  const double lo = dbl_of(mm[6 * b + j]), hi = dbl_of(mm[6 * b + 3 + j]);
  double c1, c2;
  if (form == 0) {   // trainMerge: length = M - m; full_scale - length -+ 0.001
    const double len = __dsub_rn(hi, lo);
    c1 = fmax(__dsub_rn(__dsub_rn(S, len), 0.001), 0.0);
    c2 = fmin(__dadd_rn(__dsub_rn(S, len), 0.001), 0.0);
  } else {           // valMerge: full_scale - M + m -+ 0.001
    const double t0 = __dadd_rn(__dsub_rn(S, hi), lo);
    c1 = fmax(__dsub_rn(t0, 0.001), 0.0);
    c2 = fmin(__dadd_rn(t0, 0.001), 0.0);
  }
  // -m + clip(..) * r1 + clip(..) * r2, evaluated left to right like the numpy expression (no contraction)
  offset[t] = __dadd_rn(__dadd_rn(-lo, __dmul_rn(c1, r1[t])), __dmul_rn(c2, r2[t]));
}

struct KeepLoader {
  const float *xyz; const int32_t *start; int B; const double *mats; double pre0; const double *pre, *offset; double S;
  __device__ int live(int n) const { return n; }
  __device__ int operator()(int64_t i) const {
    const int b = scene_of(start, B, i);
    double a[3];
    transform(xyz, i, mats + 9 * b, pre0, pre ? pre + 3 * b : nullptr, a);
    bool keep = true;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const double v = a[j] + offset[3 * b + j];
      keep = keep && v >= 0.0 && v < S;
    }
    return keep ? 1 : 0;
  }
};
struct KeepWriter {
  const float *xyz; const int32_t *start; int B; const double *mats; double pre0; const double *pre, *offset;
  uint64_t *keys; int32_t *kept_rows; int32_t *kept_per_scene;
  __device__ void operator()(int64_t i, int flag, int pos) const {
    if (!flag) return;
    const int b = scene_of(start, B, i);
    double a[3];
    transform(xyz, i, mats + 9 * b, pre0, pre ? pre + 3 * b : nullptr, a);
    const uint32_t x = (uint32_t)(long long)(a[0] + offset[3 * b]);       // truncation toward zero == floor (a >= 0)
    const uint32_t y = (uint32_t)(long long)(a[1] + offset[3 * b + 1]);
    const uint32_t z = (uint32_t)(long long)(a[2] + offset[3 * b + 2]);
    keys[pos] = make_key(x, y, z, (uint32_t)b);
    kept_rows[pos] = (int32_t)i;
    atomicAdd(kept_per_scene + b, 1);
  }
};

// out[j, :] = src[rows[j], :] (+ add[scene of rows[j], :]) for j < *n_dev : features of the kept points, with the reference's
// per-scene colour jitter (data.py:200: feats + torch.randn(3) * 0.1, one 3-vector per scene) folded in
__global__ void gather_rows_kernel(const float *__restrict__ src, int64_t lds, const int32_t *__restrict__ rows,
                                   const int32_t *__restrict__ n_dev, int64_t n_max, int C,
                                   const float *__restrict__ add, const int32_t *__restrict__ start, int B,
                                   float *__restrict__ out, int64_t ldo) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n = n_dev ? (int64_t)*n_dev : n_max;
  const int64_t j = t / C;
  const int c = (int)(t - j * C);
  if (j >= n) return;
  const int r = __ldg(rows + j);
  float v = __ldg(src + (int64_t)r * lds + c);
  if (add) v += __ldg(add + (int64_t)scene_of(start, B, r) * C + c);
  out[j * ldo + c] = v;
}

}  // namespace b200scn

using namespace b200scn;

extern "C" {

size_t b200scn_augment_scratch_bytes(int64_t P, int B) {
  return sizeof(unsigned long long) * 6 * (size_t)B + sizeof(int32_t) * scan_scratch_ints(P) + 256;
}

int b200scn_augment_voxelize(const float *xyz, int64_t P, const int32_t *scene_start, int B, const double *mats,
                             double pre0, const double *pre, const double *r1, const double *r2, int form,
                             int64_t spatial_size,
                             uint64_t *keys, int32_t *kept_rows, int32_t *n_kept_dev, int32_t *kept_per_scene,
                             double *offset_out, void *scratch, size_t scratch_bytes, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (B < 1 || B >= (1 << 15)) return set_error("augment_voxelize: %d scenes outside [1, 32768)", B);
  if (spatial_size < 1 || spatial_size > 65536) return set_error("augment_voxelize: spatial size %lld", (long long)spatial_size);
  if (P >= ((int64_t)1 << 31)) return set_error("augment_voxelize: too many points");
  if (scratch_bytes < b200scn_augment_scratch_bytes(P, B)) return set_error("augment_voxelize: scratch too small");
  unsigned long long *mm = reinterpret_cast<unsigned long long *>(scratch);
  int32_t *sums = reinterpret_cast<int32_t *>(mm + 6 * (size_t)B);
  SCN_CUDA(cudaMemsetAsync(kept_per_scene, 0, sizeof(int32_t) * B, st));
  if (P <= 0) {
    SCN_CUDA(cudaMemsetAsync(n_kept_dev, 0, sizeof(int32_t), st));
    return 0;
  }
  // min = +inf, max = -inf in the ordered encoding: all ones / all zeros per half
  SCN_CUDA(cudaMemset2DAsync(mm, 48, 0xFF, 24, (size_t)B, st));
  SCN_CUDA(cudaMemset2DAsync(reinterpret_cast<uint8_t *>(mm) + 24, 48, 0x00, 24, (size_t)B, st));
  augment_minmax_kernel<<<(unsigned)ceil_div(P, 256), 256, 0, st>>>(xyz, P, scene_start, B, mats, pre0, pre, mm);
  augment_offset_kernel<<<(unsigned)ceil_div(3 * B, 128), 128, 0, st>>>(mm, B, r1, r2, (double)spatial_size, form, offset_out);
  SCN_CHECK_LAUNCH("augment_minmax");
  count_launch(2);
  KeepLoader ld{xyz, scene_start, B, mats, pre0, pre, offset_out, (double)spatial_size};
  KeepWriter wr{xyz, scene_start, B, mats, pre0, pre, offset_out, keys, kept_rows, kept_per_scene};
  return scan_flags(ld, wr, P, nullptr, sums, n_kept_dev, st);
}

int b200scn_gather_rows(const float *src, int64_t lds, const int32_t *rows, const int32_t *n_dev, int64_t n_max, int C,
                        const float *add_per_scene, const int32_t *scene_start, int B, float *out, int64_t ldo,
                        void *stream) {
  if (n_max <= 0 || C <= 0) return 0;
  if (add_per_scene && (!scene_start || B < 1)) return set_error("gather_rows: per-scene addend needs scene_start");
  const int64_t total = n_max * C;
  gather_rows_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(src, lds, rows, n_dev, n_max, C,
                                                                                     add_per_scene, scene_start, B, out, ldo);
  SCN_CHECK_LAUNCH("gather_rows");
  count_launch(1);
  return 0;
}

}  // extern "C"
