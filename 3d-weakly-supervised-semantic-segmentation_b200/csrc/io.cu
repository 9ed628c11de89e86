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
