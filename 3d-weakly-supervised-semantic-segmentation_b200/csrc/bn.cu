// bn.cu -- BatchNormReLU / BatchNormLeakyReLU over the active rows.
//
// Replaces upstream scn's BatchNormalization_f_train kernel, whose grid is nPlanes/32 blocks (ONE block
// at m=32).  Here the per-channel reductions run over a full grid (4 CTAs per SM), each CTA reducing a
// contiguous slab of rows with 16-byte loads, fp32 thread partials, double accumulation across CTAs.
// Pure HBM-bound: forward = read x twice + write y, backward = read x,dy twice + write dx.
// Two launches per direction: the reduction kernel's LAST block (ticket counter) finalises the statistics and leaves the
// accumulators zeroed for the next call (self-cleaning scratch: no memset, no separate finalize launch); then the apply
// kernel.  Sums are taken about a per-channel pivot (the first row) so that E[(x-p)^2] - E[x-p]^2 does not cancel when
// |mean| >> std.
#include "common.cuh"

namespace b200scn {

// nearest TF32 (the tensor core would truncate, a one-sided error): used when the output feeds a tcgen05 convolution
__device__ __forceinline__ float rna_tf32(float v) {
  uint32_t t;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v));
  return __uint_as_float(t);
}

__device__ __forceinline__ float bn_affine(float x, float mean, float invstd, float w, float b) {
  return fmaf((x - mean) * invstd, w, b);
}

struct BnCfg {
  int tpr;   // threads per row (each thread owns VEC consecutive channels)
  int rps;   // rows per block step
};
static inline BnCfg bn_cfg(int C, int vec) {
  BnCfg c;
  c.tpr = (int)ceil_div(C, vec);
  c.rps = 256 / c.tpr;
  if (c.rps < 1) c.rps = 1;
  return c;
}

struct BnFinal {
  // MODE 0 (forward, training statistics)
  float *running_mean, *running_var, *save_mean, *save_invstd;
  float momentum, eps;
  // MODE 1 (backward)
  float *d_weight, *d_bias;
};

// MODE 0: sums of (x-p) and (x-p)^2, p = row 0 (pivot).   MODE 1: sums of g and (x-mean)*g with g = dy * relu'(y).
// scratch: [0] ticket counter (as unsigned), [1 .. 2C] accumulators; all zero on entry, all zero again on exit.
template <int VEC, int MODE>
__global__ void __launch_bounds__(256)
bn_reduce_kernel(const float *__restrict__ x, int64_t ldx, const float *__restrict__ dy, int64_t lddy,
                 int64_t n, int C, int tpr, int rps, int64_t rows_per_block,
                 const float *__restrict__ mean, const float *__restrict__ invstd,
                 const float *__restrict__ weight, const float *__restrict__ bias, float leak,
                 double *__restrict__ scratch /*1 + 2*C*/, BnFinal fin) {
  extern __shared__ float red[];  // rps * C * 2
  __shared__ bool is_last;
  double *sums = scratch + 1;
  const int tid = threadIdx.x;
  const int rg = tid / tpr, ct = tid - rg * tpr;
  const int c0 = ct * VEC;
  const bool active = rg < rps && c0 < C;
  float s0[VEC], s1[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) s0[v] = s1[v] = 0.f;
  float m[VEC], is[VEC], w[VEC], b[VEC];
  if (active) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      if (MODE == 1) {
        m[v] = mean[c0 + v]; is[v] = invstd[c0 + v]; w[v] = weight[c0 + v]; b[v] = bias[c0 + v];
      } else {
        m[v] = __ldg(x + c0 + v);   // pivot: row 0
      }
    }
  }
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = min(r0 + rows_per_block, n);
  if (active) {
#pragma unroll 4
    for (int64_t r = r0 + rg; r < r1; r += rps) {
      float xv[VEC], gv[VEC];
      if (VEC == 4) {
        float4 t = __ldg(reinterpret_cast<const float4 *>(x + r * ldx + c0));
        xv[0] = t.x; xv[1] = t.y; xv[2] = t.z; xv[3] = t.w;
        if (MODE == 1) {
          float4 u = __ldg(reinterpret_cast<const float4 *>(dy + r * lddy + c0));
          gv[0] = u.x; gv[1] = u.y; gv[2] = u.z; gv[3] = u.w;
        }
      } else {
        xv[0] = __ldg(x + r * ldx + c0);
        if (MODE == 1) gv[0] = __ldg(dy + r * lddy + c0);
      }
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        if (MODE == 0) {
          const float d = xv[v] - m[v];
          s0[v] += d;
          s1[v] = fmaf(d, d, s1[v]);
        } else {
          float y = bn_affine(xv[v], m[v], is[v], w[v], b[v]);
          float g = y > 0.f ? gv[v] : leak * gv[v];
          s0[v] += g;
          s1[v] = fmaf(xv[v] - m[v], g, s1[v]);
        }
      }
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      red[(rg * C + c0 + v) * 2 + 0] = s0[v];
      red[(rg * C + c0 + v) * 2 + 1] = s1[v];
    }
  }
  __syncthreads();
  for (int e = tid; e < 2 * C; e += 256) {
    int c = e >> 1, which = e & 1;
    double acc = 0.0;
    for (int g = 0; g < rps; ++g) acc += (double)red[(g * C + c) * 2 + which];
    atomicAdd(sums + which * C + c, acc);
  }
  // last block to finish finalises and cleans up (threadfence reduction)
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned ticket = atomicAdd(reinterpret_cast<unsigned *>(scratch), 1u);
    is_last = ticket == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  for (int c = tid; c < C; c += 256) {
    const double a0 = __ldcg(sums + c), a1 = __ldcg(sums + C + c);
    sums[c] = 0.0;
    sums[C + c] = 0.0;
    if (MODE == 0) {
      const double d = a0 / (double)n;
      const double mu = (double)__ldg(x + c) + d;
      double var = a1 / (double)n - d * d;
      if (var < 0) var = 0;
      fin.save_mean[c] = (float)mu;
      fin.save_invstd[c] = (float)(1.0 / sqrt(var + (double)fin.eps));
      const double unbiased = n > 1 ? var * (double)n / (double)(n - 1) : var;
      fin.running_mean[c] = fin.momentum * fin.running_mean[c] + (1.f - fin.momentum) * (float)mu;
      fin.running_var[c] = fin.momentum * fin.running_var[c] + (1.f - fin.momentum) * (float)unbiased;
    } else {
      fin.d_bias[c] = (float)a0;
      fin.d_weight[c] = (float)(a1 * (double)invstd[c]);
    }
  }
  if (tid == 0) *reinterpret_cast<unsigned *>(scratch) = 0u;
}

__global__ void bn_eval_stats_kernel(int C, const float *running_mean, const float *running_var,
                                     float eps, float *save_mean, float *save_invstd) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  save_mean[c] = running_mean[c];
  save_invstd[c] = rsqrtf(running_var[c] + eps);
}

// y = leaky_relu(bn(x)): thread (tx, ty) owns VEC channels of rows ty, ty + rps, ... of its block's slab, so the
// per-channel constants live in registers and there is no index arithmetic beyond one add per row
template <int VEC>
__global__ void __launch_bounds__(256)
bn_apply_kernel(const float *__restrict__ x, int64_t ldx, int64_t n, int C, int tpr, int rps,
                int64_t rows_per_block, const float *__restrict__ mean, const float *__restrict__ invstd,
                const float *__restrict__ weight, const float *__restrict__ bias, float leak,
                float *__restrict__ y, int64_t ldy, int round_tf32) {
  const int tid = threadIdx.x;
  const int rg = tid / tpr, c0 = (tid - rg * tpr) * VEC;
  if (rg >= rps || c0 >= C) return;
  float m[VEC], is[VEC], w[VEC], b[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    m[v] = __ldg(mean + c0 + v); is[v] = __ldg(invstd + c0 + v); w[v] = __ldg(weight + c0 + v); b[v] = __ldg(bias + c0 + v);
  }
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = min(r0 + rows_per_block, n);
#pragma unroll 4
  for (int64_t r = r0 + rg; r < r1; r += rps) {
    if (VEC == 4) {
      float4 t = __ldg(reinterpret_cast<const float4 *>(x + r * ldx + c0));
      float4 o;
      o.x = bn_affine(t.x, m[0], is[0], w[0], b[0]); o.x = o.x > 0.f ? o.x : leak * o.x;
      o.y = bn_affine(t.y, m[1], is[1], w[1], b[1]); o.y = o.y > 0.f ? o.y : leak * o.y;
      o.z = bn_affine(t.z, m[2], is[2], w[2], b[2]); o.z = o.z > 0.f ? o.z : leak * o.z;
      o.w = bn_affine(t.w, m[3], is[3], w[3], b[3]); o.w = o.w > 0.f ? o.w : leak * o.w;
      if (round_tf32) { o.x = rna_tf32(o.x); o.y = rna_tf32(o.y); o.z = rna_tf32(o.z); o.w = rna_tf32(o.w); }
      *reinterpret_cast<float4 *>(y + r * ldy + c0) = o;
    } else {
      float o = bn_affine(__ldg(x + r * ldx + c0), m[0], is[0], w[0], b[0]);
      o = o > 0.f ? o : leak * o;
      y[r * ldy + c0] = round_tf32 ? rna_tf32(o) : o;
    }
  }
}

template <int VEC>
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float *__restrict__ x, int64_t ldx, const float *__restrict__ dy, int64_t lddy,
                    int64_t n, int C, int tpr, int rps, int64_t rows_per_block, const float *__restrict__ mean,
                    const float *__restrict__ invstd, const float *__restrict__ weight,
                    const float *__restrict__ bias, float leak, const float *__restrict__ d_weight,
                    const float *__restrict__ d_bias, int train, const float *__restrict__ addend, int64_t ldadd,
                    float *__restrict__ dx, int64_t lddx) {
  const int tid = threadIdx.x;
  const int rg = tid / tpr, c0 = (tid - rg * tpr) * VEC;
  if (rg >= rps || c0 >= C) return;
  const float inv_n = 1.f / (float)n;
  float m[VEC], is[VEC], w[VEC], b[VEC], k0[VEC], k1[VEC], sc[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    const int c = c0 + v;
    m[v] = __ldg(mean + c); is[v] = __ldg(invstd + c); w[v] = __ldg(weight + c); b[v] = __ldg(bias + c);
    // training statistics depend on x: d_in = (g - mean(g) - xhat * dotp * invstd / n) * invstd * w; with the running
    // statistics of eval mode they are constants and only the first term remains (upstream asserts train here)
    k0[v] = train ? d_bias[c] * inv_n : 0.f;                          // mean of g
    k1[v] = train ? d_weight[c] * is[v] * inv_n : 0.f;                // dotp * invstd^2 / n  (d_weight = dotp * invstd)
    sc[v] = is[v] * w[v];
  }
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = min(r0 + rows_per_block, n);
#pragma unroll 2
  for (int64_t r = r0 + rg; r < r1; r += rps) {
    float xv[VEC], gv[VEC], ov[VEC];
    if (VEC == 4) {
      float4 a = __ldg(reinterpret_cast<const float4 *>(x + r * ldx + c0));
      float4 g = __ldg(reinterpret_cast<const float4 *>(dy + r * lddy + c0));
      xv[0] = a.x; xv[1] = a.y; xv[2] = a.z; xv[3] = a.w;
      gv[0] = g.x; gv[1] = g.y; gv[2] = g.z; gv[3] = g.w;
    } else {
      xv[0] = __ldg(x + r * ldx + c0);
      gv[0] = __ldg(dy + r * lddy + c0);
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const float yv = bn_affine(xv[v], m[v], is[v], w[v], b[v]);
      const float g = yv > 0.f ? gv[v] : leak * gv[v];
      ov[v] = (g - k0[v] - (xv[v] - m[v]) * k1[v]) * sc[v];
    }
    if (addend) {   // gradient arriving at x through another consumer (the residual skip): summed here, not by a separate pass
      if (VEC == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4 *>(addend + r * ldadd + c0));
        ov[0] += t.x; ov[1] += t.y; ov[2] += t.z; ov[3] += t.w;
      } else {
        ov[0] += __ldg(addend + r * ldadd + c0);
      }
    }
    if (VEC == 4) *reinterpret_cast<float4 *>(dx + r * lddx + c0) = make_float4(ov[0], ov[1], ov[2], ov[3]);
    else dx[r * lddx + c0] = ov[0];
  }
}

static bool vec_ok(int C, int64_t ld, const void *p) {
  return C % 4 == 0 && ld % 4 == 0 && (reinterpret_cast<uintptr_t>(p) & 15) == 0;
}

template <int MODE>
static int launch_reduce(const float *x, int64_t ldx, const float *dy, int64_t lddy, int64_t n, int C,
                         const float *mean, const float *invstd, const float *weight, const float *bias,
                         float leak, double *sums, const BnFinal &fin, bool vec, cudaStream_t st) {
  BnCfg cfg = bn_cfg(C, vec ? 4 : 1);
  if (cfg.tpr > 256) return set_error("batchnorm: internal slice of %d planes too wide", C);
  // few enough CTAs that the 2*C double atomics per CTA do not pile up on the same addresses at small n
  int64_t blocks = kNumSMs * 4;
  int64_t rows_per_block = ceil_div(n, blocks);
  if (rows_per_block < (int64_t)cfg.rps * 16) rows_per_block = (int64_t)cfg.rps * 16;
  blocks = ceil_div(n, rows_per_block);
  size_t smem = sizeof(float) * 2 * (size_t)cfg.rps * C;
  if (smem > 48 * 1024) return set_error("batchnorm: shared memory %zu too large", smem);
  if (vec)
    bn_reduce_kernel<4, MODE><<<(unsigned)blocks, 256, smem, st>>>(x, ldx, dy, lddy, n, C, cfg.tpr, cfg.rps, rows_per_block, mean, invstd, weight, bias, leak, sums, fin);
  else
    bn_reduce_kernel<1, MODE><<<(unsigned)blocks, 256, smem, st>>>(x, ldx, dy, lddy, n, C, cfg.tpr, cfg.rps, rows_per_block, mean, invstd, weight, bias, leak, sums, fin);
  SCN_CHECK_LAUNCH("bn_reduce");
  count_launch(1);
  return 0;
}

// One call covers up to 1024 planes (256 when rows are not 16-byte aligned); wider inputs are processed in column slices
// (the scratch is self-cleaning, so consecutive slices reuse it).
static int bn_forward_slice(const float *x, int64_t ldx, int64_t n, int C, const float *weight, const float *bias,
                            float *running_mean, float *running_var, float *save_mean, float *save_invstd, float eps,
                            float momentum, int train, float leak, float *y, int64_t ldy, double *scratch, bool vec,
                            int round_tf32, cudaStream_t st) {
  const unsigned cb = (unsigned)ceil_div(C, 128);
  if (train && n > 0) {
    BnFinal fin = {running_mean, running_var, save_mean, save_invstd, momentum, eps, nullptr, nullptr};
    if (launch_reduce<0>(x, ldx, nullptr, 0, n, C, nullptr, nullptr, nullptr, nullptr, 0.f, scratch, fin, vec, st)) return 1;
  } else {
    bn_eval_stats_kernel<<<cb, 128, 0, st>>>(C, running_mean, running_var, eps, save_mean, save_invstd);
    count_launch(1);
  }
  if (n > 0) {
    BnCfg cfg = bn_cfg(C, vec ? 4 : 1);
    int64_t blocks = kNumSMs * 8;
    int64_t rows_per_block = ceil_div(n, blocks);
    if (rows_per_block < cfg.rps) rows_per_block = cfg.rps;
    blocks = ceil_div(n, rows_per_block);
    if (vec) bn_apply_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(x, ldx, n, C, cfg.tpr, cfg.rps, rows_per_block, save_mean, save_invstd, weight, bias, leak, y, ldy, round_tf32);
    else bn_apply_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(x, ldx, n, C, cfg.tpr, cfg.rps, rows_per_block, save_mean, save_invstd, weight, bias, leak, y, ldy, round_tf32);
    count_launch(1);
  }
  SCN_CHECK_LAUNCH("bn_forward");
  return 0;
}

static int bn_backward_slice(const float *x, int64_t ldx, const float *dy, int64_t lddy, int64_t n, int C,
                             const float *weight, const float *bias, const float *save_mean, const float *save_invstd,
                             float leak, int train, const float *addend, int64_t ldadd, float *dx, int64_t lddx,
                             float *d_weight, float *d_bias, double *scratch, bool vec, cudaStream_t st) {
  BnFinal fin = {nullptr, nullptr, nullptr, nullptr, 0.f, 0.f, d_weight, d_bias};
  if (launch_reduce<1>(x, ldx, dy, lddy, n, C, save_mean, save_invstd, weight, bias, leak, scratch, fin, vec, st)) return 1;
  BnCfg cfg = bn_cfg(C, vec ? 4 : 1);
  int64_t blocks = kNumSMs * 8;
  int64_t rows_per_block = ceil_div(n, blocks);
  if (rows_per_block < cfg.rps) rows_per_block = cfg.rps;
  blocks = ceil_div(n, rows_per_block);
  if (vec) bn_bwd_apply_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(x, ldx, dy, lddy, n, C, cfg.tpr, cfg.rps, rows_per_block, save_mean, save_invstd, weight, bias, leak, d_weight, d_bias, train, addend, ldadd, dx, lddx);
  else bn_bwd_apply_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(x, ldx, dy, lddy, n, C, cfg.tpr, cfg.rps, rows_per_block, save_mean, save_invstd, weight, bias, leak, d_weight, d_bias, train, addend, ldadd, dx, lddx);
  SCN_CHECK_LAUNCH("bn_backward");
  count_launch(1);
  return 0;
}

}  // namespace b200scn

using namespace b200scn;

extern "C" {

/* scratch doubles b200scn_bn_forward / _backward need for C planes: 1 ticket + 2 accumulators per plane of one slice */
size_t b200scn_bn_scratch_doubles(int C) { return 1 + 2 * (size_t)(C < 1024 ? C : 1024); }

int b200scn_bn_forward(const float *x, int64_t ldx, int64_t n, int C, const float *weight,
                       const float *bias, float *running_mean, float *running_var, float *save_mean,
                       float *save_invstd, float eps, float momentum, int train, float leak, float *y,
                       int64_t ldy, double *scratch, int round_tf32, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (C <= 0) return set_error("batchnorm: C must be positive");
  const bool vec = vec_ok(C, ldx, x) && vec_ok(C, ldy, y) && vec_ok(C, 4, weight) && vec_ok(C, 4, bias) &&
                   vec_ok(C, 4, save_mean) && vec_ok(C, 4, save_invstd);
  const int width = vec ? 1024 : 256;
  for (int c0 = 0; c0 < C; c0 += width) {
    const int cs = C - c0 < width ? C - c0 : width;
    if (bn_forward_slice(x + c0, ldx, n, cs, weight + c0, bias + c0, running_mean + c0, running_var + c0, save_mean + c0,
                         save_invstd + c0, eps, momentum, train, leak, y + c0, ldy, scratch, vec, round_tf32, st))
      return 1;
  }
  return 0;
}

int b200scn_bn_backward(const float *x, int64_t ldx, const float *dy, int64_t lddy, int64_t n, int C,
                        const float *weight, const float *bias, const float *save_mean,
                        const float *save_invstd, float leak, int train, const float *addend, int64_t ldadd,
                        float *dx, int64_t lddx, float *d_weight, float *d_bias, double *scratch, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (C <= 0) return set_error("batchnorm: C must be positive");
  if (n <= 0) {
    SCN_CUDA(cudaMemsetAsync(d_weight, 0, sizeof(float) * C, st));
    SCN_CUDA(cudaMemsetAsync(d_bias, 0, sizeof(float) * C, st));
    return 0;
  }
  const bool vec = vec_ok(C, ldx, x) && vec_ok(C, lddy, dy) && vec_ok(C, lddx, dx) && (!addend || vec_ok(C, ldadd, addend));
  const int width = vec ? 1024 : 256;
  for (int c0 = 0; c0 < C; c0 += width) {
    const int cs = C - c0 < width ? C - c0 : width;
    if (bn_backward_slice(x + c0, ldx, dy + c0, lddy, n, cs, weight + c0, bias + c0, save_mean + c0, save_invstd + c0, leak,
                          train, addend ? addend + c0 : nullptr, ldadd, dx + c0, lddx, d_weight + c0, d_bias + c0, scratch, vec,
                          st))
      return 1;
  }
  return 0;
}

}  // extern "C"
