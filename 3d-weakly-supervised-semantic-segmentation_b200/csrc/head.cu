// head.cu -- scene-level classification head fused after the per-scene mean pooling:
//   logits = pooled @ W^T + b           (nn.Linear(embed, NUM_CLASSES), models/MultiLabelContrastive.py:59,66)
//   loss   = F.multilabel_soft_margin_loss(logits, labels)   (utils/loss.py:21-30, the labels.ndim == 2 branch)
//          = mean_b mean_c -( y log sigmoid(x) + (1 - y) log sigmoid(-x) )
// One launch forward (logits + loss, deterministic: the last block to finish sums the per-scene terms in scene order),
// one launch backward (d_pooled, d_W, d_b from d_loss and/or an external d_logits).  B <= a few dozen scenes, C <= 448,
// 20 classes: latency-bound, so the point is launch count -- PyTorch runs ~12 kernels for the same chain.
#include "common.cuh"

namespace b200scn {

constexpr int kHeadThreads = 256;

__device__ __forceinline__ float log_sigmoid(float x) {   // = -softplus(-x), stable for both signs (ATen's formula)
  return fminf(x, 0.f) - log1pf(expf(-fabsf(x)));
}

__global__ void __launch_bounds__(kHeadThreads)
head_fwd_kernel(const float *__restrict__ pooled, const float *__restrict__ W, const float *__restrict__ bias,
                const float *__restrict__ labels, int B, int C, int NC, float *__restrict__ logits,
                float *__restrict__ loss_part, float *__restrict__ loss, unsigned *__restrict__ done) {
  extern __shared__ float hsm[];   // logits of this scene [NC]
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float *p = pooled + (int64_t)b * C;
  for (int c = warp; c < NC; c += kHeadThreads / 32) {
    const float *w = W + (int64_t)c * C;
    float acc = 0.f;
    for (int i = lane; i < C; i += 32) acc = fmaf(__ldg(p + i), __ldg(w + i), acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      const float v = acc + (bias ? __ldg(bias + c) : 0.f);
      hsm[c] = v;
      logits[(int64_t)b * NC + c] = v;
    }
  }
  __syncthreads();
  if (labels == nullptr) return;
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int c = 0; c < NC; ++c) {   // fixed order: bit-reproducible
      const float x = hsm[c], y = __ldg(labels + (int64_t)b * NC + c);
      s -= y * log_sigmoid(x) + (1.f - y) * log_sigmoid(-x);
    }
    loss_part[b] = s / (float)NC;
    __threadfence();
    if (atomicAdd(done, 1u) == (unsigned)B - 1) {   // last scene: ordered sum of the per-scene terms
      __threadfence();
      float t = 0.f;
      for (int i = 0; i < B; ++i) t += __ldcg(loss_part + i);
      *loss = t / (float)B;
      *done = 0;   // self-cleaning for the next call
    }
  }
}

// blocks [0, B): d_pooled[b,:] = sum_c dl[b,c] W[c,:];   blocks [B, B + NC): d_W[c,:] = sum_b dl[b,c] pooled[b,:],
// d_b[c] = sum_b dl[b,c];   dl[b,c] = d_logits[b,c] (if given) + d_loss * (sigmoid(x) - y) / (B NC) (if labels given)
__global__ void __launch_bounds__(kHeadThreads)
head_bwd_kernel(const float *__restrict__ pooled, const float *__restrict__ W, const float *__restrict__ labels,
                const float *__restrict__ logits, const float *__restrict__ d_loss, const float *__restrict__ d_logits,
                int B, int C, int NC, float *__restrict__ d_pooled, float *__restrict__ d_W, float *__restrict__ d_b) {
  extern __shared__ float hsm[];   // dl of this block's scene [NC] or class [B]
  const float gl = (labels && d_loss) ? __ldg(d_loss) / (float)(B * NC) : 0.f;
  auto dl = [&](int b, int c) {
    float v = d_logits ? __ldg(d_logits + (int64_t)b * NC + c) : 0.f;
    if (labels && d_loss) {
      const float x = __ldg(logits + (int64_t)b * NC + c);
      v += gl * (1.f / (1.f + expf(-x)) - __ldg(labels + (int64_t)b * NC + c));
    }
    return v;
  };
  if ((int)blockIdx.x < B) {
    const int b = blockIdx.x;
    for (int c = threadIdx.x; c < NC; c += kHeadThreads) hsm[c] = dl(b, c);
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += kHeadThreads) {
      float acc = 0.f;
      for (int c = 0; c < NC; ++c) acc = fmaf(hsm[c], __ldg(W + (int64_t)c * C + i), acc);
      d_pooled[(int64_t)b * C + i] = acc;
    }
  } else {
    const int c = blockIdx.x - B;
    for (int b = threadIdx.x; b < B; b += kHeadThreads) hsm[b] = dl(b, c);
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += kHeadThreads) {
      float acc = 0.f;
      for (int b = 0; b < B; ++b) acc = fmaf(hsm[b], __ldg(pooled + (int64_t)b * C + i), acc);
      d_W[(int64_t)c * C + i] = acc;
    }
    if (threadIdx.x == 0 && d_b) {
      float s = 0.f;
      for (int b = 0; b < B; ++b) s += hsm[b];
      d_b[c] = s;
    }
  }
}

}  // namespace b200scn

using namespace b200scn;

extern "C" {

int b200scn_head_multilabel(const float *pooled, const float *W, const float *bias, const float *labels, int B, int C,
                            int NC, float *logits, float *loss, float *scratch, void *stream) {
  if (B <= 0 || C <= 0 || NC <= 0 || NC > 1024) return set_error("head_multilabel: bad sizes B=%d C=%d NC=%d", B, C, NC);
  if (labels && (!loss || !scratch)) return set_error("head_multilabel: labels given without loss/scratch buffers");
  // scratch: B floats of per-scene loss terms followed by one unsigned completion counter, zeroed once by the caller
  head_fwd_kernel<<<B, kHeadThreads, sizeof(float) * NC, (cudaStream_t)stream>>>(
      pooled, W, bias, labels, B, C, NC, logits, scratch, loss, reinterpret_cast<unsigned *>(scratch + B));
  SCN_CHECK_LAUNCH("head_multilabel");
  count_launch(1);
  return 0;
}

int b200scn_head_multilabel_bwd(const float *pooled, const float *W, const float *labels, const float *logits,
                                const float *d_loss, const float *d_logits, int B, int C, int NC, float *d_pooled,
                                float *d_W, float *d_b, void *stream) {
  if (B <= 0 || C <= 0 || NC <= 0) return set_error("head_multilabel_bwd: bad sizes B=%d C=%d NC=%d", B, C, NC);
  const int sm = (int)sizeof(float) * (B > NC ? B : NC);
  if (sm > 48 * 1024) return set_error("head_multilabel_bwd: B=%d too large", B);
  head_bwd_kernel<<<B + NC, kHeadThreads, sm, (cudaStream_t)stream>>>(pooled, W, labels, logits, d_loss, d_logits, B, C,
                                                                      NC, d_pooled, d_W, d_b);
  SCN_CHECK_LAUNCH("head_multilabel_bwd");
  count_launch(1);
  return 0;
}

}  // extern "C"
