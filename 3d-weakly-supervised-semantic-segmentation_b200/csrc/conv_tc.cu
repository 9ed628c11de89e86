// conv_tc.cu -- TF32 tensor-core gather convolution for sm_100a (tcgen05.mma, accumulators in TMEM).
//
//   out[o,:] = sum_k A[map[o][k],:] . W[k]      (replaces upstream scn's
//                                                 dConvolution_KMxKN_forward* fp32 FMA tiles)
//
// One CTA owns 128 output rows (one tcgen05 M=128 tile, N = Cout, accumulator = 128 lanes x Cout TMEM columns).
// For every kernel offset k that is present in the tile and every 32-channel K block:
//   warps 0-3 (producers)  gather the neighbour rows of A (16-byte loads, 8 lanes per 128-byte row, zeros for absent
//                          neighbours) and the matching slice of W[k] into a SWIZZLE_128B shared-memory stage,
//                          fence.proxy.async, arrive on the stage's "full" mbarrier;
//   warp 4, one lane       waits "full", issues Cin_block/8 tcgen05.mma.kind::tf32 (K = 8 each), tcgen05.commit ->
//                          the stage's "empty" mbarrier, so the producers run several stages ahead of the tensor pipe;
//   epilogue (warps 0-3)   tcgen05.ld the accumulator (warp w owns TMEM lanes 32w..32w+31 = output rows), adds the
//                          optional residual addend and writes every output row exactly once (no atomics, no
//                          read-modify-write across offsets).
// Weights arrive K-major: Wkm[k][n][c] (c contiguous), i.e. the B operand is read exactly as stored.
#include "common.cuh"
#include "tc_common.cuh"

namespace b200scn {

using namespace tc;

constexpr int kTcRows = 128;
constexpr int kTcThreads = 160;
constexpr int kMaxStages = 4;

struct TcSmemLayout {
  uint32_t a_bytes, b_bytes, stage_bytes, map_off, klist_off, bar_off, total;
};

static TcSmemLayout tc_layout(int Cout, int K, int nstages) {
  TcSmemLayout L;
  L.a_bytes = kTcRows * 128;
  L.b_bytes = (uint32_t)Cout * 128;
  L.stage_bytes = L.a_bytes + L.b_bytes;
  L.map_off = (uint32_t)nstages * L.stage_bytes;
  L.klist_off = L.map_off + (uint32_t)kTcRows * K * 4;
  L.bar_off = (L.klist_off + (uint32_t)(2 * K + 4) * 4 + 15) & ~15u;
  L.total = L.bar_off + (2 * kMaxStages + 1) * 8 + 16 + 1024;  // + alignment slack
  return L;
}

template <uint32_t NT>
__global__ void __launch_bounds__(kTcThreads)
gather_conv_tc_kernel(const float *__restrict__ A, int64_t lda, const int32_t *__restrict__ map, int n_rows, int K,
                      const float *__restrict__ Wkm, int Cin, int Cout, const float *__restrict__ addend,
                      int64_t ldadd, float *__restrict__ out, int64_t ldo, int nstages, uint32_t idesc,
                      uint32_t map_off, uint32_t klist_off, uint32_t bar_off) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *sm = smem_raw + (base - raw);
  const uint32_t a_bytes = kTcRows * 128, b_bytes = (uint32_t)Cout * 128;
  const uint32_t a_base = base, b_base = base + (uint32_t)nstages * a_bytes;
  int *smap = reinterpret_cast<int *>(sm + map_off);
  int *kflag = reinterpret_cast<int *>(sm + klist_off);
  int *klist = kflag + K;
  int *nk_p = klist + K;
  uint64_t *full = reinterpret_cast<uint64_t *>(sm + bar_off);
  uint64_t *empty = full + kMaxStages;
  uint64_t *accum = empty + kMaxStages;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(accum + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * kTcRows;

  for (int k = tid; k < K; k += kTcThreads) kflag[k] = 0;
  __syncthreads();
  for (int e = tid; e < kTcRows * K; e += kTcThreads) {
    int r = e / K, k = e - r * K;
    int v = -1;
    if (row0 + r < n_rows) v = map ? __ldg(map + (int64_t)(row0 + r) * K + k) : row0 + r;
    smap[e] = v;
    if (v >= 0) kflag[k] = 1;
  }
  __syncthreads();
  if (tid == 0) {
    int n = 0;
    for (int k = 0; k < K; ++k)
      if (kflag[k]) klist[n++] = k;
    *nk_p = n;
    for (int s = 0; s < nstages; ++s) {
      mbar_init(full + s, 128);
      mbar_init(empty + s, 1);
    }
    mbar_init(accum, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc<NT>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int nk = *nk_p;
  const int nkb = (Cin + 31) >> 5;
  const int T = nk * nkb;

  if (warp < 4) {
    // ------------------------------------------------------------ producers
    const int c = tid & 7, r0 = tid >> 3;
    for (int it = 0; it < T; ++it) {
      const int s = it % nstages;
      const uint32_t ph = (uint32_t)(it / nstages) & 1u;
      mbar_wait(empty + s, ph ^ 1u);
      const int k = klist[it / nkb], kb = it - (it / nkb) * nkb;
      const int chan = kb * 32 + c * 4;
      const bool cvalid = chan < Cin;
      const uint32_t a_st = a_base + (uint32_t)s * a_bytes, b_st = b_base + (uint32_t)s * b_bytes;
      float4 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int idx = smap[(r0 + 16 * i) * K + k];
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx >= 0 && cvalid) v[i] = ldg_f4(A + (int64_t)idx * lda + chan);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) sts_f4(a_st + sw128(r0 + 16 * i, c), v[i]);
      const float *wk = Wkm + (int64_t)k * Cout * Cin + chan;
      for (int n = r0; n < Cout; n += 16) {
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cvalid) w = ldg_f4(wk + (int64_t)n * Cin);
        sts_f4(b_st + sw128(n, c), w);
      }
      fence_proxy_async();
      mbar_arrive(full + s);
    }
  } else if (lane == 0) {
    // ------------------------------------------------------------ MMA issuer (one thread)
    for (int it = 0; it < T; ++it) {
      const int s = it % nstages;
      const uint32_t ph = (uint32_t)(it / nstages) & 1u;
      mbar_wait(full + s, ph);
      tc_fence_after();
      const int kb = it % nkb;
      const int kvalid = min(32, Cin - kb * 32);
      const uint32_t a_st = a_base + (uint32_t)s * a_bytes, b_st = b_base + (uint32_t)s * b_bytes;
      for (int j = 0; j < (kvalid >> 3); ++j) {
        const uint64_t ad = make_smem_desc(a_st + j * 32, 16, 1024);
        const uint64_t bd = make_smem_desc(b_st + j * 32, 16, 1024);
        mma_tf32(tmem, ad, bd, idesc, (it > 0 || j > 0) ? 1u : 0u);
      }
      mma_commit(empty + s);
    }
    mma_commit(accum);
  }

  if (warp < 4) {
    // ------------------------------------------------------------ epilogue: TMEM -> registers -> global
    if (T > 0) {
      mbar_wait(accum, 0);
      tc_fence_after();
    }
    const int row = row0 + warp * 32 + lane;
    for (int c0 = 0; c0 < Cout; c0 += 16) {
      float v[16];
      if (T > 0) {
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0.f;
      }
      if (row < n_rows) {
        if (addend) {
          const float *ad = addend + (int64_t)row * ldadd + c0;
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += __ldg(ad + i);
        }
        float *o = out + (int64_t)row * ldo + c0;
        if ((ldo & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<float4 *>(o + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = v[i];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<NT>(tmem);
}

bool gather_conv_tc_supported(const float *A, int64_t lda, int K, int Cin, int Cout, const float *W) {
  return (Cin % 8 == 0) && (Cout % 16 == 0) && Cout >= 16 && Cout <= 256 && (lda % 4 == 0) && K <= 64 &&
         ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
}

int gather_conv_tc(const float *A, int64_t lda, const int32_t *map, int64_t n_out, int K, const float *Wkm, int Cin,
                   int Cout, const float *addend, int64_t ldadd, float *out, int64_t ldo, cudaStream_t st) {
  if (n_out <= 0) return 0;
  // stage count: prefer two CTAs per SM (<= ~110 KB each), never below 2 stages
  int nstages = kMaxStages;
  TcSmemLayout L = tc_layout(Cout, K, nstages);
  while (nstages > 3 && L.total > 112 * 1024) L = tc_layout(Cout, K, --nstages);
  if (L.total > 112 * 1024) {
    nstages = kMaxStages;
    L = tc_layout(Cout, K, nstages);
    while (nstages > 2 && L.total > 224 * 1024) L = tc_layout(Cout, K, --nstages);
  }
  if (L.total > 227 * 1024) return set_error("gather_conv_tc: shared memory %u too large", L.total);
  const uint32_t idesc = make_idesc_tf32(128, Cout, 0, 0);
  dim3 grid((unsigned)ceil_div(n_out, kTcRows));
#define SCN_LAUNCH_TC(NT)                                                                                         \
  do {                                                                                                            \
    auto kern = gather_conv_tc_kernel<NT>;                                                                        \
    SCN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));              \
    kern<<<grid, kTcThreads, L.total, st>>>(A, lda, map, (int)n_out, K, Wkm, Cin, Cout, addend, ldadd, out, ldo,  \
                                            nstages, idesc, L.map_off, L.klist_off, L.bar_off);                   \
  } while (0)
  if (Cout <= 32) SCN_LAUNCH_TC(32);
  else if (Cout <= 64) SCN_LAUNCH_TC(64);
  else if (Cout <= 128) SCN_LAUNCH_TC(128);
  else SCN_LAUNCH_TC(256);
#undef SCN_LAUNCH_TC
  SCN_CHECK_LAUNCH("gather_conv_tc");
  count_launch(1);
  return 0;
}

}  // namespace b200scn
