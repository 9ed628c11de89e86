// conv_simt.cu -- fp32 CUDA-core gather-GEMM kernels (exact-fp32 path).
//
// Replaces upstream scn's dConvolution_KMxKN_forwardA/B + backward_dW kernels (27 launches per layer,
// output read-modify-written per offset).  Here one launch per layer/direction:
//   gather  : out[o]  = sum_k A[map[o][k]] . W[k]     output-stationary, each out row written once
//   scatter : out[map[j][k]] = A[j] . W[k]            input-stationary, each A tile reused over k
//   pair_dw : dW[k]   = sum_pairs A[pa]^T (x) G[pg]    split over pair chunks, fp32 atomics
// The TF32 tensor-core variants live in conv_tc.cu; this file is the precision==0 path and the path
// for shapes the tensor kernel does not cover (Cin = 3 stem).
#include "common.cuh"

namespace b200scn {

constexpr int BK = 16;

template <int BM, int BN, int MODE>  // MODE 0 gather, 1 scatter
__global__ void __launch_bounds__(256)
conv_simt_kernel(const float *__restrict__ A, int64_t lda, const int32_t *__restrict__ map, int n_rows,
                 int K, const float *__restrict__ W, int Cin, int Cout,
                 const float *__restrict__ addend, int64_t ldadd, float *__restrict__ out, int64_t ldo) {
  constexpr int TX = BN / 4;
  constexpr int TY = 256 / TX;
  constexpr int TM = BM / TY;
  static_assert(TM == 4, "thread tile is 4x4");
  extern __shared__ int dyn_smem[];
  int *smap = dyn_smem;            // BM*K
  int *kflag = dyn_smem + BM * K;  // K
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];

  const int tid = threadIdx.x;
  const int row0 = blockIdx.x * BM;
  const int col0 = blockIdx.y * BN;
  const int tx = tid % TX, ty = tid / TX;

  for (int k = tid; k < K; k += 256) kflag[k] = 0;
  __syncthreads();
  for (int e = tid; e < BM * K; e += 256) {
    int r = e / K, k = e - r * K;
    int v = -1;
    if (row0 + r < n_rows) v = map ? __ldg(map + (int64_t)(row0 + r) * K + k) : row0 + r;
    smap[e] = v;
    if (v >= 0) kflag[k] = 1;
  }
  __syncthreads();

  const bool a_vec = (Cin % 4 == 0) && (lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
  const bool b_vec = (Cout % 4 == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k = 0; k < K; ++k) {
    if (!kflag[k]) continue;  // block-uniform
    const float *Wk = W + (int64_t)k * Cin * Cout;
    for (int c0 = 0; c0 < Cin; c0 += BK) {
      // ---- A tile: BM rows x 16 channels, transposed into As[kk][r]
      for (int e = tid; e < BM * 4; e += 256) {
        int r = e >> 2, q = e & 3;
        int src = (MODE == 0) ? smap[r * K + k] : ((row0 + r < n_rows && smap[r * K + k] >= 0) ? row0 + r : -1);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        int c = c0 + q * 4;
        if (src >= 0 && c < Cin) {
          const float *p = A + (int64_t)src * lda + c;
          if (a_vec) {
            v = __ldg(reinterpret_cast<const float4 *>(p));
          } else {
            v.x = __ldg(p);
            if (c + 1 < Cin) v.y = __ldg(p + 1);
            if (c + 2 < Cin) v.z = __ldg(p + 2);
            if (c + 3 < Cin) v.w = __ldg(p + 3);
          }
        }
        As[q * 4 + 0][r] = v.x; As[q * 4 + 1][r] = v.y; As[q * 4 + 2][r] = v.z; As[q * 4 + 3][r] = v.w;
      }
      // ---- B tile: 16 x BN slice of W[k]
      for (int e = tid; e < BK * (BN / 4); e += 256) {
        int kk = e / (BN / 4), n4 = e - kk * (BN / 4);
        int ci = c0 + kk, co = col0 + n4 * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ci < Cin && co < Cout) {
          const float *p = Wk + (int64_t)ci * Cout + co;
          if (b_vec) {
            v = __ldg(reinterpret_cast<const float4 *>(p));
          } else {
            v.x = __ldg(p);
            if (co + 1 < Cout) v.y = __ldg(p + 1);
            if (co + 2 < Cout) v.z = __ldg(p + 2);
            if (co + 3 < Cout) v.w = __ldg(p + 3);
          }
        }
        *reinterpret_cast<float4 *>(&Bs[kk][n4 * 4]) = v;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float4 a = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
        float4 b = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
        float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
    if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int r = ty * 4 + i;
        int dst = smap[r * K + k];
        if (dst >= 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            int co = col0 + tx * 4 + j;
            if (co < Cout) out[(int64_t)dst * ldo + co] = acc[i][j];
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      }
    }
  }
  if (MODE == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int row = row0 + ty * 4 + i;
      if (row < n_rows) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int co = col0 + tx * 4 + j;
          if (co < Cout) {
            float v = acc[i][j];
            if (addend) v += __ldg(addend + (int64_t)row * ldadd + co);
            out[(int64_t)row * ldo + co] = v;
          }
        }
      }
    }
  }
}

// dW[k] tile (64 x 64) over one chunk of the pair list of offset k
__global__ void __launch_bounds__(256)
pair_dw_simt_kernel(const float *__restrict__ A, int64_t lda, const float *__restrict__ G, int64_t ldg,
                    const int32_t *__restrict__ pair_a, const int32_t *__restrict__ pair_g,
                    const int32_t *__restrict__ offsets, int n_single, int chunk, int Ca, int Cg,
                    int tiles_g, float *__restrict__ dW) {
  __shared__ __align__(16) float As[BK][64 + 4];
  __shared__ __align__(16) float Gs[BK][64 + 4];
  const int tid = threadIdx.x;
  const int k = blockIdx.y;
  const int ta = blockIdx.z / tiles_g, tg = blockIdx.z - ta * tiles_g;
  const int a0 = ta * 64, g0 = tg * 64;
  int beg = offsets ? offsets[k] : 0;
  int end = offsets ? offsets[k + 1] : n_single;
  int p0 = beg + blockIdx.x * chunk;
  int p1 = min(p0 + chunk, end);
  if (p0 >= p1) return;
  const int tx = tid & 15, ty = tid >> 4;
  const bool a_vec = (Ca % 4 == 0) && (lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
  const bool g_vec = (Cg % 4 == 0) && (ldg % 4 == 0) && ((reinterpret_cast<uintptr_t>(G) & 15) == 0);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int p = p0; p < p1; p += BK) {
    {
      int r = tid >> 4, q = tid & 15;  // 16 pairs x 16 float4
      int pp = p + r;
      float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vg = va;
      if (pp < p1) {
        int ra = pair_a ? __ldg(pair_a + pp) : pp;
        int rg = pair_g ? __ldg(pair_g + pp) : pp;
        int ca = a0 + q * 4, cg = g0 + q * 4;
        if (ca < Ca) {
          const float *pa = A + (int64_t)ra * lda + ca;
          if (a_vec) va = __ldg(reinterpret_cast<const float4 *>(pa));
          else {
            va.x = __ldg(pa);
            if (ca + 1 < Ca) va.y = __ldg(pa + 1);
            if (ca + 2 < Ca) va.z = __ldg(pa + 2);
            if (ca + 3 < Ca) va.w = __ldg(pa + 3);
          }
        }
        if (cg < Cg) {
          const float *pg = G + (int64_t)rg * ldg + cg;
          if (g_vec) vg = __ldg(reinterpret_cast<const float4 *>(pg));
          else {
            vg.x = __ldg(pg);
            if (cg + 1 < Cg) vg.y = __ldg(pg + 1);
            if (cg + 2 < Cg) vg.z = __ldg(pg + 2);
            if (cg + 3 < Cg) vg.w = __ldg(pg + 3);
          }
        }
      }
      *reinterpret_cast<float4 *>(&As[r][q * 4]) = va;
      *reinterpret_cast<float4 *>(&Gs[r][q * 4]) = vg;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
      float4 b = *reinterpret_cast<const float4 *>(&Gs[kk][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float *Wk = dW + (int64_t)k * Ca * Cg;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int ca = a0 + ty * 4 + i;
    if (ca >= Ca) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int cg = g0 + tx * 4 + j;
      if (cg < Cg) atomicAdd(Wk + (int64_t)ca * Cg + cg, acc[i][j]);
    }
  }
}

template <int VEC>
__global__ void __launch_bounds__(256)
unpool_kernel(const float *__restrict__ in, int64_t ldi, const int32_t *__restrict__ parent, int64_t n_fine, int C,
              float *__restrict__ out, int64_t ldo) {
  const int cv = C / VEC;
  const int64_t total = n_fine * cv;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / cv;
    const int c = (int)(t - i * cv) * VEC;
    const float *src = in + (int64_t)__ldg(parent + i) * ldi + c;
    if (VEC == 4) *reinterpret_cast<float4 *>(out + i * ldo + c) = __ldg(reinterpret_cast<const float4 *>(src));
    else out[i * ldo + c] = __ldg(src);
  }
}

template <int VEC>
__global__ void __launch_bounds__(256)
unpool_bwd_kernel(const float *__restrict__ d_out, int64_t ldd, const int32_t *__restrict__ child, int64_t n_coarse,
                  int K, int C, float *__restrict__ d_in, int64_t ldi) {
  const int cv = C / VEC;
  const int64_t total = n_coarse * cv;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = t / cv;
    const int c = (int)(t - j * cv) * VEC;
    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
    for (int k = 0; k < K; ++k) {
      const int i = __ldg(child + j * K + k);
      if (i < 0) continue;
      if (VEC == 4) {
        float4 a = __ldg(reinterpret_cast<const float4 *>(d_out + (int64_t)i * ldd + c));
        acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
      } else {
        acc[0] += __ldg(d_out + (int64_t)i * ldd + c);
      }
    }
    if (VEC == 4) *reinterpret_cast<float4 *>(d_in + j * ldi + c) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    else d_in[j * ldi + c] = acc[0];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Stem kernels (the Cin = 3 SubmanifoldConvolution at models/SparseConvNet.py:62 and its backward).
// K below the tensor-core granularity and pure bandwidth: one warp per site, lanes over the wide channel dimension.
// ---------------------------------------------------------------------------------------------------------------
// out[o][co] = sum_k sum_{ci<CI} A[map[o][k]][ci] * W[k][ci][co]          (CI <= 4, lanes over co)
template <int CI>
__global__ void __launch_bounds__(256)
gather_smallcin_kernel(const float *__restrict__ A, int64_t lda, const int32_t *__restrict__ map, int n_rows, int K,
                       const float *__restrict__ W, int Cout, const float *__restrict__ addend, int64_t ldadd,
                       float *__restrict__ out, int64_t ldo) {
  extern __shared__ float sw[];  // K*CI*Cout
  for (int e = threadIdx.x; e < K * CI * Cout; e += blockDim.x) sw[e] = W[e];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t o = warp0; o < n_rows; o += nwarps) {
    // lane k owns neighbour k: its index and its CI input values arrive in ONE dependent global round trip for the whole
    // site (a per-neighbour loop of broadcast loads costs one round trip per present neighbour)
    int mine = map ? (lane < K ? __ldg(map + o * K + lane) : -1) : (lane == 0 ? (int)o : -1);
    float av[CI];
#pragma unroll
    for (int ci = 0; ci < CI; ++ci) av[ci] = mine >= 0 ? __ldg(A + (int64_t)mine * lda + ci) : 0.f;
    const unsigned present = __ballot_sync(0xffffffffu, mine >= 0);
    for (int c0 = 0; c0 < Cout; c0 += 32) {
      const int co = c0 + lane;
      float acc = 0.f;
      for (unsigned m = present; m; m &= m - 1) {
        const int k = __ffs(m) - 1;
#pragma unroll
        for (int ci = 0; ci < CI; ++ci) {
          const float x = __shfl_sync(0xffffffffu, av[ci], k);
          if (co < Cout) acc = fmaf(x, sw[(k * CI + ci) * Cout + co], acc);
        }
      }
      if (co < Cout) {
        if (addend) acc += __ldg(addend + o * ldadd + co);
        out[o * ldo + co] = acc;
      }
    }
  }
}

// out[o][co] = sum_k sum_ci A[map[o][k]][ci] * W[k][ci][co]   with CO <= 4 outputs (backward-input of the stem):
// lanes over ci, warp reduction at the end
template <int CO>
__global__ void __launch_bounds__(256)
gather_smallcout_kernel(const float *__restrict__ A, int64_t lda, const int32_t *__restrict__ map, int n_rows, int K,
                        const float *__restrict__ W, int Cin, float *__restrict__ out, int64_t ldo) {
  extern __shared__ float sw[];  // K*Cin*CO
  for (int e = threadIdx.x; e < K * Cin * CO; e += blockDim.x) sw[e] = W[e];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t o = warp0; o < n_rows; o += nwarps) {
    const int32_t *mrow = map ? map + o * K : nullptr;
    int mine = (mrow && lane < K) ? __ldg(mrow + lane) : -1;
    float acc[CO];
#pragma unroll
    for (int j = 0; j < CO; ++j) acc[j] = 0.f;
    // present neighbours four at a time: their row loads are issued together (one round trip per four rows)
    unsigned m = mrow ? __ballot_sync(0xffffffffu, mine >= 0) : 1u;
    while (m) {
      int kk[4], id[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        kk[u] = m ? __ffs(m) - 1 : -1;
        m &= m - 1;
        const int sh = __shfl_sync(0xffffffffu, mine, kk[u] & 31);
        id[u] = kk[u] < 0 ? -1 : (mrow ? sh : (int)o);
      }
      for (int ci = lane; ci < Cin; ci += 32) {
        float x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) x[u] = id[u] >= 0 ? __ldg(A + (int64_t)id[u] * lda + ci) : 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (id[u] >= 0) {
#pragma unroll
            for (int j = 0; j < CO; ++j) acc[j] = fmaf(x[u], sw[(kk[u] * Cin + ci) * CO + j], acc[j]);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < CO; ++j) {
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], d);
    }
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < CO; ++j) out[o * ldo + j] = acc[j];
    }
  }
}

// dW[k][ca][cg] = sum_pairs A[pa][ca] * G[pg][cg]   with CA <= 4 (the stem's weight gradient): lanes over cg
template <int CA>
__global__ void __launch_bounds__(256)
pair_dw_smallca_kernel(const float *__restrict__ A, int64_t lda, const float *__restrict__ G, int64_t ldg,
                       const int32_t *__restrict__ pair_a, const int32_t *__restrict__ pair_g,
                       const int32_t *__restrict__ offsets, int n_single, int chunk, int Cg, float *__restrict__ dW) {
  const int k = blockIdx.y;
  const int beg = offsets ? offsets[k] : 0;
  const int end = offsets ? offsets[k + 1] : n_single;
  const int p0 = beg + blockIdx.x * chunk;
  const int p1 = min(p0 + chunk, end);
  if (p0 >= p1) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int c0 = 0; c0 < Cg; c0 += 32) {
    const int cg = c0 + lane;
    float acc[CA];
#pragma unroll
    for (int j = 0; j < CA; ++j) acc[j] = 0.f;
    for (int pb = p0 + warp * 4; pb < p1; pb += nw * 4) {   // four pairs per trip: index loads, then data loads, then FMAs
      int ra[4], rg[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int p = pb + u;
        ra[u] = p < p1 ? (pair_a ? __ldg(pair_a + p) : p) : -1;
        rg[u] = p < p1 ? (pair_g ? __ldg(pair_g + p) : p) : 0;
      }
      float g[4], a[4][CA];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        g[u] = (ra[u] >= 0 && cg < Cg) ? __ldg(G + (int64_t)rg[u] * ldg + cg) : 0.f;
#pragma unroll
        for (int j = 0; j < CA; ++j) a[u][j] = ra[u] >= 0 ? __ldg(A + (int64_t)ra[u] * lda + j) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < CA; ++j) acc[j] = fmaf(a[u][j], g[u], acc[j]);
    }
    if (cg < Cg) {
#pragma unroll
      for (int j = 0; j < CA; ++j) atomicAdd(dW + ((int64_t)k * CA + j) * Cg + cg, acc[j]);
    }
  }
}

template <int MODE>
static int launch_conv_simt(const float *A, int64_t lda, const int32_t *map, int64_t n_rows, int K,
                            const float *W, int Cin, int Cout, const float *addend, int64_t ldadd,
                            float *out, int64_t ldo, cudaStream_t st) {
  if (n_rows <= 0) return 0;
  if (Cout <= 32) {
    constexpr int BM = 128, BN = 32;
    size_t dyn = sizeof(int) * ((size_t)BM * K + K);
    auto kern = conv_simt_kernel<BM, BN, MODE>;
    if (dyn > 40 * 1024) SCN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    dim3 grid((unsigned)ceil_div(n_rows, BM), (unsigned)ceil_div(Cout, BN));
    kern<<<grid, 256, dyn, st>>>(A, lda, map, (int)n_rows, K, W, Cin, Cout, addend, ldadd, out, ldo);
  } else {
    constexpr int BM = 64, BN = 64;
    size_t dyn = sizeof(int) * ((size_t)BM * K + K);
    auto kern = conv_simt_kernel<BM, BN, MODE>;
    dim3 grid((unsigned)ceil_div(n_rows, BM), (unsigned)ceil_div(Cout, BN));
    kern<<<grid, 256, dyn, st>>>(A, lda, map, (int)n_rows, K, W, Cin, Cout, addend, ldadd, out, ldo);
  }
  SCN_CHECK_LAUNCH("conv_simt");
  count_launch(1);
  return 0;
}

int gather_conv_simt(const float *A, int64_t lda, const int32_t *map, int64_t n_out, int K,
                     const float *W, int Cin, int Cout, const float *addend, int64_t ldadd, float *out,
                     int64_t ldo, cudaStream_t st) {
  if (n_out > 0 && K <= 32 && (Cin <= 4 || (Cout <= 4 && !addend))) {
    const unsigned blocks = (unsigned)min((int64_t)kNumSMs * 16, ceil_div(n_out, 8));
    const size_t smem = sizeof(float) * (size_t)K * Cin * Cout;
    if (smem <= 48 * 1024) {
      if (Cin <= 4) {
        switch (Cin) {
          case 1: gather_smallcin_kernel<1><<<blocks, 256, smem, st>>>(A, lda, map, (int)n_out, K, W, Cout, addend, ldadd, out, ldo); break;
          case 2: gather_smallcin_kernel<2><<<blocks, 256, smem, st>>>(A, lda, map, (int)n_out, K, W, Cout, addend, ldadd, out, ldo); break;
          case 3: gather_smallcin_kernel<3><<<blocks, 256, smem, st>>>(A, lda, map, (int)n_out, K, W, Cout, addend, ldadd, out, ldo); break;
          default: gather_smallcin_kernel<4><<<blocks, 256, smem, st>>>(A, lda, map, (int)n_out, K, W, Cout, addend, ldadd, out, ldo); break;
        }
      } else {
        switch (Cout) {
          case 1: gather_smallcout_kernel<1><<<blocks, 256, smem, st>>>(A, lda, map, (int)n_out, K, W, Cin, out, ldo); break;
          case 2: gather_smallcout_kernel<2><<<blocks, 256, smem, st>>>(A, lda, map, (int)n_out, K, W, Cin, out, ldo); break;
          case 3: gather_smallcout_kernel<3><<<blocks, 256, smem, st>>>(A, lda, map, (int)n_out, K, W, Cin, out, ldo); break;
          default: gather_smallcout_kernel<4><<<blocks, 256, smem, st>>>(A, lda, map, (int)n_out, K, W, Cin, out, ldo); break;
        }
      }
      SCN_CHECK_LAUNCH("gather_small");
      count_launch(1);
      return 0;
    }
  }
  return launch_conv_simt<0>(A, lda, map, n_out, K, W, Cin, Cout, addend, ldadd, out, ldo, st);
}
int scatter_conv_simt(const float *A, int64_t lda, const int32_t *map, int64_t n_in, int K,
                      const float *W, int Cin, int Cout, float *out, int64_t ldo, cudaStream_t st) {
  return launch_conv_simt<1>(A, lda, map, n_in, K, W, Cin, Cout, nullptr, 0, out, ldo, st);
}

int pair_dw_simt(const float *A, int64_t lda, const float *G, int64_t ldg, const int32_t *pair_a,
                 const int32_t *pair_g, const int32_t *offsets_dev, int K, int64_t n_pairs_max, int Ca,
                 int Cg, float *dW, cudaStream_t st) {
  SCN_CUDA(cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)K * Ca * Cg, st));
  if (n_pairs_max <= 0) return 0;
  if (Ca <= 4) {
    int64_t want = ceil_div((int64_t)kNumSMs * 8, (int64_t)K);
    int64_t ch = ceil_div(n_pairs_max, want > 0 ? want : 1);
    if (ch < 256) ch = 256;
    if (ch > 2048) ch = 2048;   // n_pairs_max only bounds the lists (most are ~10x shorter): small chunks keep every SM busy
    dim3 g((unsigned)ceil_div(n_pairs_max, ch), (unsigned)K);
    switch (Ca) {
      case 1: pair_dw_smallca_kernel<1><<<g, 256, 0, st>>>(A, lda, G, ldg, pair_a, pair_g, offsets_dev, (int)n_pairs_max, (int)ch, Cg, dW); break;
      case 2: pair_dw_smallca_kernel<2><<<g, 256, 0, st>>>(A, lda, G, ldg, pair_a, pair_g, offsets_dev, (int)n_pairs_max, (int)ch, Cg, dW); break;
      case 3: pair_dw_smallca_kernel<3><<<g, 256, 0, st>>>(A, lda, G, ldg, pair_a, pair_g, offsets_dev, (int)n_pairs_max, (int)ch, Cg, dW); break;
      default: pair_dw_smallca_kernel<4><<<g, 256, 0, st>>>(A, lda, G, ldg, pair_a, pair_g, offsets_dev, (int)n_pairs_max, (int)ch, Cg, dW); break;
    }
    SCN_CHECK_LAUNCH("pair_dw_smallca");
    count_launch(1);
    return 0;
  }
  const int tiles_a = (int)ceil_div(Ca, 64), tiles_g = (int)ceil_div(Cg, 64);
  // aim for ~8 CTAs per SM overall; chunk is a multiple of BK
  int64_t want_chunks = ceil_div((int64_t)kNumSMs * 8, (int64_t)K * tiles_a * tiles_g);
  int64_t chunk = ceil_div(n_pairs_max, want_chunks > 0 ? want_chunks : 1);
  if (chunk < 256) chunk = 256;
  if (chunk > 8192) chunk = 8192;
  chunk = ceil_div(chunk, BK) * BK;
  dim3 grid((unsigned)ceil_div(n_pairs_max, chunk), (unsigned)K, (unsigned)(tiles_a * tiles_g));
  pair_dw_simt_kernel<<<grid, 256, 0, st>>>(A, lda, G, ldg, pair_a, pair_g, offsets_dev, (int)n_pairs_max,
                                            (int)chunk, Ca, Cg, tiles_g, dW);
  SCN_CHECK_LAUNCH("pair_dw_simt");
  count_launch(1);
  return 0;
}

}  // namespace b200scn

using namespace b200scn;

extern "C" {

static bool vec4_ok(int C, int64_t lda, int64_t ldb, const void *a, const void *b) {
  return C % 4 == 0 && lda % 4 == 0 && ldb % 4 == 0 &&
         ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
}

int b200scn_unpool(const float *in, int64_t ldi, const int32_t *parent, int64_t n_fine, int C,
                   float *out, int64_t ldo, void *stream) {
  if (n_fine <= 0) return 0;
  const bool vec = vec4_ok(C, ldi, ldo, in, out);
  const int64_t total = n_fine * (vec ? C / 4 : C);
  const unsigned blocks = (unsigned)min((int64_t)1 << 30, ceil_div(total, 256));   // one element group per thread
  if (vec) unpool_kernel<4><<<blocks, 256, 0, (cudaStream_t)stream>>>(in, ldi, parent, n_fine, C, out, ldo);
  else unpool_kernel<1><<<blocks, 256, 0, (cudaStream_t)stream>>>(in, ldi, parent, n_fine, C, out, ldo);
  SCN_CHECK_LAUNCH("unpool");
  count_launch(1);
  return 0;
}

int b200scn_unpool_bwd(const float *d_out, int64_t ldd, const int32_t *child, int64_t n_coarse, int K,
                       int C, float *d_in, int64_t ldi, void *stream) {
  if (n_coarse <= 0) return 0;
  const bool vec = vec4_ok(C, ldd, ldi, d_out, d_in);
  const int64_t total = n_coarse * (vec ? C / 4 : C);
  const unsigned blocks = (unsigned)min((int64_t)1 << 30, ceil_div(total, 256));   // one element group per thread
  if (vec) unpool_bwd_kernel<4><<<blocks, 256, 0, (cudaStream_t)stream>>>(d_out, ldd, child, n_coarse, K, C, d_in, ldi);
  else unpool_bwd_kernel<1><<<blocks, 256, 0, (cudaStream_t)stream>>>(d_out, ldd, child, n_coarse, K, C, d_in, ldi);
  SCN_CHECK_LAUNCH("unpool_bwd");
  count_launch(1);
  return 0;
}

}  // extern "C"
