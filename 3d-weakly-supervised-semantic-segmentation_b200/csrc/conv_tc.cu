// conv_tc.cu -- TF32 tensor-core kernels for sm_100a (tcgen05.mma, accumulators in TMEM).
//
//   gather conv:  out[o,:] = sum_k A[map[o][k],:] . W[k]      (replaces upstream scn's
//                                                               dConvolution_KMxKN_forward* fp32 FMA tiles)
//
// One CTA owns MSUB x 128 output rows (MSUB tcgen05 M=128 tiles sharing every weight stage, N = Cout, accumulators =
// 128 lanes x MSUB*Cout TMEM columns).  For every kernel offset k present in the tile and every 32-channel K block:
//   warps 0-3 (producers)  each warp owns one pipeline stage: it gathers the neighbour rows of A with 16-byte cp.async
//                          (8 lanes per 128-byte row, zero-fill for absent neighbours) plus the matching slice of W[k]
//                          into a SWIZZLE_128B shared-memory image, then fence.proxy.async + mbarrier arrive ("full");
//   warp 4, one lane       waits "full", issues MSUB * Cin_block/8 tcgen05.mma.kind::tf32 (K = 8 each) and
//                          tcgen05.commit -> the stage's "empty" mbarrier, so gathers run stages ahead of the tensor pipe;
//   epilogue (warps 0-3)   tcgen05.ld the accumulators (warp w owns TMEM lanes 32w..32w+31 = output rows), adds the
//                          optional residual addend and writes every output row exactly once.
// Tiny levels (fewer tiles than SMs) additionally split the offsets over gridDim.y CTAs that reduce into a zeroed
// output with fp32 atomics, because there the kernel is bound by the length of one CTA's dependent stage chain.
// Weights arrive K-major: Wkm[k][n][c] (c contiguous), i.e. the B operand is read exactly as stored.
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace b200scn {

using namespace tc;

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D fp32 tensor map: rows x cols (cols contiguous, row pitch ld floats), box = box_rows x 32 floats, 128-byte swizzle
// cuTensorMapEncodeTiled costs a few microseconds of host time per call and the weight operands of a step are prepared
// into buffers the caching allocator hands back at the same addresses step after step: remember the last encodings
// (a tensor map describes address + shape only, never contents, so a hit is always valid).
struct TmapKey { const float *base; int64_t rows, ld; int cols, box_rows; };
static thread_local struct { TmapKey key; CUtensorMap map; bool used; } g_tmap_cache[256];

static int make_tmap(CUtensorMap *m, const float *base, int64_t rows, int cols, int64_t ld, int box_rows) {
  const uint64_t h = (reinterpret_cast<uintptr_t>(base) >> 8) * 0x9E3779B97F4A7C15ull + (uint64_t)rows * 31 + (uint64_t)box_rows;
  auto &slot = g_tmap_cache[(h >> 32) & 255];
  if (slot.used && slot.key.base == base && slot.key.rows == rows && slot.key.ld == ld && slot.key.cols == cols &&
      slot.key.box_rows == box_rows) {
    *m = slot.map;
    return 0;
  }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error("cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error("cuTensorMapEncodeTiled failed (%d): rows %lld cols %d ld %lld box %d", (int)r,
                                          (long long)rows, cols, (long long)ld, box_rows);
  slot.key = TmapKey{base, rows, ld, cols, box_rows};
  slot.map = *m;
  slot.used = true;
  return 0;
}

// shared with conv_halo.cu
int make_weight_tmap(CUtensorMap *m, const float *base, int64_t rows, int cols, int64_t ld, int box_rows) {
  return make_tmap(m, base, rows, cols, ld, box_rows);
}

constexpr int kTcThreads = 160;
constexpr int kMaxStages = 4;

struct TcSmemLayout {
  uint32_t a_bytes, b_bytes, stage_bytes, map_off, klist_off, bar_off, total;
};

static TcSmemLayout tc_layout(int msub, int Cout, int K, int nstages) {
  TcSmemLayout L;
  L.a_bytes = (uint32_t)msub * 128 * 128;
  L.b_bytes = (uint32_t)Cout * 128;
  L.stage_bytes = L.a_bytes + L.b_bytes;
  L.map_off = (uint32_t)nstages * L.stage_bytes;
  L.klist_off = L.map_off + (uint32_t)msub * 128 * K * 4;
  L.bar_off = (L.klist_off + (uint32_t)(2 * K + 4) * 4 + 15) & ~15u;
  L.total = L.bar_off + (2 * kMaxStages + 1) * 8 + 16 + 1024;  // + alignment slack
  return L;
}

// NPW producer warps, NPW/4 per stage, each gathering an equal share of the stage's rows and weight rows
// (the producers are bound by their own dependent issue chain, so 16 warps beat 4 by ~3x: profiles/).
template <uint32_t NT, int MSUB, int NPW>
__global__ void __launch_bounds__(32 * (NPW + 2))
gather_conv_tc_kernel(const __grid_constant__ CUtensorMap tmW, const float *__restrict__ A, int64_t lda,
                      const int32_t *__restrict__ map, int n_rows, int K, int Cin, int Cout,
                      const float *__restrict__ addend,
                      int64_t ldadd, float *__restrict__ out, int64_t ldo, int nstages, uint32_t idesc,
                      uint32_t map_off, uint32_t klist_off, uint32_t bar_off, int w_rows_per_k, int w_row0,
                      const int32_t *__restrict__ grp_in, const int32_t *__restrict__ grp_out,
                      const int4 *__restrict__ grp_tab) {
  constexpr int ROWS = MSUB * 128;
  constexpr int NTHREADS = 32 * (NPW + 2);   // producers, MMA-issuing warp, weight-TMA warp
  // GROUPED mode (strided Deconvolution forward / Convolution backward-input on the fine side, where every output row
  // has exactly ONE rule): the rules arrive sorted by offset (the scn-form rulebook), a tile is a run of <= 128 rules of
  // ONE offset -- grp_tab[tile] = {offset, first rule, count} -- so the tile is a single dense K-slice: input row
  // grp_in[rule], output row grp_out[rule].  (The one-hot map this replaces wasted 7/8 of every staged A tile.)
  const bool grouped = grp_tab != nullptr;
  const int KM = grouped ? 1 : K;             // columns of the shared-memory map slice
  int g_k = 0, g_start = 0, g_cnt = ROWS;
  if (grouped) {
    const int4 t = __ldg(grp_tab + blockIdx.x);
    g_k = t.x; g_start = t.y; g_cnt = t.z;
    if (g_cnt <= 0) return;                   // table slack beyond the live tiles (uniform for the CTA)
  }
  constexpr int WPS = NPW / 4;  // producer warps per stage
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *sm = smem_raw + (base - raw);
  const uint32_t a_bytes = ROWS * 128, b_bytes = (uint32_t)Cout * 128;
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
  const int row0 = blockIdx.x * ROWS;
  const int nsplit = gridDim.y, split = blockIdx.y;

  // The tile's slice of the map is one contiguous run of ROWS*K ints.  Full tiles copy it with 16-byte cp.async, all
  // in flight at once (a dependent load->store loop with a division per element cost ~15 % of the tile, measured with a
  // clock64 timeline of one CTA); the ragged last tile takes the scalar path.
  if (grouped) {
    for (int e = tid; e < ROWS; e += NTHREADS) smap[e] = e < g_cnt ? __ldg(grp_in + g_start + e) : -1;
  } else {
    const int live = min(ROWS, n_rows - row0) * K;   // entries that belong to real rows
    const int32_t *msrc = map ? map + (int64_t)row0 * K : nullptr;
    if (msrc && live == ROWS * K && ((ROWS * K) & 3) == 0 && (reinterpret_cast<uintptr_t>(msrc) & 15) == 0) {
      for (int e = tid * 4; e < ROWS * K; e += NTHREADS * 4) cp_async16(smem_u32(smap + e), msrc + e, 16u);
      cp_async_wait_all();
    } else {
      for (int e = tid; e < ROWS * K; e += NTHREADS)
        smap[e] = e < live ? (msrc ? __ldg(msrc + e) : row0 + e) : -1;
    }
  }
  __syncthreads();
  // offset k is present in this tile if any of its ROWS entries is: one warp-strided pass per offset
  if (!grouped) {
    for (int k = warp; k < K; k += NTHREADS / 32) {
      int any = 0;
      for (int r = lane; r < ROWS; r += 32) any |= (smap[r * K + k] >= 0);
      any = __any_sync(0xffffffffu, any);
      if (lane == 0) kflag[k] = any;
    }
  }
  __syncthreads();
  if (tid == 0) {
    int n = 0;
    if (grouped) klist[n++] = g_k;
    else
      for (int k = 0; k < K; ++k)
        if (kflag[k]) klist[n++] = k;
    // this CTA's share of the present offsets
    const int per = (n + nsplit - 1) / nsplit;
    const int lo = min(n, split * per), hi = min(n, lo + per);
    for (int i = lo; i < hi; ++i) klist[i - lo] = klist[i];
    *nk_p = hi - lo;
    for (int s = 0; s < nstages; ++s) {
      mbar_init(full + s, 32 * WPS + 1);   // every gathering lane + the weight TMA's expect_tx arrival
      mbar_init(empty + s, 1);
    }
    mbar_init(accum, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmW);
  }
  if (warp == NPW) tmem_alloc<NT>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int nk = *nk_p;
  const int nkb = (Cin + 31) >> 5;
  const int T = nk * nkb;

  if (warp < NPW) {
    // ------------------------------------------------------------ producers
    // stage s is always filled by the same warp(s) (s, and s + nstages when WPS == 2): consecutive uses of a stage
    // are ordered by that warp's program order, so the one-bit mbarrier phase parity can never alias.  Every lane
    // keeps ROWS/(4*WPS) (+ Cout/(4*WPS)) 16-byte copies in flight and the warps work on different stages at once.
    const int c = lane & 7, rl = lane >> 3;
    const int my_stage = warp % nstages, half = warp / nstages;
    if (half < WPS) {
      constexpr int RPW = ROWS / WPS;  // rows per producer warp
      constexpr int NI = RPW / 4;      // row slots per lane (8 lanes share a row)
      const int rbase = half * RPW;
      // Absent neighbours need a ZERO row, not a copy: `dirty` remembers which of this lane's row slots of this stage
      // buffer currently hold data (the warp owns the same rows of the same buffer for the whole tile), so a slot is
      // touched only when it receives data or when stale data must be cleared -- at 2 cm voxels 60-80 % of the
      // (row, offset) slots are absent and stay untouched.
      uint32_t dirty = 0xFFFFFFFFu;    // unknown contents on first use
      // per-lane constants of the tile (the producers are bound by their own issue chain, so keep the loop lean)
      uint32_t soff[NI];               // swizzled byte offset of this lane's chunk in each of its rows
      const int *mrow[NI];             // this lane's rows of the map slice
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int r = rbase + rl + 4 * i;
        soff[i] = sw128(r, c);
        mrow[i] = smap + r * KM;
      }
      const int64_t lda4 = lda;
      for (int it = my_stage; it < T; it += nstages) {
        const int s = my_stage;
        const uint32_t ph = (uint32_t)(it / nstages) & 1u;
        mbar_wait_sleep(empty + s, ph ^ 1u, 500);
        const int k = klist[it / nkb], kb = it - (it / nkb) * nkb;
        const int chan = kb * 32 + c * 4;
        const uint32_t a_st = a_base + (uint32_t)s * a_bytes, b_st = b_base + (uint32_t)s * b_bytes;
        if (chan < Cin) {
          const float *acol = A + chan;
          int idx[NI];
#pragma unroll
          for (int i = 0; i < NI; ++i) idx[i] = mrow[i][grouped ? 0 : k];   // all map reads first, then the copies
#pragma unroll
          for (int i = 0; i < NI; ++i) {
            const bool valid = idx[i] >= 0;
            if (valid || ((dirty >> i) & 1u))
              cp_async16(a_st + soff[i], valid ? (const void *)(acol + (int64_t)idx[i] * lda4) : (const void *)A,
                         valid ? 16u : 0u);
            dirty = valid ? (dirty | (1u << i)) : (dirty & ~(1u << i));
          }
        }
        // (measured and rejected, profiles/r1_experiments.md: L2 prefetch one round ahead, software-pipelined producer
        //  groups, one elected mbarrier arrival per warp, two stages with more CTAs per SM)
        cp_async_wait_all();
        fence_proxy_async();
        mbar_arrive(full + s);
      }
    }
  } else if (warp == NPW + 1) {
    // ------------------------------------------------------------ weight TMA (one thread of its own warp)
    // W[k][:, kb*32 .. +32) per stage: one tiled TMA load, counted in bytes on the stage's "full" barrier.  Kept off
    // the gathering warps: the arrive.expect_tx + TMA issue sat on the critical path of the stage's slowest producer.
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int it = 0; it < T; ++it) {
        mbar_wait_sleep(empty + s, ph ^ 1u, 500);
        const int k = klist[it / nkb], kb = it - (it / nkb) * nkb;
        mbar_arrive_expect_tx(full + s, b_bytes);
        tma_load_2d(b_base + (uint32_t)s * b_bytes, &tmW, kb * 32, k * w_rows_per_k + w_row0, full + s);
        if (++s == nstages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (elect_one()) {
    // ------------------------------------------------------------ MMA issuer (one elected thread, see elect_one)
    // Descriptors differ only in the 14-bit start-address field: build the constant part once and patch the low word
    // with 32-bit adds (this single thread's dependent instruction chain is on the consumer's critical path).
    const uint64_t desc_hi = make_smem_desc(0, 16, 1024) & 0xFFFFFFFF00000000ull;
    const uint32_t desc_lo0 = (uint32_t)(make_smem_desc(0, 16, 1024) & 0xFFFFFFFFull);
    int s = 0, kb = 0;
    uint32_t ph = 0;
    for (int it = 0; it < T; ++it) {
      mbar_wait(full + s, ph);
      tc_fence_after();
      const int nj = min(32, Cin - kb * 32) >> 3;
      const uint32_t a_lo = desc_lo0 + ((a_base + (uint32_t)s * a_bytes) >> 4);
      const uint32_t b_lo = desc_lo0 + ((b_base + (uint32_t)s * b_bytes) >> 4);
#pragma unroll
      for (int m = 0; m < MSUB; ++m) {
#pragma unroll 4
        for (int j = 0; j < nj; ++j) {
          const uint64_t ad = desc_hi | (uint64_t)(a_lo + (uint32_t)m * (128 * 128 / 16) + 2 * j);
          const uint64_t bd = desc_hi | (uint64_t)(b_lo + 2 * j);
          mma_tf32(tmem + (uint32_t)(m * Cout), ad, bd, idesc, (it | j) ? 1u : 0u);
        }
      }
      mma_commit(empty + s);
      if (++kb == nkb) kb = 0;
      if (++s == nstages) { s = 0; ph ^= 1u; }
    }
    mma_commit(accum);
  }

  if (warp < NPW) {
    // ------------------------------------------------------------ epilogue: TMEM -> registers -> global
    // warp w may only read TMEM lanes 32*(w%4)..+31; with 8 producer warps, warps 4-7 take the second sub-tile
    const bool has_epilogue = (warp >> 2) < MSUB;   // the other producer warps go straight to the final barrier
    if (T > 0 && has_epilogue) {
      mbar_wait_sleep(accum, 0, 2000);
      tc_fence_after();
    }
    const bool vec = (ldo & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    const int q = warp & 3;
    for (int m = warp >> 2; m < MSUB; m += NPW / 4) {
      int row = row0 + m * 128 + q * 32 + lane;
      if (grouped) {
        const int pos = m * 128 + q * 32 + lane;
        row = pos < g_cnt ? __ldg(grp_out + g_start + pos) : n_rows;   // n_rows = "no row"
      }
      for (int c0 = 0; c0 < Cout; c0 += 16) {
        float v[16];
        if (T > 0) {
          tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(m * Cout + c0), v);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = 0.f;
        }
        if (row < n_rows) {
          if (addend && split == 0) {
            const float *ad = addend + (int64_t)row * ldadd + c0;
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += __ldg(ad + i);
          }
          float *o = out + (int64_t)row * ldo + c0;
          if (nsplit > 1) {
            if (T > 0 || (addend && split == 0)) {
#pragma unroll
              for (int i = 0; i < 16; ++i) atomicAdd(o + i, v[i]);
            }
          } else if (vec) {
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
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NPW) tmem_dealloc<NT>(tmem);
}

bool gather_conv_tc_supported(const float *A, int64_t lda, int K, int Cin, int Cout, const float *W) {
  return (Cin % 8 == 0) && (Cout % 16 == 0) && Cout >= 16 && Cout <= 1024 && (lda % 4 == 0) && K <= 64 &&
         ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
}

template <uint32_t NT, int MSUB, int NPW>
static int launch_gather_tc(dim3 grid, const TcSmemLayout &L, int nstages, const float *A, int64_t lda,
                            const int32_t *map, int64_t n_out, int K, const float *Wkm, int Cin, int Cout,
                            const float *addend, int64_t ldadd, float *out, int64_t ldo, int w_rows_per_k, int w_row0,
                            cudaStream_t st, const int32_t *grp_in = nullptr, const int32_t *grp_out = nullptr,
                            const int4 *grp_tab = nullptr) {
  auto kern = gather_conv_tc_kernel<NT, MSUB, NPW>;
  static bool smem_set = false;   // per template instantiation: opt in to the full 227 KB once, not on every launch
  if (!smem_set) {
    SCN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    smem_set = true;
  }
  const uint32_t idesc = make_idesc_tf32(128, Cout, 0, 0);
  alignas(64) CUtensorMap tmW;   // weight slice box: Cout rows x 32 channels of the (K*Cout_total, Cin) K-major stack
  if (make_tmap(&tmW, Wkm, (int64_t)K * w_rows_per_k, Cin, Cin, Cout)) return 1;
  kern<<<grid, 32 * (NPW + 2), L.total, st>>>(tmW, A, lda, map, (int)n_out, K, Cin, Cout, addend, ldadd, out, ldo,
                                              nstages, idesc, L.map_off, L.klist_off, L.bar_off, w_rows_per_k, w_row0,
                                              grp_in, grp_out, grp_tab);
  return 0;
}

// tile table of the grouped mode: one entry {offset, first rule, rules} per run of <= 128 rules of one offset, in offset
// order; entries beyond the live tiles get count 0.  offsets[K+1] is the rulebook's per-offset prefix (device memory).
__global__ void group_tiles_kernel(const int32_t *__restrict__ offsets, int K, int max_tiles, int4 *__restrict__ tab) {
  __shared__ int tbase[65];
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int k = 0; k < K; ++k) {
      tbase[k] = acc;
      acc += (offsets[k + 1] - offsets[k] + 127) >> 7;
    }
    tbase[K] = acc;
  }
  __syncthreads();
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < max_tiles; t += gridDim.x * blockDim.x) {
    int4 e = make_int4(0, 0, 0, 0);
    if (t < tbase[K]) {
      int k = 0;
      while (t >= tbase[k + 1]) ++k;
      const int beg = offsets[k], cnt = offsets[k + 1] - beg, first = (t - tbase[k]) << 7;
      e = make_int4(k, beg + first, min(128, cnt - first), 0);
    }
    tab[t] = e;
  }
}

int group_tiles(const int32_t *offsets_dev, int K, int64_t max_tiles, int32_t *tab, cudaStream_t st) {
  if (K < 1 || K > 64) return set_error("group_tiles: K=%d outside [1,64]", K);
  if (max_tiles <= 0) return 0;
  group_tiles_kernel<<<(unsigned)ceil_div(max_tiles, 256), 256, 0, st>>>(offsets_dev, K, (int)max_tiles,
                                                                         reinterpret_cast<int4 *>(tab));
  SCN_CHECK_LAUNCH("group_tiles");
  count_launch(1);
  return 0;
}

// out[out_rows[p]] = A[in_rows[p]] . W[offset of p] over an offset-sorted rule list (every output row named once)
int grouped_conv_tc(const float *A, int64_t lda, const int32_t *in_rows, const int32_t *out_rows, const int32_t *tab,
                    int64_t max_tiles, int64_t n_out, int K, const float *Wkm, int Cin, int Cout, float *out, int64_t ldo,
                    cudaStream_t st) {
  if (max_tiles <= 0) return 0;
  for (int n0 = 0; n0 < Cout; n0 += 256) {
    const int nc = Cout - n0 < 256 ? Cout - n0 : 256;
    int nstages = kMaxStages;
    TcSmemLayout L = tc_layout(1, nc, 1, nstages);
    while (nstages > 2 && L.total > 227 * 1024) L = tc_layout(1, nc, 1, --nstages);
    if (L.total > 227 * 1024) return set_error("grouped_conv_tc: shared memory %u too large", L.total);
    dim3 grid((unsigned)max_tiles, 1);
    int rc;
#define SCN_ARGS grid, L, nstages, A, lda, nullptr, n_out, K, Wkm, Cin, nc, nullptr, 0, out + n0, ldo, Cout, n0, st, in_rows, out_rows, reinterpret_cast<const int4 *>(tab)
    if (nc <= 32) rc = launch_gather_tc<32, 1, 16>(SCN_ARGS);
    else if (nc <= 64) rc = launch_gather_tc<64, 1, 16>(SCN_ARGS);
    else if (nc <= 128) rc = launch_gather_tc<128, 1, 16>(SCN_ARGS);
    else rc = launch_gather_tc<256, 1, 16>(SCN_ARGS);
#undef SCN_ARGS
    if (rc) return rc;
    SCN_CHECK_LAUNCH("grouped_conv_tc");
    count_launch(1);
  }
  return 0;
}

static int gather_conv_tc_part(const float *A, int64_t lda, const int32_t *map, int64_t n_out, int K, const float *Wkm,
                               int Cin, int Cout, const float *addend, int64_t ldadd, float *out, int64_t ldo,
                               int w_rows_per_k, int w_row0, cudaStream_t st);

// output channels beyond 256 (the widest tcgen05 N) are produced by separate launches over column slices
int gather_conv_tc(const float *A, int64_t lda, const int32_t *map, int64_t n_out, int K, const float *Wkm, int Cin,
                   int Cout, const float *addend, int64_t ldadd, float *out, int64_t ldo, cudaStream_t st) {
  for (int n0 = 0; n0 < Cout; n0 += 256) {
    const int nc = Cout - n0 < 256 ? Cout - n0 : 256;
    if (gather_conv_tc_part(A, lda, map, n_out, K, Wkm, Cin, nc, addend ? addend + n0 : nullptr, ldadd, out + n0, ldo,
                            Cout, n0, st))
      return 1;
  }
  return 0;
}

static int gather_conv_tc_part(const float *A, int64_t lda, const int32_t *map, int64_t n_out, int K, const float *Wkm,
                               int Cin, int Cout, const float *addend, int64_t ldadd, float *out, int64_t ldo,
                               int w_rows_per_k, int w_row0, cudaStream_t st) {
  if (n_out <= 0) return 0;
  // 256-row CTAs (two accumulators share each weight stage) once there are enough rows to fill the chip twice over
  // (eight producer warps then gather for one CTA per SM; the weight slice of every stage is fetched once per 256 rows;
  //  measured: a win from Cout = 96 up, while Cout <= 64 is faster as two 128-row CTAs per SM)
  int msub = (n_out >= (int64_t)256 * kNumSMs * 2 && Cout > 64 && 2 * Cout <= 512) ? 2 : 1;
  if (g_opt.tc_msub) msub = (g_opt.tc_msub == 2 && 2 * Cout <= 512) ? 2 : 1;  // test hook
  int nstages = kMaxStages;
  TcSmemLayout L = tc_layout(msub, Cout, K, nstages);
  while (nstages > 2 && L.total > 227 * 1024) L = tc_layout(msub, Cout, K, --nstages);
  if (L.total > 227 * 1024) return set_error("gather_conv_tc: shared memory %u too large", L.total);
  const int64_t tiles = ceil_div(n_out, msub * 128);
  // tiny levels: split the offsets over several CTAs per tile (atomic reduction into a zeroed output)
  int nsplit = 1;
  if (K > 1 && tiles * 2 <= kNumSMs) {
    nsplit = (int)min((int64_t)K, (int64_t)kNumSMs / tiles);
    if (nsplit > 9) nsplit = 9;
  }
  if (g_opt.tc_nsplit > 0) nsplit = max(1, min(K, g_opt.tc_nsplit));  // test hook
  if (nsplit > 1) {
    if (ldo == Cout) SCN_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)n_out * Cout, st));
    else SCN_CUDA(cudaMemset2DAsync(out, sizeof(float) * ldo, 0, sizeof(float) * Cout, (size_t)n_out, st));
  }
  dim3 grid((unsigned)tiles, (unsigned)nsplit);
  const int cols = msub * Cout;
  int rc;
#define SCN_ARGS grid, L, nstages, A, lda, map, n_out, K, Wkm, Cin, Cout, addend, ldadd, out, ldo, w_rows_per_k, w_row0, st
  if (msub == 2) {
    if (cols <= 64) rc = launch_gather_tc<64, 2, 16>(SCN_ARGS);
    else if (cols <= 128) rc = launch_gather_tc<128, 2, 16>(SCN_ARGS);
    else if (cols <= 256) rc = launch_gather_tc<256, 2, 16>(SCN_ARGS);
    else rc = launch_gather_tc<512, 2, 16>(SCN_ARGS);
  } else {
    if (cols <= 32) rc = launch_gather_tc<32, 1, 16>(SCN_ARGS);
    else if (cols <= 64) rc = launch_gather_tc<64, 1, 16>(SCN_ARGS);
    else if (cols <= 128) rc = launch_gather_tc<128, 1, 16>(SCN_ARGS);
    else rc = launch_gather_tc<256, 1, 16>(SCN_ARGS);
  }
#undef SCN_ARGS
  if (rc) return rc;
  SCN_CHECK_LAUNCH("gather_conv_tc");
  count_launch(1);
  return 0;
}

}  // namespace b200scn

// =====================================================================================================================
// Weight gradient on the tensor cores:  dW[k] = sum over the pair list of offset k of  A[pa[p],:]^T (x) G[pg[p],:]
// (replaces upstream's dConvolution_KMxKN_backward_dW atomicAdd kernels).
//
// The reduction dimension is the pair index, so both operands are "MN-major": a gathered feature row (128 bytes = 32
// channels) IS one K-row of the shared-memory image (SWIZZLE_128B_BASE32B, the only tf32 MN-major layout).  One CTA reduces a
// chunk of one offset's pair list into a Ca x Cg accumulator in TMEM (M = 128-channel tiles, N = Cg, K = 8 pairs per
// tcgen05.mma), then adds it to dW[k] with fp32 reductions.  Same producer / MMA-issuer / epilogue roles as above.
// =====================================================================================================================
namespace b200scn {

// PAIRS (template): pairs per stage, 32 or 64 (4 or 8 MMAs of K = 8 per stage and 128-channel tile)
constexpr int kDwWarps = 16;  // producer warps, four per stage

// A stage holds only the VALID 32-channel blocks of A (ab of them) followed by the gb blocks of G.  The M = 128 MMA of
// tile t still reads four MN blocks starting at block 4t; blocks past Ca alias whatever follows in shared memory
// (G blocks, the next stage, the tail pad) -- finite or not, they only feed accumulator rows >= Ca, which are never read.
template <uint32_t NT, int PAIRS>
__global__ void __launch_bounds__(32 * (kDwWarps + 1))
pair_dw_tc_kernel(const float *__restrict__ A, int64_t lda, const float *__restrict__ G, int64_t ldg,
                  const int32_t *__restrict__ pair_a, const int32_t *__restrict__ pair_g,
                  const int32_t *__restrict__ offsets, int n_single, int chunk, int Ca, int Cg, int nstages,
                  uint32_t idesc, float *__restrict__ dW, int64_t dw_kstride, const int32_t *__restrict__ blk_tab, int nblk,
                  uint32_t pad_bytes, int dbg) {
  constexpr int NPW = kDwWarps;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *sm = smem_raw + (base - raw);
  const int mt = (Ca + 127) >> 7;            // 128-channel M tiles
  const int ab = (Ca + 31) >> 5;             // valid 32-channel blocks of A
  const int gb = (Cg + 31) >> 5;
  const uint32_t blk = PAIRS * 128;       // one 32-channel block of one stage: 32 pair rows x 128 B
  const uint32_t a_bytes = (uint32_t)ab * blk, g_bytes = (uint32_t)gb * blk;
  const uint32_t stage_bytes = a_bytes + g_bytes;
  // barriers sit after the stages and a 3-block pad (the aliasing reads above may run 3 blocks past the last stage)
  uint64_t *full = reinterpret_cast<uint64_t *>(sm + (uint32_t)nstages * stage_bytes + pad_bytes);
  uint64_t *empty = full + kMaxStages;
  uint64_t *accum = empty + kMaxStages;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(accum + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // offsets vary fastest over the grid: CTAs resident at the same time work on the same stretch of every offset's list,
  // i.e. (with Morton-ordered lists) on the same region of space, so gathered rows are shared in L2
  const int k = blockIdx.x;
  int p0, p1;
  if (blk_tab) {
    // one CTA per (offset, row block) of a b200scn_pair_lists_blocked table: block b of every offset is the same stretch of
    // the Morton curve, so the CTAs resident at the same time (offsets vary fastest over the grid) gather the same rows
    p0 = __ldg(blk_tab + (int64_t)k * nblk + blockIdx.y);
    p1 = __ldg(blk_tab + (int64_t)k * nblk + blockIdx.y + 1);
  } else {
    const int beg = offsets ? offsets[k] : 0;
    const int end = offsets ? offsets[k + 1] : n_single;
    p0 = beg + blockIdx.y * chunk;
    p1 = min(p0 + chunk, end);
  }
  if (p0 >= p1) return;  // uniform for the whole CTA
  const int T = (p1 - p0 + PAIRS - 1) / PAIRS;
  const int wps = NPW / nstages;  // producer warps per stage (nstages is 2 or 4)

  if (tid == 0) {
    for (int s = 0; s < nstages; ++s) {
      mbar_init(full + s, wps);   // one arrival per producer warp (128 lanes arriving one by one on the same word cost ~500 cycles per stage)
      mbar_init(empty + s, 1);
    }
    mbar_init(accum, 1);
    fence_barrier_init();
  }
  if (warp == NPW) tmem_alloc<NT>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < NPW) {
    // producers: stage s is always filled by warps s, s + nstages, ... (each takes an equal share of the 32 pairs).
    // The loop is bound by instruction issue and copy latency (ncu: ~220 instructions per warp and stage before this
    // rewrite, half of all issue slots), so everything that does not depend on the pair is computed once: the swizzled
    // shared-memory offset of each of the lane's rows, the channel-block masks, the stage's base address; the ring phase
    // is toggled instead of divided out.
    const int c = lane & 7, rl = lane >> 3;
    const int lg = nstages == 4 ? 2 : 1;          // nstages is 2 or 4
    const int my_stage = warp & (nstages - 1), part = warp >> lg;
    const int rows_per_warp = PAIRS / wps;
    constexpr int kMaxSlots = PAIRS / 16;   // slots per lane with four producer warps per stage (nstages = 4); half of them with eight
    const int nslots = rows_per_warp / 4;
    uint32_t soff[kMaxSlots];
#pragma unroll
    for (int i = 0; i < kMaxSlots; ++i) soff[i] = sw128_32b((uint32_t)(part * rows_per_warp + rl + 4 * i), (uint32_t)c);
    // channel blocks this lane's 16-byte chunk exists in (the last block of a width that is not a multiple of 32 is partial)
    const int na = min(ab, (Ca - c * 4 + 31) >> 5), ng = min(gb, (Cg - c * 4 + 31) >> 5);
    const uint32_t a_st = base + (uint32_t)my_stage * stage_bytes, g_st = a_st + a_bytes;
    const float *Ac = A + c * 4, *Gc = G + c * 4;
    // the pair indices of the NEXT iteration are fetched while the current one is being staged (the index load ->
    // address -> copy chain is otherwise a serial global round trip at the head of every stage)
    int nra[kMaxSlots], nrg[kMaxSlots];
    const int pl = p0 + part * rows_per_warp + rl;
    auto fetch = [&](int it_) {
      const int pb = pl + it_ * PAIRS;
#pragma unroll
      for (int i = 0; i < kMaxSlots; ++i) {
        if (i < nslots) {
          const int p = pb + 4 * i;
          const bool live = p < p1;
          nra[i] = live ? (pair_a ? __ldg(pair_a + p) : p) : -1;
          nrg[i] = live ? (pair_g ? __ldg(pair_g + p) : p) : -1;
        }
      }
    };
    fetch(my_stage);
    uint32_t ph = 1;   // parity of `empty` to wait for (first pass: the slot is free)
    for (int it = my_stage; it < T; it += nstages) {
      int cra[kMaxSlots], crg[kMaxSlots];
#pragma unroll
      for (int i = 0; i < kMaxSlots; ++i) { cra[i] = nra[i]; crg[i] = nrg[i]; }
      fetch(it + nstages);
      mbar_wait(empty + my_stage, ph);
      ph ^= 1u;
#pragma unroll
      for (int i = 0; i < kMaxSlots; ++i) {
        if (i >= nslots) break;
        const bool live = cra[i] >= 0;
        const uint32_t sz = live ? 16u : 0u;      // a dead pair (list tail) zero-fills its rows
        const float *arow = Ac + (int64_t)(live ? cra[i] : 0) * lda;
        const float *grow = Gc + (int64_t)(live ? crg[i] : 0) * ldg;
        const uint32_t so = soff[i];
        if (dbg & 2) continue;   // knockout experiment (b200scn_set_option "dw_dbg"): no gathers
        for (int b = 0; b < na; ++b) cp_async16(a_st + (uint32_t)b * blk + so, arow + b * 32, sz);
        for (int b = na; b < ab; ++b) cp_async16(a_st + (uint32_t)b * blk + so, A, 0u);
        for (int b = 0; b < ng; ++b) cp_async16(g_st + (uint32_t)b * blk + so, grow + b * 32, sz);
        for (int b = ng; b < gb; ++b) cp_async16(g_st + (uint32_t)b * blk + so, G, 0u);
      }
      cp_async_wait_all();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(full + my_stage);
    }
  } else if (elect_one()) {
    // constant descriptor part once, 32-bit patching of the start address per MMA (see the gather kernel)
    const uint64_t desc_hi = make_smem_desc(0, blk, 512, 1) & 0xFFFFFFFF00000000ull;
    const uint32_t desc_lo0 = (uint32_t)(make_smem_desc(0, blk, 512, 1) & 0xFFFFFFFFull);
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < T; ++it) {
      mbar_wait(full + s, ph);
      tc_fence_after();
      const uint32_t a_lo = desc_lo0 + ((base + (uint32_t)s * stage_bytes) >> 4);
      const uint32_t g_lo = a_lo + (a_bytes >> 4);
      for (int t = 0; t < mt && !(dbg & 1); ++t) {   // (dbg 1: knockout experiment, no MMA issue)
#pragma unroll
        for (int j = 0; j < PAIRS / 8; ++j) {
          const uint64_t ad = desc_hi | (uint64_t)(a_lo + (uint32_t)(4 * t) * (blk >> 4) + j * 64);
          const uint64_t gd = desc_hi | (uint64_t)(g_lo + j * 64);
          mma_tf32(tmem + (uint32_t)(t * Cg), ad, gd, idesc, (it | j) ? 1u : 0u);
        }
      }
      mma_commit(empty + s);
      if (++s == nstages) { s = 0; ph ^= 1u; }
    }
    mma_commit(accum);
  }

  if (warp < NPW) {
    mbar_wait(accum, 0);
    tc_fence_after();
    float *Wk = dW + (int64_t)k * dw_kstride;
    const int q = warp & 3;
    for (int t = warp >> 2; t < mt; t += NPW / 4) {   // warps 0-3: even tiles' lane quarters, warps 4-7: odd tiles'
      const int ca = t * 128 + q * 32 + lane;
      for (int c0 = 0; c0 < Cg; c0 += 16) {
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * Cg + c0), v);
        if (ca < Ca) {
          float *o = Wk + (int64_t)ca * Cg + c0;
#pragma unroll
          for (int i = 0; i < 16; ++i) atomicAdd(o + i, v[i]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NPW) tmem_dealloc<NT>(tmem);
}

bool pair_dw_tc_supported(const float *A, int64_t lda, const float *G, int64_t ldg, int Ca, int Cg) {
  return (Ca % 4 == 0) && (Cg % 16 == 0) && Cg >= 16 && Cg <= 256 && Ca >= 4 && (lda % 4 == 0) && (ldg % 4 == 0) &&
         ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && ((reinterpret_cast<uintptr_t>(G) & 15) == 0);
}

// one launch for channels [ca0, ca0 + Ca) of A (the accumulator of a launch must fit the 512 TMEM columns)
static int pair_dw_tc_part(const float *A, int64_t lda, const float *G, int64_t ldg, const int32_t *pair_a,
                           const int32_t *pair_g, const int32_t *offsets_dev, int K, int64_t n_pairs_max, int Ca, int Cg,
                           float *dW, int64_t dw_kstride, cudaStream_t st, const int32_t *blk_tab = nullptr, int nblk = 0) {
  const int mt = (Ca + 127) >> 7, ab = (Ca + 31) >> 5, gb = (Cg + 31) >> 5;
  const int pairs = g_opt.dw_pairs == 64 ? 64 : 32;   // pairs per stage (64 measured 5-20 % slower wherever fewer stages fit)
  const uint32_t blk = (uint32_t)pairs * 128;
  const uint32_t stage_bytes = (uint32_t)(ab + gb) * blk;
  // an M = 128 MMA reads four 32-channel blocks of A whatever Ca is: with fewer A blocks the reads run on into the G blocks
  // of the stage and, in the last stage, past it -- the rows they produce belong to channels >= Ca and are never stored
  const int over = 4 * mt - ab - gb;
  const uint32_t pad = over > 0 ? (uint32_t)over * blk : 0u;
  int nstages = kMaxStages;
  // registers allow two CTAs per SM: four stages whenever two CTAs of them fit in shared memory (113 KB each)
  if ((uint32_t)nstages * stage_bytes + pad + 256 + 1024 > 113 * 1024) nstages = 2;
  const uint32_t smem = (uint32_t)nstages * stage_bytes + pad + 256 + 1024;
  if (smem > 227 * 1024) return set_error("pair_dw_tc: shared memory %u too large", smem);
  // chunk of pairs per CTA: enough CTAs for ~4 per SM overall, a multiple of the stage size
  int64_t want_chunks = ceil_div((int64_t)kNumSMs * 4, (int64_t)K);
  int64_t chunk = ceil_div(n_pairs_max, want_chunks > 0 ? want_chunks : 1);
  if (chunk < 512) chunk = 512;
  const int64_t chunk_max = g_opt.dw_chunk;   // small chunks keep the region the resident CTAs work on (27 offsets x ~11 chunks) inside L2
  if (chunk > chunk_max) chunk = chunk_max;
  if (ceil_div(n_pairs_max, chunk) > 65535) chunk = ceil_div(n_pairs_max, 65535);
  chunk = ceil_div(chunk, (int64_t)pairs) * pairs;
  dim3 grid((unsigned)K, (unsigned)ceil_div(n_pairs_max, chunk));
  if (blk_tab) grid.y = (unsigned)nblk;
  const uint32_t idesc = make_idesc_tf32(128, Cg, 1, 1);
  const int cols = mt * Cg;
#define SCN_LAUNCH_DW(NT)                                                                                        \
  do {                                                                                                           \
    if (pairs == 32) SCN_LAUNCH_DW2(NT, 32);                                                                     \
    else SCN_LAUNCH_DW2(NT, 64);                                                                                 \
  } while (0)
#define SCN_LAUNCH_DW2(NT, PR)                                                                                   \
  do {                                                                                                           \
    auto kern = pair_dw_tc_kernel<NT, PR>;                                                                       \
    static bool smem_set = false;                                                                                \
    if (!smem_set) {                                                                                             \
      SCN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));             \
      smem_set = true;                                                                                           \
    }                                                                                                            \
    kern<<<grid, 32 * (kDwWarps + 1), smem, st>>>(A, lda, G, ldg, pair_a, pair_g, offsets_dev, (int)n_pairs_max, \
                                                  (int)chunk, Ca, Cg, nstages, idesc, dW, dw_kstride, blk_tab, nblk, pad, g_opt.dw_dbg); \
  } while (0)
  if (cols <= 32) SCN_LAUNCH_DW(32);
  else if (cols <= 64) SCN_LAUNCH_DW(64);
  else if (cols <= 128) SCN_LAUNCH_DW(128);
  else if (cols <= 256) SCN_LAUNCH_DW(256);
  else SCN_LAUNCH_DW(512);
#undef SCN_LAUNCH_DW
#undef SCN_LAUNCH_DW2
  SCN_CHECK_LAUNCH("pair_dw_tc");
  count_launch(1);
  return 0;
}

int pair_dw_tc(const float *A, int64_t lda, const float *G, int64_t ldg, const int32_t *pair_a,
               const int32_t *pair_g, const int32_t *offsets_dev, int K, int64_t n_pairs_max, int Ca, int Cg,
               float *dW, cudaStream_t st, const int32_t *blk_tab, int nblk) {
  SCN_CUDA(cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)K * Ca * Cg, st));
  if (n_pairs_max <= 0) return 0;
  if (blk_tab && nblk > 65535) return set_error("pair_dw_blocked: %d row blocks exceed the grid limit", nblk);
  // channel slices of A whose accumulators fit TMEM: 128-channel tiles x Cg columns <= 512
  const int max_tiles = 512 / Cg >= 1 ? 512 / Cg : 1;
  const int slice = 128 * (max_tiles > 4 ? 4 : max_tiles);
  for (int ca0 = 0; ca0 < Ca; ca0 += slice) {
    const int ca = Ca - ca0 < slice ? Ca - ca0 : slice;
    if (pair_dw_tc_part(A + ca0, lda, G, ldg, pair_a, pair_g, offsets_dev, K, n_pairs_max, ca, Cg,
                        dW + (int64_t)ca0 * Cg, (int64_t)Ca * Cg, st, blk_tab, nblk))
      return 1;
  }
  return 0;
}

}  // namespace b200scn

// =====================================================================================================================
// TMA variant of the gather convolution (opt-in, B200SCN_TC_TMA=1; measured slower, see abi_conv.cu): the per-lane cp.async producers are replaced by ONE warp that
// issues tile::gather4 TMA copies -- each lane names four neighbour rows (absent neighbour = an out-of-range row index,
// which the TMA unit zero-fills), so one warp instruction stages 128 gathered rows x 128 bytes, already in the
// SWIZZLE_128B image tcgen05.mma reads -- plus one tiled TMA load for the weight slice.  Completion is counted in bytes
// on the stage's mbarrier (no LSU traffic, no proxy fence), so the producer runs `nstages` stages ahead of the tensor
// pipe.  Warps: 0-3 epilogue, 4 TMA producer, 5 MMA issuer.
// =====================================================================================================================
namespace b200scn {

constexpr int kTmaThreads = 192;

template <uint32_t NT, int MSUB>
__global__ void __launch_bounds__(kTmaThreads)
gather_conv_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                       const int32_t *__restrict__ map, int n_rows, int n_in, int K, int Cin, int Cout,
                       const float *__restrict__ addend, int64_t ldadd, float *__restrict__ out, int64_t ldo,
                       int nstages, uint32_t idesc, uint32_t map_off, uint32_t klist_off, uint32_t bar_off) {
  constexpr int ROWS = MSUB * 128;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *sm = smem_raw + (base - raw);
  const uint32_t a_bytes = ROWS * 128, b_bytes = (uint32_t)Cout * 128;
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
  const int row0 = blockIdx.x * ROWS;
  const int nsplit = gridDim.y, split = blockIdx.y;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
  }
  for (int k = tid; k < K; k += kTmaThreads) kflag[k] = 0;
  __syncthreads();
  for (int e = tid; e < ROWS * K; e += kTmaThreads) {
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
    const int per = (n + nsplit - 1) / nsplit;
    const int lo = min(n, split * per), hi = min(n, lo + per);
    for (int i = lo; i < hi; ++i) klist[i - lo] = klist[i];
    *nk_p = hi - lo;
    for (int s = 0; s < nstages; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    mbar_init(accum, 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc<NT>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int nk = *nk_p;
  const int nkb = (Cin + 31) >> 5;
  const int T = nk * nkb;

  if (warp == 4) {
    // ------------------------------------------------------------ TMA producer (one warp, never waits for data)
    for (int it = 0; it < T; ++it) {
      const int s = it % nstages;
      const uint32_t ph = (uint32_t)(it / nstages) & 1u;
      mbar_wait(empty + s, ph ^ 1u);
      const int k = klist[it / nkb], kb = it - (it / nkb) * nkb;
      const uint32_t a_st = a_base + (uint32_t)s * a_bytes, b_st = b_base + (uint32_t)s * b_bytes;
      if (lane == 0) {
        mbar_arrive_expect_tx(full + s, a_bytes + b_bytes);
        tma_load_2d(b_st, &tmW, kb * 32, k * Cout, full + s);
      }
      __syncwarp();
#pragma unroll
      for (int m = 0; m < MSUB; ++m) {
        const int r = m * 128 + lane * 4;
        int i0 = smap[(r + 0) * K + k], i1 = smap[(r + 1) * K + k], i2 = smap[(r + 2) * K + k], i3 = smap[(r + 3) * K + k];
        i0 = i0 < 0 ? n_in : i0; i1 = i1 < 0 ? n_in : i1; i2 = i2 < 0 ? n_in : i2; i3 = i3 < 0 ? n_in : i3;
        tma_gather4(a_st + (uint32_t)r * 128, &tmA, kb * 32, i0, i1, i2, i3, full + s);
      }
    }
  } else if (warp == 5 && lane == 0) {
    // ------------------------------------------------------------ MMA issuer (one thread)
    for (int it = 0; it < T; ++it) {
      const int s = it % nstages;
      const uint32_t ph = (uint32_t)(it / nstages) & 1u;
      mbar_wait(full + s, ph);
      tc_fence_after();
      const int kb = it % nkb;
      const int kvalid = min(32, Cin - kb * 32);
      const uint32_t a_st = a_base + (uint32_t)s * a_bytes, b_st = b_base + (uint32_t)s * b_bytes;
#pragma unroll
      for (int m = 0; m < MSUB; ++m)
        for (int j = 0; j < (kvalid >> 3); ++j) {
          const uint64_t ad = make_smem_desc(a_st + (uint32_t)m * (128 * 128) + j * 32, 16, 1024);
          const uint64_t bd = make_smem_desc(b_st + j * 32, 16, 1024);
          mma_tf32(tmem + (uint32_t)(m * Cout), ad, bd, idesc, (it > 0 || j > 0) ? 1u : 0u);
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
    const bool vec = (ldo & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
#pragma unroll
    for (int m = 0; m < MSUB; ++m) {
      const int row = row0 + m * 128 + warp * 32 + lane;
      for (int c0 = 0; c0 < Cout; c0 += 16) {
        float v[16];
        if (T > 0) {
          tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(m * Cout + c0), v);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = 0.f;
        }
        if (row < n_rows) {
          if (addend && split == 0) {
            const float *ad = addend + (int64_t)row * ldadd + c0;
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += __ldg(ad + i);
          }
          float *o = out + (int64_t)row * ldo + c0;
          if (nsplit > 1) {
            if (T > 0 || (addend && split == 0)) {
#pragma unroll
              for (int i = 0; i < 16; ++i) atomicAdd(o + i, v[i]);
            }
          } else if (vec) {
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
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<NT>(tmem);
}

template <uint32_t NT, int MSUB>
static int launch_gather_tma(dim3 grid, const TcSmemLayout &L, int nstages, const CUtensorMap &tmA,
                             const CUtensorMap &tmW, const int32_t *map, int64_t n_out, int64_t n_in, int K, int Cin,
                             int Cout, const float *addend, int64_t ldadd, float *out, int64_t ldo, cudaStream_t st) {
  auto kern = gather_conv_tma_kernel<NT, MSUB>;
  static bool smem_set = false;   // per template instantiation: opt in to the full 227 KB once, not on every launch
  if (!smem_set) {
    SCN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    smem_set = true;
  }
  const uint32_t idesc = make_idesc_tf32(128, Cout, 0, 0);
  kern<<<grid, kTmaThreads, L.total, st>>>(tmA, tmW, map, (int)n_out, (int)n_in, K, Cin, Cout, addend, ldadd, out, ldo,
                                           nstages, idesc, L.map_off, L.klist_off, L.bar_off);
  return 0;
}

int gather_conv_tma(const float *A, int64_t lda, int64_t n_in, const int32_t *map, int64_t n_out, int K,
                    const float *Wkm, int Cin, int Cout, const float *addend, int64_t ldadd, float *out, int64_t ldo,
                    cudaStream_t st) {
  if (n_out <= 0) return 0;
  int msub = (n_out >= (int64_t)256 * kNumSMs * 2 && 2 * Cout <= 512) ? 2 : 1;
  if (g_opt.tc_msub) msub = (g_opt.tc_msub == 2 && 2 * Cout <= 512) ? 2 : 1;  // test hook
  int nstages = kMaxStages;
  TcSmemLayout L = tc_layout(msub, Cout, K, nstages);
  while (nstages > 2 && L.total > 227 * 1024) L = tc_layout(msub, Cout, K, --nstages);
  if (L.total > 227 * 1024) return set_error("gather_conv_tma: shared memory %u too large", L.total);
  const int64_t tiles = ceil_div(n_out, msub * 128);
  int nsplit = 1;
  if (K > 1 && tiles * 2 <= kNumSMs) {
    nsplit = (int)min((int64_t)K, (int64_t)kNumSMs / tiles);
    if (nsplit > 9) nsplit = 9;
  }
  if (g_opt.tc_nsplit > 0) nsplit = max(1, min(K, g_opt.tc_nsplit));  // test hook
  if (nsplit > 1) {
    if (ldo == Cout) SCN_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)n_out * Cout, st));
    else SCN_CUDA(cudaMemset2DAsync(out, sizeof(float) * ldo, 0, sizeof(float) * Cout, (size_t)n_out, st));
  }
  alignas(64) CUtensorMap tmA, tmW;
  if (make_tmap(&tmA, A, n_in, Cin, lda, 1)) return 1;                       // gather4: one row per box
  if (make_tmap(&tmW, Wkm, (int64_t)K * Cout, Cin, Cin, Cout)) return 1;      // weight slice: Cout rows x 32 channels
  dim3 grid((unsigned)tiles, (unsigned)nsplit);
  const int cols = msub * Cout;
#define SCN_ARGS grid, L, nstages, tmA, tmW, map, n_out, n_in, K, Cin, Cout, addend, ldadd, out, ldo, st
  int rc;
  if (msub == 2) {
    if (cols <= 64) rc = launch_gather_tma<64, 2>(SCN_ARGS);
    else if (cols <= 128) rc = launch_gather_tma<128, 2>(SCN_ARGS);
    else if (cols <= 256) rc = launch_gather_tma<256, 2>(SCN_ARGS);
    else rc = launch_gather_tma<512, 2>(SCN_ARGS);
  } else {
    if (cols <= 32) rc = launch_gather_tma<32, 1>(SCN_ARGS);
    else if (cols <= 64) rc = launch_gather_tma<64, 1>(SCN_ARGS);
    else if (cols <= 128) rc = launch_gather_tma<128, 1>(SCN_ARGS);
    else rc = launch_gather_tma<256, 1>(SCN_ARGS);
  }
#undef SCN_ARGS
  if (rc) return rc;
  SCN_CHECK_LAUNCH("gather_conv_tma");
  count_launch(1);
  return 0;
}

}  // namespace b200scn
