// dw_tile.cu -- TILE-STATIONARY weight gradient of the submanifold 3x3x3 convolution (replaces
// upstream scn's dConvolution_KMxKN_backward_dW: one launch per offset, global atomicAdd per rule).
//
//   dW[k] = sum over rules (in, out) of offset k of  A[in,:]^T (x) G[out,:]
//
// The pair-list kernel (conv_tc.cu: pair_dw_tc_kernel) fetches TWO rows from L2 for every rule: ~8.7 TB/s of gathered rows
// averaged over a level-1 launch, i.e. it sits on the L2 throughput cap (profiles/r2d_dw_knockout.txt).  Here the work is
// cut along the same Morton-ordered 128-row tiles and per-tile halo lists the forward kernel uses (conv_halo.cu:
// b200scn_tile_plan): a persistent CTA stages a tile's gradient rows G (the tile's own rows) and its DISTINCT input rows
// (halo, one 32-channel block) in shared memory ONCE, and every rule of the tile is then served from shared memory:
//   * the contraction runs over the tile's 128 rows (K = 128 = 16 tcgen05.mma of K = 8), M = 128 = 4 kernel offsets x 32
//     input channels, N = Cg;
//   * the A^T operand [M lanes x K columns] lives in TENSOR MEMORY (as the forward kernel's A operand does): builder warp
//     (slot, q) owns the 32 TMEM lanes of offset 4 g + q -- lane = input channel -- and for every tile row reads ONE
//     32-bit word per lane from the row's halo slot: the 32 lanes of a warp read 128 consecutive bytes, so every
//     shared-memory wavefront is full and conflict-free whatever rows the rulebook names, and absent neighbours read an
//     all-zero row (no branch).  16 rows go to tensor memory with one tcgen05.st.x16;
//   * the G tile is the B operand of every MMA of the tile as it lies (MN-major, SWIZZLE_128B_BASE32B);
//   * accumulators for ALL offsets of the CTA stay in tensor memory across all its tiles (up to 7 groups of 4 offsets x
//     Cg columns beside two 128-column A^T slots); each CTA writes its partial ONCE, and a second kernel sums the partials
//     in a fixed order: the result is bit-reproducible (no fp32 atomics).
// A CTA owns one 32-channel block of A (and, when its accumulators would exceed tensor memory, a subset of the offset
// groups); tiles are dealt round-robin to the CTAs of a split.  Both operands are rounded to the nearest TF32 as they
// land in shared memory (by the thread that copied them), so the tensor core's truncation never sees an unrounded value.
#include "common.cuh"
#include "tc_common.cuh"

namespace b200scn {

using namespace tc;

constexpr int kDT = 128;                  // tile rows
constexpr int kDtMap = 27 * kDT;          // lmap entries per tile
constexpr int kDtLoaders = 8, kDtBuilders = 8;
constexpr int kDtThreads = 32 * (1 + kDtLoaders + kDtBuilders);
constexpr uint32_t kBlkG = kDT * 128;     // one 32-column block of the G tile: 128 rows x 128 bytes
constexpr uint32_t kSlotCols = 128;       // TMEM columns of one A^T slot (K = 128 tile rows)
constexpr uint16_t kDtAbsent = 0xFFFF, kDtOverflow = 0xFFFE;

struct DwLayout {
  uint32_t g_off[2], halo_off[2], tab_off[2], misc_off[2], bar_off, total;
  int nbuf;
};
// tab of a tile buffer: [27][128] uint16 = halo row of the rule's input in 16-byte units (row hcap = the all-zero row)
// misc of a tile buffer: sorow[128] int | hids[hcap] int | flags[4] int (flags[0]: some rule lies beyond the halo capacity)
static DwLayout dw_layout(int hcap, int Cg, int nbuf) {
  DwLayout L;
  const uint32_t gb = (uint32_t)(Cg + 31) / 32;
  uint32_t o = 0;
  for (int b = 0; b < 2; ++b) {
    const bool live = b < nbuf;
    if (live) o = (o + 1023) & ~1023u;            // the swizzled operand image needs a 1024-byte aligned base
    L.g_off[b] = live ? o : L.g_off[0];
    if (live) o += gb * kBlkG;                    // 1024-byte aligned blocks
    L.halo_off[b] = live ? o : L.halo_off[0];
    if (live) o += ((uint32_t)hcap + 1) * 128;    // + the all-zero row
    L.tab_off[b] = live ? o : L.tab_off[0];
    if (live) o += (kDtMap * 2 + 15) & ~15;
    L.misc_off[b] = live ? o : L.misc_off[0];
    if (live) o += ((512 + (uint32_t)hcap * 4 + 16) + 127) & ~127u;
  }
  L.bar_off = (o + 15) & ~15u;
  L.total = L.bar_off + 256 + 1024;   // barriers + alignment slack
  L.nbuf = nbuf;
  return L;
}

__device__ __forceinline__ float4 dt_lds_f4(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ uint4 dt_lds_u4(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ float dt_lds_f1(uint32_t saddr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ float dt_rna(float v) {   // nearest TF32, ties away, as two full-rate integer instructions
  return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ void dt_rna4(float4 &v) { v.x = dt_rna(v.x); v.y = dt_rna(v.y); v.z = dt_rna(v.z); v.w = dt_rna(v.w); }
__device__ __forceinline__ void dt_named_bar(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns, registers -> TMEM (lane t of the warp writes TMEM lane base + t)
__device__ __forceinline__ void dt_tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]),
      "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15])
      : "memory");
}
__device__ __forceinline__ void dt_tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem desc], kind::tf32, issued by ONE thread
__device__ __forceinline__ void dt_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Diagnostic: clock64 timeline of CTA 0 (b200scn_debug_dw_timeline): per tile it, slots 16*it + {0 loader waits for the
// buffer, 1 buffer free, 2 plan slice in, 3 copies landed, 4 tile published, 5 builder (slot 0, q 1) sees the tile, 6 that
// builder done, 7 MMA thread sees the tile, 8 MMA thread done issuing}.
__device__ long long *g_dw_timeline = nullptr;
#define DW_TL(slot) do { if (tl && it < 60) tl[16 * it + (slot)] = clock64(); } while (0)

// Warp 0: MMA issuer (and TMEM owner).  Warps 1..8: tile loaders (global -> shared, one tile ahead).
// Warps 9..16: A^T builders (shared -> registers -> tensor memory), then the epilogue; builder warp w owns TMEM lane
// quarter w % 4 (a hardware rule) = offset 4 g + (w % 4) of every group, and A^T slot (w - 9) / 4.
__global__ void __launch_bounds__(kDtThreads, 1)
dw_tile_kernel(const float *__restrict__ A, int64_t lda, const float *__restrict__ G, int64_t ldg,
               const int32_t *__restrict__ nbr, const int32_t *__restrict__ perm, const uint16_t *__restrict__ lmap,
               const int32_t *__restrict__ halo_ids, const int32_t *__restrict__ halo_n, int hcap, int n_rows, int Ca,
               int Cg, int ncb, int nsplit, int gper, float *__restrict__ part, int64_t slab, DwLayout L,
               uint32_t idesc) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *sm = smem_raw + (base - raw);
  const int nbuf = L.nbuf;
  uint64_t *tfull = reinterpret_cast<uint64_t *>(sm + L.bar_off);   // tile buffer loaded (one arrival per loader warp)
  uint64_t *tempty = tfull + 2;                                     // tile buffer free (builder warps + the MMA commit)
  uint64_t *afull = tempty + 2;                                     // A^T slot written (its four builder warps)
  uint64_t *aempty = afull + 2;                                     // A^T slot consumed (tcgen05.commit)
  uint64_t *done = aempty + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done + 1);
  uint32_t *acc_mask_s = tmem_slot + 1;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int sp = blockIdx.x % nsplit, ci = blockIdx.x / nsplit, nper = gridDim.x / nsplit;
  const int cb = sp % ncb, ks = sp / ncb;
  const int kbase = 4 * gper * ks;
  const int kend = min(27, kbase + 4 * gper);
  const int ng = kend > kbase ? (kend - kbase + 3) >> 2 : 0;
  const int ntiles = (n_rows + kDT - 1) / kDT;
  const int my_n = ci < ntiles ? (ntiles - ci + nper - 1) / nper : 0;
  const uint32_t acc_cols = (uint32_t)(gper * Cg);   // A^T slots follow the accumulators (the host sized TMEM for both)
  long long *tl = (g_dw_timeline && blockIdx.x == 0 && lane == 0) ? g_dw_timeline : nullptr;

  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull + b, kDtLoaders);
      mbar_init(tempty + b, kDtBuilders + 1);
      mbar_init(afull + b, 4);
      mbar_init(aempty + b, 1);
    }
    mbar_init(done, 1);
    *acc_mask_s = 0;
    fence_barrier_init();
  }
  // the all-zero halo row of every buffer (absent neighbours read it; never overwritten: live rows are < hcap)
  for (int e = tid; e < 2 * 8; e += kDtThreads)
    sts_f4(base + L.halo_off[e >> 3] + (uint32_t)hcap * 128 + (e & 7) * 16, make_float4(0.f, 0.f, 0.f, 0.f));
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp >= 1 && warp <= kDtLoaders) {
    // ------------------------------------------------------------------------------------------ tile loaders
    const int lt = tid - 32;                      // 0 .. 255
    constexpr int NL = 32 * kDtLoaders;
    const int cpr = Cg >> 2;                      // 16-byte chunks per G row
    const int cvalid = min(32, Ca - cb * 32);     // live channels of this CTA's block of A
    for (int it = 0; it < my_n; ++it) {
      const int tile = ci + it * nper, b = it % nbuf, row0 = tile * kDT;
      if (lt == 0) DW_TL(0);
      // long waits suspend (try_wait with a time hint) instead of spinning: loader warps polling in a tight loop take the
      // issue slots the MMA-issuing lane and the builders need
      mbar_wait_sleep(tempty + b, (((uint32_t)(it / nbuf)) & 1u) ^ 1u, 2000);
      if (lt == 0) DW_TL(1);
      const uint32_t g_b = base + L.g_off[b], halo_b = base + L.halo_off[b], tab_b = base + L.tab_off[b];
      int *sorow = reinterpret_cast<int *>(sm + L.misc_off[b]);
      int *hids = sorow + kDT;
      int *flags = hids + hcap;
      const int hn = __ldg(halo_n + tile);
      // phase 1: the tile's plan slice -> the builders' table (halo row of every rule in 16-byte units; absent neighbours
      // and rules beyond the halo capacity -> the all-zero row, the latter flagged for the builders' slow path), and the
      // row ids phase 2 gathers through
      {
        if (lt == 0) flags[0] = 0;
        const uint4 *src = reinterpret_cast<const uint4 *>(lmap + (int64_t)tile * kDtMap);   // 6912 bytes, 16-byte aligned
        bool ovf = false;
        for (int e = lt; e < kDtMap * 2 / 16; e += NL) {
          uint4 v = __ldg(src + e);
          uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t lo = w[i] & 0xFFFFu, hi = w[i] >> 16;
            ovf = ovf || lo == kDtOverflow || hi == kDtOverflow;
            w[i] = (min(lo, (uint32_t)hcap) * 8u) | ((min(hi, (uint32_t)hcap) * 8u) << 16);
          }
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(tab_b + e * 16), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
        }
        if (lt < kDT) sorow[lt] = row0 + lt < n_rows ? __ldg(perm + row0 + lt) : -1;
        const int32_t *ids = halo_ids + (int64_t)tile * hcap;
        for (int h = lt; h < hn; h += NL) hids[h] = __ldg(ids + h);
        dt_named_bar(2, NL);   // flags[0] = 0 is ordered before the atomicOr below; sorow / hids are complete
        if (ovf) atomicOr(flags, 1);
      }
      if (lt == 0) DW_TL(2);
      // phase 2: gradient rows of the tile (B operand image) and halo rows of A (this CTA's 32 channels): every 16-byte
      // copy of the tile is in flight at once (cp.async needs no registers), one global round trip for the whole tile
      for (int e = lt; e < kDT * cpr; e += NL) {
        const int r = e / cpr, c = e - r * cpr;
        const int row = sorow[r];
        cp_async16(g_b + (uint32_t)(c >> 3) * kBlkG + sw128_32b((uint32_t)r, (uint32_t)(c & 7)),
                   row >= 0 ? (const void *)(G + (int64_t)row * ldg + c * 4) : (const void *)G, row >= 0 ? 16u : 0u);
      }
      for (int e = lt; e < hn * 8; e += NL) {
        const int h = e >> 3, c = e & 7;
        const bool ok = c * 4 < cvalid;
        cp_async16(halo_b + (uint32_t)h * 128 + c * 16, ok ? (const void *)(A + (int64_t)hids[h] * lda + cb * 32 + c * 4) : (const void *)A,
                   ok ? 16u : 0u);
      }
      cp_async_wait_all();
      if (lt == 0) DW_TL(3);
      // both operands are consumed as they lie: every thread rounds the chunks it copied itself (its own copies are
      // visible to it after wait_all) to the nearest TF32, in place
      for (int e = lt; e < kDT * cpr; e += NL) {
        const int r = e / cpr, c = e - r * cpr;
        const uint32_t ad = g_b + (uint32_t)(c >> 3) * kBlkG + sw128_32b((uint32_t)r, (uint32_t)(c & 7));
        float4 v = dt_lds_f4(ad);
        dt_rna4(v);
        sts_f4(ad, v);
      }
      for (int e = lt; e < hn * 8; e += NL) {
        const uint32_t ad = halo_b + (uint32_t)(e >> 3) * 128 + (e & 7) * 16;
        float4 v = dt_lds_f4(ad);
        dt_rna4(v);
        sts_f4(ad, v);
      }
      fence_proxy_async();   // the G tile is read by the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(tfull + b);
      if (lt == 0) DW_TL(4);
    }
  } else if (warp > kDtLoaders) {
    // ------------------------------------------------------------------------------------------ A^T builders
    const int q = warp & 3, sl = (warp - 1 - kDtLoaders) >> 2;
    const uint32_t a_tm = tmem + ((uint32_t)(q * 32) << 16) + acc_cols + kSlotCols * (uint32_t)sl;
    const int cvalid = min(32, Ca - cb * 32);
    const bool tl_w = sl == 0 && q == 1;
    uint32_t eph = 1;   // parity of `aempty[sl]` to wait for (first use: free)
    for (int it = 0; it < my_n; ++it) {
      const int b = it % nbuf;
      mbar_wait_sleep(tfull + b, ((uint32_t)(it / nbuf)) & 1u, 1000);
      if (tl_w) DW_TL(5);
      const uint32_t halo_l = base + L.halo_off[b] + (uint32_t)lane * 4;
      const uint32_t tab_b = base + L.tab_off[b];
      const int *sorow = reinterpret_cast<const int *>(sm + L.misc_off[b]);
      const bool ovf = sorow[kDT + hcap] != 0;   // flags[0]
      for (int g = 0; g < ng; ++g) {
        if (((it * ng + g) & 1) != sl) continue;
        const int k = kbase + 4 * g + q;
        mbar_wait(aempty + sl, eph);
        eph ^= 1u;
        tc_fence_after();
        if (k < kend) {
          const uint32_t tab_k = tab_b + (uint32_t)k * (kDT * 2);
#pragma unroll 2
          for (int rc = 0; rc < 8; ++rc) {
            // halo rows (16-byte units) of tile rows 16 rc .. 16 rc + 15: the same 32 bytes for every lane (broadcast reads)
            const uint4 t0 = dt_lds_u4(tab_k + rc * 32), t1 = dt_lds_u4(tab_k + rc * 32 + 16);
            const uint32_t w[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
            float v[16];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              v[2 * i] = dt_lds_f1(halo_l + ((w[i] & 0xFFFFu) << 4));
              v[2 * i + 1] = dt_lds_f1(halo_l + ((w[i] >> 16) << 4));
            }
            if (ovf) {   // (rare) some rule of the tile lies beyond the halo capacity: those rows come through the global map
              const uint16_t *lm = lmap + ((int64_t)(ci + it * nper) * 27 + k) * kDT + rc * 16;
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                if (__ldg(lm + i) == kDtOverflow) {
                  const int idx = __ldg(nbr + (int64_t)sorow[rc * 16 + i] * 27 + k);
                  v[i] = lane < cvalid ? dt_rna(__ldg(A + (int64_t)idx * lda + cb * 32 + lane)) : 0.f;
                }
              }
            }
            dt_tmem_st16(a_tm + 16 * rc, v);
          }
          dt_tmem_st_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(afull + sl);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty + b);   // this warp no longer reads the tile's halo / table
      if (tl_w) DW_TL(6);
    }
  } else if (elect_one()) {
    // ------------------------------------------------------------------------------------------ MMA issuer
    const uint64_t dg_hi = make_smem_desc(0, kBlkG, 512, 1) & 0xFFFFFFFF00000000ull;
    const uint32_t dg_lo0 = (uint32_t)(make_smem_desc(0, kBlkG, 512, 1) & 0xFFFFFFFFull);   // (the LBO sits in the low word)
    uint32_t acc_mask = 0;
    for (int it = 0; it < my_n; ++it) {
      const int b = it % nbuf;
      mbar_wait(tfull + b, ((uint32_t)(it / nbuf)) & 1u);
      if (g_dw_timeline && blockIdx.x == 0 && it < 60) g_dw_timeline[16 * it + 7] = clock64();
      const uint32_t g_lo = dg_lo0 + ((base + L.g_off[b]) >> 4);
      for (int g = 0; g < ng; ++g) {
        const int vis = it * ng + g, sl = vis & 1;
        mbar_wait(afull + sl, ((uint32_t)(vis >> 1)) & 1u);
        tc_fence_after();
        const uint32_t a_tm = tmem + acc_cols + kSlotCols * (uint32_t)sl;
        const uint32_t d_tm = tmem + (uint32_t)(g * Cg);
        uint32_t accf = (acc_mask >> g) & 1u;
#pragma unroll
        for (int j = 0; j < kDT / 8; ++j) {   // K = 8 tile rows per MMA: 8 TMEM columns of A^T, 8 rows (1024 bytes) of the G image
          dt_mma_ts(d_tm, a_tm + 8 * j, dg_hi | (uint64_t)(g_lo + j * 64), idesc, accf);
          accf = 1u;
        }
        acc_mask |= 1u << g;
        mma_commit(aempty + sl);
      }
      mma_commit(tempty + b);   // arrives once every MMA reading this G tile has completed
      if (g_dw_timeline && blockIdx.x == 0 && it < 60) g_dw_timeline[16 * it + 8] = clock64();
    }
    *acc_mask_s = acc_mask;
    mma_commit(done);
  }
  __syncwarp();

  // ---------------------------------------------------------------------------------------------- epilogue
  // every CTA writes its whole share of its slab (zeros for accumulators it never touched): the reduction kernel then
  // sums the slabs in a fixed order
  mbar_wait_sleep(done, 0, 1000);
  tc_fence_after();
  __syncthreads();
  if (warp > kDtLoaders) {
    const int q = warp & 3;   // TMEM lane quarter this warp may read
    const uint32_t acc_mask = *acc_mask_s;
    float *slab_p = part + (int64_t)ci * slab;
    const int ca = cb * 32 + lane;
    // two warps share a lane quarter: they alternate over the groups
    const int first = (warp - 1 - kDtLoaders) >> 2;
    for (int g = first; g < ng; g += 2) {
      const int k = kbase + 4 * g + q;
      for (int c0 = 0; c0 < Cg; c0 += 16) {
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * Cg + c0), v);
        if (k < kend && ca < Ca) {
          float *o = slab_p + ((int64_t)k * Ca + ca) * Cg + c0;
          const bool live = (acc_mask >> g) & 1u;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<float4 *>(o + 4 * i) = live ? make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3])
                                                          : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// dW[e] = sum over slabs (ascending: a fixed order, so the result is bit-reproducible)
__global__ void dw_reduce_kernel(const float *__restrict__ part, int nslab, int64_t slab, float *__restrict__ dW) {
  const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (e >= slab) return;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < nslab; ++s) {
    const float4 v = __ldg(reinterpret_cast<const float4 *>(part + (int64_t)s * slab + e));
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  *reinterpret_cast<float4 *>(dW + e) = acc;
}

struct DwPlan {
  int ncb, nsplit_k, gper, nsplit, nper;
  DwLayout L;
  bool ok;
};

static DwPlan dw_plan(int64_t n, int hcap, int Ca, int Cg) {
  DwPlan p;
  p.ok = false;
  if (Cg % 16 != 0 || Cg < 16 || Cg > 256 || Ca % 4 != 0 || Ca < 4 || n <= 0) return p;
  p.ncb = (Ca + 31) / 32;
  int gmax = (512 - 2 * (int)kSlotCols) / Cg;   // accumulator groups beside the two A^T slots
  if (gmax > 7) gmax = 7;
  if (gmax < 1) return p;
  p.nsplit_k = (7 + gmax - 1) / gmax;
  p.gper = (7 + p.nsplit_k - 1) / p.nsplit_k;
  p.nsplit = p.ncb * p.nsplit_k;
  if (p.nsplit > 32) return p;   // each tile would be staged by too many CTAs: the pair-list kernel is the better choice
  const int64_t ntiles = ceil_div(n, kDT);
  p.nper = (int)(kNumSMs / p.nsplit < ntiles ? kNumSMs / p.nsplit : ntiles);
  if (p.nper < 1) p.nper = 1;
  // double-buffered tile data if it fits, else one buffer
  for (int nbuf = 2; nbuf >= 1; --nbuf) {
    p.L = dw_layout(hcap, Cg, nbuf);
    if (p.L.total <= 227 * 1024) { p.ok = true; break; }
  }
  return p;
}

}  // namespace b200scn

using namespace b200scn;

extern "C" {

/* diagnostic (not in the public header): later b200scn_subm_dw_tiled launches record a clock64 timeline of CTA 0 into buf
 * (1024 int64, device); buf = NULL switches it off */
int b200scn_debug_dw_timeline(long long *buf) {
  SCN_CUDA(cudaMemcpyToSymbol(g_dw_timeline, &buf, sizeof(buf)));
  return 0;
}

size_t b200scn_subm_dw_tiled_scratch_bytes(int64_t n, int hcap, int Ca, int Cg) {
  const DwPlan p = dw_plan(n, hcap, Ca, Cg);
  if (!p.ok) return 0;
  return sizeof(float) * (size_t)p.nper * 27 * (size_t)Ca * (size_t)Cg;
}

int b200scn_subm_dw_tiled(const float *A, int64_t lda, const float *G, int64_t ldg, const int32_t *nbr,
                          const int32_t *perm, const uint16_t *lmap, const int32_t *halo_ids, const int32_t *halo_n,
                          int hcap, int64_t n, int Ca, int Cg, float *dW, float *scratch, size_t scratch_bytes,
                          void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= 0) {
    SCN_CUDA(cudaMemsetAsync(dW, 0, sizeof(float) * 27 * (size_t)Ca * Cg, st));
    return 0;
  }
  const DwPlan p = dw_plan(n, hcap, Ca, Cg);
  if (!p.ok) return set_error("subm_dw_tiled: unsupported shape %d x %d (hcap %d)", Ca, Cg, hcap);
  if ((lda & 3) || (ldg & 3) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(G) & 15) ||
      (reinterpret_cast<uintptr_t>(scratch) & 15) || (reinterpret_cast<uintptr_t>(dW) & 15) ||
      (reinterpret_cast<uintptr_t>(lmap) & 15))
    return set_error("subm_dw_tiled: rows must be 16-byte aligned");
  const int64_t slab = (int64_t)27 * Ca * Cg;
  if (scratch_bytes < sizeof(float) * (size_t)p.nper * slab) return set_error("subm_dw_tiled: scratch too small");
  if (n >= ((int64_t)1 << 31)) return set_error("subm_dw_tiled: too many rows");
  const uint32_t idesc = make_idesc_tf32(128, Cg, 0, 1);   // A^T from tensor memory (K along the columns), G MN-major
  const unsigned grid = (unsigned)(p.nsplit * p.nper);
  static bool smem_set = false;
  if (!smem_set) {
    SCN_CUDA(cudaFuncSetAttribute(dw_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    smem_set = true;
  }
  dw_tile_kernel<<<grid, kDtThreads, p.L.total, st>>>(A, lda, G, ldg, nbr, perm, lmap, halo_ids, halo_n, hcap, (int)n, Ca, Cg,
                                                      p.ncb, p.nsplit, p.gper, scratch, slab, p.L, idesc);
  SCN_CHECK_LAUNCH("subm_dw_tiled");
  dw_reduce_kernel<<<(unsigned)ceil_div(slab / 4, 256), 256, 0, st>>>(scratch, p.nper, slab, dW);
  SCN_CHECK_LAUNCH("subm_dw_reduce");
  count_launch(2);
  return 0;
}

}  // extern "C"
