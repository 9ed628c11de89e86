// dw_tile.cu -- TILE-STATIONARY weight gradient of the submanifold 3x3x3 convolution (replaces
// upstream scn's dConvolution_KMxKN_backward_dW: one launch per offset, global atomicAdd per rule).
//
//   dW[k] = sum over rules (in, out) of offset k of  A[in,:]^T (x) G[out,:]
//
// The pair-list kernel (conv_tc.cu: pair_dw_tc_kernel) fetches TWO rows from L2 for every rule -- 4.8-5.8 TB/s of gathered
// rows at levels 1-2 of the benchmark, its bound.  Here the work is cut along the same Morton-ordered 128-row tiles and
// per-tile halo lists the forward kernel uses (conv_halo.cu: b200scn_tile_plan): a persistent CTA stages a tile's
// gradient rows G (the tile's own rows) and its DISTINCT input rows (halo) in shared memory ONCE, and every rule of the
// tile is then served from shared memory:
//   * the G tile is the B operand of every MMA of the tile as it lies (MN-major, SWIZZLE_128B_BASE32B);
//   * the A^T operand of a stage = [32 tile rows (K)] x [4 offsets x 32 input channels (M = 128)] is assembled by copying
//     halo rows shared -> shared; a quarter-warp copies one whole 128-byte row, so reads and writes are conflict-free
//     whatever rows the rulebook names (unlike lane = row layouts);
//   * accumulators for ALL offsets of the CTA stay in tensor memory across all its tiles (7 groups of 4 offsets x Cg
//     columns); each CTA writes its partial ONCE, and a second kernel sums the partials in a fixed order: the result is
//     bit-reproducible (no fp32 atomics).
// A CTA owns one 32-channel block of A (and, when 7 x Cg columns exceed tensor memory, a subset of the offsets); tiles are
// dealt round-robin to the CTAs of a split.  Both operands are rounded to the nearest TF32 as they are staged (the loaders
// go through registers), so the tensor core's truncation never sees an unrounded value.
#include "common.cuh"
#include "tc_common.cuh"

namespace b200scn {

using namespace tc;

constexpr int kDT = 128;                  // tile rows
constexpr int kDtMap = 27 * kDT;          // lmap entries per tile
constexpr int kDtLoaders = 8, kDtProducers = 8;
constexpr int kDtThreads = 32 * (1 + kDtLoaders + kDtProducers);
constexpr uint32_t kBlkA = 32 * 128;      // one M-block of an A^T stage: 32 rows x 128 bytes
constexpr uint32_t kBlkG = kDT * 128;     // one 32-column block of the G tile: 128 rows x 128 bytes
constexpr uint32_t kStageBytes = 4 * kBlkA;
constexpr uint16_t kDtAbsent = 0xFFFF, kDtOverflow = 0xFFFE;

struct DwLayout {
  uint32_t halo_off[2], g_off[2], lmap_off[2], misc_off[2], stage_off, bar_off, total;
  int nbuf, nst;
};
// misc of a tile buffer: sorow[128] int | hids[hcap] int | cm[27*4] bytes (chunk s of offset k has a present row)
static DwLayout dw_layout(int hcap, int Cg, int nbuf, int nst) {
  DwLayout L;
  const uint32_t gb = (uint32_t)(Cg + 31) / 32;
  uint32_t o = 0;
  for (int b = 0; b < 2; ++b) {
    const bool live = b < nbuf;
    if (live) o = (o + 1023) & ~1023u;            // the swizzled operand image needs a 1024-byte aligned base
    L.g_off[b] = live ? o : L.g_off[0];
    if (live) o += gb * kBlkG;                    // 1024-byte aligned blocks
    L.halo_off[b] = live ? o : L.halo_off[0];
    if (live) o += (uint32_t)hcap * 128;
    L.lmap_off[b] = live ? o : L.lmap_off[0];
    if (live) o += (kDtMap * 2 + 15) & ~15;
    L.misc_off[b] = live ? o : L.misc_off[0];
    if (live) o += ((512 + (uint32_t)hcap * 4 + 27 * 4) + 127) & ~127u;
  }
  o = (o + 1023) & ~1023u;
  L.stage_off = o;
  o += (uint32_t)nst * kStageBytes;
  L.bar_off = o;
  L.total = o + 256 + 1024;   // barriers + alignment slack
  L.nbuf = nbuf;
  L.nst = nst;
  return L;
}

__device__ __forceinline__ float4 dt_lds_f4(uint32_t saddr) {
  float4 v;
  // not volatile (independent loads of a batch may be scheduled together) but a memory reader: never hoisted or merged
  // across the barrier waits and stores around it
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ float dt_rna(float v) {   // nearest TF32, ties away, as two full-rate integer instructions
  return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ void dt_rna4(float4 &v) { v.x = dt_rna(v.x); v.y = dt_rna(v.y); v.z = dt_rna(v.z); v.w = dt_rna(v.w); }
__device__ __forceinline__ void dt_named_bar(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Diagnostic: clock64 timeline of CTA 0 (b200scn_debug_dw_timeline): per tile it, slots 16*it + {0 loader waits for the
// buffer, 1 buffer free, 2 plan slice in, 3 copies landed, 4 tile published, 5 producer 0 sees the tile, 6 producer 0 done,
// 7 MMA thread sees the tile, 8 MMA thread done issuing, 9 stages of the tile}.
__device__ long long *g_dw_timeline = nullptr;
#define DW_TL(slot) do { if (tl && it < 60) tl[16 * it + (slot)] = clock64(); } while (0)

// Warp 0: MMA issuer (and TMEM owner).  Warps 1..8: tile loaders (global -> registers -> shared, one tile ahead).
// Warps 9..16: A^T stage producers (shared -> shared), then the epilogue.
template <uint32_t NT>
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
  const int nbuf = L.nbuf, nst = L.nst;
  uint64_t *tfull = reinterpret_cast<uint64_t *>(sm + L.bar_off);   // tile buffer loaded (one arrival per loader warp)
  uint64_t *tempty = tfull + 2;                                     // tile buffer free (producer warps + the MMA commit)
  uint64_t *full = tempty + 2;                                      // A^T stage written
  uint64_t *empty = full + 4;                                       // A^T stage consumed (tcgen05.commit)
  uint64_t *done = empty + 4;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done + 1);
  uint32_t *acc_mask_s = tmem_slot + 1;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int sp = blockIdx.x % nsplit, ci = blockIdx.x / nsplit, nper = gridDim.x / nsplit;
  const int cb = sp % ncb, ks = sp / ncb;
  const int kbase = 4 * gper * ks;
  const int kend = min(27, kbase + 4 * gper);
  const int ng = (kend - kbase + 3) >> 2;
  const int ntiles = (n_rows + kDT - 1) / kDT;
  const int my_n = ci < ntiles ? (ntiles - ci + nper - 1) / nper : 0;
  const int wps = kDtProducers / nst;   // producer warps per stage
  long long *tl = (g_dw_timeline && blockIdx.x == 0 && lane == 0) ? g_dw_timeline : nullptr;

  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull + b, kDtLoaders);
      mbar_init(tempty + b, kDtProducers + 1);
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(full + s, wps);
      mbar_init(empty + s, 1);
    }
    mbar_init(done, 1);
    *acc_mask_s = 0;
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<NT>(tmem_slot);
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
      // long waits suspend (try_wait with a time hint) instead of spinning: 8 loader warps polling in a tight loop took the
      // issue slots the MMA-issuing lane and the producers needed (measured: 1180 cycles per stage with spinning waits)
      mbar_wait_sleep(tempty + b, (((uint32_t)(it / nbuf)) & 1u) ^ 1u, 2000);
      if (lt == 0) DW_TL(1);
      const uint32_t g_b = base + L.g_off[b], halo_b = base + L.halo_off[b], lmap_b = base + L.lmap_off[b];
      int *sorow = reinterpret_cast<int *>(sm + L.misc_off[b]);
      int *hids = sorow + kDT;
      uint8_t *cm = reinterpret_cast<uint8_t *>(hids + hcap);
      const uint16_t *slmap = reinterpret_cast<const uint16_t *>(sm + L.lmap_off[b]);
      const int hn = __ldg(halo_n + tile);
      // phase 1: the tile's plan slice (asynchronous copies) and the row ids phase 2 gathers through
      {
        const uint16_t *src = lmap + (int64_t)tile * kDtMap;   // 6912 bytes, 16-byte aligned
        for (int e = lt; e < kDtMap * 2 / 16; e += NL) cp_async16(lmap_b + e * 16, src + e * 8, 16u);
        if (lt < kDT) sorow[lt] = row0 + lt < n_rows ? __ldg(perm + row0 + lt) : -1;
        const int32_t *ids = halo_ids + (int64_t)tile * hcap;
        for (int h = lt; h < hn; h += NL) hids[h] = __ldg(ids + h);
      }
      dt_named_bar(2, NL);
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
      dt_named_bar(2, NL);
      if (lt == 0) DW_TL(3);
      // the G tile is consumed by the tensor core as it lies: round it to the nearest TF32 in place (the A rows are
      // rounded by the stage producers on their way through registers)
      for (int e0 = lt; e0 < kDT * cpr; e0 += 4 * NL) {
        float4 v[4];
        uint32_t ad[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int e = e0 + u * NL;
          const int r = e / cpr, c = e - r * cpr;
          ad[u] = g_b + (uint32_t)(c >> 3) * kBlkG + sw128_32b((uint32_t)r, (uint32_t)(c & 7));
          if (e < kDT * cpr) v[u] = dt_lds_f4(ad[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (e0 + u * NL < kDT * cpr) {
            dt_rna4(v[u]);
            sts_f4(ad[u], v[u]);
          }
        }
      }
      // which 32-row chunks of which offsets hold any rule
      for (int idx = (lt >> 5); idx < 27 * 4; idx += kDtLoaders) {
        const int k = idx >> 2, s = idx & 3;
        const unsigned any = __ballot_sync(0xffffffffu, slmap[k * kDT + 32 * s + lane] != kDtAbsent);
        if (lane == 0) cm[idx] = any ? 1 : 0;
      }
      fence_proxy_async();   // the G tile is read by the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(tfull + b);
      if (lt == 0) DW_TL(4);
    }
  } else if (warp > kDtLoaders) {
    // ------------------------------------------------------------------------------------------ A^T stage producers
    const int pw = warp - 1 - kDtLoaders;         // 0 .. 7
    const int my_stage = pw % nst, part_i = pw / nst;
    const int rpw = 32 / wps;                     // rows of a stage this warp copies (per M-block)
    const int c = lane & 7, rl = lane >> 3;
    const int cvalid = min(32, Ca - cb * 32);
    int st = 0;
    for (int it = 0; it < my_n; ++it) {
      const int b = it % nbuf;
      mbar_wait_sleep(tfull + b, ((uint32_t)(it / nbuf)) & 1u, 1000);
      if (pw == 0) DW_TL(5);
      const uint32_t halo_b = base + L.halo_off[b];
      const int *sorow = reinterpret_cast<const int *>(sm + L.misc_off[b]);
      const uint8_t *cm = reinterpret_cast<const uint8_t *>(sorow + kDT + hcap);
      const uint16_t *slmap = reinterpret_cast<const uint16_t *>(sm + L.lmap_off[b]);
      for (int g = 0; g < ng; ++g) {
        const int k0 = kbase + 4 * g;
        for (int s = 0; s < 4; ++s) {
          bool pres[4];
          bool any = false;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            pres[q] = (k0 + q < kend) && cm[(k0 + q) * 4 + s];
            any = any || pres[q];
          }
          if (!any) continue;
          if (st % nst == my_stage) {
            const int buf = my_stage;
            mbar_wait_sleep(empty + buf, (((uint32_t)(st / nst)) & 1u) ^ 1u, 200);
            const uint32_t st_base = base + L.stage_off + (uint32_t)buf * kStageBytes;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int k = k0 + q;
              // up to 4 rows per lane and M-block: all slots, then all loads, then all stores (independent chains)
              uint32_t slot[4];
              float4 v[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int rt = 32 * s + part_i * rpw + rl + 4 * i;
                slot[i] = (pres[q] && i < rpw / 4) ? slmap[k * kDT + rt] : kDtAbsent;
              }
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (slot[i] < kDtOverflow) v[i] = dt_lds_f4(halo_b + slot[i] * 128 + c * 16);
              }
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                if (i < rpw / 4) {
                  const int r = part_i * rpw + rl + 4 * i;
                  if (slot[i] == kDtOverflow) {              // beyond the halo capacity: through the global map (rare)
                    const int idx = __ldg(nbr + (int64_t)sorow[32 * s + r] * 27 + k);
                    if (c * 4 < cvalid) v[i] = ldg_f4(A + (int64_t)idx * lda + cb * 32 + c * 4);
                  }
                  dt_rna4(v[i]);
                  sts_f4(st_base + (uint32_t)q * kBlkA + sw128_32b((uint32_t)r, (uint32_t)c), v[i]);
                }
              }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(full + buf);
          }
          ++st;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty + b);   // this warp no longer reads the tile's halo / plan
      if (pw == 0) DW_TL(6);
    }
  } else if (elect_one()) {
    // ------------------------------------------------------------------------------------------ MMA issuer
    const uint64_t da_hi = make_smem_desc(0, kBlkA, 512, 1) & 0xFFFFFFFF00000000ull;
    const uint64_t dg_hi = make_smem_desc(0, kBlkG, 512, 1) & 0xFFFFFFFF00000000ull;
    const uint32_t da_lo0 = (uint32_t)(make_smem_desc(0, kBlkA, 512, 1) & 0xFFFFFFFFull);   // (the LBO sits in the low word)
    const uint32_t dg_lo0 = (uint32_t)(make_smem_desc(0, kBlkG, 512, 1) & 0xFFFFFFFFull);
    uint32_t acc_mask = 0;
    int st = 0;
    for (int it = 0; it < my_n; ++it) {
      const int b = it % nbuf;
      mbar_wait(tfull + b, ((uint32_t)(it / nbuf)) & 1u);
      if (g_dw_timeline && blockIdx.x == 0 && it < 60) g_dw_timeline[16 * it + 7] = clock64();
      const int st_begin = st;
      const uint8_t *cm = reinterpret_cast<const uint8_t *>(reinterpret_cast<const int *>(sm + L.misc_off[b]) + kDT + hcap);
      const uint32_t g_b = base + L.g_off[b];
      for (int g = 0; g < ng; ++g) {
        const int k0 = kbase + 4 * g;
        for (int s = 0; s < 4; ++s) {
          bool any = false;
#pragma unroll
          for (int q = 0; q < 4; ++q) any = any || ((k0 + q < kend) && cm[(k0 + q) * 4 + s]);
          if (!any) continue;
          const int buf = st % nst;
          mbar_wait(full + buf, ((uint32_t)(st / nst)) & 1u);
          tc_fence_after();
          const uint32_t a_lo = da_lo0 + ((base + L.stage_off + (uint32_t)buf * kStageBytes) >> 4);
          const uint32_t g_lo = dg_lo0 + ((g_b + (uint32_t)s * 32u * 128u) >> 4);
          uint32_t accf = (acc_mask >> g) & 1u;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            mma_tf32(tmem + (uint32_t)(g * Cg), da_hi | (uint64_t)(a_lo + j * 64), dg_hi | (uint64_t)(g_lo + j * 64), idesc, accf);
            accf = 1u;
          }
          acc_mask |= 1u << g;
          mma_commit(empty + buf);
          ++st;
        }
      }
      mma_commit(tempty + b);   // arrives once every MMA reading this G tile has completed
      if (g_dw_timeline && blockIdx.x == 0 && it < 60) {
        g_dw_timeline[16 * it + 8] = clock64();
        g_dw_timeline[16 * it + 9] = st - st_begin;
      }
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
    const int pw = warp - 1 - kDtLoaders;
    const int q = warp & 3;   // TMEM lane quarter this warp may read
    const uint32_t acc_mask = *acc_mask_s;
    float *slab_p = part + (int64_t)ci * slab;
    const int ca = cb * 32 + lane;
    // two warps share a lane quarter: they alternate over the groups
    int first = 0;
    for (int w2 = 9; w2 < warp; ++w2) first += ((w2 & 3) == q);
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
    (void)pw;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<NT>(tmem);
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
  uint32_t nt;
  bool ok;
};

static DwPlan dw_plan(int64_t n, int hcap, int Ca, int Cg) {
  DwPlan p;
  p.ok = false;
  if (Cg % 16 != 0 || Cg < 16 || Cg > 256 || Ca % 4 != 0 || Ca < 4 || n <= 0) return p;
  p.ncb = (Ca + 31) / 32;
  int gmax = 512 / Cg;
  if (gmax > 7) gmax = 7;
  if (gmax < 1) return p;
  p.nsplit_k = (7 + gmax - 1) / gmax;
  p.gper = (7 + p.nsplit_k - 1) / p.nsplit_k;
  p.nsplit = p.ncb * p.nsplit_k;
  if (p.nsplit > 8) return p;   // each tile would be staged by too many CTAs: the pair-list kernel is the better choice
  const int64_t ntiles = ceil_div(n, kDT);
  p.nper = (int)(kNumSMs / p.nsplit < ntiles ? kNumSMs / p.nsplit : ntiles);
  if (p.nper < 1) p.nper = 1;
  // double-buffered tile data and four stages if they fit, else fewer
  const int tries[4][2] = {{2, 4}, {2, 2}, {1, 4}, {1, 2}};
  for (int t = 0; t < 4; ++t) {
    p.L = dw_layout(hcap, Cg, tries[t][0], tries[t][1]);
    if (p.L.total <= 227 * 1024) { p.ok = true; break; }
  }
  const int cols = p.gper * Cg;
  p.nt = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
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
      (reinterpret_cast<uintptr_t>(scratch) & 15) || (reinterpret_cast<uintptr_t>(dW) & 15))
    return set_error("subm_dw_tiled: rows must be 16-byte aligned");
  const int64_t slab = (int64_t)27 * Ca * Cg;
  if (scratch_bytes < sizeof(float) * (size_t)p.nper * slab) return set_error("subm_dw_tiled: scratch too small");
  if (n >= ((int64_t)1 << 31)) return set_error("subm_dw_tiled: too many rows");
  const uint32_t idesc = make_idesc_tf32(128, Cg, 1, 1);
  const unsigned grid = (unsigned)(p.nsplit * p.nper);
#define SCN_LAUNCH_DT(NT)                                                                                          \
  do {                                                                                                             \
    auto kern = dw_tile_kernel<NT>;                                                                                \
    static bool smem_set = false;                                                                                  \
    if (!smem_set) {                                                                                               \
      SCN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));               \
      smem_set = true;                                                                                             \
    }                                                                                                              \
    kern<<<grid, kDtThreads, p.L.total, st>>>(A, lda, G, ldg, nbr, perm, lmap, halo_ids, halo_n, hcap, (int)n, Ca, \
                                              Cg, p.ncb, p.nsplit, p.gper, scratch, slab, p.L, idesc);             \
  } while (0)
  switch (p.nt) {
    case 32: SCN_LAUNCH_DT(32); break;
    case 64: SCN_LAUNCH_DT(64); break;
    case 128: SCN_LAUNCH_DT(128); break;
    case 256: SCN_LAUNCH_DT(256); break;
    default: SCN_LAUNCH_DT(512); break;
  }
#undef SCN_LAUNCH_DT
  SCN_CHECK_LAUNCH("subm_dw_tiled");
  dw_reduce_kernel<<<(unsigned)ceil_div(slab / 4, 256), 256, 0, st>>>(scratch, p.nper, slab, dW);
  SCN_CHECK_LAUNCH("subm_dw_reduce");
  count_launch(2);
  return 0;
}

}  // extern "C"
