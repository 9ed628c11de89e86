// common.cuh -- shared helpers for the b200scn CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/b200scn.h"

namespace b200scn {

extern thread_local char g_err[512];
int set_error(const char *fmt, ...);
void count_launch(int n);  // kernels enqueued by this library (b200scn_launch_count)

#define SCN_CHECK_LAUNCH(name)                                              \
  do {                                                                      \
    cudaError_t e__ = cudaGetLastError();                                   \
    if (e__ != cudaSuccess)                                                 \
      return b200scn::set_error("%s: %s", name, cudaGetErrorString(e__));   \
  } while (0)

#define SCN_CUDA(call)                                                      \
  do {                                                                      \
    cudaError_t e__ = (call);                                               \
    if (e__ != cudaSuccess)                                                 \
      return b200scn::set_error("%s: %s", #call, cudaGetErrorString(e__));  \
  } while (0)

// Tuning / test knobs, set through b200scn_set_option (never read from the environment on the launch path).
struct Options {
  int tc_tma = 0;        // 1: gather kernel with TMA tile::gather4 producers (measured 2.2x slower; kept for comparison)
  int tc_msub = 0;       // 0 auto, 1 / 2: accumulators per CTA of the gather kernel
  int tc_nsplit = 0;     // 0 auto, >0: split the offsets of the gather kernel over this many CTAs
  int dw_chunk = 4096;   // pairs per CTA of the pair-list weight-gradient kernel
  int dw_pairs = 32;     // pairs per pipeline stage of that kernel (32 or 64)
  int dw_dbg = 0;        // knockout experiments of that kernel (WRONG RESULTS): 1 no MMA issue, 2 no gathers
  int halo_pf = -1;      // L2 prefetch distance of the tiled kernel in tiles (-1 auto, 0 off)
  int halo_dbg = 0;      // knockout experiments of the tiled kernel (WRONG RESULTS): 1 no A build, 2 no MMA, 4 no halo reads, 8 no TMEM stores
  int halo_one_cta = 0;  // 1: force one CTA per SM in the tiled kernel
};
extern Options g_opt;

constexpr int kNumSMs = 148;
constexpr uint64_t kEmptyKey = 0xFFFFFFFFFFFFFFFFull;

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return k;
}

__device__ __forceinline__ uint64_t make_key(uint32_t x, uint32_t y, uint32_t z, uint32_t b) {
  return ((uint64_t)b << 48) | ((uint64_t)x << 32) | ((uint64_t)y << 16) | (uint64_t)z;
}
__device__ __forceinline__ void split_key(uint64_t k, int &x, int &y, int &z, int &b) {
  z = (int)(k & 0xFFFF); y = (int)((k >> 16) & 0xFFFF); x = (int)((k >> 32) & 0xFFFF);
  b = (int)(k >> 48);
}

// read-only hash lookup: id or -1
__device__ __forceinline__ int hash_lookup(const uint64_t *__restrict__ hkeys,
                                           const int32_t *__restrict__ hvals, int64_t cap,
                                           uint64_t key) {
  uint64_t s = mix64(key) & (uint64_t)(cap - 1);
  for (;;) {
    uint64_t k = __ldg(hkeys + s);
    if (k == key) return __ldg(hvals + s);
    if (k == kEmptyKey) return -1;
    s = (s + 1) & (uint64_t)(cap - 1);
  }
}

__device__ __forceinline__ int live_count(const int32_t *n_dev, int64_t n_max) {
  return n_dev ? *n_dev : (int)n_max;
}

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// --------------------------------------------------------------------------------------------
// Exclusive prefix sum over int flags produced on the fly by Loader(i) (i < n), with a fused
// Writer(i, flag, exclusive_position).  Three launches: chunk reduce, spine, chunk scan.
// Chunk = 1024 threads x 4 items.
constexpr int kScanThreads = 1024;
constexpr int kScanItems = 4;
constexpr int kScanChunk = kScanThreads * kScanItems;

__device__ __forceinline__ int block_exclusive_scan(int v, int *smem /*>=32*/, int &block_total) {
  // inclusive warp scan
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = lane < (blockDim.x >> 5) ? smem[lane] : 0;
    int winc = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, winc, d);
      if (lane >= d) winc += t;
    }
    smem[lane] = winc - w;  // exclusive warp offsets
    if (lane == 31) smem[32] = winc;
  }
  __syncthreads();
  int res = inc - v + smem[warp];
  block_total = smem[32];
  __syncthreads();
  return res;
}

template <class Loader>
__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(Loader ld, int64_t n_max,
                                                                   const int32_t *n_dev,
                                                                   int32_t *block_sums) {
  __shared__ int sm[33];
  const int64_t n = n_dev ? (int64_t)ld.live(*n_dev) : n_max;
  const int64_t base = (int64_t)blockIdx.x * kScanChunk + (int64_t)threadIdx.x * kScanItems;
  int s = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j)
    if (base + j < n) s += ld(base + j);
  int tot;
  block_exclusive_scan(s, sm, tot);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

// single block: exclusive scan of block_sums[0..nb) in place; total -> total_dev (may be NULL)
__global__ void __launch_bounds__(kScanThreads) scan_spine_kernel(int32_t *block_sums, int nb,
                                                                  int32_t *total_dev);

template <class Loader, class Writer>
__global__ void __launch_bounds__(kScanThreads) scan_down_kernel(Loader ld, Writer wr, int64_t n_max,
                                                                 const int32_t *n_dev,
                                                                 const int32_t *block_sums) {
  __shared__ int sm[33];
  const int64_t n = n_dev ? (int64_t)ld.live(*n_dev) : n_max;
  const int64_t base = (int64_t)blockIdx.x * kScanChunk + (int64_t)threadIdx.x * kScanItems;
  int f[kScanItems];
  int s = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    f[j] = (base + j < n) ? ld(base + j) : 0;
    s += f[j];
  }
  int tot;
  int pos = block_exclusive_scan(s, sm, tot) + block_sums[blockIdx.x];
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    if (base + j < n) wr(base + j, f[j], pos);
    pos += f[j];
  }
}

static inline size_t scan_scratch_ints(int64_t n_max) { return (size_t)ceil_div(n_max, kScanChunk) + 1; }

template <class Loader, class Writer>
int scan_flags(Loader ld, Writer wr, int64_t n_max, const int32_t *n_dev, int32_t *block_sums,
               int32_t *total_dev, cudaStream_t st) {
  if (n_max <= 0) {
    if (total_dev) SCN_CUDA(cudaMemsetAsync(total_dev, 0, sizeof(int32_t), st));
    return 0;
  }
  const int nb = (int)ceil_div(n_max, kScanChunk);
  scan_reduce_kernel<<<nb, kScanThreads, 0, st>>>(ld, n_max, n_dev, block_sums);
  scan_spine_kernel<<<1, kScanThreads, 0, st>>>(block_sums, nb, total_dev);
  scan_down_kernel<<<nb, kScanThreads, 0, st>>>(ld, wr, n_max, n_dev, block_sums);
  SCN_CHECK_LAUNCH("scan_flags");
  count_launch(3);
  return 0;
}

}  // namespace b200scn
