// grid.cu -- active-site numbering and rulebooks on the GPU.
//
// Replaces upstream scn's CPU-only Metadata / dense_hash_map rulebook builders that sit
// behind scn.InputLayer / SubmanifoldConvolution / Convolution (models/SparseConvNet.py:61,62,137).
// Everything here is HBM/L2-bound integer work: one open-addressing hash over packed 64-bit site keys,
// an atomicMin on the first row that touches a site, and a prefix sum over "I am the first row" flags,
// which yields first-occurrence ids with no sort at all.
#include <stdarg.h>

#include "common.cuh"

namespace b200scn {

thread_local char g_err[512] = "";

int set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

static unsigned long long g_launches = 0;
void count_launch(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }
unsigned long long launch_count() { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

__global__ void __launch_bounds__(kScanThreads) scan_spine_kernel(int32_t *block_sums, int nb,
                                                                  int32_t *total_dev) {
  __shared__ int sm[33];
  int carry = 0;
  for (int base = 0; base < nb; base += kScanThreads) {
    int i = base + threadIdx.x;
    int v = i < nb ? block_sums[i] : 0;
    int tot;
    int ex = block_exclusive_scan(v, sm, tot);
    if (i < nb) block_sums[i] = ex + carry;
    carry += tot;
  }
  if (threadIdx.x == 0 && total_dev) *total_dev = carry;
}

// ------------------------------------------------------------------------------------------ pack
__global__ void pack_coords_kernel(const int64_t *__restrict__ coords, int64_t P, int ncols,
                                   int64_t spatial_size, uint64_t *__restrict__ keys,
                                   int32_t *err_flag) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= P) return;
  const int64_t *c = coords + r * ncols;
  int64_t x = c[0], y = c[1], z = c[2], b = ncols == 4 ? c[3] : 0;
  bool bad = x < 0 || y < 0 || z < 0 || b < 0 || x >= spatial_size || y >= spatial_size ||
             z >= spatial_size || b > 32767;
  if (bad) {
    atomicOr(err_flag, 1);
    x = y = z = 0; b = 0;
  }
  keys[r] = make_key((uint32_t)x, (uint32_t)y, (uint32_t)z, (uint32_t)b);
}

// ------------------------------------------------------------------------------------------ dedup
__global__ void hash_insert_kernel(const uint64_t *__restrict__ keys, int64_t n_max,
                                   const int32_t *n_dev, unsigned long long *hkeys, int32_t *hvals,
                                   int64_t cap, int32_t *slot_of_row) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= live_count(n_dev, n_max)) return;
  const uint64_t key = keys[r];
  uint64_t s = mix64(key) & (uint64_t)(cap - 1);
  for (;;) {
    unsigned long long prev = *(volatile unsigned long long *)(hkeys + s);
    if (prev == kEmptyKey) prev = atomicCAS(hkeys + s, (unsigned long long)kEmptyKey, (unsigned long long)key);
    if (prev == kEmptyKey || prev == key) break;
    s = (s + 1) & (uint64_t)(cap - 1);
  }
  atomicMin(hvals + s, (int32_t)r);
  slot_of_row[r] = (int32_t)s;
}

struct FirstRowLoader {
  const int32_t *hvals, *slot_of_row;
  __device__ int live(int n) const { return n; }
  __device__ int operator()(int64_t i) const { return hvals[slot_of_row[i]] == (int32_t)i; }
};
struct FirstRowWriter {
  const uint64_t *keys;
  uint64_t *ukeys;
  int32_t *rank, *first_row;
  __device__ void operator()(int64_t i, int flag, int pos) const {
    rank[i] = pos;
    if (flag) { ukeys[pos] = keys[i]; first_row[pos] = (int32_t)i; }
  }
};

__global__ void assign_ids_kernel(int64_t n_max, const int32_t *n_dev, const int32_t *__restrict__ hvals,
                                  const int32_t *__restrict__ slot_of_row,
                                  const int32_t *__restrict__ rank, int32_t *__restrict__ id_of_row,
                                  int32_t *count, int32_t *last_row) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= live_count(n_dev, n_max)) return;
  int id = rank[hvals[slot_of_row[r]]];
  id_of_row[r] = id;
  if (count) atomicAdd(count + id, 1);
  if (last_row) atomicMax(last_row + id, (int32_t)r);
}

__global__ void finalize_table_kernel(int64_t n_max, const int32_t *n_unique_dev,
                                      const int32_t *__restrict__ first_row,
                                      const int32_t *__restrict__ slot_of_row, int32_t *hvals) {
  int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= *n_unique_dev) return;
  hvals[slot_of_row[first_row[id]]] = (int32_t)id;
}

// ------------------------------------------------------------------------------------------ strided
__global__ void coarse_keys_kernel(const uint64_t *__restrict__ ukeys, int64_t n_max,
                                   const int32_t *n_dev, int s, uint64_t *__restrict__ ckeys,
                                   uint8_t *__restrict__ off) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= live_count(n_dev, n_max)) return;
  int x, y, z, b;
  split_key(ukeys[i], x, y, z, b);
  int cx = x / s, cy = y / s, cz = z / s;
  ckeys[i] = make_key(cx, cy, cz, b);
  off[i] = (uint8_t)(((x - cx * s) * s + (y - cy * s)) * s + (z - cz * s));
}

__global__ void child_map_kernel(const int32_t *__restrict__ parent, const uint8_t *__restrict__ off,
                                 int64_t nf_max, const int32_t *nf_dev, int K, int32_t *child) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= live_count(nf_dev, nf_max)) return;
  child[(int64_t)parent[i] * K + off[i]] = (int32_t)i;
}

// ------------------------------------------------------------------------------------------ subm map
__global__ void __launch_bounds__(256) subm_map_kernel(const uint64_t *__restrict__ ukeys, int64_t n_max,
                                                       const int32_t *n_dev,
                                                       const uint64_t *__restrict__ hkeys,
                                                       const int32_t *__restrict__ hvals, int64_t cap,
                                                       int spatial_size, int32_t *__restrict__ nbr,
                                                       int32_t *counts27) {
  __shared__ int cnt[27];
  if (threadIdx.x < 27) cnt[threadIdx.x] = 0;
  __syncthreads();
  const int64_t n = live_count(n_dev, n_max);
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n * 27) {
    int64_t o = t / 27;
    int k = (int)(t - o * 27);
    int x, y, z, b;
    split_key(__ldg(ukeys + o), x, y, z, b);
    int dx = k / 9 - 1, dy = (k / 3) % 3 - 1, dz = k % 3 - 1;
    int v = -1;
    if (k == 13) {
      v = (int)o;
    } else {
      int nx = x + dx, ny = y + dy, nz = z + dz;
      if (nx >= 0 && ny >= 0 && nz >= 0 && nx < spatial_size && ny < spatial_size && nz < spatial_size)
        v = hash_lookup(hkeys, hvals, cap, make_key(nx, ny, nz, b));
    }
    nbr[t] = v;
    if (v >= 0 && counts27) atomicAdd(&cnt[k], 1);
  }
  __syncthreads();
  if (counts27 && threadIdx.x < 27 && cnt[threadIdx.x]) atomicAdd(counts27 + threadIdx.x, cnt[threadIdx.x]);
}

// ------------------------------------------------------------------------------------------ pair lists
// `order` (optional): the list of offset k enumerates rows order[0], order[1], ... instead of 0, 1, ... -- the canonical
// scn form (ascending row) is order == NULL; the weight-gradient kernel walks the rows along the Morton curve instead.
struct PairLoader {
  const int32_t *map, *order;
  int64_t n;
  int K;
  __device__ int live(int nn) const { return nn; }
  __device__ int operator()(int64_t t) const {
    int64_t k = t / n, o = t - k * n;
    const int64_t r = order ? order[o] : o;
    return map[r * K + k] >= 0;
  }
};
struct PairWriter {
  const int32_t *map, *order;
  int64_t n;
  int K;
  int32_t *pair_in, *pair_out, *offsets;
  int32_t *blk;          // optional: list position at which row block b (row_block rows of `order`) of offset k starts
  int row_block, nblk;   // row_block is a power of two
  __device__ void operator()(int64_t t, int flag, int pos) const {
    int64_t k = t / n, o = t - k * n;
    if (o == 0) offsets[k] = pos;
    if (blk && (o & (row_block - 1)) == 0) blk[k * nblk + o / row_block] = pos;
    const int64_t r = order ? order[o] : o;
    if (flag) { pair_in[pos] = map[r * K + k]; pair_out[pos] = (int32_t)r; }
  }
};

}  // namespace b200scn

using namespace b200scn;

extern "C" {

const char *b200scn_last_error(void) { return g_err; }
int b200scn_version(void) { return 1; }
unsigned long long b200scn_launch_count(void) { return launch_count(); }

int64_t b200scn_hash_capacity(int64_t n) {
  int64_t cap = 1024;
  while (cap < 2 * n) cap <<= 1;
  return cap;
}

size_t b200scn_grid_scratch_bytes(int64_t n_max) {
  // slot_of_row, rank, first_row : n_max ints each ; scan block sums
  return sizeof(int32_t) * (3 * (size_t)(n_max > 0 ? n_max : 1) + scan_scratch_ints(n_max) + 64);
}

int b200scn_pack_coords(const int64_t *coords, int64_t P, int ncols, int64_t spatial_size,
                        uint64_t *keys, int32_t *err_flag, void *stream) {
  if (ncols != 3 && ncols != 4) return set_error("pack_coords: coords must have 3 or 4 columns, got %d", ncols);
  if (spatial_size < 1 || spatial_size > 65536) return set_error("pack_coords: spatial_size %lld outside [1,65536]", (long long)spatial_size);
  if (P == 0) return 0;
  pack_coords_kernel<<<(unsigned)ceil_div(P, 256), 256, 0, (cudaStream_t)stream>>>(coords, P, ncols, spatial_size, keys, err_flag);
  SCN_CHECK_LAUNCH("pack_coords");
  count_launch(1);
  return 0;
}

int b200scn_grid_build(const uint64_t *keys, int64_t n_max, const int32_t *n_dev, uint64_t *hkeys,
                       int32_t *hvals, int64_t cap, int32_t *id_of_row, uint64_t *ukeys,
                       int32_t *first_row, int32_t *last_row, int32_t *count,
                       int32_t *n_unique_dev, void *scratch, size_t scratch_bytes, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (cap < 2 * n_max || (cap & (cap - 1))) return set_error("grid_build: cap %lld must be a power of two >= 2*n_max", (long long)cap);
  if (scratch_bytes < b200scn_grid_scratch_bytes(n_max)) return set_error("grid_build: scratch too small");
  if (!n_unique_dev) return set_error("grid_build: n_unique_dev is required");
  SCN_CUDA(cudaMemsetAsync(hkeys, 0xFF, sizeof(uint64_t) * cap, st));
  SCN_CUDA(cudaMemsetAsync(hvals, 0x7F, sizeof(int32_t) * cap, st));
  if (n_max <= 0) { SCN_CUDA(cudaMemsetAsync(n_unique_dev, 0, sizeof(int32_t), st)); return 0; }
  const size_t nm = (size_t)n_max;
  int32_t *slot_of_row = (int32_t *)scratch;
  int32_t *rank = slot_of_row + nm;
  int32_t *first_tmp = rank + nm;
  int32_t *block_sums = first_tmp + nm;
  int32_t *first = first_row ? first_row : first_tmp;
  if (count) SCN_CUDA(cudaMemsetAsync(count, 0, sizeof(int32_t) * nm, st));
  if (last_row) SCN_CUDA(cudaMemsetAsync(last_row, 0xFF, sizeof(int32_t) * nm, st));
  const unsigned nb = (unsigned)ceil_div(n_max, 256);
  hash_insert_kernel<<<nb, 256, 0, st>>>(keys, n_max, n_dev, (unsigned long long *)hkeys, hvals, cap, slot_of_row);
  SCN_CHECK_LAUNCH("hash_insert");
  FirstRowLoader ld{hvals, slot_of_row};
  FirstRowWriter wr{keys, ukeys, rank, first};
  if (scan_flags(ld, wr, n_max, n_dev, block_sums, n_unique_dev, st)) return 1;
  assign_ids_kernel<<<nb, 256, 0, st>>>(n_max, n_dev, hvals, slot_of_row, rank, id_of_row, count, last_row);
  finalize_table_kernel<<<nb, 256, 0, st>>>(n_max, n_unique_dev, first, slot_of_row, hvals);
  SCN_CHECK_LAUNCH("grid_build");
  count_launch(3);
  return 0;
}

int b200scn_coarse_keys(const uint64_t *ukeys, int64_t n_max, const int32_t *n_dev, int s,
                        uint64_t *ckeys, uint8_t *off, void *stream) {
  if (s < 2 || s > 4) return set_error("coarse_keys: stride %d unsupported (2..4)", s);
  if (n_max <= 0) return 0;
  coarse_keys_kernel<<<(unsigned)ceil_div(n_max, 256), 256, 0, (cudaStream_t)stream>>>(ukeys, n_max, n_dev, s, ckeys, off);
  SCN_CHECK_LAUNCH("coarse_keys");
  count_launch(1);
  return 0;
}

int b200scn_subm_map(const uint64_t *ukeys, int64_t n_max, const int32_t *n_dev,
                     const uint64_t *hkeys, const int32_t *hvals, int64_t cap, int64_t spatial_size,
                     int32_t *nbr, int32_t *counts27_dev, void *stream) {
  if (n_max <= 0) return 0;
  subm_map_kernel<<<(unsigned)ceil_div(n_max * 27, 256), 256, 0, (cudaStream_t)stream>>>(
      ukeys, n_max, n_dev, hkeys, hvals, cap, (int)spatial_size, nbr, counts27_dev);
  SCN_CHECK_LAUNCH("subm_map");
  count_launch(1);
  return 0;
}

int b200scn_child_map(const int32_t *parent, const uint8_t *off, int64_t nf_max,
                      const int32_t *nf_dev, int K, int32_t *child, int64_t nc_max, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (nc_max > 0) SCN_CUDA(cudaMemsetAsync(child, 0xFF, sizeof(int32_t) * (size_t)nc_max * K, st));
  if (nf_max <= 0) return 0;
  child_map_kernel<<<(unsigned)ceil_div(nf_max, 256), 256, 0, st>>>(parent, off, nf_max, nf_dev, K, child);
  SCN_CHECK_LAUNCH("child_map");
  count_launch(1);
  return 0;
}

size_t b200scn_pair_scratch_bytes(int64_t n, int K) {
  return sizeof(int32_t) * (scan_scratch_ints(n * K) + 64);
}

int b200scn_pair_lists(const int32_t *map, int64_t n, int K, int32_t *pair_in, int32_t *pair_out,
                       int32_t *offsets_dev, void *scratch, size_t scratch_bytes, void *stream) {
  return b200scn_pair_lists_ordered(map, nullptr, n, K, pair_in, pair_out, offsets_dev, scratch, scratch_bytes, stream);
}

int b200scn_pair_lists_ordered(const int32_t *map, const int32_t *order, int64_t n, int K, int32_t *pair_in,
                               int32_t *pair_out, int32_t *offsets_dev, void *scratch, size_t scratch_bytes,
                               void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (scratch_bytes < b200scn_pair_scratch_bytes(n, K)) return set_error("pair_lists: scratch too small");
  if (n * K >= (int64_t)1 << 31) return set_error("pair_lists: n*K overflows int32");
  if (n <= 0) { SCN_CUDA(cudaMemsetAsync(offsets_dev, 0, sizeof(int32_t) * (K + 1), st)); return 0; }
  PairLoader ld{map, order, n, K};
  PairWriter wr{map, order, n, K, pair_in, pair_out, offsets_dev, nullptr, 1, 0};
  return scan_flags(ld, wr, n * K, nullptr, (int32_t *)scratch, offsets_dev + K, st);
}

int b200scn_pair_lists_blocked(const int32_t *map, const int32_t *order, int64_t n, int K, int row_block,
                               int32_t *pair_in, int32_t *pair_out, int32_t *offsets_dev, int32_t *blk_offsets,
                               void *scratch, size_t scratch_bytes, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (row_block < 32 || (row_block & (row_block - 1))) return set_error("pair_lists_blocked: row_block %d must be a power of two >= 32", row_block);
  if (scratch_bytes < b200scn_pair_scratch_bytes(n, K)) return set_error("pair_lists: scratch too small");
  if (n * K >= (int64_t)1 << 31) return set_error("pair_lists: n*K overflows int32");
  const int nblk = (int)ceil_div(n > 0 ? n : 1, (int64_t)row_block);
  if (n <= 0) {
    SCN_CUDA(cudaMemsetAsync(offsets_dev, 0, sizeof(int32_t) * (K + 1), st));
    SCN_CUDA(cudaMemsetAsync(blk_offsets, 0, sizeof(int32_t) * ((size_t)K * nblk + 1), st));
    return 0;
  }
  PairLoader ld{map, order, n, K};
  PairWriter wr{map, order, n, K, pair_in, pair_out, offsets_dev, blk_offsets, row_block, nblk};
  if (scan_flags(ld, wr, n * K, nullptr, (int32_t *)scratch, offsets_dev + K, st)) return 1;
  // closing entry = total number of pairs (segment i of the flat table ends where segment i + 1 starts)
  SCN_CUDA(cudaMemcpyAsync(blk_offsets + (size_t)K * nblk, offsets_dev + K, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  return 0;
}

}  // extern "C"
