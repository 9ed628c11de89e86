// sort.cu -- Morton ordering of a level's sites inside the library (closes the boundary: the tile plan no longer needs a
// torch.sort between two C-ABI calls).  perm = site ids sorted by (sample, interleaved x/y/z): one LSD radix sort of
// (key, id) pairs with CUB's DeviceRadixSort over exactly the significant key bits.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace b200scn {

__device__ __forceinline__ uint64_t spread3_s(uint64_t v) {   // 16 bits -> every third bit
  v &= 0x1fffffull;
  v = (v | v << 32) & 0x1f00000000ffffull;
  v = (v | v << 16) & 0x1f0000ff0000ffull;
  v = (v | v << 8) & 0x100f00f00f00f00full;
  v = (v | v << 4) & 0x10c30c30c30c30c3ull;
  v = (v | v << 2) & 0x1249249249249249ull;
  return v;
}

__global__ void morton_pairs_kernel(const uint64_t *__restrict__ ukeys, int64_t n, int bshift, uint64_t *__restrict__ mk,
                                    int32_t *__restrict__ ids) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int x, y, z, b;
  split_key(ukeys[i], x, y, z, b);
  // sample index right above the interleaved coordinate bits: one sort over [0, bshift + batch_bits)
  mk[i] = ((uint64_t)b << bshift) | (spread3_s((uint64_t)x) << 2) | (spread3_s((uint64_t)y) << 1) | spread3_s((uint64_t)z);
  ids[i] = (int32_t)i;
}

static size_t sort_temp_bytes(int64_t n) {
  size_t bytes = 0;
  cub::DoubleBuffer<uint64_t> k(nullptr, nullptr);
  cub::DoubleBuffer<int32_t> v(nullptr, nullptr);
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, k, v, (int)n, 0, 64);
  return bytes;
}

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace b200scn

using namespace b200scn;

extern "C" {

size_t b200scn_morton_perm_scratch_bytes(int64_t n) {
  if (n <= 0) return 256;
  return 2 * align256(sizeof(uint64_t) * n) + align256(sizeof(int32_t) * n) + align256(sort_temp_bytes(n)) + 256;
}

/* perm[0..n) = site ids along the Morton curve (sample index major).  spatial_size bounds the coordinate bits that are
 * sorted (3 * ceil(log2(size)) low bits) and batch_bits the sample bits above bit 48. */
int b200scn_morton_perm(const uint64_t *ukeys, int64_t n, int64_t spatial_size, int batch_bits, int32_t *perm,
                        void *scratch, size_t scratch_bytes, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= 0) return 0;
  if (n >= ((int64_t)1 << 31)) return set_error("morton_perm: too many rows");
  if (scratch_bytes < b200scn_morton_perm_scratch_bytes(n)) return set_error("morton_perm: scratch too small");
  if (reinterpret_cast<uintptr_t>(scratch) & 255) return set_error("morton_perm: scratch must be 256-byte aligned");
  uint8_t *p = reinterpret_cast<uint8_t *>(scratch);
  uint64_t *k0 = reinterpret_cast<uint64_t *>(p); p += align256(sizeof(uint64_t) * n);
  uint64_t *k1 = reinterpret_cast<uint64_t *>(p); p += align256(sizeof(uint64_t) * n);
  int32_t *v1 = reinterpret_cast<int32_t *>(p); p += align256(sizeof(int32_t) * n);
  size_t temp = sort_temp_bytes(n);
  int cbits = 1;
  while (((int64_t)1 << cbits) < spatial_size) ++cbits;
  if (batch_bits < 0) batch_bits = 0;
  if (3 * cbits + batch_bits > 64) return set_error("morton_perm: %d coordinate + %d sample bits exceed 64", 3 * cbits, batch_bits);
  morton_pairs_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(ukeys, n, 3 * cbits, k0, perm);
  SCN_CHECK_LAUNCH("morton_pairs");
  cub::DoubleBuffer<uint64_t> keys(k0, k1);
  cub::DoubleBuffer<int32_t> vals(perm, v1);
  SCN_CUDA(cub::DeviceRadixSort::SortPairs(p, temp, keys, vals, (int)n, 0, 3 * cbits + batch_bits, st));
  if (vals.Current() != perm)
    SCN_CUDA(cudaMemcpyAsync(perm, vals.Current(), sizeof(int32_t) * n, cudaMemcpyDeviceToDevice, st));
  count_launch(2);
  return 0;
}

}  // extern "C"
