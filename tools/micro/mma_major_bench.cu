// Microbenchmark: throughput of tcgen05.mma kind::tf32 (M = 128, K = 8) as a function of the operand MAJORNESS.
// The weight-gradient contraction runs over the rule (row) index, so gathered feature rows are "MN-major" operands
// (SWIZZLE_128B_BASE32B, the only tf32 MN-major layout); the forward contraction uses K-major operands (SWIZZLE_128B).
// 512 back-to-back MMAs into one accumulator, one elected lane, one CTA; reports cycles per MMA until completion.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_major_bench mma_major_bench.cu && ./mma_major_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../3d-weakly-supervised-semantic-segmentation_b200/csrc/tc_common.cuh"
using namespace b200scn::tc;

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// a_src: 0 smem K-major, 1 smem MN-major, 2 TMEM;  b_mn: 0 K-major, 1 MN-major
__global__ void __launch_bounds__(128) bench(int N, int a_src, int b_mn, int reps, long long *out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (32768 + 65536) / 4; i += 128) reinterpret_cast<float *>(smem_raw + (base - smem_u32(smem_raw)))[i] = 1.0f;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1) {
    const uint32_t idesc = make_idesc_tf32(128, N, a_src == 1, b_mn);
    // K-major: SBO = 1024 (8-row groups), layout 2.  MN-major: LBO = 4096 between 32-element MN blocks, SBO = 512, layout 1;
    // a K = 8 step advances 1024 bytes (8 rows of 128 bytes) in the MN-major image, 32 bytes in the K-major one
    const uint64_t adesc = a_src == 1 ? make_smem_desc(base, 4096, 512, 1) : make_smem_desc(base, 16, 1024);
    const uint64_t bdesc = b_mn ? make_smem_desc(base + 32768, 4096, 512, 1) : make_smem_desc(base + 32768, 16, 1024);
    const uint32_t astep = a_src == 1 ? 64 : 2, bstep = b_mn ? 64 : 2;
    long long t0 = clock64();
    if (elect_one()) {
      for (int i = 0; i < reps; ++i) {
        const uint64_t b = bdesc + bstep * (i & 3);
        if (a_src == 2) mma_ts(tmem, tmem + 480 + 8 * (i & 3), b, idesc, i >= 1);
        else mma_tf32(tmem, adesc + astep * (i & 3), b, idesc, i >= 1);
      }
      mma_commit(&bar);
    }
    __syncwarp();
    long long t1 = clock64();
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if ((threadIdx.x & 31) == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

int main() {
  long long *d, h[2];
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
  const int reps = 512;
  const char *an[3] = {"smem K-major ", "smem MN-major", "TMEM         "};
  for (int N : {32, 64, 128})
    for (int a_src = 0; a_src < 3; ++a_src)
      for (int b_mn = 0; b_mn < 2; ++b_mn) {
        for (int w = 0; w < 2; ++w) bench<<<1, 128, 128 * 1024>>>(N, a_src, b_mn, reps, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("M=128 N=%3d K=8  A %s  B %s : issue %.1f cyc/mma, complete %.1f cyc/mma\n", N, an[a_src],
               b_mn ? "MN-major" : "K-major ", (double)h[0] / reps, (double)h[1] / reps);
      }
  return 0;
}
