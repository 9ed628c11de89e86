// Microbenchmark: cost of back-to-back tcgen05.mma kind::tf32 (M = 128, K = 8) issued by one elected lane, as a function
// of N, operand source of A (shared memory descriptor vs tensor memory) and accumulator rotation.  B200, one CTA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_issue_bench mma_issue_bench.cu && ./mma_issue_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../3d-weakly-supervised-semantic-segmentation_b200/csrc/tc_common.cuh"
using namespace b200scn::tc;

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts_plain(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(ok));
  return ok != 0;
}
__device__ __forceinline__ void commit_elect(uint64_t *bar) {
  asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
               "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}

// mode 0: A in smem, 1: A in TMEM; nacc accumulators used round-robin; reps MMAs
__global__ void __launch_bounds__(128) bench(int M, int N, int mode, int variant, int reps, long long *out) {
  const int nacc = 1;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<float *>(smem_raw + (base - smem_u32(smem_raw)))[i] = 1.0f;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1) {
    const uint32_t idesc = make_idesc_tf32(M, N, 0, 0);
    const uint64_t adesc = make_smem_desc(base, 16, 1024), bdesc = make_smem_desc(base + 16384, 16, 1024);
    const int stride = N;   // accumulator a at columns a * N (nacc * N + 32 <= 512)
    long long t0 = clock64();
    int acc = 0;
    if (variant == 0) {
      for (int i = 0; i < reps; ++i) {
        const uint32_t d = tmem + acc * stride;
        if (mode == 0) mma_ss(d, adesc + 2 * (i & 3), bdesc + 2 * (i & 3), idesc, i >= nacc);
        else mma_ts(d, tmem + 480 + 8 * (i & 3), bdesc + 2 * (i & 3), idesc, i >= nacc);
        if (++acc == nacc) acc = 0;
      }
    } else if (variant == 1) {
      if (elect_one()) {
        for (int i = 0; i < reps; ++i) {
          if (mode == 0) mma_tf32(tmem, adesc + 2 * (i & 3), bdesc + 2 * (i & 3), idesc, i >= 1);
          else mma_ts_plain(tmem, tmem + 480 + 8 * (i & 3), bdesc + 2 * (i & 3), idesc, i >= 1);
        }
      }
      __syncwarp();
    } else {
      if (elect_one()) {
        const uint64_t b0 = bdesc, b1 = bdesc + 2, b2 = bdesc + 4, b3 = bdesc + 6;
        const uint32_t a0 = tmem + 480, a1 = tmem + 488, a2 = tmem + 496, a3 = tmem + 504;
        mma_ts_plain(tmem, a0, b0, idesc, 0);
        for (int i = 0; i < reps / 4; ++i) {
          mma_ts_plain(tmem, a1, b1, idesc, 1);
          mma_ts_plain(tmem, a2, b2, idesc, 1);
          mma_ts_plain(tmem, a3, b3, idesc, 1);
          mma_ts_plain(tmem, a0, b0, idesc, 1);
        }
      }
      __syncwarp();
    }
    long long t1 = clock64();
    commit_elect(&bar);
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
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int reps = 512;
  for (int M : {128})
  for (int mode = 0; mode < 2; ++mode)
    for (int N : {32, 64, 128, 256})
      for (int nacc : {0, 1, 2}) {
        if (mode == 0 && nacc == 2) continue;
        for (int w = 0; w < 2; ++w) bench<<<1, 128, 64 * 1024>>>(M, N, mode, nacc, reps, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("M=%3d A in %s  N=%3d variant=%d : issue %.1f cyc/mma, complete %.1f cyc/mma\n", M, mode ? "TMEM" : "smem", N, nacc,
               (double)h[0] / reps, (double)h[1] / reps);
      }
  return 0;
}
