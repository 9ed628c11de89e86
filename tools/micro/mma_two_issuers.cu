// Can two threads of one CTA issue tcgen05.mma (accumulate) into the SAME TMEM accumulator concurrently without losing
// updates?  A = 1 (TMEM), B = 1 (shared memory), every MMA (M = 128, N, K = 8, kind::tf32) adds exactly 8 to each element.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_two_issuers mma_two_issuers.cu && ./mma_two_issuers
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../3d-weakly-supervised-semantic-segmentation_b200/csrc/tc_common.cuh"
using namespace b200scn::tc;

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void st16(uint32_t taddr, float v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
               ::"r"(taddr), "f"(v) : "memory");
}

// warps 0-3: TMEM lane quarters (init + check); warps 4, 5: issuers
__global__ void __launch_bounds__(192) k(int N, int reps, int issuers, int *bad, long long *cyc) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar[2];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float *sm = reinterpret_cast<float *>(smem_raw + (base - smem_u32(smem_raw)));
  for (int i = threadIdx.x; i < 32768 / 4; i += 192) sm[i] = 1.0f;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp < 4) {
    const uint32_t t = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < N; c += 16) st16(t + c, 0.0f);      // accumulator = 0
    st16(t + 480, 1.0f);                                     // A operand (16 columns of ones, K = 8 uses 8)
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp >= 4 && warp - 4 < issuers) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_tf32(128, N, 0, 0);
      const uint64_t bdesc = make_smem_desc(base, 16, 1024);
      long long t0 = clock64();
      for (int i = 0; i < reps; ++i) mma_ts(tmem, tmem + 480, bdesc, idesc, 1u);
      mma_commit(&bar[warp - 4]);
      mbar_wait(&bar[warp - 4], 0);
      cyc[warp - 4] = clock64() - t0;
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp < 4) {
    const float want = 8.0f * reps * issuers;
    int nbad = 0;
    for (int c = 0; c < N; c += 16) {
      float v[16];
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
      for (int i = 0; i < 16; ++i) nbad += (v[i] != want);
    }
    if (nbad) atomicAdd(bad, nbad);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

int main() {
  int *bad; long long *cyc;
  cudaMalloc(&bad, 4); cudaMalloc(&cyc, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int issuers : {1, 2})
    for (int N : {32, 64, 128}) {
      int total_bad = 0; long long c[2] = {0, 0};
      for (int rep = 0; rep < 20; ++rep) {
        cudaMemset(bad, 0, 4);
        k<<<1, 192, 64 * 1024>>>(N, 512, issuers, bad, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        int h; cudaMemcpy(&h, bad, 4, cudaMemcpyDeviceToHost); total_bad += h;
        cudaMemcpy(c, cyc, 16, cudaMemcpyDeviceToHost);
      }
      printf("issuers %d N=%3d: wrong elements over 20 runs = %d ; cycles per MMA (issuer 0) %.1f\n", issuers, N, total_bad,
             (double)c[0] / 512);
    }
  return 0;
}
