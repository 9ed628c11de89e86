// abi_conv.cu -- C-ABI dispatch for the convolution entry points (fp32 CUDA-core vs TF32 tcgen05 paths).
#include <stdlib.h>

#include "common.cuh"

namespace b200scn {
int gather_conv_simt(const float *A, int64_t lda, const int32_t *map, int64_t n_out, int K, const float *W,
                     int Cin, int Cout, const float *addend, int64_t ldadd, float *out, int64_t ldo,
                     cudaStream_t st);
int scatter_conv_simt(const float *A, int64_t lda, const int32_t *map, int64_t n_in, int K, const float *W,
                      int Cin, int Cout, float *out, int64_t ldo, cudaStream_t st);
int pair_dw_simt(const float *A, int64_t lda, const float *G, int64_t ldg, const int32_t *pair_a,
                 const int32_t *pair_g, const int32_t *offsets_dev, int K, int64_t n_pairs_max, int Ca, int Cg,
                 float *dW, cudaStream_t st);
bool gather_conv_tc_supported(const float *A, int64_t lda, int K, int Cin, int Cout, const float *W);
int gather_conv_tc(const float *A, int64_t lda, const int32_t *map, int64_t n_out, int K, const float *Wkm, int Cin,
                   int Cout, const float *addend, int64_t ldadd, float *out, int64_t ldo, cudaStream_t st);
int gather_conv_tma(const float *A, int64_t lda, int64_t n_in, const int32_t *map, int64_t n_out, int K,
                    const float *Wkm, int Cin, int Cout, const float *addend, int64_t ldadd, float *out, int64_t ldo,
                    cudaStream_t st);
int group_tiles(const int32_t *offsets_dev, int K, int64_t max_tiles, int32_t *tab, cudaStream_t st);
int grouped_conv_tc(const float *A, int64_t lda, const int32_t *in_rows, const int32_t *out_rows, const int32_t *tab,
                    int64_t max_tiles, int64_t n_out, int K, const float *Wkm, int Cin, int Cout, float *out, int64_t ldo,
                    cudaStream_t st);
bool pair_dw_tc_supported(const float *A, int64_t lda, const float *G, int64_t ldg, int Ca, int Cg);
int pair_dw_tc(const float *A, int64_t lda, const float *G, int64_t ldg, const int32_t *pair_a, const int32_t *pair_g,
               const int32_t *offsets_dev, int K, int64_t n_pairs_max, int Ca, int Cg, float *dW, cudaStream_t st,
               const int32_t *blk_tab = nullptr, int nblk = 0);
}  // namespace b200scn

using namespace b200scn;

extern "C" {

int b200scn_gather_conv(const float *A, int64_t lda, int64_t n_in, const int32_t *map, int64_t n_out, int K,
                        const float *W, int Cin, int Cout, const float *addend, int64_t ldadd,
                        float *out, int64_t ldo, int precision, void *stream) {
  if (n_out <= 0) return 0;
  if (!map && K != 1) return set_error("gather_conv: identity map requires K == 1");
  if (K < 1 || K > 64) return set_error("gather_conv: K=%d outside [1,64]", K);
  if (Cin < 1 || Cout < 1) return set_error("gather_conv: bad channel counts %d -> %d", Cin, Cout);
  if (n_out >= ((int64_t)1 << 31)) return set_error("gather_conv: too many rows");
  if (precision == 1) {
    if (!gather_conv_tc_supported(A, lda, K, Cin, Cout, W))
      return set_error("gather_conv: TF32 path needs Cin %% 8 == 0, Cout %% 16 == 0, Cout <= 1024, 16-byte aligned rows "
                       "(got %d -> %d, lda %lld)", Cin, Cout, (long long)lda);
    // option tc_tma=1 selects the TMA tile::gather4 producer variant.  Measured on B200 (profiles/r1_tma_gather4_shapes.txt):
    // 128-byte gather4 boxes run at ~1 row / 15 cycles / SM, 2.2x slower than per-lane cp.async, so it is not the default.
    if (g_opt.tc_tma == 1 && Cout <= 256)
      return gather_conv_tma(A, lda, n_in, map, n_out, K, W, Cin, Cout, addend, ldadd, out, ldo, (cudaStream_t)stream);
    return gather_conv_tc(A, lda, map, n_out, K, W, Cin, Cout, addend, ldadd, out, ldo, (cudaStream_t)stream);
  }
  return gather_conv_simt(A, lda, map, n_out, K, W, Cin, Cout, addend, ldadd, out, ldo, (cudaStream_t)stream);
}

int b200scn_scatter_conv(const float *A, int64_t lda, const int32_t *map, int64_t n_in, int K,
                         const float *W, int Cin, int Cout, float *out, int64_t ldo, int precision,
                         void *stream) {
  if (n_in <= 0) return 0;
  if (!map) return set_error("scatter_conv: map is required");
  if (K < 1 || K > 64) return set_error("scatter_conv: K=%d outside [1,64]", K);
  if (Cin < 1 || Cout < 1) return set_error("scatter_conv: bad channel counts %d -> %d", Cin, Cout);
  (void)precision;
  return scatter_conv_simt(A, lda, map, n_in, K, W, Cin, Cout, out, ldo, (cudaStream_t)stream);
}

int b200scn_group_tiles(const int32_t *offsets_dev, int K, int64_t max_tiles, int32_t *tab, void *stream) {
  return group_tiles(offsets_dev, K, max_tiles, tab, (cudaStream_t)stream);
}

int b200scn_grouped_conv(const float *A, int64_t lda, const int32_t *in_rows, const int32_t *out_rows,
                         const int32_t *tab, int64_t max_tiles, int64_t n_out, int K, const float *Wkm, int Cin,
                         int Cout, float *out, int64_t ldo, void *stream) {
  if (K < 1 || K > 64) return set_error("grouped_conv: K=%d outside [1,64]", K);
  if (!gather_conv_tc_supported(A, lda, K, Cin, Cout, Wkm))
    return set_error("grouped_conv: needs Cin %% 8 == 0, Cout %% 16 == 0, 16-byte aligned rows (got %d -> %d)", Cin, Cout);
  if (reinterpret_cast<uintptr_t>(tab) & 15) return set_error("grouped_conv: tile table must be 16-byte aligned");
  return grouped_conv_tc(A, lda, in_rows, out_rows, tab, max_tiles, n_out, K, Wkm, Cin, Cout, out, ldo,
                         (cudaStream_t)stream);
}

int b200scn_pair_dw(const float *A, int64_t lda, const float *G, int64_t ldg, const int32_t *pair_a,
                    const int32_t *pair_g, const int32_t *offsets_dev, int K, int64_t n_pairs_max,
                    int Ca, int Cg, float *dW, int precision, void *stream) {
  if (K < 1 || K > 64) return set_error("pair_dw: K=%d outside [1,64]", K);
  if (!offsets_dev && K != 1) return set_error("pair_dw: a single list requires K == 1");
  if (precision == 1 && pair_dw_tc_supported(A, lda, G, ldg, Ca, Cg))
    return pair_dw_tc(A, lda, G, ldg, pair_a, pair_g, offsets_dev, K, n_pairs_max, Ca, Cg, dW, (cudaStream_t)stream);
  return pair_dw_simt(A, lda, G, ldg, pair_a, pair_g, offsets_dev, K, n_pairs_max, Ca, Cg, dW, (cudaStream_t)stream);
}

int b200scn_pair_dw_blocked(const float *A, int64_t lda, const float *G, int64_t ldg, const int32_t *pair_a,
                            const int32_t *pair_g, const int32_t *blk_offsets, int K, int nblk, int Ca, int Cg,
                            float *dW, void *stream) {
  if (K < 1 || K > 64 || nblk < 1 || !blk_offsets) return set_error("pair_dw_blocked: bad K=%d / nblk=%d / table", K, nblk);
  if (!pair_dw_tc_supported(A, lda, G, ldg, Ca, Cg))
    return set_error("pair_dw_blocked: shape %d x %d not taken by the tensor-core kernel (use b200scn_pair_dw)", Ca, Cg);
  return pair_dw_tc(A, lda, G, ldg, pair_a, pair_g, nullptr, K, 1, Ca, Cg, dW, (cudaStream_t)stream, blk_offsets, nblk);
}

}  // extern "C"

/* 1 if b200scn_gather_conv(precision = 1) accepts this shape */
extern "C" int b200scn_gather_conv_tf32_ok(int Cin, int Cout, int64_t lda) {
  return (Cin % 8 == 0) && (Cout % 16 == 0) && Cout >= 16 && Cout <= 1024 && (lda % 4 == 0);
}

namespace b200scn { Options g_opt; }

extern "C" int b200scn_set_option(const char *name, int value) {
  if (!name) return set_error("set_option: NULL name");
  if (!strcmp(name, "tc_tma")) g_opt.tc_tma = value;
  else if (!strcmp(name, "tc_msub")) g_opt.tc_msub = value;
  else if (!strcmp(name, "tc_nsplit")) g_opt.tc_nsplit = value;
  else if (!strcmp(name, "dw_chunk")) g_opt.dw_chunk = value >= 512 ? value : 512;
  else if (!strcmp(name, "dw_pairs")) g_opt.dw_pairs = value == 64 ? 64 : 32;
  else if (!strcmp(name, "dw_dbg")) g_opt.dw_dbg = value;
  else if (!strcmp(name, "halo_pf")) g_opt.halo_pf = value;
  else if (!strcmp(name, "halo_dbg")) g_opt.halo_dbg = value;
  else if (!strcmp(name, "halo_one_cta")) g_opt.halo_one_cta = value;
  else return set_error("set_option: unknown option '%s'", name);
  return 0;
}

extern "C" int b200scn_set_device(int device) {
  SCN_CUDA(cudaSetDevice(device));
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Weight preparation for the tensor-core kernels: one launch turns the parameter stack w0 (K, a, b) into the K-major
// operand (K, Cout_g, Cin_g) of one GEMM direction -- offsets optionally mirrored (k -> K-1-k), matrices optionally
// transposed -- and rounds every value to the nearest TF32 (cvt.rna; the tensor core itself would truncate the low 13
// mantissa bits, a bias of up to 2^-10 per weight).
namespace b200scn {
__global__ void prep_weight_kernel(const float *__restrict__ w0, int K, int a, int b, int transposed, int flip,
                                   float *__restrict__ out) {
  // transposed == 0: GEMM multiplies by w0[k] (Cin = a, Cout = b): out[k][co][ci] = w0[kk][ci][co]
  // transposed == 1: GEMM multiplies by w0[k]^T (Cin = b, Cout = a): out[k][co][ci] = w0[kk][co][ci]
  const int64_t n = (int64_t)K * a * b;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(e / ((int64_t)a * b));
    const int r = (int)(e - (int64_t)k * a * b);
    const int kk = flip ? K - 1 - k : k;
    float v;
    if (transposed) {
      v = __ldg(w0 + (int64_t)kk * a * b + r);
    } else {
      const int co = r / a, ci = r - co * a;   // out is (b, a) row-major
      v = __ldg(w0 + ((int64_t)kk * a + ci) * b + co);
    }
    uint32_t t;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v));
    out[e] = __uint_as_float(t);
  }
}
// Both GEMM directions of one layer in ONE launch (forward operand: w0[k]; backward-input operand: w0[k]^T, offsets
// mirrored when flip_bwd): out_fwd (K,b,a), out_bwd (K,a,b); each parameter element is read once and written twice.
__global__ void prep_weight_both_kernel(const float *__restrict__ w0, int K, int a, int b, int flip_bwd,
                                        float *__restrict__ out_fwd, float *__restrict__ out_bwd) {
  const int64_t n = (int64_t)K * a * b;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(e / ((int64_t)a * b));
    const int r = (int)(e - (int64_t)k * a * b);
    uint32_t t;
    // backward operand: coalesced read and write of w0[kk] as stored
    const int kk = flip_bwd ? K - 1 - k : k;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(__ldg(w0 + (int64_t)kk * a * b + r)));
    out_bwd[e] = __uint_as_float(t);
    // forward operand: out_fwd[k][co][ci] = w0[k][ci][co] (strided read of a matrix that sits in L1/L2)
    const int co = r / a, ci = r - co * a;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(__ldg(w0 + ((int64_t)k * a + ci) * b + co)));
    out_fwd[e] = __uint_as_float(t);
  }
}
}  // namespace b200scn

extern "C" int b200scn_prep_weight_tf32_both(const float *w0, int K, int a, int b, int flip_bwd, float *out_fwd,
                                             float *out_bwd, void *stream) {
  const int64_t n = (int64_t)K * a * b;
  if (n <= 0) return 0;
  const unsigned blocks = (unsigned)min((int64_t)kNumSMs * 8, ceil_div(n, 256));
  prep_weight_both_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w0, K, a, b, flip_bwd, out_fwd, out_bwd);
  SCN_CHECK_LAUNCH("prep_weight_tf32_both");
  count_launch(1);
  return 0;
}

// Every layer of a network in ONE launch: items[i] = {w0, out_fwd, out_bwd, K, a, b, flip_bwd, first element} (device table,
// ascending first element); element e of the concatenation is located by a binary search over the table.
namespace b200scn {
__global__ void prep_weight_batch_kernel(const b200scn_prep_item *__restrict__ items, int n_items, int64_t total) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int lo = 0, hi = n_items - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (__ldg(&items[mid].first) <= e) lo = mid; else hi = mid - 1;
    }
    const b200scn_prep_item it = items[lo];
    const int64_t le = e - it.first, ab = (int64_t)it.a * it.b;
    const int k = (int)(le / ab);
    const int r = (int)(le - (int64_t)k * ab);
    uint32_t t;
    const int kk = it.flip_bwd ? it.K - 1 - k : k;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(__ldg(it.w0 + (int64_t)kk * ab + r)));
    it.out_bwd[le] = __uint_as_float(t);
    const int co = r / it.a, ci = r - co * it.a;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(__ldg(it.w0 + ((int64_t)k * it.a + ci) * it.b + co)));
    it.out_fwd[le] = __uint_as_float(t);
  }
}
}  // namespace b200scn

extern "C" int b200scn_prep_weight_tf32_batch(const b200scn_prep_item *items_dev, int n_items, int64_t total, void *stream) {
  if (n_items <= 0 || total <= 0) return 0;
  const unsigned blocks = (unsigned)min((int64_t)kNumSMs * 16, ceil_div(total, 256));
  prep_weight_batch_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(items_dev, n_items, total);
  SCN_CHECK_LAUNCH("prep_weight_tf32_batch");
  count_launch(1);
  return 0;
}

extern "C" int b200scn_prep_weight_tf32(const float *w0, int K, int a, int b, int transposed, int flip, float *out,
                                        void *stream) {
  const int64_t n = (int64_t)K * a * b;
  if (n <= 0) return 0;
  const unsigned blocks = (unsigned)min((int64_t)kNumSMs * 8, ceil_div(n, 256));
  prep_weight_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w0, K, a, b, transposed, flip, out);
  SCN_CHECK_LAUNCH("prep_weight_tf32");
  count_launch(1);
  return 0;
}
