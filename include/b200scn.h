/*
 * b200scn.h -- C ABI of the B200-native sparse voxel convolution backbone.
 *
 * Drop-in boundary for the `sparseconvnet` ("scn") operator layer that the reference's encoders
 * compose (models/SparseConvNet.py:5,59-71,73-88,107-158).  Upstream scn binds a pybind module
 * `sparseconvnet.SCN` whose entry points take ATen tensors; every function below
 * replaces one of those, with plain device pointers, sizes in elements, leading dimensions in
 * elements and an opaque `cudaStream_t` passed as `void*`.  No torch/ATen type crosses this ABI.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *  - every call only enqueues work on `stream` and never synchronises;
 *  - return value 0 = ok, nonzero = error (text via b200scn_last_error());
 *  - `n_dev` arguments: optional device int32 holding the live row count when the host only knows an
 *    upper bound `n_max` (rulebook pyramids are built without host round trips); NULL means n_max is exact;
 *  - site key layout: b<<48 | x<<32 | y<<16 | z  (x,y,z < 65536, sample index b < 32768);
 *  - feature maps are fp32 row-major (rows = active sites, ids in first-occurrence order).
 */
#ifndef B200SCN_H
#define B200SCN_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char *b200scn_last_error(void);
int b200scn_version(void);
/* bind the calling thread of this library's CUDA runtime to `device` (one process per GPU) */
int b200scn_set_device(int device);
/* tuning / test knobs (process-wide; never read from the environment on the launch path): "tc_tma", "tc_msub",
 * "tc_nsplit", "dw_chunk", "halo_pf", "halo_one_cta" -- kernel variants the heuristics would not pick at test sizes */
int b200scn_set_option(const char *name, int value);
/* number of kernels this library has enqueued since load (bench.py's gpu_launches) */
unsigned long long b200scn_launch_count(void);

/* ------------------------------------------------------------------ grids / rulebooks (A1-A4) */
/* hash slots for n keys (power of two >= 2n) */
int64_t b200scn_hash_capacity(int64_t n);
/* scratch bytes b200scn_grid_build needs for n_max rows */
size_t b200scn_grid_scratch_bytes(int64_t n_max);

/* scn.InputLayer coordinate intake (models/SparseConvNet.py:61; dataset/data.py:186,198):
 * coords (P,ncols) int64 [x,y,z(,b)] -> keys[P]; *err_flag |= 1 if any coordinate is outside
 * [0,spatial_size) or the sample index is negative / too large. */
int b200scn_pack_coords(const int64_t *coords, int64_t P, int ncols, int64_t spatial_size,
                        uint64_t *keys, int32_t *err_flag, void *stream);

/* Active-site numbering (upstream Metadata::inputLayer / Convolution_InputSgsToRulesAndOutputSgs):
 * unique keys get ids in order of first occurrence.  hkeys/hvals: hash table of `cap` slots (the call
 * initialises it); id_of_row[n_max]; ukeys[n_max] keys in id order; first_row/last_row/count[n_max] are
 * optional per-site statistics (NULL to skip); n_unique_dev receives the number of sites. */
int b200scn_grid_build(const uint64_t *keys, int64_t n_max, const int32_t *n_dev, uint64_t *hkeys,
                       int32_t *hvals, int64_t cap, int32_t *id_of_row, uint64_t *ukeys,
                       int32_t *first_row, int32_t *last_row, int32_t *count,
                       int32_t *n_unique_dev, void *scratch, size_t scratch_bytes, void *stream);

/* Strided grid, filter size == stride == s (scn.Convolution(3,a,b,s,s), models/SparseConvNet.py:137):
 * ckeys[i] = key of site//s, off[i] = ((x%s)*s+(y%s))*s+(z%s). */
int b200scn_coarse_keys(const uint64_t *ukeys, int64_t n_max, const int32_t *n_dev, int s,
                        uint64_t *ckeys, uint8_t *off, void *stream);

/* Submanifold 3x3x3 neighbour table (upstream SubmanifoldConvolution_SgsToRules):
 * nbr[o*27+k] = id of site o+d_k (k = 9(dx+1)+3(dy+1)+(dz+1)) or -1; counts27_dev[k] += #present. */
int b200scn_subm_map(const uint64_t *ukeys, int64_t n_max, const int32_t *n_dev,
                     const uint64_t *hkeys, const int32_t *hvals, int64_t cap, int64_t spatial_size,
                     int32_t *nbr, int32_t *counts27_dev, void *stream);

/* child[j*K+off[i]] = i for every fine site i with parent[i] == j; other entries -1. */
int b200scn_child_map(const int32_t *parent, const uint8_t *off, int64_t nf_max,
                      const int32_t *nf_dev, int K, int32_t *child, int64_t nc_max, void *stream);

/* scn-form rulebook: compact map[n][K] (-1 = absent) into per-offset pair lists, ascending row:
 * for k: pairs p in [offsets[k], offsets[k+1]) : pair_in[p] = map[row][k], pair_out[p] = row.
 * offsets_dev has K+1 entries.  scratch: b200scn_pair_scratch_bytes(n,K). */
size_t b200scn_pair_scratch_bytes(int64_t n, int K);
int b200scn_pair_lists(const int32_t *map, int64_t n, int K, int32_t *pair_in, int32_t *pair_out,
                       int32_t *offsets_dev, void *scratch, size_t scratch_bytes, void *stream);
/* Same lists with the rows of every offset enumerated in the order order[0], order[1], ... (a permutation of 0..n-1)
 * instead of ascending: the same pairs, used by the weight gradient so that concurrently running CTAs touch rows that are
 * neighbours in space (order = the Morton permutation of b200scn_tile_plan).  order NULL = the canonical form above. */
int b200scn_pair_lists_ordered(const int32_t *map, const int32_t *order, int64_t n, int K, int32_t *pair_in,
                               int32_t *pair_out, int32_t *offsets_dev, void *scratch, size_t scratch_bytes,
                               void *stream);
/* Same lists plus a table of row-block boundaries for the weight-gradient kernel: blk_offsets[k * nblk + b] = position
 * in the lists at which the pairs of offset k whose row lies in order[b * row_block .. (b+1) * row_block) start
 * (nblk = ceil(n / row_block), row_block a power of two); blk_offsets has K * nblk + 1 entries, the last = total pairs,
 * so segment i ends where segment i + 1 starts.  With `order` = the Morton permutation, block b of EVERY offset covers
 * the same region of space: CTAs (k, b) that run at the same time share their gathered rows in L2. */
int b200scn_pair_lists_blocked(const int32_t *map, const int32_t *order, int64_t n, int K, int row_block,
                               int32_t *pair_in, int32_t *pair_out, int32_t *offsets_dev, int32_t *blk_offsets,
                               void *scratch, size_t scratch_bytes, void *stream);

/* ------------------------------------------------------------------ convolutions (A5-A7, A10) */
/* out[o,:] = sum_k A[map[o*K+k],:] . W[k]  (+ addend[o,:] if addend)      W: (K,Cin,Cout) row-major.
 * A has n_in rows (map values are in [0,n_in) or -1).
 * map NULL => K must be 1 and row o reads A[o] (NetworkInNetwork).
 * Replaces SubmanifoldConvolution_updateOutput / Convolution_updateOutput and, with the transposed
 * mirrored weights, SubmanifoldConvolution / Deconvolution backward-input.
 * precision: 0 = fp32 CUDA cores, W as above;
 *            1 = TF32 tensor cores (tcgen05.mma, fp32 accumulate in TMEM), W given K-major: (K,Cout,Cin) row-major;
 *                needs b200scn_gather_conv_tf32_ok(Cin, Cout, lda). */
int b200scn_gather_conv_tf32_ok(int Cin, int Cout, int64_t lda);
int b200scn_gather_conv(const float *A, int64_t lda, int64_t n_in, const int32_t *map, int64_t n_out, int K,
                        const float *W, int Cin, int Cout, const float *addend, int64_t ldadd,
                        float *out, int64_t ldo, int precision, void *stream);

/* ---- spatially tiled submanifold convolution (TF32 tcgen05; same result as b200scn_gather_conv(precision = 1) on the
 * 3x3x3 neighbour map, replaces SubmanifoldConvolution_updateOutput / backward-input).
 * Plan, once per level:  mkeys[i] = b<<48 | Morton(x,y,z) of ukeys[i]; the caller sorts them and passes the sorting
 * permutation `perm` (site ids in curve order); tile t = rows perm[128t .. 128t+127].  b200scn_tile_plan fills, per tile,
 *   halo_ids[t*hcap + s]  the distinct neighbour ids the tile references (first hcap of them), halo_n[t] their number,
 *   lmap[t*3456 + k*128 + r]  halo slot of nbr[perm[128t+r]*27+k]; 0xFFFF absent, 0xFFFE beyond hcap (fetched through nbr),
 *   kmask[t]  bit k set iff offset k occurs in the tile.
 * hcap: multiple of 8 in [8,512].  lmap must be 16-byte aligned and hold ceil(n/128)*3456 entries. */
int b200scn_morton_keys(const uint64_t *ukeys, int64_t n, uint64_t *mkeys, void *stream);
/* the whole ordering inside the library: perm[0..n) = site ids sorted by (sample, Morton code of x,y,z) -- key build + LSD
 * radix sort of (key, id) pairs over the significant bits only (3*ceil(log2(spatial_size)) coordinate bits, batch_bits sample
 * bits).  scratch: b200scn_morton_perm_scratch_bytes(n) bytes, 256-byte aligned. */
size_t b200scn_morton_perm_scratch_bytes(int64_t n);
int b200scn_morton_perm(const uint64_t *ukeys, int64_t n, int64_t spatial_size, int batch_bits, int32_t *perm,
                        void *scratch, size_t scratch_bytes, void *stream);
int b200scn_tile_plan(const int32_t *nbr, const int32_t *perm, int64_t n, int hcap, uint16_t *lmap,
                      int32_t *halo_ids, int32_t *halo_n, uint32_t *kmask, void *stream);
/* out[o,:] = sum_k A[nbr[o*27+k],:] . W[k] (+ addend[o,:]);  W K-major (27,Cout,Cin); shapes as for precision 1 above.
 * round_a != 0: rows of A are rounded to the nearest TF32 on their way into tensor memory (the tensor core would truncate);
 * 0 when the producer already wrote them rounded (b200scn_bn_forward with round_tf32). */
int b200scn_subm_conv_tiled(const float *A, int64_t lda, const int32_t *nbr, const int32_t *perm,
                            const uint16_t *lmap, const int32_t *halo_ids, const int32_t *halo_n,
                            const uint32_t *kmask, int hcap, int64_t n, const float *Wkm, int Cin, int Cout,
                            const float *addend, int64_t ldadd, float *out, int64_t ldo, int round_a, void *stream);

/* Tile-stationary weight gradient of the submanifold convolution (weight-gradient half of
 * SubmanifoldConvolution_backward) on the SAME plan as b200scn_subm_conv_tiled: dW (27,Ca,Cg) = sum over the rules of
 * A[in]^T (x) G[out].  Each tile's G rows and distinct A rows are staged in shared memory once; accumulators for all offsets
 * stay in tensor memory; per-CTA partials go to `scratch` and are summed in a fixed order (bit-reproducible, no atomics).
 * scratch_bytes() returns 0 for shapes it does not take (use b200scn_pair_dw then). */
size_t b200scn_subm_dw_tiled_scratch_bytes(int64_t n, int hcap, int Ca, int Cg);
int b200scn_subm_dw_tiled(const float *A, int64_t lda, const float *G, int64_t ldg, const int32_t *nbr,
                          const int32_t *perm, const uint16_t *lmap, const int32_t *halo_ids, const int32_t *halo_n,
                          int hcap, int64_t n, int Ca, int Cg, float *dW, float *scratch, size_t scratch_bytes,
                          void *stream);

/* K-major TF32 operand of one GEMM direction from the parameter stack w0 (K,a,b) in one launch: offsets mirrored if flip,
 * matrices transposed unless `transposed` (transposed = 0: multiply by w0[k], out (K,b,a); = 1: by w0[k]^T, out (K,a,b)),
 * values rounded to the nearest TF32 (the tensor core would truncate).  Input of every precision = 1 entry point. */
int b200scn_prep_weight_tf32(const float *w0, int K, int a, int b, int transposed, int flip, float *out, void *stream);
/* both directions of one layer in one launch: out_fwd (K,b,a) = operand of "multiply by w0[k]" (forward), out_bwd (K,a,b) =
 * operand of "multiply by w0[k]^T" with the offsets mirrored if flip_bwd (backward-input; flip for submanifold 3^3) */
int b200scn_prep_weight_tf32_both(const float *w0, int K, int a, int b, int flip_bwd, float *out_fwd, float *out_bwd,
                                  void *stream);
/* The same for MANY layers in one launch (a training step prepares every layer's operands once, after the optimiser has
 * changed the weights): a device-resident table, ascending `first` = index of the item's first element in the
 * concatenation of all items (K*a*b elements each); total = sum of all K*a*b. */
typedef struct b200scn_prep_item {
  const float *w0;
  float *out_fwd, *out_bwd;
  int32_t K, a, b, flip_bwd;
  int64_t first;
} b200scn_prep_item;
int b200scn_prep_weight_tf32_batch(const b200scn_prep_item *items_dev, int n_items, int64_t total, void *stream);

/* Offset-sorted ("grouped") strided convolution on the tensor cores, for the directions in which every output row has
 * exactly one rule (Deconvolution_updateOutput, backward-input of Convolution): the scn-form rulebook
 * (in_rows[p], out_rows[p], offsets[K+1], rules sorted by offset) is cut into runs of <= 128 rules of ONE offset
 * (b200scn_group_tiles: tab[4*t] = {offset, first rule, rules, 0}, max_tiles >= ceil(n/128) + K entries), and each run is
 * one dense [rules x Cin] . W[offset] tile: out[out_rows[p],:] = A[in_rows[p],:] . W[k(p)].  TF32 operand Wkm as for
 * b200scn_gather_conv(precision = 1). */
int b200scn_group_tiles(const int32_t *offsets_dev, int K, int64_t max_tiles, int32_t *tab, void *stream);
int b200scn_grouped_conv(const float *A, int64_t lda, const int32_t *in_rows, const int32_t *out_rows,
                         const int32_t *tab, int64_t max_tiles, int64_t n_out, int K, const float *Wkm, int Cin,
                         int Cout, float *out, int64_t ldo, void *stream);

/* out[map[j*K+k],:] = A[j,:] . W[k] for every present (j,k)   (Deconvolution_updateOutput,
 * Convolution backward-input).  Rows of `out` not addressed by map are left untouched. */
int b200scn_scatter_conv(const float *A, int64_t lda, const int32_t *map, int64_t n_in, int K,
                         const float *W, int Cin, int Cout, float *out, int64_t ldo, int precision,
                         void *stream);

/* dW[k] = sum over pairs p of list k of A[pair_a[p],:]^T (x) G[pair_g[p],:]   (K,Ca,Cg) row-major.
 * pair_a / pair_g NULL => identity (row p).  offsets_dev NULL => one list [0,n_pairs_max).
 * n_pairs_max bounds the length of any single list.  dW is overwritten.
 * precision 1: TF32 tcgen05 tiles when the shape allows (Cg % 16 == 0, channels <= 256), else the fp32 kernel. */
int b200scn_pair_dw(const float *A, int64_t lda, const float *G, int64_t ldg, const int32_t *pair_a,
                    const int32_t *pair_g, const int32_t *offsets_dev, int K, int64_t n_pairs_max,
                    int Ca, int Cg, float *dW, int precision, void *stream);
/* Same with one CTA per (offset k, row block b) of a b200scn_pair_lists_blocked table (tensor-core path only: returns an
 * error for shapes b200scn_pair_dw would send to the CUDA-core kernel -- call that one then). */
int b200scn_pair_dw_blocked(const float *A, int64_t lda, const float *G, int64_t ldg, const int32_t *pair_a,
                            const int32_t *pair_g, const int32_t *blk_offsets, int K, int nblk, int Ca, int Cg,
                            float *dW, void *stream);

/* out[i,:] = in[parent[i],:]  (UnPooling_updateOutput) and its transpose via the child map. */
int b200scn_unpool(const float *in, int64_t ldi, const int32_t *parent, int64_t n_fine, int C,
                   float *out, int64_t ldo, void *stream);
int b200scn_unpool_bwd(const float *d_out, int64_t ldd, const int32_t *child, int64_t n_coarse, int K,
                       int C, float *d_in, int64_t ldi, void *stream);

/* ------------------------------------------------------------------ data path on the GPU (f2) */
/* dataset/data.py:165-200 (trainMerge, form 0) / :266-290 (valMerge, form 1) ending in the packed keys b200scn_grid_build
 * consumes: per scene b (points scene_start[b]..scene_start[b+1], B+1 device ints) a = xyz . mats[b] (+ pre0 + pre[b] if pre)
 * in float64, offset from the scene's min/max and the host-drawn r1, r2 (B x 3 doubles each) exactly as the reference
 * computes it, rows outside [0, spatial_size)^3 dropped, coordinates truncated.  Outputs (device): keys / kept_rows of the
 * kept points in input order (capacity P), *n_kept_dev, kept_per_scene[B] (-> batch_offsets), offset_out[B x 3].
 * scratch: b200scn_augment_scratch_bytes(P, B). */
size_t b200scn_augment_scratch_bytes(int64_t P, int B);
int b200scn_augment_voxelize(const float *xyz, int64_t P, const int32_t *scene_start, int B, const double *mats,
                             double pre0, const double *pre, const double *r1, const double *r2, int form,
                             int64_t spatial_size, uint64_t *keys, int32_t *kept_rows, int32_t *n_kept_dev,
                             int32_t *kept_per_scene, double *offset_out, void *scratch, size_t scratch_bytes,
                             void *stream);
/* out[j,:] = src[rows[j],:] (+ add_per_scene[scene of rows[j],:]) for j < *n_dev (n_max if n_dev is NULL): features of the
 * kept points, with the per-scene colour jitter of data.py:200 folded in */
int b200scn_gather_rows(const float *src, int64_t lds, const int32_t *rows, const int32_t *n_dev, int64_t n_max, int C,
                        const float *add_per_scene, const int32_t *scene_start, int B, float *out, int64_t ldo,
                        void *stream);

/* ------------------------------------------------------------------ BatchNormReLU (A8) */
/* scratch: b200scn_bn_scratch_doubles(C) doubles, ZERO on entry and left zero on exit (self-cleaning: the reduction
 * kernel's last block finalises the statistics and clears the accumulators, so one persistent zero-initialised buffer per
 * stream serves every call; two launches per direction, no memset). */
size_t b200scn_bn_scratch_doubles(int C);
/* BatchNormalization_updateOutput: train!=0 => batch statistics over the n rows (biased var, sums taken about a pivot row),
 * running stats updated as r = momentum*r + (1-momentum)*batch (unbiased var); else running stats.
 * y = leaky_relu((x-mean)*invstd*weight+bias, leak).  Any C (column slices of 1024 inside).
 * round_tf32 != 0: y is written rounded to the nearest TF32 -- for outputs consumed ONLY by tcgen05 convolutions, whose
 * kind::tf32 operand fetch would otherwise truncate (operand rounding fused into the producer). */
int b200scn_bn_forward(const float *x, int64_t ldx, int64_t n, int C, const float *weight,
                       const float *bias, float *running_mean, float *running_var, float *save_mean,
                       float *save_invstd, float eps, float momentum, int train, float leak,
                       float *y, int64_t ldy, double *scratch, int round_tf32, void *stream);
/* BatchNormalization_backward; the ReLU mask is recomputed from x, which equals the sign of the forward output bit for
 * bit.  train != 0: batch-statistics formula; train == 0 (forward used the running statistics, which do not depend on x):
 * d_in = g * invstd * weight, d_weight / d_bias from the same sums.  addend (may be NULL): added to d_in in the same pass --
 * the gradient that reaches x through its other consumer (the skip of a residual block), instead of a separate add kernel. */
int b200scn_bn_backward(const float *x, int64_t ldx, const float *dy, int64_t lddy, int64_t n, int C,
                        const float *weight, const float *bias, const float *save_mean,
                        const float *save_invstd, float leak, int train, const float *addend, int64_t ldadd,
                        float *dx, int64_t lddx, float *d_weight, float *d_bias, double *scratch, void *stream);

/* ------------------------------------------------------------------ I/O layers (A1, A9) */
/* InputLayer_updateOutput feature half: out[pv[r]] += mult(r)*feats[r]; mode 1 last, 2 first, 3 sum,
 * 4 mean (Function_test.py:38-44).  out must be zeroed by the caller. */
int b200scn_input_features(const float *feats, int64_t P, int C, const int32_t *pv,
                           const int32_t *count, const int32_t *first_row, const int32_t *last_row,
                           int mode, float *out, void *stream);
int b200scn_input_features_bwd(const float *d_out, int64_t P, int C, const int32_t *pv,
                               const int32_t *count, const int32_t *first_row,
                               const int32_t *last_row, int mode, float *d_in, void *stream);
/* OutputLayer_updateOutput: out[r,:] = feats[pv[r],:] (no averaging); backward sums rows per site
 * (d_feats must be zeroed by the caller). */
int b200scn_output_features(const float *feats, int64_t ldf, int64_t P, int C, const int32_t *pv,
                            const int32_t *first_row, const int32_t *last_row, int mode, float *out,
                            void *stream);
int b200scn_output_features_bwd(const float *d_out, int64_t P, int C, const int32_t *pv,
                                const int32_t *first_row, const int32_t *last_row, int mode,
                                float *d_feats, int64_t ldf, void *stream);

/* Rows grouped by site (counting sort over pv): start[v] (exclusive scan of count), rows[start[v]..+count[v]).
 * With it the OutputLayer backward is a gather (each d_out row read once, no atomics). */
size_t b200scn_site_rows_scratch_bytes(int64_t n_sites);
int b200scn_site_rows(const int32_t *pv, int64_t P, const int32_t *count, int64_t n_sites, int32_t *start,
                      int32_t *rows, void *scratch, size_t scratch_bytes, void *stream);
int b200scn_output_features_bwd_csr(const float *d_out, int64_t n_sites, int C, const int32_t *start,
                                    const int32_t *count, const int32_t *rows, const int32_t *first_row,
                                    const int32_t *last_row, int mode, float *d_feats, int64_t ldf, void *stream);

/* Head pooling (A11: SparseConvBase_.postProcessing, models/SparseConvNet.py:20-26; MultiLabelContrastive.py:35-40):
 * out[b,:] = mean over the POINTS of scene b of the OutputLayer result, computed from the level-0 voxel features without
 * materialising the per-point tensor: (1/P_b) sum_v w(v) feats[v,:], w(v) = count[v] (modes 3,4) or 1 (modes 1,2).
 * ukeys/count: level-0 site keys (scene = key >> 48) and points per site; npts[B] receives P_b (kept for the backward). */
int b200scn_scene_mean(const float *feats, int64_t ldf, const uint64_t *ukeys, const int32_t *count, int mode,
                       int64_t n, int C, int B, float *out, float *npts, void *stream);
int b200scn_scene_mean_bwd(const float *g, const uint64_t *ukeys, const int32_t *count, int mode, const float *npts,
                           int64_t n, int C, float *d_feats, int64_t ldd, void *stream);

/* Scene-level head fused after the pooling (f1: nn.Linear(embed, NUM_CLASSES), models/MultiLabelContrastive.py:59,66;
 * F.multilabel_soft_margin_loss, utils/loss.py:21-30).  logits[B,NC] = pooled[B,C] W[NC,C]^T + bias; when labels[B,NC] is
 * given, *loss = mean_b mean_c -(y log sigmoid(x) + (1-y) log sigmoid(-x)), summed in a fixed order (bit-reproducible).
 * scratch: B + 1 floats, zeroed ONCE by the caller (the kernel leaves its completion counter at zero).
 * Backward: dl = d_logits (optional) + *d_loss (sigmoid(x) - y) / (B NC) (when labels and d_loss are given);
 * d_pooled = dl W, d_W = dl^T pooled, d_b = column sums of dl (d_b may be NULL). */
int b200scn_head_multilabel(const float *pooled, const float *W, const float *bias, const float *labels, int B, int C,
                            int NC, float *logits, float *loss, float *scratch, void *stream);
int b200scn_head_multilabel_bwd(const float *pooled, const float *W, const float *labels, const float *logits,
                                const float *d_loss, const float *d_logits, int B, int C, int NC, float *d_pooled,
                                float *d_W, float *d_b, void *stream);

/* scn.MaxPooling with size == stride (MaxPooling_updateOutput / _updateGradInput; models/projector/components.py:78-100):
 * out[j,:] = max(0, max over children of in[child,:]) -- upstream zero-initialises the output; the gradient goes to every
 * child that equals the pooled value. */
int b200scn_maxpool(const float *in, int64_t ldi, const int32_t *child, int64_t n_coarse, int K, int C, float *out,
                    int64_t ldo, void *stream);
int b200scn_maxpool_bwd(const float *g, int64_t ldg, const float *in, int64_t ldi, const float *out, int64_t ldo,
                        const int32_t *parent, int64_t n_fine, int C, float *d_in, int64_t ldd, void *stream);
/* scn.SparseToDense (SparseToDense_updateOutput / _updateGradInput; Function_test.py:46): dense is (B, C, S, S, S),
 * zero-filled by the caller; the backward gathers d_feats[v,:] from d_dense. */
int b200scn_sparse_to_dense(const float *feats, int64_t ldf, const uint64_t *ukeys, int64_t n, int C, int64_t spatial_size,
                            float *dense, void *stream);
int b200scn_sparse_to_dense_bwd(const float *d_dense, const uint64_t *ukeys, int64_t n, int C, int64_t spatial_size,
                                float *d_feats, int64_t ldf, void *stream);

/* ------------------------------------------------------------------ point2mask (A12) */
/* ops/point2mask/_ext_src/src/ball_query.cpp:8-33 (+ ball_query_gpu.cu:9-45). idx is fully written
 * (-1 sentinel included). */
int b200scn_p2m_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xy,
                           const float *xy, const int32_t *pointnums, int32_t *idx, void *stream);
/* Cell-bucketed ball query with IDENTICAL results (first nsample hits in ascending point index among the first n - ptnum
 * points, -1 sentinel): candidates are counting-sorted into cells of side `radius` over the queries' bounding box and each
 * query merges the index-sorted lists of the <= 4 x 4 cells covering [q - r, q + r]^2 -- O(points near the query) instead of
 * O(n) per query (production shape: 64 x 65 536 queries x 200 k points).  max_side >= (query extent / radius) + 3 cells per
 * axis; scratch: b200scn_p2m_ball_query_scratch_bytes(b, n, m, max_side). */
size_t b200scn_p2m_ball_query_scratch_bytes(int b, int n, int m, int max_side);
int b200scn_p2m_ball_query_bucketed(int b, int n, int m, float radius, int nsample, const float *new_xy,
                                    const float *xy, const int32_t *pointnums, int32_t *idx, int max_side,
                                    void *scratch, size_t scratch_bytes, void *stream);
/* ops/point2mask/_ext_src/src/group_points.cpp:12-36 / 38-62. out / grad_points fully written. */
int b200scn_p2m_group_points(int b, int c, int n, int npoints, int nsample, const float *points,
                             const int32_t *idx, float *out, void *stream);
int b200scn_p2m_group_points_grad(int b, int c, int n, int npoints, int nsample,
                                  const float *grad_out, const int32_t *idx, float *grad_points,
                                  void *stream);

#ifdef __cplusplus
}
#endif
#endif
