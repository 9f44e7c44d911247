/*
 * rqvae_b200.h — C ABI of the B200-native RQ-VAE semantic-ID encode path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / C++ types.  The
 * reference (CatchMan1/AI-education-generative-recommendation) is pure Python, so the
 * binding a maintainer adds is a ctypes stub (INTEGRATION.md shows it); the Python
 * package in this repo mirrors the reference's RQVAE / get_indices / infer() surface on
 * top of exactly these entry points.  Reference citations are paths relative to the
 * reference root, file:line.
 *
 * Conventions
 *   - Every function returns 0 on success and a negative RQB200_E* code on failure;
 *     rqb200_last_error() returns a thread-local message for the last failure.
 *   - "dev" pointers are CUDA device pointers on the model's device; `stream` is a
 *     cudaStream_t passed as void* (NULL = legacy default stream).  Nothing synchronises
 *     the device unless stated.  Inputs are borrowed, never copied silently.
 *   - Row-major contiguous arrays; fp32 = IEEE binary32; codes are int64 like the
 *     reference's LongTensor (vq.py:75, rq.py:54).
 *   - There is no CPU fallback: every compute entry point needs a CUDA device.
 */
#ifndef RQVAE_B200_H
#define RQVAE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define RQB200_ABI_VERSION 1

#define RQB200_OK            0
#define RQB200_EINVAL       -1   /* bad argument / unsupported shape            */
#define RQB200_ECUDA        -2   /* CUDA runtime error (message has the detail) */
#define RQB200_ENOMEM       -3
#define RQB200_ESTATE       -4   /* model not fully loaded                       */

#define RQB200_MAX_LAYERS    8
#define RQB200_MAX_LEVELS    8

typedef struct rqb200_model rqb200_model;   /* opaque: device-resident weights + workspaces */

int         rqb200_abi_version(void);
const char *rqb200_last_error(void);
/* Number of CUDA devices visible (0 ⇒ every compute call will fail with RQB200_ECUDA). */
int         rqb200_device_count(void);

/* ---- instrumentation (bench.py) ----------------------------------------------------
 * rqb200_launch_count: kernels this library has launched in this process so far.
 * rqb200_profile_enable(1) makes the library record cudaEvent pairs around its main kernels on the
 * launching stream; rqb200_profile_read waits for them and returns accumulated milliseconds and
 * launch counts per slot: 0 first Linear (exact), 1 other Linears (exact), 2 quantizer, 3 dedup
 * (sort + segmented rank), 4 tensor-core first Linear, 5 Sinkhorn regroup, 6 other tensor-core Linears,
 * 7 three-pass re-run tier of the fast route (all its kernels), 8 exact rescue tier (all its kernels; its
 * quantizer launch is also counted in slot 2).                                                          */
long long rqb200_launch_count(void);
int rqb200_profile_enable(int on);
int rqb200_profile_read(double *ms_out, long long *count_out, int nslots);
/* Diagnostics: CTA 0 of the tensor-core linear kernel records clock64() at pipeline events into
 * buf_dev[8][256] (NULL switches the trace off).  Used by tools/trace_tc.py.                      */
int rqb200_debug_tc_trace(long long *buf_dev);
/* Diagnostics: ablation switches of the tensor-core linear kernel (tools/ablate_tc.py); 0 = production. */
int rqb200_debug_tc_flags(int flags);
/* Diagnostics: one tensor-core Linear (+bias, optional ReLU) of the model in isolation; passes = 3 (split-fp16),
 * 1 (fp16 screening pass) or 2 (the TMA-fed TF32 kernel, first layer only).  Used by tools/ to time and ablate.   */
int rqb200_debug_linear_tc(rqb200_model *m, int which, int layer, const float *x_dev, int64_t n, float *y_dev,
                           int passes, int relu, void *stream);
/* Diagnostics: 1 = rqb200_sinkhorn_regroup uses the shared-memory kernels only (the register-resident kernels are the
 * default where K is 256 or 1024 and e_dim 32 or 64; both give the same codes, tests compare them); 0 = default.   */
int rqb200_debug_sinkhorn_variant(int variant);
/* Diagnostics: the Sinkhorn kernels divide many matrix entries by one row / column sum through the sum's correctly rounded
 * reciprocal and two fma corrections (same bits as `a / b`, csrc/sinkhorn.cu: div_by); this compares the two on `pairs`
 * pseudo-random operand pairs spread over ±exponent_span binades and returns the number of differing quotients.       */
int rqb200_debug_check_division(unsigned long long seed, int exponent_span, unsigned long long pairs,
                                unsigned long long *mismatches_host, void *stream);
/* Diagnostics: the whole tensor-core MLP at a chosen precision — passes = 3, 1, or 2 (TF32 first layer + three-pass
 * tail: the latent the screening tier certifies).  Used by tools/calibrate_gate.py.                               */
int rqb200_debug_mlp_tc(rqb200_model *m, int which, const float *x_dev, int64_t n, float *y_dev, int passes, void *stream);

/* ---- model lifetime ---------------------------------------------------------------
 * Replaces the module tree RQVAE.__init__ builds (rqvae.py:45-58): an encoder MLP
 * `dims[0] → … → dims[n_layers]` (layers.py:18-32; ReLU after all but the last Linear),
 * L codebooks of K[l] × e_dim (vq.py:21) and the mirrored decoder.                    */
int rqb200_model_create(rqb200_model **out, int device, int n_layers, const int *dims,
                        int n_levels, const int *K);
void rqb200_model_destroy(rqb200_model *m);

/* Upload one Linear of the encoder (which=0) or decoder (which=1).  W is [out,in] row-major
 * as in nn.Linear.weight (layers.py:23); W / b may be host or device pointers (UVA).
 * kblocks (host, nblk entries summing to `in`) is the K-blocking of the reference's CPU
 * GEMM for this shape (see DESIGN.md "summation order"); NULL ⇒ the default rule.       */
int rqb200_model_set_linear(rqb200_model *m, int which, int layer, const float *W, const float *b,
                            const int *kblocks, int nblk);
/* Upload codebook `level` ([K[level], e_dim], nn.Embedding.weight, vq.py:21). */
int rqb200_model_set_codebook(rqb200_model *m, int level, const float *E);
/* Margin-gate parameters of RQB200_ENCODE_FAST: the tensor-core latent z~ is assumed to satisfy
 * |z~ - z| <= gamma * (|z| + floor_abs) per row; rows whose top-2 distance gap could be closed by
 * such an error are re-run by the exact kernels.  Defaults: gamma = 2^-15, floor_abs = 1e-3.       */
int rqb200_model_set_gate(rqb200_model *m, float gamma, float floor_abs);
/* Screening tier of the fast route (opt-in).  enabled = 2: the wide first encoder layer runs as ONE TF32 tensor pass
 * fed by TMA straight from the fp32 rows (csrc/encode_tf32.cu), the narrow layers three-pass; enabled = 1: one fp16
 * pass through all layers; 0: off.  Rows are certified with the same margin gate at the looser bound gamma1 (default
 * 2^-11; 0 keeps the current value); only rows inside that gate are re-run with the three-pass split-fp16 kernels, and
 * only rows inside the tight gate of rqb200_model_set_gate go to the exact kernels.  rqb200_model_last_tier_rows: rows
 * re-run by the three-pass tier [0] and by the exact tier [1] in the last RQB200_ENCODE_FAST call.                  */
int rqb200_model_set_screen(rqb200_model *m, int enabled, float gamma1);
int rqb200_model_last_tier_rows(rqb200_model *m, int64_t *out2);
/* Shape accessors: number of quantizer levels L and the latent width e_dim of a handle (0 + last_error for NULL). */
int rqb200_model_levels(const rqb200_model *m);
int rqb200_model_e_dim(const rqb200_model *m);
/* Copy the current codebook of `level` back (host or device destination). */
int rqb200_model_get_codebook(rqb200_model *m, int level, float *E_out);

/* ---- encoder MLP / decoder MLP (layers.py:42-43) ------------------------------------
 * y[n, dims_out] = MLP(x[n, dims_in]) with the reference's exact fp32 summation order.
 * rows (device int64, may be NULL) gathers input rows x[rows[i]] for i < n.            */
int rqb200_mlp_exact(rqb200_model *m, int which, const float *x_dev, const int64_t *rows_dev,
                     int64_t n, float *y_dev, void *stream);

/* Same MLP on the tensor cores (tcgen05 split-fp16 GEMMs, fp32-class accuracy, NOT bit-exact):
 * the building block of RQB200_ENCODE_FAST, exposed for tests and for the decoder (1e-4 tolerance). */
int rqb200_mlp_tc(rqb200_model *m, int which, const float *x_dev, int64_t n, float *y_dev, void *stream);

/* ---- residual quantizer, use_sk=False (rq.py:39-56, vq.py:63-99) --------------------
 * z[n,e] → codes[n,L] int64; optional x_q[n,e] (Σ of straight-through outputs, rq.py:48),
 * optional sumsq[L] (fp64, Σ(q-r)² per level — numerator of the mse at vq.py:90-91; added
 * to whatever is already there), optional residual_out[n,e] (the residual entering the
 * last level, used by the collision re-encode loop, infer.py:112-130).
 * rows_out (may be NULL): codes row i is written to codes[rows_out[i]].                */
int rqb200_quantize(rqb200_model *m, const float *z_dev, int64_t n, int64_t *codes_dev,
                    const int64_t *rows_out_dev, float *xq_dev, double *sumsq_dev,
                    float *last_residual_dev, void *stream);

/* ---- RQVAE.get_indices(xs, use_sk=False) (rqvae.py:67-71) ---------------------------
 * mode: RQB200_ENCODE_EXACT — SIMT fp32 in the reference's summation order;
 *       RQB200_ENCODE_FAST  — tcgen05 split-fp16 tensor-core encoder + quantizer with a margin gate, rows
 *                             inside the gate re-run by the exact kernels (same codes).
 * z_out (may be NULL) receives the latent.  stats (host, may be NULL): [0]=rows rescued.
 * Synchronisation: both modes enqueue and return.  The number of rows the margin gate hands to the exact tier stays
 * on the device (the exact kernels run persistent grids over "row tiles below the count"), so the fast route can be
 * captured in a CUDA graph; only a non-NULL stats pointer makes the call wait for the stream, and
 * rqb200_model_last_tier_rows fetches the counts of the last call on demand.                                */
#define RQB200_ENCODE_EXACT 0
#define RQB200_ENCODE_FAST  1
int rqb200_get_indices(rqb200_model *m, int mode, const float *x_dev, int64_t n,
                       int64_t *codes_dev, float *z_out_dev, int64_t *stats_host, void *stream);

/* ---- RQVAE.forward(x, use_sk=False) + compute_loss pieces (rqvae.py:60-65,73-84) ----
 * out[n,in] (may be NULL), codes[n,L], sumsq[L] as above, recon_sum[2] (fp64):
 * [0] += Σ (out-x)², [1] += Σ |out-x|  (mse / l1 numerators, rqvae.py:75-78).          */
int rqb200_forward(rqb200_model *m, const float *x_dev, int64_t n, float *out_dev,
                   int64_t *codes_dev, double *sumsq_dev, double *recon_sum_dev, void *stream);

/* ---- Sinkhorn branch (vq.py:74-83, layers.py:85-108) --------------------------------
 * Last-level re-encode of collision groups (infer.py:109-130): groups are given CSR-style,
 * items[offsets[g] .. offsets[g+1]) are row numbers into residual[n,e] (the residual
 * entering the last level) and into codes[n,L]; for each group the [|g|,K] distance
 * matrix is centred over the whole group (vq.py:51-61), run through fp64 Sinkhorn and the
 * arg-max overwrites codes[item, L-1].  One CTA per group.                              */
int rqb200_sinkhorn_regroup(rqb200_model *m, const float *residual_dev, const int64_t *items_dev,
                            const int64_t *offsets_dev, int64_t n_groups, int max_group,
                            double epsilon, int iters, int64_t *codes_dev, void *stream);
/* The part of `model.get_indices(data[collision_items], use_sk=True)` (infer.py:120-122 ≡ generate_code.py:115-117)
 * that precedes the last level, for ALL collision groups of one round in one call: the reference runs its encoder and
 * its quantizer levels again on the rows of each group ALONE, and its CPU GEMM sums in another order for such small
 * batches (2 <= rows <= 15 and 24·rows <= inner dimension; csrc/small_batch.cu) — so every member is re-encoded in the
 * order of ITS group's size.  Writes codes[item, 0..L-2] and residual[item, :] (the residual entering the last
 * level; [N,e], indexed by item like codes) for the members only; rqb200_sinkhorn_regroup on that residual finishes
 * the round.  x_dev: the catalogue [N,in] (x_is_gathered = 0: row of member i is items[i]) or the members' rows in
 * list order [n_items,in] (x_is_gathered = 1).  With x_is_gathered = 0 the call synchronises once (it reads back how
 * many members belong to groups of fewer than 16 rows).                                                           */
int rqb200_reencode_groups(rqb200_model *m, const float *x_dev, int x_is_gathered, const int64_t *items_dev,
                           const int64_t *offsets_dev, int64_t n_groups, int64_t n_items, int64_t *codes_dev,
                           float *residual_dev, void *stream);
/* rqb200_reencode_groups with a memo.  What the call computes for a member depends on the item and on the size of its
 * group only through WHICH products take the small-batch order — a few size classes (rqb200_reencode_classes: one
 * threshold min(15, K/24) per product).  (item, class) → (codes of the first L-1 levels, residual) is a pure function,
 * so a member already re-encoded in a group of the same class (an earlier round of infer.py:112-130) is answered from
 * the memo instead of running the model again; new results are added.  Caller-owned device buffers, n_catalogue = rows
 * of x / codes / residual:  memo_have_dev uint32[n_catalogue] (zero before the first call),
 * memo_residual_dev float[classes][n_catalogue][e], memo_codes_dev int32[classes][n_catalogue][L-1].
 * Same outputs, bit for bit, as rqb200_reencode_groups(x_is_gathered = 0).  Synchronises once (miss counts).          */
int rqb200_reencode_classes(const rqb200_model *m);
int rqb200_reencode_groups_memo(rqb200_model *m, const float *x_dev, const int64_t *items_dev,
                                const int64_t *offsets_dev, int64_t n_groups, int64_t n_items, int64_t *codes_dev,
                                float *residual_dev, int64_t n_catalogue, unsigned *memo_have_dev,
                                float *memo_residual_dev, int *memo_codes_dev, void *stream);
/* The same for rows whose group mates live on other GPUs (sharded catalogue): the size of every row's group is given
 * (msize_dev, int32 [n_rows]) instead of the group lists; rows_dev = the rows' indices into x / codes / residual.   */
int rqb200_reencode_rows(rqb200_model *m, const float *x_dev, int x_is_gathered, const int64_t *rows_dev,
                         const int *msize_dev, int64_t n_rows, int64_t *codes_dev, float *residual_dev, void *stream);
/* Largest group (rows) rqb200_sinkhorn_regroup handles in shared memory for this model; larger
 * groups are skipped by it and must go through rqb200_distances + rqb200_sinkhorn_assign.   */
int rqb200_sinkhorn_group_cap(rqb200_model *m);
/* Groups with more members than rqb200_sinkhorn_group_cap: same re-encode with the fp64 matrix in a caller-provided
 * global scratch.  group_ids_dev[n_listed] = indices into offsets_dev, scratch_off_dev[n_listed] = offset (in doubles)
 * of each listed group's slice, which needs B*K + (B+1)/2 doubles for a group of B members.  One launch for all.   */
int rqb200_sinkhorn_regroup_large(rqb200_model *m, const float *residual_dev, const int64_t *items_dev,
                                  const int64_t *offsets_dev, const int64_t *group_ids_dev,
                                  const int64_t *scratch_off_dev, int64_t n_listed, double *scratch_dev,
                                  double epsilon, int iters, int64_t *codes_dev, void *stream);
/* sinkhorn_algorithm(distances, epsilon, iters) (layers.py:85-108): D[B,K] fp64 distances are
 * replaced in place by Q.                                                               */
int rqb200_sinkhorn(double *D_dev, int64_t B, int K, double epsilon, int iters, void *stream);
/* Full use_sk branch for one batch at one level (vq.py:77-83): d (fp32 [B,K]) → indices[B] int64.
 * scratch_dev: B*K doubles.                                                             */
int rqb200_sinkhorn_assign(const float *d_dev, int64_t B, int K, double epsilon, int iters,
                           double *scratch_dev, int64_t *idx_dev, void *stream);
/* The fp32 distance matrix of vq.py:71-73 for one level: d[n,K[level]].                 */
int rqb200_distances(rqb200_model *m, int level, const float *r_dev, int64_t n, float *d_dev,
                     void *stream);

/* ---- collisions and the suffix column (infer.py:18-42,139-177) ----------------------
 * Workspace is owned by the model and grown on demand.
 * rqb200_suffix_dedup: codes[n,L] → out[n,L+1] with out[i,L] = #{j<i : codes[j]==codes[i]}
 *   (infer.py:152-163), bit-exact.  n_distinct_host / max_group_host (may be NULL) give the
 *   statistics printed at infer.py:132-137; reading them synchronises the stream.
 * K_host (may be NULL): codes[:,l] < K_host[l]; NULL ⇒ the column ranges are scanned on the
 *   device first (one extra stream synchronisation).                                     */
int rqb200_suffix_dedup(rqb200_model *m, const int64_t *codes_dev, int64_t n, int L,
                        const int *K_host, int64_t *out_dev, int64_t *n_distinct_host,
                        int64_t *max_group_host, void *stream);
/* get_collision_item (infer.py:29-42) on device: writes the item lists of all groups with
 * >1 member, concatenated (ascending item index inside a group, groups ordered by key),
 * items_dev[n] and offsets_dev[n+1] capacity.  Returns counts through host pointers
 * (synchronises the stream).                                                            */
int rqb200_collision_groups(rqb200_model *m, const int64_t *codes_dev, int64_t n, int L,
                            const int *K_host, int64_t *items_dev, int64_t *offsets_dev,
                            int64_t *n_groups_host,
                            int64_t *n_items_host, int64_t *max_group_host, void *stream);
/* The same, for the rounds of the collision loop (infer.py:112-130): groups whose member set is exactly a group that
 * was re-encoded in the previous round are fixed points of the round (the re-encode is a pure function of the member
 * rows) and are left out of items/offsets; n_groups_total_host still counts them.  prev_first_dev / prev_meta_dev:
 * int64[n] records owned by the caller (any content before round 0), round = 0, 1, 2 … in call order.            */
int rqb200_collision_groups_changed(rqb200_model *m, const int64_t *codes_dev, int64_t n, int L, const int *K_host,
                                    int64_t *prev_first_dev, int64_t *prev_meta_dev, int round,
                                    int64_t *items_dev, int64_t *offsets_dev, int64_t *n_groups_host,
                                    int64_t *n_items_host, int64_t *max_group_host, int64_t *n_groups_total_host,
                                    void *stream);
/* Generic building blocks used by the multi-GPU dedup exchange: pack codes to u64 keys and
 * stable-sort (key, value) pairs by key; segmented rank = position inside the equal-key run. */
int rqb200_pack_keys(const int64_t *codes_dev, int64_t n, int L, const int *K_host,
                     uint64_t *keys_dev, void *stream);
int rqb200_sort_pairs(rqb200_model *m, uint64_t *keys_dev, int64_t *vals_dev, int64_t n,
                      int key_bits, void *stream);
int rqb200_segment_rank(rqb200_model *m, const uint64_t *sorted_keys_dev, int64_t n,
                        int64_t *rank_dev, void *stream);

/* ---- sharded suffix dedup over NVLink peer memory (new design; SURVEY.md §8e row "dedup") ------------
 * One process per GPU, contiguous item shards.  Equal codes have to meet to be ranked in ascending GLOBAL
 * item order, so every packed key travels to owner = hash(key) mod world — written by the partition kernel
 * itself straight into the owner's receive buffer (peer memory mapped with cudaIpc, NVLink underneath); the
 * owner ranks them with the same radix sort + segmented rank as rqb200_suffix_dedup and writes the ranks
 * back into the sources' return buffers.  Cross-GPU ordering uses system-scope release/acquire flags; there
 * is no NCCL call and no host wait in mid-step: the owner's receive count stays on the device (the sort's grids
 * are sized for max_recv_items); the call synchronises the stream once, at its end, to return the status.
 * Result: out[n_local, L+1], bit-identical to rqb200_suffix_dedup on the concatenated catalogue.
 *
 * Setup: every rank calls rqb200_shard_create, exchanges the rqb200_shard_handle_bytes()-byte handle of
 * rqb200_shard_get_handle with all ranks (any transport: the Python host uses torch.distributed
 * all_gather_object), then rqb200_shard_connect with the world*handle_bytes concatenation in rank order.
 * max_recv_items <= 0 selects 2*max_local_items + 65536; if an owner would receive more keys than that,
 * every rank returns RQB200_ENOMEM (nothing is written out of bounds).  All ranks must make the same
 * sequence of rqb200_shard_suffix_dedup calls (it is a collective).                                   */
typedef struct rqb200_shard rqb200_shard;
int  rqb200_shard_create(rqb200_shard **out, rqb200_model *m, int rank, int world,
                         int64_t max_local_items, int64_t max_recv_items);
int  rqb200_shard_handle_bytes(void);
int  rqb200_shard_get_handle(rqb200_shard *sh, void *handle_out);
int  rqb200_shard_connect(rqb200_shard *sh, const void *all_handles);
int  rqb200_shard_suffix_dedup(rqb200_shard *sh, const int64_t *codes_dev, int64_t n, int L,
                               const int *K_host, int64_t *out_dev, void *stream);
void rqb200_shard_destroy(rqb200_shard *sh);

/* ---- k-means codebook init (layers.py:69-82 via vq.py:40-49) ------------------------
 * Lloyd steps on device.  kmeans_assign gives each sample its nearest centre (the quantizer's
 * distance / first-index argmin); kmeans_accumulate adds per-cluster sums[K,e] (fp64),
 * counts[K] (int64) and the inertia — the statistics that are all-reduced over NCCL when the
 * samples are sharded; kmeans_update turns them into centres (empty clusters keep their
 * previous centre) and adds the squared centre shift to *shift_dev.                       */
int rqb200_kmeans_assign(const float *x_dev, int64_t n, int e, const float *centers_dev, int K,
                         float *cnorm_scratch_dev /*[K]*/, int64_t *assign_dev, void *stream);
/* d[n,K]: squared distances of every sample to K candidate centres (k-means++ seeding). */
int rqb200_kmeans_distances(const float *x_dev, int64_t n, int e, const float *centers_dev, int K,
                            float *cnorm_scratch_dev /*[K]*/, float *d_dev, void *stream);
int rqb200_kmeans_accumulate(const float *x_dev, int64_t n, int e, const int64_t *assign_dev,
                             const float *centers_dev, int K, double *sums_dev, int64_t *counts_dev,
                             double *inertia_dev, void *stream);
int rqb200_kmeans_update(float *centers_dev, int K, int e, const double *sums_dev,
                         const int64_t *counts_dev, double *shift_dev, void *stream);

/* ---- training step (reference RQ-VAE/train.py:97-124) ----------
 * The Python RQVAE mirror calls these from torch.autograd.Functions when gradients are enabled, so the
 * reference's own loop (model(data) → compute_loss → backward → clip → optimizer.step) runs unchanged.
 * All pointers are device pointers; nothing synchronises.                                  */
/* nn.Dropout in training mode (layers.py:21): y = x * keep / (1 - p), keep a pure function of (seed, index);
 * calling it again on a gradient with the same seed applies the same mask (backward).     */
int rqb200_dropout(const float *x_dev, int64_t count, float p, uint64_t seed, float *y_dev, void *stream);
/* Same with a second, device-resident seed word mixed in (may be NULL): a captured CUDA graph replays identical launch
 * arguments, so the per-step part of the seed has to live in device memory.                */
int rqb200_dropout_dev(const float *x_dev, int64_t count, float p, uint64_t seed, const uint64_t *seed_dev, float *y_dev,
                       void *stream);
/* nn.Linear (+ReLU) with caller-owned parameters, the reference's fp32 summation order (layers.py:23).      */
int rqb200_linear_forward(const float *x_dev, const float *W_dev, const float *b_dev, int64_t n, int in_dim,
                          int out_dim, int relu, float *y_dev, void *stream);
int64_t rqb200_linear_backward_scratch_floats(int64_t n, int in_dim, int out_dim);
/* Backward of y = relu?(x Wᵀ + b): dy_dev is masked in place by (y > 0) when relu, then dW[out,in] = dyᵀ x,
 * db[out] = Σ_rows dy, dx[n,in] = dy W (dx_dev may be NULL for the first layer).  Deterministic.             */
int rqb200_linear_backward(const float *x_dev, const float *W_dev, const float *y_dev, float *dy_dev, int64_t n,
                           int in_dim, int out_dim, int relu, float *dx_dev, float *dW_dev, float *db_dev,
                           float *scratch_dev, int64_t scratch_floats, void *stream);
/* One residual level once its codes are chosen (vq.py:87-95, rq.py:47-48): sumsq += Σ (E[idx] - r)²,
 * x_res = r + (E[idx] - r), r_next = r - x_res, x_q = (first ? 0 : x_q) + x_res.           */
int rqb200_rq_level_apply(const float *r_dev, const int64_t *idx_dev, const float *cb_dev, int64_t n, int e,
                          int first, float *xq_dev, float *r_next_dev, double *sumsq_dev, void *stream);
/* Gradient of the codebook loss mse(E[idx], r.detach()) (vq.py:91): dE[k] = coef * g[0] * Σ_{idx[i]=k} (E[k] - r[i]). */
int rqb200_vq_codebook_grad(const float *r_dev, const int64_t *idx_dev, const float *cb_dev, int64_t n, int e, int K,
                            float coef, const float *g_dev, float *dE_dev, void *stream);
/* Gradient reaching the latent through the straight-through estimator (vq.py:95) and the commitment loss of
 * level 0 (vq.py:90): dz = g_xq + coef * g[0] * (z - E0[idx0]); g_xq may be NULL.          */
int rqb200_rq_latent_grad(const float *z_dev, const int64_t *idx0_dev, const float *cb0_dev, const float *gxq_dev,
                          int64_t n, int e, float coef, const float *g_dev, float *dz_dev, void *stream);
/* compute_loss (rqvae.py:73-84): sums2[0] += Σ (out - x)², sums2[1] += Σ |out - x|; and its gradient
 * d_out = g[0] * 2 (out - x) / count (mse) or g[0] * sign(out - x) / count (l1).           */
int rqb200_recon_loss(const float *out_dev, const float *x_dev, int64_t count, double *sums2_dev, void *stream);
int rqb200_recon_grad(const float *out_dev, const float *x_dev, int64_t count, int l1, const float *g_dev,
                      float *d_out_dev, void *stream);
/* clip_grad_norm_(params, max_norm) + AdamW.step() (train.py:116-117) over every parameter in three launches.
 * chunks_dev[n_chunks][5] = {param, grad, exp_avg, exp_avg_sq (device addresses), element count}; gradients
 * are read as grad * grad_scale (data-parallel averaging); stats_dev[0] = total norm, [1] = clip coefficient;
 * step counts from 1; max_norm <= 0 disables clipping.                                     */
int rqb200_adamw_clip_step(const int64_t *chunks_dev, int n_chunks, double *partial_dev, float *stats_dev,
                           float grad_scale, float max_norm, float lr, float beta1, float beta2, float eps,
                           float weight_decay, int64_t step, void *stream);

/* Graph-replayable variant: the per-step scalars come from device memory, hyper_dev[6] = {1 - lr*wd, lr / (1 - beta1^t),
 * sqrt(1 - beta2^t), beta1, beta2, eps} — the host refreshes them before each replay.                 */
/* dst_dev[0..n) = values_host[0..n), n <= 8, as one tiny launch whose arguments carry the values (safe however far the host
 * runs ahead of the stream) — how the per-step scalars above reach the device.            */
int rqb200_set_floats(float *dst_dev, int n, const float *values_host, void *stream);
int rqb200_adamw_clip_step_dev(const int64_t *chunks_dev, int n_chunks, double *partial_dev, float *stats_dev,
                               float grad_scale, float max_norm, const float *hyper_dev, void *stream);

/* ---- consumers of the semantic ids -------------------------------
 * token = id + column * codebook_size + 1 for every column of the [n, n_cols] id array (item_to_offset_code,
 * RQVAE-T5/data_read.ipynb cell 2; ranges per position: check_data_alignment.py:105-121), int32 like the TIGER files. */
int rqb200_offset_tokens(const int64_t *ids_dev, int64_t n, int n_cols, int codebook_size, int32_t *tokens_dev,
                         void *stream);
/* out[j, :] = tokens[item_ids[j] - 1, :] for 1-indexed item ids (same cell: data[item_id - 1]); an id outside
 * [1, n] sets *bad_flag_dev to 1 and yields a zero row.                                     */
int rqb200_gather_item_tokens(const int32_t *tokens_dev, int64_t n, int n_cols, const int64_t *item_ids_dev,
                              int64_t count, int32_t *out_dev, int *bad_flag_dev, void *stream);

/* ---- synthetic catalogue (bench / tests) --------------------------------------------
 * Integer-hash generator of clustered BERT-like item embeddings; the same function exists in
 * numpy (package `synth.py`) so CPU and GPU produce identical bytes for any row range.    */
int rqb200_synth_items(uint64_t seed, int64_t first_row, int64_t n, int dim, int64_t n_total,
                       float *x_dev, void *stream);

/* ---- host-buffer end-to-end call (what infer.py:88-177 does, use_sk=False pass + suffix)
 * x_host[n,in] pinned or pageable host memory → codes_host[n,L+1] int64; chunks of
 * `chunk_rows` are copied H2D on a side stream while the previous chunk is encoded.
 * Kernels and the final D2H copy run on `stream` (NULL = the legacy default stream); the call synchronises that
 * stream before returning.  stats_host (may be NULL): [0] rows rescued by the exact
 * kernels, [1] distinct codes, [2] largest collision group (infer.py:132-137).           */
int rqb200_generate_codes_host(rqb200_model *m, int mode, const float *x_host, int64_t n,
                               int64_t chunk_rows, int64_t *codes_host, int64_t *stats_host, void *stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* RQVAE_B200_H */
