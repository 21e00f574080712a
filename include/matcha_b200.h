/*
 * matcha_b200 — C ABI of the B200-native Hyper-SAGNN hyperedge-scoring hot path.
 *
 * This is the drop-in boundary.  The reference (ma-compbio/MATCHA) is pure Python/PyTorch and has no
 * FFI of its own; each entry point below names the reference function whose device work it replaces
 * (paths relative to the reference's Code/ directory).  A maintainer binds it with ctypes (see
 * INTEGRATION.md); our own host mirror lives in matcha_b200/Modules.py.
 *
 * Conventions
 *   - every pointer marked "dev" is a CUDA device pointer allocated by the CALLER; the library never
 *     allocates or frees device memory and keeps no global state besides lazily set function attributes;
 *   - all work is enqueued on the `stream` argument (a cudaStream_t passed as void*) and is
 *     stream-ordered; nothing synchronises the device;
 *   - matrices are row-major fp32 unless stated; node ids are int64, 0 = padding, 1..N real bins;
 *   - return value 0 = success, negative = error (message via matcha_last_error()); nothing throws
 *     or calls exit().
 */
#ifndef MATCHA_B200_H
#define MATCHA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MATCHA_OK 0
#define MATCHA_ERR_ARG (-1)
#define MATCHA_ERR_CUDA (-2)
#define MATCHA_ERR_UNSUPPORTED (-3)

#define MATCHA_MAX_CHROM 64
#define MATCHA_MAX_WIDTH 8 /* hyperedge width L (reference uses 2..5, config.JSON:14) */

const char* matcha_last_error(void);
int matcha_version(void);

/* Optional per-call-site timing with CUDA events on the launching stream (bench.py's roofline figures).
 * matcha_profile_read synchronises on the recorded events, fills ms / calls / kernel-launch counts per
 * label (n >= matcha_profile_labels()) and resets the counters. */
void matcha_profile_enable(int32_t on);
int32_t matcha_profile_labels(void);
const char* matcha_profile_label_name(int32_t i);
int matcha_profile_read(float* ms, int64_t* calls, int64_t* kernels, int32_t n);
/* 0 = SIMT fp32 contractions only, 1 = tcgen05 (bf16x3) where the shape is eligible */
void matcha_set_gemm_impl(int32_t impl);
/* 1 (default) = fused hyperedge-tile kernels (QKG projection + attention in one tcgen05 kernel) where eligible,
 * 0 = decomposed pipeline (projection, attention, ... as separate launches); also MATCHA_FUSED=0 */
void matcha_set_fused(int32_t on);
/* 1 (default) = the 64-wide layers around the attention block run as tcgen05 row-chain kernels (needs the fused
 * path), 0 = SIMT fp32 contractions; also MATCHA_CHAIN=0 */
void matcha_set_chain(int32_t on);
/* 1 (default) = the reconstruction head (Modules.py:192-199) and its backward run as one fused tcgen05 kernel per pass
 * (needs the fused path), 0 = four SIMT launches through a [T, n_r] buffer; also MATCHA_RECON_TC=0 */
void matcha_set_recon_tc(int32_t on);
void matcha_set_recon_pipe(int32_t on);   /* pipelined gradient pass of the reconstruction head (default on) */
/* pipelined node-encoder kernels (csrc/enc_tc.cu): forward 0 = unit kernel, 1 = pipelined, 2 = by feature-row width (default:
 * pipelined from 512 bins per chromosome on average); backward 0 = unit, 1 = pipelined (default); also MATCHA_ENC_PIPE,
 * MATCHA_ENC_PIPE_BWD */
void matcha_set_enc_pipe(int32_t fwd, int32_t bwd);
/* 1 (default) = embed_dim-128 models run their contractions on the general tcgen05 kernel (csrc/gemm_tcg.cu, bf16x3) from
 * 1024 token rows on, 0 = fp32 SIMT kernel; also MATCHA_GEMM_TCG=0 */
void matcha_set_gemm_tcg(int32_t on);
/* 1 (default) = both node-encoder layers (Modules.py:104-122) run as one tcgen05 kernel over the chromosome-bucketed
 * token list (dense feature rows, embed_dim 64, >= 1024 tokens), 0 = two grouped SIMT launches; also MATCHA_ENC_TC=0 */
void matcha_set_enc_tc(int32_t on);
/* 1 (default) = the fused attention FORWARD runs in "X-form" for widths L <= 5 (csrc/attn_xform.cu: scores and value mixing
 * re-associated onto the neighbours' input rows, held in registers: no per-head warp shuffles), 0 = the Q/K/G shuffle form
 * of csrc/attn_fused.cu; also MATCHA_XFORM=0 */
void matcha_set_xform(int32_t on);
/* bf16 products per tcgen05 contraction step of the attention block: 3 (default) = hi*hi + hi*lo + lo*hi of the bf16 hi|lo
 * operand split (fp32-accurate: ~2^-16 relative), 1 = hi*hi only ("bf16 mode", BASELINE's stated-tolerance option:
 * measured logit / gradient error in DESIGN.md); also MATCHA_BF16=1 */
void matcha_set_mma_passes(int32_t passes);

/* ---------------------------------------------------------------------------------------------
 * Model description: where every live tensor of Modules.Classifier sits.
 * `params` / `grads` are flat fp32 device buffers with IDENTICAL layout; off_* are element offsets.
 * --------------------------------------------------------------------------------------------- */
typedef struct matcha_model_desc {
  int32_t d;        /* embed_dim (main.py:517); 64 supported */
  int32_t n_head;   /* 8 (main.py:616) */
  int32_t n_chrom;  /* C */
  int32_t attr_dim; /* C + 1 (main.py:497-512) */
  int64_t n_nodes;  /* N */
  float* params;    /* dev */
  float* grads;     /* dev (may be NULL for inference-only use) */

  /* Classifier.attribute_nn, Classifier.next_w (Modules.py:242-249) */
  int64_t off_attr_w, off_attr_b, off_next_w, off_next_b;
  /* encode1.mul_head_attn: layer_norm1/2/3, w_qs/w_ks/w_vs, fc1 (Modules.py:481-500) */
  int64_t off_lnq_g, off_lnq_b, off_lnk_g, off_lnk_b, off_lnv_g, off_lnv_b;
  int64_t off_wq, off_wk, off_wv, off_fc1_w, off_fc1_b;
  /* encode1.pff_n1: PWF_Conv0/1, layer_norm (Modules.py:604-605) */
  int64_t off_pff_w0, off_pff_b0, off_pff_w1, off_pff_b1, off_pff_g, off_pff_b;
  /* Classifier.layer_norm1/2, pff_classifier.PWF_Conv0 (Modules.py:218,240-241) */
  int64_t off_ln1_g, off_ln1_b, off_ln2_g, off_ln2_b, off_cls_w, off_cls_b;

  /* per chromosome c (Modules.py:163-171): ids [chrom_start[c], chrom_end[c]) , n_c = end - start */
  int64_t chrom_start[MATCHA_MAX_CHROM];
  int64_t chrom_end[MATCHA_MAX_CHROM];
  int64_t off_w0[MATCHA_MAX_CHROM];  /* 'tied weight_0' [d, n_c]  */
  int64_t off_w1[MATCHA_MAX_CHROM];  /* 'tied weight_1' [d, d]    */
  int64_t off_rw[MATCHA_MAX_CHROM];  /* Embedding_recon{c}.FF_Linear0.weight [n_c, d] */
  int64_t off_rb[MATCHA_MAX_CHROM];  /* Embedding_recon{c}.FF_Linear0.bias   [n_c]    */
  const float* feat[MATCHA_MAX_CHROM]; /* dev: SparseEmbedding.embedding, dense [n_c, feat_ld[c]] (Modules.py:48-52) */
  int64_t feat_ld[MATCHA_MAX_CHROM];
  /* optional CSR form of the same feature rows (Modules.py:58-65, sparse=True); used when feat[c]==NULL */
  const int64_t* feat_indptr[MATCHA_MAX_CHROM];  /* dev [n_c + 1] */
  const int32_t* feat_indices[MATCHA_MAX_CHROM]; /* dev */
  const float* feat_values[MATCHA_MAX_CHROM];    /* dev */

  const float* attr_table; /* dev [N+1, attr_dim]: attribute_dict_embedding.weight (frozen, Modules.py:245-247) */
  const float* inter;      /* dev [N, inter_ld] z-scored inter matrix (Modules.py:147-154) or NULL */
  int64_t inter_ld;

  /* effective ("derived") weights written by matcha_prepare: dev buffers of matcha_derived_elems() floats */
  float* derived;
  float* derived_grad;

  float p_feature, p_attn, p_pff; /* dropout probabilities (Modules.py:174,226-227) */
} matcha_model_desc;

int64_t matcha_derived_elems(const matcha_model_desc* m);
/* bytes of scratch needed for a batch of B hyperedges of padded width L (training=1 keeps activations) */
int64_t matcha_workspace_bytes(const matcha_model_desc* m, int64_t B, int32_t L, int32_t training);

/* Fold LayerNorm affine / temperature / fc1 into the effective projection weights.  Call after every
 * parameter update and before forward.  Replaces nothing in the reference (a B200-side re-association
 * of Modules.py:519-529,572); exact in real arithmetic.
 * Stream contract: it may be queued on a DIFFERENT stream than the pass that follows (it reads the weights only, so a
 * trainer can run it beside batch assembly).  matcha_forward / matcha_node_embeddings / matcha_pair_tables wait, on
 * their own stream and after their weight-free token bucketing, for the most recent matcha_prepare queued on this
 * device.  The caller orders matcha_prepare after the weight update and after the previous backward pass. */
int matcha_prepare(const matcha_model_desc* m, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Classifier.forward(x, return_recon=True)  — Modules.py:278-318 (and everything it calls).
 *   x          dev int64 [B, L]
 *   logits     dev fp32  [B]      raw logits (callers apply sigmoid: main.py:58)
 *   recon      dev fp32  [1]      reconstruction loss for chromosome `random_chrom` (Modules.py:192-199);
 *                                 random_chrom < 0 skips it and writes 0
 *   training   0 = eval (no dropout), 1 = train (dropout from the counter RNG keyed by `seed`)
 *   workspace  dev, matcha_workspace_bytes() bytes; after a training forward it holds the tape that
 *              matcha_backward consumes
 * --------------------------------------------------------------------------------------------- */
int matcha_forward(const matcha_model_desc* m, const int64_t* x, int64_t B, int32_t L, int32_t training,
                   uint64_t seed, int32_t random_chrom, float* logits, float* recon, void* workspace,
                   int64_t workspace_bytes, void* stream);

/* loss = alpha * BCEWithLogits(logits, y, weight=w).mean() + beta * recon   (main.py:56,166)
 * writes loss_out (dev fp32 [3]) = {bce, recon, loss} and dlogit (dev fp32 [B]) = d loss / d logits.
 * recon may be NULL (treated as 0). */
int matcha_bce_loss(const float* logits, const float* y, const float* w, int64_t B, float alpha, float beta,
                    const float* recon, float* dlogit, float* loss_out, void* stream);

/* Backward pass of everything matcha_forward(training=1) did, given d loss / d logits (dev fp32 [B]) and
 * the scalar d loss / d recon (`beta`, 0 skips the reconstruction head).  Uses the tape left in
 * `workspace` by the forward call with the same (x, B, L, seed, random_chrom).  Gradients are
 * ACCUMULATED into m->grads (zero the buffer first).
 * active (dev int32 [2 * n_chrom], optional): per-chromosome flags "encoder c received a gradient" and
 * "recon head c received a gradient" -- the tensors torch would have given a non-None .grad. */
int matcha_backward(const matcha_model_desc* m, const int64_t* x, int64_t B, int32_t L, uint64_t seed,
                    int32_t random_chrom, const float* dlogit, float beta, int32_t* active, void* workspace,
                    int64_t workspace_bytes, void* stream);

/* Classifier.get_node_embeddings in eval mode (Modules.py:252-259; rows of embeddings.npy, main.py:462-476)
 * ids dev int64 [T] -> out dev fp32 [T, d]; workspace as for matcha_forward with B = T, L = 1. */
int matcha_node_embeddings(const matcha_model_desc* m, const int64_t* ids, int64_t T, float* out,
                           void* workspace, int64_t workspace_bytes, void* stream);

/* torch.optim.AdamW step on a flat range (main.py:630,671).  seg_* (dev, optional) describe
 * conditionally-active segments: element range [seg_begin[s], seg_end[s]) is updated only when
 * active[seg_flag[s]] != 0; step counts are per segment (seg_step, dev int32, incremented here).
 * Elements outside every segment but inside [0, n_always) are always updated with step `step`. */
int matcha_adamw(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n_always,
                 int32_t step, int32_t n_seg, const int64_t* seg_begin, const int64_t* seg_end,
                 const int32_t* seg_flag, int32_t* seg_step, const int32_t* active, float lr, float beta1,
                 float beta2, float eps, float weight_decay, float grad_scale, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Data-parallel step boundary over NVLink peer memory (no counterpart in the single-process
 * reference): cross-GPU barrier -> all-reduce of the flat gradient buffers (P2P loads, summed in rank order so every
 * replica computes bit-identical values) -> mean -> AdamW on the own replica (same semantics as matcha_adamw, activity flags
 * OR-ed over ranks) -> barrier -> the own gradient buffer is zeroed for the next step.  One launch instead of
 * ncclAllReduce + AdamW + memset.
 *   grad_ptrs / active_ptrs / barrier_ptrs   HOST arrays of `world` device pointers, entry r = rank r's buffer as mapped
 *                                            into this process (cudaIpcOpenMemHandle); entry `rank` = the own buffer
 *   barrier buffer                           matcha_dp_barrier_bytes() bytes per rank, zero-filled once at set-up
 *   epoch                                    1, 2, 3, ... : the same value on every rank for the same step
 * All ranks must call this once per step with the same arguments (it spins, bounded, until every peer has arrived).
 * world = 1 (barrier_ptrs may be NULL) is the single-replica form: AdamW + gradient-buffer clear in one launch.
 * --------------------------------------------------------------------------------------------- */
int32_t matcha_dp_blocks(void);
int matcha_enable_peer_access(int32_t peer_device);   /* cudaDeviceEnablePeerAccess from the current device, idempotent */
/* Form of the data-parallel step boundary below: 0 = one-shot (every rank reads every peer's whole gradient buffer), 1 =
 * two-shot (rank r reduces the r-th part of every block's slice and writes the mean into all replicas), 2 = by world size
 * (default: two-shot from 4 ranks on); also MATCHA_DP_TWO_SHOT=0 / 1.  Both forms give bit-identical weights. */
void matcha_set_dp_two_shot(int32_t mode);
/* CUDA IPC: handle (64 bytes) of the cudaMalloc allocation starting at base_ptr; map a peer process's allocation for kernels
 * of the current device (cudaIpcMemLazyEnablePeerAccess); unmap */
int matcha_ipc_get_handle(const void* base_ptr, uint8_t* handle64);
int matcha_ipc_open(const uint8_t* handle64, void** mapped_base);
int matcha_ipc_close(void* mapped_base);
int64_t matcha_dp_barrier_bytes(void);
int matcha_dp_reduce_adamw(int32_t world, int32_t rank, const void* const* grad_ptrs, const void* const* active_ptrs,
                           void* const* barrier_ptrs, float* params, float* exp_avg, float* exp_avg_sq,
                           int32_t* active_reduced, int64_t n_always, int64_t n_flat, int32_t n_seg, int32_t n_flags,
                           const int64_t* seg_begin, const int64_t* seg_end, const int32_t* seg_flag, int32_t* seg_step,
                           int32_t step, uint32_t epoch, float lr, float beta1, float beta2, float eps, float weight_decay,
                           void* stream);

/* ---------------------------------------------------------------------------------------------
 * k-mer enumeration + counting — generate_kmers.py:8-69 (build_dict) and :86-141: the producer of
 * all_<k>_counter.npy / all_<k>_freq_counter.npy.
 *   members / offsets   dev int64 CSR of the clusters (unique ascending node ids per cluster, process.py:72-78)
 *   work_prefix         dev int64 [n_clusters + 1]: prefix sums of C(n_c, k) over the clusters with
 *                       k <= n_c <= max_cluster_size (0 work for the others, generate_kmers.py:88-91); total_work = last entry
 *   table               capacity (power of two) slots of 16 bytes, zero-filled by the caller; counts dev int32 [capacity],
 *                       zero-filled; status dev int32 [1], zero-filled (1 table full, 2 bad prefix / cluster > 64, 3 id range)
 * A k-subset is counted iff every adjacent gap of its sorted ids exceeds min_distance (:16-17,:23-32).
 * matcha_kmer_collect appends the k-mers seen >= min_freq times (:40) in arbitrary order (the reference's own
 * order is process-pool completion order); n_out may exceed max_out, in which case nothing past max_out was written.
 * --------------------------------------------------------------------------------------------- */
int matcha_kmer_count(const int64_t* members, const int64_t* offsets, int64_t n_clusters, const int64_t* work_prefix,
                      int64_t total_work, int32_t k, int32_t min_distance, void* table, int64_t capacity,
                      int32_t* counts, int32_t* status, void* stream);
int matcha_kmer_collect(const void* table, int64_t capacity, const int32_t* counts, int32_t k, int32_t min_freq,
                        int64_t* rows, int32_t* freq, int64_t max_out, uint64_t* n_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Positive k-mer hash set + negative sampler — main.py:361-459 (generate_negative), :345-346
 * (neighbor_check), utils.py:75-97 (build_hash; exact set instead of a Bloom filter).
 * Table: `capacity` (power of two) slots of 16 bytes PLUS one trailing status slot, i.e.
 * (capacity + 1) * 16 bytes, caller-allocated and zero-filled; load factor must stay <= 0.5.
 * Keys pack up to 6 ids of 21 bits (ids < 2^21).  chrom_start / chrom_end are HOST arrays.
 * --------------------------------------------------------------------------------------------- */
int matcha_hashset_insert(void* table, int64_t capacity, const int64_t* kmers, int64_t n, int32_t width,
                          void* stream);
int matcha_hashset_contains(const void* table, int64_t capacity, const int64_t* kmers, int64_t n,
                            int32_t width, uint8_t* out, void* stream);
/* pos dev int64 [P, L] (rows sorted ascending, zero padded) -> neg dev int64 [P * neg_num, L];
 * valid dev uint8 [P * neg_num] = 0 where max_rounds candidate rounds were exhausted (row = the positive). */
int matcha_neg_sample(const void* table, int64_t capacity, const int64_t* pos, int64_t P, int32_t L,
                      int32_t neg_num, const int64_t* chrom_start, const int64_t* chrom_end, int32_t n_chrom,
                      int32_t min_dis, uint64_t seed, uint64_t step, int32_t max_rounds, int64_t* neg,
                      uint8_t* valid, int32_t* rounds_used, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Streaming scorers — denoise_contact.py:67-88 (generate_pair_wise + predict), predict_multiway.py:74-87.
 * Pair path: per-node tables D, S [N+1, d] (k = 2 closed form, exact because with two tokens the
 * diagonal-masked attention weight is 1) then logits for the row-major upper-triangular pair range.
 * --------------------------------------------------------------------------------------------- */
int matcha_pair_tables(const matcha_model_desc* m, float* D, float* S, void* workspace,
                       int64_t workspace_bytes, void* stream);
/* pairs (i, j), lo <= i <= j - min_dis, j < hi  enumerated as generate_pair_wise does; writes
 * out[p] for global pair indices p in [p_begin, p_end) (sigmoid applied when apply_sigmoid != 0). */
int matcha_pair_score_range(const float* D, const float* S, const float* cls_w, const float* cls_b, int32_t d,
                            int64_t lo, int64_t hi, int32_t min_dis, int64_t p_begin, int64_t p_end,
                            int32_t apply_sigmoid, float* out, void* stream);
int64_t matcha_pair_count(int64_t lo, int64_t hi, int32_t min_dis);
/* Tensor-core form of the same scorer (tcgen05, bf16x3 split): logit = u_i + u_j - PA_i . PB_j + b with
 * PA = [w.S | D], PB = [D | w.S]; u and b ride in 16 extra k columns as exact bf16 pieces.  matcha_pair_tc_prepare packs the tables of one chromosome (ids [lo, hi)) into
 * `workspace` (dev, 256-byte aligned, matcha_pair_tc_workspace_bytes(lo, hi) bytes); matcha_pair_tc_score_range then
 * writes the same out[p] as matcha_pair_score_range (apply_sigmoid uses the fast exponential: 1e-6 relative). */
int64_t matcha_pair_tc_workspace_bytes(int64_t lo, int64_t hi);
int matcha_pair_tc_prepare(const float* D, const float* S, const float* cls_w, const float* cls_b, int32_t d, int64_t lo,
                           int64_t hi, void* workspace, int64_t workspace_bytes, void* stream);
int matcha_pair_tc_score_range(const void* workspace, int64_t lo, int64_t hi, int32_t min_dis, int64_t p_begin, int64_t p_end,
                               int32_t apply_sigmoid, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Validation metrics on the device — utils.py:32-72 (roc_auc_cuda: sklearn roc_auc_score / average_precision_score
 * over all samples and per hyperedge size; accuracy), called from main.py:189-195,246-252.  SURVEY.md section 8f rank 3.
 *   score, label  dev fp32 [n] (label > 0.5 = positive; accuracy compares score >= 0.5 with label >= 0.5)
 *   cls           dev int32 [n] class id (hyperedge size) in [0, n_classes), or NULL with n_classes = 0
 *   out           dev fp64 [(1 + n_classes) * 4]: row 0 = all samples, row 1 + c = class c;
 *                 columns {AUROC, AUPR, accuracy, count}; AUROC / AUPR are NaN where a row has one label only
 * Ties share one threshold exactly as sklearn's distinct-value curves do.
 * --------------------------------------------------------------------------------------------- */
int64_t matcha_metrics_workspace_bytes(int64_t n);
int matcha_binary_metrics(const float* score, const float* label, const int32_t* cls, int64_t n, int32_t n_classes,
                          double* out, void* workspace, int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Denoise post-processing — denoise_contact.py:31-61 (proba2matrix) and :160-192 (per-chromosome tail).  SURVEY 8f rank 2.
 *   proba      dev fp32 [matcha_pair_count] sigmoid scores of ONE chromosome of n bins, generate_pair_wise order
 *              (what matcha_pair_tc_score_range(apply_sigmoid = 1) writes)
 *   origin     dev fp32: observed contacts of the same bins, element (i, j) at origin[i * origin_ld + j]
 *              (a view into intra_adj.npy; only j >= i + min_dis is read, as :160 does)
 *   my         dev fp32 [n, n]: the reference's `my` before the quantile transform (:177-185)
 * matcha_quantile_uniform applies sklearn QuantileTransformer(output_distribution="uniform")._transform_col in place
 * given the fitted table (quantiles, references: dev fp64 [nq]); matcha_pair_gather reads my at the scored pairs
 * (the `balanced` pixel column, :205); matcha_gather_f32 fetches the fit's subsample.
 * --------------------------------------------------------------------------------------------- */
int64_t matcha_denoise_workspace_bytes(int64_t n);
int matcha_denoise_matrix(const float* proba, const float* origin, int64_t origin_ld, int64_t n, int32_t min_dis, float* my,
                          void* workspace, int64_t workspace_bytes, void* stream);
int matcha_quantile_uniform(float* x, int64_t n, const double* quantiles, const double* references, int32_t nq, void* stream);
int matcha_gather_f32(const float* src, const int64_t* idx, int64_t n, float* out, void* stream);
int matcha_pair_gather(const float* my, int64_t n, int32_t min_dis, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Feature construction — SURVEY.md section 8f rank 4.
 *   matcha_corrcoef              np.corrcoef(A) with NaN -> 0 of one chromosome's intra-contact block (main.py:572-577):
 *                                A dev fp32 [n, lda] (rows = variables), out dev fp32 [n, ldo]; centred and contracted in
 *                                float64 like numpy; workspace matcha_corrcoef_workspace_bytes(n) (an n x n fp64 matrix)
 *   matcha_zscore_positive_rows  Modules.py:147-152 in place on dev fp32 [nrows, ld]: the positive entries of each row are
 *                                z-scored among themselves (ddof 0, fp32 like scipy.stats.mstats.zscore), NaN -> 0
 *   matcha_adj_from_pixels       process.py:148-170: cooler pixels (bin1, bin2 dev int64, count dev fp64, NaN skipped) ->
 *                                intra / inter dev fp64 [N, N] (zero-filled by the caller); cool2node dev int64
 *                                [n_cool_bins] holds the 1-based node id of a cooler bin or <= 0; node2chrom dev int32 [N + 1]
 *   matcha_adj_from_clusters     process.py:90-105: CSR clusters (1-based node ids) -> adj dev fp64 [N, N] += 1 for every
 *                                ordered pair i != j of a cluster
 * --------------------------------------------------------------------------------------------- */
int64_t matcha_corrcoef_workspace_bytes(int64_t n);
int matcha_corrcoef(const float* A, int64_t lda, int64_t n, float* out, int64_t ldo, void* workspace, int64_t workspace_bytes,
                    void* stream);
int matcha_zscore_positive_rows(float* M, int64_t ld, int64_t nrows, int64_t ncols, void* stream);
int matcha_adj_from_pixels(const int64_t* bin1, const int64_t* bin2, const double* count, int64_t n_pixels,
                           const int64_t* cool2node, int64_t n_cool_bins, const int32_t* node2chrom, int64_t n_nodes,
                           double* intra, double* inter, void* stream);
int matcha_adj_from_clusters(const int64_t* members, const int64_t* offsets, int64_t n_clusters, int64_t n_nodes, double* adj,
                             void* stream);
int matcha_f64_to_f32(const double* in, float* out, int64_t n, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Building blocks exposed for tests (dense fp32 contractions used by the passes above).
 *   form 0: C[M,N]  = A[M,K] . B[N,K]^T (+bias)      form 1: C[M,N] = A[M,K] . B[K,N]
 *   form 2: C[M,N] += A[K,M]^T . B[K,N]
 * impl 0 = SIMT fp32, 1 = tcgen05 (bf16x3 split, fp32 accumulate in TMEM) kernels specialised on the embed_dim-64 shapes,
 * 2 = general tcgen05 kernel (any shape with >= 1024 token rows and N >= 64; csrc/gemm_tcg.cu)
 * --------------------------------------------------------------------------------------------- */
int matcha_gemm(int32_t form, int32_t impl, const float* A, const float* B, float* C, const float* bias,
                int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc, float* scratch,
                int64_t scratch_floats, void* stream);
/* floats of scratch the tcgen05 form-2 kernel needs for M output rows (split-K partial sums) */
int64_t matcha_gemm_scratch_floats(int64_t M);

#ifdef __cplusplus
}
#endif
#endif /* MATCHA_B200_H */
