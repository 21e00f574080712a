// Launchers of the row-wise (per token / per hyperedge) kernels; implementation in rowwise.cu.
#pragma once
#include "common.cuh"

namespace matcha {

constexpr int kD = 64;            // embed_dim handled by this build
constexpr int kH = 8;             // heads
constexpr int kQKG = 3 * kH * kD; // 1536 columns: [Q (8x64) | K (8x64) | G (8x64)]

// ------------------------------------------------------------------------------------------
// MMA-ready tiles written by producers for the tcgen05 kernels of qkg_tiles.cu.
// A tile covers 128 tokens x 64 columns as bf16 hi | bf16 lo, each [plane = column/8][token][8 columns]
// (2048-byte planes).  xhat tiles carry two extra planes per half: a ones column (column 64) that makes the
// weight-gradient MMA also produce the bias gradient, and a zero plane (UMMA N must be a multiple of 16).
// ------------------------------------------------------------------------------------------
constexpr int kTileTok = 128;
constexpr int kPlaneBytes = 2048;
constexpr int kXPlanes = 10;
constexpr int kXHalfBytes = kXPlanes * kPlaneBytes;    // 20480
constexpr int kXTileBytes = 2 * kXHalfBytes;           // 40960 per 128 tokens
constexpr int kGHalfBytes = 8 * kPlaneBytes;           // 16384
constexpr int kGTileBytes = 2 * kGHalfBytes;           // 32768 per (128 tokens, 64-column chunk)
constexpr int kGChunks = kQKG / 64;                    // 24 chunks: Q heads 0-7, K heads 0-7, G heads 0-7
constexpr int64_t kTilePathMinTokens = 1024;           // below this the SIMT path (fp32 tensors) is used
__host__ __device__ inline int64_t num_token_tiles(int64_t T) { return (T + kTileTok - 1) / kTileTok; }
// Hyperedge-aligned tiles of the fused attention kernels (attn_fused.cu): warp-row q of tile i holds the
// rows_per_warp(L) consecutive tokens starting at (4 i + q) * rows_per_warp(L); the other lanes are zero rows.
__host__ __device__ inline int rows_per_warp(int L) { return (32 / L) * L; }
__host__ __device__ inline int64_t num_atiles(int64_t T, int L) { return (T + 4 * rows_per_warp(L) - 1) / (4 * rows_per_warp(L)); }
constexpr int kHeadWBytes = 2 * 192 * 64 * 2;          // per-head [Q_h | K_h | G_h] weights, bf16 hi | lo (49152)

struct ChromMeta {
  int32_t n;
  int64_t start[MATCHA_MAX_CHROM];
  int64_t end[MATCHA_MAX_CHROM];
};

// counts[0..C-1] tokens per chromosome, counts[C] = pads (counts holds MATCHA_MAX_CHROM + 2 ints: the last one is the
// kernel's rendezvous counter); group_off[C+1]; perm lists token indices by chromosome; hist: bucket_hist_ints() ints
int bucket_hist_ints();
int launch_bucket(const int64_t* x, int64_t T, const ChromMeta& cm, int32_t* counts, int32_t* group_off,
                  int32_t* hist, int32_t* perm, cudaStream_t s);

// xhat_tiles (optional): also emit the pre-split tiles (num_token_tiles(T) * kXTileBytes bytes, tail rows zeroed)
int launch_ln_fwd(int d, const float* X, float* xhat, float* rstd, int64_t T, uint8_t* xhat_tiles, cudaStream_t s);
int launch_attn_fwd(int d, const float* QKG, const int64_t* x, const float* b_dyn, float* U, int64_t B, int L,
                    DropCfg drop, cudaStream_t s);
struct ScoreParams {
  const float *pff_g, *pff_b, *ln1_g, *ln1_b, *ln2_g, *ln2_b, *cls_w, *cls_b;
};
int launch_score_fwd(int d, const float* H2, const float* xhat, const int64_t* x, ScoreParams p, float* logits,
                     int64_t B, int L, cudaStream_t s);
int launch_bce(const float* logits, const float* y, const float* w, float alpha, float* dlogit, float* loss_out,
               int64_t B, cudaStream_t s);
int launch_finalize_loss(float* loss_out, const float* recon, float alpha, float beta, cudaStream_t s);
// pred [T, n_r] -> in place (pred - target) * 200 / (T' * n_r); recon_out[0] = 100 * mean mean (..)^2
int launch_recon_diff(float* pred, int64_t ld, const int64_t* x, int64_t T, const float* inter, int64_t inter_ld,
                      int64_t r_start, int64_t r_end, const int32_t* counts, int rchrom, int n_chrom,
                      float* recon_out, cudaStream_t s);

struct ScoreGrads {
  float *pff_g, *pff_b, *ln1_g, *ln1_b, *ln2_g, *ln2_b, *cls_w, *cls_b;
};
int launch_score_bwd(int d, const float* H2, const float* xhat, const float* rstd_x, const int64_t* x, ScoreParams p,
                     const float* dlogit, float* dH2, float* dXs, ScoreGrads g, int64_t B, int L, cudaStream_t s);
// dQKG is written as fp32 [T, 1536], or -- when dqkg_tiles != NULL -- only as pre-split tiles
// [token tile][24 chunks][kGTileBytes] (the caller zeroes the tail rows of the last tile)
int launch_attn_bwd(int d, const float* QKG, const float* dU, const int64_t* x, float* dQKG, uint8_t* dqkg_tiles, float* db_dyn,
                    int64_t B, int L, DropCfg drop, cudaStream_t s);

// fused attention kernels (attn_fused.cu)
int launch_split_w_heads(const float* W, void* out, cudaStream_t s);   // W_qkg [1536, 64] -> 8 per-head chunks (kHeadWBytes)
int launch_ln_fwd_atiles(const float* X, float* xhat, float* rstd, int64_t T, int L, uint8_t* xhat_tiles, cudaStream_t s);
// U = dropout(sum_h softmax(Q_h K_h^T) G_h + b_dyn) * non_pad_mask straight from hyperedge-aligned xhat tiles;
// probs (optional) [T, 8, 4 (L <= 5) or 8] keeps the attention weights for the backward pass
int launch_attn_fused_fwd(const uint8_t* xhat_tiles, const uint8_t* wheads, const float* bq, const float* b_dyn,
                          const int64_t* x, float* U, float* probs, int64_t B, int L, DropCfg drop, cudaStream_t s);

// X-form forward (attn_xform.cu, L <= 5): neighbours' xhat rows stay in registers, no per-head shuffles.
//   wx  8 per-head blocks of kXformWBytes (N_h = Wk_h^T Wq_h | Wg_h, bf16 hi | lo), vx [8, 64] = Wk_h^T bq_h: launch_prep_xform
//   passes  3 = bf16x3 split contractions (fp32-accurate), 1 = single bf16 pass (stated-tolerance mode)
constexpr int kXformWBytes = 32768;
int launch_prep_xform(const float* W, const float* bq, void* wx, float* vx, cudaStream_t s);
int launch_attn_xform_fwd(const uint8_t* xhat_tiles, const float* xhat, const uint8_t* wx, const float* vx, const float* b_dyn,
                          const int64_t* x, float* U, float* probs, int64_t B, int L, DropCfg drop, int passes, cudaStream_t s);

// backward of the same block: recompute per head pair, shuffle attention backward, tcgen05 data and weight gradients.
//   wpairs       launch_split_w_pairs output (4 head pairs x 3 pieces x 32 KB)
//   dxhat_parts  [4, T, 64] per-head-pair partial data gradients (summed by launch_ln_tanh_bwd)
//   part         attn_fused_bwd_scratch_floats() floats of split-K partials; dW [1536, 64] += their sum
//   dbq [512], db_dyn [64] accumulate (atomics); dU is masked / dropout-scaled IN PLACE first unless `premasked`
int launch_split_w_pairs(const float* W, void* out, cudaStream_t s);
int64_t attn_fused_bwd_scratch_floats();
int launch_attn_fused_bwd(const uint8_t* xhat_tiles, const uint8_t* wpairs, const float* bq, const int64_t* x, float* dU,
                          const float* probs, float* dxhat_parts, float* part, float* dW, float* dbq, float* db_dyn,
                          int64_t B, int L, DropCfg drop, int premasked, int passes, cudaStream_t s);
constexpr int kPairWBytes = 3 * 32768;                 // per head pair: G | K | Q piece pairs, bf16 hi | lo

// row-chain kernels (chain.cu): the 64-wide layers around the attention block on tensor cores, thread = tile row
constexpr int kChainWBytes = 16384;                    // one 64x64 weight, bf16 hi 8 KB | lo 8 KB
// next_w, pff_w0, pff_w1 [64, 64] -> K-major and MN-major bf16 hi | lo copies, one launch
int launch_split_w64x3(const float* W0, const float* W1, const float* W2, void* k0, void* mn0, void* k1, void* mn1, void* k2,
                       void* mn2, cudaStream_t s);
// V0 = E + attribute_nn(attr[id]); X = tanh(next_w V0 + b); xhat / rstd; hyperedge-aligned xhat tiles
// (optional: V0, V0 tiles and attribute-row tiles for the backward pass)
int launch_chain_mix_fwd(const float* E, const int64_t* x, const float* attr_table, int attr_dim, const float* attr_w,
                         const float* attr_b, const void* w_next_k, const float* next_b, float* V0, float* X, float* xhat,
                         float* rstd, uint8_t* xhat_tiles, uint8_t* v0_tiles, uint8_t* attr_tiles, int64_t B, int L,
                         cudaStream_t s);
// pff_n1 + scorer: U -> H1d -> H2 -> logits (H1d / H2 / tiles optional: kept for the backward pass)
int launch_chain_pff_fwd(const float* U, const float* xhat, const int64_t* x, const void* w0_k, const void* w1_k,
                         const float* b0, const float* b1, ScoreParams p, DropCfg drop, float* H1d, float* H2, float* logits,
                         uint8_t* u_tiles, uint8_t* h1_tiles, int64_t B, int L, cudaStream_t s);
// backward of pff_n1 from dH2: writes the masked / dropout-scaled gradient of the attention output (dd) and the dH2 /
// dH1pre tiles (32 KB each) consumed by launch_wgrad_pair
int launch_chain_pff_bwd(const float* dH2, const float* H1d, const int64_t* x, const void* w1_mn, const void* w0_mn,
                         DropCfg dpff, DropCfg dattn, float* dd, uint8_t* dh2_tiles, uint8_t* dh1_tiles, int64_t B, int L,
                         cudaStream_t s);
// LayerNorm / tanh / next_w backward: dE = dV0 + beta * dtE * (1 - tanh(E)^2), plus the dP / dV0 tiles
int launch_chain_mix_bwd(const float* dxhat, int nparts, int64_t part_stride, const float* dXs, const float* xhat,
                         const float* rstd, const float* X, const float* dtE, const float* E, float beta, const void* wn_mn,
                         float* dE, uint8_t* dp_tiles, uint8_t* dv0_tiles, int64_t B, int L, cudaStream_t s);
// two stacked 64-row weight gradients (+ bias gradients) from pre-split tiles: w1 [64, n1] += A1^T B1, w2 [64, n2] += A2^T B2
int64_t wgrad_pair_scratch_floats();
int launch_wgrad_pair(const uint8_t* a1, const uint8_t* a2, const uint8_t* b1, const uint8_t* b2, int b2_planes, int64_t ntiles,
                      float* scratch, float* w1, int ld1, int n1, float* bias1, float* w2, int ld2, int n2, float* bias2,
                      cudaStream_t s);

// fused reconstruction head on tensor cores (recon_tc.cu): the loss is added to recon_out[0] (if given); mode 1 also adds
// beta * the weight / bias gradients to dRw [n_r, 64] / drb [n_r] and gdiff . Rw to dtE [T, 64] (zeroed by the caller).
// perm / group_off / counts: the chromosome-bucketed token list of launch_bucket (the kernel walks the eligible tokens only)
int launch_recon_tc(const float* E, const int64_t* x, int64_t T, const float* inter, int64_t inter_ld, int64_t rs, int64_t re,
                    const float* Rw, const float* rb, const int32_t* counts, int rchrom, int n_chrom, const int32_t* perm,
                    const int32_t* group_off, float* recon_out, float* dRw, float* drb, float* dtE, float beta, int mode,
                    cudaStream_t s);

int launch_axpy(const float* in, float scale, float* out, int64_t n, cudaStream_t s);      // out += scale * in

// pipelined gradient pass of the same head (recon_pipe.cu; same outputs as mode 1 with beta = 1): a pre-pass writes tanh(E) of
// the eligible tokens as operand tiles into `tiles` (recon_pipe_tile_bytes(T) bytes), one persistent CTA per SM walks an equal
// share of the (token tile x column block) units with the target loads of the next unit in flight
int64_t recon_pipe_tile_bytes(int64_t T);
int launch_recon_pipe(const float* E, const int64_t* x, int64_t T, const float* inter, int64_t inter_ld, int64_t rs, int64_t re,
                      const float* Rw, const float* rb, const int32_t* counts, int rchrom, int n_chrom, const int32_t* perm,
                      const int32_t* group_off, float* recon_out, float* dRw, float* drb, float* dtE, uint8_t* tiles,
                      cudaStream_t s);

// node encoder forward on tensor cores (enc_tc.cu): dense feature rows, embed_dim 64.  The pre-split weight chunks live in
// the derived buffer from float offset split_base (enc_tc_split_floats(m) floats, rebuilt by launch_enc_tc_prepare)
int64_t enc_tc_split_floats(const matcha_model_desc* m);
int launch_enc_tc_prepare(const matcha_model_desc* m, int64_t split_base, cudaStream_t s);
int launch_enc_tc_fwd(const matcha_model_desc* m, int64_t split_base, const int64_t* x, int64_t T, const int32_t* perm,
                      const int32_t* group_off, float* H0, float* E, DropCfg drop, cudaStream_t s);

// backward of the same two layers: dW1_c and dW0_c accumulated into m->grads (chromosomes up to 384 bins: 6 feature
// chunks resident in TMEM; enc_tc_bwd_fits says whether the model qualifies)
bool enc_tc_bwd_fits(const matcha_model_desc* m);
void set_enc_pipe(int fwd, int bwd);     // pipelined encoder kernels: fwd 0 / 1 / 2 (by row width), bwd 0 / 1; -1 = environment
int launch_enc_tc_bwd(const matcha_model_desc* m, int64_t split_base, const int64_t* x, int64_t T, const int32_t* perm,
                      const int32_t* group_off, const float* dE, const float* H0, DropCfg drop, cudaStream_t s);

// CSR first encoder layer (csr_encoder.cu): feature rows given as CSR (feat[c] == NULL, feat_indptr/indices/values set).
// W0T_c [n_c, 64] copies live in the derived buffer from float offset w0t_base (chromosome after chromosome); the same
// offsets of derived_grad accumulate dW0T_c.
bool model_uses_csr(const matcha_model_desc* m);
int launch_csr_prepare(const matcha_model_desc* m, int64_t w0t_base, cudaStream_t s);
int launch_enc0_csr_fwd(const matcha_model_desc* m, int64_t w0t_base, const int64_t* x, int64_t T, const int32_t* perm,
                        const int32_t* group_off, float* H0, DropCfg drop, cudaStream_t s);
int launch_enc0_csr_wgrad(const matcha_model_desc* m, int64_t w0t_base, const int64_t* x, int64_t T, const int32_t* perm,
                          const int32_t* group_off, const float* dH0pre, DropCfg drop, cudaStream_t s);

// tcgen05 tile kernels (qkg_tiles.cu)
int launch_split_wT(const float* W, void* out, cudaStream_t s);     // W [1536, 64] fp32 -> MN-major chunks for the dgrad
int tc_qkg_forward_tiles(const uint8_t* xhat_tiles, const uint8_t* w_split, const float* bias, float* QKG, int64_t T,
                         cudaStream_t s);
int tc_qkg_dgrad_tiles(const uint8_t* dqkg_tiles, const uint8_t* wT_split, float* dxhat, int64_t T, cudaStream_t s);
int tc_qkg_wgrad_tiles(const uint8_t* dqkg_tiles, const uint8_t* xhat_tiles, float* scratch, int64_t scratch_floats,
                       float* dW, float* dbias, int64_t dbias_n, int64_t T, cudaStream_t s);
// dP = (LNbwd(sum of nparts dxhat partials, part_stride floats apart) + dXs) * (1 - X^2)
int launch_ln_tanh_bwd(int d, const float* dxhat, int nparts, int64_t part_stride, const float* dXs, const float* xhat,
                       const float* rstd, const float* X, float* dP, int64_t T, cudaStream_t s);
// dE = dV0 + beta * dtE * (1 - tanh(E)^2)   (dtE may be NULL)
int launch_enc_combine_bwd(const float* dV0, const float* dtE, const float* E, float beta, float* dE, int64_t n,
                           cudaStream_t s);
// active[c] = counts[c] > 0 ; active[C + c] = (c == rchrom && eligible > 0)
int launch_active_flags(const int32_t* counts, int n_chrom, int rchrom, int64_t T, int32_t* active, cudaStream_t s);
// per-node tables for the k = 2 closed form: U[n] = sum_h G_h[n] + b_dyn (the other token's attention output)
int launch_pair_u(int d, const float* QKG, const float* b_dyn, float* U, int64_t T, cudaStream_t s);
// D = LN1(LN_pff(H2)), S = LN2 affine of xhat
int launch_pair_ds(int d, const float* H2, const float* xhat, ScoreParams p, float* D, float* S, int64_t T, cudaStream_t s);

int launch_iota_i64(int64_t* out, int64_t n, cudaStream_t s);

}  // namespace matcha
