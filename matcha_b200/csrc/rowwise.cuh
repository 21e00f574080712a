// Launchers of the row-wise (per token / per hyperedge) kernels; implementation in rowwise.cu.
#pragma once
#include "common.cuh"

namespace matcha {

constexpr int kD = 64;            // embed_dim handled by this build
constexpr int kH = 8;             // heads
constexpr int kQKG = 3 * kH * kD; // 1536 columns: [Q (8x64) | K (8x64) | G (8x64)]

struct ChromMeta {
  int32_t n;
  int64_t start[MATCHA_MAX_CHROM];
  int64_t end[MATCHA_MAX_CHROM];
};

// counts[0..C-1] tokens per chromosome, counts[C] = pads; group_off[C+1]; perm lists token indices by chromosome
int launch_bucket(const int64_t* x, int64_t T, const ChromMeta& cm, int32_t* counts, int32_t* group_off,
                  int32_t* cursor, int32_t* perm, cudaStream_t s);

int launch_ln_fwd(const float* X, float* xhat, float* rstd, int64_t T, cudaStream_t s);
int launch_attn_fwd(const float* QKG, const int64_t* x, const float* b_dyn, float* U, int64_t B, int L,
                    DropCfg drop, cudaStream_t s);
struct ScoreParams {
  const float *pff_g, *pff_b, *ln1_g, *ln1_b, *ln2_g, *ln2_b, *cls_w, *cls_b;
};
int launch_score_fwd(const float* H2, const float* xhat, const int64_t* x, ScoreParams p, float* logits,
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
int launch_score_bwd(const float* H2, const float* xhat, const float* rstd_x, const int64_t* x, ScoreParams p,
                     const float* dlogit, float* dH2, float* dXs, ScoreGrads g, int64_t B, int L, cudaStream_t s);
int launch_attn_bwd(const float* QKG, const float* dU, const int64_t* x, float* dQKG, float* db_dyn, int64_t B,
                    int L, DropCfg drop, cudaStream_t s);
// dP = (LNbwd(dxhat) + dXs) * (1 - X^2)
int launch_ln_tanh_bwd(const float* dxhat, const float* dXs, const float* xhat, const float* rstd, const float* X,
                       float* dP, int64_t T, cudaStream_t s);
// dE = dV0 + beta * dtE * (1 - tanh(E)^2)   (dtE may be NULL)
int launch_enc_combine_bwd(const float* dV0, const float* dtE, const float* E, float beta, float* dE, int64_t n,
                           cudaStream_t s);
// active[c] = counts[c] > 0 ; active[C + c] = (c == rchrom && eligible > 0)
int launch_active_flags(const int32_t* counts, int n_chrom, int rchrom, int64_t T, int32_t* active, cudaStream_t s);
// per-node tables for the k = 2 closed form: U[n] = sum_h G_h[n] + b_dyn (the other token's attention output)
int launch_pair_u(const float* QKG, const float* b_dyn, float* U, int64_t T, cudaStream_t s);
// D = LN1(LN_pff(H2)), S = LN2 affine of xhat
int launch_pair_ds(const float* H2, const float* xhat, ScoreParams p, float* D, float* S, int64_t T, cudaStream_t s);

int launch_iota_i64(int64_t* out, int64_t n, cudaStream_t s);

}  // namespace matcha
