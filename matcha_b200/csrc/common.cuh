// Shared device/host helpers for the matcha_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/matcha_b200.h"

namespace matcha {

void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);

#define MATCHA_CHECK_LAUNCH(what)                                         \
  do {                                                                    \
    int _rc = ::matcha::check_cuda(cudaGetLastError(), what);             \
    if (_rc) return _rc;                                                  \
  } while (0)

#define MATCHA_REQUIRE(cond, ...)                                         \
  do {                                                                    \
    if (!(cond)) {                                                        \
      ::matcha::set_error(__VA_ARGS__);                                   \
      return MATCHA_ERR_ARG;                                              \
    }                                                                     \
  } while (0)

constexpr int kSMs = 148;  // B200: 2 dies x 74 SMs

// ------------------------------------------------------------------------------------------
// Optional per-call-site timing (CUDA events on the launching stream) -- off unless
// matcha_profile_enable(1) was called; used by bench.py for the roofline figures.
// ------------------------------------------------------------------------------------------
enum ProfLabel : int {
  P_BUCKET = 0, P_ENC0, P_ENC1, P_RECON_PRED, P_RECON_DIFF, P_ATTR, P_MIX, P_LN, P_QKG, P_ATTN_FWD, P_PFF0, P_PFF1,
  P_SCORE_FWD, P_BCE, P_SCORE_BWD, P_W_PFF1, P_D_PFF1, P_W_PFF0, P_D_PFF0, P_ATTN_BWD, P_W_QKG, P_D_QKG, P_LN_BWD,
  P_W_NEXT, P_D_NEXT, P_W_ATTR, P_W_RECON, P_D_RECON, P_ENC_COMBINE, P_W_ENC1, P_D_ENC1, P_W_ENC0, P_PREP, P_PREP_BWD,
  P_ADAMW, P_SAMPLER, P_PAIR_SCORE, P_MISC, P_COUNT
};
void prof_begin(int label, cudaStream_t s);
void prof_end(int label, int n_kernels, cudaStream_t s);
#define PROF(label, nk, call)                 \
  ([&]() -> int {                             \
    ::matcha::prof_begin(label, s);           \
    int _prc = (call);                        \
    ::matcha::prof_end(label, nk, s);         \
    return _prc;                              \
  }())

// ------------------------------------------------------------------------------------------
// Counter-based RNG, restated bit-exactly in oracle/hypersagnn_oracle.py (splitmix64).
// ------------------------------------------------------------------------------------------
constexpr uint64_t kGolden = 0x9E3779B97F4A7C15ull;
constexpr uint64_t kSiteMult = 0xD1B54A32D192ED03ull;
enum DropSite : int { SITE_FEATURE = 1, SITE_ATTN = 2, SITE_PFF = 3 };

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
  z += kGolden;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t site_key(uint64_t seed, int site) {
  return splitmix64(seed ^ ((uint64_t)site * kSiteMult));
}
__host__ __device__ __forceinline__ uint32_t dropout_thr16(float p) { return (uint32_t)(p * 65536.0f + 0.5f); }

struct DropCfg {
  uint64_t key;    // site_key(seed, site)
  uint32_t thr;    // keep iff 16-bit draw >= thr ; thr == 0 disables dropout
  float scale;     // 1 / (1 - p)
};
__host__ __device__ inline DropCfg make_drop(uint64_t seed, int site, float p, bool on) {
  DropCfg c;
  c.key = site_key(seed, site);
  c.thr = (on && p > 0.f) ? dropout_thr16(p) : 0u;
  c.scale = (on && p > 0.f) ? 1.0f / (1.0f - p) : 1.0f;
  return c;
}
// 64 random bits shared by columns [4q, 4q+4) of row `row`
__device__ __forceinline__ uint64_t drop_word(const DropCfg& c, uint64_t row, uint32_t col) {
  return splitmix64(c.key + (row << 24) + (uint64_t)(col >> 2));
}
__device__ __forceinline__ float drop_apply(const DropCfg& c, uint64_t word, uint32_t col, float v) {
  uint32_t bits = (uint32_t)(word >> (16 * (col & 3))) & 0xFFFFu;
  return bits >= c.thr ? v * c.scale : 0.f;
}
__device__ __forceinline__ float4 drop_apply4(const DropCfg& c, uint64_t row, uint32_t col4, float4 v) {
  if (c.thr == 0u) return v;
  uint64_t w = drop_word(c, row, col4);
  v.x = drop_apply(c, w, 0, v.x);
  v.y = drop_apply(c, w, 1, v.y);
  v.z = drop_apply(c, w, 2, v.z);
  v.w = drop_apply(c, w, 3, v.w);
  return v;
}
// keep * scale factor only (for backward passes)
__device__ __forceinline__ float4 drop_factor4(const DropCfg& c, uint64_t row, uint32_t col4) {
  return drop_apply4(c, row, col4, make_float4(1.f, 1.f, 1.f, 1.f));
}

// ------------------------------------------------------------------------------------------
// Generic fp32 contraction descriptor (see gemm_simt.cu)
// ------------------------------------------------------------------------------------------
struct GemmGroup {       // one problem of a grouped launch (one chromosome)
  const float* A;        // form NT/NN: gathered operand base (feature table) or NULL -> use GemmDesc.A
  int64_t lda;
  int64_t a_id_off;      // physical row = ids[t] - a_id_off
  const float* B;        // per-group weight
  int64_t ldb;
  float* C;              // TN form: per-group output
  int64_t ldc;
  int32_t dim;           // NT/NN: K of this group; TN: N of this group
  int32_t pad;
};

enum GemmForm : int { FORM_NT = 0, FORM_NN = 1, FORM_TN = 2 };

struct GemmDesc {
  int form;
  int64_t M, N, K;
  const float* A; int64_t lda;
  const float* B; int64_t ldb;
  float* C; int64_t ldc;
  // token indirection: logical row r (or k for TN) -> token t = perm ? perm[off + r] : r
  const int32_t* perm;
  // operand row gather through node ids: physical row = ids[t] - id_off (NULL -> physical row = t)
  const int64_t* a_ids; int64_t a_id_off;
  const int64_t* b_ids; int64_t b_id_off;   // TN form only
  int a_act, b_act;                         // 1 = tanh applied on load
  int drop_on;                              // 0 none, 1 = A operand, 2 = B operand (row = token, col = k / n index)
  DropCfg drop;
  // epilogue (NT / NN)
  const float* bias;
  const float* addend; int64_t ld_add;
  int epi_act;                              // 0 none, 1 = tanh, 2 = multiply by tanh'(.) recovered from `aux`
  const float* aux; int64_t ld_aux;         //   aux holds tanh output (after dropout when epi_drop is set)
  int epi_drop; DropCfg edrop;              // dropout after activation (row = token, col = n); with epi_act 2
                                            //   the same mask/scale is applied to the incoming gradient
  float out_scale;                          // accumulator scale (0 is treated as 1)
  // TN: also accumulate column sums of A (= bias gradient) for the first colsum_n columns
  float* colsum; int64_t colsum_n;
  // grouped launch
  int ngroups; const GemmGroup* groups; const int32_t* group_off;
  int64_t max_group_dim;                    // TN grouped: max N over groups (grid sizing)
  int64_t total_rows;                       // grouped: total tokens (upper bound for grid sizing)
  // tcgen05 TN path: split-K partial sums live here (gemm_tc_scratch_floats(M) floats)
  float* scratch; int64_t scratch_floats;
  // tcgen05 NT path: B pre-split into bf16 hi|lo canonical chunks by launch_split_weights_k64 (N*64*4 bytes)
  const uint8_t* b_split;
};

constexpr int kTcMaxSplits = 128;
int64_t gemm_tc_scratch_floats(int64_t M);
int launch_split_weights_k64(const float* W, int64_t ldw, int64_t N, void* out, cudaStream_t stream);

int launch_gemm_simt(const GemmDesc& d, cudaStream_t stream);
int launch_gemm_tc(const GemmDesc& d, cudaStream_t stream, bool* handled);
// general tcgen05 kernel (gemm_tcg.cu): gathered / grouped / ragged problems with >= 1024 token rows
int launch_gemm_tcg(const GemmDesc& d, cudaStream_t stream, bool* handled);

}  // namespace matcha
