// Host-side sequencing of the Hyper-SAGNN hot path behind the C ABI (include/matcha_b200.h):
// parameter preparation, forward, backward, AdamW.  Every dense product is a GemmDesc handed to
// the contraction kernels; everything else is a row-wise kernel from rowwise.cu.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "rowwise.cuh"

namespace matcha {

// ------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return MATCHA_OK;
  set_error("%s: %s", what, cudaGetErrorString(e));
  return MATCHA_ERR_CUDA;
}

// ------------------------------------------------------------------------------------------
// profiling (see common.cuh): event pairs are created lazily and reused
// ------------------------------------------------------------------------------------------
namespace {
struct ProfState {
  bool on = false;
  std::vector<cudaEvent_t> free_events;
  struct Rec { int label; cudaEvent_t b, e; };
  std::vector<Rec> recs;
  cudaEvent_t open_begin[P_COUNT] = {};
  int64_t kernels[P_COUNT] = {};
  cudaEvent_t get() {
    if (!free_events.empty()) { cudaEvent_t e = free_events.back(); free_events.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
  }
};
ProfState g_prof;
const char* kProfNames[P_COUNT] = {
    "bucket_tokens", "enc0_gather_gemm", "enc1_gemm", "recon_pred_gemm", "recon_diff", "attr_gemm", "mix_gemm", "ln_fwd",
    "qkg_gemm", "attn_fwd", "pff0_gemm", "pff1_gemm", "score_fwd", "bce_loss", "score_bwd", "pff1_wgrad", "pff1_dgrad",
    "pff0_wgrad", "pff0_dgrad", "attn_bwd", "qkg_wgrad", "qkg_dgrad", "ln_tanh_bwd", "mix_wgrad", "mix_dgrad",
    "attr_wgrad", "recon_wgrad", "recon_dgrad", "enc_combine_bwd", "enc1_wgrad", "enc1_dgrad", "enc0_wgrad", "param_prep",
    "param_prep_bwd", "adamw", "neg_sampler", "pair_score", "misc"};
}  // namespace
void prof_begin(int label, cudaStream_t s) {
  if (!g_prof.on) return;
  cudaEvent_t b = g_prof.get();
  cudaEventRecord(b, s);
  g_prof.open_begin[label] = b;
}
void prof_end(int label, int n_kernels, cudaStream_t s) {
  if (!g_prof.on) return;
  cudaEvent_t e = g_prof.get();
  cudaEventRecord(e, s);
  g_prof.recs.push_back({label, g_prof.open_begin[label], e});
  g_prof.kernels[label] += n_kernels;
}

static int g_gemm_impl = -1;  // -1 = read MATCHA_GEMM_IMPL on first use; 0 = SIMT only; 1 = tcgen05 where eligible
static int gemm_impl() {
  if (g_gemm_impl < 0) {
    const char* e = getenv("MATCHA_GEMM_IMPL");
    g_gemm_impl = (e && e[0] == '0') ? 0 : 1;   // tcgen05 for the eligible shapes unless MATCHA_GEMM_IMPL=0
  }
  return g_gemm_impl;
}
static bool use_tiles(const matcha_model_desc* m, int64_t T) { return m->d == kD && gemm_impl() == 1 && T >= kTilePathMinTokens; }
static int g_fused = -1;      // -1 = read MATCHA_FUSED on first use; 0 = decomposed pipeline; 1 = fused hyperedge-tile kernels
static int fused_impl() {
  if (g_fused < 0) {
    const char* e = getenv("MATCHA_FUSED");
    g_fused = (e && e[0] == '0') ? 0 : 1;
  }
  return g_fused;
}
// fused attention forward (eval and training) and backward: QKG and its gradient never leave the SM
enum { CW_NEXT_K = 0, CW_NEXT_MN, CW_PFF0_K, CW_PFF0_MN, CW_PFF1_K, CW_PFF1_MN };
static int g_chain = -1;      // row-chain kernels for the 64-wide layers (MATCHA_CHAIN=0 keeps the SIMT contractions)
static bool use_fused(const matcha_model_desc* m, int64_t T, int L);
static bool use_chain(const matcha_model_desc* m, int64_t T, int L) {
  if (g_chain < 0) {
    const char* e = getenv("MATCHA_CHAIN");
    g_chain = (e && e[0] == '0') ? 0 : 1;
  }
  return g_chain == 1 && use_fused(m, T, L) && m->attr_dim <= 32;
}
static bool use_fused(const matcha_model_desc* m, int64_t T, int L) {
  return m->d == kD && fused_impl() == 1 && gemm_impl() == 1 && L >= 2 && L <= 6 && T >= kTilePathMinTokens;
}

static int g_xform = -1;      // X-form fused attention forward (MATCHA_XFORM=0 keeps the Q/K/G shuffle form of attn_fused.cu)
static bool use_xform(const matcha_model_desc* m, int64_t T, int L) {
  if (g_xform < 0) {
    const char* e = getenv("MATCHA_XFORM");
    g_xform = (e && e[0] == '0') ? 0 : 1;
  }
  return g_xform == 1 && use_fused(m, T, L) && L <= 5;
}
static int g_passes = -1;     // bf16 products per tcgen05 contraction step in the attention block: 3 (default) or 1 (MATCHA_BF16=1)
static int mma_passes() {
  if (g_passes < 0) {
    const char* e = getenv("MATCHA_BF16");
    g_passes = (e && e[0] == '1') ? 1 : 3;
  }
  return g_passes;
}

static int g_enc_tc = -1;     // tensor-core node encoder forward (MATCHA_ENC_TC=0 keeps the two grouped SIMT launches)
static bool enc_tc_eligible(const matcha_model_desc* m) { return m->d == kD && !model_uses_csr(m); }
static bool use_enc_tc(const matcha_model_desc* m, int64_t T) {
  if (g_enc_tc < 0) {
    const char* e = getenv("MATCHA_ENC_TC");
    g_enc_tc = (e && e[0] == '0') ? 0 : 1;
  }
  return g_enc_tc == 1 && enc_tc_eligible(m) && gemm_impl() == 1 && T >= kTilePathMinTokens;
}
static int g_recon_tc = -1;   // fused tensor-core reconstruction head (MATCHA_RECON_TC=0 keeps the four SIMT launches)
static bool use_recon_tc(const matcha_model_desc* m, int64_t T, int L) {
  if (g_recon_tc < 0) {
    const char* e = getenv("MATCHA_RECON_TC");
    g_recon_tc = (e && e[0] == '0') ? 0 : 1;
  }
  return g_recon_tc == 1 && use_fused(m, T, L);
}
static int g_recon_pipe = -1;   // pipelined gradient pass of that head (MATCHA_RECON_PIPE=0 keeps the unit-serial kernel)
static bool use_recon_pipe() {
  if (g_recon_pipe < 0) {
    const char* e = getenv("MATCHA_RECON_PIPE");
    g_recon_pipe = (e && e[0] == '0') ? 0 : 1;
  }
  return g_recon_pipe == 1;
}

// general tcgen05 contraction kernel (gemm_tcg.cu) for the models the fused embed_dim-64 kernels do not cover (embed_dim 128):
// MATCHA_GEMM_TCG=0 keeps the fp32 SIMT kernel.  g_tcg_model is set per call by validate()
static int g_tcg = -1;
static bool g_tcg_model = false;
static bool use_tcg() {
  if (g_tcg < 0) {
    const char* e = getenv("MATCHA_GEMM_TCG");
    g_tcg = (e && e[0] == '0') ? 0 : 1;
  }
  return g_tcg == 1 && g_tcg_model;
}

static int run_gemm(const GemmDesc& d, cudaStream_t s, int label, bool allow_tc = true) {
  prof_begin(label, s);
  int rc = MATCHA_OK;
  bool handled = false;
  if (allow_tc && gemm_impl() == 1) rc = launch_gemm_tc(d, s, &handled);
  if (!rc && !handled && allow_tc && gemm_impl() == 1 && use_tcg()) rc = launch_gemm_tcg(d, s, &handled);
  if (!rc && !handled) rc = launch_gemm_simt(d, s);
  prof_end(label, 1, s);
  return rc;
}

// ------------------------------------------------------------------------------------------
// derived-parameter layout
// ------------------------------------------------------------------------------------------
struct DerivedLayout {
  int64_t wqkg, bqkg, bdyn, bdyn_part, wsplit, wtsplit, wheads, wpairs, wchain, wx, vx, tables, total;  // float offsets
};
static __host__ __device__ DerivedLayout derived_layout(int D) {
  DerivedLayout l;
  const int64_t QKG = 3 * kH * D;
  l.wqkg = 0;
  l.bqkg = l.wqkg + QKG * D;
  l.bdyn = l.bqkg + QKG;
  l.bdyn_part = l.bdyn + D;         // per-head partial sums of b_dyn (summed in a fixed order: deterministic)
  l.wsplit = (l.bdyn_part + kH * D + 255) / 256 * 256;    // W_qkg pre-split to bf16 hi|lo chunks (same byte count)
  l.wtsplit = l.wsplit + QKG * D;                        // W_qkg^T pre-split, MN-major chunks (data-gradient B operand)
  l.wheads = l.wtsplit + QKG * D;                        // per-head [Q_h | K_h | G_h] chunks for the fused attention kernels
  l.wpairs = l.wheads + (int64_t)kH * kHeadWBytes / 4;   // per-head-pair G | K | Q piece pairs for the fused backward
  l.wchain = l.wpairs + (int64_t)4 * kPairWBytes / 4;    // next_w, pff_w0, pff_w1: K-major and MN-major pre-split copies
  l.wx = l.wchain + (int64_t)6 * kChainWBytes / 4;       // X-form attention forward: per-head N_h | Wg_h blocks, then v_h
  l.vx = l.wx + (int64_t)kH * kXformWBytes / 4;
  l.tables = l.vx + kH * D;
  l.tables = (l.tables + 63) / 64 * 64;
  const int64_t table_floats = (int64_t)(sizeof(GemmGroup) * MATCHA_MAX_CHROM + 3) / 4;
  l.total = l.tables + 4 * table_floats;
  return l;
}
enum { TAB_ENC0 = 0, TAB_ENC1 = 1, TAB_GW1 = 2, TAB_GW0 = 3 };
static __host__ __device__ const GemmGroup* table_ptr(const float* derived, int D, int which) {
  const DerivedLayout l = derived_layout(D);
  const int64_t table_floats = (int64_t)(sizeof(GemmGroup) * MATCHA_MAX_CHROM + 3) / 4;
  return reinterpret_cast<const GemmGroup*>(derived + l.tables + which * table_floats);
}

// ------------------------------------------------------------------------------------------
// parameter preparation kernels
//   Q = LN_q(X) Wq^T / sqrt(d) = xhat (Wq diag(g_q) / sqrt(d))^T + Wq b_q / sqrt(d)
//   K = xhat (Wk diag(g_k))^T                      (+ Wk b_k, constant over keys -> softmax-invariant, dropped)
//   G_h = xhat (fc1_h Wv_h diag(g_v))^T            (fc1 folded into the value projection, per head)
//   b_dyn = fc1.bias + sum_h fc1_h Wv_h b_v        (softmax rows sum to 1)
// ------------------------------------------------------------------------------------------
__global__ void prep_qk_kernel(const matcha_model_desc m) {
  const int D = m.d;
  const DerivedLayout l = derived_layout(D);
  const float* P = m.params;
  float* W = m.derived + l.wqkg;
  float* bq = m.derived + l.bqkg;
  const float inv = rsqrtf((float)D);
  const int r = blockIdx.x;  // 0 .. 2*H*D-1
  const int c = threadIdx.x; // 0 .. D-1
  if (r < kH * D) {
    const float w = P[m.off_wq + (int64_t)r * D + c];
    W[(int64_t)r * D + c] = w * P[m.off_lnq_g + c] * inv;
    float part = w * P[m.off_lnq_b + c] * inv;
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    __shared__ float s2[4];
    if ((c & 31) == 0) s2[c >> 5] = part;
    __syncthreads();
    if (c == 0) { float t = 0.f; for (int i = 0; i < D / 32; ++i) t += s2[i]; bq[r] = t; }
  } else if (r < 2 * kH * D) {
    const int rk = r - kH * D;
    W[(int64_t)r * D + c] = P[m.off_wk + (int64_t)rk * D + c] * P[m.off_lnk_g + c];
    if (c == 0) bq[r] = 0.f;
  } else {
    // (Wv b_v)[h*D + mm], parked in the (otherwise zero) G part of the folded bias until prep_tables clears it
    const int rv = r - 2 * kH * D;
    float part = P[m.off_wv + (int64_t)rv * D + c] * P[m.off_lnv_b + c];
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    __shared__ float s3[4];
    if ((c & 31) == 0) s3[c >> 5] = part;
    __syncthreads();
    if (c == 0) { float t = 0.f; for (int i = 0; i < D / 32; ++i) t += s3[i]; bq[r] = t; }
  }
}
// one thread per (head h, output row o, column c): Wg_h[o][c] = sum_mm fc1[o][h*D+mm] * Wv[h*D+mm][c] * g_v[c]; the
// operands are tiny (L2 / L1 resident), a warp reads one fc1 value (broadcast) and one coalesced Wv row per step
__global__ void prep_g_kernel(const matcha_model_desc m) {
  const int D = m.d;
  const DerivedLayout l = derived_layout(D);
  const float* P = m.params;
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= (int64_t)kH * D * D) return;
  const int c = (int)(u % D), o = (int)((u / D) % D), h = (int)(u / ((int64_t)D * D));
  const float* f = P + m.off_fc1_w + (int64_t)o * (kH * D) + h * D;
  const float* v = P + m.off_wv + (int64_t)(h * D) * D;
  float s = 0.f;
  for (int mm = 0; mm < D; ++mm) s = fmaf(__ldg(f + mm), __ldg(v + (int64_t)mm * D + c), s);
  m.derived[l.wqkg + (int64_t)(2 * kH * D + h * D + o) * D + c] = s * P[m.off_lnv_g + c];
  if (c == 0) {
    // b_dyn partial of head h, row o: sum_mm fc1_h[o][mm] * (Wv_h b_v)[mm]
    const float* vb = m.derived + l.bqkg + 2 * kH * D + h * D;
    float acc = 0.f;
    for (int mm = 0; mm < D; ++mm) acc = fmaf(__ldg(f + mm), vb[mm], acc);
    m.derived[l.bdyn_part + h * D + o] = acc;
  }
}
__global__ void prep_bdyn_kernel(const matcha_model_desc m) {
  const int D = m.d;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {   // b_dyn = fc1.bias + sum_h (fc1_h Wv_h b_v), heads added in a fixed order
    const DerivedLayout l = derived_layout(D);
    float s = m.params[m.off_fc1_b + c];
    for (int h = 0; h < kH; ++h) s += m.derived[l.bdyn_part + h * D + c];
    m.derived[l.bdyn + c] = s;
    for (int h = 0; h < kH; ++h) m.derived[l.bqkg + 2 * kH * D + h * D + c] = 0.f;   // the G part of the folded bias is zero
  }
}
// grouped-contraction tables of the node encoder (pointers and leading dimensions only: no dependence on the other folds)
__global__ void prep_tables_kernel(const matcha_model_desc m) {
  const int D = m.d;
  const int c = threadIdx.x;
  if (c >= m.n_chrom) return;
  GemmGroup* t0 = const_cast<GemmGroup*>(table_ptr(m.derived, D, TAB_ENC0));
  GemmGroup* t1 = const_cast<GemmGroup*>(table_ptr(m.derived, D, TAB_ENC1));
  GemmGroup* t2 = const_cast<GemmGroup*>(table_ptr(m.derived, D, TAB_GW1));
  GemmGroup* t3 = const_cast<GemmGroup*>(table_ptr(m.derived, D, TAB_GW0));
  const int32_t n_c = (int32_t)(m.chrom_end[c] - m.chrom_start[c]);
  GemmGroup g;
  g.pad = 0;
  // enc0: H0 = tanh(drop(F_c[id - start]) W0_c^T)
  g.A = m.feat[c]; g.lda = m.feat_ld[c]; g.a_id_off = m.chrom_start[c];
  g.B = m.params + m.off_w0[c]; g.ldb = n_c; g.C = nullptr; g.ldc = 0; g.dim = n_c;
  t0[c] = g;
  // enc1: E = H0 W1_c^T   (and, read as [K, N], dH0 = dE W1_c)
  g.A = nullptr; g.lda = 0; g.a_id_off = 0; g.B = m.params + m.off_w1[c]; g.ldb = D; g.dim = D;
  t1[c] = g;
  if (m.grads) {
    // dW1_c += dE^T H0
    g.A = nullptr; g.B = nullptr; g.ldb = 0; g.C = m.grads + m.off_w1[c]; g.ldc = D; g.dim = D;
    t2[c] = g;
    // dW0_c += dH0pre^T drop(F_c[id - start])
    g.B = m.feat[c]; g.ldb = m.feat_ld[c]; g.a_id_off = m.chrom_start[c];
    g.C = m.grads + m.off_w0[c]; g.ldc = n_c; g.dim = n_c;
    t3[c] = g;
  }
}

// Side stream of the backward pass: the fold of the derived-weight gradients back onto the reference's parameters only
// needs the attention backward, so it runs under the row-chain / encoder backward kernels that follow (fork after the
// attention backward, join before matcha_backward returns: stream-ordered on the caller's stream as before).
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  int device = -1;
};
static int side_stream(SideStream** out) {
  static SideStream g[16];
  int dev = 0;
  if (int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return rc;
  if (dev < 0 || dev >= 16) { set_error("side_stream: device index %d out of range", dev); return MATCHA_ERR_ARG; }
  SideStream& ss = g[dev];
  if (ss.stream == nullptr) {
    if (int rc = check_cuda(cudaStreamCreateWithFlags(&ss.stream, cudaStreamNonBlocking), "cudaStreamCreate")) return rc;
    if (int rc = check_cuda(cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming), "cudaEventCreate")) return rc;
    if (int rc = check_cuda(cudaEventCreateWithFlags(&ss.join, cudaEventDisableTiming), "cudaEventCreate")) return rc;
    ss.device = dev;
  }
  *out = &ss;
  return MATCHA_OK;
}

// matcha_prepare may be issued on a different stream than the passes that consume the derived weights (it depends on the
// weights only, so the trainer runs it beside batch assembly).  It records two events: PREP_ENCODER once everything the
// node encoder reads is queued (grouped-contraction tables, pre-split encoder weights -- queued first), PREP_ALL at its
// end.  A pass waits for the first one AFTER its weight-free token bucketing and for the second one only before the
// attribute mix, so the attention / row-chain folds run beside the encoder and the reconstruction head.
enum { PREP_ENCODER = 0, PREP_ALL = 1 };
static int prepare_event(int which, cudaEvent_t* out) {
  static cudaEvent_t g[16][2] = {};
  int dev = 0;
  if (int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return rc;
  if (dev < 0 || dev >= 16) { set_error("prepare_event: device index %d out of range", dev); return MATCHA_ERR_ARG; }
  if (g[dev][which] == nullptr)
    if (int rc = check_cuda(cudaEventCreateWithFlags(&g[dev][which], cudaEventDisableTiming), "cudaEventCreate")) return rc;
  *out = g[dev][which];
  return MATCHA_OK;
}
static int wait_prepared(int which, cudaStream_t s) {
  cudaEvent_t ev;
  if (int rc = prepare_event(which, &ev)) return rc;
  return check_cuda(cudaStreamWaitEvent(s, ev, 0), "cudaStreamWaitEvent");
}

// derived-parameter gradients -> gradients of the reference's own parameters
__global__ void prep_bwd_qk_kernel(const matcha_model_desc m) {
  const int D = m.d;
  const DerivedLayout l = derived_layout(D);
  const float* P = m.params; float* G = m.grads;
  const float* dW = m.derived_grad + l.wqkg;
  const float* dbq = m.derived_grad + l.bqkg;
  const float inv = rsqrtf((float)D);
  // one block per column c (D blocks), threads stride over the H*D rows
  const int c = blockIdx.x, tid = threadIdx.x;
  const float gq = P[m.off_lnq_g + c], bq = P[m.off_lnq_b + c], gk = P[m.off_lnk_g + c];
  float a_gq = 0.f, a_bq = 0.f, a_gk = 0.f;
  for (int r = tid; r < kH * D; r += blockDim.x) {
    const float wq = P[m.off_wq + (int64_t)r * D + c], dwq = dW[(int64_t)r * D + c], dcq = dbq[r];
    atomicAdd(&G[m.off_wq + (int64_t)r * D + c], (dwq * gq + dcq * bq) * inv);
    a_gq = fmaf(dwq, wq * inv, a_gq);
    a_bq = fmaf(dcq, wq * inv, a_bq);
    const float wk = P[m.off_wk + (int64_t)r * D + c], dwk = dW[(int64_t)(kH * D + r) * D + c];
    atomicAdd(&G[m.off_wk + (int64_t)r * D + c], dwk * gk);
    a_gk = fmaf(dwk, wk, a_gk);
  }
  __shared__ float red[3][32];
  for (int o = 16; o > 0; o >>= 1) {
    a_gq += __shfl_xor_sync(0xffffffffu, a_gq, o);
    a_bq += __shfl_xor_sync(0xffffffffu, a_bq, o);
    a_gk += __shfl_xor_sync(0xffffffffu, a_gk, o);
  }
  if ((tid & 31) == 0) { red[0][tid >> 5] = a_gq; red[1][tid >> 5] = a_bq; red[2][tid >> 5] = a_gk; }
  __syncthreads();
  if (tid == 0) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { s0 += red[0][i]; s1 += red[1][i]; s2 += red[2][i]; }
    atomicAdd(&G[m.off_lnq_g + c], s0);
    atomicAdd(&G[m.off_lnq_b + c], s1);
    atomicAdd(&G[m.off_lnk_g + c], s2);   // layer_norm2.bias (key side): exactly zero gradient
  }
}
// block = (head h, row a of the head): thread b.  With M_h = fc1_h^T dWg_h, u = fc1_h^T db_dyn, z = Wv_h b_v:
//   dWv_h[a][b] += M_h[a][b] g_v[b] + u[a] b_v[b];   dg_v[b] += M_h[a][b] Wv_h[a][b];   db_v[b] += u[a] Wv_h[a][b]
//   dfc1[a][h*D + b] += sum_c dWg_h[a][c] g_v[c] Wv_h[b][c] + db_dyn[a] z[b]
__global__ void prep_bwd_g_kernel(const matcha_model_desc m) {
  const int D = m.d;
  const DerivedLayout l = derived_layout(D);
  const float* P = m.params; float* G = m.grads;
  const int h = blockIdx.x / D, a = blockIdx.x % D, b = threadIdx.x;      // blockDim.x == D
  const float* dWg = m.derived_grad + l.wqkg + (int64_t)(2 * kH * D + h * D) * D;
  const float* F = P + m.off_fc1_w + h * D;                   // fc1_h[o][mm] = F[o * H*D + mm]
  const float* V = P + m.off_wv + (int64_t)(h * D) * D;       // Wv_h[mm][c]
  const float* dbd = m.derived_grad + l.bdyn;
  __shared__ float s_dw[128], s_gv[128];                      // dWg_h[a][:], g_v
  s_dw[b] = dWg[(int64_t)a * D + b];
  s_gv[b] = P[m.off_lnv_g + b];
  __syncthreads();
  float mh = 0.f, u = 0.f;
  for (int o = 0; o < D; ++o) {
    const float f = __ldg(F + (int64_t)o * (kH * D) + a);
    mh = fmaf(f, __ldg(dWg + (int64_t)o * D + b), mh);
    u = fmaf(f, __ldg(dbd + o), u);
  }
  const float gv = s_gv[b], bv = P[m.off_lnv_b + b], vab = V[(int64_t)a * D + b];
  atomicAdd(&G[m.off_wv + (int64_t)(h * D + a) * D + b], mh * gv + u * bv);
  atomicAdd(&G[m.off_lnv_g + b], mh * vab);
  atomicAdd(&G[m.off_lnv_b + b], u * vab);
  float df = 0.f, z = 0.f;
  for (int c = 0; c < D; ++c) {
    const float vbc = __ldg(V + (int64_t)b * D + c);
    df = fmaf(s_dw[c] * s_gv[c], vbc, df);
    z = fmaf(vbc, P[m.off_lnv_b + c], z);
  }
  atomicAdd(&G[m.off_fc1_w + (int64_t)a * (kH * D) + h * D + b], df + __ldg(dbd + a) * z);
  if (blockIdx.x < D && b == 0) atomicAdd(&G[m.off_fc1_b + blockIdx.x], __ldg(dbd + blockIdx.x));
}

// ------------------------------------------------------------------------------------------
// workspace carving (identical on the sizing and the execution path)
// ------------------------------------------------------------------------------------------
struct Workspace {
  int32_t *counts, *group_off, *cursor, *perm;
  float *H0, *E, *V0, *X, *xhat, *rstd, *QKG, *U, *H1d, *H2, *pred, *recon, *probs;
  uint8_t *xhat_t, *dqkg_t;   // MMA-ready tiles (rowwise.cuh); dqkg_t aliases dQKG
  uint8_t *v0_t, *attr_t, *u_t, *h1_t, *dh2_t, *dh1_t, *dp_t, *dv0_t;   // row-chain tiles kept for the weight-gradient kernel
  float* wpair_scratch;
  float *dlogit, *dH2, *dXs, *dH1pre, *dU, *dQKG, *dxhat, *dP, *dV0, *dtE, *dE, *dH0pre, *tc_scratch;
  uint8_t* recon_tiles;    // tanh(E) operand tiles of the eligible tokens (recon_pipe.cu)
  float* recon_stash;      // unscaled dRw [n_r, 64] | drb [n_r] left by the training forward of the fused recon head
  int64_t tc_scratch_floats;
  int64_t pred_ld;
  int64_t bytes;
};
static int64_t max_chrom_len(const matcha_model_desc* m) {
  int64_t mx = 0;
  for (int c = 0; c < m->n_chrom; ++c) mx = mx > (m->chrom_end[c] - m->chrom_start[c]) ? mx : (m->chrom_end[c] - m->chrom_start[c]);
  return mx;
}
static Workspace carve(const matcha_model_desc* m, int64_t B, int L, int training, void* base) {
  Workspace w;
  memset(&w, 0, sizeof(w));
  const int64_t T = B * L;
  const int D = m->d;
  const int64_t QKG = 3 * (int64_t)kH * D;
  const bool tc = D == kD;                       // tensor-core tile buffers exist only for embed_dim 64
  char* p = reinterpret_cast<char*>(base);
  int64_t off = 0;
  auto take = [&](int64_t nbytes) -> void* {
    void* r = p ? (void*)(p + off) : nullptr;
    off += (nbytes + 255) / 256 * 256;
    return r;
  };
  w.counts = (int32_t*)take(sizeof(int32_t) * (MATCHA_MAX_CHROM + 2));
  w.group_off = (int32_t*)take(sizeof(int32_t) * (MATCHA_MAX_CHROM + 2));
  w.cursor = (int32_t*)take(sizeof(int32_t) * (int64_t)bucket_hist_ints());      // per-block histograms of the bucketing kernel
  w.perm = (int32_t*)take(sizeof(int32_t) * (T + 1));
  w.recon = (float*)take(sizeof(float) * 4);
  const int64_t row = sizeof(float) * T * D;
  w.H0 = (float*)take(row); w.E = (float*)take(row); w.V0 = (float*)take(row); w.X = (float*)take(row);
  w.xhat = (float*)take(row); w.rstd = (float*)take(sizeof(float) * (T + 1));
  if (tc) {
    int64_t nt = num_token_tiles(T);
    if (L >= 2 && L <= 6 && num_atiles(T, L) > nt) nt = num_atiles(T, L);
    w.xhat_t = (uint8_t*)take(nt * (int64_t)kXTileBytes);
  }
  w.QKG = (float*)take(sizeof(float) * T * QKG);
  w.U = (float*)take(row); w.H1d = (float*)take(row); w.H2 = (float*)take(row);
  w.pred_ld = m->inter ? (max_chrom_len(m) + 3) / 4 * 4 : 0;
  w.pred = (float*)take(sizeof(float) * T * (w.pred_ld > 0 ? w.pred_ld : 1));
  if (training) {
    w.dlogit = (float*)take(sizeof(float) * (B + 1));
    w.dH2 = (float*)take(row); w.dXs = (float*)take(row); w.dH1pre = (float*)take(row); w.dU = (float*)take(row);
    w.dQKG = (float*)take(tc ? num_token_tiles(T) * (int64_t)kGChunks * kGTileBytes      // >= T * 1536 * 4 bytes
                             : (int64_t)sizeof(float) * T * QKG);
    w.dqkg_t = reinterpret_cast<uint8_t*>(w.dQKG);
    w.dxhat = (float*)take(row); w.dP = (float*)take(row); w.dV0 = (float*)take(row); w.dtE = (float*)take(row);
    w.dE = (float*)take(row); w.dH0pre = (float*)take(row);
    w.probs = (float*)take(sizeof(float) * T * kH * 8);       // attention weights kept by the fused forward
    if (tc && L >= 2 && L <= 6) {
      const int64_t nat = num_atiles(T, L);
      w.v0_t = (uint8_t*)take(nat * 32768); w.u_t = (uint8_t*)take(nat * 32768); w.h1_t = (uint8_t*)take(nat * 32768);
      w.dh2_t = (uint8_t*)take(nat * 32768); w.dh1_t = (uint8_t*)take(nat * 32768); w.dp_t = (uint8_t*)take(nat * 32768);
      w.dv0_t = (uint8_t*)take(nat * 32768); w.attr_t = (uint8_t*)take(nat * 16384);
      w.wpair_scratch = (float*)take(sizeof(float) * wgrad_pair_scratch_floats());
    }
    if (m->inter) w.recon_stash = (float*)take(sizeof(float) * max_chrom_len(m) * (D + 1));
    if (m->inter && tc) w.recon_tiles = (uint8_t*)take(recon_pipe_tile_bytes(T));
    w.tc_scratch_floats = gemm_tc_scratch_floats(QKG);
    w.tc_scratch = (float*)take(sizeof(float) * w.tc_scratch_floats);
  }
  w.bytes = off;
  return w;
}

static int validate(const matcha_model_desc* m) {
  MATCHA_REQUIRE(m != nullptr, "model descriptor is NULL");
  if ((m->d != 64 && m->d != 128) || m->n_head != kH) {
    set_error("this build supports embed_dim 64 (tensor-core path) or 128 (fp32 SIMT path) with n_head=%d (got %d, %d)", kH, m->d,
              m->n_head);
    return MATCHA_ERR_UNSUPPORTED;
  }
  MATCHA_REQUIRE(m->n_chrom >= 1 && m->n_chrom <= MATCHA_MAX_CHROM, "n_chrom=%d out of range", m->n_chrom);
  g_tcg_model = m->d != kD;
  MATCHA_REQUIRE(m->params && m->derived, "params / derived buffers missing");
  MATCHA_REQUIRE(m->attr_dim >= 1 && m->attr_table, "attribute table missing");
  const bool csr = model_uses_csr(m);
  if (csr && m->d != kD) { set_error("CSR feature rows need embed_dim %d", kD); return MATCHA_ERR_UNSUPPORTED; }
  for (int c = 0; c < m->n_chrom; ++c) {
    const bool has_csr = m->feat_indptr[c] && m->feat_indices[c] && m->feat_values[c];
    if (csr ? (m->feat[c] || !has_csr) : !m->feat[c]) {
      set_error("chromosome %d: feature rows must be all dense or all CSR", c);
      return MATCHA_ERR_UNSUPPORTED;
    }
  }
  return MATCHA_OK;
}

static GemmDesc gemm_base(int form, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
                          int64_t ldb, float* C, int64_t ldc) {
  GemmDesc d;
  memset(&d, 0, sizeof(d));
  d.form = form; d.M = M; d.N = N; d.K = K; d.A = A; d.lda = lda; d.B = B; d.ldb = ldb; d.C = C; d.ldc = ldc;
  d.out_scale = 1.f;
  return d;
}

static ChromMeta chrom_meta(const matcha_model_desc* m) {
  ChromMeta cm;
  cm.n = m->n_chrom;
  for (int c = 0; c < m->n_chrom; ++c) { cm.start[c] = m->chrom_start[c]; cm.end[c] = m->chrom_end[c]; }
  return cm;
}

// encoder: bucket tokens, two grouped contractions.  Outputs H0, E (zero rows for pads)
static int run_encoder(const matcha_model_desc* m, const int64_t* x, int64_t T, int training, uint64_t seed,
                       const Workspace& w, float* E_out, cudaStream_t s) {
  const int Dm = m->d;
  const int64_t QKGm = 3 * (int64_t)kH * Dm;
  (void)QKGm;
  int rc;
  if ((rc = PROF(P_BUCKET, 1, launch_bucket(x, T, chrom_meta(m), w.counts, w.group_off, w.cursor, w.perm, s)))) return rc;
  if ((rc = wait_prepared(PREP_ENCODER, s))) return rc;     // the most recent matcha_prepare (possibly queued on another stream)
  // E rows of pad tokens are read (the attribute mix walks every token): zero them.  H0 is only ever read through the
  // chromosome-bucketed token lists (real tokens), so its pad rows need no fill.
  if ((rc = check_cuda(cudaMemsetAsync(E_out, 0, sizeof(float) * T * Dm, s), "memset E"))) return rc;
  if (use_enc_tc(m, T)) {    // both encoder layers in one tcgen05 kernel over the bucketed token list
    const DropCfg fd = make_drop(seed, SITE_FEATURE, m->p_feature, training != 0);
    return PROF(P_ENC0, 1, launch_enc_tc_fwd(m, derived_layout(m->d).total, x, T, w.perm, w.group_off, w.H0, E_out, fd, s));
  }
  if (model_uses_csr(m)) {
    const DropCfg fd = make_drop(seed, SITE_FEATURE, m->p_feature, training != 0);
    if ((rc = PROF(P_ENC0, 1, launch_enc0_csr_fwd(m, derived_layout(m->d).total, x, T, w.perm, w.group_off, w.H0, fd, s)))) return rc;
  } else {
    GemmDesc d = gemm_base(FORM_NT, 0, Dm, 0, nullptr, 0, nullptr, 0, w.H0, Dm);
    d.perm = w.perm; d.a_ids = x; d.ngroups = m->n_chrom; d.groups = table_ptr(m->derived, m->d, TAB_ENC0);
    d.group_off = w.group_off; d.total_rows = T; d.epi_act = 1;
    if (training && m->p_feature > 0.f) { d.drop_on = 1; d.drop = make_drop(seed, SITE_FEATURE, m->p_feature, true); }
    if ((rc = run_gemm(d, s, P_ENC0))) return rc;
  }
  GemmDesc e = gemm_base(FORM_NT, 0, Dm, Dm, w.H0, Dm, nullptr, 0, E_out, Dm);
  e.perm = w.perm; e.ngroups = m->n_chrom; e.groups = table_ptr(m->derived, m->d, TAB_ENC1);
  e.group_off = w.group_off; e.total_rows = T;
  return run_gemm(e, s, P_ENC1);
}

// X = tanh(next_w(E + attribute_nn(attr[id]))), xhat, rstd, QKG
static int run_mix_qkg(const matcha_model_desc* m, const int64_t* x, int64_t T, const Workspace& w, cudaStream_t s,
                       int fused_L = 0, int training = 0) {
  const int Dm = m->d;
  const int64_t QKGm = 3 * (int64_t)kH * Dm;
  (void)QKGm;
  int rc;
  const float* P = m->params;
  const DerivedLayout l = derived_layout(m->d);
  if ((rc = wait_prepared(PREP_ALL, s))) return rc;          // attention / row-chain folds of the most recent matcha_prepare
  if (fused_L > 0 && use_chain(m, T, fused_L))     // attribute mix + next_w + LayerNorm statistics + tiles in one kernel
    return PROF(P_MIX, 1, launch_chain_mix_fwd(w.E, x, m->attr_table, m->attr_dim, P + m->off_attr_w, P + m->off_attr_b,
                                               m->derived + l.wchain + CW_NEXT_K * (kChainWBytes / 4), P + m->off_next_b, nullptr,
                                               w.X, w.xhat, w.rstd, w.xhat_t, training ? w.v0_t : nullptr, training ? w.attr_t : nullptr,
                                               T / fused_L, fused_L, s));
  GemmDesc a = gemm_base(FORM_NT, T, Dm, m->attr_dim, m->attr_table, m->attr_dim, P + m->off_attr_w, m->attr_dim, w.V0, Dm);
  a.a_ids = x; a.bias = P + m->off_attr_b; a.addend = w.E; a.ld_add = Dm;
  if ((rc = run_gemm(a, s, P_ATTR))) return rc;
  GemmDesc b = gemm_base(FORM_NT, T, Dm, Dm, w.V0, Dm, P + m->off_next_w, Dm, w.X, Dm);
  b.bias = P + m->off_next_b; b.epi_act = 1;
  if ((rc = run_gemm(b, s, P_MIX))) return rc;
  if (fused_L > 0)   // fused attention: hyperedge-aligned tiles; the QKG projection happens inside the attention kernel
    return PROF(P_LN, 1, launch_ln_fwd_atiles(w.X, w.xhat, w.rstd, T, fused_L, w.xhat_t, s));
  const bool tiles = use_tiles(m, T);
  if (m->d == kD && gemm_impl() == 1 &&
      (rc = launch_split_weights_k64(m->derived + l.wqkg, kD, kQKG, m->derived + l.wsplit, s))) return rc;   // decomposed pipeline only
  if ((rc = PROF(P_LN, 1, launch_ln_fwd(m->d, w.X, w.xhat, w.rstd, T, tiles ? w.xhat_t : nullptr, s)))) return rc;
  if (tiles)
    return PROF(P_QKG, 1, tc_qkg_forward_tiles(w.xhat_t, reinterpret_cast<const uint8_t*>(m->derived + l.wsplit),
                                               m->derived + l.bqkg, w.QKG, T, s));
  GemmDesc q = gemm_base(FORM_NT, T, QKGm, Dm, w.xhat, Dm, m->derived + l.wqkg, Dm, w.QKG, QKGm);
  q.bias = m->derived + l.bqkg;
  q.b_split = reinterpret_cast<const uint8_t*>(m->derived + l.wsplit);
  return run_gemm(q, s, P_QKG);
}

// pff_n1 (two 1x1 convolutions with residual), input U, output H2 (pre-LayerNorm)
static int run_pff(const matcha_model_desc* m, int64_t T, int training, uint64_t seed, const Workspace& w, cudaStream_t s) {
  const int Dm = m->d;
  const int64_t QKGm = 3 * (int64_t)kH * Dm;
  (void)QKGm;
  int rc;
  const float* P = m->params;
  GemmDesc a = gemm_base(FORM_NT, T, Dm, Dm, w.U, Dm, P + m->off_pff_w0, Dm, w.H1d, Dm);
  a.bias = P + m->off_pff_b0; a.epi_act = 1;
  if (training && m->p_pff > 0.f) { a.epi_drop = 1; a.edrop = make_drop(seed, SITE_PFF, m->p_pff, true); }
  if ((rc = run_gemm(a, s, P_PFF0))) return rc;
  GemmDesc b = gemm_base(FORM_NT, T, Dm, Dm, w.H1d, Dm, P + m->off_pff_w1, Dm, w.H2, Dm);
  b.bias = P + m->off_pff_b1; b.addend = w.U; b.ld_add = Dm;
  return run_gemm(b, s, P_PFF1);
}

static ScoreParams score_params(const matcha_model_desc* m) {
  const float* P = m->params;
  ScoreParams p;
  p.pff_g = P + m->off_pff_g; p.pff_b = P + m->off_pff_b; p.ln1_g = P + m->off_ln1_g; p.ln1_b = P + m->off_ln1_b;
  p.ln2_g = P + m->off_ln2_g; p.ln2_b = P + m->off_ln2_b; p.cls_w = P + m->off_cls_w; p.cls_b = P + m->off_cls_b;
  return p;
}

}  // namespace matcha

using namespace matcha;

// ==========================================================================================
// C ABI
// ==========================================================================================
extern "C" {

const char* matcha_last_error(void) { return g_err; }

void matcha_profile_enable(int32_t on) { g_prof.on = on != 0; }
int32_t matcha_profile_labels(void) { return P_COUNT; }
const char* matcha_profile_label_name(int32_t i) { return (i >= 0 && i < P_COUNT) ? kProfNames[i] : ""; }
int matcha_profile_read(float* ms, int64_t* calls, int64_t* kernels, int32_t n) {
  MATCHA_REQUIRE(ms && calls && kernels && n >= P_COUNT, "matcha_profile_read: need %d slots", (int)P_COUNT);
  for (int i = 0; i < n; ++i) { ms[i] = 0.f; calls[i] = 0; kernels[i] = 0; }
  for (auto& r : g_prof.recs) {
    if (int rc = check_cuda(cudaEventSynchronize(r.e), "profile sync")) return rc;
    float t = 0.f;
    if (int rc = check_cuda(cudaEventElapsedTime(&t, r.b, r.e), "profile elapsed")) return rc;
    ms[r.label] += t;
    calls[r.label] += 1;
    g_prof.free_events.push_back(r.b);
    g_prof.free_events.push_back(r.e);
  }
  g_prof.recs.clear();
  for (int i = 0; i < P_COUNT; ++i) { kernels[i] = g_prof.kernels[i]; g_prof.kernels[i] = 0; }
  return MATCHA_OK;
}
void matcha_set_gemm_impl(int32_t impl) { g_gemm_impl = impl; }
void matcha_set_fused(int32_t on) { g_fused = on != 0; }
void matcha_set_chain(int32_t on) { g_chain = on != 0; }
void matcha_set_recon_tc(int32_t on) { g_recon_tc = on != 0; }
void matcha_set_recon_pipe(int32_t on) { g_recon_pipe = on != 0; }
void matcha_set_gemm_tcg(int32_t on) { g_tcg = on != 0; }
void matcha_set_enc_pipe(int32_t fwd, int32_t bwd) { set_enc_pipe(fwd, bwd); }
void matcha_set_enc_tc(int32_t on) { g_enc_tc = on != 0; }
void matcha_set_xform(int32_t on) { g_xform = on != 0; }
void matcha_set_mma_passes(int32_t passes) { g_passes = passes == 1 ? 1 : 3; }
int matcha_version(void) { return 100; }

// CSR models keep W0T_c [n_c, 64] (and, in derived_grad, its gradient) after the fixed-size part
static int64_t w0t_floats(const matcha_model_desc* m) {
  const int Dm = m->d;
  const int64_t QKGm = 3 * (int64_t)kH * Dm;
  (void)QKGm;
  if (!m || !model_uses_csr(m)) return 0;
  int64_t n = 0;
  for (int c = 0; c < m->n_chrom; ++c) n += (m->chrom_end[c] - m->chrom_start[c]) * Dm;
  return n;
}
// after the fixed-size part: W0T copies (CSR models) or the pre-split encoder weight chunks (dense rows, embed_dim 64)
int64_t matcha_derived_elems(const matcha_model_desc* m) {
  return derived_layout(m->d).total + w0t_floats(m) + (enc_tc_eligible(m) ? enc_tc_split_floats(m) : 0);
}

int64_t matcha_workspace_bytes(const matcha_model_desc* m, int64_t B, int32_t L, int32_t training) {
  if (!m || B < 0 || L < 1) return -1;
  return carve(m, B, L, training, nullptr).bytes + 256;
}

int matcha_prepare(const matcha_model_desc* m, void* stream) {
  int rc = validate(m);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const DerivedLayout l = derived_layout(m->d);
  (void)l;
  prof_begin(P_PREP, s);
  cudaEvent_t ev;
  // (1) what the node encoder reads: queued first, so a pass can start its encoder while the rest is still folding
  prep_tables_kernel<<<1, MATCHA_MAX_CHROM, 0, s>>>(*m);
  MATCHA_CHECK_LAUNCH("prep_tables");
  if (model_uses_csr(m) && (rc = launch_csr_prepare(m, l.total, s))) return rc;
  if (enc_tc_eligible(m) && (rc = launch_enc_tc_prepare(m, l.total, s))) return rc;
  if ((rc = prepare_event(PREP_ENCODER, &ev))) return rc;
  if ((rc = check_cuda(cudaEventRecord(ev, s), "cudaEventRecord"))) return rc;
  // (2) attention and row-chain folds
  prep_qk_kernel<<<3 * kH * m->d, m->d, 0, s>>>(*m);
  MATCHA_CHECK_LAUNCH("prep_qk");
  prep_g_kernel<<<(kH * m->d * m->d + 255) / 256, 256, 0, s>>>(*m);
  MATCHA_CHECK_LAUNCH("prep_g");
  prep_bdyn_kernel<<<1, 128, 0, s>>>(*m);
  MATCHA_CHECK_LAUNCH("prep_bdyn");
  if (m->d == kD) {      // pre-split bf16 hi | lo operand copies for the tcgen05 kernels (embed_dim 64 only)
    // (the K-major / MN-major copies of the whole W_qkg are only read by the decomposed pipeline: built there, on demand)
    if ((rc = launch_split_w_heads(m->derived + l.wqkg, m->derived + l.wheads, s))) return rc;
    if ((rc = launch_prep_xform(m->derived + l.wqkg, m->derived + l.bqkg, m->derived + l.wx, m->derived + l.vx, s))) return rc;
    if (m->grads && (rc = launch_split_w_pairs(m->derived + l.wqkg, m->derived + l.wpairs, s))) return rc;
    {
      float* wc = m->derived + l.wchain;
      const int64_t st = kChainWBytes / 4;
      if ((rc = launch_split_w64x3(m->params + m->off_next_w, m->params + m->off_pff_w0, m->params + m->off_pff_w1, wc + CW_NEXT_K * st,
                                 wc + CW_NEXT_MN * st, wc + CW_PFF0_K * st, wc + CW_PFF0_MN * st, wc + CW_PFF1_K * st,
                                 wc + CW_PFF1_MN * st, s))) return rc;
    }
  }
  prof_end(P_PREP, 10, s);
  if ((rc = prepare_event(PREP_ALL, &ev))) return rc;
  return check_cuda(cudaEventRecord(ev, s), "cudaEventRecord");
}

int matcha_forward(const matcha_model_desc* m, const int64_t* x, int64_t B, int32_t L, int32_t training,
                   uint64_t seed, int32_t random_chrom, float* logits, float* recon, void* workspace,
                   int64_t workspace_bytes, void* stream) {
  const int Dm = m->d;
  const int64_t QKGm = 3 * (int64_t)kH * Dm;
  (void)QKGm;
  int rc = validate(m);
  if (rc) return rc;
  MATCHA_REQUIRE(x && logits && workspace, "matcha_forward: NULL argument");
  MATCHA_REQUIRE(L >= 2 && L <= 6, "matcha_forward: padded width L=%d unsupported (2..6)", L);
  MATCHA_REQUIRE(random_chrom < m->n_chrom, "random_chrom=%d out of range", random_chrom);
  cudaStream_t s = (cudaStream_t)stream;
  if (B == 0) return MATCHA_OK;
  uintptr_t base = ((uintptr_t)workspace + 255) / 256 * 256;
  Workspace w = carve(m, B, L, training, (void*)base);
  MATCHA_REQUIRE((int64_t)(base - (uintptr_t)workspace) + w.bytes <= workspace_bytes,
                 "workspace too small: need %lld bytes", (long long)(w.bytes + 256));
  const int64_t T = B * L;
  const DerivedLayout l = derived_layout(m->d);

  if ((rc = run_encoder(m, x, T, training, seed, w, w.E, s))) return rc;
  {
    if ((rc = check_cuda(cudaMemsetAsync(w.recon, 0, sizeof(float), s), "memset recon"))) return rc;
    if (random_chrom >= 0 && m->inter) {
      const int64_t rs = m->chrom_start[random_chrom], re = m->chrom_end[random_chrom];
      if (use_recon_tc(m, T, L)) {
        // training: loss AND the (unscaled) gradients in one pass -- the backward pass only applies d loss / d recon
        float* st = training ? w.recon_stash : nullptr;
        if (training) {
          if ((rc = check_cuda(cudaMemsetAsync(st, 0, sizeof(float) * (re - rs) * (Dm + 1), s), "memset recon stash"))) return rc;
          if ((rc = check_cuda(cudaMemsetAsync(w.dtE, 0, sizeof(float) * T * Dm, s), "memset dtE"))) return rc;
        }
        if (training && use_recon_pipe()) {
          if ((rc = PROF(P_RECON_PRED, 2, launch_recon_pipe(w.E, x, T, m->inter, m->inter_ld, rs, re, m->params + m->off_rw[random_chrom],
                                                            m->params + m->off_rb[random_chrom], w.counts, random_chrom, m->n_chrom,
                                                            w.perm, w.group_off, w.recon, st, st + (re - rs) * Dm, w.dtE,
                                                            w.recon_tiles, s)))) return rc;
        } else if ((rc = PROF(P_RECON_PRED, 1, launch_recon_tc(w.E, x, T, m->inter, m->inter_ld, rs, re, m->params + m->off_rw[random_chrom],
                                                        m->params + m->off_rb[random_chrom], w.counts, random_chrom, m->n_chrom,
                                                        w.perm, w.group_off, w.recon, st, training ? st + (re - rs) * Dm : nullptr,
                                                        training ? w.dtE : nullptr, 1.f, training ? 1 : 0, s)))) return rc;
      } else {
        GemmDesc p = gemm_base(FORM_NT, T, re - rs, Dm, w.E, Dm, m->params + m->off_rw[random_chrom], Dm, w.pred, w.pred_ld);
        p.a_act = 1; p.bias = m->params + m->off_rb[random_chrom];
        if ((rc = run_gemm(p, s, P_RECON_PRED))) return rc;
        if ((rc = PROF(P_RECON_DIFF, 1, launch_recon_diff(w.pred, w.pred_ld, x, T, m->inter, m->inter_ld, rs, re, w.counts, random_chrom,
                                    m->n_chrom, w.recon, s)))) return rc;
      }
    }
    if (recon && (rc = check_cuda(cudaMemcpyAsync(recon, w.recon, sizeof(float), cudaMemcpyDeviceToDevice, s), "copy recon")))
      return rc;
  }
  const bool fused = use_fused(m, T, L);
  if ((rc = run_mix_qkg(m, x, T, w, s, fused ? L : 0, training))) return rc;
  DropCfg dattn = make_drop(seed, SITE_ATTN, m->p_attn, training != 0);
  if (fused && use_xform(m, T, L)) {
    if ((rc = PROF(P_ATTN_FWD, 1, launch_attn_xform_fwd(w.xhat_t, w.xhat, reinterpret_cast<const uint8_t*>(m->derived + l.wx),
                                                        m->derived + l.vx, m->derived + l.bdyn, x, w.U, training ? w.probs : nullptr, B, L,
                                                        dattn, mma_passes(), s))))
      return rc;
  } else if (fused) {
    if ((rc = PROF(P_ATTN_FWD, 1, launch_attn_fused_fwd(w.xhat_t, reinterpret_cast<const uint8_t*>(m->derived + l.wheads),
                                                        m->derived + l.bqkg, m->derived + l.bdyn, x, w.U, training ? w.probs : nullptr, B, L,
                                                        dattn, s))))
      return rc;
  } else if ((rc = PROF(P_ATTN_FWD, 1, launch_attn_fwd(m->d, w.QKG, x, m->derived + l.bdyn, w.U, B, L, dattn, s)))) return rc;
  if (fused && use_chain(m, T, L)) {
    const float* wc = m->derived + l.wchain;
    return PROF(P_PFF0, 1, launch_chain_pff_fwd(w.U, w.xhat, x, wc + CW_PFF0_K * (kChainWBytes / 4), wc + CW_PFF1_K * (kChainWBytes / 4),
                                                m->params + m->off_pff_b0, m->params + m->off_pff_b1, score_params(m),
                                                make_drop(seed, SITE_PFF, m->p_pff, training != 0), training ? w.H1d : nullptr,
                                                training ? w.H2 : nullptr, logits, training ? w.u_t : nullptr,
                                                training ? w.h1_t : nullptr, B, L, s));
  }
  if ((rc = run_pff(m, T, training, seed, w, s))) return rc;
  return PROF(P_SCORE_FWD, 1, launch_score_fwd(m->d, w.H2, w.xhat, x, score_params(m), logits, B, L, s));
}

int matcha_bce_loss(const float* logits, const float* y, const float* wgt, int64_t B, float alpha, float beta,
                    const float* recon, float* dlogit, float* loss_out, void* stream) {
  MATCHA_REQUIRE(logits && y && wgt && dlogit && loss_out && B > 0, "matcha_bce_loss: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  int rc;
  if ((rc = check_cuda(cudaMemsetAsync(loss_out, 0, 3 * sizeof(float), s), "memset loss"))) return rc;
  if ((rc = PROF(P_BCE, 1, launch_bce(logits, y, wgt, alpha, dlogit, loss_out, B, s)))) return rc;
  return PROF(P_BCE, 1, launch_finalize_loss(loss_out, recon, alpha, beta, s));
}

int matcha_backward(const matcha_model_desc* m, const int64_t* x, int64_t B, int32_t L, uint64_t seed,
                    int32_t random_chrom, const float* dlogit, float beta, int32_t* active, void* workspace,
                    int64_t workspace_bytes, void* stream) {
  const int Dm = m->d;
  const int64_t QKGm = 3 * (int64_t)kH * Dm;
  (void)QKGm;
  int rc = validate(m);
  if (rc) return rc;
  MATCHA_REQUIRE(m->grads && m->derived_grad, "matcha_backward: grads / derived_grad buffers missing");
  MATCHA_REQUIRE(x && dlogit && workspace, "matcha_backward: NULL argument");
  MATCHA_REQUIRE(L >= 2 && L <= 6, "matcha_backward: padded width L=%d unsupported (2..6)", L);
  cudaStream_t s = (cudaStream_t)stream;
  if (B == 0) return MATCHA_OK;
  uintptr_t base = ((uintptr_t)workspace + 255) / 256 * 256;
  Workspace w = carve(m, B, L, 1, (void*)base);
  MATCHA_REQUIRE((int64_t)(base - (uintptr_t)workspace) + w.bytes <= workspace_bytes,
                 "workspace too small: need %lld bytes", (long long)(w.bytes + 256));
  const int64_t T = B * L;
  const DerivedLayout l = derived_layout(m->d);
  const float* P = m->params;
  float* G = m->grads;
  float* DG = m->derived_grad;
  const bool recon_on = random_chrom >= 0 && m->inter && beta != 0.f;

  if ((rc = check_cuda(cudaMemsetAsync(DG, 0, sizeof(float) * l.tables, s), "memset derived_grad"))) return rc;

  // scorer + pff LayerNorm backward
  ScoreGrads sg;
  sg.pff_g = G + m->off_pff_g; sg.pff_b = G + m->off_pff_b; sg.ln1_g = G + m->off_ln1_g; sg.ln1_b = G + m->off_ln1_b;
  sg.ln2_g = G + m->off_ln2_g; sg.ln2_b = G + m->off_ln2_b; sg.cls_w = G + m->off_cls_w; sg.cls_b = G + m->off_cls_b;
  if ((rc = PROF(P_SCORE_BWD, 1, launch_score_bwd(m->d, w.H2, w.xhat, w.rstd, x, score_params(m), dlogit, w.dH2, w.dXs, sg, B, L, s)))) return rc;

  // pff_n1 backward
  DropCfg dpff = make_drop(seed, SITE_PFF, m->p_pff, true);
  DropCfg dattn = make_drop(seed, SITE_ATTN, m->p_attn, true);
  const bool chain = use_chain(m, T, L);
  const int64_t nat = chain ? num_atiles(T, L) : 0;
  const float* wc = m->derived + l.wchain;
  if (chain) {
    // dH2 -> dH1pre -> masked gradient of the attention output in one tcgen05 row-chain kernel; both weight gradients
    // (and bias gradients) of pff_n1 in one stacked tile kernel
    if ((rc = PROF(P_D_PFF1, 1, launch_chain_pff_bwd(w.dH2, w.H1d, x, wc + CW_PFF1_MN * (kChainWBytes / 4),
                                                     wc + CW_PFF0_MN * (kChainWBytes / 4), dpff, dattn, w.dU, w.dh2_t, w.dh1_t, B,
                                                     L, s)))) return rc;
    if ((rc = PROF(P_W_PFF1, 2, launch_wgrad_pair(w.dh2_t, w.dh1_t, w.h1_t, w.u_t, 8, nat, w.wpair_scratch, G + m->off_pff_w1, Dm,
                                                  Dm, G + m->off_pff_b1, G + m->off_pff_w0, Dm, Dm, G + m->off_pff_b0, s))))
      return rc;
  } else {
    GemmDesc d = gemm_base(FORM_TN, Dm, Dm, T, w.dH2, Dm, w.H1d, Dm, G + m->off_pff_w1, Dm);
    d.colsum = G + m->off_pff_b1; d.colsum_n = Dm;
    if ((rc = run_gemm(d, s, P_W_PFF1))) return rc;
    GemmDesc e = gemm_base(FORM_NN, T, Dm, Dm, w.dH2, Dm, P + m->off_pff_w1, Dm, w.dH1pre, Dm);
    e.epi_act = 2; e.aux = w.H1d; e.ld_aux = Dm;
    if (m->p_pff > 0.f) { e.epi_drop = 1; e.edrop = dpff; }
    if ((rc = run_gemm(e, s, P_D_PFF1))) return rc;
    GemmDesc f = gemm_base(FORM_TN, Dm, Dm, T, w.dH1pre, Dm, w.U, Dm, G + m->off_pff_w0, Dm);
    f.colsum = G + m->off_pff_b0; f.colsum_n = Dm;
    if ((rc = run_gemm(f, s, P_W_PFF0))) return rc;
    GemmDesc g = gemm_base(FORM_NN, T, Dm, Dm, w.dH1pre, Dm, P + m->off_pff_w0, Dm, w.dU, Dm);
    g.addend = w.dH2; g.ld_add = Dm;
    if ((rc = run_gemm(g, s, P_D_PFF0))) return rc;
  }
  // attention backward
  int dx_parts = 1;
  if (use_fused(m, T, L)) {
    // fused path: recompute + attention backward + data / weight gradients of the QKG projection in one kernel; the
    // per-head-pair dxhat partials live in the (otherwise unused) dQKG area
    dx_parts = 4;
    if ((rc = PROF(P_ATTN_BWD, 1, launch_attn_fused_bwd(w.xhat_t, reinterpret_cast<const uint8_t*>(m->derived + l.wpairs),
                                                        m->derived + l.bqkg, x, w.dU, w.probs, w.dQKG, w.tc_scratch,
                                                        DG + l.wqkg, DG + l.bqkg, DG + l.bdyn, B, L, dattn, chain ? 1 : 0, mma_passes(), s)))) return rc;
  } else if (use_tiles(m, T)) {
    // tile path: the attention backward emits dQKG directly as bf16 hi|lo MMA tiles; rows of the last tile beyond T are zero
    const int64_t nt = num_token_tiles(T);
    if (T % kTileTok != 0 &&
        (rc = check_cuda(cudaMemsetAsync(w.dqkg_t + (nt - 1) * (int64_t)kGChunks * kGTileBytes, 0, (size_t)kGChunks * kGTileBytes, s),
                         "memset dQKG tail tile")))
      return rc;
    if ((rc = PROF(P_ATTN_BWD, 1, launch_attn_bwd(m->d, w.QKG, w.dU, x, nullptr, w.dqkg_t, DG + l.bdyn, B, L, dattn, s)))) return rc;
    if ((rc = PROF(P_W_QKG, 2, tc_qkg_wgrad_tiles(w.dqkg_t, w.xhat_t, w.tc_scratch, w.tc_scratch_floats, DG + l.wqkg,
                                                  DG + l.bqkg, kH * Dm, T, s)))) return rc;
    if ((rc = launch_split_wT(m->derived + l.wqkg, m->derived + l.wtsplit, s))) return rc;
    if ((rc = PROF(P_D_QKG, 1, tc_qkg_dgrad_tiles(w.dqkg_t, reinterpret_cast<const uint8_t*>(m->derived + l.wtsplit), w.dxhat,
                                                  T, s)))) return rc;
  } else {
    if ((rc = PROF(P_ATTN_BWD, 1, launch_attn_bwd(m->d, w.QKG, w.dU, x, w.dQKG, nullptr, DG + l.bdyn, B, L, dattn, s)))) return rc;
    GemmDesc d = gemm_base(FORM_TN, QKGm, Dm, T, w.dQKG, QKGm, w.xhat, Dm, DG + l.wqkg, Dm);
    d.colsum = DG + l.bqkg; d.colsum_n = kH * Dm;
    d.scratch = w.tc_scratch; d.scratch_floats = w.tc_scratch_floats;
    if ((rc = run_gemm(d, s, P_W_QKG))) return rc;
    GemmDesc e = gemm_base(FORM_NN, T, Dm, QKGm, w.dQKG, QKGm, m->derived + l.wqkg, Dm, w.dxhat, Dm);
    if ((rc = run_gemm(e, s, P_D_QKG))) return rc;
  }
  // derived -> reference parameters (weights / LayerNorm affine / fc1 of the attention block), on the side stream
  SideStream* side = nullptr;
  if ((rc = side_stream(&side))) return rc;
  if ((rc = check_cuda(cudaEventRecord(side->fork, s), "cudaEventRecord"))) return rc;
  if ((rc = check_cuda(cudaStreamWaitEvent(side->stream, side->fork, 0), "cudaStreamWaitEvent"))) return rc;
  prof_begin(P_PREP_BWD, side->stream);
  prep_bwd_qk_kernel<<<m->d, 256, 0, side->stream>>>(*m);
  MATCHA_CHECK_LAUNCH("prep_bwd_qk");
  prep_bwd_g_kernel<<<kH * m->d, m->d, 0, side->stream>>>(*m);
  MATCHA_CHECK_LAUNCH("prep_bwd_g");
  prof_end(P_PREP_BWD, 2, side->stream);
  if ((rc = check_cuda(cudaEventRecord(side->join, side->stream), "cudaEventRecord"))) return rc;
  // reconstruction head backward (gdiff was left in w.pred by the forward pass, without the beta factor)
  const float* dtE = nullptr;
  if (recon_on) {
    const int64_t rs = m->chrom_start[random_chrom], re = m->chrom_end[random_chrom], nr = re - rs;
    if (use_recon_tc(m, T, L)) {
      // the training forward left gdiff . Rw in w.dtE and the unscaled weight / bias gradients in the stash
      if ((rc = PROF(P_W_RECON, 2, launch_axpy(w.recon_stash, beta, G + m->off_rw[random_chrom], nr * Dm, s)))) return rc;
      if ((rc = launch_axpy(w.recon_stash + nr * Dm, beta, G + m->off_rb[random_chrom], nr, s))) return rc;
    } else {
      GemmDesc d = gemm_base(FORM_TN, nr, Dm, T, w.pred, w.pred_ld, w.E, Dm, G + m->off_rw[random_chrom], Dm);
      d.b_act = 1; d.out_scale = beta; d.colsum = G + m->off_rb[random_chrom]; d.colsum_n = nr;
      if ((rc = run_gemm(d, s, P_W_RECON))) return rc;
      GemmDesc e = gemm_base(FORM_NN, T, Dm, nr, w.pred, w.pred_ld, P + m->off_rw[random_chrom], Dm, w.dtE, Dm);
      if ((rc = run_gemm(e, s, P_D_RECON))) return rc;
    }
    dtE = w.dtE;
  }
  if (chain) {
    // LayerNorm / tanh / next_w backward and the encoder-gradient combine in one row-chain kernel; the next_w and
    // attribute_nn weight / bias gradients in one stacked tile kernel
    if ((rc = PROF(P_LN_BWD, 1, launch_chain_mix_bwd(dx_parts > 1 ? w.dQKG : w.dxhat, dx_parts, T * Dm, w.dXs, w.xhat, w.rstd, w.X,
                                                     dtE, w.E, beta, wc + CW_NEXT_MN * (kChainWBytes / 4), w.dE, w.dp_t, w.dv0_t, B,
                                                     L, s)))) return rc;
    if ((rc = PROF(P_W_NEXT, 2, launch_wgrad_pair(w.dp_t, w.dv0_t, w.v0_t, w.attr_t, 4, nat, w.wpair_scratch, G + m->off_next_w, Dm,
                                                  Dm, G + m->off_next_b, G + m->off_attr_w, m->attr_dim, m->attr_dim,
                                                  G + m->off_attr_b, s)))) return rc;
  } else {
    if ((rc = PROF(P_LN_BWD, 1, launch_ln_tanh_bwd(m->d, dx_parts > 1 ? w.dQKG : w.dxhat, dx_parts, T * Dm, w.dXs, w.xhat, w.rstd, w.X,
                                                   w.dP, T, s)))) return rc;
    GemmDesc d = gemm_base(FORM_TN, Dm, Dm, T, w.dP, Dm, w.V0, Dm, G + m->off_next_w, Dm);
    d.colsum = G + m->off_next_b; d.colsum_n = Dm;
    if ((rc = run_gemm(d, s, P_W_NEXT))) return rc;
    GemmDesc e = gemm_base(FORM_NN, T, Dm, Dm, w.dP, Dm, P + m->off_next_w, Dm, w.dV0, Dm);
    if ((rc = run_gemm(e, s, P_D_NEXT))) return rc;
    GemmDesc f = gemm_base(FORM_TN, Dm, m->attr_dim, T, w.dV0, Dm, m->attr_table, m->attr_dim, G + m->off_attr_w, m->attr_dim);
    f.b_ids = x; f.colsum = G + m->off_attr_b; f.colsum_n = Dm;
    if ((rc = run_gemm(f, s, P_W_ATTR))) return rc;
    if ((rc = PROF(P_ENC_COMBINE, 1, launch_enc_combine_bwd(w.dV0, dtE, w.E, beta, w.dE, T * Dm, s)))) return rc;
  }
  // encoder backward (grouped by chromosome; token lists from the forward pass are still in the workspace)
  if (use_enc_tc(m, T) && enc_tc_bwd_fits(m)) {
    // dH0pre and both weight gradients in one tcgen05 kernel (dH0pre never leaves the SM)
    if ((rc = PROF(P_W_ENC0, 1, launch_enc_tc_bwd(m, l.total, x, T, w.perm, w.group_off, w.dE, w.H0,
                                                  make_drop(seed, SITE_FEATURE, m->p_feature, true), s)))) return rc;
  } else {
    GemmDesc d = gemm_base(FORM_TN, Dm, Dm, 0, w.dE, Dm, w.H0, Dm, nullptr, Dm);
    d.perm = w.perm; d.ngroups = m->n_chrom; d.groups = table_ptr(m->derived, m->d, TAB_GW1); d.group_off = w.group_off;
    d.total_rows = T; d.max_group_dim = Dm;
    if ((rc = run_gemm(d, s, P_W_ENC1))) return rc;
    GemmDesc e = gemm_base(FORM_NN, 0, Dm, Dm, w.dE, Dm, nullptr, 0, w.dH0pre, Dm);
    e.perm = w.perm; e.ngroups = m->n_chrom; e.groups = table_ptr(m->derived, m->d, TAB_ENC1); e.group_off = w.group_off;
    e.total_rows = T; e.epi_act = 2; e.aux = w.H0; e.ld_aux = Dm;
    if ((rc = run_gemm(e, s, P_D_ENC1))) return rc;
    if (model_uses_csr(m)) {
      if ((rc = check_cuda(cudaMemsetAsync(DG + l.total, 0, sizeof(float) * w0t_floats(m), s), "memset dW0T"))) return rc;
      if ((rc = PROF(P_W_ENC0, 2, launch_enc0_csr_wgrad(m, l.total, x, T, w.perm, w.group_off, w.dH0pre,
                                                        make_drop(seed, SITE_FEATURE, m->p_feature, true), s)))) return rc;
    } else {
      GemmDesc f = gemm_base(FORM_TN, Dm, 0, 0, w.dH0pre, Dm, nullptr, 0, nullptr, 0);
      f.perm = w.perm; f.b_ids = x; f.ngroups = m->n_chrom; f.groups = table_ptr(m->derived, m->d, TAB_GW0);
      f.group_off = w.group_off; f.total_rows = T; f.max_group_dim = max_chrom_len(m);
      if (m->p_feature > 0.f) { f.drop_on = 2; f.drop = make_drop(seed, SITE_FEATURE, m->p_feature, true); }
      if ((rc = run_gemm(f, s, P_W_ENC0))) return rc;
    }
  }
  if ((rc = check_cuda(cudaStreamWaitEvent(s, side->join, 0), "cudaStreamWaitEvent"))) return rc;     // join
  if (active) {
    if ((rc = PROF(P_MISC, 1, launch_active_flags(w.counts, m->n_chrom, recon_on ? random_chrom : -1, T, active, s)))) return rc;
  }
  return MATCHA_OK;
}

int matcha_node_embeddings(const matcha_model_desc* m, const int64_t* ids, int64_t T, float* out, void* workspace,
                           int64_t workspace_bytes, void* stream) {
  int rc = validate(m);
  if (rc) return rc;
  MATCHA_REQUIRE(ids && out && workspace, "matcha_node_embeddings: NULL argument");
  if (T == 0) return MATCHA_OK;
  uintptr_t base = ((uintptr_t)workspace + 255) / 256 * 256;
  Workspace w = carve(m, T, 1, 0, (void*)base);
  MATCHA_REQUIRE((int64_t)(base - (uintptr_t)workspace) + w.bytes <= workspace_bytes,
                 "workspace too small: need %lld bytes", (long long)(w.bytes + 256));
  return run_encoder(m, ids, T, 0, 0, w, out, (cudaStream_t)stream);
}

int matcha_pair_tables(const matcha_model_desc* m, float* D, float* S, void* workspace, int64_t workspace_bytes,
                       void* stream) {
  const int Dm = m->d;
  const int64_t QKGm = 3 * (int64_t)kH * Dm;
  (void)QKGm;
  int rc = validate(m);
  if (rc) return rc;
  MATCHA_REQUIRE(D && S && workspace, "matcha_pair_tables: NULL argument");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t T = m->n_nodes + 1;
  uintptr_t base = ((uintptr_t)workspace + 255) / 256 * 256;
  Workspace w = carve(m, T, 1, 0, (void*)base);
  MATCHA_REQUIRE((int64_t)(base - (uintptr_t)workspace) + w.bytes <= workspace_bytes,
                 "workspace too small: need %lld bytes", (long long)(w.bytes + 256));
  // node ids 0..N are materialised in the H1d area (T*64 floats >= T int64); they are last read by
  // run_mix_qkg, before run_pff overwrites H1d
  int64_t* ids = reinterpret_cast<int64_t*>(w.H1d);
  if ((rc = PROF(P_MISC, 1, launch_iota_i64(ids, T, s)))) return rc;
  const DerivedLayout l = derived_layout(m->d);
  if ((rc = run_encoder(m, ids, T, 0, 0, w, w.E, s))) return rc;
  if ((rc = run_mix_qkg(m, ids, T, w, s))) return rc;
  if ((rc = PROF(P_MISC, 1, launch_pair_u(m->d, w.QKG, m->derived + l.bdyn, w.U, T, s)))) return rc;
  if ((rc = run_pff(m, T, 0, 0, w, s))) return rc;
  return PROF(P_MISC, 1, launch_pair_ds(m->d, w.H2, w.xhat, score_params(m), D, S, T, s));
}

int64_t matcha_gemm_scratch_floats(int64_t M) { return gemm_tc_scratch_floats(M); }

int matcha_gemm(int32_t form, int32_t impl, const float* A, const float* B, float* C, const float* bias, int64_t M,
                int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc, float* scratch, int64_t scratch_floats,
                void* stream) {
  MATCHA_REQUIRE(form >= 0 && form <= 2 && A && B && C, "matcha_gemm: bad arguments");
  GemmDesc d = gemm_base(form, M, N, K, A, lda, B, ldb, C, ldc);
  d.bias = bias;
  d.scratch = scratch; d.scratch_floats = scratch_floats;
  if (impl == 1 && form == 0 && K == 64 && N % 128 == 0 && scratch && scratch_floats >= N * 64 && ldb % 4 == 0) {
    if (int rc = launch_split_weights_k64(B, ldb, N, scratch, (cudaStream_t)stream)) return rc;   // v2 kernel path
    d.b_split = reinterpret_cast<const uint8_t*>(scratch);
  }
  if (impl == 2) {      // general tcgen05 kernel (gemm_tcg.cu)
    bool handled = false;
    int rc = launch_gemm_tcg(d, (cudaStream_t)stream, &handled);
    if (rc) return rc;
    if (!handled) { set_error("matcha_gemm: problem too small for the general tcgen05 kernel"); return MATCHA_ERR_UNSUPPORTED; }
    return MATCHA_OK;
  }
  if (impl == 1) {
    bool handled = false;
    int rc = launch_gemm_tc(d, (cudaStream_t)stream, &handled);
    if (rc) return rc;
    if (!handled) { set_error("matcha_gemm: shape not eligible for the tcgen05 path"); return MATCHA_ERR_UNSUPPORTED; }
    return MATCHA_OK;
  }
  return launch_gemm_simt(d, (cudaStream_t)stream);
}

}  // extern "C"
