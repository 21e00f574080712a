// fp32 SIMT contraction kernel: the general path for every dense product of the Hyper-SAGNN hot path
// (gathered / grouped / ragged shapes, fused epilogues).  The tcgen05 kernel in gemm_tc.cu takes over
// the large regular shapes; this one remains the fall-through for ragged K/N and is the on-device
// cross-check for it.
//
// 64x64 output tile, 256 threads, 4x4 register micro-tile, K step 16, register prefetch of the next
// K tile.  Operand tiles are stored k-major in shared memory so the inner loop is two LDS.128 per
// 16 FFMA.
#include "common.cuh"

namespace matcha {

namespace {
constexpr int TM = 64, TN = 64, TK = 16, NTHREADS = 256, PAD = 4;
constexpr int KC_GROUPED = 512;  // most tokens per CTA in grouped TN (weight-gradient) launches

__device__ __forceinline__ float4 ld4_guard(const float* __restrict__ p, int64_t i, int64_t n) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p == nullptr || i >= n) return v;
  const float* q = p + i;
  if (i + 3 < n && ((reinterpret_cast<uintptr_t>(q) & 15) == 0)) return __ldg(reinterpret_cast<const float4*>(q));
  v.x = __ldg(q);
  if (i + 1 < n) v.y = __ldg(q + 1);
  if (i + 2 < n) v.z = __ldg(q + 2);
  if (i + 3 < n) v.w = __ldg(q + 3);
  return v;
}
__device__ __forceinline__ float4 tanh4(float4 v) {
  return make_float4(tanhf(v.x), tanhf(v.y), tanhf(v.z), tanhf(v.w));
}

__global__ void __launch_bounds__(NTHREADS) gemm_simt_kernel(const GemmDesc d, const int64_t kc) {
  __shared__ __align__(16) float As[TK][TM + PAD];
  __shared__ __align__(16) float Bs[TK][TN + PAD];

  const int tid = threadIdx.x;
  const bool tn = d.form == FORM_TN;
  const bool b_rowk = d.form == FORM_NT;  // B given as [N, K]

  // ---- resolve this CTA's problem -------------------------------------------------------
  int64_t Mg = d.M, Ng = d.N, Kg = d.K;
  const float* A = d.A; const float* B = d.B; float* C = d.C;
  int64_t lda = d.lda, ldb = d.ldb, ldc = d.ldc;
  int64_t a_id_off = d.a_id_off, b_id_off = d.b_id_off;
  int64_t tok_off = 0, m0, n0 = (int64_t)blockIdx.x * TN, k_begin = 0, k_end;
  if (d.ngroups > 0) {
    const int64_t unit = tn ? kc : TM;
    int64_t tile = tn ? blockIdx.z : blockIdx.y, before = 0;
    int g = 0, cnt = 0;
    for (; g < d.ngroups; ++g) {
      cnt = d.group_off[g + 1] - d.group_off[g];
      int64_t nt = (cnt + unit - 1) / unit;
      if (tile < before + nt) break;
      before += nt;
    }
    if (g == d.ngroups) return;
    const GemmGroup gr = d.groups[g];
    tok_off = d.group_off[g];
    if (tn) {
      Kg = cnt; Ng = gr.dim; C = gr.C; ldc = gr.ldc;
      if (gr.B) { B = gr.B; ldb = gr.ldb; b_id_off = gr.a_id_off; }
      m0 = (int64_t)blockIdx.y * TM;
      k_begin = (tile - before) * kc;
      k_end = min((int64_t)cnt, k_begin + kc);
      if (n0 >= Ng) return;
    } else {
      Mg = cnt; Kg = gr.dim;
      if (gr.A) { A = gr.A; lda = gr.lda; a_id_off = gr.a_id_off; }
      B = gr.B; ldb = gr.ldb;
      m0 = (tile - before) * TM;
      k_end = Kg;
    }
  } else {
    m0 = (int64_t)blockIdx.y * TM;
    if (tn) { k_begin = (int64_t)blockIdx.z * kc; k_end = min(Kg, k_begin + kc); if (k_begin >= k_end) return; }
    else k_end = Kg;
  }

  // ---- per-thread load coordinates ------------------------------------------------------
  // row-k operands ([rows, K], k contiguous): thread -> (row = tid & 63, kq = (tid >> 6) * 4)
  // col-k operands ([K, cols], cols contiguous): thread -> (k = tid >> 4, cq = (tid & 15) * 4)
  const int lr = tid & 63, kq = (tid >> 6) * 4;
  const int lk = tid >> 4, cq = (tid & 15) * 4;

  const float* a_row = nullptr; int64_t a_tok = -1;   // NT / NN: fixed A row for this thread
  if (!tn) {
    int64_t r = m0 + lr;
    if (r < Mg) {
      a_tok = d.perm ? d.perm[tok_off + r] : r;
      int64_t phys = d.a_ids ? (d.a_ids[a_tok] - a_id_off) : a_tok;
      a_row = A + phys * lda;
    }
  }
  const float* b_row = nullptr;                       // NT: fixed B row (= output column n)
  if (b_rowk) { int64_t n = n0 + lr; if (n < Ng) b_row = B + n * ldb; }

  auto load_a = [&](int64_t k0) -> float4 {
    float4 v;
    if (!tn) {
      v = ld4_guard(a_row, k0 + kq, k_end);
      if (d.a_act) v = tanh4(v);
      if (d.drop_on == 1 && a_row) v = drop_apply4(d.drop, (uint64_t)a_tok, (uint32_t)(k0 + kq), v);
    } else {
      int64_t k = k0 + lk;
      v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < k_end) {
        int64_t t = d.perm ? d.perm[tok_off + k] : k;
        v = ld4_guard(A + t * lda, m0 + cq, Mg);
        if (d.a_act) v = tanh4(v);
      }
    }
    return v;
  };
  auto load_b = [&](int64_t k0) -> float4 {
    float4 v;
    if (b_rowk) {
      v = ld4_guard(b_row, k0 + kq, k_end);
    } else {
      int64_t k = k0 + lk;
      v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < k_end) {
        if (tn) {
          int64_t t = d.perm ? d.perm[tok_off + k] : k;
          int64_t phys = d.b_ids ? (d.b_ids[t] - b_id_off) : t;
          v = ld4_guard(B + phys * ldb, n0 + cq, Ng);
          if (d.b_act) v = tanh4(v);
          if (d.drop_on == 2) v = drop_apply4(d.drop, (uint64_t)t, (uint32_t)(n0 + cq), v);
        } else {
          v = ld4_guard(B + k * ldb, n0 + cq, Ng);
        }
      }
    }
    return v;
  };

  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float csum = 0.f;
  const bool do_colsum = tn && d.colsum != nullptr && blockIdx.x == 0 && tid < TM;

  float4 ra = load_a(k_begin), rb = load_b(k_begin);
  for (int64_t k0 = k_begin; k0 < k_end; k0 += TK) {
    if (!tn) { As[kq + 0][lr] = ra.x; As[kq + 1][lr] = ra.y; As[kq + 2][lr] = ra.z; As[kq + 3][lr] = ra.w; }
    else *reinterpret_cast<float4*>(&As[lk][cq]) = ra;
    if (b_rowk) { Bs[kq + 0][lr] = rb.x; Bs[kq + 1][lr] = rb.y; Bs[kq + 2][lr] = rb.z; Bs[kq + 3][lr] = rb.w; }
    else *reinterpret_cast<float4*>(&Bs[lk][cq]) = rb;
    __syncthreads();
    if (k0 + TK < k_end) { ra = load_a(k0 + TK); rb = load_b(k0 + TK); }
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w};
      const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (do_colsum) {
#pragma unroll
      for (int k = 0; k < TK; ++k) csum += As[k][tid];
    }
    __syncthreads();
  }

  // ---- epilogue ---------------------------------------------------------------------------
  const float oscale = d.out_scale == 0.f ? 1.f : d.out_scale;
  if (tn) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int64_t m = m0 + ty * 4 + i;
      if (m >= Mg) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int64_t n = n0 + tx * 4 + j;
        if (n < Ng) atomicAdd(C + m * ldc + n, acc[i][j] * oscale);
      }
    }
    if (do_colsum) { int64_t m = m0 + tid; if (m < Mg && m < d.colsum_n) atomicAdd(d.colsum + m, csum * oscale); }
    return;
  }
  const int64_t nb = n0 + tx * 4;
  if (nb >= Ng) return;
  float bias[4] = {0.f, 0.f, 0.f, 0.f};
  if (d.bias) {
#pragma unroll
    for (int j = 0; j < 4; ++j) if (nb + j < Ng) bias[j] = __ldg(d.bias + nb + j);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t r = m0 + ty * 4 + i;
    if (r >= Mg) continue;
    int64_t t = d.perm ? d.perm[tok_off + r] : r;
    float4 v = make_float4(fmaf(acc[i][0], oscale, bias[0]), fmaf(acc[i][1], oscale, bias[1]),
                           fmaf(acc[i][2], oscale, bias[2]), fmaf(acc[i][3], oscale, bias[3]));
    if (d.epi_act == 2) {
      // gradient through y = dropout(tanh(.)): dy * f * (1 - (y / f)^2), f = keep * scale
      const float4 yv = ld4_guard(d.aux + t * d.ld_aux, nb, Ng);
      float4 f = make_float4(1.f, 1.f, 1.f, 1.f);
      if (d.epi_drop) f = drop_factor4(d.edrop, (uint64_t)t, (uint32_t)nb);
      const float yy[4] = {yv.x, yv.y, yv.z, yv.w}, ff[4] = {f.x, f.y, f.z, f.w};
      float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float h = ff[j] > 0.f ? yy[j] / ff[j] : 0.f;
        vv[j] = vv[j] * ff[j] * (1.f - h * h);
      }
      v = make_float4(vv[0], vv[1], vv[2], vv[3]);
    }
    if (d.addend) {
      float4 ad = ld4_guard(d.addend + t * d.ld_add, nb, Ng);
      v.x += ad.x; v.y += ad.y; v.z += ad.z; v.w += ad.w;
    }
    if (d.epi_act == 1) {
      v = tanh4(v);
      if (d.epi_drop) v = drop_apply4(d.edrop, (uint64_t)t, (uint32_t)nb, v);
    }
    float* out = C + t * ldc + nb;
    if (nb + 3 < Ng && ((reinterpret_cast<uintptr_t>(out) & 15) == 0)) {
      *reinterpret_cast<float4*>(out) = v;
    } else {
      out[0] = v.x;
      if (nb + 1 < Ng) out[1] = v.y;
      if (nb + 2 < Ng) out[2] = v.z;
      if (nb + 3 < Ng) out[3] = v.w;
    }
  }
}
}  // namespace

int launch_gemm_simt(const GemmDesc& d, cudaStream_t stream) {
  const bool tn = d.form == FORM_TN;
  dim3 grid;
  int64_t kc = 0;
  if (d.ngroups > 0) {
    if (tn) {
      grid.x = (unsigned)((d.max_group_dim + TN - 1) / TN);
      grid.y = (unsigned)((d.M + TM - 1) / TM);
      // tokens per CTA: aim at >= 4 CTAs per SM, between 128 and KC_GROUPED tokens each
      const int64_t xy = (int64_t)grid.x * grid.y;
      kc = d.total_rows * xy / (4 * kSMs);
      kc = kc < 128 ? 128 : (kc > KC_GROUPED ? KC_GROUPED : kc / TK * TK);
      grid.z = (unsigned)((d.total_rows + kc - 1) / kc + d.ngroups);
    } else {
      grid.x = (unsigned)((d.N + TN - 1) / TN);
      grid.y = (unsigned)((d.total_rows + TM - 1) / TM + d.ngroups);
      grid.z = 1;
    }
  } else {
    if (d.M <= 0 || d.N <= 0) return MATCHA_OK;
    grid.x = (unsigned)((d.N + TN - 1) / TN);
    grid.y = (unsigned)((d.M + TM - 1) / TM);
    grid.z = 1;
    if (tn) {
      if (d.K <= 0) return MATCHA_OK;
      int64_t tiles = (int64_t)grid.x * grid.y;
      int64_t want = (4 * kSMs + tiles - 1) / tiles;
      int64_t maxsplit = (d.K + 4 * TK - 1) / (4 * TK);
      int64_t split = want < 1 ? 1 : (want > maxsplit ? maxsplit : want);
      kc = (d.K + split - 1) / split;
      kc = (kc + TK - 1) / TK * TK;
      grid.z = (unsigned)((d.K + kc - 1) / kc);
    }
  }
  if (grid.x == 0 || grid.y == 0 || grid.z == 0) return MATCHA_OK;
  if (grid.y > 65535u || grid.z > 65535u) {
    set_error("gemm: problem too large for one launch (grid %u x %u x %u)", grid.x, grid.y, grid.z);
    return MATCHA_ERR_ARG;
  }
  gemm_simt_kernel<<<grid, NTHREADS, 0, stream>>>(d, kc);
  MATCHA_CHECK_LAUNCH("gemm_simt_kernel");
  return MATCHA_OK;
}

}  // namespace matcha
