// General tcgen05 contraction kernel: every shape the specialised kernels of gemm_tc.cu do not take -- gathered / grouped /
// ragged operands, fused activation / dropout on load, the epilogues of gemm_simt.cu -- on the tensor cores.  It exists for
// embed_dim 128 (BASELINE configs[4]), where the fused embed_dim-64 kernels do not apply and every contraction of the step
// used to run through the fp32 SIMT kernel (profiles/r02_bench_cfg5_1gpu.json: eight launches of 2.1-2.8 ms each).
//
// Same problem description (GemmDesc), same operand accessors and the same epilogue arithmetic as gemm_simt.cu, so the two
// kernels are interchangeable per launch and cross-check each other; the inner product is a 128 x BN x 64 tcgen05 step
// (bf16x3 split, fp32 accumulation in TMEM, tc_common.cuh) instead of a 4 x 4 register tile:
//   NT  C[M, N]  = A[M, K] . B[N, K]^T     A, B K-major operand tiles
//   NN  C[M, N]  = A[M, K] . B[K, N]       B MN-major
//   TN  C[M, N] += A[K, M]^T . B[K, N]     both MN-major, split over K (tokens) with atomic accumulation
// 256 threads stage fp32 -> bf16 hi | lo operands through registers (the next K step's loads are in flight while the tensor
// pipe works): B tiles (and the transposed A tile of the TN form) into shared memory, the A rows of the NT / NN forms straight
// into TENSOR MEMORY (thread = row = TMEM lane; A-from-TMEM MMAs: no A tile in shared memory).  One thread issues the MMAs,
// thread (r, h) reads TMEM lane r and half h of the BN accumulator columns.
#include "tc_common.cuh"

namespace matcha {
namespace {

constexpr int kGT = 256;
constexpr int kGBM = 128, kGBK = 64;

__device__ __forceinline__ float4 ldg4_guard(const float* __restrict__ p, int64_t i, int64_t n) {
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
__device__ __forceinline__ float4 tanh4g(float4 v) { return make_float4(tanhf(v.x), tanhf(v.y), tanhf(v.z), tanhf(v.w)); }

template <int FORM, int BN>
__global__ void __launch_bounds__(kGT, 2) gemm_tcg_kernel(const GemmDesc d, const int64_t kc) {
  constexpr bool tn = FORM == FORM_TN;
  constexpr int kAHalf = 16384;                    // A tile: 128 (m) x 64 (k) bf16
  constexpr int kBHalf = BN * 128;                 // B tile: BN (n) x 64 (k) bf16
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sA = smem;                              // hi | lo
  uint8_t* sB = smem + 2 * kAHalf;                 // hi | lo
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_cs[kGBM];
  const int tid = threadIdx.x, warp = tid >> 5;

  // ---- resolve this CTA's problem (as gemm_simt.cu, 128-row tiles) ----------------------
  int64_t Mg = d.M, Ng = d.N, Kg = d.K;
  const float* A = d.A; const float* B = d.B; float* C = d.C;
  int64_t lda = d.lda, ldb = d.ldb, ldc = d.ldc;
  int64_t a_id_off = d.a_id_off, b_id_off = d.b_id_off;
  int64_t tok_off = 0, m0, k_begin = 0, k_end;
  const int64_t n0 = (int64_t)blockIdx.x * BN;
  bool dead = false;
  if (d.ngroups > 0) {
    const int64_t unit = tn ? kc : kGBM;
    int64_t tile = tn ? blockIdx.z : blockIdx.y, before = 0;
    int g = 0, cnt = 0;
    for (; g < d.ngroups; ++g) {
      cnt = d.group_off[g + 1] - d.group_off[g];
      const int64_t nt = (cnt + unit - 1) / unit;
      if (tile < before + nt) break;
      before += nt;
    }
    if (g == d.ngroups) return;
    const GemmGroup gr = d.groups[g];
    tok_off = d.group_off[g];
    if (tn) {
      Kg = cnt; Ng = gr.dim; C = gr.C; ldc = gr.ldc;
      if (gr.B) { B = gr.B; ldb = gr.ldb; b_id_off = gr.a_id_off; }
      m0 = (int64_t)blockIdx.y * kGBM;
      k_begin = (tile - before) * kc;
      k_end = min((int64_t)cnt, k_begin + kc);
      if (n0 >= Ng) dead = true;
    } else {
      Mg = cnt; Kg = gr.dim;
      if (gr.A) { A = gr.A; lda = gr.lda; a_id_off = gr.a_id_off; }
      B = gr.B; ldb = gr.ldb;
      m0 = (tile - before) * kGBM;
      k_end = Kg;
    }
  } else {
    m0 = (int64_t)blockIdx.y * kGBM;
    if (tn) { k_begin = (int64_t)blockIdx.z * kc; k_end = min(Kg, k_begin + kc); if (k_begin >= k_end) dead = true; }
    else k_end = Kg;
  }
  if (dead) return;                                // uniform per CTA, before any barrier / allocation

  constexpr uint32_t kTmemCols = tn ? BN : (BN == 128 ? 256 : 128);     // NT / NN: + 64 columns of A operand (power of two)
  constexpr uint32_t kColA = BN;                                        // A operand in tensor memory: hi 32 columns | lo 32
  if (warp == 0) tmem_alloc(&tmem_base_s, kTmemCols);
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < kGBM) s_cs[tid] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  // ---- operand loaders: registers of one K step (64 wide) ------------------------------------------------------------
  // row-k operand (rows x 64 k, k contiguous in memory): thread -> (row, span of 64 / (256 / rows) k)
  // col-k operand (64 k x cols, cols contiguous in memory): thread -> granules (k, 8 consecutive columns)
  constexpr int kARegs = 8;                        // float4 per thread for the A tile (128 x 64 / 256 / 4)
  constexpr int kBRegs = BN / 16;                  // float4 per thread for the B tile (BN x 64 / 256 / 4)
  float4 ra[kARegs], rb[kBRegs];
  float csum[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) csum[i] = 0.f;
  const bool do_colsum = tn && d.colsum != nullptr && blockIdx.x == 0;

  // NT / NN: this thread's fixed A row
  const int lr = tid & 127, kh = tid >> 7;
  const float* a_row = nullptr; int64_t a_tok = -1;
  if (!tn) {
    const int64_t r = m0 + lr;
    if (r < Mg) {
      a_tok = d.perm ? d.perm[tok_off + r] : r;
      const int64_t phys = d.a_ids ? (d.a_ids[a_tok] - a_id_off) : a_tok;
      a_row = A + phys * lda;
    }
  }
  // NT: this thread's fixed B row (= output column n)
  constexpr int kBThrPerRow = kGT / BN;            // 2 (BN = 128) or 4 (BN = 64)
  constexpr int kBSpan = kGBK / kBThrPerRow;       // 32 or 16 k per thread
  const int br = tid % BN, bq = tid / BN;
  const float* b_row = nullptr;
  if (FORM == FORM_NT) { const int64_t n = n0 + br; if (n < Ng) b_row = B + n * ldb; }

  auto load_tiles = [&](int64_t k0) {
    if (!tn) {
#pragma unroll
      for (int j = 0; j < kARegs; ++j) {
        const int64_t k = k0 + kh * 32 + 4 * j;
        float4 v = ldg4_guard(a_row, k, k_end);
        if (d.a_act) v = tanh4g(v);
        if (d.drop_on == 1 && a_row) v = drop_apply4(d.drop, (uint64_t)a_tok, (uint32_t)k, v);
        ra[j] = v;
      }
    } else {                                       // A given [K tokens, M]: granule (k, 8 consecutive m)
      const int mg = tid & 15, kk = tid >> 4;
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int64_t k = k0 + kk + 16 * it;
        float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
        if (k < k_end) {
          const int64_t t = d.perm ? d.perm[tok_off + k] : k;
          v0 = ldg4_guard(A + t * lda, m0 + mg * 8, Mg);
          v1 = ldg4_guard(A + t * lda, m0 + mg * 8 + 4, Mg);
          if (d.a_act) { v0 = tanh4g(v0); v1 = tanh4g(v1); }
        }
        ra[2 * it] = v0; ra[2 * it + 1] = v1;
        if (do_colsum) {
          csum[0] += v0.x; csum[1] += v0.y; csum[2] += v0.z; csum[3] += v0.w;
          csum[4] += v1.x; csum[5] += v1.y; csum[6] += v1.z; csum[7] += v1.w;
        }
      }
    }
    if (FORM == FORM_NT) {
#pragma unroll
      for (int j = 0; j < kBRegs; ++j) rb[j] = ldg4_guard(b_row, k0 + bq * kBSpan + 4 * j, k_end);
    } else {                                       // B given [K, N]: granule (k, 8 consecutive n)
      constexpr int kNG = BN / 8;                  // column groups
      constexpr int kKPer = kGT / kNG;             // k rows covered per pass (32 or 16)
      const int ng = tid % kNG, kk = tid / kNG;
#pragma unroll
      for (int it = 0; it < kGBK / kKPer; ++it) {
        const int64_t k = k0 + kk + kKPer * it;
        float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
        if (k < k_end) {
          if (tn) {
            const int64_t t = d.perm ? d.perm[tok_off + k] : k;
            const int64_t phys = d.b_ids ? (d.b_ids[t] - b_id_off) : t;
            v0 = ldg4_guard(B + phys * ldb, n0 + ng * 8, Ng);
            v1 = ldg4_guard(B + phys * ldb, n0 + ng * 8 + 4, Ng);
            if (d.b_act) { v0 = tanh4g(v0); v1 = tanh4g(v1); }
            if (d.drop_on == 2) {
              v0 = drop_apply4(d.drop, (uint64_t)t, (uint32_t)(n0 + ng * 8), v0);
              v1 = drop_apply4(d.drop, (uint64_t)t, (uint32_t)(n0 + ng * 8 + 4), v1);
            }
          } else {
            v0 = ldg4_guard(B + k * ldb, n0 + ng * 8, Ng);
            v1 = ldg4_guard(B + k * ldb, n0 + ng * 8 + 4, Ng);
          }
        }
        rb[2 * it] = v0; rb[2 * it + 1] = v1;
      }
    }
  };
  auto store_tiles = [&]() {
    if (!tn) {                                     // A operand straight into tensor memory: thread = row = TMEM lane, this
                                                   // thread's 32 k elements = 16 columns of packed bf16 pairs (hi and lo)
      const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + kColA + kh * 16;
#pragma unroll
      for (int g8 = 0; g8 < 2; ++g8) {
        uint32_t ph[8], pl[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 v = ra[g8 * 4 + j];
          split2(v.x, v.y, ph[2 * j], pl[2 * j]);
          split2(v.z, v.w, ph[2 * j + 1], pl[2 * j + 1]);
        }
        tmem_st8(ta + g8 * 8, ph);
        tmem_st8(ta + 32 + g8 * 8, pl);
      }
      tmem_st_wait();
    } else {                                       // MN-major [k/8][m/8][k%8][8 m]
      const int mg = tid & 15, kk = tid >> 4;
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int kl = kk + 16 * it;
        uint4 hi, lo;
        split8(ra[2 * it], ra[2 * it + 1], hi, lo);
        const int off = (kl >> 3) * 2048 + mg * 128 + (kl & 7) * 16;
        sts16(sA + off, hi);
        sts16(sA + kAHalf + off, lo);
      }
    }
    if (FORM == FORM_NT) {                         // K-major [k/8][n][8]
#pragma unroll
      for (int j = 0; j < kBRegs / 2; ++j) {
        uint4 hi, lo;
        split8(rb[2 * j], rb[2 * j + 1], hi, lo);
        const int plane = bq * (kBSpan / 8) + j;
        sts16(sB + plane * (BN * 16) + br * 16, hi);
        sts16(sB + kBHalf + plane * (BN * 16) + br * 16, lo);
      }
    } else {                                       // MN-major [k/8][n/8][k%8][8 n]
      constexpr int kNG = BN / 8, kKPer = kGT / kNG;
      const int ng = tid % kNG, kk = tid / kNG;
#pragma unroll
      for (int it = 0; it < kGBK / kKPer; ++it) {
        const int kl = kk + kKPer * it;
        uint4 hi, lo;
        split8(rb[2 * it], rb[2 * it + 1], hi, lo);
        const int off = (kl >> 3) * (BN * 16) + ng * 128 + (kl & 7) * 16;
        sts16(sB + off, hi);
        sts16(sB + kBHalf + off, lo);
      }
    }
  };

  constexpr uint32_t idesc = make_idesc(kGBM, BN, tn, FORM != FORM_NT);
  const uint32_t ah = smem_u32(sA), al = ah + kAHalf, bh = smem_u32(sB), bl = bh + kBHalf;
  // descriptor strides: K-major  LBO = next 8 k = rows * 16, SBO = 128;  MN-major  LBO = next 8 k = (mn / 8) * 128, SBO = 128
  constexpr uint32_t a_lbo = 2048, b_lbo = BN * 16;
  uint32_t phase = 0;
  load_tiles(k_begin);
  for (int64_t k0 = k_begin; k0 < k_end; k0 += kGBK) {
    if (k0 > k_begin) {                            // the previous step's MMAs have read the tiles
      mbar_wait(&bar, phase);
      phase ^= 1;
    }
    store_tiles();
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      if (tn) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_x3s(tmem_base, ah + ks * 2 * a_lbo, al + ks * 2 * a_lbo, bh + ks * 2 * b_lbo, bl + ks * 2 * b_lbo, a_lbo, 128, b_lbo, 128,
                   idesc, k0 == k_begin && ks == 0);
      } else {                                     // A from tensor memory (8 columns per K = 16 step), B from shared memory
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t a_hi = tmem_base + kColA + ks * 8, a_lo = a_hi + 32;
          const uint64_t b_hi = make_smem_desc(bh + ks * 2 * b_lbo, b_lbo, 128), b_lo = make_smem_desc(bl + ks * 2 * b_lbo, b_lbo, 128);
          umma_ts_bf16(tmem_base, a_lo, b_hi, idesc, (k0 == k_begin && ks == 0) ? 0u : 1u);
          umma_ts_bf16(tmem_base, a_hi, b_lo, idesc, 1u);
          umma_ts_bf16(tmem_base, a_hi, b_hi, idesc, 1u);
        }
      }
      umma_commit(&bar);
    }
    if (k0 + kGBK < k_end) load_tiles(k0 + kGBK);
  }
  mbar_wait(&bar, phase);
  tc_fence_after();

  // ---- epilogue: thread (r, h) = TMEM lane r, accumulator columns [h * BN / 2, (h + 1) * BN / 2) --------------------
  const float oscale = d.out_scale == 0.f ? 1.f : d.out_scale;
  const int h = tid >> 7;
  const uint32_t tl = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + h * (BN / 2);
  const int64_t mrow = m0 + lr;
#pragma unroll 1
  for (int cc = 0; cc < BN / 64; ++cc) {
    float v[32];
    tmem_ld32(tl + cc * 32, v);
    const int64_t nb0 = n0 + h * (BN / 2) + cc * 32;
    if (mrow >= Mg) continue;
    if (tn) {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (nb0 + i < Ng) atomicAdd(C + mrow * ldc + nb0 + i, v[i] * oscale);
      continue;
    }
    const int64_t t = d.perm ? d.perm[tok_off + mrow] : mrow;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int64_t nb = nb0 + 4 * j;
      if (nb >= Ng) break;
      float bias[4] = {0.f, 0.f, 0.f, 0.f};
      if (d.bias) {
#pragma unroll
        for (int e = 0; e < 4; ++e) if (nb + e < Ng) bias[e] = __ldg(d.bias + nb + e);
      }
      float4 o = make_float4(fmaf(v[4 * j], oscale, bias[0]), fmaf(v[4 * j + 1], oscale, bias[1]), fmaf(v[4 * j + 2], oscale, bias[2]),
                             fmaf(v[4 * j + 3], oscale, bias[3]));
      if (d.epi_act == 2) {      // gradient through y = dropout(tanh(.)): dy * f * (1 - (y / f)^2), f = keep * scale
        const float4 yv = ldg4_guard(d.aux + t * d.ld_aux, nb, Ng);
        float4 f = make_float4(1.f, 1.f, 1.f, 1.f);
        if (d.epi_drop) f = drop_factor4(d.edrop, (uint64_t)t, (uint32_t)nb);
        const float yy[4] = {yv.x, yv.y, yv.z, yv.w}, ff[4] = {f.x, f.y, f.z, f.w};
        float vv[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float hh = ff[e] > 0.f ? yy[e] / ff[e] : 0.f;
          vv[e] = vv[e] * ff[e] * (1.f - hh * hh);
        }
        o = make_float4(vv[0], vv[1], vv[2], vv[3]);
      }
      if (d.addend) {
        const float4 ad = ldg4_guard(d.addend + t * d.ld_add, nb, Ng);
        o.x += ad.x; o.y += ad.y; o.z += ad.z; o.w += ad.w;
      }
      if (d.epi_act == 1) {
        o = tanh4g(o);
        if (d.epi_drop) o = drop_apply4(d.edrop, (uint64_t)t, (uint32_t)nb, o);
      }
      float* out = C + t * ldc + nb;
      if (nb + 3 < Ng && ((reinterpret_cast<uintptr_t>(out) & 15) == 0)) {
        *reinterpret_cast<float4*>(out) = o;
      } else {
        out[0] = o.x;
        if (nb + 1 < Ng) out[1] = o.y;
        if (nb + 2 < Ng) out[2] = o.z;
        if (nb + 3 < Ng) out[3] = o.w;
      }
    }
  }
  if (do_colsum) {     // bias gradient: column sums of A over this CTA's tokens (16 threads share each group of 8 columns)
    const int mg = tid & 15;
#pragma unroll
    for (int e = 0; e < 8; ++e) atomicAdd(&s_cs[mg * 8 + e], csum[e]);
    __syncthreads();
    if (tid < kGBM) {
      const int64_t m = m0 + tid;
      if (m < Mg && m < d.colsum_n) atomicAdd(d.colsum + m, s_cs[tid] * oscale);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

template <int FORM, int BN>
int launch_form(const GemmDesc& d, dim3 grid, int64_t kc, cudaStream_t stream) {
  constexpr int smem = 32768 + 2 * BN * 128;
  static bool once = false;
  if (!once) {
    if (int rc = check_cuda(cudaFuncSetAttribute(gemm_tcg_kernel<FORM, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem),
                            "cudaFuncSetAttribute"))
      return rc;
    once = true;
  }
  gemm_tcg_kernel<FORM, BN><<<grid, kGT, smem, stream>>>(d, kc);
  MATCHA_CHECK_LAUNCH("gemm_tcg_kernel");
  return MATCHA_OK;
}

}  // namespace

// Takes the launch when the problem is large enough for 128-row tensor-core tiles; *handled = false leaves it to the SIMT kernel
int launch_gemm_tcg(const GemmDesc& d, cudaStream_t stream, bool* handled) {
  *handled = false;
  const bool tn = d.form == FORM_TN;
  const int64_t rows = d.ngroups > 0 ? d.total_rows : (tn ? d.K : d.M);      // token dimension
  const int64_t maxN = (tn && d.ngroups > 0) ? d.max_group_dim : d.N;
  if (rows < 1024 || maxN < 64) return MATCHA_OK;
  if ((tn || d.ngroups == 0) && d.M <= 0) return MATCHA_OK;     // grouped NT / NN launches carry their row counts in group_off
  if (!tn && d.ngroups == 0 && d.K < 32) return MATCHA_OK;
  if (tn && d.M < 64) return MATCHA_OK;
  const int BN = maxN >= 128 ? 128 : 64;
  dim3 grid;
  int64_t kc = 0;
  grid.x = (unsigned)((maxN + BN - 1) / BN);
  if (d.ngroups > 0) {
    if (tn) {
      grid.y = (unsigned)((d.M + kGBM - 1) / kGBM);
      const int64_t xy = (int64_t)grid.x * grid.y;
      kc = d.total_rows * xy / (2 * kSMs);                     // aim at two CTAs per SM
      kc = kc < 256 ? 256 : (kc > 2048 ? 2048 : kc / kGBK * kGBK);
      grid.z = (unsigned)((d.total_rows + kc - 1) / kc + d.ngroups);
    } else {
      grid.y = (unsigned)((d.total_rows + kGBM - 1) / kGBM + d.ngroups);
      grid.z = 1;
    }
  } else {
    grid.y = (unsigned)((d.M + kGBM - 1) / kGBM);
    grid.z = 1;
    if (tn) {
      const int64_t tiles = (int64_t)grid.x * grid.y;
      const int64_t want = (2 * kSMs + tiles - 1) / tiles;
      const int64_t maxsplit = (d.K + 255) / 256;
      const int64_t split = want < 1 ? 1 : (want > maxsplit ? maxsplit : want);
      kc = (d.K + split - 1) / split;
      kc = (kc + kGBK - 1) / kGBK * kGBK;
      grid.z = (unsigned)((d.K + kc - 1) / kc);
    }
  }
  if (grid.x == 0 || grid.y == 0 || grid.z == 0) return MATCHA_OK;
  if (grid.y > 65535u || grid.z > 65535u) return MATCHA_OK;     // SIMT kernel reports the size error
  int rc;
  if (d.form == FORM_NT) rc = BN == 128 ? launch_form<FORM_NT, 128>(d, grid, kc, stream) : launch_form<FORM_NT, 64>(d, grid, kc, stream);
  else if (d.form == FORM_NN) rc = BN == 128 ? launch_form<FORM_NN, 128>(d, grid, kc, stream) : launch_form<FORM_NN, 64>(d, grid, kc, stream);
  else rc = BN == 128 ? launch_form<FORM_TN, 128>(d, grid, kc, stream) : launch_form<FORM_TN, 64>(d, grid, kc, stream);
  if (rc) return rc;
  *handled = true;
  return MATCHA_OK;
}

}  // namespace matcha
