// CSR form of the first encoder layer (Modules.py:58-65 sparse=True rows + :104-113 first tied layer):
//   H0[t] = tanh( W0_c . dropout(f_t) )  with f_t the CSR row of node id_t in chromosome c
//         = tanh( sum over nonzeros (col, val) of  drop(val) * W0T_c[col, :] )
// an SpMM whose dense operand is the TRANSPOSED weight W0T_c [n_c, 64] (written by matcha_prepare), so that every nonzero
// reads one contiguous 256-byte weight row: 16 lanes x 128-bit loads.  Tokens arrive bucketed by chromosome (rowwise.cu),
// a CTA works on one chromosome and stages W0T_c in shared memory when it fits (n_c <= kStageRows); the CSR arrays are
// the only HBM stream (8 bytes per nonzero), the evidence for this kernel is achieved GB/s.
// Backward: dW0T_c[col, :] += drop(val) * dH0pre[t, :] with 128-bit vector reductions, then a transpose-add into the
// reference layout dW0_c [64, n_c].
#include "rowwise.cuh"

namespace matcha {
namespace {

constexpr int kCsrThreads = 256;
constexpr int kStageRows = 768;                 // W0T rows staged in shared memory (768 x 256 B = 192 KB)

struct CsrChrom {
  const int64_t* indptr; const int32_t* indices; const float* values;
  int64_t start;        // first node id of the chromosome
  int32_t n;            // bins
  int64_t w0t_off;      // float offset of W0T_c inside the derived buffer
  int64_t off_w0;       // float offset of W0_c [64, n] inside params / grads
};
struct CsrMeta {
  int32_t n_chrom;
  CsrChrom c[MATCHA_MAX_CHROM];
};

// W0_c [64, n_c] -> W0T_c [n_c, 64]
__global__ void transpose_w0_kernel(const float* __restrict__ params, const CsrMeta m, float* __restrict__ derived) {
  const int c = blockIdx.y;
  const int n = m.c[c].n;
  const float* W = params + m.c[c].off_w0;
  float* WT = derived + m.c[c].w0t_off;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (int64_t)n * kD; i += (int64_t)gridDim.x * blockDim.x) {
    const int col = (int)(i / kD), o = (int)(i % kD);
    WT[i] = W[(int64_t)o * n + col];
  }
}

template <bool STAGED>
__global__ void __launch_bounds__(kCsrThreads) enc0_csr_fwd_kernel(const int64_t* __restrict__ x, const int32_t* __restrict__ perm,
                                                                   const int32_t* __restrict__ group_off, const CsrMeta m,
                                                                   const float* __restrict__ derived, float* __restrict__ H0,
                                                                   int chunks, const DropCfg drop) {
  extern __shared__ __align__(16) float sW[];
  const int c = blockIdx.x / chunks, chunk = blockIdx.x % chunks;
  const CsrChrom cc = m.c[c];
  const int32_t g0 = group_off[c], g1 = group_off[c + 1];
  if (g0 == g1) return;
  const float* WT = derived + cc.w0t_off;
  const bool staged = STAGED && cc.n <= kStageRows;
  if (staged) {
    for (int i = threadIdx.x; i < cc.n * (kD / 4); i += kCsrThreads)
      reinterpret_cast<float4*>(sW)[i] = __ldg(reinterpret_cast<const float4*>(WT) + i);
    __syncthreads();
  }
  const float* Wsrc = staged ? sW : WT;
  const int hl = threadIdx.x & 15, hw = threadIdx.x >> 4;
  const unsigned hmask = (threadIdx.x & 16) ? 0xffff0000u : 0x0000ffffu;     // the two half-warps of a warp run different rows
  const int64_t per = (g1 - g0 + chunks - 1) / chunks;
  const int64_t r0 = g0 + chunk * per, r1 = (r0 + per < g1) ? r0 + per : g1;
  for (int64_t r = r0 + hw; r < r1; r += kCsrThreads / 16) {
    const int64_t t = perm[r];
    const int64_t row = x[t] - cc.start;
    const int64_t p0 = cc.indptr[row], p1 = cc.indptr[row + 1];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t p = p0; p < p1; p += 16) {
      // the half-warp fetches 16 nonzeros at once (coalesced 64-byte index and value reads), then walks them
      const int64_t pp = p + hl;
      const int col_l = pp < p1 ? __ldg(cc.indices + pp) : 0;
      float val_l = pp < p1 ? __ldg(cc.values + pp) : 0.f;
      if (drop.thr != 0u && pp < p1) val_l = drop_apply(drop, drop_word(drop, (uint64_t)t, (uint32_t)col_l), (uint32_t)col_l, val_l);
      const int cnt = (p1 - p < 16) ? (int)(p1 - p) : 16;
      for (int k = 0; k < cnt; ++k) {
        const int col = __shfl_sync(hmask, col_l, k, 16);
        const float val = __shfl_sync(hmask, val_l, k, 16);
        const float4 w = staged ? *reinterpret_cast<const float4*>(Wsrc + (int64_t)col * kD + hl * 4)
                                : __ldg(reinterpret_cast<const float4*>(Wsrc + (int64_t)col * kD + hl * 4));
        acc.x = fmaf(val, w.x, acc.x); acc.y = fmaf(val, w.y, acc.y); acc.z = fmaf(val, w.z, acc.z); acc.w = fmaf(val, w.w, acc.w);
      }
    }
    *reinterpret_cast<float4*>(H0 + t * kD + hl * 4) = make_float4(tanhf(acc.x), tanhf(acc.y), tanhf(acc.z), tanhf(acc.w));
  }
}

// dW0T_c[col, :] += drop(val) * dH0pre[t, :]
__global__ void __launch_bounds__(kCsrThreads) enc0_csr_wgrad_kernel(const int64_t* __restrict__ x, const int32_t* __restrict__ perm,
                                                                     const int32_t* __restrict__ group_off, const CsrMeta m,
                                                                     const float* __restrict__ dH0pre, float* __restrict__ dgrad,
                                                                     int chunks, const DropCfg drop) {
  const int c = blockIdx.x / chunks, chunk = blockIdx.x % chunks;
  const CsrChrom cc = m.c[c];
  const int32_t g0 = group_off[c], g1 = group_off[c + 1];
  if (g0 == g1) return;
  float* dWT = dgrad + cc.w0t_off;
  const int hl = threadIdx.x & 15, hw = threadIdx.x >> 4;
  const unsigned hmask = (threadIdx.x & 16) ? 0xffff0000u : 0x0000ffffu;     // the two half-warps of a warp run different rows
  const int64_t per = (g1 - g0 + chunks - 1) / chunks;
  const int64_t r0 = g0 + chunk * per, r1 = (r0 + per < g1) ? r0 + per : g1;
  for (int64_t r = r0 + hw; r < r1; r += kCsrThreads / 16) {
    const int64_t t = perm[r];
    const int64_t row = x[t] - cc.start;
    const int64_t p0 = cc.indptr[row], p1 = cc.indptr[row + 1];
    const float4 g = __ldg(reinterpret_cast<const float4*>(dH0pre + t * kD + hl * 4));
    for (int64_t p = p0; p < p1; p += 16) {
      const int64_t pp = p + hl;
      const int col_l = pp < p1 ? __ldg(cc.indices + pp) : 0;
      float val_l = pp < p1 ? __ldg(cc.values + pp) : 0.f;
      if (drop.thr != 0u && pp < p1) val_l = drop_apply(drop, drop_word(drop, (uint64_t)t, (uint32_t)col_l), (uint32_t)col_l, val_l);
      const int cnt = (p1 - p < 16) ? (int)(p1 - p) : 16;
      for (int k = 0; k < cnt; ++k) {
        const int col = __shfl_sync(hmask, col_l, k, 16);
        const float val = __shfl_sync(hmask, val_l, k, 16);
        if (val != 0.f) {
          float* dst = dWT + (int64_t)col * kD + hl * 4;
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(val * g.x), "f"(val * g.y), "f"(val * g.z),
                       "f"(val * g.w)
                       : "memory");
        }
      }
    }
  }
}

// dW0_c [64, n_c] += dW0T_c^T
__global__ void transpose_add_kernel(const float* __restrict__ dgrad, const CsrMeta m, float* __restrict__ grads) {
  const int c = blockIdx.y;
  const int n = m.c[c].n;
  const float* dWT = dgrad + m.c[c].w0t_off;
  float* dW = grads + m.c[c].off_w0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (int64_t)n * kD; i += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(i / n), col = (int)(i % n);
    dW[i] += dWT[(int64_t)col * kD + o];
  }
}

CsrMeta make_meta(const matcha_model_desc* m, int64_t w0t_base) {
  CsrMeta cm;
  cm.n_chrom = m->n_chrom;
  int64_t off = w0t_base;
  for (int c = 0; c < m->n_chrom; ++c) {
    cm.c[c].indptr = m->feat_indptr[c]; cm.c[c].indices = m->feat_indices[c]; cm.c[c].values = m->feat_values[c];
    cm.c[c].start = m->chrom_start[c];
    cm.c[c].n = (int32_t)(m->chrom_end[c] - m->chrom_start[c]);
    cm.c[c].w0t_off = off;
    cm.c[c].off_w0 = m->off_w0[c];
    off += (int64_t)cm.c[c].n * kD;
  }
  return cm;
}
int chunks_for(int n_chrom) {
  int k = (2 * kSMs + n_chrom - 1) / n_chrom;
  return k < 1 ? 1 : k;
}

}  // namespace

bool model_uses_csr(const matcha_model_desc* m) {
  for (int c = 0; c < m->n_chrom; ++c)
    if (!m->feat[c] && m->feat_indptr[c] && m->feat_indices[c] && m->feat_values[c]) return true;
  return false;
}

int launch_csr_prepare(const matcha_model_desc* m, int64_t w0t_base, cudaStream_t s) {
  const CsrMeta cm = make_meta(m, w0t_base);
  dim3 grid(8, (unsigned)m->n_chrom);
  transpose_w0_kernel<<<grid, 256, 0, s>>>(m->params, cm, m->derived);
  MATCHA_CHECK_LAUNCH("transpose_w0");
  return MATCHA_OK;
}

int launch_enc0_csr_fwd(const matcha_model_desc* m, int64_t w0t_base, const int64_t* x, const int32_t* perm,
                        const int32_t* group_off, float* H0, DropCfg drop, cudaStream_t s) {
  const CsrMeta cm = make_meta(m, w0t_base);
  int max_n = 0;
  for (int c = 0; c < m->n_chrom; ++c) max_n = max_n > cm.c[c].n ? max_n : cm.c[c].n;
  const int chunks = chunks_for(m->n_chrom);
  const int stage_rows = max_n < kStageRows ? max_n : kStageRows;
  const int smem = stage_rows * kD * 4;
  static int set_for = 0;
  if (set_for < smem) {
    if (int rc = check_cuda(cudaFuncSetAttribute(enc0_csr_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "cudaFuncSetAttribute"))
      return rc;
    set_for = smem;
  }
  enc0_csr_fwd_kernel<true><<<m->n_chrom * chunks, kCsrThreads, smem, s>>>(x, perm, group_off, cm, m->derived, H0, chunks, drop);
  MATCHA_CHECK_LAUNCH("enc0_csr_fwd");
  return MATCHA_OK;
}

int launch_enc0_csr_wgrad(const matcha_model_desc* m, int64_t w0t_base, const int64_t* x, const int32_t* perm,
                          const int32_t* group_off, const float* dH0pre, DropCfg drop, cudaStream_t s) {
  const CsrMeta cm = make_meta(m, w0t_base);
  const int chunks = chunks_for(m->n_chrom);
  enc0_csr_wgrad_kernel<<<m->n_chrom * chunks, kCsrThreads, 0, s>>>(x, perm, group_off, cm, dH0pre, m->derived_grad, chunks, drop);
  MATCHA_CHECK_LAUNCH("enc0_csr_wgrad");
  dim3 grid(8, (unsigned)m->n_chrom);
  transpose_add_kernel<<<grid, 256, 0, s>>>(m->derived_grad, cm, m->grads);
  MATCHA_CHECK_LAUNCH("transpose_add");
  return MATCHA_OK;
}

}  // namespace matcha
