// CSR form of the first encoder layer (Modules.py:58-65 sparse=True rows + :104-113 first tied layer):
//   H0[t] = tanh( W0_c . dropout(f_t) )  with f_t the CSR row of node id_t in chromosome c
//         = tanh( sum over nonzeros (col, val) of  drop(val) * W0T_c[col, :] )
// an SpMM whose dense operand is the TRANSPOSED weight W0T_c [n_c, 64] (written by matcha_prepare), so that every nonzero
// reads one contiguous 256-byte weight row: 16 lanes x 128-bit loads.  Tokens arrive bucketed by chromosome (rowwise.cu),
// a CTA works on 128 tokens of one chromosome and sweeps the columns in shared-memory tiles of W0T_c; the CSR arrays are
// the only HBM stream (8 bytes per nonzero), the evidence for this kernel is achieved GB/s.
// Backward: dW0T_c[col, :] += drop(val) * dH0pre[t, :] with 128-bit vector reductions, then a transpose-add into the
// reference layout dW0_c [64, n_c].
#include "rowwise.cuh"

namespace matcha {
namespace {

constexpr int kCsrThreads = 256;

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

// A CTA owns kTokCta tokens of one chromosome (kTph per half-warp, accumulators in registers) and sweeps the columns
// in tiles of kTileRows weight rows staged in shared memory; CSR rows are column-sorted, so every token keeps a running
// pointer into its row.  L2 traffic for the weights drops from 256 B per nonzero to n_c x 256 B per CTA.
constexpr int kTph = 8;
constexpr int kTokCta = (kCsrThreads / 16) * kTph;     // 128
constexpr int kTileRows = 384;                          // 96 KB of weight rows -> two CTAs per SM

__device__ __forceinline__ int half_ballot_count(unsigned hmask, bool pred, int shift) {
  return __popc((__ballot_sync(hmask, pred) >> shift) & 0xffffu);
}

__global__ void __launch_bounds__(kCsrThreads) enc0_csr_fwd_kernel(const int64_t* __restrict__ x, const int32_t* __restrict__ perm,
                                                                   const int32_t* __restrict__ group_off, const CsrMeta m,
                                                                   const float* __restrict__ derived, float* __restrict__ H0,
                                                                   int chunks, const DropCfg drop) {
  extern __shared__ __align__(16) float sW[];
  const int c = blockIdx.x / chunks, chunk = blockIdx.x % chunks;
  const CsrChrom cc = m.c[c];
  const int32_t g0 = group_off[c], g1 = group_off[c + 1];
  const int64_t r0 = g0 + (int64_t)chunk * kTokCta;
  if (r0 >= g1) return;
  const float* WT = derived + cc.w0t_off;
  const int hl = threadIdx.x & 15, hw = threadIdx.x >> 4;
  const int shift = (threadIdx.x & 16);
  const unsigned hmask = shift ? 0xffff0000u : 0x0000ffffu;     // the two half-warps of a warp run different rows
  int64_t tok[kTph], pcur[kTph], pend[kTph];
  float4 acc[kTph];
#pragma unroll
  for (int j = 0; j < kTph; ++j) {
    const int64_t r = r0 + hw * kTph + j;
    tok[j] = -1; pcur[j] = 0; pend[j] = 0;
    acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < g1) {
      tok[j] = perm[r];
      const int64_t row = x[tok[j]] - cc.start;
      pcur[j] = cc.indptr[row]; pend[j] = cc.indptr[row + 1];
    }
  }
  for (int c0 = 0; c0 < cc.n; c0 += kTileRows) {
    const int c1 = (c0 + kTileRows < cc.n) ? c0 + kTileRows : cc.n;
    __syncthreads();
    for (int i = threadIdx.x; i < (c1 - c0) * (kD / 4); i += kCsrThreads)
      reinterpret_cast<float4*>(sW)[i] = __ldg(reinterpret_cast<const float4*>(WT + (int64_t)c0 * kD) + i);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kTph; ++j) {
      bool more = tok[j] >= 0;
      while (more) {
        // the half-warp fetches up to 16 nonzeros (coalesced), keeps those whose column lies in this tile
        const int64_t pp = pcur[j] + hl;
        const bool in = pp < pend[j];
        const int col_l = in ? __ldg(cc.indices + pp) : 0x7fffffff;
        float val_l = in ? __ldg(cc.values + pp) : 0.f;
        const int cnt = half_ballot_count(hmask, in && col_l < c1, shift);
        if (drop.thr != 0u && in) val_l = drop_apply(drop, drop_word(drop, (uint64_t)tok[j], (uint32_t)col_l), (uint32_t)col_l, val_l);
        for (int k = 0; k < cnt; ++k) {
          const int col = __shfl_sync(hmask, col_l, k, 16);
          const float val = __shfl_sync(hmask, val_l, k, 16);
          const float4 w = *reinterpret_cast<const float4*>(sW + (col - c0) * kD + hl * 4);
          acc[j].x = fmaf(val, w.x, acc[j].x); acc[j].y = fmaf(val, w.y, acc[j].y);
          acc[j].z = fmaf(val, w.z, acc[j].z); acc[j].w = fmaf(val, w.w, acc[j].w);
        }
        pcur[j] += cnt;
        more = cnt == 16;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < kTph; ++j)
    if (tok[j] >= 0)
      *reinterpret_cast<float4*>(H0 + tok[j] * kD + hl * 4) =
          make_float4(tanhf(acc[j].x), tanhf(acc[j].y), tanhf(acc[j].z), tanhf(acc[j].w));
}

// dW0T_c[col, :] += drop(val) * dH0pre[t, :]: same sweep, the column tile of the gradient accumulates in shared memory
// (fp32 shared atomics) and is flushed with 128-bit vector reductions
__global__ void __launch_bounds__(kCsrThreads) enc0_csr_wgrad_kernel(const int64_t* __restrict__ x, const int32_t* __restrict__ perm,
                                                                     const int32_t* __restrict__ group_off, const CsrMeta m,
                                                                     const float* __restrict__ dH0pre, float* __restrict__ dgrad,
                                                                     int chunks, const DropCfg drop) {
  extern __shared__ __align__(16) float sW[];
  const int c = blockIdx.x / chunks, chunk = blockIdx.x % chunks;
  const CsrChrom cc = m.c[c];
  const int32_t g0 = group_off[c], g1 = group_off[c + 1];
  const int64_t r0 = g0 + (int64_t)chunk * kTokCta;
  if (r0 >= g1) return;
  float* dWT = dgrad + cc.w0t_off;
  const int hl = threadIdx.x & 15, hw = threadIdx.x >> 4;
  const int shift = (threadIdx.x & 16);
  const unsigned hmask = shift ? 0xffff0000u : 0x0000ffffu;
  int64_t tok[kTph], pcur[kTph], pend[kTph];
  float4 g[kTph];
#pragma unroll
  for (int j = 0; j < kTph; ++j) {
    const int64_t r = r0 + hw * kTph + j;
    tok[j] = -1; pcur[j] = 0; pend[j] = 0;
    g[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < g1) {
      tok[j] = perm[r];
      const int64_t row = x[tok[j]] - cc.start;
      pcur[j] = cc.indptr[row]; pend[j] = cc.indptr[row + 1];
      g[j] = __ldg(reinterpret_cast<const float4*>(dH0pre + tok[j] * kD + hl * 4));
    }
  }
  for (int c0 = 0; c0 < cc.n; c0 += kTileRows) {
    const int c1 = (c0 + kTileRows < cc.n) ? c0 + kTileRows : cc.n;
    __syncthreads();
    for (int i = threadIdx.x; i < (c1 - c0) * (kD / 4); i += kCsrThreads) reinterpret_cast<float4*>(sW)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    bool any = false;
#pragma unroll
    for (int j = 0; j < kTph; ++j) {
      bool more = tok[j] >= 0;
      while (more) {
        const int64_t pp = pcur[j] + hl;
        const bool in = pp < pend[j];
        const int col_l = in ? __ldg(cc.indices + pp) : 0x7fffffff;
        float val_l = in ? __ldg(cc.values + pp) : 0.f;
        const int cnt = half_ballot_count(hmask, in && col_l < c1, shift);
        if (drop.thr != 0u && in) val_l = drop_apply(drop, drop_word(drop, (uint64_t)tok[j], (uint32_t)col_l), (uint32_t)col_l, val_l);
        for (int k = 0; k < cnt; ++k) {
          const int col = __shfl_sync(hmask, col_l, k, 16);
          const float val = __shfl_sync(hmask, val_l, k, 16);
          if (val != 0.f) {
            float* dst = sW + (col - c0) * kD + hl * 4;
            atomicAdd(dst, val * g[j].x); atomicAdd(dst + 1, val * g[j].y); atomicAdd(dst + 2, val * g[j].z); atomicAdd(dst + 3, val * g[j].w);
            any = true;
          }
        }
        pcur[j] += cnt;
        more = cnt == 16;
      }
    }
    if (__syncthreads_or(any)) {
      for (int i = threadIdx.x; i < (c1 - c0) * (kD / 4); i += kCsrThreads) {
        const float4 v = reinterpret_cast<const float4*>(sW)[i];
        if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) {
          float* dst = dWT + (int64_t)c0 * kD + (int64_t)i * 4;
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
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
// upper bound of token chunks any chromosome can need; CTAs beyond a chromosome's token count exit at once
int chunks_for(int64_t T) {
  const int64_t k = (T + kTokCta - 1) / kTokCta;
  return (int)(k < 1 ? 1 : k);
}
constexpr int kCsrSmem = kTileRows * kD * 4;

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

int launch_enc0_csr_fwd(const matcha_model_desc* m, int64_t w0t_base, const int64_t* x, int64_t T, const int32_t* perm,
                        const int32_t* group_off, float* H0, DropCfg drop, cudaStream_t s) {
  const CsrMeta cm = make_meta(m, w0t_base);
  const int chunks = chunks_for(T);
  static bool once = false;
  if (!once) {
    if (int rc = check_cuda(cudaFuncSetAttribute(enc0_csr_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCsrSmem), "cudaFuncSetAttribute"))
      return rc;
    once = true;
  }
  enc0_csr_fwd_kernel<<<m->n_chrom * chunks, kCsrThreads, kCsrSmem, s>>>(x, perm, group_off, cm, m->derived, H0, chunks, drop);
  MATCHA_CHECK_LAUNCH("enc0_csr_fwd");
  return MATCHA_OK;
}

int launch_enc0_csr_wgrad(const matcha_model_desc* m, int64_t w0t_base, const int64_t* x, int64_t T, const int32_t* perm,
                          const int32_t* group_off, const float* dH0pre, DropCfg drop, cudaStream_t s) {
  const CsrMeta cm = make_meta(m, w0t_base);
  const int chunks = chunks_for(T);
  static bool once = false;
  if (!once) {
    if (int rc = check_cuda(cudaFuncSetAttribute(enc0_csr_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCsrSmem), "cudaFuncSetAttribute"))
      return rc;
    once = true;
  }
  enc0_csr_wgrad_kernel<<<m->n_chrom * chunks, kCsrThreads, kCsrSmem, s>>>(x, perm, group_off, cm, dH0pre, m->derived_grad, chunks, drop);
  MATCHA_CHECK_LAUNCH("enc0_csr_wgrad");
  dim3 grid(8, (unsigned)m->n_chrom);
  transpose_add_kernel<<<grid, 256, 0, s>>>(m->derived_grad, cm, m->grads);
  MATCHA_CHECK_LAUNCH("transpose_add");
  return MATCHA_OK;
}

}  // namespace matcha
