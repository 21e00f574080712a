// Denoise post-processing on the device (reference: denoise_contact.py:31-61 proba2matrix and
// the per-chromosome tail :160-192).  Input: the packed upper-triangular pair scores the all-pairs scorer leaves in HBM
// (generate_pair_wise order, denoise_contact.py:67-74) and the observed intra-chromosomal contact block.  Output: the
// reference's `my` matrix [n, n].  The reference does this on the host with a 3e8-iteration Python loop (:160), np.add.at
// into dense matrices and six full-matrix numpy passes; here it is four streaming kernels over n^2 fp32 (HBM-bound:
// ~40 bytes per matrix element in total) plus the element-wise quantile map.
//   fill      M = U + U^T for the score and the observed matrix (U = strict band-limited upper fill; the diagonal doubles,
//             as m + m.T does at :46), 32 x 32 tiles mirrored through shared memory so both triangles are written coalesced
//   rowstat   row means in fp64 -> sqrt-coverage vectors (:163-166, :171-174), gap flags (:169-170)
//   combine   my0 = max(p o, p) on the normalised matrices (:177) + its row means (:178-179)
//   final     my = my0 / c / c, gap rows / columns zeroed (:180-185)
// Row and column means coincide for these symmetric matrices, so one vector serves both divisions (numpy's two reductions
// differ only in fp32 summation order).  Division order and the `+ 1e-15` in fp32 follow the reference literally.
#include "common.cuh"

namespace matcha {
namespace {

__device__ __forceinline__ int64_t pair_index(int64_t i, int64_t j, int64_t full, int md) {
  return i * full - i * (i - 1) / 2 + (j - i - md);
}

__global__ void __launch_bounds__(256) denoise_fill_kernel(const float* __restrict__ proba, const float* __restrict__ origin, int64_t ld,
                                                           int n, int md, float* __restrict__ MP, float* __restrict__ MW) {
  __shared__ float tp[32][33], tw[32][33];
  const int bx = blockIdx.x, by = blockIdx.y;
  if (bx < by) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t full = n - md;
  for (int r = ty; r < 32; r += 8) {
    const int i = by * 32 + r, j = bx * 32 + tx;
    float p = 0.f, w = 0.f;
    if (i < n && j < n && j >= i && j - i >= md) {
      p = __ldg(proba + pair_index(i, j, full, md));
      w = __ldg(origin + (int64_t)i * ld + j);
    }
    tp[r][tx] = p;
    tw[r][tx] = w;
  }
  __syncthreads();
  const bool diag = bx == by;
  for (int r = ty; r < 32; r += 8) {
    const int i = by * 32 + r, j = bx * 32 + tx;
    if (i < n && j < n) {
      MP[(int64_t)i * n + j] = tp[r][tx] + (diag ? tp[tx][r] : 0.f);
      MW[(int64_t)i * n + j] = tw[r][tx] + (diag ? tw[tx][r] : 0.f);
    }
    if (!diag) {
      const int i2 = bx * 32 + r, j2 = by * 32 + tx;
      if (i2 < n && j2 < n) {
        MP[(int64_t)i2 * n + j2] = tp[tx][r];
        MW[(int64_t)i2 * n + j2] = tw[tx][r];
      }
    }
  }
}

__device__ __forceinline__ double block_sum(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < 32) {
    t = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  __syncthreads();
  return t;      // valid in thread 0
}

// one block per row: coverage of the score matrix and of the observed matrix, gap flag
__global__ void __launch_bounds__(256) denoise_rowstat_kernel(const float* __restrict__ MP, const float* __restrict__ MW, int n,
                                                              float* __restrict__ cP, float* __restrict__ cW, uint8_t* __restrict__ gap) {
  __shared__ double sh[8];
  const int i = blockIdx.x;
  double sp = 0.0, sw = 0.0;
  for (int j = threadIdx.x; j < n; j += 256) { sp += (double)MP[(int64_t)i * n + j]; sw += (double)MW[(int64_t)i * n + j]; }
  sp = block_sum(sp, sh);
  sw = block_sum(sw, sh);
  if (threadIdx.x == 0) {
    cP[i] = sqrtf((float)(sp / (double)n)) + 1e-15f;
    cW[i] = sqrtf((float)(sw / (double)n)) + 1e-15f;
    gap[i] = sw == 0.0 ? 1 : 0;
  }
}
__global__ void __launch_bounds__(256) denoise_combine_kernel(const float* __restrict__ MP, const float* __restrict__ MW, int n,
                                                              const float* __restrict__ cP, const float* __restrict__ cW,
                                                              float* __restrict__ MY, float* __restrict__ cY) {
  __shared__ double sh[8];
  const int i = blockIdx.x;
  const float ai = cP[i], bi = cW[i];
  double s = 0.0;
  for (int j = threadIdx.x; j < n; j += 256) {
    const float p = __fdiv_rn(__fdiv_rn(MP[(int64_t)i * n + j], ai), cP[j]);
    const float o = __fdiv_rn(__fdiv_rn(MW[(int64_t)i * n + j], bi), cW[j]);
    const float v = fmaxf(__fmul_rn(p, o), p);
    MY[(int64_t)i * n + j] = v;
    s += (double)v;
  }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) cY[i] = sqrtf((float)(s / (double)n)) + 1e-15f;
}
__global__ void __launch_bounds__(256) denoise_final_kernel(float* __restrict__ MY, int n, const float* __restrict__ cY,
                                                            const uint8_t* __restrict__ gap) {
  const int i = blockIdx.x;
  const float ci = cY[i];
  const bool gi = gap[i] != 0;
  for (int j = threadIdx.x; j < n; j += 256) {
    const float v = __fdiv_rn(__fdiv_rn(MY[(int64_t)i * n + j], ci), cY[j]);
    MY[(int64_t)i * n + j] = (gi || gap[j]) ? 0.f : v;
  }
}

// np.interp (numpy/core/src/multiarray/compiled_base.c arr_interp) on a monotone table reached through an index map, so the
// forward table (q, r) and sklearn's mirrored one (-q[::-1], -r[::-1]) share one code path and one arithmetic
template <bool MIRROR>
__device__ __forceinline__ double interp_tab(double x, const double* __restrict__ q, const double* __restrict__ r, int nq) {
  auto X = [&](int k) { return MIRROR ? -q[nq - 1 - k] : q[k]; };
  auto Y = [&](int k) { return MIRROR ? -r[nq - 1 - k] : r[k]; };
  if (x > X(nq - 1)) return Y(nq - 1);
  if (x < X(0)) return Y(0);
  int lo = 0, hi = nq;                      // largest j with X(j) <= x
  while (hi - lo > 1) { const int m = (lo + hi) >> 1; if (X(m) <= x) lo = m; else hi = m; }
  const int j = lo;
  if (j == nq - 1) return Y(j);
  if (X(j) == x) return Y(j);
  const double slope = (Y(j + 1) - Y(j)) / (X(j + 1) - X(j));
  double res = slope * (x - X(j)) + Y(j);
  if (res != res) {                         // numpy's non-finite rescue path
    res = slope * (x - X(j + 1)) + Y(j + 1);
    if (res != res && Y(j) == Y(j + 1)) res = Y(j);
  }
  return res;
}
// sklearn.preprocessing.QuantileTransformer(output_distribution="uniform")._transform_col, forward direction
__global__ void __launch_bounds__(256) quantile_uniform_kernel(float* __restrict__ x, int64_t n, const double* __restrict__ q,
                                                               const double* __restrict__ r, int nq) {
  extern __shared__ double tab[];
  double* sq = tab;
  double* sr = tab + nq;
  for (int k = threadIdx.x; k < nq; k += blockDim.x) { sq[k] = q[k]; sr[k] = r[k]; }
  __syncthreads();
  const double lo_x = sq[0], hi_x = sq[nq - 1];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float xf = x[i];
    if (xf != xf) continue;
    const double xv = (double)xf;
    double v = 0.5 * (interp_tab<false>(xv, sq, sr, nq) - interp_tab<true>(-xv, sq, sr, nq));
    if (xv == hi_x) v = 1.0;
    if (xv == lo_x) v = 0.0;
    x[i] = (float)v;
  }
}

__global__ void __launch_bounds__(256) gather_f32_kernel(const float* __restrict__ src, const int64_t* __restrict__ idx, int64_t n,
                                                         float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = src[idx[i]];
}
// balanced[p] = my[i(p), j(p)] in generate_pair_wise order (denoise_contact.py:205): one block per row i, coalesced both ways
__global__ void __launch_bounds__(256) pair_gather_kernel(const float* __restrict__ my, int n, int md, float* __restrict__ out) {
  const int i = blockIdx.x;
  const int64_t full = n - md;
  if (i >= full) return;
  const int64_t base = pair_index(i, i + md, full, md);
  for (int j = i + md + threadIdx.x; j < n; j += 256) out[base + (j - i - md)] = my[(int64_t)i * n + j];
}

}  // namespace
}  // namespace matcha

using namespace matcha;

extern "C" {

int64_t matcha_denoise_workspace_bytes(int64_t n) {
  if (n < 1) return -1;
  const int64_t mat = (n * n * 4 + 255) / 256 * 256, vec = (n * 4 + 255) / 256 * 256;
  return 2 * mat + 3 * vec + (n + 255) / 256 * 256 + 256;
}

int matcha_denoise_matrix(const float* proba, const float* origin, int64_t origin_ld, int64_t n, int32_t min_dis, float* my,
                          void* workspace, int64_t workspace_bytes, void* stream) {
  MATCHA_REQUIRE(proba && origin && my && workspace, "matcha_denoise_matrix: NULL argument");
  MATCHA_REQUIRE(n >= 1 && n < 65536 * 32 && min_dis >= 0 && min_dis < n && origin_ld >= n, "matcha_denoise_matrix: n=%lld min_dis=%d ld=%lld out of range",
                 (long long)n, (int)min_dis, (long long)origin_ld);
  MATCHA_REQUIRE(workspace_bytes >= matcha_denoise_workspace_bytes(n), "matcha_denoise_matrix: workspace too small: need %lld bytes",
                 (long long)matcha_denoise_workspace_bytes(n));
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t mat = (n * n * 4 + 255) / 256 * 256, vec = (n * 4 + 255) / 256 * 256;
  char* p = reinterpret_cast<char*>(((uintptr_t)workspace + 255) / 256 * 256);
  float* MP = (float*)p; p += mat;
  float* MW = (float*)p; p += mat;
  float* cP = (float*)p; p += vec;
  float* cW = (float*)p; p += vec;
  float* cY = (float*)p; p += vec;
  uint8_t* gap = (uint8_t*)p;
  const int nt = (int)((n + 31) / 32);
  denoise_fill_kernel<<<dim3(nt, nt), 256, 0, s>>>(proba, origin, origin_ld, (int)n, min_dis, MP, MW);
  MATCHA_CHECK_LAUNCH("denoise_fill");
  denoise_rowstat_kernel<<<(unsigned)n, 256, 0, s>>>(MP, MW, (int)n, cP, cW, gap);
  MATCHA_CHECK_LAUNCH("denoise_rowstat");
  denoise_combine_kernel<<<(unsigned)n, 256, 0, s>>>(MP, MW, (int)n, cP, cW, my, cY);
  MATCHA_CHECK_LAUNCH("denoise_combine");
  denoise_final_kernel<<<(unsigned)n, 256, 0, s>>>(my, (int)n, cY, gap);
  MATCHA_CHECK_LAUNCH("denoise_final");
  return MATCHA_OK;
}

int matcha_quantile_uniform(float* x, int64_t n, const double* quantiles, const double* references, int32_t nq, void* stream) {
  MATCHA_REQUIRE(x && quantiles && references && n >= 0 && nq >= 1 && nq <= 2048, "matcha_quantile_uniform: bad arguments (nq <= 2048)");
  if (n == 0) return MATCHA_OK;
  int64_t blocks = (n + 255) / 256;
  if (blocks > kSMs * 8) blocks = kSMs * 8;
  quantile_uniform_kernel<<<(unsigned)blocks, 256, 2 * nq * sizeof(double), (cudaStream_t)stream>>>(x, n, quantiles, references, nq);
  MATCHA_CHECK_LAUNCH("quantile_uniform");
  return MATCHA_OK;
}

int matcha_gather_f32(const float* src, const int64_t* idx, int64_t n, float* out, void* stream) {
  MATCHA_REQUIRE(src && idx && out && n >= 0, "matcha_gather_f32: bad arguments");
  if (n == 0) return MATCHA_OK;
  int64_t blocks = (n + 255) / 256;
  if (blocks > kSMs * 8) blocks = kSMs * 8;
  gather_f32_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, idx, n, out);
  MATCHA_CHECK_LAUNCH("gather_f32");
  return MATCHA_OK;
}

int matcha_pair_gather(const float* my, int64_t n, int32_t min_dis, float* out, void* stream) {
  MATCHA_REQUIRE(my && out && n >= 1 && min_dis >= 0 && min_dis < n, "matcha_pair_gather: bad arguments");
  pair_gather_kernel<<<(unsigned)n, 256, 0, (cudaStream_t)stream>>>(my, (int)n, min_dis, out);
  MATCHA_CHECK_LAUNCH("pair_gather");
  return MATCHA_OK;
}

}  // extern "C"
