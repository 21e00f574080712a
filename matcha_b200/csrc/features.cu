// Feature construction on the device: what feeds the node encoder and the reconstruction head.
//   corrcoef       np.corrcoef of every chromosome's intra-contact block, NaN -> 0 (main.py:572-577): rows are variables;
//                  numpy centres and contracts in float64, so does this kernel (fp64 FMA: 64 x 64 output tiles, 4 x 4 per
//                  thread, only tiles on or above the diagonal are computed and mirrored)
//   zscore rows    row-wise z-score (ddof 0) of the POSITIVE entries of the inter-contact matrix, NaN -> 0, in place
//                  (Modules.py:147-152: a Python loop over N rows around scipy.stats.mstats.zscore)
//   pixels -> adj  process.py:148-170 (parse_cool_contact): every cooler pixel (bin1, bin2, count) adds count to
//                  [node1, node2] and [node2, node1] of intra_adj or inter_adj (float64, like the reference's arrays)
//   clusters -> adj process.py:90-105 (edgelist2adj): every ordered pair i != j of a cluster adds 1
#include "common.cuh"

namespace matcha {
namespace {

// ---------------- corrcoef ----------------
__global__ void __launch_bounds__(256) row_mean_kernel(const float* __restrict__ A, int64_t ld, int n, double* __restrict__ mean) {
  __shared__ double sh[8];
  const int i = blockIdx.x;
  double s = 0.0;
  for (int j = threadIdx.x; j < n; j += 256) s += (double)A[(int64_t)i * ld + j];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) { double t = 0.0; for (int w = 0; w < 8; ++w) t += sh[w]; mean[i] = t / (double)n; }
}
// C[i][j] = sum_k (A[i][k] - mean_i) (A[j][k] - mean_j) for the 64 x 64 tile (bi, bj), bj >= bi; mirrored into (bj, bi)
constexpr int kCT = 64, kCK = 16;
__global__ void __launch_bounds__(256) cov_tile_kernel(const float* __restrict__ A, int64_t ld, int n, const double* __restrict__ mean,
                                                       double* __restrict__ C) {
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bj < bi) return;
  __shared__ double sa[kCK][kCT + 1], sb[kCK][kCT + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;       // 16 x 16 threads, 4 x 4 outputs each
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
  const int lr = threadIdx.x >> 2, lk = (threadIdx.x & 3) * 4;   // loader: row lr (0..63), 4 consecutive k
  const int ri = bi * kCT + lr, rj = bj * kCT + lr;
  const double mi = ri < n ? mean[ri] : 0.0, mj = rj < n ? mean[rj] : 0.0;
  for (int k0 = 0; k0 < n; k0 += kCK) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k = k0 + lk + u;
      sa[lk + u][lr] = (ri < n && k < n) ? (double)A[(int64_t)ri * ld + k] - mi : 0.0;
      sb[lk + u][lr] = (rj < n && k < n) ? (double)A[(int64_t)rj * ld + k] - mj : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kCK; ++k) {
      double av[4], bv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) av[a] = sa[k][ty * 4 + a];
#pragma unroll
      for (int b = 0; b < 4; ++b) bv[b] = sb[k][tx * 4 + b];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int i = bi * kCT + ty * 4 + a, j = bj * kCT + tx * 4 + b;
      if (i < n && j < n) {
        C[(int64_t)i * n + j] = acc[a][b];
        if (bi != bj) C[(int64_t)j * n + i] = acc[a][b];
      }
    }
}
// np.corrcoef's normalisation (numpy/lib/_function_base_impl.py: c /= stddev[:, None]; c /= stddev[None, :]; clip to [-1, 1])
// with c = cov / (n - 1); NaN (zero-variance rows) -> 0 as main.py:576 does; fp32 out
__global__ void __launch_bounds__(256) corr_finish_kernel(const double* __restrict__ C, int n, float* __restrict__ out, int64_t ldo) {
  const int i = blockIdx.x;
  const double fact = (double)(n - 1);
  const double di = sqrt(C[(int64_t)i * n + i] / fact);
  for (int j = threadIdx.x; j < n; j += 256) {
    const double dj = sqrt(C[(int64_t)j * n + j] / fact);
    double v = C[(int64_t)i * n + j] / fact;
    v /= di;
    v /= dj;
    v = fmin(fmax(v, -1.0), 1.0);             // np.clip propagates NaN; so do fmin / fmax only if written this way:
    if (!(di > 0.0) || !(dj > 0.0) || v != v) v = 0.0;
    out[(int64_t)i * ldo + j] = (float)v;
  }
}

// ---------------- z-score of the positive entries of each row ----------------
__device__ __forceinline__ double block_sum256(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < 8; ++w) t += sh[w];
  __syncthreads();
  return t;
}
__global__ void __launch_bounds__(256) zscore_rows_kernel(float* __restrict__ M, int64_t ld, int64_t ncols) {
  __shared__ double sh[8];
  float* row = M + (int64_t)blockIdx.x * ld;
  double s = 0.0, c = 0.0;
  for (int64_t j = threadIdx.x; j < ncols; j += 256) { const float v = row[j]; if (v > 0.f) { s += (double)v; c += 1.0; } }
  s = block_sum256(s, sh);
  c = block_sum256(c, sh);
  if (c == 0.0) {
    for (int64_t j = threadIdx.x; j < ncols; j += 256) { const float v = row[j]; if (v != v) row[j] = 0.f; }   // NaN -> 0 (:152)
    return;
  }
  const double mean = s / c;
  double q = 0.0;
  for (int64_t j = threadIdx.x; j < ncols; j += 256) { const float v = row[j]; if (v > 0.f) { const double d = (double)v - mean; q += d * d; } }
  q = block_sum256(q, sh);
  // scipy.stats.mstats.zscore works in the array's dtype (float32): (a - mean) / std with fp32 mean and std
  const float mean32 = (float)mean, std32 = (float)sqrt(q / c);
  for (int64_t j = threadIdx.x; j < ncols; j += 256) {
    const float v = row[j];
    if (v > 0.f) {
      const float z = __fdiv_rn(v - mean32, std32);
      row[j] = (z != z) ? 0.f : z;            // 0 / 0 (all positives equal) -> NaN -> 0; x / 0 -> inf stays, as in numpy
    } else if (v != v) {
      row[j] = 0.f;
    }
  }
}

// ---------------- adjacency builders (float64 accumulators, like np.zeros((N, N))) ----------------
__global__ void __launch_bounds__(256) adj_pixels_kernel(const int64_t* __restrict__ bin1, const int64_t* __restrict__ bin2,
                                                         const double* __restrict__ count, int64_t n_pix,
                                                         const int64_t* __restrict__ cool2node, int64_t n_cool,
                                                         const int32_t* __restrict__ node2chrom, int64_t N, double* __restrict__ intra,
                                                         double* __restrict__ inter) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_pix; p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i1 = bin1[p], i2 = bin2[p];
    if (i1 < 0 || i1 >= n_cool || i2 < 0 || i2 >= n_cool) continue;
    const int64_t a = cool2node[i1], b = cool2node[i2];           // 1-based node ids, <= 0 = bin not in chrom_list (:148-150)
    if (a <= 0 || b <= 0 || a > N || b > N) continue;
    const double c = count[p];
    if (c != c) continue;                                         // :156 NaN pixels (unbalanced bins) are skipped
    double* dst = node2chrom[a] == node2chrom[b] ? intra : inter;
    atomicAdd(dst + (a - 1) * N + (b - 1), c);
    atomicAdd(dst + (b - 1) * N + (a - 1), c);                    // a == b: the diagonal receives the count twice, as :161-162
  }
}
// one warp per cluster: all ordered pairs i != j (process.py:99-103)
__global__ void __launch_bounds__(256) adj_clusters_kernel(const int64_t* __restrict__ members, const int64_t* __restrict__ offsets,
                                                           int64_t n_clusters, int64_t N, double* __restrict__ adj) {
  const int lane = threadIdx.x & 31;
  const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t c = w0; c < n_clusters; c += nw) {
    const int64_t b = offsets[c], m = offsets[c + 1] - b;
    for (int64_t u = lane; u < m * m; u += 32) {
      const int64_t a = u / m, d = u - a * m;
      const int64_t i = members[b + a], j = members[b + d];
      if (i != j && i >= 1 && j >= 1 && i <= N && j <= N) atomicAdd(adj + (i - 1) * N + (j - 1), 1.0);
    }
  }
}
__global__ void __launch_bounds__(256) f64_to_f32_kernel(const double* __restrict__ in, float* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = (float)in[i];
}

int grid_for(int64_t n, int per_block = 256, int cap = kSMs * 8) {
  int64_t b = (n + per_block - 1) / per_block;
  if (b < 1) b = 1;
  return (int)(b > cap ? cap : b);
}

}  // namespace
}  // namespace matcha

using namespace matcha;

extern "C" {

int64_t matcha_corrcoef_workspace_bytes(int64_t n) { return n < 1 ? -1 : (n * n * 8 + 255) / 256 * 256 + (n * 8 + 255) / 256 * 256 + 256; }

int matcha_corrcoef(const float* A, int64_t lda, int64_t n, float* out, int64_t ldo, void* workspace, int64_t workspace_bytes,
                    void* stream) {
  MATCHA_REQUIRE(A && out && workspace && n >= 2 && lda >= n && ldo >= n, "matcha_corrcoef: bad arguments (n >= 2)");
  MATCHA_REQUIRE(n <= 65535 * 64, "matcha_corrcoef: n too large");
  MATCHA_REQUIRE(workspace_bytes >= matcha_corrcoef_workspace_bytes(n), "matcha_corrcoef: workspace too small: need %lld bytes",
                 (long long)matcha_corrcoef_workspace_bytes(n));
  cudaStream_t s = (cudaStream_t)stream;
  char* p = reinterpret_cast<char*>(((uintptr_t)workspace + 255) / 256 * 256);
  double* C = (double*)p;
  double* mean = (double*)(p + (n * n * 8 + 255) / 256 * 256);
  row_mean_kernel<<<(unsigned)n, 256, 0, s>>>(A, lda, (int)n, mean);
  MATCHA_CHECK_LAUNCH("row_mean");
  const int nt = (int)((n + kCT - 1) / kCT);
  cov_tile_kernel<<<dim3(nt, nt), 256, 0, s>>>(A, lda, (int)n, mean, C);
  MATCHA_CHECK_LAUNCH("cov_tile");
  corr_finish_kernel<<<(unsigned)n, 256, 0, s>>>(C, (int)n, out, ldo);
  MATCHA_CHECK_LAUNCH("corr_finish");
  return MATCHA_OK;
}

int matcha_zscore_positive_rows(float* M, int64_t ld, int64_t nrows, int64_t ncols, void* stream) {
  MATCHA_REQUIRE(M && nrows >= 0 && ncols >= 1 && ld >= ncols && nrows < (1ll << 31), "matcha_zscore_positive_rows: bad arguments");
  if (nrows == 0) return MATCHA_OK;
  zscore_rows_kernel<<<(unsigned)nrows, 256, 0, (cudaStream_t)stream>>>(M, ld, ncols);
  MATCHA_CHECK_LAUNCH("zscore_rows");
  return MATCHA_OK;
}

int matcha_adj_from_pixels(const int64_t* bin1, const int64_t* bin2, const double* count, int64_t n_pixels, const int64_t* cool2node,
                           int64_t n_cool_bins, const int32_t* node2chrom, int64_t n_nodes, double* intra, double* inter, void* stream) {
  MATCHA_REQUIRE(bin1 && bin2 && count && cool2node && node2chrom && intra && inter && n_pixels >= 0 && n_nodes >= 1,
                 "matcha_adj_from_pixels: bad arguments");
  if (n_pixels == 0) return MATCHA_OK;
  adj_pixels_kernel<<<grid_for(n_pixels), 256, 0, (cudaStream_t)stream>>>(bin1, bin2, count, n_pixels, cool2node, n_cool_bins, node2chrom,
                                                                         n_nodes, intra, inter);
  MATCHA_CHECK_LAUNCH("adj_pixels");
  return MATCHA_OK;
}

int matcha_adj_from_clusters(const int64_t* members, const int64_t* offsets, int64_t n_clusters, int64_t n_nodes, double* adj, void* stream) {
  MATCHA_REQUIRE(members && offsets && adj && n_clusters >= 0 && n_nodes >= 1, "matcha_adj_from_clusters: bad arguments");
  if (n_clusters == 0) return MATCHA_OK;
  adj_clusters_kernel<<<grid_for(n_clusters, 8), 256, 0, (cudaStream_t)stream>>>(members, offsets, n_clusters, n_nodes, adj);
  MATCHA_CHECK_LAUNCH("adj_clusters");
  return MATCHA_OK;
}

int matcha_f64_to_f32(const double* in, float* out, int64_t n, void* stream) {
  MATCHA_REQUIRE(in && out && n >= 0, "matcha_f64_to_f32: bad arguments");
  if (n == 0) return MATCHA_OK;
  f64_to_f32_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(in, out, n);
  MATCHA_CHECK_LAUNCH("f64_to_f32");
  return MATCHA_OK;
}

}  // extern "C"
