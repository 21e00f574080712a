// All-pairs (k = 2) scorer for denoise_contact.py:67-88.
//
// With two tokens and the diagonal masked, each token attends to the other with weight exactly 1, so
// the dynamic embedding of token i is a function of node j alone.  matcha_pair_tables (engine.cu)
// evaluates the network once per NODE (tables D, S of [N+1, d]); this kernel then streams the
// upper-triangular pair range, generating (i, j) on the device:
//   logit(i, j) = 0.5 * sum_c w_c [ (D[j,c] - S[i,c])^2 + (D[i,c] - S[j,c])^2 ] + b
// 64 x 64 pair tile per CTA, tables of the tile staged in shared memory (c-major), 4 x 4 pairs per thread.
#include "common.cuh"

namespace matcha {
namespace {

constexpr int PT = 64;   // pairs tile edge
constexpr int PD = 64;   // embed dim

__host__ __device__ __forceinline__ int64_t row_prefix(int64_t r, int64_t n, int64_t md) {
  // number of pairs in rows [0, r) when row q holds max(0, n - md - q) pairs
  const int64_t full = n - md;           // pairs in row 0
  if (full <= 0) return 0;
  if (r > full) r = full;
  return r * full - r * (r - 1) / 2;
}

__global__ void __launch_bounds__(256) pair_score_kernel(const float* __restrict__ D, const float* __restrict__ S,
                                                         const float* __restrict__ cls_w, const float* __restrict__ cls_b,
                                                         int64_t lo, int64_t n, int md, int64_t ti0, int64_t p_begin,
                                                         int64_t p_end, int apply_sigmoid, float* __restrict__ out) {
  const int64_t ti = ti0 + blockIdx.y, tj = blockIdx.x;
  // tile (ti, tj) holds pairs r in [ti*64, +64), c in [tj*64, +64) with c >= r + md
  if ((tj + 1) * PT - 1 < ti * PT + md) return;
  constexpr int PC = PD / 2;   // columns staged per pass (keeps static shared memory under 48 KB)
  __shared__ __align__(16) float sDi[PC][PT + 4], sSi[PC][PT + 4], sDj[PC][PT + 4], sSj[PC][PT + 4];
  __shared__ float sw[PD];
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  if (tid < PD) sw[tid] = __ldg(cls_w + tid);
  for (int pass = 0; pass < PD / PC; ++pass) {
    __syncthreads();
    for (int e = tid; e < PT * (PC / 4); e += 256) {
      const int node = e / (PC / 4), c4 = (e % (PC / 4)) * 4, gc = pass * PC + c4;
      const int64_t ri = ti * PT + node, rj = tj * PT + node;
      float4 di = make_float4(0.f, 0.f, 0.f, 0.f), si = di, dj = di, sj = di;
      if (ri < n) { di = __ldg(reinterpret_cast<const float4*>(D + (lo + ri) * PD + gc)); si = __ldg(reinterpret_cast<const float4*>(S + (lo + ri) * PD + gc)); }
      if (rj < n) { dj = __ldg(reinterpret_cast<const float4*>(D + (lo + rj) * PD + gc)); sj = __ldg(reinterpret_cast<const float4*>(S + (lo + rj) * PD + gc)); }
      sDi[c4 + 0][node] = di.x; sDi[c4 + 1][node] = di.y; sDi[c4 + 2][node] = di.z; sDi[c4 + 3][node] = di.w;
      sSi[c4 + 0][node] = si.x; sSi[c4 + 1][node] = si.y; sSi[c4 + 2][node] = si.z; sSi[c4 + 3][node] = si.w;
      sDj[c4 + 0][node] = dj.x; sDj[c4 + 1][node] = dj.y; sDj[c4 + 2][node] = dj.z; sDj[c4 + 3][node] = dj.w;
      sSj[c4 + 0][node] = sj.x; sSj[c4 + 1][node] = sj.y; sSj[c4 + 2][node] = sj.z; sSj[c4 + 3][node] = sj.w;
    }
    __syncthreads();
#pragma unroll 4
    for (int c = 0; c < PC; ++c) {
      const float w = sw[pass * PC + c];
      const float4 di = *reinterpret_cast<const float4*>(&sDi[c][ty * 4]);
      const float4 si = *reinterpret_cast<const float4*>(&sSi[c][ty * 4]);
      const float4 dj = *reinterpret_cast<const float4*>(&sDj[c][tx * 4]);
      const float4 sj = *reinterpret_cast<const float4*>(&sSj[c][tx * 4]);
      const float dI[4] = {di.x, di.y, di.z, di.w}, sI[4] = {si.x, si.y, si.z, si.w};
      const float dJ[4] = {dj.x, dj.y, dj.z, dj.w}, sJ[4] = {sj.x, sj.y, sj.z, sj.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const float t1 = dJ[b] - sI[a], t2 = dI[a] - sJ[b];
          acc[a][b] = fmaf(w * t1, t1, acc[a][b]);
          acc[a][b] = fmaf(w * t2, t2, acc[a][b]);
        }
    }
  }
  const float bias = __ldg(cls_b);
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int64_t r = ti * PT + ty * 4 + a;
    if (r >= n) continue;
    const int64_t rowbase = row_prefix(r, n, md) - (r + md);   // p(r, c) = rowbase + c
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int64_t c = tj * PT + tx * 4 + b;
      if (c >= n || c < r + md) continue;
      const int64_t p = rowbase + c;
      if (p < p_begin || p >= p_end) continue;
      float v = 0.5f * acc[a][b] + bias;
      if (apply_sigmoid) v = 1.0f / (1.0f + expf(-v));
      out[p - p_begin] = v;
    }
  }
}
}  // namespace
}  // namespace matcha

using namespace matcha;

extern "C" {

int64_t matcha_pair_count(int64_t lo, int64_t hi, int32_t min_dis) {
  const int64_t n = hi - lo;
  if (n <= 0) return 0;
  return row_prefix(n, n, min_dis < 0 ? 0 : min_dis);
}

int matcha_pair_score_range(const float* D, const float* S, const float* cls_w, const float* cls_b, int32_t d,
                            int64_t lo, int64_t hi, int32_t min_dis, int64_t p_begin, int64_t p_end,
                            int32_t apply_sigmoid, float* out, void* stream) {
  MATCHA_REQUIRE(D && S && cls_w && cls_b && out, "pair_score_range: NULL argument");
  if (d != PD) { set_error("pair_score_range: embed_dim %d unsupported (64)", d); return MATCHA_ERR_UNSUPPORTED; }
  MATCHA_REQUIRE(min_dis >= 0, "pair_score_range: negative min_distance");
  const int64_t n = hi - lo;
  const int64_t total = matcha_pair_count(lo, hi, min_dis);
  MATCHA_REQUIRE(p_begin >= 0 && p_end <= total && p_begin <= p_end, "pair_score_range: pair range [%lld, %lld) outside [0, %lld)",
                 (long long)p_begin, (long long)p_end, (long long)total);
  if (p_begin == p_end) return MATCHA_OK;
  // rows touched by [p_begin, p_end)
  auto row_of = [&](int64_t p) {
    int64_t a = 0, b = n;  // largest r with prefix(r) <= p
    while (b - a > 1) { int64_t mid = (a + b) / 2; if (row_prefix(mid, n, min_dis) <= p) a = mid; else b = mid; }
    return a;
  };
  const int64_t r0 = row_of(p_begin), r1 = row_of(p_end - 1);
  const int64_t ti0 = r0 / PT, ti1 = r1 / PT;
  const int64_t ntj = (n + PT - 1) / PT;
  int64_t done = ti0;
  int nk = 0;
  prof_begin(P_PAIR_SCORE, (cudaStream_t)stream);
  while (done <= ti1) {   // gridDim.y <= 65535
    const int64_t chunk = (ti1 - done + 1) > 32768 ? 32768 : (ti1 - done + 1);
    dim3 grid((unsigned)ntj, (unsigned)chunk);
    pair_score_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(D, S, cls_w, cls_b, lo, n, min_dis, done, p_begin, p_end,
                                                              apply_sigmoid, out);
    MATCHA_CHECK_LAUNCH("pair_score");
    done += chunk;
    ++nk;
  }
  prof_end(P_PAIR_SCORE, nk, (cudaStream_t)stream);
  return MATCHA_OK;
}

}  // extern "C"
