// Row-wise kernels of the Hyper-SAGNN hot path: everything that is not a dense contraction.
// Layout rule: one HALF-WARP owns one token row (64 floats = 16 lanes x float4, one coalesced 256 B
// access) or one hyperedge (its L <= 8 rows in turn); row reductions are 4 xor-shuffles.
#include <cuda_bf16.h>
#include <cooperative_groups.h>

#include "rowwise.cuh"

namespace matcha {
namespace {

constexpr int kBlock = 256;
constexpr float kLnEps = 1e-5f;

// sum over the D / 4 lanes that share one row (16 lanes for d = 64: a half-warp; 32 for d = 128: the warp)
template <int D>
__device__ __forceinline__ float redrow(float v) {
  if (D == 128) v += __shfl_xor_sync(0xffffffffu, v, 16);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float dot4(float4 a, float4 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w))); }
__device__ __forceinline__ float sum4(float4 a) { return (a.x + a.y) + (a.z + a.w); }
__device__ __forceinline__ float4 f4(float s) { return make_float4(s, s, s, s); }
__device__ __forceinline__ float4 operator+(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 operator-(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ float4 operator*(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 operator*(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float4 fma4(float s, float4 a, float4 c) {
  return make_float4(fmaf(s, a.x, c.x), fmaf(s, a.y, c.y), fmaf(s, a.z, c.z), fmaf(s, a.w, c.w));
}

struct LnOut { float4 xh; float rstd; };
// LayerNorm statistics of one D-wide row spread over D / 4 lanes (biased variance, eps 1e-5)
template <int D>
__device__ __forceinline__ LnOut ln_norm(float4 v) {
  const float mean = redrow<D>(sum4(v)) * (1.0f / D);
  const float4 c = v - f4(mean);
  const float var = redrow<D>(dot4(c, c)) * (1.0f / D);
  LnOut o;
  o.rstd = 1.0f / sqrtf(var + kLnEps);
  o.xh = c * o.rstd;
  return o;
}
// dx of y = LN(x) given gy = dy * gamma, normalised row xh and rstd
template <int D>
__device__ __forceinline__ float4 ln_bwd(float4 gy, float4 xh, float rstd) {
  const float m1 = redrow<D>(sum4(gy)) * (1.0f / D);
  const float m2 = redrow<D>(dot4(gy, xh)) * (1.0f / D);
  return (gy - f4(m1) - xh * m2) * rstd;
}

// block-level column reduction: every thread holds a float4 for columns [4*hl, 4*hl+4); add to dst[D]
template <int D>
__device__ __forceinline__ void block_colsum_atomic(float4 v, float* smemD, float* dst) {
  const int lane = threadIdx.x & 31, hl = threadIdx.x & (D / 4 - 1);
  if (D == 64) {
    v.x += __shfl_xor_sync(0xffffffffu, v.x, 16);
    v.y += __shfl_xor_sync(0xffffffffu, v.y, 16);
    v.z += __shfl_xor_sync(0xffffffffu, v.z, 16);
    v.w += __shfl_xor_sync(0xffffffffu, v.w, 16);
  }
  if (threadIdx.x < D) smemD[threadIdx.x] = 0.f;
  __syncthreads();
  if (lane < D / 4) {
    atomicAdd(&smemD[hl * 4 + 0], v.x);
    atomicAdd(&smemD[hl * 4 + 1], v.y);
    atomicAdd(&smemD[hl * 4 + 2], v.z);
    atomicAdd(&smemD[hl * 4 + 3], v.w);
  }
  __syncthreads();
  if (threadIdx.x < D && dst) atomicAdd(dst + threadIdx.x, smemD[threadIdx.x]);
  __syncthreads();
}

__device__ __forceinline__ int64_t num_token_tiles_dev(int64_t T) { return (T + kTileTok - 1) / kTileTok; }

inline int grid_for_rows(int64_t n, int d, int cap_blocks) {       // d / 4 lanes per row
  int64_t blocks = (n * (d / 4) + kBlock - 1) / kBlock;
  if (blocks < 1) blocks = 1;
  if (blocks > cap_blocks) blocks = cap_blocks;
  return (int)blocks;
}

// ------------------------------------------------------------------------------------------
// token bucketing by chromosome (replaces the per-chromosome mask/nonzero loop of Modules.py:180-188)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int chrom_of(const ChromMeta& cm, int64_t id) {
  if (id <= 0) return cm.n;  // pad bucket
  for (int c = 0; c < cm.n; ++c)
    if (id >= cm.start[c] && id < cm.end[c]) return c;
  return cm.n;               // ids outside every range behave like padding (zero encoder output)
}
// count + scan + scatter in ONE launch (was memset + two kernels, ~40 us of latency in front of every encoder pass).
// Every block histograms its contiguous token slice in shared memory and publishes the histogram; after a grid-wide
// barrier thread c of every block derives, for chromosome c, the group offset (totals of the chromosomes before it) plus
// the tokens that lower-numbered blocks place in it -- a private, contention-free range -- and the block scatters its
// tokens with shared-memory atomics.  The barrier is cooperative_groups' grid.sync() under cudaLaunchCooperativeKernel:
// the runtime guarantees co-residency of the whole grid (or fails the launch with an error code) whatever else runs on
// other streams, under MPS / MIG, or on a part with fewer SMs -- no hand-rolled spin, no trap.
// warp-aggregated shared-memory counter increment: lanes with the same bucket elect a leader that adds their count once;
// returns this lane's slot (old value + rank among its peers), -1 for inactive lanes.  Called by all 32 lanes.
__device__ __forceinline__ int32_t warp_agg_inc(int32_t* ctr, int c, bool valid, int lane) {
  const unsigned active = __ballot_sync(0xffffffffu, valid);
  if (!valid) return -1;
  const unsigned peers = __match_any_sync(active, c);
  const int leader = __ffs(peers) - 1;
  int32_t old = 0;
  if (lane == leader) old = atomicAdd(&ctr[c], __popc(peers));
  old = __shfl_sync(peers, old, leader);
  return old + __popc(peers & ((1u << lane) - 1u));
}
__global__ void __launch_bounds__(256) bucket_fused_kernel(const int64_t* __restrict__ x, int64_t T, const ChromMeta cm,
                                                           int32_t* __restrict__ counts, int32_t* __restrict__ group_off,
                                                           int32_t* __restrict__ perm, int32_t* __restrict__ hist) {
  __shared__ int32_t h[MATCHA_MAX_CHROM + 1], base[MATCHA_MAX_CHROM + 1], tot[MATCHA_MAX_CHROM + 1];
  const int nb = cm.n + 1;                                   // buckets: chromosomes + the pad bucket
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) { h[i] = 0; tot[i] = 0; base[i] = 0; }
  __syncthreads();
  const int64_t per = (T + gridDim.x - 1) / gridDim.x;
  const int64_t t0 = (int64_t)blockIdx.x * per, t1 = (t0 + per < T) ? t0 + per : T;
  const int iters = (int)((t1 - t0 + blockDim.x - 1) / blockDim.x);      // block-uniform trip count (warp collectives inside)
  // the bucket of each of this thread's tokens is kept for the scatter phase (slices are <= 8 tokens per thread except for
  // very long batches, which recompute)
  constexpr int kKeep = 8;
  int8_t mine[kKeep];
  for (int it = 0; it < iters; ++it) {
    const int64_t t = t0 + (int64_t)it * blockDim.x + threadIdx.x;
    const bool valid = t < t1;
    const int c = valid ? chrom_of(cm, x[t]) : 0;
    if (it < kKeep) mine[it] = (int8_t)c;
    warp_agg_inc(h, c, valid, lane);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nb; i += blockDim.x) hist[(int64_t)blockIdx.x * nb + i] = h[i];
  cooperative_groups::this_grid().sync();
  // totals per bucket and the tokens that lower-numbered blocks place in each bucket: all threads stream the published
  // histograms (independent coalesced loads) into shared-memory sums
  for (int i = threadIdx.x; i < (int)gridDim.x * nb; i += blockDim.x) {
    const int32_t v = __ldcg(hist + i);
    if (v != 0) {
      const int bk = i / nb, c = i - bk * nb;
      atomicAdd(&tot[c], v);
      if (bk < (int)blockIdx.x) atomicAdd(&base[c], v);
    }
  }
  __syncthreads();
  if (threadIdx.x < nb) {
    const int c = threadIdx.x;
    int32_t off = 0;
    for (int cc = 0; cc < c; ++cc) off += tot[cc];
    h[c] = base[c] + off;                                   // this block's first slot of bucket c
    if (blockIdx.x == 0) { counts[c] = tot[c]; group_off[c] = off; }
  }
  __syncthreads();
  for (int it = 0; it < iters; ++it) {
    const int64_t t = t0 + (int64_t)it * blockDim.x + threadIdx.x;
    const bool valid = t < t1;
    int c = 0;
    if (valid) c = it < kKeep ? (int)mine[it] : chrom_of(cm, x[t]);
    const bool real = valid && c < cm.n;                    // pads are counted, not listed
    const int32_t slot = warp_agg_inc(h, c, real, lane);
    if (real) perm[slot] = (int32_t)t;
  }
}
__global__ void active_flags_kernel(const int32_t* counts, int n_chrom, int rchrom, int64_t T, int32_t* active) {
  int c = threadIdx.x;
  if (c < n_chrom) {
    active[c] = counts[c] > 0;
    int64_t elig = T - counts[n_chrom] - (rchrom >= 0 ? counts[rchrom] : 0);
    active[n_chrom + c] = (c == rchrom && elig > 0);
  }
}

// ------------------------------------------------------------------------------------------
// LayerNorm statistics of X (shared by MHA layer_norm1/2/3 and Classifier.layer_norm2: same input row)
// ------------------------------------------------------------------------------------------
// fp32 x4 -> 4 bf16 hi (8 B) + 4 bf16 lo (8 B), lo = bf16(x - float(hi))
__device__ __forceinline__ void split4(float4 v, uint2& hi, uint2& lo) {
  const float x[4] = {v.x, v.y, v.z, v.w};
  uint32_t h[2], l[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(x[2 * i]), h1 = __float2bfloat16_rn(x[2 * i + 1]);
    const __nv_bfloat16 l0 = __float2bfloat16_rn(x[2 * i] - __bfloat162float(h0));
    const __nv_bfloat16 l1 = __float2bfloat16_rn(x[2 * i + 1] - __bfloat162float(h1));
    h[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    l[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
  }
  hi = make_uint2(h[0], h[1]);
  lo = make_uint2(l[0], l[1]);
}
// byte offset, inside a tile half, of columns [4*hl, 4*hl+4) of token `tok` (0..127)
__device__ __forceinline__ int tile_off(int hl, int tok) { return (hl >> 1) * kPlaneBytes + tok * 16 + (hl & 1) * 8; }

template <int D>
__global__ void __launch_bounds__(kBlock) ln_fwd_kernel(const float* __restrict__ X, float* __restrict__ xhat,
                                                         float* __restrict__ rstd, int64_t T, uint8_t* __restrict__ xt) {
  const int hl = threadIdx.x & (D / 4 - 1);
  const int64_t hw0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / (D / 4), nhw = ((int64_t)gridDim.x * blockDim.x) / (D / 4);
  const int64_t Tp = xt ? num_token_tiles_dev(T) * kTileTok : T;     // tile path: also zero the tail rows of the last tile
  const int64_t iters = (Tp + nhw - 1) / nhw;
  for (int64_t it = 0; it < iters; ++it) {
    const int64_t t = hw0 + it * nhw;
    const bool valid = t < T;
    const int64_t tt = valid ? t : T - 1;
    LnOut o = ln_norm<D>(ldg4(X + tt * D + hl * 4));
    if (valid) {
      st4(xhat + t * D + hl * 4, o.xh);
      if (hl == 0) rstd[t] = o.rstd;
    }
    if (xt && t < Tp) {
      uint8_t* tile = xt + (t / kTileTok) * (int64_t)kXTileBytes;
      const int tok = (int)(t % kTileTok);
      uint2 hi = make_uint2(0u, 0u), lo = make_uint2(0u, 0u);
      if (valid) split4(o.xh, hi, lo);
      *reinterpret_cast<uint2*>(tile + tile_off(hl, tok)) = hi;
      *reinterpret_cast<uint2*>(tile + kXHalfBytes + tile_off(hl, tok)) = lo;
      if (hl < 4) {   // planes 8 (ones column) and 9 (zeros) of both halves: 16 bytes each per token
        uint4 v = make_uint4((hl == 0 && valid) ? 0x00003F80u : 0u, 0u, 0u, 0u);     // bf16(1.0) in column 64
        uint8_t* dst = tile + (hl >> 1) * kXHalfBytes + (8 + (hl & 1)) * kPlaneBytes + tok * 16;
        *reinterpret_cast<uint4*>(dst) = v;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// attention over the L tokens of one hyperedge (Modules.py:448-460, 563-572 with fc1 folded into G)
//   S_h[i][j] = Q_h[i].K_h[j]  (1/sqrt(d) and LayerNorm affine folded into the projection),
//   diagonal masked, softmax over j (pads are live keys), dyn_i = b_dyn + sum_h sum_j A_h[i][j] G_h[j]
// ------------------------------------------------------------------------------------------
template <int L>
__device__ __forceinline__ void softmax_offdiag(float (&A)[L][L]) {
#pragma unroll
  for (int i = 0; i < L; ++i) {
    float mx = -3.0e38f;
#pragma unroll
    for (int j = 0; j < L; ++j) if (j != i) mx = fmaxf(mx, A[i][j]);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < L; ++j) if (j != i) { A[i][j] = expf(A[i][j] - mx); sum += A[i][j]; }
    const float inv = 1.0f / sum;
#pragma unroll
    for (int j = 0; j < L; ++j) A[i][j] = (j != i) ? A[i][j] * inv : 0.f;
  }
}

template <int D, int L>
__global__ void __launch_bounds__(kBlock) attn_fwd_kernel(const float* __restrict__ QKG, const int64_t* __restrict__ x,
                                                           const float* __restrict__ b_dyn, float* __restrict__ U,
                                                           int64_t B, const DropCfg drop) {
  const int hl = threadIdx.x & (D / 4 - 1);
  const int64_t hw0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / (D / 4), nhw = ((int64_t)gridDim.x * blockDim.x) / (D / 4);
  const int64_t iters = (B + nhw - 1) / nhw;
  const float4 bd = ldg4(b_dyn + hl * 4);
  for (int64_t it = 0; it < iters; ++it) {
    const int64_t b = hw0 + it * nhw;
    const bool valid = b < B;
    const int64_t bb = valid ? b : B - 1;
    const float* base = QKG + bb * L * (3 * kH * D) + hl * 4;
    float4 o[L];
#pragma unroll
    for (int i = 0; i < L; ++i) o[i] = bd;
#pragma unroll 1
    for (int h = 0; h < kH; ++h) {
      float4 q[L], k[L];
#pragma unroll
      for (int i = 0; i < L; ++i) {
        q[i] = ldg4(base + i * (3 * kH * D) + h * D);
        k[i] = ldg4(base + i * (3 * kH * D) + kH * D + h * D);
      }
      float A[L][L];
#pragma unroll
      for (int i = 0; i < L; ++i)
#pragma unroll
        for (int j = 0; j < L; ++j) A[i][j] = (i != j) ? redrow<D>(dot4(q[i], k[j])) : 0.f;
      softmax_offdiag<L>(A);
#pragma unroll
      for (int j = 0; j < L; ++j) {
        const float4 g = ldg4(base + j * (3 * kH * D) + 2 * kH * D + h * D);
#pragma unroll
        for (int i = 0; i < L; ++i) if (i != j) o[i] = fma4(A[i][j], g, o[i]);
      }
    }
    if (valid) {
#pragma unroll
      for (int i = 0; i < L; ++i) {
        const int64_t t = b * L + i;
        const float m = x[t] != 0 ? 1.f : 0.f;                       // non_pad_mask (Modules.py:614)
        float4 v = drop_apply4(drop, (uint64_t)t, (uint32_t)(hl * 4), o[i]);   // dropout after fc1 (:572)
        st4(U + t * D + hl * 4, v * m);
      }
    }
  }
}

// store 4 consecutive gradient columns of token t: fp32 row-major, or pre-split tile (chunk fc, columns 4*hl..)
template <int D, bool SPLIT>
__device__ __forceinline__ void store_dqkg(float* dQKG, uint8_t* gt, int64_t t, int fc, int hl, float4 v) {
  if (!SPLIT) {
    st4(dQKG + t * (3 * kH * D) + fc * D + hl * 4, v);
  } else {
    uint2 hi, lo;
    split4(v, hi, lo);
    uint8_t* tile = gt + ((t / kTileTok) * kGChunks + fc) * (int64_t)kGTileBytes;
    const int off = tile_off(hl, (int)(t % kTileTok));
    *reinterpret_cast<uint2*>(tile + off) = hi;
    *reinterpret_cast<uint2*>(tile + kGHalfBytes + off) = lo;
  }
}

template <int D, int L, bool SPLIT>
__global__ void __launch_bounds__(kBlock) attn_bwd_kernel(const float* __restrict__ QKG, const float* __restrict__ dU,
                                                           const int64_t* __restrict__ x, float* __restrict__ dQKG,
                                                           uint8_t* __restrict__ gt, float* __restrict__ db_dyn, int64_t B,
                                                           const DropCfg drop) {
  __shared__ float red[D];
  const int hl = threadIdx.x & (D / 4 - 1);
  const int64_t hw0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / (D / 4), nhw = ((int64_t)gridDim.x * blockDim.x) / (D / 4);
  const int64_t iters = (B + nhw - 1) / nhw;
  float4 acc_b = f4(0.f);
  for (int64_t it = 0; it < iters; ++it) {
    const int64_t b = hw0 + it * nhw;
    const bool valid = b < B;
    const int64_t bb = valid ? b : B - 1;
    const float* base = QKG + bb * L * (3 * kH * D) + hl * 4;
    float4 dd[L];
#pragma unroll
    for (int i = 0; i < L; ++i) {
      const int64_t t = bb * L + i;
      const float m = x[t] != 0 ? 1.f : 0.f;
      dd[i] = ldg4(dU + t * D + hl * 4) * drop_factor4(drop, (uint64_t)t, (uint32_t)(hl * 4)) * m;
      if (valid) acc_b = acc_b + dd[i];
    }
#pragma unroll 1
    for (int h = 0; h < kH; ++h) {
      float4 q[L], k[L], g[L];
#pragma unroll
      for (int i = 0; i < L; ++i) {
        q[i] = ldg4(base + i * (3 * kH * D) + h * D);
        k[i] = ldg4(base + i * (3 * kH * D) + kH * D + h * D);
        g[i] = ldg4(base + i * (3 * kH * D) + 2 * kH * D + h * D);
      }
      float A[L][L], dA[L][L];
#pragma unroll
      for (int i = 0; i < L; ++i)
#pragma unroll
        for (int j = 0; j < L; ++j) {
          A[i][j] = (i != j) ? redrow<D>(dot4(q[i], k[j])) : 0.f;
          dA[i][j] = (i != j) ? redrow<D>(dot4(dd[i], g[j])) : 0.f;
        }
      softmax_offdiag<L>(A);
      // dG_h[j] = sum_i A[i][j] * ddyn_i
#pragma unroll
      for (int j = 0; j < L; ++j) {
        float4 dg = f4(0.f);
#pragma unroll
        for (int i = 0; i < L; ++i) if (i != j) dg = fma4(A[i][j], dd[i], dg);
        if (valid) store_dqkg<D, SPLIT>(dQKG, gt, bb * L + j, 2 * kH + h, hl, dg);
      }
      // softmax backward -> dS (stored in dA)
#pragma unroll
      for (int i = 0; i < L; ++i) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < L; ++j) s = fmaf(A[i][j], dA[i][j], s);
#pragma unroll
        for (int j = 0; j < L; ++j) dA[i][j] = A[i][j] * (dA[i][j] - s);
      }
#pragma unroll
      for (int i = 0; i < L; ++i) {
        float4 dq = f4(0.f), dk = f4(0.f);
#pragma unroll
        for (int j = 0; j < L; ++j) if (j != i) { dq = fma4(dA[i][j], k[j], dq); dk = fma4(dA[j][i], q[j], dk); }
        if (valid) {
          store_dqkg<D, SPLIT>(dQKG, gt, bb * L + i, h, hl, dq);
          store_dqkg<D, SPLIT>(dQKG, gt, bb * L + i, kH + h, hl, dk);
        }
      }
    }
  }
  block_colsum_atomic<D>(acc_b, red, db_dyn);
}

// ------------------------------------------------------------------------------------------
// scorer: pff_n1 LayerNorm + mask, Classifier.layer_norm1/2, (dyn - static)^2, Conv1d(d->1), masked mean
// (Modules.py:373-374, 614, 290-311)
// ------------------------------------------------------------------------------------------
struct ScoreVec { float4 gp, bp, g1, b1, g2, b2, w; float cb; };
__device__ __forceinline__ ScoreVec load_score_params(const ScoreParams& p, int hl) {
  ScoreVec v;
  v.gp = ldg4(p.pff_g + hl * 4); v.bp = ldg4(p.pff_b + hl * 4);
  v.g1 = ldg4(p.ln1_g + hl * 4); v.b1 = ldg4(p.ln1_b + hl * 4);
  v.g2 = ldg4(p.ln2_g + hl * 4); v.b2 = ldg4(p.ln2_b + hl * 4);
  v.w = ldg4(p.cls_w + hl * 4);  v.cb = __ldg(p.cls_b);
  return v;
}

template <int D, int L>
__global__ void __launch_bounds__(kBlock) score_fwd_kernel(const float* __restrict__ H2, const float* __restrict__ xhat,
                                                            const int64_t* __restrict__ x, const ScoreParams p,
                                                            float* __restrict__ logits, int64_t B) {
  const int hl = threadIdx.x & (D / 4 - 1);
  const int64_t hw0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / (D / 4), nhw = ((int64_t)gridDim.x * blockDim.x) / (D / 4);
  const int64_t iters = (B + nhw - 1) / nhw;
  const ScoreVec sv = load_score_params(p, hl);
  for (int64_t it = 0; it < iters; ++it) {
    const int64_t b = hw0 + it * nhw;
    const bool valid = b < B;
    const int64_t bb = valid ? b : B - 1;
    float zsum = 0.f, msum = 0.f;
#pragma unroll
    for (int i = 0; i < L; ++i) {
      const int64_t t = bb * L + i;
      const float m = x[t] != 0 ? 1.f : 0.f;
      LnOut a = ln_norm<D>(ldg4(H2 + t * D + hl * 4));
      const float4 dyn2 = (a.xh * sv.gp + sv.bp) * m;
      LnOut c = ln_norm<D>(dyn2);
      const float4 Dv = c.xh * sv.g1 + sv.b1;
      const float4 S = ldg4(xhat + t * D + hl * 4) * sv.g2 + sv.b2;
      const float4 df = Dv - S;
      const float z = redrow<D>(dot4(df * df, sv.w)) + sv.cb;
      zsum += z * m;
      msum += m;
    }
    if (valid && hl == 0) logits[b] = zsum / (msum + 1e-15f);
  }
}

template <int D, int L>
__global__ void __launch_bounds__(kBlock) score_bwd_kernel(const float* __restrict__ H2, const float* __restrict__ xhat,
                                                            const float* __restrict__ rstd_x, const int64_t* __restrict__ x,
                                                            const ScoreParams p, const float* __restrict__ dlogit,
                                                            float* __restrict__ dH2, float* __restrict__ dXs,
                                                            const ScoreGrads g, int64_t B) {
  __shared__ float red[D];
  const int hl = threadIdx.x & (D / 4 - 1);
  const int64_t hw0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / (D / 4), nhw = ((int64_t)gridDim.x * blockDim.x) / (D / 4);
  const int64_t iters = (B + nhw - 1) / nhw;
  const ScoreVec sv = load_score_params(p, hl);
  float4 a_gp = f4(0.f), a_bp = f4(0.f), a_g1 = f4(0.f), a_b1 = f4(0.f), a_g2 = f4(0.f), a_b2 = f4(0.f), a_w = f4(0.f);
  float a_cb = 0.f;
  for (int64_t it = 0; it < iters; ++it) {
    const int64_t b = hw0 + it * nhw;
    const bool valid = b < B;
    const int64_t bb = valid ? b : B - 1;
    float msum = 0.f;
#pragma unroll
    for (int i = 0; i < L; ++i) msum += x[bb * L + i] != 0 ? 1.f : 0.f;
    const float dl = valid ? dlogit[bb] / (msum + 1e-15f) : 0.f;
#pragma unroll
    for (int i = 0; i < L; ++i) {
      const int64_t t = bb * L + i;
      const float m = x[t] != 0 ? 1.f : 0.f;
      const float dz = dl * m;
      LnOut a = ln_norm<D>(ldg4(H2 + t * D + hl * 4));
      const float4 dyn2 = (a.xh * sv.gp + sv.bp) * m;
      LnOut c = ln_norm<D>(dyn2);
      const float4 Dv = c.xh * sv.g1 + sv.b1;
      const float4 xh = ldg4(xhat + t * D + hl * 4);
      const float4 S = xh * sv.g2 + sv.b2;
      const float4 df = Dv - S;
      if (hl == 0) a_cb += dz;
      a_w = a_w + df * df * dz;
      const float4 dD = df * sv.w * (2.f * dz);
      // Classifier.layer_norm1 backward
      a_g1 = a_g1 + dD * c.xh; a_b1 = a_b1 + dD;
      float4 ddyn2 = ln_bwd<D>(dD * sv.g1, c.xh, c.rstd) * m;
      // pff_n1.layer_norm backward
      a_gp = a_gp + ddyn2 * a.xh; a_bp = a_bp + ddyn2;
      const float4 dh2 = ln_bwd<D>(ddyn2 * sv.gp, a.xh, a.rstd);
      // Classifier.layer_norm2 (static branch) backward, directly w.r.t. X
      const float4 dS = f4(0.f) - dD;
      a_g2 = a_g2 + dS * xh; a_b2 = a_b2 + dS;
      const float4 dxs = ln_bwd<D>(dS * sv.g2, xh, rstd_x[t]);
      if (valid) { st4(dH2 + t * D + hl * 4, dh2); st4(dXs + t * D + hl * 4, dxs); }
    }
  }
  block_colsum_atomic<D>(a_gp, red, g.pff_g); block_colsum_atomic<D>(a_bp, red, g.pff_b);
  block_colsum_atomic<D>(a_g1, red, g.ln1_g); block_colsum_atomic<D>(a_b1, red, g.ln1_b);
  block_colsum_atomic<D>(a_g2, red, g.ln2_g); block_colsum_atomic<D>(a_b2, red, g.ln2_b);
  block_colsum_atomic<D>(a_w, red, g.cls_w);
  // scalar bias gradient
  float s = a_cb;
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0 && s != 0.f) atomicAdd(g.cls_b, s);
}

// ------------------------------------------------------------------------------------------
// losses
// ------------------------------------------------------------------------------------------
__global__ void bce_kernel(const float* __restrict__ logits, const float* __restrict__ y, const float* __restrict__ w,
                           float alpha, float* __restrict__ dlogit, float* __restrict__ loss_out, int64_t B) {
  __shared__ float red[32];
  float part = 0.f;
  const float invB = 1.0f / (float)B;
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (int64_t)gridDim.x * blockDim.x) {
    const float z = logits[b], yy = y[b], ww = w[b];
    // F.binary_cross_entropy_with_logits(pred, y, weight=w) (main.py:56)
    part += ww * (fmaxf(z, 0.f) - z * yy + log1pf(expf(-fabsf(z))));
    const float sig = 1.0f / (1.0f + expf(-z));
    dlogit[b] = alpha * ww * (sig - yy) * invB;
  }
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) atomicAdd(loss_out, v * invB);
  }
}
__global__ void finalize_loss_kernel(float* loss_out, const float* recon, float alpha, float beta) {
  const float r = recon ? recon[0] : 0.f;
  loss_out[1] = r;
  loss_out[2] = alpha * loss_out[0] + beta * r;
}

// recon target gather + squared error (Modules.py:194-199); one warp per token row
__global__ void __launch_bounds__(kBlock) recon_diff_kernel(float* __restrict__ pred, int64_t ld, const int64_t* __restrict__ x,
                                                             int64_t T, const float* __restrict__ inter, int64_t inter_ld,
                                                             int64_t r_start, int64_t r_end, const int32_t* __restrict__ counts,
                                                             int rchrom, int n_chrom, float* __restrict__ recon_out) {
  __shared__ float red[kBlock / 32];
  const int lane = threadIdx.x & 31;
  const int64_t n_r = r_end - r_start;
  const int64_t elig = T - counts[n_chrom] - counts[rchrom];
  const float gscale = elig > 0 ? 200.0f / ((float)elig * (float)n_r) : 0.f;
  float part = 0.f;
  for (int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < T; t += ((int64_t)gridDim.x * blockDim.x) >> 5) {
    const int64_t id = x[t];
    const bool ok = id != 0 && (id < r_start || id >= r_end);
    float* prow = pred + t * ld;
    const float* trow = inter + (id - 1) * inter_ld + (r_start - 1);
    for (int64_t j = lane; j < n_r; j += 32) {
      float dv = 0.f;
      if (ok) { dv = prow[j] - __ldg(trow + j); part = fmaf(dv, dv, part); }
      prow[j] = dv * gscale;
    }
  }
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if (lane == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
    for (int i = 0; i < kBlock / 32; ++i) v += red[i];
    if (elig > 0 && v != 0.f) atomicAdd(recon_out, v * 100.0f / ((float)elig * (float)n_r));
  }
}

// ------------------------------------------------------------------------------------------
// small backward helpers
// ------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(kBlock) ln_tanh_bwd_kernel(const float* __restrict__ dxhat, int nparts, int64_t part_stride,
                                                              const float* __restrict__ dXs,
                                                              const float* __restrict__ xhat, const float* __restrict__ rstd,
                                                              const float* __restrict__ X, float* __restrict__ dP, int64_t T) {
  const int hl = threadIdx.x & (D / 4 - 1);
  const int64_t hw0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / (D / 4), nhw = ((int64_t)gridDim.x * blockDim.x) / (D / 4);
  const int64_t iters = (T + nhw - 1) / nhw;
  for (int64_t it = 0; it < iters; ++it) {
    const int64_t t = hw0 + it * nhw;
    const bool valid = t < T;
    const int64_t tt = valid ? t : T - 1;
    const float4 xh = ldg4(xhat + tt * D + hl * 4);
    float4 gy = ldg4(dxhat + tt * D + hl * 4);
    for (int p = 1; p < nparts; ++p) gy = gy + ldg4(dxhat + p * part_stride + tt * D + hl * 4);
    float4 dx = ln_bwd<D>(gy, xh, rstd[tt]) + ldg4(dXs + tt * D + hl * 4);
    const float4 xv = ldg4(X + tt * D + hl * 4);
    dx = dx * (f4(1.f) - xv * xv);                      // X = tanh(P)  (Modules.py:270)
    if (valid) st4(dP + t * D + hl * 4, dx);
  }
}
__global__ void enc_combine_bwd_kernel(const float4* __restrict__ dV0, const float4* __restrict__ dtE,
                                       const float4* __restrict__ E, float beta, float4* __restrict__ dE, int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = dV0[i];
    if (dtE) {
      const float4 e = E[i];
      const float4 te = make_float4(tanhf(e.x), tanhf(e.y), tanhf(e.z), tanhf(e.w));
      v = v + dtE[i] * (f4(1.f) - te * te) * beta;
    }
    dE[i] = v;
  }
}

// ------------------------------------------------------------------------------------------
// k = 2 closed-form tables
// ------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(kBlock) pair_u_kernel(const float* __restrict__ QKG, const float* __restrict__ b_dyn,
                                                         float* __restrict__ U, int64_t T) {
  const int hl = threadIdx.x & (D / 4 - 1);
  const int64_t hw0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / (D / 4), nhw = ((int64_t)gridDim.x * blockDim.x) / (D / 4);
  for (int64_t t = hw0; t < T; t += nhw) {
    float4 o = ldg4(b_dyn + hl * 4);
#pragma unroll
    for (int h = 0; h < kH; ++h) o = o + ldg4(QKG + t * (3 * kH * D) + 2 * kH * D + h * D + hl * 4);
    st4(U + t * D + hl * 4, o);
  }
}
template <int D>
__global__ void __launch_bounds__(kBlock) pair_ds_kernel(const float* __restrict__ H2, const float* __restrict__ xhat,
                                                          const ScoreParams p, float* __restrict__ Dt, float* __restrict__ St,
                                                          int64_t T) {
  const int hl = threadIdx.x & (D / 4 - 1);
  const int64_t hw0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / (D / 4), nhw = ((int64_t)gridDim.x * blockDim.x) / (D / 4);
  const int64_t iters = (T + nhw - 1) / nhw;
  const ScoreVec sv = load_score_params(p, hl);
  for (int64_t it = 0; it < iters; ++it) {
    const int64_t t = hw0 + it * nhw;
    const bool valid = t < T;
    const int64_t tt = valid ? t : T - 1;
    LnOut a = ln_norm<D>(ldg4(H2 + tt * D + hl * 4));
    LnOut c = ln_norm<D>(a.xh * sv.gp + sv.bp);
    if (valid) {
      st4(Dt + t * D + hl * 4, c.xh * sv.g1 + sv.b1);
      st4(St + t * D + hl * 4, ldg4(xhat + t * D + hl * 4) * sv.g2 + sv.b2);
    }
  }
}

__global__ void iota_i64_kernel(int64_t* out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = i;
}

}  // namespace

// ==========================================================================================
// launchers
// ==========================================================================================
int bucket_hist_ints() { return kSMs * (MATCHA_MAX_CHROM + 1); }
// counts [n + 1], group_off [n + 1], perm [T], hist: bucket_hist_ints() ints of scratch
int launch_bucket(const int64_t* x, int64_t T, const ChromMeta& cm, int32_t* counts, int32_t* group_off,
                  int32_t* hist, int32_t* perm, cudaStream_t s) {
  // grid limit of the cooperative launch: what the device can keep resident at once (queried, not assumed), capped by
  // the histogram scratch (one row per block)
  static int max_blocks[16] = {};
  int dev = 0;
  if (int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return rc;
  if (dev < 0 || dev >= 16) { set_error("launch_bucket: device index %d out of range", dev); return MATCHA_ERR_ARG; }
  if (max_blocks[dev] == 0) {
    int per_sm = 0, sms = 0, coop = 0;
    if (int rc = check_cuda(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev), "cudaDeviceGetAttribute")) return rc;
    if (!coop) { set_error("launch_bucket: device %d does not support cooperative launches", dev); return MATCHA_ERR_UNSUPPORTED; }
    if (int rc = check_cuda(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bucket_fused_kernel, 256, 0), "occupancy")) return rc;
    if (int rc = check_cuda(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "cudaDeviceGetAttribute")) return rc;
    int lim = per_sm * sms;
    if (lim < 1) { set_error("launch_bucket: the bucketing kernel does not fit on an SM"); return MATCHA_ERR_CUDA; }
    max_blocks[dev] = lim < kSMs ? lim : kSMs;
  }
  int blocks = (int)((T + 1023) / 1024);
  if (blocks < 1) blocks = 1;
  if (blocks > max_blocks[dev]) blocks = max_blocks[dev];
  ChromMeta cm_arg = cm;
  void* args[] = {(void*)&x, (void*)&T, (void*)&cm_arg, (void*)&counts, (void*)&group_off, (void*)&perm, (void*)&hist};
  if (int rc = check_cuda(cudaLaunchCooperativeKernel((const void*)bucket_fused_kernel, dim3(blocks), dim3(256), args, 0, s),
                          "cudaLaunchCooperativeKernel(bucket_fused)")) return rc;
  return MATCHA_OK;
}

int launch_active_flags(const int32_t* counts, int n_chrom, int rchrom, int64_t T, int32_t* active, cudaStream_t s) {
  active_flags_kernel<<<1, MATCHA_MAX_CHROM, 0, s>>>(counts, n_chrom, rchrom, T, active);
  MATCHA_CHECK_LAUNCH("active_flags");
  return MATCHA_OK;
}


#define MATCHA_DISPATCH_D(d, CALL)                                              \
  if ((d) == 64) { constexpr int DD = 64; CALL; }                               \
  else if ((d) == 128) { constexpr int DD = 128; CALL; }                        \
  else { set_error("embed_dim %d unsupported (64, 128)", (int)(d)); return MATCHA_ERR_UNSUPPORTED; }
#define MATCHA_DISPATCH_L(L, CALL)                                              \
  switch (L) {                                                                  \
    case 2: { constexpr int LL = 2; CALL; } break;                              \
    case 3: { constexpr int LL = 3; CALL; } break;                              \
    case 4: { constexpr int LL = 4; CALL; } break;                              \
    case 5: { constexpr int LL = 5; CALL; } break;                              \
    case 6: { constexpr int LL = 6; CALL; } break;                              \
    default:                                                                    \
      set_error("hyperedge width L=%d unsupported (2..6)", (int)(L));           \
      return MATCHA_ERR_UNSUPPORTED;                                            \
  }

int launch_ln_fwd(int d, const float* X, float* xhat, float* rstd, int64_t T, uint8_t* xhat_tiles, cudaStream_t s) {
  if (T <= 0) return MATCHA_OK;
  if (xhat_tiles && d != 64) { set_error("ln_fwd: pre-split tiles need embed_dim 64"); return MATCHA_ERR_UNSUPPORTED; }
  MATCHA_DISPATCH_D(d, (ln_fwd_kernel<DD><<<grid_for_rows(T, d, kSMs * 16), kBlock, 0, s>>>(X, xhat, rstd, T, xhat_tiles)));
  MATCHA_CHECK_LAUNCH("ln_fwd");
  return MATCHA_OK;
}

int launch_attn_fwd(int d, const float* QKG, const int64_t* x, const float* b_dyn, float* U, int64_t B, int L, DropCfg drop,
                    cudaStream_t s) {
  if (B <= 0) return MATCHA_OK;
  const int grid = grid_for_rows(B, d, kSMs * 8);
  MATCHA_DISPATCH_D(d, MATCHA_DISPATCH_L(L, (attn_fwd_kernel<DD, LL><<<grid, kBlock, 0, s>>>(QKG, x, b_dyn, U, B, drop))));
  MATCHA_CHECK_LAUNCH("attn_fwd");
  return MATCHA_OK;
}
int launch_attn_bwd(int d, const float* QKG, const float* dU, const int64_t* x, float* dQKG, uint8_t* dqkg_tiles, float* db_dyn,
                    int64_t B, int L, DropCfg drop, cudaStream_t s) {
  if (B <= 0) return MATCHA_OK;
  const int grid = grid_for_rows(B, d, kSMs * 4);
  if (dqkg_tiles) {
    if (d != 64) { set_error("attn_bwd: pre-split tiles need embed_dim 64"); return MATCHA_ERR_UNSUPPORTED; }
    MATCHA_DISPATCH_L(L, (attn_bwd_kernel<64, LL, true><<<grid, kBlock, 0, s>>>(QKG, dU, x, dQKG, dqkg_tiles, db_dyn, B, drop)));
  } else {
    MATCHA_DISPATCH_D(d, MATCHA_DISPATCH_L(L, (attn_bwd_kernel<DD, LL, false><<<grid, kBlock, 0, s>>>(QKG, dU, x, dQKG, dqkg_tiles, db_dyn, B, drop))));
  }
  MATCHA_CHECK_LAUNCH("attn_bwd");
  return MATCHA_OK;
}
int launch_score_fwd(int d, const float* H2, const float* xhat, const int64_t* x, ScoreParams p, float* logits, int64_t B, int L,
                     cudaStream_t s) {
  if (B <= 0) return MATCHA_OK;
  const int grid = grid_for_rows(B, d, kSMs * 8);
  MATCHA_DISPATCH_D(d, MATCHA_DISPATCH_L(L, (score_fwd_kernel<DD, LL><<<grid, kBlock, 0, s>>>(H2, xhat, x, p, logits, B))));
  MATCHA_CHECK_LAUNCH("score_fwd");
  return MATCHA_OK;
}
int launch_score_bwd(int d, const float* H2, const float* xhat, const float* rstd_x, const int64_t* x, ScoreParams p,
                     const float* dlogit, float* dH2, float* dXs, ScoreGrads g, int64_t B, int L, cudaStream_t s) {
  if (B <= 0) return MATCHA_OK;
  const int grid = grid_for_rows(B, d, kSMs * 2);      // fewer blocks: fewer contended parameter-gradient atomics
  MATCHA_DISPATCH_D(d, MATCHA_DISPATCH_L(L, (score_bwd_kernel<DD, LL><<<grid, kBlock, 0, s>>>(H2, xhat, rstd_x, x, p, dlogit, dH2, dXs, g, B))));
  MATCHA_CHECK_LAUNCH("score_bwd");
  return MATCHA_OK;
}
int launch_bce(const float* logits, const float* y, const float* w, float alpha, float* dlogit, float* loss_out, int64_t B,
               cudaStream_t s) {
  int blocks = (int)((B + 255) / 256);
  if (blocks < 1) blocks = 1;
  if (blocks > kSMs * 2) blocks = kSMs * 2;
  bce_kernel<<<blocks, 256, 0, s>>>(logits, y, w, alpha, dlogit, loss_out, B);
  MATCHA_CHECK_LAUNCH("bce");
  return MATCHA_OK;
}
int launch_finalize_loss(float* loss_out, const float* recon, float alpha, float beta, cudaStream_t s) {
  finalize_loss_kernel<<<1, 1, 0, s>>>(loss_out, recon, alpha, beta);
  MATCHA_CHECK_LAUNCH("finalize_loss");
  return MATCHA_OK;
}
int launch_recon_diff(float* pred, int64_t ld, const int64_t* x, int64_t T, const float* inter, int64_t inter_ld,
                      int64_t r_start, int64_t r_end, const int32_t* counts, int rchrom, int n_chrom, float* recon_out,
                      cudaStream_t s) {
  if (T <= 0) return MATCHA_OK;
  int64_t blocks = (T * 32 + kBlock - 1) / kBlock;
  if (blocks > kSMs * 8) blocks = kSMs * 8;
  recon_diff_kernel<<<(int)blocks, kBlock, 0, s>>>(pred, ld, x, T, inter, inter_ld, r_start, r_end, counts, rchrom,
                                                   n_chrom, recon_out);
  MATCHA_CHECK_LAUNCH("recon_diff");
  return MATCHA_OK;
}
int launch_ln_tanh_bwd(int d, const float* dxhat, int nparts, int64_t part_stride, const float* dXs, const float* xhat,
                       const float* rstd, const float* X, float* dP, int64_t T, cudaStream_t s) {
  if (T <= 0) return MATCHA_OK;
  MATCHA_DISPATCH_D(d, (ln_tanh_bwd_kernel<DD><<<grid_for_rows(T, d, kSMs * 16), kBlock, 0, s>>>(dxhat, nparts, part_stride, dXs, xhat, rstd, X, dP, T)));
  MATCHA_CHECK_LAUNCH("ln_tanh_bwd");
  return MATCHA_OK;
}
int launch_enc_combine_bwd(const float* dV0, const float* dtE, const float* E, float beta, float* dE, int64_t n,
                           cudaStream_t s) {
  if (n <= 0) return MATCHA_OK;
  const int64_t n4 = n / 4;
  int64_t blocks = (n4 + 255) / 256;
  if (blocks > kSMs * 16) blocks = kSMs * 16;
  enc_combine_bwd_kernel<<<(int)blocks, 256, 0, s>>>(reinterpret_cast<const float4*>(dV0),
                                                     reinterpret_cast<const float4*>(dtE),
                                                     reinterpret_cast<const float4*>(E), beta,
                                                     reinterpret_cast<float4*>(dE), n4);
  MATCHA_CHECK_LAUNCH("enc_combine_bwd");
  return MATCHA_OK;
}
int launch_iota_i64(int64_t* out, int64_t n, cudaStream_t s) {
  if (n <= 0) return MATCHA_OK;
  int64_t blocks = (n + 255) / 256;
  if (blocks > kSMs * 8) blocks = kSMs * 8;
  iota_i64_kernel<<<(int)blocks, 256, 0, s>>>(out, n);
  MATCHA_CHECK_LAUNCH("iota_i64");
  return MATCHA_OK;
}
int launch_pair_u(int d, const float* QKG, const float* b_dyn, float* U, int64_t T, cudaStream_t s) {
  if (T <= 0) return MATCHA_OK;
  MATCHA_DISPATCH_D(d, (pair_u_kernel<DD><<<grid_for_rows(T, d, kSMs * 16), kBlock, 0, s>>>(QKG, b_dyn, U, T)));
  MATCHA_CHECK_LAUNCH("pair_u");
  return MATCHA_OK;
}
int launch_pair_ds(int d, const float* H2, const float* xhat, ScoreParams p, float* D, float* S, int64_t T, cudaStream_t s) {
  if (T <= 0) return MATCHA_OK;
  MATCHA_DISPATCH_D(d, (pair_ds_kernel<DD><<<grid_for_rows(T, d, kSMs * 16), kBlock, 0, s>>>(H2, xhat, p, D, S, T)));
  MATCHA_CHECK_LAUNCH("pair_ds");
  return MATCHA_OK;
}

}  // namespace matcha
