// Fused "hyperedge tile" attention kernels (Modules.py:513-575 with the folds of engine.cu's matcha_prepare):
//   forward : xhat tile -> per head { tcgen05: [Q_h | K_h | G_h] = xhat . W_h^T into TMEM } ->
//             diagonal-masked softmax over the L tokens of each hyperedge and A.G with warp shuffles ->
//             U = dropout(sum_h A_h G_h + b_dyn) * non_pad_mask.            QKG never touches HBM.
//
// Tiles are HYPEREDGE ALIGNED: a tile has 4 warp-rows of 32 TMEM lanes; warp-row q of tile i holds the
// RPW = floor(32 / L) * L consecutive tokens starting at (4 i + q) * RPW, so every hyperedge lives inside one
// warp and the thread that owns TMEM lane r (tcgen05.ld 32x32b) owns token row r.  Lanes >= RPW are zero rows.
//
// Roles (320 threads, 1 CTA / SM, persistent over tile PAIRS so a head's weights are fetched once per 2 tiles):
//   warp 0      producer: cp.async.bulk of xhat tiles and per-head weight chunks (48 KB, bf16 hi | lo)
//   warp 1      one thread issues tcgen05.mma (bf16x3: lo*hi + hi*lo + hi*hi), N = 192 per (tile, head)
//   warps 2-9   two groups of 4 warps, one per tile of the pair: tcgen05.ld + shuffles + softmax
// The two tiles of a pair alternate on the tensor pipe, so the MMA of one overlaps the shuffle phase of the other.
#include "rowwise.cuh"
#include "tc_common.cuh"

namespace matcha {
#ifdef MATCHA_ATTNB_TRACE
__device__ unsigned long long g_abtrace[4096];
// stage index n (first 250 of CTA 0) x 16 event slots
#define ABTRACE(ev) do { if (blockIdx.x == 0 && n < 250) { unsigned long long _t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t)); g_abtrace[n * 16 + (ev)] = _t; } } while (0)
#else
#define ABTRACE(ev) do { } while (0)
#endif

namespace {

constexpr int kAThreads = 320;
constexpr int kAWHalf = 192 * 64 * 2;              // 24576: one head's [Q_h | K_h | G_h] rows as bf16
constexpr int kAWBytes = 2 * kAWHalf;              // hi | lo
constexpr int kAXBytes = 32768;                    // first 8 planes of both halves of an xhat tile
constexpr int kASmem = 2 * kAWBytes + 2 * kAXBytes;
constexpr float kLnEpsA = 1e-5f;

// ------------------------------------------------------------------------------------------
// parameter preparation: W_qkg [1536, 64] fp32 (rows Q 0..511 | K 512..1023 | G 1024..1535) -> per head
// K-major canonical chunks [k/8][192 rows][8] bf16, hi then lo (LBO = 3072 B, SBO = 128 B)
// ------------------------------------------------------------------------------------------
__global__ void split_w_heads_kernel(const float* __restrict__ W, uint8_t* __restrict__ out) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;       // unit = (head, row r, k-group g)
  if (u >= kH * 192 * 8) return;
  const int g = u & 7, r = (u >> 3) % 192, h = u / (192 * 8);
  const int src_row = (r >> 6) * (kH * kD) + h * kD + (r & 63);
  const float* src = W + (int64_t)src_row * kD + g * 8;
  uint4 hi, lo;
  split8(__ldg(reinterpret_cast<const float4*>(src)), __ldg(reinterpret_cast<const float4*>(src + 4)), hi, lo);
  uint8_t* chunk = out + (int64_t)h * kAWBytes;
  const int off = g * (192 * 16) + r * 16;
  *reinterpret_cast<uint4*>(chunk + off) = hi;
  *reinterpret_cast<uint4*>(chunk + kAWHalf + off) = lo;
}

// ------------------------------------------------------------------------------------------
// LayerNorm statistics + hyperedge-aligned tile emission (one half-warp per TILE ROW)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float red16a(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}
__global__ void __launch_bounds__(256) ln_fwd_atiles_kernel(const float* __restrict__ X, float* __restrict__ xhat,
                                                            float* __restrict__ rstd, int64_t T, int rpw,
                                                            uint8_t* __restrict__ xt) {
  const int hl = threadIdx.x & 15;
  const int64_t hw0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4, nhw = ((int64_t)gridDim.x * blockDim.x) >> 4;
  const int64_t rows = ((T + 4 * rpw - 1) / (4 * rpw)) * 128;
  const int64_t iters = (rows + nhw - 1) / nhw;
  for (int64_t it = 0; it < iters; ++it) {
    const int64_t R = hw0 + it * nhw;
    if (R >= rows) continue;                      // whole half-warp leaves together
    const int r = (int)(R & 127), lane_r = r & 31;
    const int64_t t = ((R >> 7) * 4 + (r >> 5)) * rpw + lane_r;
    const bool live = lane_r < rpw && t < T;
    uint2 hi = make_uint2(0u, 0u), lo = make_uint2(0u, 0u);
    if (live) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(X + t * kD + hl * 4));
      const float mean = red16a((v.x + v.y) + (v.z + v.w)) * (1.0f / kD);
      const float4 c = make_float4(v.x - mean, v.y - mean, v.z - mean, v.w - mean);
      const float var = red16a(fmaf(c.x, c.x, fmaf(c.y, c.y, fmaf(c.z, c.z, c.w * c.w)))) * (1.0f / kD);
      const float rs = 1.0f / sqrtf(var + kLnEpsA);
      const float4 xh = make_float4(c.x * rs, c.y * rs, c.z * rs, c.w * rs);
      *reinterpret_cast<float4*>(xhat + t * kD + hl * 4) = xh;
      if (hl == 0) rstd[t] = rs;
      const float x4[4] = {xh.x, xh.y, xh.z, xh.w};
      uint32_t h2[2], l2[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const __nv_bfloat16 h0 = __float2bfloat16_rn(x4[2 * i]), h1 = __float2bfloat16_rn(x4[2 * i + 1]);
        const __nv_bfloat16 l0 = __float2bfloat16_rn(x4[2 * i] - __bfloat162float(h0));
        const __nv_bfloat16 l1 = __float2bfloat16_rn(x4[2 * i + 1] - __bfloat162float(h1));
        h2[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
        l2[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
      }
      hi = make_uint2(h2[0], h2[1]);
      lo = make_uint2(l2[0], l2[1]);
    }
    uint8_t* tile = xt + (R >> 7) * (int64_t)kXTileBytes;
    const int off = (hl >> 1) * kPlaneBytes + r * 16 + (hl & 1) * 8;
    *reinterpret_cast<uint2*>(tile + off) = hi;
    *reinterpret_cast<uint2*>(tile + kXHalfBytes + off) = lo;
    if (hl < 4) {   // planes 8 (ones column) and 9 (zeros) of both halves
      uint4 v = make_uint4((hl == 0 && live) ? 0x00003F80u : 0u, 0u, 0u, 0u);
      uint8_t* dst = tile + (hl >> 1) * kXHalfBytes + (8 + (hl & 1)) * kPlaneBytes + r * 16;
      *reinterpret_cast<uint4*>(dst) = v;
    }
  }
}

// ------------------------------------------------------------------------------------------
// fused attention forward
// ------------------------------------------------------------------------------------------
template <int L>
__global__ void __launch_bounds__(kAThreads, 1)
attn_fused_fwd_kernel(const uint8_t* __restrict__ xt, const uint8_t* __restrict__ wheads, const float* __restrict__ bq,
                      const float* __restrict__ b_dyn, const int64_t* __restrict__ x, float* __restrict__ U,
                      float* __restrict__ probs, int64_t T, const DropCfg drop) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sW = smem;
  uint8_t* sX = smem + 2 * kAWBytes;
  __shared__ uint64_t w_full[2], w_empty[2], x_full[2], x_empty[2], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float sBq[kH * kD];
  __shared__ __align__(16) float sBd[kD];
  constexpr int RPW = (32 / L) * L;
  constexpr int PL = (L - 1 <= 4) ? 4 : 8;          // probabilities stored per (token, head)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = (T + 4 * RPW - 1) / (4 * RPW);
  const int64_t npairs = (ntiles + 1) / 2;

  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 32) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1);
      mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < kH * kD; i += kAThreads) sBq[i] = __ldg(bq + i);
  if (tid < kD) sBd[tid] = __ldg(b_dyn + tid);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    if (elect_one()) {
      int ws = 0; uint32_t wp = 0, xp[2] = {0u, 0u};
      for (int64_t pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
        for (int s = 0; s < 2; ++s) {
          const int64_t tile = 2 * pair + s;
          if (tile >= ntiles) break;
          mbar_wait_backoff(&x_empty[s], xp[s] ^ 1);
          mbar_expect_tx(&x_full[s], kAXBytes);
          const uint8_t* src = xt + tile * (int64_t)kXTileBytes;
          bulk_g2s(sX + s * kAXBytes, src, 16384, &x_full[s]);
          bulk_g2s(sX + s * kAXBytes + 16384, src + kXHalfBytes, 16384, &x_full[s]);
          xp[s] ^= 1;
        }
        for (int h = 0; h < kH; ++h) {
          mbar_wait_backoff(&w_empty[ws], wp ^ 1);
          mbar_expect_tx(&w_full[ws], kAWBytes);
          bulk_g2s(sW + ws * kAWBytes, wheads + (int64_t)h * kAWBytes, kAWBytes, &w_full[ws]);
          if (++ws == 2) { ws = 0; wp ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(128, 192, false, false);
      int ws = 0; uint32_t wp = 0, xp[2] = {0u, 0u}, ap[2] = {0u, 0u};
      for (int64_t pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
        const int nsub = (2 * pair + 1 < ntiles) ? 2 : 1;
        for (int h = 0; h < kH; ++h) {
          mbar_wait_backoff(&w_full[ws], wp);
          const uint32_t wh = smem_u32(sW + ws * kAWBytes), wl = wh + kAWHalf;
          for (int s = 0; s < nsub; ++s) {
            if (h == 0) { mbar_wait_backoff(&x_full[s], xp[s]); xp[s] ^= 1; }
            mbar_wait_backoff(&acc_empty[s], ap[s] ^ 1);
            ap[s] ^= 1;
            tc_fence_after();
            const uint32_t xh = smem_u32(sX + s * kAXBytes), xl = xh + 16384;
            const uint32_t d = tmem_base + s * 256;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)    // A = xhat tile (128 tokens, K-major), B = W_h (192 rows, K-major)
              umma_x3s(d, xh + ks * 4096, xl + ks * 4096, wh + ks * 6144, wl + ks * 6144, 2048, 128, 3072, 128, idesc, ks == 0);
            umma_commit(&acc_full[s]);
            if (h == kH - 1) umma_commit(&x_empty[s]);
          }
          umma_commit(&w_empty[ws]);
          if (++ws == 2) { ws = 0; wp ^= 1; }
        }
      }
    }
  } else {
    const int sub = (warp - 2) >> 2, q = warp & 3;
    const bool live_lane = lane < RPW;
    const int g = lane / L, pos = lane - g * L;
    int src[L - 1];
#pragma unroll
    for (int s = 1; s < L; ++s) src[s - 1] = live_lane ? g * L + (pos + s) % L : lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + sub * 256;
    uint32_t ap = 0;
    for (int64_t pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
      const int64_t tile = 2 * pair + sub;
      if (tile >= ntiles) continue;
      const int64_t t = (tile * 4 + q) * RPW + lane;
      const bool live = live_lane && t < T;
      float acc[kD];
#pragma unroll
      for (int c = 0; c < kD; ++c) acc[c] = 0.f;
#pragma unroll 1
      for (int h = 0; h < kH; ++h) {
        mbar_wait(&acc_full[sub], ap);
        ap ^= 1;
        tc_fence_after();
        float S[L - 1];
#pragma unroll
        for (int s = 0; s < L - 1; ++s) S[s] = 0.f;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t qv[32], kv[32];
          tmem_ld32_issue(taddr + half * 32, qv);
          tmem_ld32_issue(taddr + 64 + half * 32, kv);
          tmem_ld_wait(qv);
          tmem_ld_wait(kv);
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const float qc = __uint_as_float(qv[c]) + sBq[h * kD + half * 32 + c];
#pragma unroll
            for (int s = 0; s < L - 1; ++s)
              S[s] = fmaf(qc, __uint_as_float(__shfl_sync(0xffffffffu, kv[c], src[s])), S[s]);
          }
        }
        // softmax over the L-1 other tokens of the hyperedge (diagonal masked, pads are live keys)
        float mx = S[0];
#pragma unroll
        for (int s = 1; s < L - 1; ++s) mx = fmaxf(mx, S[s]);
        float sum = 0.f;
#pragma unroll
        for (int s = 0; s < L - 1; ++s) { S[s] = expf(S[s] - mx); sum += S[s]; }
        const float inv = 1.0f / sum;
#pragma unroll
        for (int s = 0; s < L - 1; ++s) S[s] *= inv;
        if (probs && live) {
          float pr[PL];
#pragma unroll
          for (int s = 0; s < PL; ++s) pr[s] = (s < L - 1) ? S[s < L - 1 ? s : 0] : 0.f;
          float* pp = probs + (t * kH + h) * PL;
#pragma unroll
          for (int s = 0; s < PL; s += 4) *reinterpret_cast<float4*>(pp + s) = make_float4(pr[s], pr[s + 1], pr[s + 2], pr[s + 3]);
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t gv[32];
          tmem_ld32_issue(taddr + 128 + half * 32, gv);
          tmem_ld_wait(gv);
#pragma unroll
          for (int c = 0; c < 32; ++c) {
#pragma unroll
            for (int s = 0; s < L - 1; ++s)
              acc[half * 32 + c] = fmaf(S[s], __uint_as_float(__shfl_sync(0xffffffffu, gv[c], src[s])), acc[half * 32 + c]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[sub]);
      }
      if (live) {
        const float m = x[t] != 0 ? 1.f : 0.f;                      // non_pad_mask (Modules.py:614)
        float* dst = U + t * kD;
#pragma unroll
        for (int c = 0; c < kD; c += 4) {
          float4 v = make_float4(acc[c] + sBd[c], acc[c + 1] + sBd[c + 1], acc[c + 2] + sBd[c + 2], acc[c + 3] + sBd[c + 3]);
          v = drop_apply4(drop, (uint64_t)t, (uint32_t)c, v);       // dropout after fc1 (Modules.py:572)
          *reinterpret_cast<float4*>(dst + c) = make_float4(v.x * m, v.y * m, v.z * m, v.w * m);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ==========================================================================================
// fused attention backward
//   grid = 4 head pairs x S token splits (one CTA per SM).  A CTA keeps its head pair's weights resident in shared
//   memory (one copy serves the recompute, K-major, and the data gradient, MN-major) and its 384 x 64 slice of the
//   weight gradient resident in TMEM across all of its tiles.  Per tile, three stages (G, K, Q):
//     tcgen05  R = xhat . W_piece^T                  (recompute, N = 128: both heads of the pair)
//     SIMT     thread = (token row, head): shuffles within the hyperedge -> dG / dQ / dK rows -> bf16 hi|lo d-tile
//     tcgen05  dxhat += d-tile . W_piece   and   dW_piece += d-tile^T . xhat
//   TMEM columns: dxhat 0..63 | R0 64..191 | R1 192..319 | dW_G, dW_K, dW_Q 320..511.
// ==========================================================================================
constexpr int kBThreads = 320;
constexpr int kBWPiece = 32768;                 // piece pair: 128 rows x 64 k, hi 16 KB | lo 16 KB
constexpr int kBWBytes = 3 * kBWPiece;          // G | K | Q
constexpr int kBXBytes = 32768;
constexpr int kBDHalf = 32768;                  // d-tile pair: 16 planes x 2048
constexpr int kBDBytes = 2 * kBDHalf;
constexpr int kBSmem = kBWBytes + 2 * kBXBytes + kBDBytes;   // 229376
constexpr uint32_t kColDX = 0, kColR = 64, kColDW = 320;

// W_qkg [1536, 64] -> per head pair hp, piece p (0 = G, 1 = K, 2 = Q): K-major [k/8][128 rows][8] hi | lo,
// rows 0..63 = head 2 hp, rows 64..127 = head 2 hp + 1
__global__ void split_w_pairs_kernel(const float* __restrict__ W, uint8_t* __restrict__ out) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;       // unit = (hp, piece, row r, k-group g)
  if (u >= 4 * 3 * 128 * 8) return;
  const int g = u & 7, r = (u >> 3) & 127, p = (u >> 10) % 3, hp = u / (3 * 1024);
  const int head = 2 * hp + (r >> 6);
  const int base = (p == 0) ? 2 * kH * kD : (p == 1 ? kH * kD : 0);
  const float* src = W + (int64_t)(base + head * kD + (r & 63)) * kD + g * 8;
  uint4 hi, lo;
  split8(__ldg(reinterpret_cast<const float4*>(src)), __ldg(reinterpret_cast<const float4*>(src + 4)), hi, lo);
  uint8_t* chunk = out + (int64_t)hp * kBWBytes + p * kBWPiece;
  const int off = g * 2048 + r * 16;
  *reinterpret_cast<uint4*>(chunk + off) = hi;
  *reinterpret_cast<uint4*>(chunk + 16384 + off) = lo;
}

// column sums of a [32 lanes x 64] register tile: 62 shuffles; lane l ends with the sums of columns col0, col0 + 1
__device__ __forceinline__ void warp_colsum64(const float (&v)[64], int lane, float& s0, float& s1, int& col0) {
  float a[32], b[16], c[8], d[4];
  bool up = (lane & 16) != 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const float send = up ? v[i] : v[32 + i], keep = up ? v[32 + i] : v[i];
    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  up = (lane & 8) != 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float send = up ? a[i] : a[16 + i], keep = up ? a[16 + i] : a[i];
    b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  up = (lane & 4) != 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float send = up ? b[i] : b[8 + i], keep = up ? b[8 + i] : b[i];
    c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  up = (lane & 2) != 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = up ? c[i] : c[4 + i], keep = up ? c[4 + i] : c[i];
    d[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  up = (lane & 1) != 0;
  {
    const float send0 = up ? d[0] : d[2], keep0 = up ? d[2] : d[0];
    const float send1 = up ? d[1] : d[3], keep1 = up ? d[3] : d[1];
    s0 = keep0 + __shfl_xor_sync(0xffffffffu, send0, 1);
    s1 = keep1 + __shfl_xor_sync(0xffffffffu, send1, 1);
  }
  col0 = ((lane & 16) ? 32 : 0) + ((lane & 8) ? 16 : 0) + ((lane & 4) ? 8 : 0) + ((lane & 2) ? 4 : 0) + ((lane & 1) ? 2 : 0);
}

// one 64-wide fp32 row -> bf16 hi | lo planes of the d-tile pair (thread = tile row r, planes p0 .. p0 + 7)
__device__ __forceinline__ void store_drow(uint8_t* sD, int p0, int r, const float (&o)[64]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    uint4 hi, lo;
    split8(make_float4(o[8 * j], o[8 * j + 1], o[8 * j + 2], o[8 * j + 3]),
           make_float4(o[8 * j + 4], o[8 * j + 5], o[8 * j + 6], o[8 * j + 7]), hi, lo);
    sts16(sD + (p0 + j) * 2048 + r * 16, hi);
    sts16(sD + kBDHalf + (p0 + j) * 2048 + r * 16, lo);
  }
}

// umma_x3s with a run-time pass count: 3 = fp32-accurate bf16 split, 1 = hi * hi only (bf16 mode)
__device__ __forceinline__ void umma_xps(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, uint32_t a_lbo,
                                         uint32_t a_sbo, uint32_t b_lbo, uint32_t b_sbo, uint32_t idesc, bool first, int passes) {
  if (passes == 3) {
    umma_x3s(tmem_d, a_hi, a_lo, b_hi, b_lo, a_lbo, a_sbo, b_lbo, b_sbo, idesc, first);
  } else {
    umma_bf16(tmem_d, make_smem_desc(a_hi, a_lbo, a_sbo), make_smem_desc(b_hi, b_lbo, b_sbo), idesc, first ? 0u : 1u);
  }
}

template <int L>
__global__ void __launch_bounds__(kBThreads, 1)
attn_fused_bwd_kernel(const uint8_t* __restrict__ xt, const uint8_t* __restrict__ wpairs, const float* __restrict__ bq,
                      const float* __restrict__ dd_in, const float* __restrict__ probs, float* __restrict__ dxhat_parts,
                      float* __restrict__ dW, float* __restrict__ dbq, float* __restrict__ db_dyn, int64_t T, int passes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sW = smem;
  uint8_t* sX = smem + kBWBytes;
  uint8_t* sD = smem + kBWBytes + 2 * kBXBytes;
  __shared__ uint64_t w_full, x_full[2], x_empty[2], r_full[2], r_empty[2], d_full, d_empty, dx_full, dx_empty, done;
  __shared__ uint32_t tmem_base_s;
  __shared__ float sBq[2 * kD], sDbq[2 * kD], sDbd[kD];
  constexpr int RPW = (32 / L) * L;
  constexpr int PL = (L - 1 <= 4) ? 4 : 8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int hp = blockIdx.x & 3, sp = blockIdx.x >> 2, S = gridDim.x >> 2;
  const int64_t ntiles = (T + 4 * RPW - 1) / (4 * RPW);
  const int64_t my_tiles = (ntiles - sp + S - 1) / S;          // tiles sp, sp + S, ...

  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 32) {
    mbar_init(&w_full, 1); mbar_init(&d_full, 8); mbar_init(&d_empty, 1); mbar_init(&dx_full, 1); mbar_init(&dx_empty, 8);
    mbar_init(&done, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); mbar_init(&r_full[i], 1); mbar_init(&r_empty[i], 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 2 * kD) { sBq[tid] = __ldg(bq + (2 * hp) * kD + tid); sDbq[tid] = 0.f; }
  if (tid < kD) sDbd[tid] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(&w_full, kBWBytes);
      for (int p = 0; p < 3; ++p)
        bulk_g2s(sW + p * kBWPiece, wpairs + (int64_t)hp * kBWBytes + p * kBWPiece, kBWPiece, &w_full);
      for (int64_t k = 0; k < my_tiles; ++k) {
        const int b = (int)(k & 1);
        mbar_wait_backoff(&x_empty[b], (uint32_t)((k >> 1) & 1) ^ 1u);
        mbar_expect_tx(&x_full[b], kBXBytes);
        const uint8_t* src = xt + (sp + k * S) * (int64_t)kXTileBytes;
        bulk_g2s(sX + b * kBXBytes, src, 16384, &x_full[b]);
        bulk_g2s(sX + b * kBXBytes + 16384, src + kXHalfBytes, 16384, &x_full[b]);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idescR = make_idesc(128, 128, false, false);
      constexpr uint32_t idescD = make_idesc(128, 64, false, true);
      constexpr uint32_t idescW = make_idesc(128, 64, true, true);
      const int64_t N = 3 * my_tiles;
      const uint32_t dh = smem_u32(sD), dl = dh + kBDHalf;
      mbar_wait_backoff(&w_full, 0);
      auto recompute = [&](int64_t n) {
        const int64_t k = n / 3;
        const int g = (int)(n - 3 * k), xb = (int)(k & 1), rb = (int)(n & 1);
        if (g == 0) mbar_wait_backoff(&x_full[xb], (uint32_t)((k >> 1) & 1));
        mbar_wait_backoff(&r_empty[rb], (uint32_t)((n >> 1) & 1) ^ 1u);
        tc_fence_after();
        const uint32_t xh = smem_u32(sX + xb * kBXBytes), xl = xh + 16384;
        const uint32_t wh = smem_u32(sW + g * kBWPiece), wl = wh + 16384;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_xps(tmem_base + kColR + rb * 128, xh + ks * 4096, xl + ks * 4096, wh + ks * 4096, wl + ks * 4096, 2048, 128,
                   2048, 128, idescR, ks == 0, passes);
        umma_commit(&r_full[rb]);
      };
      if (N > 0) recompute(0);
      for (int64_t n = 0; n < N; ++n) {
        if (n + 1 < N) recompute(n + 1);
        const int64_t k = n / 3;
        const int g = (int)(n - 3 * k), xb = (int)(k & 1);
        ABTRACE(8);
        mbar_wait_backoff(&d_full, (uint32_t)(n & 1));
        if (g == 0) mbar_wait_backoff(&dx_empty, (uint32_t)(k & 1) ^ 1u);
        ABTRACE(9);
        tc_fence_after();
        // stage g recomputes piece g (G, K, Q) but its d-tile holds the gradient of piece gp (dG, dQ, dK)
        const int gp = (g == 0) ? 0 : (g == 1 ? 2 : 1);
        const uint32_t xh = smem_u32(sX + xb * kBXBytes), xl = xh + 16384;
        const uint32_t wh = smem_u32(sW + gp * kBWPiece), wl = wh + 16384;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)   // dxhat[128 tok, 64] += d[128 tok, 128 f] . W_piece[128 f, 64]   (B MN-major)
          umma_xps(tmem_base + kColDX, dh + ks * 4096, dl + ks * 4096, wh + ks * 256, wl + ks * 256, 2048, 128, 128, 2048,
                   idescD, g == 0 && ks == 0, passes);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)   // dW_piece[128 f, 64] += d^T[128 f, 128 tok] . xhat[128 tok, 64]  (both MN-major)
          umma_xps(tmem_base + kColDW + gp * 64, dh + ks * 256, dl + ks * 256, xh + ks * 256, xl + ks * 256, 128, 2048, 128,
                   2048, idescW, k == 0 && ks == 0, passes);
        umma_commit(&d_empty);
        if (g == 2) { umma_commit(&dx_full); umma_commit(&x_empty[xb]); }
        ABTRACE(10);
      }
      umma_commit(&done);
    }
  } else {
    const int hl = (warp - 2) >> 2, q = warp & 3, head = 2 * hp + hl;
    const int r = q * 32 + lane;
    const bool live_lane = lane < RPW;
    const int gI = lane / L, pos = lane - gI * L;
    int src[L - 1];
#pragma unroll
    for (int s = 1; s < L; ++s) src[s - 1] = live_lane ? gI * L + (pos + s) % L : lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    int64_t n = 0;
    int64_t t_prev = 0;
    bool live_prev = false;
    // dxhat of the previous tile is drained one stage late, so the tensor pipe finishes that tile's last data-gradient
    // MMAs underneath this tile's first shuffle phase
    auto drain_dxhat = [&](int64_t kk, int64_t tt, bool lv) {
      mbar_wait(&dx_full, (uint32_t)(kk & 1));
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32_issue(tlane + kColDX + hl * 32, v);
      tmem_ld_wait(v);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&dx_empty);
      if (lv) {
        float* dst = dxhat_parts + ((int64_t)hp * T + tt) * kD + hl * 32;
#pragma unroll
        for (int c = 0; c < 32; c += 4)
          *reinterpret_cast<float4*>(dst + c) = make_float4(__uint_as_float(v[c]), __uint_as_float(v[c + 1]),
                                                            __uint_as_float(v[c + 2]), __uint_as_float(v[c + 3]));
      }
    };
    // rows of the next tile are fetched as soon as the current tile's registers are free, so their latency hides
    // under the K and Q stages
    float dd[kD];                                    // gradient wrt the attention output, already masked / dropout-scaled
    float pr[PL];
    auto fetch_rows = [&](int64_t kk) {
      const int64_t tt = ((sp + kk * S) * 4 + q) * RPW + lane;
      const bool lv = live_lane && tt < T;
      const int64_t tl = lv ? tt : 0;               // clamped row: loads are issued unconditionally, results masked on use
#pragma unroll
      for (int c = 0; c < kD; c += 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(dd_in + tl * kD + c));
        dd[c] = v.x; dd[c + 1] = v.y; dd[c + 2] = v.z; dd[c + 3] = v.w;
      }
      const float* pp = probs + (tl * kH + head) * PL;
#pragma unroll
      for (int s = 0; s < PL; s += 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(pp + s));
        pr[s] = v.x; pr[s + 1] = v.y; pr[s + 2] = v.z; pr[s + 3] = v.w;
      }
    };
    if (my_tiles > 0) fetch_rows(0);
    for (int64_t k = 0; k < my_tiles; ++k) {
      const int64_t tile = sp + k * S;
      const int64_t t = (tile * 4 + q) * RPW + lane;
      const bool live = live_lane && t < T;
      float A[L - 1], AT[L - 1], dS[L - 1];
      float o[kD];
      // ---------------- stage G: dA = dd . G_j,  dG_i = sum_j A_ji dd_j ----------------
      {
        if (!live) {
#pragma unroll
          for (int c = 0; c < kD; ++c) dd[c] = 0.f;
        }
#pragma unroll
        for (int s = 0; s < L - 1; ++s) A[s] = live ? pr[s] : 0.f;
#pragma unroll
        for (int s = 1; s < L; ++s) AT[s - 1] = __shfl_sync(0xffffffffu, A[L - s - 1], src[s - 1]);   // weight of row (i+s) on row i
        if (hp == 0 && hl == 0) {          // db_dyn = sum over tokens of the masked, dropout-scaled gradient
          // (spreading these 64 column sums over all head pairs / heads, 8 columns per warp, was measured slower: 295 -> 320 us
          // -- every warp then adds 40 shuffles per tile to the shared-memory pipe that already bounds the kernel)
          float s0, s1; int col0;
          warp_colsum64(dd, lane, s0, s1, col0);
          atomicAdd(&sDbd[col0], s0);
          atomicAdd(&sDbd[col0 + 1], s1);
        }
        float dA[L - 1];
#pragma unroll
        for (int s = 0; s < L - 1; ++s) dA[s] = 0.f;
        const int rb = (int)(n & 1);
        if (tid == 64) ABTRACE(0);
        mbar_wait(&r_full[rb], (uint32_t)((n >> 1) & 1));
        if (tid == 64) ABTRACE(1);
        tc_fence_after();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t gv[32];
          tmem_ld32_issue(tlane + kColR + rb * 128 + hl * 64 + half * 32, gv);
          tmem_ld_wait(gv);
#pragma unroll
          for (int c = 0; c < 32; ++c)
#pragma unroll
            for (int s = 0; s < L - 1; ++s)
              dA[s] = fmaf(dd[half * 32 + c], __uint_as_float(__shfl_sync(0xffffffffu, gv[c], src[s])), dA[s]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&r_empty[rb]);
#pragma unroll
        for (int c = 0; c < kD; ++c) {
          float acc = 0.f;
#pragma unroll
          for (int s = 0; s < L - 1; ++s) acc = fmaf(AT[s], __shfl_sync(0xffffffffu, dd[c], src[s]), acc);
          o[c] = acc;
        }
        float dot = 0.f;
#pragma unroll
        for (int s = 0; s < L - 1; ++s) dot = fmaf(A[s], dA[s], dot);
#pragma unroll
        for (int s = 0; s < L - 1; ++s) dS[s] = A[s] * (dA[s] - dot);
        if (tid == 64) ABTRACE(2);
        if (k + 1 < my_tiles) fetch_rows(k + 1);                // (fetching at the top of the next tile instead: 294 -> 309 us;
                                                                //  second half of the row only at the end of stage Q: -> 319 us)
        if (k > 0) drain_dxhat(k - 1, t_prev, live_prev);      // (draining first: 320 -> 330 us)
        if (tid == 64) ABTRACE(3);
        mbar_wait(&d_empty, (uint32_t)(n & 1) ^ 1u);
        if (tid == 64) ABTRACE(4);
        store_drow(sD, hl * 8, r, o);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&d_full);
        if (tid == 64) ABTRACE(5);
        ++n;
      }
      // ---------------- stage K: dQ_i = sum_j dS_ij K_j ----------------
      {
        const int rb = (int)(n & 1);
        if (tid == 64) ABTRACE(0);
        mbar_wait(&r_full[rb], (uint32_t)((n >> 1) & 1));
        if (tid == 64) ABTRACE(1);
        tc_fence_after();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t kv[32];
          tmem_ld32_issue(tlane + kColR + rb * 128 + hl * 64 + half * 32, kv);
          tmem_ld_wait(kv);
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            float acc = 0.f;
#pragma unroll
            for (int s = 0; s < L - 1; ++s) acc = fmaf(dS[s], __uint_as_float(__shfl_sync(0xffffffffu, kv[c], src[s])), acc);
            o[half * 32 + c] = acc;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&r_empty[rb]);
        {                                   // gradient of the folded Q bias
          float s0, s1; int col0;
          warp_colsum64(o, lane, s0, s1, col0);
          atomicAdd(&sDbq[hl * kD + col0], s0);
          atomicAdd(&sDbq[hl * kD + col0 + 1], s1);
        }
        if (tid == 64) ABTRACE(3);
        mbar_wait(&d_empty, (uint32_t)(n & 1) ^ 1u);
        if (tid == 64) ABTRACE(4);
        store_drow(sD, hl * 8, r, o);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&d_full);
        if (tid == 64) ABTRACE(5);
        ++n;
      }
      // ---------------- stage Q: dK_i = sum_j dS_ji Q_j ----------------
      {
        float dST[L - 1];
#pragma unroll
        for (int s = 1; s < L; ++s) dST[s - 1] = __shfl_sync(0xffffffffu, dS[L - s - 1], src[s - 1]);
        const int rb = (int)(n & 1);
        if (tid == 64) ABTRACE(0);
        mbar_wait(&r_full[rb], (uint32_t)((n >> 1) & 1));
        if (tid == 64) ABTRACE(1);
        tc_fence_after();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t qv[32];
          tmem_ld32_issue(tlane + kColR + rb * 128 + hl * 64 + half * 32, qv);
          tmem_ld_wait(qv);
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const float qc = live_lane ? __uint_as_float(qv[c]) + sBq[hl * kD + half * 32 + c] : 0.f;
            float acc = 0.f;
#pragma unroll
            for (int s = 0; s < L - 1; ++s) acc = fmaf(dST[s], __shfl_sync(0xffffffffu, qc, src[s]), acc);
            o[half * 32 + c] = acc;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&r_empty[rb]);
        if (tid == 64) ABTRACE(3);
        mbar_wait(&d_empty, (uint32_t)(n & 1) ^ 1u);
        if (tid == 64) ABTRACE(4);
        store_drow(sD, hl * 8, r, o);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&d_full);
        if (tid == 64) ABTRACE(5);
        ++n;
      }
      t_prev = t;
      live_prev = live;
    }
    if (my_tiles > 0) drain_dxhat(my_tiles - 1, t_prev, live_prev);
    // ---------------- weight-gradient slice of this CTA: reduced over the token splits with 128-bit atomics ----------------
    // (the S = 37 CTAs of a head pair finish within microseconds of each other; their 98 KB slices land in L2's atomic
    // units instead of a partial buffer + a reduction launch.  dW was zeroed with the rest of the derived gradients.)
    mbar_wait(&done, 0);
    tc_fence_after();
    if (my_tiles > 0) {
      const int f = q * 32 + lane;                 // TMEM lane = feature row of the piece pair
      const int hrow = (2 * hp + (f >> 6)) * kD + (f & 63);
#pragma unroll 1
      for (int g = 0; g < 3; ++g) {
        uint32_t v[32];
        tmem_ld32_issue(tlane + kColDW + g * 64 + hl * 32, v);
        tmem_ld_wait(v);
        const int row = ((g == 0) ? 2 * kH * kD : (g == 1 ? kH * kD : 0)) + hrow;
        float* dst = dW + (int64_t)row * kD + hl * 32;
#pragma unroll
        for (int c = 0; c < 32; c += 4)
          atomicAdd(reinterpret_cast<float4*>(dst + c), make_float4(__uint_as_float(v[c]), __uint_as_float(v[c + 1]),
                                                                    __uint_as_float(v[c + 2]), __uint_as_float(v[c + 3])));
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const int i = tid - 64;
    if (i < 2 * kD) atomicAdd(dbq + (2 * hp) * kD + i, sDbq[i]);
    if (hp == 0 && i >= 2 * kD && i < 3 * kD) atomicAdd(db_dyn + (i - 2 * kD), sDbd[i - 2 * kD]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

template <typename K>
int set_smem_attr_a(K kernel, int bytes) {
  return check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes), "cudaFuncSetAttribute");
}

template <int L>
int launch_fwd_L(const uint8_t* xt, const uint8_t* wheads, const float* bq, const float* b_dyn, const int64_t* x, float* U,
                 float* probs, int64_t T, DropCfg drop, cudaStream_t s) {
  static bool once = false;
  if (!once) { if (int rc = set_smem_attr_a(attn_fused_fwd_kernel<L>, kASmem)) return rc; once = true; }
  const int64_t npairs = (num_atiles(T, L) + 1) / 2;
  const unsigned grid = (unsigned)(npairs < kSMs ? npairs : kSMs);
  attn_fused_fwd_kernel<L><<<grid, kAThreads, kASmem, s>>>(xt, wheads, bq, b_dyn, x, U, probs, T, drop);
  MATCHA_CHECK_LAUNCH("attn_fused_fwd");
  return MATCHA_OK;
}

template <int L>
int launch_bwd_L(const uint8_t* xt, const uint8_t* wpairs, const float* bq, const float* dd,
                 const float* probs, float* dxhat_parts, float* part, float* dW, float* dbq, float* db_dyn, int64_t T,
                 int passes, cudaStream_t s) {
  static bool once = false;
  if (!once) { if (int rc = set_smem_attr_a(attn_fused_bwd_kernel<L>, kBSmem)) return rc; once = true; }
  const int64_t ntiles = num_atiles(T, L);
  const int S = (int)(ntiles < kSMs / 4 ? ntiles : kSMs / 4);
  (void)part;       // split-K partial buffer of earlier builds: the slices are reduced with atomics now
  attn_fused_bwd_kernel<L><<<4 * S, kBThreads, kBSmem, s>>>(xt, wpairs, bq, dd, probs, dxhat_parts, dW, dbq, db_dyn, T, passes);
  MATCHA_CHECK_LAUNCH("attn_fused_bwd");
  return MATCHA_OK;
}

}  // namespace

#ifdef MATCHA_ATTNB_TRACE
extern "C" int matcha_attnb_trace(unsigned long long* out) {
  return cudaMemcpyFromSymbol(out, g_abtrace, sizeof(unsigned long long) * 4096) == cudaSuccess ? 0 : -2;
}
#endif

int launch_split_w_pairs(const float* W, void* out, cudaStream_t s) {
  split_w_pairs_kernel<<<(4 * 3 * 128 * 8 + 255) / 256, 256, 0, s>>>(W, reinterpret_cast<uint8_t*>(out));
  MATCHA_CHECK_LAUNCH("split_w_pairs");
  return MATCHA_OK;
}

int64_t attn_fused_bwd_scratch_floats() { return (int64_t)(kSMs / 4) * kQKG * kD; }

// dU <- dU * dropout factor * non_pad_mask, in place (the gradient wrt the pre-dropout attention output)
__global__ void mask_drop_kernel(float* __restrict__ dU, const int64_t* __restrict__ x, int64_t T, const DropCfg drop) {
  const int64_t n4 = T * (kD / 4);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = i >> 4;
    const int c = (int)(i & 15) * 4;
    float4 v = reinterpret_cast<float4*>(dU)[i];
    const float m = x[t] != 0 ? 1.f : 0.f;
    const float4 f = drop_factor4(drop, (uint64_t)t, (uint32_t)c);
    reinterpret_cast<float4*>(dU)[i] = make_float4(v.x * f.x * m, v.y * f.y * m, v.z * f.z * m, v.w * f.w * m);
  }
}

int launch_attn_fused_bwd(const uint8_t* xhat_tiles, const uint8_t* wpairs, const float* bq, const int64_t* x, float* dU,
                          const float* probs, float* dxhat_parts, float* part, float* dW, float* dbq, float* db_dyn,
                          int64_t B, int L, DropCfg drop, int premasked, int passes, cudaStream_t s) {
  if (B <= 0) return MATCHA_OK;
  const int64_t T = B * L;
  if (!premasked) {
    int64_t blocks = (T * 16 + 255) / 256;
    if (blocks > kSMs * 16) blocks = kSMs * 16;
    mask_drop_kernel<<<(unsigned)blocks, 256, 0, s>>>(dU, x, T, drop);
    MATCHA_CHECK_LAUNCH("mask_drop");
  }
  switch (L) {
    case 2: return launch_bwd_L<2>(xhat_tiles, wpairs, bq, dU, probs, dxhat_parts, part, dW, dbq, db_dyn, T, passes, s);
    case 3: return launch_bwd_L<3>(xhat_tiles, wpairs, bq, dU, probs, dxhat_parts, part, dW, dbq, db_dyn, T, passes, s);
    case 4: return launch_bwd_L<4>(xhat_tiles, wpairs, bq, dU, probs, dxhat_parts, part, dW, dbq, db_dyn, T, passes, s);
    case 5: return launch_bwd_L<5>(xhat_tiles, wpairs, bq, dU, probs, dxhat_parts, part, dW, dbq, db_dyn, T, passes, s);
    case 6: return launch_bwd_L<6>(xhat_tiles, wpairs, bq, dU, probs, dxhat_parts, part, dW, dbq, db_dyn, T, passes, s);
    default: set_error("attn_fused_bwd: padded width L=%d unsupported (2..6)", L); return MATCHA_ERR_ARG;
  }
}

int launch_split_w_heads(const float* W, void* out, cudaStream_t s) {
  split_w_heads_kernel<<<(kH * 192 * 8 + 255) / 256, 256, 0, s>>>(W, reinterpret_cast<uint8_t*>(out));
  MATCHA_CHECK_LAUNCH("split_w_heads");
  return MATCHA_OK;
}

int launch_ln_fwd_atiles(const float* X, float* xhat, float* rstd, int64_t T, int L, uint8_t* xhat_tiles, cudaStream_t s) {
  if (T <= 0) return MATCHA_OK;
  const int64_t rows = num_atiles(T, L) * 128;
  int64_t blocks = (rows * 16 + 255) / 256;
  if (blocks > kSMs * 16) blocks = kSMs * 16;
  ln_fwd_atiles_kernel<<<(unsigned)blocks, 256, 0, s>>>(X, xhat, rstd, T, rows_per_warp(L), xhat_tiles);
  MATCHA_CHECK_LAUNCH("ln_fwd_atiles");
  return MATCHA_OK;
}

int launch_attn_fused_fwd(const uint8_t* xhat_tiles, const uint8_t* wheads, const float* bq, const float* b_dyn,
                          const int64_t* x, float* U, float* probs, int64_t B, int L, DropCfg drop, cudaStream_t s) {
  if (B <= 0) return MATCHA_OK;
  const int64_t T = B * L;
  switch (L) {
    case 2: return launch_fwd_L<2>(xhat_tiles, wheads, bq, b_dyn, x, U, probs, T, drop, s);
    case 3: return launch_fwd_L<3>(xhat_tiles, wheads, bq, b_dyn, x, U, probs, T, drop, s);
    case 4: return launch_fwd_L<4>(xhat_tiles, wheads, bq, b_dyn, x, U, probs, T, drop, s);
    case 5: return launch_fwd_L<5>(xhat_tiles, wheads, bq, b_dyn, x, U, probs, T, drop, s);
    case 6: return launch_fwd_L<6>(xhat_tiles, wheads, bq, b_dyn, x, U, probs, T, drop, s);
    default: set_error("attn_fused_fwd: padded width L=%d unsupported (2..6)", L); return MATCHA_ERR_ARG;
  }
}

}  // namespace matcha
