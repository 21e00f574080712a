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
// TMEM loads with the wait decoupled from the issue (several loads in flight per warp)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// wait for all outstanding tcgen05.ld of this thread; the "+r" pass-through pins every later use of the
// loaded registers behind the wait
template <int N>
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[N]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < N; ++i) asm volatile("" : "+r"(r[i]));
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
    if (lane == 0) {
      int ws = 0; uint32_t wp = 0, xp[2] = {0u, 0u};
      for (int64_t pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
        for (int s = 0; s < 2; ++s) {
          const int64_t tile = 2 * pair + s;
          if (tile >= ntiles) break;
          mbar_wait(&x_empty[s], xp[s] ^ 1);
          mbar_expect_tx(&x_full[s], kAXBytes);
          const uint8_t* src = xt + tile * (int64_t)kXTileBytes;
          bulk_g2s(sX + s * kAXBytes, src, 16384, &x_full[s]);
          bulk_g2s(sX + s * kAXBytes + 16384, src + kXHalfBytes, 16384, &x_full[s]);
          xp[s] ^= 1;
        }
        for (int h = 0; h < kH; ++h) {
          mbar_wait(&w_empty[ws], wp ^ 1);
          mbar_expect_tx(&w_full[ws], kAWBytes);
          bulk_g2s(sW + ws * kAWBytes, wheads + (int64_t)h * kAWBytes, kAWBytes, &w_full[ws]);
          if (++ws == 2) { ws = 0; wp ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(128, 192, false, false);
      int ws = 0; uint32_t wp = 0, xp[2] = {0u, 0u}, ap[2] = {0u, 0u};
      for (int64_t pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
        const int nsub = (2 * pair + 1 < ntiles) ? 2 : 1;
        for (int h = 0; h < kH; ++h) {
          mbar_wait(&w_full[ws], wp);
          const uint32_t wh = smem_u32(sW + ws * kAWBytes), wl = wh + kAWHalf;
          for (int s = 0; s < nsub; ++s) {
            if (h == 0) { mbar_wait(&x_full[s], xp[s]); xp[s] ^= 1; }
            mbar_wait(&acc_empty[s], ap[s] ^ 1);
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
          float* pp = probs + (t * kH + h) * PL;
#pragma unroll
          for (int s = 0; s < L - 1; ++s) pp[s] = S[s];
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

}  // namespace

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
