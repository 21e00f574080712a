// "X-form" fused attention forward (Modules.py:513-575 with the folds of matcha_prepare), hyperedge widths L <= 5.
//
// The attention of Hyper-SAGNN mixes only the L <= 5 tokens of one hyperedge, so the L x L products are far too small for
// the tensor cores; in the Q/K/G form (attn_fused.cu) every token pulls the K_h and G_h rows of its L - 1 neighbours through
// warp shuffles for every head: 2 (L - 1) 64 = 512 shuffles per token and head, and the shared-memory data pipe that
// executes them is the binding unit (profiles/r01b_ncu_hot.md: 57 % busy, tensor pipe 24 %).  Both products are re-associated
// here so that the only per-neighbour data is the neighbour's INPUT row xhat_j, which does not depend on the head:
//     S_h[i, j] = Q_h[i] . K_h[j] = (N_h xhat_i + v_h) . xhat_j          N_h = Wk_h^T Wq_h,  v_h = Wk_h^T bq_h
//     O_h[i]    = sum_j A_h[i, j] G_h[j] = Wg_h (sum_j A_h[i, j] xhat_j) = Wg_h Z_h[i]
// A thread keeps its neighbours' xhat rows in REGISTERS for all 8 heads (two threads per token: column halves), so per
// head it exchanges 4 partial scores with its partner through shared memory instead of 512 values through shuffles.
//   tcgen05  Y_h = xhat . N_h^T                        (N = 64, per head, TMEM double buffered)
//   SIMT     S, diagonal-free softmax over the L - 1 neighbours, Z_h = sum_j A_ij xhat_j  -> bf16 hi | lo tile in smem
//   tcgen05  U += Z_h . Wg_h^T                         (accumulated over the 8 heads in TMEM)
// then U + b_dyn -> dropout -> non-pad mask.  Exact re-association in real arithmetic; same bf16x3 contraction accuracy.
//
// Roles (320 threads, 1 CTA / SM, persistent over hyperedge-aligned tiles of 128 rows, rowwise.cuh):
//   warp 0 producer (bulk copies of xhat tiles and per-head weight blocks), warp 1 MMA issuer,
//   warps 2..: compute, (TMEM lane quarter q = warp & 3, column part = (warp - 2) >> 2): kXParts threads per token.
#include "rowwise.cuh"
#include "tc_common.cuh"

namespace matcha {
#ifdef MATCHA_XFORM_TRACE
__device__ unsigned long long g_xtrace[4096];
#define XTRACE(slot) do { if (blockIdx.x == 0 && k == 1) { unsigned long long _t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t)); g_xtrace[slot] = _t; } } while (0)
#else
#define XTRACE(slot) do { } while (0)
#endif
namespace {

constexpr int kXParts = 2;                   // threads per token: column parts of 64 / kXParts
constexpr int kXCols = kD / kXParts;
constexpr int kXCW = 4 * kXParts;            // compute warps per tile
constexpr int kXThreads = 64 + 32 * kXCW;
constexpr int kXWBytes = 32768;              // per head: N_h hi 8 KB | N_h lo 8 KB | Wg_h hi 8 KB | Wg_h lo 8 KB
constexpr int kXTile = 32768;                // xhat / Z tile: 8 planes hi (16 KB) | 8 planes lo (16 KB)
constexpr int kXNStages = 3, kXGStages = 2;   // N_h blocks are released after the first contraction of a head, Wg_h blocks after the second
constexpr int kXYBufs = 3;                   // Y accumulators in TMEM: the first contraction runs two heads ahead of the SIMT phase
constexpr int kXHalfW = 16384;               // one 64 x 64 weight, bf16 hi 8 KB | lo 8 KB
constexpr int kXSmem = (kXNStages + kXGStages) * kXHalfW + 2 * kXTile + 2 * kXTile;      // 212992
constexpr uint32_t kXColY = 0, kXColU = 64 * kXYBufs;

// derived W_qkg [1536, 64] (+ folded Q bias [512]) -> per head: N_h = Wk_h^T Wq_h (rows a, k = b), Wg_h, both as K-major
// canonical bf16 hi | lo blocks [k/8][64 rows][8]; v_h = Wk_h^T bq_h.  grid = 8 heads x 64 rows, 64 threads
__global__ void __launch_bounds__(64) prep_xform_kernel(const float* __restrict__ W, const float* __restrict__ bq, uint8_t* __restrict__ wx,
                                                        float* __restrict__ vx) {
  const int h = blockIdx.x >> 6, a = blockIdx.x & 63, b = threadIdx.x;
  const float* Wq = W + (int64_t)(h * kD) * kD;
  const float* Wk = W + (int64_t)(kH * kD + h * kD) * kD;
  const float* Wg = W + (int64_t)(2 * kH * kD + h * kD) * kD;
  float n = 0.f, v = 0.f;
  for (int m = 0; m < kD; ++m) {
    const float wk = __ldg(Wk + m * kD + a);
    n = fmaf(wk, __ldg(Wq + m * kD + b), n);
    v = fmaf(wk, __ldg(bq + h * kD + m), v);
  }
  if (b == 0) vx[h * kD + a] = v;
  uint8_t* blk = wx + (int64_t)h * kXWBytes;
  const int off = (b >> 3) * (64 * 16) + a * 16 + (b & 7) * 2;
  auto put = [&](uint8_t* base, float val) {
    const __nv_bfloat16 hi = __float2bfloat16_rn(val);
    const __nv_bfloat16 lo = __float2bfloat16_rn(val - __bfloat162float(hi));
    *reinterpret_cast<__nv_bfloat16*>(base + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(base + 8192 + off) = lo;
  };
  put(blk, n);
  put(blk + 16384, __ldg(Wg + a * kD + b));
}

// One K = 64 contraction (4 steps of K = 16) of a [128 x 64] A tile (planes of 2048 B, hi then lo 16 KB apart) with a
// [64 x 64] K-major weight block (k-groups of 1024 B, hi then lo 8 KB apart).  `passes` bf16 products per step: 3 =
// lo*hi + hi*lo + hi*hi (fp32-accurate split), 1 = hi*hi only.  Descriptors are built once and stepped by address.
__device__ __forceinline__ void umma_k64(uint32_t d, uint32_t a_addr, uint32_t b_addr, uint32_t idesc, bool fresh, int passes) {
  const uint64_t a_hi = make_smem_desc(a_addr, 2048, 128), b_hi = make_smem_desc(b_addr, 1024, 128);
  const uint64_t a_lo = desc_at(a_hi, 16384), b_lo = desc_at(b_hi, 8192);
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    if (passes == 3) {
      umma_bf16(d, desc_at(a_lo, ks * 4096), desc_at(b_hi, ks * 2048), idesc, (fresh && ks == 0) ? 0u : 1u);
      umma_bf16(d, desc_at(a_hi, ks * 4096), desc_at(b_lo, ks * 2048), idesc, 1u);
      umma_bf16(d, desc_at(a_hi, ks * 4096), desc_at(b_hi, ks * 2048), idesc, 1u);
    } else {
      umma_bf16(d, desc_at(a_hi, ks * 4096), desc_at(b_hi, ks * 2048), idesc, (fresh && ks == 0) ? 0u : 1u);
    }
  }
}

template <int L>
__global__ void __launch_bounds__(kXThreads, 1)
attn_xform_fwd_kernel(const uint8_t* __restrict__ xt, const float* __restrict__ xhat, const uint8_t* __restrict__ wx,
                      const float* __restrict__ vx, const float* __restrict__ b_dyn, const int64_t* __restrict__ x, float* __restrict__ U,
                      float* __restrict__ probs, int64_t T, const DropCfg drop, int passes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sN = smem;
  uint8_t* sG = smem + kXNStages * kXHalfW;
  uint8_t* sX = smem + (kXNStages + kXGStages) * kXHalfW;
  uint8_t* sZ = sX + 2 * kXTile;
  __shared__ uint64_t n_full[kXNStages], n_empty[kXNStages], g_full[kXGStages], g_empty[kXGStages];
  __shared__ uint64_t x_full[2], x_empty[2], y_full[kXYBufs], y_empty[kXYBufs], z_full[2], z_empty[2], u_full, u_empty;
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float sV[kH * kD];
  __shared__ __align__(16) float sBd[kD];
  __shared__ __align__(16) float sS[2][128][kXParts][4];
  constexpr int RPW = (32 / L) * L;
  constexpr int NB = L - 1;                           // neighbours per token
  static_assert(NB >= 1 && NB <= 4, "X-form attention keeps the neighbour rows in registers: L <= 5");
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = (T + 4 * RPW - 1) / (4 * RPW);

  if (warp == 0) tmem_alloc(&tmem_base_s, 256);
  if (tid == 32) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1);
      mbar_init(&z_full[i], kXCW); mbar_init(&z_empty[i], 1);
    }
    for (int i = 0; i < kXYBufs; ++i) { mbar_init(&y_full[i], 1); mbar_init(&y_empty[i], kXCW); }
    for (int i = 0; i < kXNStages; ++i) { mbar_init(&n_full[i], 1); mbar_init(&n_empty[i], 1); }
    for (int i = 0; i < kXGStages; ++i) { mbar_init(&g_full[i], 1); mbar_init(&g_empty[i], 1); }
    mbar_init(&u_full, 1); mbar_init(&u_empty, kXCW);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < kH * kD; i += kXThreads) sV[i] = __ldg(vx + i);
  if (tid < kD) sBd[tid] = __ldg(b_dyn + tid);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    if (elect_one()) {
      // N_h blocks run two heads ahead of the Wg_h blocks (global head numbering g = 8 k + h as in the MMA warp): the N ring
      // is released early in a head, the Wg ring only after the head's second contraction, and one thread feeds both
      const int64_t my_tiles = (ntiles - (int64_t)blockIdx.x + gridDim.x - 1) / gridDim.x;
      const int64_t total = my_tiles * kH;
      constexpr int kSkew = 2;
      for (int64_t i = 0; i < total + kSkew; ++i) {
        if (i < total) {
          const int64_t kk = i >> 3;
          const int h = (int)(i & 7), ns = (int)(i % kXNStages);
          if (h == 0) {
            const int xb = (int)(kk & 1);
            const int64_t tile = (int64_t)blockIdx.x + kk * gridDim.x;
            mbar_wait_backoff(&x_empty[xb], (uint32_t)((kk >> 1) & 1) ^ 1u);
            mbar_expect_tx(&x_full[xb], kXTile);
            const uint8_t* src = xt + tile * (int64_t)kXTileBytes;
            bulk_g2s(sX + xb * kXTile, src, 16384, &x_full[xb]);
            bulk_g2s(sX + xb * kXTile + 16384, src + kXHalfBytes, 16384, &x_full[xb]);
          }
          mbar_wait_backoff(&n_empty[ns], (uint32_t)((i / kXNStages) & 1) ^ 1u);
          mbar_expect_tx(&n_full[ns], kXHalfW);
          bulk_g2s(sN + ns * kXHalfW, wx + (int64_t)h * kXWBytes, kXHalfW, &n_full[ns]);
        }
        const int64_t gg = i - kSkew;
        if (gg >= 0) {
          const int gs = (int)(gg % kXGStages);
          mbar_wait_backoff(&g_empty[gs], (uint32_t)((gg / kXGStages) & 1) ^ 1u);
          mbar_expect_tx(&g_full[gs], kXHalfW);
          bulk_g2s(sG + gs * kXHalfW, wx + (int64_t)(gg & 7) * kXWBytes + kXHalfW, kXHalfW, &g_full[gs]);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(128, 64, false, false);
      // heads are numbered globally over this CTA's tiles (g = 8 k + h): the first contraction of head g + 2 is issued
      // before the second contraction of head g, so Y is always ready when the SIMT warps reach a head -- across tiles too
      const int64_t my_tiles = (ntiles - (int64_t)blockIdx.x + gridDim.x - 1) / gridDim.x;
      const int64_t total = my_tiles * kH;
      auto mma1 = [&](int64_t g) {                     // Y = xhat . N_h^T into Y buffer g % 3
        const int64_t kk = g >> 3;
        const int xb = (int)(kk & 1), yb = (int)(g % kXYBufs), ns = (int)(g % kXNStages);
        if ((g & 7) == 0) mbar_wait_backoff(&x_full[xb], (uint32_t)((kk >> 1) & 1));
        mbar_wait_backoff(&n_full[ns], (uint32_t)((g / kXNStages) & 1));
        mbar_wait_backoff(&y_empty[yb], (uint32_t)((g / kXYBufs) & 1) ^ 1u);
        tc_fence_after();
        umma_k64(tmem_base + kXColY + yb * 64, smem_u32(sX + xb * kXTile), smem_u32(sN + ns * kXHalfW), idesc, true, passes);
        umma_commit(&y_full[yb]);
        umma_commit(&n_empty[ns]);
      };
      if (total > 0) mma1(0);
      if (total > 1) mma1(1);
      for (int64_t g = 0; g < total; ++g) {
        if (g + 2 < total) mma1(g + 2);
        const int64_t k = g >> 3;
        const int h = (int)(g & 7);
        XTRACE(100 + h * 4 + 0);
        const int st = (int)(g & 1), gs = (int)(g % kXGStages);
        mbar_wait_backoff(&g_full[gs], (uint32_t)((g / kXGStages) & 1));
        mbar_wait_backoff(&z_full[st], (uint32_t)((g >> 1) & 1));
        if (h == 0) mbar_wait_backoff(&u_empty, (uint32_t)(k & 1) ^ 1u);
        XTRACE(100 + h * 4 + 1);
        tc_fence_after();
        umma_k64(tmem_base + kXColU, smem_u32(sZ + st * kXTile), smem_u32(sG + gs * kXHalfW), idesc, h == 0, passes);   // U += Z_h . Wg_h^T
        umma_commit(&z_empty[st]);
        umma_commit(&g_empty[gs]);
        XTRACE(100 + h * 4 + 2);
        if (h == kH - 1) { umma_commit(&u_full); umma_commit(&x_empty[k & 1]); }
      }
    }
  } else {
    const int part = (warp - 2) >> 2, q = warp & 3;
    const int r = q * 32 + lane;
    const bool live_lane = lane < RPW;
    const int gI = lane / L, pos = lane - gI * L;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int bar_id = 1 + q;
    int64_t k = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++k) {
      const int64_t t = (tile * 4 + q) * RPW + lane;
      const bool live = live_lane && t < T;
      // the L - 1 other tokens of this hyperedge: their xhat rows (this thread's column part) stay in registers for all heads.
      // They are read from the tile the producer already staged in shared memory for the tensor cores (bf16 hi + lo: the same
      // 2^-17 operand precision the contraction uses), not from HBM
      float nx[NB][kXCols];
      {
        const int xb = (int)(k & 1);
        mbar_wait(&x_full[xb], (uint32_t)((k >> 1) & 1));
        const uint8_t* xs = sX + xb * kXTile;
#pragma unroll
        for (int s = 0; s < NB; ++s) {
          const int rn = live_lane ? (r - pos + (pos + s + 1) % L) : r;       // tile row of the neighbour (same warp quarter)
#pragma unroll
          for (int j = 0; j < kXCols / 8; ++j) {
            const int plane = part * (kXCols / 8) + j;
            const uint4 hi = *reinterpret_cast<const uint4*>(xs + plane * 2048 + rn * 16);
            const uint4 lo = *reinterpret_cast<const uint4*>(xs + 16384 + plane * 2048 + rn * 16);
            const uint32_t hw[4] = {hi.x, hi.y, hi.z, hi.w}, lw[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              nx[s][j * 8 + 2 * e] = __uint_as_float(hw[e] << 16) + __uint_as_float(lw[e] << 16);
              nx[s][j * 8 + 2 * e + 1] = __uint_as_float(hw[e] & 0xFFFF0000u) + __uint_as_float(lw[e] & 0xFFFF0000u);
            }
          }
        }
      }
#pragma unroll 1
      for (int h = 0; h < kH; ++h) {
        const int64_t g = k * kH + h;
        const int st = (int)(g & 1), yb = (int)(g % kXYBufs);
        const uint32_t ph = (uint32_t)((g >> 1) & 1);
        if (tid == 64) XTRACE(200 + h * 8 + 0);
        mbar_wait(&y_full[yb], (uint32_t)((g / kXYBufs) & 1));
        if (tid == 64) XTRACE(200 + h * 8 + 1);
        tc_fence_after();
        float S[NB];
#pragma unroll
        for (int s = 0; s < NB; ++s) S[s] = 0.f;
#pragma unroll
        for (int hh = 0; hh < kXCols / 16; ++hh) {
          uint32_t yv[16];
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                       : "=r"(yv[0]), "=r"(yv[1]), "=r"(yv[2]), "=r"(yv[3]), "=r"(yv[4]), "=r"(yv[5]), "=r"(yv[6]), "=r"(yv[7]), "=r"(yv[8]),
                         "=r"(yv[9]), "=r"(yv[10]), "=r"(yv[11]), "=r"(yv[12]), "=r"(yv[13]), "=r"(yv[14]), "=r"(yv[15])
                       : "r"(taddr + kXColY + yb * 64 + part * kXCols + hh * 16)
                       : "memory");
          tmem_ld_wait(yv);
          if (hh == kXCols / 16 - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&y_empty[yb]);
          }
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const float yc = __uint_as_float(yv[c]) + sV[h * kD + part * kXCols + hh * 16 + c];
#pragma unroll
            for (int s = 0; s < NB; ++s) S[s] = fmaf(yc, nx[s][hh * 16 + c], S[s]);
          }
        }
        // the other column parts of the same row live in the warps of the same lane quarter: exchange the partial scores and
        // add them in a FIXED order, so every part of a row computes bit-identical attention weights
        if (tid == 64) XTRACE(200 + h * 8 + 2);
        {
          float4 mine = make_float4(S[0], NB > 1 ? S[NB > 1 ? 1 : 0] : 0.f, NB > 2 ? S[NB > 2 ? 2 : 0] : 0.f, NB > 3 ? S[NB > 3 ? 3 : 0] : 0.f);
          *reinterpret_cast<float4*>(&sS[st][r][part][0]) = mine;
          asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * kXParts) : "memory");
          float4 tot = *reinterpret_cast<const float4*>(&sS[st][r][0][0]);
#pragma unroll
          for (int pp = 1; pp < kXParts; ++pp) {
            const float4 o = *reinterpret_cast<const float4*>(&sS[st][r][pp][0]);
            tot.x += o.x; tot.y += o.y; tot.z += o.z; tot.w += o.w;
          }
          S[0] = tot.x;
          if (NB > 1) S[NB > 1 ? 1 : 0] = tot.y;
          if (NB > 2) S[NB > 2 ? 2 : 0] = tot.z;
          if (NB > 3) S[NB > 3 ? 3 : 0] = tot.w;
        }
        if (tid == 64) XTRACE(200 + h * 8 + 3);
        float mx = S[0];
#pragma unroll
        for (int s = 1; s < NB; ++s) mx = fmaxf(mx, S[s]);
        float sum = 0.f;
#pragma unroll
        for (int s = 0; s < NB; ++s) { S[s] = expf(S[s] - mx); sum += S[s]; }
        const float inv = 1.0f / sum;
#pragma unroll
        for (int s = 0; s < NB; ++s) S[s] *= inv;
        if (probs && live && part == 0) {
          float pr[4];
#pragma unroll
          for (int s = 0; s < 4; ++s) pr[s] = (s < NB) ? S[s < NB ? s : 0] : 0.f;
          *reinterpret_cast<float4*>(probs + (t * kH + h) * 4) = make_float4(pr[0], pr[1], pr[2], pr[3]);
        }
        // Z_h[i] = sum_j A_ij xhat_j (this thread's columns) -> bf16 hi | lo planes of the MMA operand tile
        if (tid == 64) XTRACE(200 + h * 8 + 4);
        mbar_wait(&z_empty[st], ph ^ 1u);
        if (tid == 64) XTRACE(200 + h * 8 + 5);
        uint8_t* zt = sZ + st * kXTile;
#pragma unroll
        for (int j = 0; j < kXCols / 8; ++j) {
          float z[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float acc = 0.f;
#pragma unroll
            for (int s = 0; s < NB; ++s) acc = fmaf(S[s], nx[s][j * 8 + c], acc);
            z[c] = acc;
          }
          uint4 hi, lo;
          split8(make_float4(z[0], z[1], z[2], z[3]), make_float4(z[4], z[5], z[6], z[7]), hi, lo);
          const int plane = part * (kXCols / 8) + j;
          sts16(zt + plane * 2048 + r * 16, hi);
          sts16(zt + 16384 + plane * 2048 + r * 16, lo);
        }
        if (tid == 64) XTRACE(200 + h * 8 + 6);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&z_full[st]);
        if (tid == 64) XTRACE(200 + h * 8 + 7);
      }
      // U = sum_h Z_h Wg_h^T: bias, dropout after fc1 (Modules.py:572), non-pad mask (Modules.py:614)
      if (tid == 64) XTRACE(300);
      mbar_wait(&u_full, (uint32_t)(k & 1));
      if (tid == 64) XTRACE(301);
      tc_fence_after();
      uint32_t uv[kXCols];
#pragma unroll
      for (int hh = 0; hh < kXCols / 16; ++hh) {
        uint32_t tv[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                     : "=r"(tv[0]), "=r"(tv[1]), "=r"(tv[2]), "=r"(tv[3]), "=r"(tv[4]), "=r"(tv[5]), "=r"(tv[6]), "=r"(tv[7]), "=r"(tv[8]),
                       "=r"(tv[9]), "=r"(tv[10]), "=r"(tv[11]), "=r"(tv[12]), "=r"(tv[13]), "=r"(tv[14]), "=r"(tv[15])
                     : "r"(taddr + kXColU + part * kXCols + hh * 16)
                     : "memory");
        tmem_ld_wait(tv);
#pragma unroll
        for (int c = 0; c < 16; ++c) uv[hh * 16 + c] = tv[c];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&u_empty);
      if (live) {
        const float m = x[t] != 0 ? 1.f : 0.f;
        float* dst = U + t * kD + part * kXCols;
#pragma unroll
        for (int c = 0; c < kXCols; c += 4) {
          const int cc = part * kXCols + c;
          float4 v = make_float4(__uint_as_float(uv[c]) + sBd[cc], __uint_as_float(uv[c + 1]) + sBd[cc + 1],
                                 __uint_as_float(uv[c + 2]) + sBd[cc + 2], __uint_as_float(uv[c + 3]) + sBd[cc + 3]);
          v = drop_apply4(drop, (uint64_t)t, (uint32_t)cc, v);
          *reinterpret_cast<float4*>(dst + c) = make_float4(v.x * m, v.y * m, v.z * m, v.w * m);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

template <int L>
int launch_xform_L(const uint8_t* xt, const float* xhat, const uint8_t* wx, const float* vx, const float* b_dyn, const int64_t* x, float* U,
                   float* probs, int64_t T, DropCfg drop, int passes, cudaStream_t s) {
  static bool once = false;
  if (!once) {
    if (int rc = check_cuda(cudaFuncSetAttribute(attn_xform_fwd_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, kXSmem),
                            "cudaFuncSetAttribute")) return rc;
    once = true;
  }
  const int64_t nt = num_atiles(T, L);
  const unsigned grid = (unsigned)(nt < kSMs ? nt : kSMs);
  attn_xform_fwd_kernel<L><<<grid, kXThreads, kXSmem, s>>>(xt, xhat, wx, vx, b_dyn, x, U, probs, T, drop, passes);
  MATCHA_CHECK_LAUNCH("attn_xform_fwd");
  return MATCHA_OK;
}

}  // namespace

#ifdef MATCHA_XFORM_TRACE
extern "C" int matcha_xform_trace(unsigned long long* out) {
  return cudaMemcpyFromSymbol(out, g_xtrace, sizeof(unsigned long long) * 4096) == cudaSuccess ? 0 : -2;
}
#endif

int launch_prep_xform(const float* W, const float* bq, void* wx, float* vx, cudaStream_t s) {
  prep_xform_kernel<<<kH * kD, 64, 0, s>>>(W, bq, reinterpret_cast<uint8_t*>(wx), vx);
  MATCHA_CHECK_LAUNCH("prep_xform");
  return MATCHA_OK;
}

int launch_attn_xform_fwd(const uint8_t* xhat_tiles, const float* xhat, const uint8_t* wx, const float* vx, const float* b_dyn,
                          const int64_t* x, float* U, float* probs, int64_t B, int L, DropCfg drop, int passes, cudaStream_t s) {
  if (B <= 0) return MATCHA_OK;
  const int64_t T = B * L;
  switch (L) {
    case 2: return launch_xform_L<2>(xhat_tiles, xhat, wx, vx, b_dyn, x, U, probs, T, drop, passes, s);
    case 3: return launch_xform_L<3>(xhat_tiles, xhat, wx, vx, b_dyn, x, U, probs, T, drop, passes, s);
    case 4: return launch_xform_L<4>(xhat_tiles, xhat, wx, vx, b_dyn, x, U, probs, T, drop, passes, s);
    case 5: return launch_xform_L<5>(xhat_tiles, xhat, wx, vx, b_dyn, x, U, probs, T, drop, passes, s);
    default: set_error("attn_xform_fwd: padded width L=%d unsupported (2..5)", L); return MATCHA_ERR_ARG;
  }
}

}  // namespace matcha
