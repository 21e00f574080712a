// tcgen05 kernels of the QKG projection working on PRE-SPLIT TILES (see rowwise.cuh): the producers of the
// operands (LayerNorm forward, attention backward, parameter preparation) already emit bf16 hi | lo tiles in
// the canonical UMMA layout, so these kernels contain no staging arithmetic at all:
//     warp 0   one thread issues cp.async.bulk (TMA engine, no tensor map: tiles are contiguous) into a ring
//     warp 1   one thread issues tcgen05.mma (three bf16 passes: lo*hi + hi*lo + hi*hi) into TMEM accumulators
//     warps 2+ drain accumulators with tcgen05.ld and write fp32 results
// synchronised only by mbarriers (expect_tx for the copies, tcgen05.commit for the MMAs).
//   forward : QKG[T,1536]  = xhat . Wqkg^T + bias        D[128 features, 128 tokens] per (tile, chunk): coalesced stores
//   dgrad   : dxhat[T,64]  = dQKG . Wqkg                  D[128 tokens, 64], K = 1536 streamed in 24 chunks
//   wgrad   : dWqkg[1536,64] += dQKG^T . xhat             D[128 features, 80] x 4 per CTA, K = tokens; column 64 of
//             the xhat tile is a ones column, so the same MMAs also produce the bias gradient
#include "rowwise.cuh"
#include "tc_common.cuh"

namespace matcha {
namespace {

// ------------------------------------------------------------------------------------------
// parameter preparation: W [1536, 64] fp32 -> per 64-feature chunk, MN-major (n = output column c, k = feature f)
// [f/8][c/8][f%8][8 c] bf16 hi (8 KB) | lo (8 KB): the B operand of the data-gradient MMA
// ------------------------------------------------------------------------------------------
constexpr int kWTChunkBytes = 16384;
__global__ void split_wT_kernel(const float* __restrict__ W, uint8_t* __restrict__ out) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;    // unit = (feature f, column group g)
  if (u >= kQKG * 8) return;
  const int f = u >> 3, g = u & 7;
  const float* src = W + (int64_t)f * kD + g * 8;
  uint4 hi, lo;
  split8(__ldg(reinterpret_cast<const float4*>(src)), __ldg(reinterpret_cast<const float4*>(src + 4)), hi, lo);
  uint8_t* chunk = out + (f >> 6) * kWTChunkBytes;
  const int fl = f & 63;
  const int off = (fl >> 3) * 1024 + g * 128 + (fl & 7) * 16;
  *reinterpret_cast<uint4*>(chunk + off) = hi;
  *reinterpret_cast<uint4*>(chunk + 8192 + off) = lo;
}

// ==========================================================================================
// forward
// ==========================================================================================
constexpr int kFThreads = 320;                       // producer warp, MMA warp, 8 epilogue warps
constexpr int kFWStages = 3, kFAccStages = 4;
constexpr int kFWBytes = 32768;                      // 128 features x 64: hi 16 KB | lo 16 KB (gemm_tc.cu split_weights_k64)
constexpr int kFXBytes = 32768;                      // first 8 planes of both halves of an xhat tile
constexpr int kFSmem = kFWStages * kFWBytes + 2 * kFXBytes;

__global__ void __launch_bounds__(kFThreads, 1) qkg_fwd_tiles_kernel(const uint8_t* __restrict__ xt,
                                                                     const uint8_t* __restrict__ wsplit,
                                                                     const float* __restrict__ bias, float* __restrict__ C,
                                                                     int64_t T) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sW = smem;
  uint8_t* sX = smem + kFWStages * kFWBytes;
  __shared__ uint64_t w_full[kFWStages], w_empty[kFWStages], x_full[2], x_empty[2], acc_full[kFAccStages],
      acc_empty[kFAccStages];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int nchunk = kQKG / 128;
  const int64_t ntiles = num_token_tiles(T);
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 32) {
    for (int i = 0; i < kFWStages; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); }
    for (int i = 0; i < kFAccStages; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    if (elect_one()) {
      int ws = 0, xs = 0; uint32_t wp = 0, xp = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        mbar_wait(&x_empty[xs], xp ^ 1);
        mbar_expect_tx(&x_full[xs], kFXBytes);
        const uint8_t* src = xt + tile * (int64_t)kXTileBytes;
        bulk_g2s(sX + xs * kFXBytes, src, 16384, &x_full[xs]);
        bulk_g2s(sX + xs * kFXBytes + 16384, src + kXHalfBytes, 16384, &x_full[xs]);
        if (++xs == 2) { xs = 0; xp ^= 1; }
        for (int fc = 0; fc < nchunk; ++fc) {
          mbar_wait(&w_empty[ws], wp ^ 1);
          mbar_expect_tx(&w_full[ws], kFWBytes);
          bulk_g2s(sW + ws * kFWBytes, wsplit + (int64_t)fc * kFWBytes, kFWBytes, &w_full[ws]);
          if (++ws == kFWStages) { ws = 0; wp ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(128, 128, false, false);
      int ws = 0, as = 0, xs = 0; uint32_t wp = 0, ap = 0, xp = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        mbar_wait(&x_full[xs], xp);
        const uint32_t xh = smem_u32(sX + xs * kFXBytes), xl = xh + 16384;
        for (int fc = 0; fc < nchunk; ++fc) {
          mbar_wait(&w_full[ws], wp);
          mbar_wait(&acc_empty[as], ap ^ 1);
          tc_fence_after();
          const uint32_t wh = smem_u32(sW + ws * kFWBytes), wl = wh + 16384;
          const uint32_t d = tmem_base + as * 128;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_x3(d, wh + ks * 4096, wl + ks * 4096, xh + ks * 4096, xl + ks * 4096, 2048, 2048, idesc, ks == 0);
          umma_commit(&w_empty[ws]);
          umma_commit(&acc_full[as]);
          if (++ws == kFWStages) { ws = 0; wp ^= 1; }
          if (++as == kFAccStages) { as = 0; ap ^= 1; }
        }
        umma_commit(&x_empty[xs]);
        if (++xs == 2) { xs = 0; xp ^= 1; }
      }
    }
  } else {
    const int ew = warp - 2;               // 0..7
    const int lane_grp = warp & 3;         // TMEM lane quarter readable by this warp
    const int tok_half = ew >> 2;
    int as = 0; uint32_t ap = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      for (int fc = 0; fc < nchunk; ++fc) {
        mbar_wait(&acc_full[as], ap);
        tc_fence_after();
        const int feat = fc * 128 + lane_grp * 32 + lane;
        const float b = bias ? __ldg(bias + feat) : 0.f;
        float v0[32], v1[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + as * 128 + tok_half * 64;
        tmem_ld32(taddr, v0);
        tmem_ld32(taddr + 32, v1);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[as]);
        const int64_t t0 = tile * 128 + tok_half * 64;
        float* dst = C + t0 * kQKG + feat;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (t0 + j < T) dst[(int64_t)j * kQKG] = v0[j] + b;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (t0 + 32 + j < T) dst[(int64_t)(32 + j) * kQKG] = v1[j] + b;
        if (++as == kFAccStages) { as = 0; ap ^= 1; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ==========================================================================================
// data gradient
// ==========================================================================================
constexpr int kDThreads = 192;                       // producer, MMA, 4 epilogue warps
constexpr int kDStages = 4;
constexpr int kDStageBytes = kGTileBytes + kWTChunkBytes;      // 48 KB
constexpr int kDSmem = kDStages * kDStageBytes;                // 192 KB

__global__ void __launch_bounds__(kDThreads, 1) qkg_dgrad_tiles_kernel(const uint8_t* __restrict__ gt,
                                                                       const uint8_t* __restrict__ wT,
                                                                       float* __restrict__ dxhat, int64_t T) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[kDStages], empty[kDStages], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = num_token_tiles(T);
  if (warp == 0) tmem_alloc(&tmem_base_s, 128);
  if (tid == 32) {
    for (int i = 0; i < kDStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    if (elect_one()) {
      int st = 0; uint32_t ph = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int fc = 0; fc < kGChunks; ++fc) {
          mbar_wait(&empty[st], ph ^ 1);
          mbar_expect_tx(&full[st], kDStageBytes);
          uint8_t* dst = smem + st * kDStageBytes;
          bulk_g2s(dst, gt + (tile * kGChunks + fc) * (int64_t)kGTileBytes, kGTileBytes, &full[st]);
          bulk_g2s(dst + kGTileBytes, wT + (int64_t)fc * kWTChunkBytes, kWTChunkBytes, &full[st]);
          if (++st == kDStages) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(128, 64, false, true);
      int st = 0, as = 0; uint32_t ph = 0, ap = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        mbar_wait(&acc_empty[as], ap ^ 1);
        const uint32_t d = tmem_base + as * 64;
        for (int fc = 0; fc < kGChunks; ++fc) {
          mbar_wait(&full[st], ph);
          tc_fence_after();
          const uint32_t gh = smem_u32(smem + st * kDStageBytes), gl = gh + kGHalfBytes;
          const uint32_t wh = gh + kGTileBytes, wl = wh + 8192;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)   // A: tokens x features, K-major.  B: W chunk, MN-major (n = column, k = feature)
            umma_x3s(d, gh + ks * 4096, gl + ks * 4096, wh + ks * 2048, wl + ks * 2048, 2048, 128, 1024, 128, idesc,
                     fc == 0 && ks == 0);
          umma_commit(&empty[st]);
          if (++st == kDStages) { st = 0; ph ^= 1; }
        }
        umma_commit(&acc_full[as]);
        if (++as == 2) { as = 0; ap ^= 1; }
      }
    }
  } else {
    const int lane_grp = warp & 3;
    int as = 0; uint32_t ap = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      mbar_wait(&acc_full[as], ap);
      tc_fence_after();
      float v0[32], v1[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + as * 64;
      tmem_ld32(taddr, v0);
      tmem_ld32(taddr + 32, v1);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[as]);
      const int64_t t = tile * 128 + lane_grp * 32 + lane;
      if (t < T) {
        float* dst = dxhat + t * kD;
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v0[j], v0[j + 1], v0[j + 2], v0[j + 3]);
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + 32 + j) = make_float4(v1[j], v1[j + 1], v1[j + 2], v1[j + 3]);
      }
      if (++as == 2) { as = 0; ap ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 128);
}

// ==========================================================================================
// weight gradient (+ bias gradient through the ones column)
// ==========================================================================================
constexpr int kWThreads = 192;
constexpr int kWGStage = 65536;                      // two 64-feature chunks: hi0 | hi1 | lo0 | lo1
constexpr int kWSmem = 2 * kWGStage + 2 * kXTileBytes;          // 128 KB + 80 KB
constexpr int kWN = 80;                              // 64 columns + ones column + padding to a multiple of 16

__global__ void __launch_bounds__(kWThreads, 1) qkg_wgrad_tiles_kernel(const uint8_t* __restrict__ gt,
                                                                       const uint8_t* __restrict__ xt, int64_t T,
                                                                       int64_t tiles_per_split, float* __restrict__ part,
                                                                       float* __restrict__ part_cs) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sG = smem;
  uint8_t* sX = smem + 2 * kWGStage;
  __shared__ uint64_t g_full[2], g_empty[2], x_full[2], x_empty[2], done;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mg = blockIdx.x, split = blockIdx.y;      // 512-feature group, token split
  const int64_t ntiles = num_token_tiles(T);
  const int64_t t_begin = (int64_t)split * tiles_per_split;
  const int64_t t_end = (t_begin + tiles_per_split < ntiles) ? t_begin + tiles_per_split : ntiles;
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 32) {
    for (int i = 0; i < 2; ++i) { mbar_init(&g_full[i], 1); mbar_init(&g_empty[i], 1); mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); }
    mbar_init(&done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    if (elect_one()) {
      int gs = 0, xs = 0; uint32_t gp = 0, xp = 0;
      for (int64_t tile = t_begin; tile < t_end; ++tile) {
        mbar_wait(&x_empty[xs], xp ^ 1);
        mbar_expect_tx(&x_full[xs], kXTileBytes);
        bulk_g2s(sX + xs * kXTileBytes, xt + tile * (int64_t)kXTileBytes, kXTileBytes, &x_full[xs]);
        if (++xs == 2) { xs = 0; xp ^= 1; }
        for (int j = 0; j < 4; ++j) {
          mbar_wait(&g_empty[gs], gp ^ 1);
          mbar_expect_tx(&g_full[gs], kWGStage);
          const uint8_t* c0 = gt + (tile * kGChunks + mg * 8 + 2 * j) * (int64_t)kGTileBytes;
          const uint8_t* c1 = c0 + kGTileBytes;
          uint8_t* dst = sG + gs * kWGStage;
          bulk_g2s(dst, c0, kGHalfBytes, &g_full[gs]);
          bulk_g2s(dst + 16384, c1, kGHalfBytes, &g_full[gs]);
          bulk_g2s(dst + 32768, c0 + kGHalfBytes, kGHalfBytes, &g_full[gs]);
          bulk_g2s(dst + 49152, c1 + kGHalfBytes, kGHalfBytes, &g_full[gs]);
          if (++gs == 2) { gs = 0; gp ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(128, kWN, true, true);
      int gs = 0, xs = 0; uint32_t gp = 0, xp = 0;
      for (int64_t tile = t_begin; tile < t_end; ++tile) {
        mbar_wait(&x_full[xs], xp);
        const uint32_t xh = smem_u32(sX + xs * kXTileBytes), xl = xh + kXHalfBytes;
        for (int j = 0; j < 4; ++j) {
          mbar_wait(&g_full[gs], gp);
          tc_fence_after();
          const uint32_t gh = smem_u32(sG + gs * kWGStage), gl = gh + 32768;
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)   // K = 16 tokens per step; both operands MN-major: LBO 128 (next 8 tokens), SBO 2048 (next plane)
            umma_x3s(tmem_base + j * 128, gh + ks * 256, gl + ks * 256, xh + ks * 256, xl + ks * 256, 128, 2048, 128, 2048,
                     idesc, tile == t_begin && ks == 0);
          umma_commit(&g_empty[gs]);
          if (++gs == 2) { gs = 0; gp ^= 1; }
        }
        umma_commit(&x_empty[xs]);
        if (++xs == 2) { xs = 0; xp ^= 1; }
      }
      umma_commit(&done);
    }
  } else {
    const int lane_grp = warp & 3;
    const bool any = t_end > t_begin;
    if (any) mbar_wait(&done, 0);
    tc_fence_after();
    const int64_t M = kQKG;
    float* out = part + ((int64_t)split * M + (int64_t)mg * 512) * 64;
#pragma unroll 1
    for (int j = 0; j < 4; ++j) {
      const int row = j * 128 + lane_grp * 32 + lane;
      float* dst = out + (int64_t)row * 64;
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        float v[32];
        if (any) {
          tmem_ld32(tmem_base + ((uint32_t)(lane_grp * 32) << 16) + j * 128 + cc * 32, v);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + cc * 32 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
      float c16[16];
      if (any) {
        tmem_ld16(tmem_base + ((uint32_t)(lane_grp * 32) << 16) + j * 128 + 64, c16);
      } else {
        c16[0] = 0.f;
      }
      part_cs[(int64_t)split * M + (int64_t)mg * 512 + row] = c16[0];     // column 64 = sum over tokens (ones column)
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

__global__ void wgrad_reduce_kernel(const float* __restrict__ part, const float* __restrict__ part_cs, int splits,
                                    float* __restrict__ dW, float* __restrict__ dbias, int64_t dbias_n) {
  const int64_t total = (int64_t)kQKG * 64;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total + kQKG; i += (int64_t)gridDim.x * blockDim.x) {
    if (i < total) {
      float s = 0.f;
      for (int sp = 0; sp < splits; ++sp) s += part[(int64_t)sp * total + i];
      dW[i] += s;
    } else {
      const int64_t m = i - total;
      if (dbias && m < dbias_n) {
        float s = 0.f;
        for (int sp = 0; sp < splits; ++sp) s += part_cs[(int64_t)sp * kQKG + m];
        dbias[m] += s;
      }
    }
  }
}

template <typename K>
int set_smem_attr(K kernel, int bytes) {
  return check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes), "cudaFuncSetAttribute");
}
}  // namespace

int launch_split_wT(const float* W, void* out, cudaStream_t s) {
  split_wT_kernel<<<(kQKG * 8 + 255) / 256, 256, 0, s>>>(W, reinterpret_cast<uint8_t*>(out));
  MATCHA_CHECK_LAUNCH("split_wT");
  return MATCHA_OK;
}

int tc_qkg_forward_tiles(const uint8_t* xhat_tiles, const uint8_t* w_split, const float* bias, float* QKG, int64_t T,
                         cudaStream_t s) {
  static bool once = false;
  if (!once) { if (int rc = set_smem_attr(qkg_fwd_tiles_kernel, kFSmem)) return rc; once = true; }
  const int64_t ntiles = num_token_tiles(T);
  qkg_fwd_tiles_kernel<<<(unsigned)(ntiles < kSMs ? ntiles : kSMs), kFThreads, kFSmem, s>>>(xhat_tiles, w_split, bias, QKG, T);
  MATCHA_CHECK_LAUNCH("qkg_fwd_tiles");
  return MATCHA_OK;
}

int tc_qkg_dgrad_tiles(const uint8_t* dqkg_tiles, const uint8_t* wT_split, float* dxhat, int64_t T, cudaStream_t s) {
  static bool once = false;
  if (!once) { if (int rc = set_smem_attr(qkg_dgrad_tiles_kernel, kDSmem)) return rc; once = true; }
  const int64_t ntiles = num_token_tiles(T);
  qkg_dgrad_tiles_kernel<<<(unsigned)(ntiles < kSMs ? ntiles : kSMs), kDThreads, kDSmem, s>>>(dqkg_tiles, wT_split, dxhat, T);
  MATCHA_CHECK_LAUNCH("qkg_dgrad_tiles");
  return MATCHA_OK;
}

int tc_qkg_wgrad_tiles(const uint8_t* dqkg_tiles, const uint8_t* xhat_tiles, float* scratch, int64_t scratch_floats,
                       float* dW, float* dbias, int64_t dbias_n, int64_t T, cudaStream_t s) {
  static bool once = false;
  if (!once) { if (int rc = set_smem_attr(qkg_wgrad_tiles_kernel, kWSmem)) return rc; once = true; }
  const int64_t ntiles = num_token_tiles(T);
  int splits = kSMs / 3;                                   // 3 feature groups x 49 token splits = 147 CTAs
  if (splits > ntiles) splits = (int)ntiles;
  if (splits > kTcMaxSplits) splits = kTcMaxSplits;
  const int64_t tps = (ntiles + splits - 1) / splits;
  splits = (int)((ntiles + tps - 1) / tps);
  MATCHA_REQUIRE(scratch && scratch_floats >= gemm_tc_scratch_floats(kQKG), "qkg wgrad: scratch too small");
  float* part = scratch;
  float* part_cs = scratch + (int64_t)kTcMaxSplits * kQKG * 64;
  dim3 grid(3, (unsigned)splits);
  qkg_wgrad_tiles_kernel<<<grid, kWThreads, kWSmem, s>>>(dqkg_tiles, xhat_tiles, T, tps, part, part_cs);
  MATCHA_CHECK_LAUNCH("qkg_wgrad_tiles");
  wgrad_reduce_kernel<<<kSMs * 2, 256, 0, s>>>(part, part_cs, splits, dW, dbias, dbias_n);
  MATCHA_CHECK_LAUNCH("wgrad_reduce");
  return MATCHA_OK;
}

}  // namespace matcha
