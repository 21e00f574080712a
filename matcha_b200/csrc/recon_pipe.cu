// Pipelined form of the fused reconstruction head's gradient pass (Modules.py:192-199 and its backward):
//   pred = tanh(E) . Rw^T + rb -> diff vs the z-scored inter-chromosomal target rows -> loss,
//   dRw += gdiff^T . tanh(E),  drb += sum_t gdiff,  dtE += gdiff . Rw
// Same arithmetic as recon_tc.cu (bf16x3 split contractions, fp32 accumulation), different schedule.  recon_tc.cu runs one
// (128 tokens x 128 target columns) unit as a serial chain  stage -> MMA -> target loads -> SIMT -> MMA -> scatter  with one
// CTA per SM (profiles/r02_ncu_recon.md: 6.6 us per unit, 7 % of the HBM peak at cfg3) on a (token split x column block)
// grid whose size is not a multiple of the SM count.  Here:
//   * a pre-pass writes tanh(E) of the ELIGIBLE tokens once as bf16 hi | lo operand tiles (+ the token / target-row index of
//     every tile row), so a unit's A operand is two bulk copies instead of a scattered gather + 8192 tanhf per column block;
//   * the units (column block major) are cut into 148 equal contiguous ranges: one persistent CTA per SM, dRw resident in
//     TMEM while the CTA stays inside a column block (a range touches at most two);
//   * the first contraction is TRANSPOSED:  P^T[col, tok] = Rw_block . tanh(E)^T, so TMEM lane = target column and a warp
//     reads 32 consecutive floats of one target row per load instruction (coalesced, no shared-memory transpose); every
//     thread has its 64 target loads of the NEXT unit in flight while the tensor cores work on the current one;
//   * gdiff^T is stored [tok/8][col][8 tok]: K-major for dRw (K = tokens) and, with the strides swapped, MN-major for dtE;
//   * warp roles: 8 compute warps, one bulk-copy producer, one MMA issuer; P^T and dtE double buffered in TMEM, the operand
//     tiles in a ring of three (released by the tensor pipe's own commit); the bias gradient is summed in registers and
//     dtE leaves as 16-byte vector reductions.
// Measured (scripts/dev/recon_trace.py, cfg3, chr1 = 20 column blocks): 2.9 us per unit -- the 60 bf16x3 MMAs of a unit
// take 2.0 us at the sustained tensor rate, the SIMT phase 2.7 us while its target rows miss L2 (1.25 us when they hit).
#include "rowwise.cuh"
#include "tc_common.cuh"

namespace matcha {
#ifdef MATCHA_RECON_TRACE
__device__ unsigned long long g_rtrace[4096];
// slot layout: unit i (first 40 of CTA 0's first segment) x 16 events
#define RTRACE(ev) do { if (blockIdx.x == 0 && k == 0 && i < 40) { unsigned long long _t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t)); g_rtrace[i * 16 + (ev)] = _t; } } while (0)
#else
#define RTRACE(ev) do { } while (0)
#endif
namespace {

constexpr int kPThreads = 320;                        // warps 0-7 compute, warp 8 producer, warp 9 MMA issuer
constexpr int kPRec = 32768 + 1024;                   // tile record: hi 16 KB | lo 16 KB | tok[128] | row[128]  (int32)
constexpr int kPA = 32768, kPAHalf = 16384;           // sA buffer: hi 8 planes | lo 8 planes
constexpr int kPW = 32768;                            // sW: 128 target columns x 64 features, hi 16 KB | lo 16 KB
constexpr int kPG = 65536;                            // sG: gdiff^T, hi 32 KB | lo 32 KB
constexpr int kPABufs = 3;                            // operand ring: the bulk copies of unit u + 3 start when unit u retires
constexpr int kPTab = 1024;                           // index words of a tile (ride with the operand buffer)
constexpr int kPSmem = kPABufs * (kPA + kPTab) + kPW + kPG;            // 199 680
constexpr uint32_t kColP = 0, kColDT = 256, kColDW = 384;              // P^T 2 x 128 | dtE 2 x 64 | dRw 64

struct PipeArgs {
  const uint8_t* tiles;
  const float* inter; int64_t inter_ld;
  int64_t rs, re;
  const float* Rw; const float* rb;
  const int32_t* counts; const int32_t* group_off; int rchrom, n_chrom, ncb;
  float* recon_out; float* dRw; float* drb; float* dtE;
};

// tanh(E) rows of the eligible tokens (bucketed list minus the drawn chromosome's bucket) -> operand tiles + index tables
__global__ void __launch_bounds__(256) recon_tiles_kernel(const float* __restrict__ E, const int64_t* __restrict__ x,
                                                          const int32_t* __restrict__ perm, const int32_t* __restrict__ group_off,
                                                          const int32_t* __restrict__ counts, int rchrom, int n_chrom,
                                                          uint8_t* __restrict__ tiles) {
  const int64_t lo_n = group_off[rchrom], skip = counts[rchrom];
  const int64_t elig = (int64_t)group_off[n_chrom] - skip;
  const int64_t ntiles = (elig + 127) / 128;
  const int p = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    uint8_t* rec = tiles + tile * (int64_t)kPRec;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int row = j * 32 + lane;
      const int64_t e = tile * 128 + row;
      const bool ok = e < elig;
      const int64_t t = ok ? (int64_t)perm[e < lo_n ? e : e + skip] : -1;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
      if (ok) {
        const float4* src = reinterpret_cast<const float4*>(E + t * 64 + p * 8);
        a = __ldg(src); b = __ldg(src + 1);
        a = make_float4(tanhf(a.x), tanhf(a.y), tanhf(a.z), tanhf(a.w));
        b = make_float4(tanhf(b.x), tanhf(b.y), tanhf(b.z), tanhf(b.w));
      }
      uint4 hi, lo;
      split8(a, b, hi, lo);
      sts16(rec + p * 2048 + row * 16, hi);
      sts16(rec + 16384 + p * 2048 + row * 16, lo);
      if (p == 0) {
        reinterpret_cast<int32_t*>(rec + 32768)[row] = (int32_t)t;
        reinterpret_cast<int32_t*>(rec + 32768 + 512)[row] = ok ? (int32_t)(x[t] - 1) : 0;      // row 0: loaded, masked out
      }
    }
  }
}

// 16-byte vector reduction: one instruction adds four consecutive floats of a row
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(kPThreads, 1) recon_pipe_kernel(const PipeArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sA = smem;                                   // ring of kPABufs operand buffers
  uint8_t* sW = smem + kPABufs * kPA;
  uint8_t* sG = sW + kPW;
  uint8_t* sTabB = sG + kPG;                            // ring of index tables: [tok 128 | row 128] int32
  __shared__ uint64_t a_full[kPABufs], a_free[kPABufs], p_full[2], d_full[2], g_full;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t n_r = a.re - a.rs;
  const int64_t skip = a.counts[a.rchrom];
  const int64_t elig = (int64_t)a.group_off[a.n_chrom] - skip;
  const int64_t ntiles = (elig + 127) / 128;
  const int64_t total = ntiles * a.ncb;
  const int64_t u_begin = total * blockIdx.x / gridDim.x, u_end = total * (blockIdx.x + 1) / gridDim.x;
  const float gscale = elig > 0 ? 200.0f / ((float)elig * (float)n_r) : 0.f;

  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 256) {
    for (int i = 0; i < kPABufs; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_free[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&p_full[i], 1); mbar_init(&d_full[i], 1); }
    mbar_init(&g_full, 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t wh = smem_u32(sW), wl = wh + 16384;
  const uint32_t gh = smem_u32(sG), gl = gh + 32768;
  float loss = 0.f;
  int k = 0;                                              // units this CTA has finished (every role counts alike)

  for (int64_t u0 = u_begin; u0 < u_end;) {
    const int64_t cb = u0 / ntiles;
    const int64_t seg_end = (cb + 1) * ntiles < u_end ? (cb + 1) * ntiles : u_end;
    const int n = (int)(seg_end - u0), tile0 = (int)(u0 - cb * ntiles);
    if (warp < 8) {   // Rw block -> sW (row = target column; this thread: features [32 h, 32 h + 32)), zero rows beyond n_r
      const int r = tid & 127, h = tid >> 7;
      const int64_t col = cb * 128 + r;
      float v[32];
      if (col < n_r) {
        const float* src = a.Rw + col * 64 + h * 32;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 e = __ldg(reinterpret_cast<const float4*>(src) + j);
          v[4 * j] = e.x; v[4 * j + 1] = e.y; v[4 * j + 2] = e.z; v[4 * j + 3] = e.w;
        }
      } else {
#pragma unroll
        for (int c = 0; c < 32; ++c) v[c] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 hi, lo;
        split8(make_float4(v[8 * j], v[8 * j + 1], v[8 * j + 2], v[8 * j + 3]),
               make_float4(v[8 * j + 4], v[8 * j + 5], v[8 * j + 6], v[8 * j + 7]), hi, lo);
        sts16(sW + (h * 4 + j) * 2048 + r * 16, hi);
        sts16(sW + 16384 + (h * 4 + j) * 2048 + r * 16, lo);
      }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp == 8) {
      if (elect_one()) {
        for (int i = 0; i < n; ++i) {
          const int kk = k + i;
          const int sb = (int)(kk % kPABufs);
          mbar_wait_backoff(&a_free[sb], (uint32_t)((kk / kPABufs) & 1) ^ 1u);
          mbar_expect_tx(&a_full[sb], kPA + kPTab);
          const uint8_t* rec = a.tiles + (int64_t)(tile0 + i) * kPRec;
          bulk_g2s(sTabB + sb * kPTab, rec + 32768, kPTab, &a_full[sb]);
          bulk_g2s(sA + sb * kPA, rec, kPA, &a_full[sb]);
        }
      }
    } else if (warp == 9) {
      if (elect_one()) {
        constexpr uint32_t idescP = make_idesc(128, 128, false, false);
        constexpr uint32_t idescD = make_idesc(128, 64, true, true);
        constexpr uint32_t idescW = make_idesc(128, 64, false, true);
        auto mma1 = [&](int i) {          // P^T[128 col, 128 tok] = Rw block . tanh(E)^T   (K = 64)
          const int kk = k + i;
          const int b = (int)(kk & 1), sb = (int)(kk % kPABufs);
          mbar_wait_backoff(&a_full[sb], (uint32_t)((kk / kPABufs) & 1));
          tc_fence_after();
          const uint32_t ah = smem_u32(sA + sb * kPA), al = ah + kPAHalf;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_x3s(tmem_base + kColP + b * 128, wh + ks * 4096, wl + ks * 4096, ah + ks * 4096, al + ks * 4096, 2048, 128, 2048, 128,
                     idescP, ks == 0);
          umma_commit(&p_full[b]);
        };
        mma1(0);
        for (int i = 0; i < n; ++i) {
          if (i + 1 < n) mma1(i + 1);
          const int kk = k + i;
          const int b = (int)(kk & 1), sb = (int)(kk % kPABufs);
          RTRACE(8);
          mbar_wait_backoff(&g_full, (uint32_t)(kk & 1));
          RTRACE(9);
          tc_fence_after();
          const uint32_t ah = smem_u32(sA + sb * kPA), al = ah + kPAHalf;
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)      // dtE[128 tok, 64] = gdiff[128 tok, 128 col] . Rw[128 col, 64]   (A, B MN-major)
            umma_x3s(tmem_base + kColDT + b * 64, gh + ks * 256, gl + ks * 256, wh + ks * 256, wl + ks * 256, 128, 2048, 128, 2048,
                     idescD, ks == 0);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)      // dRw[128 col, 64] += gdiff^T[128 col, 128 tok] . tanh(E)[128 tok, 64]
            umma_x3s(tmem_base + kColDW, gh + ks * 4096, gl + ks * 4096, ah + ks * 256, al + ks * 256, 2048, 128, 128, 2048, idescW,
                     i == 0 && ks == 0);
          umma_commit(&d_full[b]);
          umma_commit(&a_free[sb]);
          RTRACE(10);
        }
      }
    } else {
      const int q = warp & 3, h = warp >> 2;
      const int c = q * 32 + lane;                           // TMEM lane: target column of the block / token row of the tile
      const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
      const int64_t col = cb * 128 + c;
      const bool col_ok = col < n_r;
      const float rbc = col_ok ? __ldg(a.rb + col) : 0.f;
      const float cmask = col_ok ? 1.f : 0.f;
      const float* tbase = a.inter + (a.rs - 1) + (col_ok ? col : n_r - 1);      // columns beyond n_r: loaded, masked out
      float tv[64];
      float bsum = 0.f;                                      // bias gradient of this thread's column over its 64 tokens per unit
      int32_t t_next = -1;
      // 64 target values of this thread's column: tokens [64 h, 64 h + 64) of a tile; one coalesced 128-byte row segment per
      // warp instruction, all 64 loads in flight.  Row indices: broadcast reads of the tile's index table.  (Unconditional
      // loads: rows past the end of the list read row 0 and are masked when the difference is formed -- a predicated load
      // makes ptxas select on the result and wait for every load where it is issued.)
      auto load_targets = [&](int kk, int half) {
        const int sb = (int)(kk % kPABufs);
        if (half == 0) mbar_wait(&a_full[sb], (uint32_t)((kk / kPABufs) & 1));
        const int32_t* tab = reinterpret_cast<const int32_t*>(sTabB + sb * kPTab);
        if (half == 0) t_next = tab[c];
        const int4* rows = reinterpret_cast<const int4*>(tab + 128 + h * 64 + half * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int4 r4 = rows[j];
          tv[half * 32 + 4 * j] = __ldg(tbase + (int64_t)r4.x * a.inter_ld);
          tv[half * 32 + 4 * j + 1] = __ldg(tbase + (int64_t)r4.y * a.inter_ld);
          tv[half * 32 + 4 * j + 2] = __ldg(tbase + (int64_t)r4.z * a.inter_ld);
          tv[half * 32 + 4 * j + 3] = __ldg(tbase + (int64_t)r4.w * a.inter_ld);
        }
      };
      // dtE rows of a finished unit: TMEM lane = token row, this thread: features [32 h, 32 h + 32) -> eight 16-byte
      // vector reductions into the token's row
      auto scatter_dte = [&](int kk, int32_t t) {
        const int b = (int)(kk & 1);
        mbar_wait(&d_full[b], (uint32_t)((kk >> 1) & 1));
        tc_fence_after();
        uint32_t d0[32];
        tmem_ld32_issue(tlane + kColDT + b * 64 + h * 32, d0);
        tmem_ld_wait(d0);
        tc_fence_before();
        if (t >= 0) {
          float* dst = a.dtE + (int64_t)t * 64 + h * 32;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            red_add_v4(dst + 4 * j, __uint_as_float(d0[4 * j]), __uint_as_float(d0[4 * j + 1]), __uint_as_float(d0[4 * j + 2]),
                       __uint_as_float(d0[4 * j + 3]));
        }
      };

      int32_t t_prev = -1;
      load_targets(k, 0);
      load_targets(k, 1);
      for (int i = 0; i < n; ++i) {
        const int kk = k + i;
        const int b = (int)(kk & 1);
        const int nvalid = (int)(elig - (int64_t)(tile0 + i) * 128 < 128 ? elig - (int64_t)(tile0 + i) * 128 : 128);
        const int32_t t_cur = t_next;
        if (tid == 0) RTRACE(0);
        mbar_wait(&p_full[b], (uint32_t)((kk >> 1) & 1));
        if (tid == 0) RTRACE(1);
        tc_fence_after();
        uint4 ghi[8], glo[8];
#pragma unroll
        for (int j2 = 0; j2 < 2; ++j2) {
          uint32_t pv[32];
          tmem_ld32_issue(tlane + kColP + b * 128 + h * 64 + j2 * 32, pv);
          tmem_ld_wait(pv);
          float g[32];
          if (nvalid >= 128) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float dv = (__uint_as_float(pv[j]) + rbc - tv[j2 * 32 + j]) * cmask;
              loss = fmaf(dv, dv, loss);
              g[j] = dv * gscale;
              bsum += g[j];
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float dv = 0.f;
              if (col_ok && h * 64 + j2 * 32 + j < nvalid) dv = __uint_as_float(pv[j]) + rbc - tv[j2 * 32 + j];
              loss = fmaf(dv, dv, loss);
              g[j] = dv * gscale;
              bsum += g[j];
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            split8(make_float4(g[8 * j], g[8 * j + 1], g[8 * j + 2], g[8 * j + 3]),
                   make_float4(g[8 * j + 4], g[8 * j + 5], g[8 * j + 6], g[8 * j + 7]), ghi[j2 * 4 + j], glo[j2 * 4 + j]);
          // this half's target registers are free: the next unit's loads go out now and fly under the rest of this unit
          if (i + 1 < n) load_targets(kk + 1, j2);
        }
        tc_fence_before();
        if (tid == 0) RTRACE(2);
        if (i > 0) mbar_wait(&d_full[b ^ 1], (uint32_t)(((kk - 1) >> 1) & 1));      // previous unit's contractions done: sG is free
        if (tid == 0) RTRACE(3);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          sts16(sG + (h * 8 + j) * 2048 + c * 16, ghi[j]);
          sts16(sG + 32768 + (h * 8 + j) * 2048 + c * 16, glo[j]);
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&g_full);
        if (tid == 0) RTRACE(4);
        if (i > 0) scatter_dte(kk - 1, t_prev);
        if (tid == 0) RTRACE(5);
        t_prev = t_cur;
      }
      scatter_dte(k + n - 1, t_prev);
      // weight / bias gradient slice of this segment: TMEM lane = target column, this thread: features [32 h, 32 h + 32)
      {
        tc_fence_after();
        uint32_t v[32];
        tmem_ld32_issue(tlane + kColDW + h * 32, v);
        tmem_ld_wait(v);
        if (col_ok) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            red_add_v4(a.dRw + col * 64 + h * 32 + 4 * j, __uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                       __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
          atomicAdd(a.drb + col, bsum);
        }
      }
    }
    tc_fence_before();
    __syncthreads();       // segment drained: sW and the dRw accumulator are rewritten by the next one
    tc_fence_after();
    k += n;
    u0 = seg_end;
  }

  if (a.recon_out != nullptr && warp < 8) {
    for (int o = 16; o > 0; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
    if (lane == 0 && elig > 0 && loss != 0.f) atomicAdd(a.recon_out, loss * 100.0f / ((float)elig * (float)n_r));
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace

#ifdef MATCHA_RECON_TRACE
extern "C" int matcha_recon_trace(unsigned long long* out) {
  return cudaMemcpyFromSymbol(out, g_rtrace, sizeof(unsigned long long) * 4096) == cudaSuccess ? 0 : -2;
}
#endif

int64_t recon_pipe_tile_bytes(int64_t T) { return (num_token_tiles(T) + 1) * (int64_t)kPRec; }

// gradient pass of the reconstruction head (mode 1 of launch_recon_tc, same outputs): the loss is added to recon_out[0];
// UNSCALED dRw [n_r, 64] / drb [n_r] and dtE [T, 64] accumulate (zeroed by the caller).  tiles: recon_pipe_tile_bytes(T)
int launch_recon_pipe(const float* E, const int64_t* x, int64_t T, const float* inter, int64_t inter_ld, int64_t rs, int64_t re,
                      const float* Rw, const float* rb, const int32_t* counts, int rchrom, int n_chrom, const int32_t* perm,
                      const int32_t* group_off, float* recon_out, float* dRw, float* drb, float* dtE, uint8_t* tiles,
                      cudaStream_t s) {
  if (T <= 0 || re <= rs) return MATCHA_OK;
  static bool once = false;
  if (!once) {
    if (int rc = check_cuda(cudaFuncSetAttribute(recon_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPSmem),
                            "cudaFuncSetAttribute"))
      return rc;
    once = true;
  }
  const int64_t nt = num_token_tiles(T);
  recon_tiles_kernel<<<(unsigned)(nt < 4 * kSMs ? nt : 4 * kSMs), 256, 0, s>>>(E, x, perm, group_off, counts, rchrom, n_chrom, tiles);
  MATCHA_CHECK_LAUNCH("recon_tiles");
  PipeArgs a;
  a.tiles = tiles; a.inter = inter; a.inter_ld = inter_ld; a.rs = rs; a.re = re; a.Rw = Rw; a.rb = rb; a.counts = counts;
  a.group_off = group_off; a.rchrom = rchrom; a.n_chrom = n_chrom; a.ncb = (int)((re - rs + 127) / 128);
  a.recon_out = recon_out; a.dRw = dRw; a.drb = drb; a.dtE = dtE;
  const int64_t max_units = nt * a.ncb;
  recon_pipe_kernel<<<(unsigned)(max_units < kSMs ? max_units : kSMs), kPThreads, kPSmem, s>>>(a);
  MATCHA_CHECK_LAUNCH("recon_pipe");
  return MATCHA_OK;
}

}  // namespace matcha
