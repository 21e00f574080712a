// Fused reconstruction head (Modules.py:192-199 and its backward) on tensor cores:
//   pred = tanh(E) . Rw^T + rb  ->  diff vs the z-scored inter-chromosomal target rows  ->  loss,
//   and in the backward pass also  dRw += beta * gdiff^T . tanh(E),  drb += beta * sum_t gdiff,  dtE = gdiff . Rw
// in ONE kernel per pass: pred / gdiff [T, n_r] never touch HBM (the decomposed path writes and re-reads them three
// times through four SIMT launches).
//
// One CTA = 256 threads = 128 ELIGIBLE tokens x one block of 128 target columns: thread (r, h) owns token row r (= TMEM
// lane r) and half h of the features / of the target columns.  Eligible = real tokens outside the drawn chromosome; they
// are taken from the chromosome-bucketed token list the encoder already built (every bucket but the drawn one and the pad
// bucket), so the ~1/3 of the rows that are padding or in-chromosome never enter a tile.  Per token tile:
//   stage   tanh(E) rows -> bf16 hi | lo tile sA [feature/8][token][8]  (+ a ones column: plane 8, zeros: plane 9)
//   MMA1    P[128 tok, 128 col] = sA . sW^T                      (K = 64; sW = Rw block, K-major)
//   SIMT    thread = token: gdiff row = (P + rb - target) * gscale on eligible tokens; loss partial; -> sG tile
//   MMA2    dtE[128 tok, 64]   = sG . sW          (K = 128 columns; sW read MN-major)
//   MMA3    dRw[128 col, 80]  += sG^T . sA        (K = 128 tokens; both MN-major; column 64 = ones -> bias gradient)
// dRw stays resident in TMEM across all tiles of the CTA.  bf16x3 split, fp32 accumulation (tc_common.cuh).
#include "rowwise.cuh"
#include "tc_common.cuh"

namespace matcha {
namespace {

constexpr int kRThreads = 256;                // two warpgroups: thread = (token row r = tid & 127, half h = tid >> 7)
constexpr int kRA = 2 * 10 * 2048;            // sA: hi 10 planes | lo 10 planes                      40 960
constexpr int kRAHalf = 10 * 2048;
constexpr int kRW = 32768;                    // sW: 128 rows x 64 k, hi 16 KB | lo 16 KB
constexpr int kRG = 65536;                    // sG: 128 tokens x 128 columns, hi 32 KB | lo 32 KB
constexpr int kRStageRow = 36;                // staging row: 32 floats + pad (coalesced half-row I/O, 32 x 32 target blocks)
constexpr int kRStage = 8 * 32 * kRStageRow * 4;   // 36 864
constexpr int kRSmem = kRA + kRW + kRG + kRStage;  // 176 128 (gradient pass); 110 592 without sG (loss pass: 2 CTAs / SM)
constexpr uint32_t kColP = 0, kColDT = 128, kColDW = 192;

struct ReconArgs {
  const float* E; const int64_t* x; int64_t T;
  const float* inter; int64_t inter_ld;
  int64_t rs, re;                 // node-id range of the drawn chromosome
  const float* Rw; const float* rb;
  const int32_t* counts; int rchrom, n_chrom;
  const int32_t* perm; const int32_t* group_off;      // chromosome-bucketed token list (rowwise.cu: launch_bucket)
  float* recon_out;               // (optional) += 100 * mean((pred - target)^2)
  float* dRw; float* drb; float* dtE; float beta;     // mode 1
  int mode;
};

// 64-bit pointer broadcast from lane `src`
__device__ __forceinline__ const float* bcast_ptr(const float* p, int src) {
  const unsigned long long v = reinterpret_cast<unsigned long long>(p);
  const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src), hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src);
  return reinterpret_cast<const float*>(((unsigned long long)hi << 32) | lo);
}
// 32 scattered rows x 32 floats (row pointer incl. column offset held by the owning lane, NULL = zero row) -> registers:
// four rows per instruction, 8 lanes x float4 each, transposed through the warp's staging area
__device__ __forceinline__ void half_rows_gather(const float* rowp, float* stage, int lane, float (&v)[32]) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int row = 4 * k + (lane >> 3), c4 = lane & 7;
    const float* p = bcast_ptr(rowp, row);
    float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p != nullptr) e = __ldg(reinterpret_cast<const float4*>(p) + c4);
    *reinterpret_cast<float4*>(stage + row * kRStageRow + c4 * 4) = e;
  }
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float4 e = *reinterpret_cast<const float4*>(stage + lane * kRStageRow + k * 4);
    v[4 * k] = e.x; v[4 * k + 1] = e.y; v[4 * k + 2] = e.z; v[4 * k + 3] = e.w;
  }
  __syncwarp();
}
// half rows -> scattered global rows with atomic adds (several column blocks contribute to the same dtE row); one row of
// 32 consecutive floats per instruction
__device__ __forceinline__ void half_rows_scatter_red(float* rowp, float* stage, int lane, const float (&v)[32]) {
#pragma unroll
  for (int k = 0; k < 8; ++k)
    *reinterpret_cast<float4*>(stage + lane * kRStageRow + k * 4) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
  __syncwarp();
#pragma unroll 4
  for (int row = 0; row < 32; ++row) {
    float* p = const_cast<float*>(bcast_ptr(rowp, row));
    if (p != nullptr) atomicAdd(p + lane, stage[row * kRStageRow + lane]);
  }
  __syncwarp();
}
// 32 floats of row r -> planes p0 .. p0 + 3 of a K-major bf16 hi | lo tile
__device__ __forceinline__ void put_planes4(uint8_t* hi_base, int lo_off, int p0, int r, const float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 hi, lo;
    split8(make_float4(v[8 * j], v[8 * j + 1], v[8 * j + 2], v[8 * j + 3]),
           make_float4(v[8 * j + 4], v[8 * j + 5], v[8 * j + 6], v[8 * j + 7]), hi, lo);
    sts16(hi_base + (p0 + j) * 2048 + r * 16, hi);
    sts16(hi_base + lo_off + (p0 + j) * 2048 + r * 16, lo);
  }
}

__global__ void __launch_bounds__(kRThreads, 2) recon_tc_kernel(const ReconArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sW = smem + kRA;
  // loss-only pass (mode 0): no gdiff tile, the staging area follows sW directly (108 KB per CTA)
  uint8_t* sG = smem + kRA + kRW;
  float* sStage = reinterpret_cast<float*>(smem + kRA + kRW + (a.mode == 1 ? kRG : 0));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_rb[128];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r = tid & 127, h = tid >> 7, wq = warp & 3;         // token row, half, TMEM lane quarter
  const int cb = blockIdx.y;                                   // column block: target columns [cb * 128, cb * 128 + 128)
  const int64_t n_r = a.re - a.rs;
  const int S = gridDim.x;
  // eligible tokens = bucketed list minus the drawn chromosome's bucket: list positions [0, lo_n) and [lo_n + skip, ...)
  const int64_t lo_n = a.group_off[a.rchrom], skip = a.counts[a.rchrom];
  const int64_t elig = (int64_t)a.group_off[a.n_chrom] - skip;
  const int64_t ntiles = (elig + 127) / 128;
  const float gscale = elig > 0 ? 200.0f / ((float)elig * (float)n_r) : 0.f;

  const uint32_t tmem_cols = a.mode == 1 ? 512u : 128u;      // loss-only pass: P alone -> two CTAs per SM
  if (warp == 0) tmem_alloc(&tmem_base_s, tmem_cols);
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {   // Rw block -> sW (row = target column; this thread: features [32 h, 32 h + 32)), zero rows beyond n_r; bias slice
    const int64_t col = (int64_t)cb * 128 + r;
    float v[32];
    if (col < n_r) {
      const float* src = a.Rw + col * 64 + h * 32;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float4 e = __ldg(reinterpret_cast<const float4*>(src) + k);
        v[4 * k] = e.x; v[4 * k + 1] = e.y; v[4 * k + 2] = e.z; v[4 * k + 3] = e.w;
      }
    } else {
#pragma unroll
      for (int c = 0; c < 32; ++c) v[c] = 0.f;
    }
    put_planes4(sW, 16384, h * 4, r, v);
    if (h == 0) {
      s_rb[r] = col < n_r ? __ldg(a.rb + col) : 0.f;
      // ones / zero planes of sA never change
      sts16(sA + 8 * 2048 + r * 16, make_uint4(0x00003F80u, 0u, 0u, 0u));          // bf16(1.0) in column 64
      sts16(sA + 9 * 2048 + r * 16, make_uint4(0u, 0u, 0u, 0u));
      sts16(sA + kRAHalf + 8 * 2048 + r * 16, make_uint4(0u, 0u, 0u, 0u));
      sts16(sA + kRAHalf + 9 * 2048 + r * 16, make_uint4(0u, 0u, 0u, 0u));
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t tlane = tmem_base + ((uint32_t)(wq * 32) << 16);
  constexpr uint32_t idescP = make_idesc(128, 128, false, false);
  constexpr uint32_t idescD = make_idesc(128, 64, false, true);
  constexpr uint32_t idescW = make_idesc(128, 80, true, true);
  const uint32_t ah = smem_u32(sA), al = ah + kRAHalf;
  const uint32_t wh = smem_u32(sW), wl = wh + 16384;
  const uint32_t gh = smem_u32(sG), gl = gh + 32768;
  float* stage = sStage + warp * (32 * kRStageRow);
  uint32_t phase = 0;
  float loss = 0.f;
  bool first = true;

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += S) {
    const int64_t e = tile * 128 + wq * 32 + lane;           // position in the eligible list
    const bool ok = e < elig;
    const int64_t t = ok ? a.perm[e < lo_n ? e : e + skip] : 0;
    const int64_t id = ok ? a.x[t] : 0;
    // ---- stage tanh(E): this thread's 32 features ----
    {
      float v[32];
      half_rows_gather(ok ? a.E + t * 64 + h * 32 : nullptr, stage, lane, v);
#pragma unroll
      for (int c = 0; c < 32; ++c) v[c] = tanhf(v[c]);
      put_planes4(sA, kRAHalf, h * 4, r, v);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        umma_x3s(tmem_base + kColP, ah + ks * 4096, al + ks * 4096, wh + ks * 4096, wl + ks * 4096, 2048, 128, 2048, 128, idescP,
                 ks == 0);
      umma_commit(&bar);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- gdiff rows: thread = (token, half): 2 chunks of 32 target columns ----
#pragma unroll 1
    for (int c2 = 0; c2 < 2; ++c2) {
      const int ch = h * 2 + c2;                               // chunk of the 128-column block
      const int64_t c0 = (int64_t)cb * 128 + ch * 32;          // first target column of the chunk
      // target block [32 tokens x 32 columns]: row k of the warp is read by all lanes (coalesced), kept transposed
      {
        float tv[32];
        const float* tbase = a.inter + (a.rs - 1) + c0 + lane;
        const bool col_ok = c0 + lane < n_r;
#pragma unroll
        for (int k = 0; k < 32; ++k) {                         // 32 independent loads in flight per lane
          const int64_t idk = __shfl_sync(0xffffffffu, id, k);
          const bool okk = __shfl_sync(0xffffffffu, ok ? 1 : 0, k) != 0;
          tv[k] = (okk && col_ok) ? __ldg(tbase + (idk - 1) * a.inter_ld) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 32; ++k) stage[k * kRStageRow + lane] = tv[k];
      }
      __syncwarp();
      uint32_t pv[32];
      tmem_ld32_issue(tlane + kColP + ch * 32, pv);
      tmem_ld_wait(pv);
      float g[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float dv = 0.f;
        if (ok && c0 + i < n_r) dv = __uint_as_float(pv[i]) + s_rb[ch * 32 + i] - stage[lane * kRStageRow + i];
        loss = fmaf(dv, dv, loss);
        g[i] = dv * gscale;
      }
      __syncwarp();
      if (a.mode == 1) put_planes4(sG, 32768, ch * 4, r, g);
    }
    tc_fence_before();
    if (a.mode == 1) {
      fence_async_smem();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)      // dtE[128 tok, 64] = gdiff[128 tok, 128 col] . Rw[128 col, 64]   (B MN-major)
          umma_x3s(tmem_base + kColDT, gh + ks * 4096, gl + ks * 4096, wh + ks * 256, wl + ks * 256, 2048, 128, 128, 2048, idescD,
                   ks == 0);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)      // dRw[128 col, 80] += gdiff^T[128 col, 128 tok] . [tanh(E) | 1][128 tok, 80]
          umma_x3s(tmem_base + kColDW, gh + ks * 256, gl + ks * 256, ah + ks * 256, al + ks * 256, 128, 2048, 128, 2048, idescW,
                   first && ks == 0);
        umma_commit(&bar);
      }
      first = false;
      mbar_wait(&bar, phase);
      phase ^= 1;
      tc_fence_after();
      {
        uint32_t d0[32];
        tmem_ld32_issue(tlane + kColDT + h * 32, d0);
        tmem_ld_wait(d0);
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(d0[i]);
        half_rows_scatter_red(ok ? a.dtE + t * 64 + h * 32 : nullptr, stage, lane, v);
      }
      tc_fence_before();
    }
    __syncthreads();       // sA / sG / staging are rewritten by the next tile
  }

  if (a.recon_out != nullptr) {
    for (int o = 16; o > 0; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
    if (lane == 0 && elig > 0 && loss != 0.f) atomicAdd(a.recon_out, loss * 100.0f / ((float)elig * (float)n_r));
  }
  if (a.mode == 1 && !first) {
    // weight / bias gradient slice of this CTA: TMEM lane = target column, this thread: features [32 h, 32 h + 32)
    tc_fence_after();
    const int64_t col = (int64_t)cb * 128 + r;
    uint32_t v[32];
    tmem_ld32_issue(tlane + kColDW + h * 32, v);
    tmem_ld_wait(v);
    if (col < n_r) {
#pragma unroll
      for (int i = 0; i < 32; ++i) atomicAdd(a.dRw + col * 64 + h * 32 + i, __uint_as_float(v[i]) * a.beta);
    }
    if (h == 1) {
      uint32_t b8[8];
      tmem_ld8_issue(tlane + kColDW + 64, b8);
      tmem_ld_wait(b8);
      if (col < n_r) atomicAdd(a.drb + col, __uint_as_float(b8[0]) * a.beta);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

// out[i] += scale * in[i]  (the training forward leaves the unscaled weight / bias gradients in the workspace; the backward
// pass applies d loss / d recon)
__global__ void axpy_kernel(const float* __restrict__ in, float scale, float* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] += scale * in[i];
}
}  // namespace
int launch_axpy(const float* in, float scale, float* out, int64_t n, cudaStream_t s) {
  if (n <= 0) return MATCHA_OK;
  int64_t blocks = (n + 255) / 256;
  if (blocks > kSMs * 4) blocks = kSMs * 4;
  axpy_kernel<<<(unsigned)blocks, 256, 0, s>>>(in, scale, out, n);
  MATCHA_CHECK_LAUNCH("axpy");
  return MATCHA_OK;
}
// mode 0: loss only (eval);  mode 1: also dRw, drb += beta * ..., dtE += gdiff . Rw (dtE zeroed by the caller).  recon_out
// (optional) receives the loss in either mode
int launch_recon_tc(const float* E, const int64_t* x, int64_t T, const float* inter, int64_t inter_ld, int64_t rs, int64_t re,
                    const float* Rw, const float* rb, const int32_t* counts, int rchrom, int n_chrom, const int32_t* perm,
                    const int32_t* group_off, float* recon_out, float* dRw, float* drb, float* dtE, float beta, int mode,
                    cudaStream_t s) {
  if (T <= 0 || re <= rs) return MATCHA_OK;
  static bool once = false;
  if (!once) {
    if (int rc = check_cuda(cudaFuncSetAttribute(recon_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRSmem),
                            "cudaFuncSetAttribute"))
      return rc;
    once = true;
  }
  ReconArgs a;
  a.E = E; a.x = x; a.T = T; a.inter = inter; a.inter_ld = inter_ld; a.rs = rs; a.re = re; a.Rw = Rw; a.rb = rb;
  a.counts = counts; a.rchrom = rchrom; a.n_chrom = n_chrom; a.perm = perm; a.group_off = group_off; a.recon_out = recon_out; a.dRw = dRw; a.drb = drb; a.dtE = dtE;
  a.beta = beta; a.mode = mode;
  const int ncb = (int)((re - rs + 127) / 128);
  const int64_t ntiles = (T + 127) / 128;
  const int per_sm = mode == 1 ? 1 : 2;                      // 172 KB (gradient pass) or 108 KB (loss pass) of shared memory per CTA
  int64_t S = (per_sm * kSMs + ncb - 1) / ncb;
  if (S > ntiles) S = ntiles;
  if (S < 1) S = 1;
  recon_tc_kernel<<<dim3((unsigned)S, (unsigned)ncb), kRThreads, mode == 1 ? kRSmem : kRSmem - kRG, s>>>(a);
  MATCHA_CHECK_LAUNCH("recon_tc");
  return MATCHA_OK;
}

}  // namespace matcha
