// Node encoder forward on tensor cores (Modules.py:104-122 via :176-188; embed_dim 64, dense feature rows):
//   H0 = tanh( dropout(F_c[id - start_c]) . W0_c^T ),   E = H0 . W1_c^T        for the tokens of chromosome c
// as ONE kernel over the chromosome-bucketed token list: a CTA owns 128 tokens of one chromosome (thread r = token
// row r = TMEM lane r), walks the feature row in chunks of 64 columns { gather rows (coalesced, through a staging
// area) -> dropout from the counter RNG -> bf16 hi | lo A tile -> tcgen05 against the pre-split W0 chunk }, applies
// tanh to the accumulator row, stores H0 (the backward pass needs it), and chains the 64 x 64 contraction with W1_c.
// Replaces two grouped SIMT launches and the H0 re-read between them.
#include <stdlib.h>
#include <string.h>

#include "rowwise.cuh"
#include "tc_common.cuh"

namespace matcha {
#ifdef MATCHA_ENC_TRACE
__device__ unsigned long long g_etrace[4096];
// first 250 steps of CTA 0: 16 event slots per step
#define ETRACE(ev) do { if (blockIdx.x == 0 && sidx < 250u) { unsigned long long _t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t)); g_etrace[sidx * 16 + (ev)] = _t; } } while (0)
#define BTRACE(ev) do { if (blockIdx.x == 0 && iidx < 250u) { unsigned long long _t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t)); g_etrace[iidx * 16 + (ev)] = _t; } } while (0)
#else
#define ETRACE(ev) do { } while (0)
#define BTRACE(ev) do { } while (0)
#endif
namespace {

constexpr int kFThreads = 256;                 // forward kernel: two warpgroups split every 64-wide row
constexpr int kHStageRow = 36;                 // half-row staging: 32 floats + pad
constexpr int kEChunk = 16384;                 // one 64 x 64 weight chunk: K-major [k/8][row][8], hi 8 KB | lo 8 KB
constexpr int kEA = 32768;                     // A tile: 128 tokens x 64 k, hi 16 KB | lo 16 KB
constexpr int kEStage = 8 * 32 * kHStageRow * 4;     // 36 864
constexpr int kESmem = kEA + 2 * kEChunk + kEStage;  // 102 400 -> two CTAs per SM

struct EncMeta {
  int32_t n;
  int32_t nc[MATCHA_MAX_CHROM];                // bins per chromosome
  int64_t start[MATCHA_MAX_CHROM];             // first node id
  int64_t ld[MATCHA_MAX_CHROM];                // feature row stride (floats, multiple of 4, zero padded)
  const float* feat[MATCHA_MAX_CHROM];
  int64_t off_w0[MATCHA_MAX_CHROM], off_w1[MATCHA_MAX_CHROM];   // element offsets into params
  int64_t woff[MATCHA_MAX_CHROM];              // byte offset of the chromosome's pre-split chunks (W0 chunks..., W1)
};

// W0_c [64, n_c] -> ceil(n_c / 64) chunks of [k/8][f][8] bf16 hi | lo (zero beyond n_c); W1_c [64, 64] -> one chunk
__global__ void enc_split_kernel(const EncMeta em, const float* __restrict__ params, uint8_t* __restrict__ out, int64_t units) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // unit = (chunk slot, row f, k-group g)
  if (u >= units) return;
  const int g = (int)(u & 7), f = (int)((u >> 3) & 63);
  int64_t slot = u >> 9;
  int c = 0;
  for (; c < em.n; ++c) {
    const int64_t ns = (em.nc[c] + 63) / 64 + 1;
    if (slot < ns) break;
    slot -= ns;
  }
  if (c >= em.n) return;
  const int nchunk = (em.nc[c] + 63) / 64;
  float v[8];
  if (slot < nchunk) {
    const float* src = params + em.off_w0[c] + (int64_t)f * em.nc[c];
    const int k0 = (int)slot * 64 + g * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (k0 + i < em.nc[c]) ? __ldg(src + k0 + i) : 0.f;
  } else {
    const float* src = params + em.off_w1[c] + (int64_t)f * 64 + g * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __ldg(src + i);
  }
  uint4 hi, lo;
  split8(make_float4(v[0], v[1], v[2], v[3]), make_float4(v[4], v[5], v[6], v[7]), hi, lo);
  uint8_t* dst = out + em.woff[c] + slot * kEChunk + g * 1024 + f * 16;
  *reinterpret_cast<uint4*>(dst) = hi;
  *reinterpret_cast<uint4*>(dst + 8192) = lo;
}

// 64-bit pointer broadcast from lane `src`
__device__ __forceinline__ const float* shfl_ptr(const float* p, int src) {
  const unsigned long long v = reinterpret_cast<unsigned long long>(p);
  const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src), hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src);
  return reinterpret_cast<const float*>(((unsigned long long)hi << 32) | lo);
}
__device__ __forceinline__ float* shfl_ptr(float* p, int src) {
  return const_cast<float*>(shfl_ptr(const_cast<const float*>(p), src));
}


// 32 scattered rows x 32 floats (row pointer incl. column offset held by the owning lane, NULL = zero row): four rows per
// instruction, 8 lanes x float4 each; transposed through the warp's staging area
__device__ __forceinline__ void gather_half_rows(const float* rowp, int64_t k_first, int64_t klimit, float* stage, int lane,
                                                 float (&v)[32]) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int row = 4 * k + (lane >> 3), c4 = lane & 7;
    const float* p = shfl_ptr(rowp, row);
    float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p != nullptr && k_first + c4 * 4 < klimit) e = __ldg(reinterpret_cast<const float4*>(p) + c4);
    *reinterpret_cast<float4*>(stage + row * kHStageRow + c4 * 4) = e;
  }
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float4 e = *reinterpret_cast<const float4*>(stage + lane * kHStageRow + k * 4);
    v[4 * k] = e.x; v[4 * k + 1] = e.y; v[4 * k + 2] = e.z; v[4 * k + 3] = e.w;
  }
  __syncwarp();
}
__device__ __forceinline__ void scatter_half_rows(float* rowp, float* stage, int lane, const float (&v)[32]) {
#pragma unroll
  for (int k = 0; k < 8; ++k)
    *reinterpret_cast<float4*>(stage + lane * kHStageRow + k * 4) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int row = 4 * k + (lane >> 3), c4 = lane & 7;
    float* p = shfl_ptr(rowp, row);
    if (p != nullptr) reinterpret_cast<float4*>(p)[c4] = *reinterpret_cast<const float4*>(stage + row * kHStageRow + c4 * 4);
  }
  __syncwarp();
}
__device__ __forceinline__ void put4(uint8_t* hi_base, int lo_off, int p0, int r, const float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 hi, lo;
    split8(make_float4(v[8 * j], v[8 * j + 1], v[8 * j + 2], v[8 * j + 3]),
           make_float4(v[8 * j + 4], v[8 * j + 5], v[8 * j + 6], v[8 * j + 7]), hi, lo);
    sts16(hi_base + (p0 + j) * 2048 + r * 16, hi);
    sts16(hi_base + lo_off + (p0 + j) * 2048 + r * 16, lo);
  }
}

// One CTA = 256 threads: thread (r, h) owns token row r (TMEM lane r) and half h of every 64-wide row / chunk
__global__ void __launch_bounds__(kFThreads, 2)
enc_tc_fwd_kernel(const EncMeta em, const uint8_t* __restrict__ wsplit, const int64_t* __restrict__ x,
                  const int32_t* __restrict__ perm, const int32_t* __restrict__ group_off, float* __restrict__ H0,
                  float* __restrict__ E, const DropCfg drop) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sW = smem + kEA;
  uint8_t* sW1 = smem + kEA + kEChunk;
  float* sStage = reinterpret_cast<float*>(smem + kEA + 2 * kEChunk);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r = tid & 127, h = tid >> 7, wq = warp & 3;
  if (warp == 0) tmem_alloc(&tmem_base_s, 64);
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t tlane = tmem_base + ((uint32_t)(wq * 32) << 16);
  constexpr uint32_t idesc = make_idesc(128, 64, false, false);
  const uint32_t ah = smem_u32(sA), al = ah + 16384;
  const uint32_t wh = smem_u32(sW), wl = wh + 8192;
  const uint32_t vh = smem_u32(sW1), vl = vh + 8192;
  float* stage = sStage + warp * (32 * kHStageRow);
  uint32_t phase = 0;
  int cur_c = -1;
  auto copy_chunk = [&](uint8_t* dst, const uint8_t* src) {     // 16 KB, all 256 threads
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int i = 0; i < kEChunk / 16 / kFThreads; ++i) d4[tid + i * kFThreads] = __ldg(s4 + tid + i * kFThreads);
  };

  // tile list: chromosome c contributes ceil(count_c / 128) tiles of its bucket (pads live in bucket em.n: skipped)
  int c = 0;
  int64_t before = 0;                       // tiles of the chromosomes before c
  for (int64_t ti = blockIdx.x;; ti += gridDim.x) {
    int cnt = 0;
    for (; c < em.n; ++c) {
      cnt = group_off[c + 1] - group_off[c];
      const int64_t nt = (cnt + 127) / 128;
      if (ti < before + nt) break;
      before += nt;
    }
    if (c >= em.n) break;
    const int off = (int)(ti - before) * 128;
    const int nrows_cta = cnt - off < 128 ? cnt - off : 128;
    const bool live = r < nrows_cta;
    const int64_t t = live ? perm[group_off[c] + off + r] : 0;
    const float* frow = live ? em.feat[c] + (x[t] - em.start[c]) * em.ld[c] : nullptr;
    const int64_t ld = em.ld[c];
    const int nchunk = (em.nc[c] + 63) / 64;
    const uint8_t* wbase = wsplit + em.woff[c];
    if (c != cur_c) {                       // W1_c stays resident while the CTA works on this chromosome
      __syncthreads();
      copy_chunk(sW1, wbase + (int64_t)nchunk * kEChunk);
      cur_c = c;
    }
    // feature chunk kc + 1 is gathered (global loads, dropout) while the tensor pipe works on chunk kc
    auto gather_chunk = [&](int kc, float (&v)[32]) {
      const int64_t kf = (int64_t)kc * 64 + h * 32;
      gather_half_rows(frow ? frow + kf : nullptr, kf, ld, stage, lane, v);
      if (drop.thr != 0u && live) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 d = drop_apply4(drop, (uint64_t)t, (uint32_t)(kf + 4 * j), make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
          v[4 * j] = d.x; v[4 * j + 1] = d.y; v[4 * j + 2] = d.z; v[4 * j + 3] = d.w;
        }
      }
    };
    float fv[32];
    gather_chunk(0, fv);
    for (int kc = 0; kc < nchunk; ++kc) {
      put4(sA, 16384, h * 4, r, fv);
      copy_chunk(sW, wbase + (int64_t)kc * kEChunk);
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_x3s(tmem_base, ah + ks * 4096, al + ks * 4096, wh + ks * 2048, wl + ks * 2048, 2048, 128, 1024, 128, idesc,
                   kc == 0 && ks == 0);
        umma_commit(&bar);
      }
      if (kc + 1 < nchunk) gather_chunk(kc + 1, fv);
      mbar_wait(&bar, phase);               // the A tile and the weight chunk are rewritten by the next chunk
      phase ^= 1;
      tc_fence_after();
    }
    // ---- H0 = tanh(acc): kept for the backward pass, and the A operand of the second contraction ----
    float hv[32];
    {
      uint32_t d0[32];
      tmem_ld32_issue(tlane + h * 32, d0);
      tmem_ld_wait(d0);
#pragma unroll
      for (int i = 0; i < 32; ++i) hv[i] = live ? tanhf(__uint_as_float(d0[i])) : 0.f;
    }
    tc_fence_before();
    put4(sA, 16384, h * 4, r, hv);
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        umma_x3s(tmem_base, ah + ks * 4096, al + ks * 4096, vh + ks * 2048, vl + ks * 2048, 2048, 128, 1024, 128, idesc, ks == 0);
      umma_commit(&bar);
    }
    scatter_half_rows(live ? H0 + t * 64 + h * 32 : nullptr, stage, lane, hv);      // under the second contraction
    mbar_wait(&bar, phase);
    phase ^= 1;
    tc_fence_after();
    {
      uint32_t d0[32];
      tmem_ld32_issue(tlane + h * 32, d0);
      tmem_ld_wait(d0);
#pragma unroll
      for (int i = 0; i < 32; ++i) hv[i] = __uint_as_float(d0[i]);
    }
    tc_fence_before();
    scatter_half_rows(live ? E + t * 64 + h * 32 : nullptr, stage, lane, hv);
    __syncthreads();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 64);
}

// ==========================================================================================
// Pipelined forward (default; enc_tc_fwd_kernel above stays as the cross-check, MATCHA_ENC_PIPE=0).
// The unit kernel above runs  gather -> split -> copy weights -> barrier -> MMA  once per 64-column chunk with the global
// loads of one chunk in flight per CTA (29 % of the HBM copy rate at cfg3).  Here the feature rows of a chunk arrive by
// asynchronous copies (cp.async, 16-byte pieces, no register staging) issued by four producer warps into a ring of four fp32
// staging tiles together with one bulk copy of the chunk's pre-split weights -- so up to three chunks of feature rows are in
// flight per SM whatever the compute warps do;
// sixteen compute warps turn a landed stage into the bf16 hi | lo A operand IN TENSOR MEMORY (thread = TMEM lane = token row:
// tcgen05.st of the packed bf16 pairs, column c = k elements 2c, 2c + 1; A-from-TMEM MMAs -- no operand tile in shared memory,
// no operand reads on the shared-memory pipe: 172 -> 157 us at cfg3; dropout from the counter RNG; two buffers so the
// conversion of chunk k + 1 runs under the MMAs of chunk k; with eight warps this conversion was the critical path: 1.2 us
// per chunk, scripts/dev/enc_trace.py); one thread issues the MMAs.  One persistent CTA per SM.
//   step = (tile, chunk kc) for kc < nchunk, then (tile, W1): H0 = tanh(acc) -> A tile -> E = H0 . W1_c^T
// ==========================================================================================
// the converted A operand lives in TMEM (tcgen05.st by the converter threads, A-from-TMEM MMAs: tc_common.cuh), not in shared memory
constexpr uint32_t kPFColA = 128;               // two operand buffers of 64 columns: hi 32 | lo 32 (bf16 pairs)
constexpr int kPFThreads = 672;                 // warps 0-15 compute, warps 16-19 producers (32 rows each), warp 20 MMA issuer
constexpr int kPFRow = 272;                     // staging row stride: 64 floats + 16 B (conflict-free 16-byte reads down a column)
constexpr int kPFStageF = 128 * kPFRow;         // 34 816
constexpr int kPFStage = kPFStageF + kEChunk;   // + the chunk's pre-split weights: 51 200
constexpr int kPFStages = 4;
constexpr int kPFSmem = kPFStages * kPFStage;                // 204 800 (the operand tiles are in TMEM)
constexpr uint32_t kPFColAcc = 0, kPFColE = 64;

__global__ void __launch_bounds__(kPFThreads, 1)
enc_pipe_fwd_kernel(const EncMeta em, const uint8_t* __restrict__ wsplit, const int64_t* __restrict__ x,
                    const int32_t* __restrict__ perm, const int32_t* __restrict__ group_off, float* __restrict__ H0,
                    float* __restrict__ E, const DropCfg drop) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sStage = smem;
  __shared__ uint64_t f_full[kPFStages], f_free[kPFStages], a_full[2], a_free[2], acc_full, e_full;
  __shared__ uint32_t tmem_base_s;
  __shared__ const float* sRow[128];              // feature-row pointers of the producer's current tile
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) tmem_alloc(&tmem_base_s, 256);
  if (tid == 512) {
    for (int i = 0; i < kPFStages; ++i) { mbar_init(&f_full[i], 129); mbar_init(&f_free[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], 16); mbar_init(&a_free[i], 1); }
    mbar_init(&acc_full, 1); mbar_init(&e_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  // every role walks the same tile list: chromosome c contributes ceil(count_c / 128) tiles of its bucket
  struct Cursor { int c; int64_t before; };
  auto next_tile = [&](Cursor& cu, int64_t ti, int& cnt) -> bool {
    for (; cu.c < em.n; ++cu.c) {
      cnt = group_off[cu.c + 1] - group_off[cu.c];
      const int64_t nt = (cnt + 127) / 128;
      if (ti < cu.before + nt) return true;
      cu.before += nt;
    }
    return false;
  };

  if (warp >= 16 && warp < 20) {
    // ---------------- producers: feature rows (32 per warp) + weight chunk of every step ----------------
    // rows: 16-byte cp.async pieces (an instruction moves two 256-byte row chunks; pieces of dead rows / beyond the row are
    // zero-filled), weights: one 16 KB bulk copy.  f_full counts the 128 lanes' cp.async arrivals + the bulk copy's expect_tx.
    // (one producer warp for all 128 rows: 4.0 us per chunk -- a single warp does not issue LDGSTS fast enough)
    // (one bulk copy per ROW was measured first: ~50 ns per copy, 6.7 us per chunk -- the copy engine is not a gather unit)
    Cursor cu{0, 0};
    uint32_t sidx = 0;
    const int half = lane >> 4, piece = lane & 15, pw = warp - 16;
    const float** myRow = sRow + pw * 32;
    for (int64_t ti = blockIdx.x;; ti += gridDim.x) {
      int cnt = 0;
      if (!next_tile(cu, ti, cnt)) break;
      const int c = cu.c;
      const int off = (int)(ti - cu.before) * 128;
      const int nrows = cnt - off < 128 ? cnt - off : 128;
      const int64_t ld = em.ld[c];
      const int nchunk = (em.nc[c] + 63) / 64;
      const uint8_t* wbase = wsplit + em.woff[c];
      __syncwarp();
      {
        const int r = pw * 32 + lane;
        const float* fr = em.feat[c];
        if (r < nrows) fr += (x[perm[group_off[c] + off + r]] - em.start[c]) * ld;
        myRow[lane] = fr;
      }
      __syncwarp();
      for (int kc = 0; kc <= nchunk; ++kc, ++sidx) {
        const int st = (int)(sidx % kPFStages);
        uint8_t* stage = sStage + st * kPFStage;
        if (pw == 0 && lane == 0) ETRACE(0);
        mbar_wait_backoff(&f_free[st], ((sidx / kPFStages) & 1u) ^ 1u);
        if (pw == 0 && lane == 0) ETRACE(1);
        if (pw == 0 && lane == 0) {
          mbar_expect_tx(&f_full[st], kEChunk);
          bulk_g2s(stage + kPFStageF, wbase + (int64_t)kc * kEChunk, kEChunk, &f_full[st]);     // kc == nchunk: W1_c
        }
        if (kc < nchunk) {
          const int64_t col = (int64_t)kc * 64 + piece * 4;
          const bool col_ok = col < ld;
          const uint32_t dst0 = smem_u32(stage) + (pw * 32 + half) * kPFRow + piece * 16;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int r = 2 * j + half;
            const float* src = myRow[r] + (col_ok ? col : 0);
            const uint32_t nbytes = (col_ok && pw * 32 + r < nrows) ? 16u : 0u;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst0 + 2 * j * kPFRow), "l"(src), "r"(nbytes) : "memory");
          }
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&f_full[st])) : "memory");
        if (pw == 0 && lane == 0) ETRACE(2);
      }
    }
  } else if (warp == 20) {
    // ---------------- MMA issuer ----------------
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(128, 64, false, false);
      Cursor cu{0, 0};
      uint32_t sidx = 0;
      for (int64_t ti = blockIdx.x;; ti += gridDim.x) {
        int cnt = 0;
        if (!next_tile(cu, ti, cnt)) break;
        const int nchunk = (em.nc[cu.c] + 63) / 64;
        for (int kc = 0; kc <= nchunk; ++kc, ++sidx) {
          const int st = (int)(sidx % kPFStages), ab = (int)(sidx & 1);
          mbar_wait_backoff(&f_full[st], (sidx / kPFStages) & 1u);
          ETRACE(8);
          mbar_wait_backoff(&a_full[ab], (sidx >> 1) & 1u);
          ETRACE(9);
          tc_fence_after();
          const uint32_t wh = smem_u32(sStage + st * kPFStage + kPFStageF), wl = wh + 8192;
          const uint32_t dcol = tmem_base + (kc < nchunk ? kPFColAcc : kPFColE);
          const uint32_t ta_hi = tmem_base + kPFColA + ab * 64, ta_lo = ta_hi + 32;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const bool fresh = (kc == 0 || kc == nchunk) && ks == 0;
            umma_ts_bf16(dcol, ta_lo + ks * 8, make_smem_desc(wh + ks * 2048, 1024, 128), idesc, fresh ? 0u : 1u);
            umma_ts_bf16(dcol, ta_hi + ks * 8, make_smem_desc(wl + ks * 2048, 1024, 128), idesc, 1u);
            umma_ts_bf16(dcol, ta_hi + ks * 8, make_smem_desc(wh + ks * 2048, 1024, 128), idesc, 1u);
          }
          umma_commit(&f_free[st]);
          umma_commit(&a_free[ab]);
          if (kc == nchunk - 1) umma_commit(&acc_full);
          if (kc == nchunk) umma_commit(&e_full);
          ETRACE(10);
        }
      }
    }
  } else {
    // ---------------- compute warps: thread (r, q4) = token row r (TMEM lane r), quarter q4 of every 64-wide row ----------------
    const int r = tid & 127, q4 = tid >> 7, wq = warp & 3;
    const uint32_t tlane = tmem_base + ((uint32_t)(wq * 32) << 16);
    Cursor cu{0, 0};
    uint32_t sidx = 0, tcount = 0;
    auto put_tile = [&](uint32_t si, const float (&v)[16]) {      // 16 floats of row r -> planes 2 q4, 2 q4 + 1 of operand tile si & 1
      const int ab = (int)(si & 1);
      mbar_wait(&a_free[ab], ((si >> 1) & 1u) ^ 1u);
      uint32_t ph[8], pl[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) split2(v[2 * j], v[2 * j + 1], ph[j], pl[j]);
      tc_fence_after();
      tmem_st8(tlane + kPFColA + ab * 64 + q4 * 8, ph);
      tmem_st8(tlane + kPFColA + ab * 64 + 32 + q4 * 8, pl);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[ab]);
    };
    for (int64_t ti = blockIdx.x;; ti += gridDim.x, ++tcount) {
      int cnt = 0;
      if (!next_tile(cu, ti, cnt)) break;
      const int c = cu.c;
      const int off = (int)(ti - cu.before) * 128;
      const int nrows = cnt - off < 128 ? cnt - off : 128;
      const bool live = r < nrows;
      const int64_t t = live ? perm[group_off[c] + off + r] : 0;
      const int nchunk = (em.nc[c] + 63) / 64;
      for (int kc = 0; kc < nchunk; ++kc, ++sidx) {
        const int st = (int)(sidx % kPFStages);
        if (tid == 0) ETRACE(4);
        mbar_wait(&f_full[st], (sidx / kPFStages) & 1u);
        if (tid == 0) ETRACE(5);
        const float* srow = reinterpret_cast<const float*>(sStage + st * kPFStage + r * kPFRow) + q4 * 16;
        const int64_t kf = (int64_t)kc * 64 + q4 * 16;
        float v[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float4 e = *reinterpret_cast<const float4*>(srow + 4 * j);
          if (drop.thr != 0u && live) e = drop_apply4(drop, (uint64_t)t, (uint32_t)(kf + 4 * j), e);      // dead rows / columns beyond the row arrive zero-filled
          v[4 * j] = e.x; v[4 * j + 1] = e.y; v[4 * j + 2] = e.z; v[4 * j + 3] = e.w;
        }
        put_tile(sidx, v);
        if (tid == 0) ETRACE(6);
      }
      // ---- H0 = tanh(acc): kept for the backward pass, and the A operand of the W1 step ----
      {
        mbar_wait(&acc_full, tcount & 1u);
        tc_fence_after();
        float d0[16];
        tmem_ld16(tlane + kPFColAcc + q4 * 16, d0);
        tc_fence_before();
        float hv[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) hv[i] = live ? tanhf(d0[i]) : 0.f;
        put_tile(sidx, hv);
        if (live) {
          float4* dst = reinterpret_cast<float4*>(H0 + t * 64 + q4 * 16);
#pragma unroll
          for (int j = 0; j < 4; ++j) dst[j] = make_float4(hv[4 * j], hv[4 * j + 1], hv[4 * j + 2], hv[4 * j + 3]);
        }
        ++sidx;
      }
      {
        mbar_wait(&e_full, tcount & 1u);
        tc_fence_after();
        float d0[16];
        tmem_ld16(tlane + kPFColE + q4 * 16, d0);
        tc_fence_before();
        if (live) {
          float4* dst = reinterpret_cast<float4*>(E + t * 64 + q4 * 16);
#pragma unroll
          for (int j = 0; j < 4; ++j) dst[j] = make_float4(d0[4 * j], d0[4 * j + 1], d0[4 * j + 2], d0[4 * j + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// ==========================================================================================
// backward:  dH0pre = (dE . W1_c) * (1 - H0^2),   dW1_c += dE^T . H0,   dW0_c += dH0pre^T . dropout(F_c rows)
// One CTA = 256 threads: thread (r, h) owns token row r (TMEM lane r) and half h of every 64-wide row.  The two weight
// gradients share ONE M = 128 contraction per operand tile: the stacked tile sS = [dE | dH0pre] (128 feature rows, K =
// tokens, MN-major view) against sB = H0 and then each feature chunk; lanes 0..63 of the H0 columns are dW1_c, lanes
// 64..127 of the chunk columns are dW0_c (the other two blocks of the 2 x 2 product are ignored).  Accumulators stay in
// TMEM while the CTA stays inside one (chromosome, column group) and are flushed with atomics.  A column group is
// kBMaxChunks feature chunks (384 bins: what fits in TMEM next to dH0 and dW1); a chromosome wider than that contributes
// its token tiles once per column group (work item = (chromosome, group, tile), group major), re-deriving dH0pre per item
// -- 512 B per token next to the 1.5 KB of feature columns the item gathers.  CTAs own contiguous item ranges.
// ==========================================================================================
constexpr int kBThreads = 256;
constexpr int kBS = 65536;                      // sS: 16 planes, hi 32 KB | lo 32 KB
constexpr int kBB = 32768;                      // sB: 8 planes, hi 16 KB | lo 16 KB
constexpr int kBStage = 8 * 32 * kHStageRow * 4;            // 36 864
constexpr int kBSmem = kBS + kBB + kEChunk + kBStage;       // 151 552
constexpr int kBMaxChunks = 6;                  // TMEM: dH0 64 | dW1 64 | 6 x 64 chunk columns = 512
constexpr uint32_t kColDH = 0, kColW1 = 64, kColW0 = 128;

__global__ void __launch_bounds__(kBThreads, 1)
enc_tc_bwd_kernel(const EncMeta em, const uint8_t* __restrict__ wsplit, const int64_t* __restrict__ x,
                  const int32_t* __restrict__ perm, const int32_t* __restrict__ group_off, const float* __restrict__ dE,
                  const float* __restrict__ H0, float* __restrict__ grads, const DropCfg drop) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sS = smem;
  uint8_t* sB = smem + kBS;
  uint8_t* sW1 = smem + kBS + kBB;
  float* sStage = reinterpret_cast<float*>(smem + kBS + kBB + kEChunk);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r = tid & 127, h = tid >> 7, wq = warp & 3;
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t tlane = tmem_base + ((uint32_t)(wq * 32) << 16);
  constexpr uint32_t idescD = make_idesc(128, 64, false, true);     // dH0 = dE . W1          (B MN-major)
  constexpr uint32_t idescW = make_idesc(128, 64, true, true);      // [dE | dH0pre]^T . tile  (both MN-major)
  const uint32_t sh = smem_u32(sS), sl = sh + 32768;
  const uint32_t bh = smem_u32(sB), bl = bh + 16384;
  const uint32_t vh = smem_u32(sW1), vl = vh + 8192;
  float* stage = sStage + warp * (32 * kHStageRow);
  uint32_t phase = 0;

  // this CTA's contiguous range of the item list (chromosome c contributes groups_c * ceil(count_c / 128) items)
  auto n_groups = [&](int cc) { return ((em.nc[cc] + 63) / 64 + kBMaxChunks - 1) / kBMaxChunks; };
  int64_t total = 0;
  for (int c = 0; c < em.n; ++c) total += (int64_t)((group_off[c + 1] - group_off[c] + 127) / 128) * n_groups(c);
  const int64_t per = (total + gridDim.x - 1) / gridDim.x;
  const int64_t ti0 = (int64_t)blockIdx.x * per, ti1 = (ti0 + per < total) ? ti0 + per : total;

  auto flush = [&](int c, int kc0, int kc1) {   // TMEM -> atomics on the weight gradients of chromosome c, chunks [kc0, kc1)
    tc_fence_after();
    if (r < 64) {                          // lanes 0..63: dW1_c[o = r][k], this thread: k in [32 h, 32 h + 32)
      if (kc0 == 0) {                      // every column group accumulates dW1_c; the first one publishes it
        uint32_t v[32];
        tmem_ld32_issue(tlane + kColW1 + h * 32, v);
        tmem_ld_wait(v);
        float* dst = grads + em.off_w1[c] + (int64_t)r * 64 + h * 32;
#pragma unroll
        for (int i = 0; i < 32; ++i) atomicAdd(dst + i, __uint_as_float(v[i]));
      }
    } else {                               // lanes 64..127: dW0_c[f = r - 64][k]
      const int nc = em.nc[c];
      float* dst = grads + em.off_w0[c] + (int64_t)(r - 64) * nc;
      for (int kc = kc0; kc < kc1; ++kc) {
        uint32_t v[32];
        tmem_ld32_issue(tlane + kColW0 + (kc - kc0) * 64 + h * 32, v);
        tmem_ld_wait(v);
        const int k0 = kc * 64 + h * 32;
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (k0 + i < nc) atomicAdd(dst + k0 + i, __uint_as_float(v[i]));
      }
    }
    tc_fence_before();
    __syncthreads();
  };

  int c = 0, cur_c = -1, cur_kc0 = 0, cur_kc1 = 0;
  int64_t before = 0;
  bool fresh = true;                       // no tile accumulated yet for (cur_c, cur_kc0)
  for (int64_t ti = ti0; ti < ti1; ++ti) {
    int cnt = 0;
    int64_t nt = 0;
    for (; c < em.n; ++c) {
      cnt = group_off[c + 1] - group_off[c];
      nt = (cnt + 127) / 128;
      const int64_t ni = nt * n_groups(c);
      if (ti < before + ni) break;
      before += ni;
    }
    if (c >= em.n) break;
    const int nchunk = (em.nc[c] + 63) / 64;
    const int grp = (int)((ti - before) / nt);
    const int kc0 = grp * kBMaxChunks, kc1 = kc0 + kBMaxChunks < nchunk ? kc0 + kBMaxChunks : nchunk;
    if (c != cur_c || kc0 != cur_kc0) {
      if (cur_c >= 0 && !fresh) flush(cur_c, cur_kc0, cur_kc1);
      __syncthreads();
      if (c != cur_c) {   // W1_c chunk (K-major for the forward; read MN-major here)
        const uint4* s4 = reinterpret_cast<const uint4*>(wsplit + em.woff[c] + (int64_t)nchunk * kEChunk);
        uint4* d4 = reinterpret_cast<uint4*>(sW1);
#pragma unroll
        for (int i = 0; i < kEChunk / 16 / kBThreads; ++i) d4[tid + i * kBThreads] = __ldg(s4 + tid + i * kBThreads);
      }
      cur_c = c; cur_kc0 = kc0; cur_kc1 = kc1; fresh = true;
    }
    const int off = (int)((ti - before) - (int64_t)grp * nt) * 128;
    const int nrows_cta = cnt - off < 128 ? cnt - off : 128;
    const bool live = r < nrows_cta;
    const int64_t t = live ? perm[group_off[c] + off + r] : 0;
    // ---- dE rows -> planes 0..7 of sS ----
    {
      float v[32];
      gather_half_rows(live ? dE + t * 64 + h * 32 : nullptr, 0, 64, stage, lane, v);
      put4(sS, 32768, h * 4, r, v);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)       // dH0[128 tok, 64] = dE[128 tok, 64 o] . W1[64 o, 64 k]
        umma_x3s(tmem_base + kColDH, sh + ks * 4096, sl + ks * 4096, vh + ks * 256, vl + ks * 256, 2048, 128, 128, 1024, idescD, ks == 0);
      umma_commit(&bar);
    }
    // H0 rows while the MMA runs
    float hv[32];
    gather_half_rows(live ? H0 + t * 64 + h * 32 : nullptr, 0, 64, stage, lane, hv);
    put4(sB, 16384, h * 4, r, hv);
    mbar_wait(&bar, phase);
    phase ^= 1;
    tc_fence_after();
    {
      uint32_t d0[32];
      tmem_ld32_issue(tlane + kColDH + h * 32, d0);
      tmem_ld_wait(d0);
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = live ? __uint_as_float(d0[i]) * (1.f - hv[i] * hv[i]) : 0.f;     // tanh'
      put4(sS, 32768, 8 + h * 4, r, v);
    }
    tc_fence_before();
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)       // [dE | dH0pre]^T[128 feat, 128 tok] . H0[128 tok, 64]: lanes 0..63 = dW1_c
        umma_x3s(tmem_base + kColW1, sh + ks * 256, sl + ks * 256, bh + ks * 256, bl + ks * 256, 128, 2048, 128, 2048, idescW,
                 fresh && ks == 0);
      umma_commit(&bar);
    }
    const float* frow = live ? em.feat[c] + (x[t] - em.start[c]) * em.ld[c] : nullptr;
    // feature chunk kc + 1 is gathered (global loads, dropout) while the tensor pipe works on chunk kc
    auto gather_chunk = [&](int kc, float (&v)[32]) {
      const int64_t kf = (int64_t)kc * 64 + h * 32;
      gather_half_rows(frow ? frow + kf : nullptr, kf, em.ld[c], stage, lane, v);
      if (drop.thr != 0u && live) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 d = drop_apply4(drop, (uint64_t)t, (uint32_t)(kf + 4 * j), make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
          v[4 * j] = d.x; v[4 * j + 1] = d.y; v[4 * j + 2] = d.z; v[4 * j + 3] = d.w;
        }
      }
    };
    float fv[32];
    gather_chunk(kc0, fv);
    mbar_wait(&bar, phase);                // dW1 contraction done: sB may be rewritten
    phase ^= 1;
    tc_fence_after();
    for (int kc = kc0; kc < kc1; ++kc) {
      put4(sB, 16384, h * 4, r, fv);
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)     // lanes 64..127 = dW0_c[:, chunk kc]
          umma_x3s(tmem_base + kColW0 + (kc - kc0) * 64, sh + ks * 256, sl + ks * 256, bh + ks * 256, bl + ks * 256, 128, 2048, 128,
                   2048, idescW, fresh && ks == 0);
        umma_commit(&bar);
      }
      if (kc + 1 < kc1) gather_chunk(kc + 1, fv);
      mbar_wait(&bar, phase);
      phase ^= 1;
      tc_fence_after();
    }
    fresh = false;
    tc_fence_before();
    __syncthreads();
  }
  if (cur_c >= 0 && !fresh) flush(cur_c, cur_kc0, cur_kc1);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ==========================================================================================
// Pipelined backward (default; enc_tc_bwd_kernel above stays as the cross-check, MATCHA_ENC_PIPE_BWD=0).  Same products, same
// work items (chromosome, column group of kBMaxChunks feature chunks, 128-token tile), same TMEM layout; different schedule:
//   * the feature chunks of an item arrive by cp.async (four producer warps, 16-byte pieces, two fp32 staging tiles) while
//     sixteen converter warps (thread = quarter row) turn the landed chunk into the bf16 hi | lo operand tile -- two tiles, so
//     the conversion of chunk k + 1 runs under the 24 MMAs of chunk k -- and one thread issues the MMAs;
//   * the dW1_c contraction is issued for the first column group of a chromosome only (the unit kernel accumulates it in
//     every group and publishes the first).
// ==========================================================================================
constexpr int kPBThreads = 672;                 // warps 0-15 converters, 16-19 producers (32 rows each), 20 MMA issuer
constexpr int kPBStages = 3;                    // two chunks of feature rows in flight while one is converted (two stages: 243 us at cfg3)
constexpr int kPBSmem = kBS + kBB + kPBStages * kPFStageF + kEChunk;          // 65 536 + 32 768 + 104 448 + 16 384 = 219 136

struct BwdItem {            // one work item of the backward list and what changes with it
  int c, grp, kc0, kc1, off, nrows, nchunk;
  bool new_group;
};

__global__ void __launch_bounds__(kPBThreads, 1)
enc_pipe_bwd_kernel(const EncMeta em, const uint8_t* __restrict__ wsplit, const int64_t* __restrict__ x,
                    const int32_t* __restrict__ perm, const int32_t* __restrict__ group_off, const float* __restrict__ dE,
                    const float* __restrict__ H0, float* __restrict__ grads, const DropCfg drop) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sS = smem;                                   // [dE | dH0pre] stacked operand tile, hi 32 KB | lo 32 KB
  uint8_t* sB = smem + kBS;                             // operand tile (H0 / feature chunk), hi 16 KB | lo 16 KB
  uint8_t* sStage = smem + kBS + kBB;                   // ring of fp32 staging tiles
  uint8_t* sW1 = sStage + kPBStages * kPFStageF;
  __shared__ uint64_t st_full[kPBStages], st_free[kPBStages], b_full, b_free, e_full, dh_full, s_full, item_done, flush_done;
  __shared__ uint32_t tmem_base_s;
  __shared__ const float* sRow[128];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 512) {
    for (int i = 0; i < kPBStages; ++i) { mbar_init(&st_full[i], 128); mbar_init(&st_free[i], 16); }
    mbar_init(&b_full, 16); mbar_init(&b_free, 1);
    mbar_init(&e_full, 16); mbar_init(&dh_full, 1); mbar_init(&s_full, 16); mbar_init(&item_done, 1); mbar_init(&flush_done, 16);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  // this CTA's contiguous range of the item list (as enc_tc_bwd_kernel)
  auto n_groups = [&](int cc) { return ((em.nc[cc] + 63) / 64 + kBMaxChunks - 1) / kBMaxChunks; };
  int64_t total = 0;
  for (int c = 0; c < em.n; ++c) total += (int64_t)((group_off[c + 1] - group_off[c] + 127) / 128) * n_groups(c);
  const int64_t per = (total + gridDim.x - 1) / gridDim.x;
  const int64_t ti0 = (int64_t)blockIdx.x * per, ti1 = (ti0 + per < total) ? ti0 + per : total;
  struct Walk { int c; int64_t before; int cur_c, cur_kc0; };
  auto next_item = [&](Walk& w, int64_t ti, BwdItem& it) -> bool {
    int cnt = 0;
    int64_t nt = 0;
    for (; w.c < em.n; ++w.c) {
      cnt = group_off[w.c + 1] - group_off[w.c];
      nt = (cnt + 127) / 128;
      const int64_t ni = nt * n_groups(w.c);
      if (ti < w.before + ni) break;
      w.before += ni;
    }
    if (w.c >= em.n) return false;
    it.c = w.c;
    it.nchunk = (em.nc[w.c] + 63) / 64;
    it.grp = (int)((ti - w.before) / nt);
    it.kc0 = it.grp * kBMaxChunks;
    it.kc1 = it.kc0 + kBMaxChunks < it.nchunk ? it.kc0 + kBMaxChunks : it.nchunk;
    it.off = (int)((ti - w.before) - (int64_t)it.grp * nt) * 128;
    it.nrows = cnt - it.off < 128 ? cnt - it.off : 128;
    it.new_group = it.c != w.cur_c || it.kc0 != w.cur_kc0;
    w.cur_c = it.c; w.cur_kc0 = it.kc0;
    return true;
  };

  if (warp >= 16 && warp < 20) {
    // ---------------- producers: the feature chunks of every item ----------------
    const int half = lane >> 4, piece = lane & 15, pw = warp - 16;
    const float** myRow = sRow + pw * 32;
    Walk w{0, 0, -1, -1};
    BwdItem it;
    uint32_t cidx = 0;
    for (int64_t ti = ti0; ti < ti1; ++ti) {
      if (!next_item(w, ti, it)) break;
      const int64_t ld = em.ld[it.c];
      __syncwarp();
      {
        const int r = pw * 32 + lane;
        const float* fr = em.feat[it.c];
        if (r < it.nrows) fr += (x[perm[group_off[it.c] + it.off + r]] - em.start[it.c]) * ld;
        myRow[lane] = fr;
      }
      __syncwarp();
      for (int kc = it.kc0; kc < it.kc1; ++kc, ++cidx) {
        const int st = (int)(cidx % kPBStages);
        mbar_wait_backoff(&st_free[st], ((cidx / kPBStages) & 1u) ^ 1u);
        const int64_t col = (int64_t)kc * 64 + piece * 4;
        const bool col_ok = col < ld;
        const uint32_t dst0 = smem_u32(sStage + st * kPFStageF) + (pw * 32 + half) * kPFRow + piece * 16;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int r = 2 * j + half;
          const float* src = myRow[r] + (col_ok ? col : 0);
          const uint32_t nbytes = (col_ok && pw * 32 + r < it.nrows) ? 16u : 0u;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst0 + 2 * j * kPFRow), "l"(src), "r"(nbytes) : "memory");
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&st_full[st])) : "memory");
      }
    }
  } else if (warp == 20) {
    // ---------------- MMA issuer ----------------
    if (elect_one()) {
      constexpr uint32_t idescD = make_idesc(128, 64, false, true);     // dH0 = dE . W1          (B MN-major)
      constexpr uint32_t idescW = make_idesc(128, 64, true, true);      // [dE | dH0pre]^T . tile  (both MN-major)
      const uint32_t sh = smem_u32(sS), sl = sh + 32768;
      const uint32_t vh = smem_u32(sW1), vl = vh + 8192;
      Walk w{0, 0, -1, -1};
      BwdItem it;
      uint32_t bidx = 0, iidx = 0, nflush = 0;
      bool fresh = true, first_item = true;
      for (int64_t ti = ti0; ti < ti1; ++ti, ++iidx) {
        if (!next_item(w, ti, it)) break;
        if (it.new_group) {
          if (!first_item) { mbar_wait_backoff(&flush_done, nflush & 1u); ++nflush; }      // the accumulators were read out
          fresh = true;
        }
        first_item = false;
        mbar_wait_backoff(&e_full, iidx & 1u);
        BTRACE(8);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)       // dH0[128 tok, 64] = dE[128 tok, 64 o] . W1[64 o, 64 k]
          umma_x3s(tmem_base + kColDH, sh + ks * 4096, sl + ks * 4096, vh + ks * 256, vl + ks * 256, 2048, 128, 128, 1024, idescD, ks == 0);
        umma_commit(&dh_full);
        mbar_wait_backoff(&s_full, iidx & 1u);
        BTRACE(9);
        tc_fence_after();
        auto wgrad = [&](uint32_t dcol) {    // [dE | dH0pre]^T[128 feat, 128 tok] . operand tile[128 tok, 64]
          mbar_wait_backoff(&b_full, bidx & 1u);
          tc_fence_after();
          const uint32_t bh = smem_u32(sB), bl = bh + 16384;
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            umma_x3s(tmem_base + dcol, sh + ks * 256, sl + ks * 256, bh + ks * 256, bl + ks * 256, 128, 2048, 128, 2048, idescW,
                     fresh && ks == 0);
          umma_commit(&b_free);
          ++bidx;
        };
        if (it.kc0 == 0) wgrad(kColW1);                                     // lanes 0..63 = dW1_c (first column group only)
        for (int kc = it.kc0; kc < it.kc1; ++kc) wgrad(kColW0 + (kc - it.kc0) * 64);     // lanes 64..127 = dW0_c[:, chunk kc]
        umma_commit(&item_done);
        BTRACE(10);
        fresh = false;
      }
    }
  } else {
    // ---------------- converters: thread (r, q4) = token row r (TMEM lane r), quarter q4 of every 64-wide row ----------------
    const int r = tid & 127, q4 = tid >> 7, wq = warp & 3;
    const uint32_t tlane = tmem_base + ((uint32_t)(wq * 32) << 16);
    Walk w{0, 0, -1, -1};
    BwdItem it;
    uint32_t bidx = 0, cidx = 0, iidx = 0;
    int cur_c = -1, prev_c = -1, prev_kc0 = 0, prev_kc1 = 0;
    bool have_prev = false;
    auto put_quarter = [&](uint8_t* tile, int lo_off, int plane0, const float (&v)[16]) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        uint4 hi, lo;
        split8(make_float4(v[8 * j], v[8 * j + 1], v[8 * j + 2], v[8 * j + 3]),
               make_float4(v[8 * j + 4], v[8 * j + 5], v[8 * j + 6], v[8 * j + 7]), hi, lo);
        sts16(tile + (plane0 + j) * 2048 + r * 16, hi);
        sts16(tile + lo_off + (plane0 + j) * 2048 + r * 16, lo);
      }
    };
    auto put_b = [&](const float (&v)[16]) {       // next operand tile: stored as soon as the previous one's MMAs have read it
      mbar_wait(&b_free, (bidx & 1u) ^ 1u);
      put_quarter(sB, 16384, q4 * 2, v);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&b_full);
      ++bidx;
    };
    // TMEM -> atomics on the weight gradients of chromosome c, chunks [kc0, kc1): TMEM lane = stacked feature row
    auto flush = [&](int c, int kc0, int kc1) {
      tc_fence_after();
      const int rr = wq * 32 + lane;
      if (rr < 64) {                         // warps of lane quarters 0, 1: dW1_c[o = rr][k], this thread: k in [16 q4, 16 q4 + 16)
        if (kc0 == 0) {
          float v[16];
          tmem_ld16(tlane + kColW1 + q4 * 16, v);
          float* dst = grads + em.off_w1[c] + (int64_t)rr * 64 + q4 * 16;
#pragma unroll
          for (int i = 0; i < 16; ++i) atomicAdd(dst + i, v[i]);
        }
      } else {                               // lane quarters 2, 3: dW0_c[f = rr - 64][k]
        const int nc = em.nc[c];
        float* dst = grads + em.off_w0[c] + (int64_t)(rr - 64) * nc;
        for (int kc = kc0; kc < kc1; ++kc) {
          float v[16];
          tmem_ld16(tlane + kColW0 + (kc - kc0) * 64 + q4 * 16, v);
          const int k0 = kc * 64 + q4 * 16;
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (k0 + i < nc) atomicAdd(dst + k0 + i, v[i]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&flush_done);
    };
    for (int64_t ti = ti0; ti < ti1; ++ti, ++iidx) {
      if (!next_item(w, ti, it)) break;
      if (tid == 0) BTRACE(0);
      if (have_prev) mbar_wait(&item_done, (iidx - 1) & 1u);      // the previous item's MMAs have read sS (and finished its group)
      if (tid == 0) BTRACE(1);
      if (it.new_group && have_prev) flush(prev_c, prev_kc0, prev_kc1);
      if (tid == 0) BTRACE(2);
      if (it.c != cur_c) {        // W1_c chunk (K-major for the forward; read MN-major here); no MMA is in flight
        asm volatile("bar.sync 1, 512;" ::: "memory");
        const uint4* s4 = reinterpret_cast<const uint4*>(wsplit + em.woff[it.c] + (int64_t)it.nchunk * kEChunk);
        uint4* d4 = reinterpret_cast<uint4*>(sW1);
#pragma unroll
        for (int i = 0; i < kEChunk / 16 / 512; ++i) d4[tid + i * 512] = __ldg(s4 + tid + i * 512);
        cur_c = it.c;
      }
      const bool live = r < it.nrows;
      const int64_t t = live ? perm[group_off[it.c] + it.off + r] : 0;
      // ---- dE rows -> planes 0..7 of sS (requesting them before the wait above was measured slower: 244 -> 268 us) ----
      float hv[16];
      {
        float v[16];
        const float4* src = reinterpret_cast<const float4*>(dE + t * 64 + q4 * 16);
        const float4* hsrc = reinterpret_cast<const float4*>(H0 + t * 64 + q4 * 16);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 e = live ? __ldg(src + j) : make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 hh = live ? __ldg(hsrc + j) : make_float4(0.f, 0.f, 0.f, 0.f);
          v[4 * j] = e.x; v[4 * j + 1] = e.y; v[4 * j + 2] = e.z; v[4 * j + 3] = e.w;
          hv[4 * j] = hh.x; hv[4 * j + 1] = hh.y; hv[4 * j + 2] = hh.z; hv[4 * j + 3] = hh.w;
        }
        put_quarter(sS, 32768, q4 * 2, v);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&e_full);
      if (tid == 0) BTRACE(3);
      if (it.kc0 == 0) put_b(hv);            // H0 tile: the B operand of dW1_c
      // ---- dH0pre = (dE . W1) * (1 - H0^2) -> planes 8..15 of sS ----
      {
        mbar_wait(&dh_full, iidx & 1u);
        if (tid == 0) BTRACE(4);
        tc_fence_after();
        float d0[16];
        tmem_ld16(tlane + kColDH + q4 * 16, d0);
        tc_fence_before();
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = live ? d0[i] * (1.f - hv[i] * hv[i]) : 0.f;      // tanh'
        put_quarter(sS, 32768, 8 + q4 * 2, v);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_full);
      if (tid == 0) BTRACE(5);
      // ---- feature chunks ----
      for (int kc = it.kc0; kc < it.kc1; ++kc, ++cidx) {
        const int st = (int)(cidx % kPBStages);
        mbar_wait(&st_full[st], (cidx / kPBStages) & 1u);
        const float* srow = reinterpret_cast<const float*>(sStage + st * kPFStageF + r * kPFRow) + q4 * 16;
        const int64_t kf = (int64_t)kc * 64 + q4 * 16;
        float v[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float4 e = *reinterpret_cast<const float4*>(srow + 4 * j);
          if (drop.thr != 0u && live) e = drop_apply4(drop, (uint64_t)t, (uint32_t)(kf + 4 * j), e);
          v[4 * j] = e.x; v[4 * j + 1] = e.y; v[4 * j + 2] = e.z; v[4 * j + 3] = e.w;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&st_free[st]);      // the staging tile is in registers
        put_b(v);
      }
      if (tid == 0) BTRACE(6);
      prev_c = it.c; prev_kc0 = it.kc0; prev_kc1 = it.kc1; have_prev = true;
    }
    if (have_prev) {
      mbar_wait(&item_done, (iidx - 1) & 1u);
      flush(prev_c, prev_kc0, prev_kc1);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

EncMeta make_meta(const matcha_model_desc* m, int64_t* total_bytes);
}  // namespace
#ifdef MATCHA_ENC_TRACE
extern "C" int matcha_enc_trace(unsigned long long* out) {
  return cudaMemcpyFromSymbol(out, g_etrace, sizeof(unsigned long long) * 4096) == cudaSuccess ? 0 : -2;
}
#endif
namespace {
EncMeta make_meta(const matcha_model_desc* m, int64_t* total_bytes) {
  EncMeta em;
  memset(&em, 0, sizeof(em));
  em.n = m->n_chrom;
  int64_t off = 0;
  for (int c = 0; c < m->n_chrom; ++c) {
    em.nc[c] = (int32_t)(m->chrom_end[c] - m->chrom_start[c]);
    em.start[c] = m->chrom_start[c];
    em.ld[c] = m->feat_ld[c];
    em.feat[c] = m->feat[c];
    em.off_w0[c] = m->off_w0[c];
    em.off_w1[c] = m->off_w1[c];
    em.woff[c] = off;
    off += ((em.nc[c] + 63) / 64 + 1) * (int64_t)kEChunk;
  }
  if (total_bytes) *total_bytes = off;
  return em;
}

}  // namespace

// forward: 0 unit kernel, 1 pipelined, 2 by feature-row width (default); backward: 0 unit, 1 pipelined (default); -1 = environment
static int g_enc_pipe_fwd = -1, g_enc_pipe_bwd = -1;
void set_enc_pipe(int fwd, int bwd) { g_enc_pipe_fwd = fwd; g_enc_pipe_bwd = bwd; }

// floats of the derived buffer taken by the pre-split encoder weights (dense feature rows, embed_dim 64)
int64_t enc_tc_split_floats(const matcha_model_desc* m) {
  int64_t bytes = 0;
  make_meta(m, &bytes);
  return bytes / 4;
}

int launch_enc_tc_prepare(const matcha_model_desc* m, int64_t split_base, cudaStream_t s) {
  int64_t bytes = 0;
  const EncMeta em = make_meta(m, &bytes);
  const int64_t units = bytes / kEChunk * 512;
  enc_split_kernel<<<(unsigned)((units + 255) / 256), 256, 0, s>>>(em, m->params, reinterpret_cast<uint8_t*>(m->derived + split_base), units);
  MATCHA_CHECK_LAUNCH("enc_split");
  return MATCHA_OK;
}

// H0, E rows of the real tokens (pad rows are left untouched: the caller zero-fills both tensors first)
int launch_enc_tc_fwd(const matcha_model_desc* m, int64_t split_base, const int64_t* x, int64_t T, const int32_t* perm,
                      const int32_t* group_off, float* H0, float* E, DropCfg drop, cudaStream_t s) {
  if (T <= 0) return MATCHA_OK;
  static bool once = false;
  if (!once) {
    if (int rc = check_cuda(cudaFuncSetAttribute(enc_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kESmem),
                            "cudaFuncSetAttribute"))
      return rc;
    once = true;
  }
  const EncMeta em = make_meta(m, nullptr);
  int64_t tiles = (T + 127) / 128 + m->n_chrom;            // upper bound; the kernel stops at the real tile count
  // pipelined kernel for wide feature rows (cfg3: 1 319 bins per chromosome on average, cfg4: 24 897); at cfg2 (133 bins: at
  // most four chunks per tile) the per-tile epilogue dominates and two unit CTAs per SM overlap it better (measured: 39 us vs
  // 46 us at cfg2, 203 us vs 155 us at cfg3).  MATCHA_ENC_PIPE = 0 / 1 forces one of them
  int& pipe = g_enc_pipe_fwd;
  if (pipe < 0) {
    const char* e = getenv("MATCHA_ENC_PIPE");
    pipe = !e ? 2 : (e[0] == '0' ? 0 : 1);
  }
  int64_t bins = 0;
  for (int c = 0; c < m->n_chrom; ++c) bins += m->chrom_end[c] - m->chrom_start[c];
  if (pipe == 1 || (pipe == 2 && bins >= 512 * (int64_t)m->n_chrom)) {
    static bool once2 = false;
    if (!once2) {
      if (int rc = check_cuda(cudaFuncSetAttribute(enc_pipe_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPFSmem),
                              "cudaFuncSetAttribute"))
        return rc;
      once2 = true;
    }
    enc_pipe_fwd_kernel<<<(unsigned)(tiles < kSMs ? tiles : kSMs), kPFThreads, kPFSmem, s>>>(
        em, reinterpret_cast<const uint8_t*>(m->derived + split_base), x, perm, group_off, H0, E, drop);
    MATCHA_CHECK_LAUNCH("enc_pipe_fwd");
    return MATCHA_OK;
  }
  const unsigned grid = (unsigned)(tiles < 2 * kSMs ? tiles : 2 * kSMs);
  enc_tc_fwd_kernel<<<grid, kFThreads, kESmem, s>>>(em, reinterpret_cast<const uint8_t*>(m->derived + split_base), x, perm,
                                                    group_off, H0, E, drop);
  MATCHA_CHECK_LAUNCH("enc_tc_fwd");
  return MATCHA_OK;
}


// the backward kernel keeps kBMaxChunks feature chunks resident in TMEM and walks wider chromosomes in column groups
bool enc_tc_bwd_fits(const matcha_model_desc*) { return true; }
static int64_t max_col_groups(const matcha_model_desc* m) {
  int64_t g = 1;
  for (int c = 0; c < m->n_chrom; ++c) {
    const int64_t nchunk = (m->chrom_end[c] - m->chrom_start[c] + 63) / 64;
    const int64_t gc = (nchunk + kBMaxChunks - 1) / kBMaxChunks;
    if (gc > g) g = gc;
  }
  return g;
}

// grads (flat, same layout as params): dW1_c += dE^T H0, dW0_c += ((dE W1_c) * (1 - H0^2))^T dropout(F_c rows)
int launch_enc_tc_bwd(const matcha_model_desc* m, int64_t split_base, const int64_t* x, int64_t T, const int32_t* perm,
                      const int32_t* group_off, const float* dE, const float* H0, DropCfg drop, cudaStream_t s) {
  if (T <= 0) return MATCHA_OK;
  static bool once = false;
  if (!once) {
    if (int rc = check_cuda(cudaFuncSetAttribute(enc_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBSmem),
                            "cudaFuncSetAttribute"))
      return rc;
    once = true;
  }
  const EncMeta em = make_meta(m, nullptr);
  const int64_t tiles = ((T + 127) / 128 + m->n_chrom) * max_col_groups(m);     // upper bound of the item count
  const unsigned grid = (unsigned)(tiles < kSMs ? tiles : kSMs);
  int& pipe = g_enc_pipe_bwd;
  if (pipe < 0) {
    const char* e = getenv("MATCHA_ENC_PIPE_BWD");
    pipe = (e && e[0] == '0') ? 0 : 1;
  }
  if (pipe == 1) {
    static bool once2 = false;
    if (!once2) {
      if (int rc = check_cuda(cudaFuncSetAttribute(enc_pipe_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPBSmem),
                              "cudaFuncSetAttribute"))
        return rc;
      once2 = true;
    }
    enc_pipe_bwd_kernel<<<grid, kPBThreads, kPBSmem, s>>>(em, reinterpret_cast<const uint8_t*>(m->derived + split_base), x, perm,
                                                         group_off, dE, H0, m->grads, drop);
    MATCHA_CHECK_LAUNCH("enc_pipe_bwd");
    return MATCHA_OK;
  }
  enc_tc_bwd_kernel<<<grid, kBThreads, kBSmem, s>>>(em, reinterpret_cast<const uint8_t*>(m->derived + split_base), x, perm,
                                                    group_off, dE, H0, m->grads, drop);
  MATCHA_CHECK_LAUNCH("enc_tc_bwd");
  return MATCHA_OK;
}

}  // namespace matcha
