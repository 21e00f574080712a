// "Row-chain" kernels: the 64-wide position-wise layers around the attention block (Modules.py:263-270 attribute mix,
// :353-376 pff_n1, :290-311 scorer) as chains of tcgen05 contractions whose activations stay in registers.
//
// One CTA = the 128 rows of one hyperedge-aligned tile (rowwise.cuh).  The forward kernels and the pff backward run 256
// threads: thread (r, h) owns half h (32 columns) of tile row r = TMEM lane r, so 16 warps per SM (at 128 registers) hide
// the load -> MMA -> read-out latency chain; row statistics meet through a 64-thread named barrier per warp pair.  The
// LayerNorm / next_w backward, which is bound by its nine row streams rather than by latency, keeps 128 threads (thread r
// owns the whole row r).  A stage = { the thread splits its fp32 row into bf16 hi | lo and stores it into the shared A tile
// (and, when a consumer needs it, into the same tile in global memory) -> one thread issues the bf16x3 MMAs against a
// pre-split 64x64 weight resident in shared memory -> every thread reads its output row back with tcgen05.ld }.
// Several CTAs per SM (<= 64 KB shared memory, 64 TMEM columns each) overlap each other's load / MMA / store phases.
#include "rowwise.cuh"
#include "tc_common.cuh"

namespace matcha {
namespace {

constexpr int kCThreads = 128;
constexpr int kC2Threads = 256;               // half a row per thread: 16 warps per SM at 128 registers
constexpr int kCTile = 32768;                 // A tile: hi 16 KB (8 planes x 128 rows x 16 B) | lo 16 KB
constexpr int kCW = kChainWBytes;             // one 64x64 weight: hi 8 KB | lo 8 KB
constexpr float kLnEpsC = 1e-5f;

// ------------------------------------------------------------------------------------------
// parameter preparation: W [64][64] row-major -> (a) K-major [k/8][n][8] for Y = X W^T, (b) the layout of W read as
// B[N = column, K = row] MN-major ([row/8][col/8][row%8][8 cols]) for dX = dY W.  LBO 1024 / SBO 128 in both.
// ------------------------------------------------------------------------------------------
struct SplitW3 { const float* W[3]; uint8_t* out_k[3]; uint8_t* out_mn[3]; };
__global__ void split_w64_kernel(const SplitW3 a) {          // grid = 2 blocks per weight
  const float* __restrict__ W = a.W[blockIdx.x >> 1];
  uint8_t* __restrict__ out_k = a.out_k[blockIdx.x >> 1];
  uint8_t* __restrict__ out_mn = a.out_mn[blockIdx.x >> 1];
  const int u = (blockIdx.x & 1) * blockDim.x + threadIdx.x;       // unit = (row r, group g of 8 columns)
  if (u >= 64 * 8) return;
  const int g = u & 7, r = u >> 3;
  const float* src = W + r * 64 + g * 8;
  uint4 hi, lo;
  split8(__ldg(reinterpret_cast<const float4*>(src)), __ldg(reinterpret_cast<const float4*>(src + 4)), hi, lo);
  const int off_k = g * 1024 + r * 16;
  *reinterpret_cast<uint4*>(out_k + off_k) = hi;
  *reinterpret_cast<uint4*>(out_k + 8192 + off_k) = lo;
  const int off_mn = (r >> 3) * 1024 + g * 128 + (r & 7) * 16;
  *reinterpret_cast<uint4*>(out_mn + off_mn) = hi;
  *reinterpret_cast<uint4*>(out_mn + 8192 + off_mn) = lo;
}

// ------------------------------------------------------------------------------------------
// stage helpers
// ------------------------------------------------------------------------------------------
struct ChainCtx {
  uint8_t* sA;
  uint64_t* bar;
  uint32_t tmem;       // TMEM base (64 columns)
  uint32_t phase;
  int r;               // tile row of this thread
};

// split one fp32 row into the shared A tile and (optionally) the same tile in global memory
__device__ __forceinline__ void put_row(uint8_t* sA, uint8_t* gT, int r, const float (&v)[64]) {   // gT: kCTile layout
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    uint4 hi, lo;
    split8(make_float4(v[8 * j], v[8 * j + 1], v[8 * j + 2], v[8 * j + 3]),
           make_float4(v[8 * j + 4], v[8 * j + 5], v[8 * j + 6], v[8 * j + 7]), hi, lo);
    sts16(sA + j * 2048 + r * 16, hi);
    sts16(sA + 16384 + j * 2048 + r * 16, lo);
    if (gT) {
      *reinterpret_cast<uint4*>(gT + j * 2048 + r * 16) = hi;
      *reinterpret_cast<uint4*>(gT + 16384 + j * 2048 + r * 16) = lo;
    }
  }
}
// only the global copy (rows that feed a later weight-gradient kernel but no MMA of this kernel)
__device__ __forceinline__ void put_row_global(uint8_t* gT, int half_bytes, int r, const float (&v)[64]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    uint4 hi, lo;
    split8(make_float4(v[8 * j], v[8 * j + 1], v[8 * j + 2], v[8 * j + 3]),
           make_float4(v[8 * j + 4], v[8 * j + 5], v[8 * j + 6], v[8 * j + 7]), hi, lo);
    *reinterpret_cast<uint4*>(gT + j * 2048 + r * 16) = hi;
    *reinterpret_cast<uint4*>(gT + half_bytes + j * 2048 + r * 16) = lo;
  }
}

__device__ __forceinline__ void tmem_ld32_issue_c(uint32_t taddr, uint32_t (&r)[32]) {
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

// the row is already in the A tile: publish it, run [128 x 64] . W (64 x 64, pre-split at shared address w_hi), read the
// output row back.  Must be called by all 128 threads.
__device__ __forceinline__ void run_stage(ChainCtx& c, uint32_t w_hi, uint32_t idesc, float (&out)[64]) {
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    tc_fence_after();
    const uint32_t ah = smem_u32(c.sA), al = ah + 16384;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
      umma_x3s(c.tmem, ah + ks * 4096, al + ks * 4096, w_hi + ks * 2048, w_hi + 8192 + ks * 2048, 2048, 128, 1024, 128, idesc,
               ks == 0);
    umma_commit(c.bar);
  }
  mbar_wait(c.bar, c.phase);
  c.phase ^= 1;
  tc_fence_after();
  const uint32_t taddr = c.tmem + ((uint32_t)((c.r >> 5) * 32) << 16);
  uint32_t a[32], b[32];
  tmem_ld32_issue_c(taddr, a);
  tmem_ld32_issue_c(taddr + 32, b);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) { asm volatile("" : "+r"(a[i])); asm volatile("" : "+r"(b[i])); }
#pragma unroll
  for (int i = 0; i < 32; ++i) { out[i] = __uint_as_float(a[i]); out[32 + i] = __uint_as_float(b[i]); }
  tc_fence_before();
}

// Coalesced row I/O.  A warp's RPW tile rows are consecutive tokens, i.e. one contiguous run of fp32 in global memory:
// the warp moves it with fully coalesced 128-bit accesses through a padded staging area (68 floats per row: conflict
// free for both access patterns) and every thread picks up / drops off its own 64-float row there.
constexpr int kStageRow = 68;
constexpr int kStageWarp = 32 * kStageRow;          // floats per warp
constexpr int kStageBytes = 4 * kStageWarp * 4;     // 34816
__device__ __forceinline__ void warp_load_rows(const float* __restrict__ g, int nrows, float* stage, int lane, float (&v)[64]) {
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int idx = k * 32 + lane, row = idx >> 4, c4 = idx & 15;
    float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < nrows) e = __ldg(reinterpret_cast<const float4*>(g) + idx);
    *reinterpret_cast<float4*>(stage + row * kStageRow + c4 * 4) = e;
  }
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const float4 e = *reinterpret_cast<const float4*>(stage + lane * kStageRow + k * 4);
    v[4 * k] = e.x; v[4 * k + 1] = e.y; v[4 * k + 2] = e.z; v[4 * k + 3] = e.w;
  }
  __syncwarp();
}
// v += rows (same traffic pattern, accumulating)
__device__ __forceinline__ void warp_add_rows(const float* __restrict__ g, int nrows, float* stage, int lane, float (&v)[64]) {
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int idx = k * 32 + lane, row = idx >> 4, c4 = idx & 15;
    float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < nrows) e = __ldg(reinterpret_cast<const float4*>(g) + idx);
    *reinterpret_cast<float4*>(stage + row * kStageRow + c4 * 4) = e;
  }
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const float4 e = *reinterpret_cast<const float4*>(stage + lane * kStageRow + k * 4);
    v[4 * k] += e.x; v[4 * k + 1] += e.y; v[4 * k + 2] += e.z; v[4 * k + 3] += e.w;
  }
  __syncwarp();
}
__device__ __forceinline__ void warp_store_rows(float* __restrict__ g, int nrows, float* stage, int lane, const float (&v)[64]) {
#pragma unroll
  for (int k = 0; k < 16; ++k)
    *reinterpret_cast<float4*>(stage + lane * kStageRow + k * 4) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int idx = k * 32 + lane, row = idx >> 4, c4 = idx & 15;
    if (row < nrows) reinterpret_cast<float4*>(g)[idx] = *reinterpret_cast<const float4*>(stage + row * kStageRow + c4 * 4);
  }
  __syncwarp();
}
// live rows of this warp's slice of the tile
__device__ __forceinline__ int warp_rows(int64_t t0, int rpw, int64_t T) {
  const int64_t n = T - t0;
  return n <= 0 ? 0 : (n < rpw ? (int)n : rpw);
}

__device__ __forceinline__ void load_weight(uint8_t* dst, const uint8_t* src) {   // 16 KB, all 128 threads
  const uint4* s = reinterpret_cast<const uint4*>(src);
  uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int i = 0; i < kCW / 16 / kCThreads; ++i) d[threadIdx.x + i * kCThreads] = __ldg(s + threadIdx.x + i * kCThreads);
}

// ---- half-row helpers for the 256-thread kernels: thread (r, h) owns token row r (= TMEM lane r) and columns [32 h, 32 h + 32)
constexpr int kHalfStageRow = 36;                          // 32 floats + pad: conflict free for both access patterns
constexpr int kHalfStageBytes = 8 * 32 * kHalfStageRow * 4;   // 36 864: one 32 x 32 staging block per warp
// 32 consecutive token rows x 32 floats (this warp's column half of a [*, 64] fp32 array) <-> registers, coalesced
__device__ __forceinline__ void half_rows_load(const float* __restrict__ g, int nrows, float* stage, int lane, float (&v)[32]) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int idx = k * 32 + lane, row = idx >> 3, c4 = idx & 7;
    float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < nrows) e = __ldg(reinterpret_cast<const float4*>(g + row * 64) + c4);
    *reinterpret_cast<float4*>(stage + row * kHalfStageRow + c4 * 4) = e;
  }
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float4 e = *reinterpret_cast<const float4*>(stage + lane * kHalfStageRow + k * 4);
    v[4 * k] = e.x; v[4 * k + 1] = e.y; v[4 * k + 2] = e.z; v[4 * k + 3] = e.w;
  }
  __syncwarp();
}
__device__ __forceinline__ void half_rows_store(float* __restrict__ g, int nrows, float* stage, int lane, const float (&v)[32]) {
#pragma unroll
  for (int k = 0; k < 8; ++k)
    *reinterpret_cast<float4*>(stage + lane * kHalfStageRow + k * 4) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int idx = k * 32 + lane, row = idx >> 3, c4 = idx & 7;
    if (row < nrows) reinterpret_cast<float4*>(g + row * 64)[c4] = *reinterpret_cast<const float4*>(stage + row * kHalfStageRow + c4 * 4);
  }
  __syncwarp();
}
// 32 floats of row r -> planes p0 .. p0 + 3 of the shared A tile and (optionally) the same tile in global memory
__device__ __forceinline__ void put_half_row(uint8_t* sA, uint8_t* gT, int p0, int r, const float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 hi, lo;
    split8(make_float4(v[8 * j], v[8 * j + 1], v[8 * j + 2], v[8 * j + 3]),
           make_float4(v[8 * j + 4], v[8 * j + 5], v[8 * j + 6], v[8 * j + 7]), hi, lo);
    sts16(sA + (p0 + j) * 2048 + r * 16, hi);
    sts16(sA + 16384 + (p0 + j) * 2048 + r * 16, lo);
    if (gT) {
      *reinterpret_cast<uint4*>(gT + (p0 + j) * 2048 + r * 16) = hi;
      *reinterpret_cast<uint4*>(gT + 16384 + (p0 + j) * 2048 + r * 16) = lo;
    }
  }
}
// the 256-thread form of run_stage: every thread reads back its 32 columns of the output row
__device__ __forceinline__ void run_stage_half(ChainCtx& c, uint32_t w_hi, uint32_t idesc, int h, float (&out)[32]) {
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    tc_fence_after();
    const uint32_t ah = smem_u32(c.sA), al = ah + 16384;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
      umma_x3s(c.tmem, ah + ks * 4096, al + ks * 4096, w_hi + ks * 2048, w_hi + 8192 + ks * 2048, 2048, 128, 1024, 128, idesc,
               ks == 0);
    umma_commit(c.bar);
  }
  mbar_wait(c.bar, c.phase);
  c.phase ^= 1;
  tc_fence_after();
  const uint32_t taddr = c.tmem + ((uint32_t)((c.r >> 5) * 32) << 16) + h * 32;
  uint32_t a[32];
  tmem_ld32_issue_c(taddr, a);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) asm volatile("" : "+r"(a[i]));
#pragma unroll
  for (int i = 0; i < 32; ++i) out[i] = __uint_as_float(a[i]);
  tc_fence_before();
}

// sums across the two threads that share a token row (warps w and w + 4 of the 256-thread kernels): one 64-thread named
// barrier per exchange; every exchange of a tile uses its own slot array, and a slot is rewritten one tile later, behind
// the CTA-wide barriers of the tile's contraction stages.  Fixed order (half 0 + half 1): both threads get the same bits.
__device__ __forceinline__ float pair_sum(float v, float* slot, int r, int h, int bar_id) {
  slot[h * 128 + r] = v;
  asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
  return slot[r] + slot[128 + r];
}
// LayerNorm statistics of a row split over two threads (biased variance, eps 1e-5); v <- normalised half row
__device__ __forceinline__ float ln_half_row(float (&v)[32], float* slot_mean, float* slot_var, int r, int h, int bar_id) {
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < 32; ++c) s += v[c];
  const float mean = pair_sum(s, slot_mean, r, h, bar_id) * (1.0f / 64);
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < 32; ++c) { v[c] -= mean; q = fmaf(v[c], v[c], q); }
  const float rstd = 1.0f / sqrtf(pair_sum(q, slot_var, r, h, bar_id) * (1.0f / 64) + kLnEpsC);
#pragma unroll
  for (int c = 0; c < 32; ++c) v[c] *= rstd;
  return rstd;
}
// only the global copy of a half row (tiles that feed a later kernel but no MMA of this one)
__device__ __forceinline__ void put_half_row_global(uint8_t* gT, int half_bytes, int p0, int r, const float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 hi, lo;
    split8(make_float4(v[8 * j], v[8 * j + 1], v[8 * j + 2], v[8 * j + 3]),
           make_float4(v[8 * j + 4], v[8 * j + 5], v[8 * j + 6], v[8 * j + 7]), hi, lo);
    *reinterpret_cast<uint4*>(gT + (p0 + j) * 2048 + r * 16) = hi;
    *reinterpret_cast<uint4*>(gT + half_bytes + (p0 + j) * 2048 + r * 16) = lo;
  }
}
__device__ __forceinline__ void load_weight2(uint8_t* dst, const uint8_t* src) {   // 16 KB, all 256 threads
  for (int i = threadIdx.x; i < kCW / 16; i += kC2Threads) reinterpret_cast<uint4*>(dst)[i] = __ldg(reinterpret_cast<const uint4*>(src) + i);
}

// ==========================================================================================
// F1: V0 = E + attribute_nn(attr[id]);  X = tanh(next_w V0 + b);  xhat, rstd = LayerNorm statistics of X
// ==========================================================================================
struct MixArgs {
  const float* E; const int64_t* x; const float* attr_table; int attr_dim; const float* attr_w; const float* attr_b;
  const uint8_t* w_next; const float* next_b;
  float* V0; float* X; float* xhat; float* rstd;
  uint8_t* xt;       // hyperedge-aligned xhat tiles (kXTileBytes each)
  uint8_t* v0t;      // V0 tiles (32 KB each: hi 16 KB | lo 16 KB) for the weight-gradient kernel, or NULL
  uint8_t* attrt;    // attribute-row tiles [128 x 32] hi 8 KB | lo 8 KB, or NULL
  int64_t T;
};

template <int L>
__global__ void __launch_bounds__(kC2Threads, 2) chain_mix_fwd_kernel(const MixArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sW = smem + kCTile;
  float* sStage = reinterpret_cast<float*>(smem + kCTile + kCW);
  float* sWt = reinterpret_cast<float*>(smem + kCTile + kCW + kHalfStageBytes);     // attribute_nn.weight^T [attr_dim][64]
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float sNb[64], sAb[64];
  __shared__ float sX[2][256];
  constexpr int RPW = (32 / L) * L;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r = tid & 127, h = tid >> 7, wq = warp & 3, c0 = h * 32, pbar = 1 + wq;
  const int64_t ntiles = (a.T + 4 * RPW - 1) / (4 * RPW);
  if (warp == 0) tmem_alloc(&tmem_base_s, 64);
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  load_weight2(sW, a.w_next);
  for (int i = tid; i < a.attr_dim * 64; i += kC2Threads) sWt[i] = __ldg(a.attr_w + (i & 63) * a.attr_dim + (i >> 6));
  if (tid < 64) { sNb[tid] = __ldg(a.next_b + tid); sAb[tid] = __ldg(a.attr_b + tid); }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  ChainCtx c{sA, &bar, tmem_base_s, 0u, r};
  constexpr uint32_t idesc = make_idesc(128, 64, false, false);
  const uint32_t w_hi = smem_u32(sW);
  float* stage = sStage + warp * (32 * kHalfStageRow);
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t t0 = (tile * 4 + wq) * RPW, t = t0 + lane;
    const int nrows = warp_rows(t0, RPW, a.T);
    const bool live = lane < nrows;
    float v[32];
    half_rows_load(a.E + t0 * 64 + c0, nrows, stage, lane, v);
    float av[32];
    {
      const int64_t id = live ? a.x[t] : 0;
      const float* arow = a.attr_table + id * a.attr_dim;
#pragma unroll
      for (int k = 0; k < 32; ++k) av[k] = (live && k < a.attr_dim) ? __ldg(arow + k) : 0.f;
    }
    if (a.attrt) {       // attribute rows for the attribute_nn weight gradient: 4 planes of 8 columns, two per half
      uint8_t* gt = a.attrt + tile * 16384;
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const int j = 2 * h + jj;
        uint4 hi, lo;
        // (selects keep the indices static)
        float e[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) e[i] = h == 0 ? av[8 * jj + i] : av[16 + 8 * jj + i];
        split8(make_float4(e[0], e[1], e[2], e[3]), make_float4(e[4], e[5], e[6], e[7]), hi, lo);
        *reinterpret_cast<uint4*>(gt + j * 2048 + r * 16) = hi;
        *reinterpret_cast<uint4*>(gt + 8192 + j * 2048 + r * 16) = lo;
      }
    }
    if (live) {
#pragma unroll
      for (int cc = 0; cc < 32; ++cc) v[cc] += sAb[c0 + cc];
#pragma unroll
      for (int k = 0; k < 32; ++k) {        // attribute rows are one-hot + one scalar (main.py:497-512): skip the zeros
        if (k < a.attr_dim && av[k] != 0.f) {
          const float* wt = sWt + k * 64 + c0;
#pragma unroll
          for (int cc = 0; cc < 32; ++cc) v[cc] = fmaf(av[k], wt[cc], v[cc]);
        }
      }
    }
    if (a.V0) half_rows_store(a.V0 + t0 * 64 + c0, nrows, stage, lane, v);
    put_half_row(sA, a.v0t ? a.v0t + tile * (int64_t)kCTile : nullptr, h * 4, r, v);
    float o[32];
    run_stage_half(c, w_hi, idesc, h, o);
    if (live) {
#pragma unroll
      for (int cc = 0; cc < 32; ++cc) o[cc] = tanhf(o[cc] + sNb[c0 + cc]);
    } else {
#pragma unroll
      for (int cc = 0; cc < 32; ++cc) o[cc] = 0.f;
    }
    half_rows_store(a.X + t0 * 64 + c0, nrows, stage, lane, o);
    const float rs = ln_half_row(o, sX[0], sX[1], r, h, pbar);
    if (live) {
      if (h == 0) a.rstd[t] = rs;
    } else {
#pragma unroll
      for (int cc = 0; cc < 32; ++cc) o[cc] = 0.f;
    }
    half_rows_store(a.xhat + t0 * 64 + c0, nrows, stage, lane, o);
    put_half_row_global(a.xt + tile * (int64_t)kXTileBytes, kXHalfBytes, h * 4, r, o);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base_s, 64);
}

// ==========================================================================================
// F3: pff_n1 (two 1x1 convolutions, tanh, dropout, residual) + scorer (LayerNorms, (dyn - static)^2, masked mean)
// ==========================================================================================
struct PffArgs {
  const float* U; const float* xhat; const int64_t* x;
  const uint8_t* w0; const uint8_t* w1; const float* b0; const float* b1;
  ScoreParams p; DropCfg drop;
  float* H1d; float* H2; float* logits;
  uint8_t* ut; uint8_t* h1t;       // U / H1d tiles for the weight-gradient kernel (training) or NULL
  int64_t T;
};

template <int L>
__global__ void __launch_bounds__(kC2Threads, 2) chain_pff_fwd_kernel(const PffArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sW0 = smem + kCTile;
  uint8_t* sW1 = smem + kCTile + kCW;
  float* sStage = reinterpret_cast<float*>(smem + kCTile + 2 * kCW);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float sP[9][64];        // b0, b1, pff_g, pff_b, ln1_g, ln1_b, ln2_g, ln2_b, cls_w
  __shared__ float sCb;
  __shared__ float sX[5][256];
  constexpr int RPW = (32 / L) * L;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r = tid & 127, h = tid >> 7, wq = warp & 3, c0 = h * 32, pbar = 1 + wq;
  const int64_t ntiles = (a.T + 4 * RPW - 1) / (4 * RPW);
  if (warp == 0) tmem_alloc(&tmem_base_s, 64);
  if (tid == 0) { mbar_init(&bar, 1); sCb = __ldg(a.p.cls_b); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  load_weight2(sW0, a.w0);
  load_weight2(sW1, a.w1);
  if (tid < 64) {
    sP[0][tid] = __ldg(a.b0 + tid); sP[1][tid] = __ldg(a.b1 + tid);
    sP[2][tid] = __ldg(a.p.pff_g + tid); sP[3][tid] = __ldg(a.p.pff_b + tid);
    sP[4][tid] = __ldg(a.p.ln1_g + tid); sP[5][tid] = __ldg(a.p.ln1_b + tid);
    sP[6][tid] = __ldg(a.p.ln2_g + tid); sP[7][tid] = __ldg(a.p.ln2_b + tid);
    sP[8][tid] = __ldg(a.p.cls_w + tid);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  ChainCtx c{sA, &bar, tmem_base_s, 0u, r};
  constexpr uint32_t idesc = make_idesc(128, 64, false, false);
  const uint32_t w0_hi = smem_u32(sW0), w1_hi = smem_u32(sW1);
  float* stage = sStage + warp * (32 * kHalfStageRow);
  const bool live_lane = lane < RPW;
  const int g = lane / L, pos = lane - g * L;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t t0 = (tile * 4 + wq) * RPW, t = t0 + lane;
    const int nrows = warp_rows(t0, RPW, a.T);
    const bool live = lane < nrows;
    float u[32];
    half_rows_load(a.U + t0 * 64 + c0, nrows, stage, lane, u);
    put_half_row(sA, a.ut ? a.ut + tile * (int64_t)kCTile : nullptr, h * 4, r, u);
    float hh[32];
    run_stage_half(c, w0_hi, idesc, h, hh);
    if (live) {
#pragma unroll
      for (int cc = 0; cc < 32; cc += 4) {
        float4 v = make_float4(tanhf(hh[cc] + sP[0][c0 + cc]), tanhf(hh[cc + 1] + sP[0][c0 + cc + 1]),
                               tanhf(hh[cc + 2] + sP[0][c0 + cc + 2]), tanhf(hh[cc + 3] + sP[0][c0 + cc + 3]));
        v = drop_apply4(a.drop, (uint64_t)t, (uint32_t)(c0 + cc), v);    // dropout inside pff_n1 (Modules.py:359-360)
        hh[cc] = v.x; hh[cc + 1] = v.y; hh[cc + 2] = v.z; hh[cc + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int cc = 0; cc < 32; ++cc) hh[cc] = 0.f;
    }
    if (a.H1d) half_rows_store(a.H1d + t0 * 64 + c0, nrows, stage, lane, hh);
    put_half_row(sA, a.h1t ? a.h1t + tile * (int64_t)kCTile : nullptr, h * 4, r, hh);
    float o[32];
    run_stage_half(c, w1_hi, idesc, h, o);
#pragma unroll
    for (int cc = 0; cc < 32; ++cc) o[cc] = o[cc] + sP[1][c0 + cc] + u[cc];      // residual (Modules.py:371-372)
    if (a.H2) half_rows_store(a.H2 + t0 * 64 + c0, nrows, stage, lane, o);
    half_rows_load(a.xhat + t0 * 64 + c0, nrows, stage, lane, u);                // u now holds the normalised layer input
    // dead rows run the same exchanges on zeros (the named barriers need both warps of a row)
    const float m = (live && a.x[t] != 0) ? 1.f : 0.f;
    ln_half_row(o, sX[0], sX[1], r, h, pbar);                                    // pff_n1.layer_norm
#pragma unroll
    for (int cc = 0; cc < 32; ++cc) o[cc] = fmaf(o[cc], sP[2][c0 + cc], sP[3][c0 + cc]) * m;
    ln_half_row(o, sX[2], sX[3], r, h, pbar);                                    // Classifier.layer_norm1
    float zp = 0.f;
#pragma unroll
    for (int cc = 0; cc < 32; ++cc) {
      const float D = fmaf(o[cc], sP[4][c0 + cc], sP[5][c0 + cc]);
      const float S = fmaf(u[cc], sP[6][c0 + cc], sP[7][c0 + cc]);               // Classifier.layer_norm2 of the layer input
      const float df = D - S;
      zp = fmaf(df * df, sP[8][c0 + cc], zp);
    }
    float z = pair_sum(zp, sX[4], r, h, pbar);
    z = live ? z + sCb : 0.f;
    // masked mean over the L tokens of the hyperedge (all inside this warp)
    float zs = z * m, ms = m;
#pragma unroll
    for (int s = 1; s < L; ++s) {
      const int src = live_lane ? g * L + (pos + s) % L : lane;
      zs += __shfl_sync(0xffffffffu, z * m, src);
      ms += __shfl_sync(0xffffffffu, m, src);
    }
    if (live && pos == 0 && h == 0) a.logits[t / L] = zs / (ms + 1e-15f);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base_s, 64);
}

// ==========================================================================================
// B1: pff_n1 backward.  dH2 -> dH1pre = (dH2 W1) * tanh' * dropout -> dU = dH1pre W0 + dH2 -> masked / dropout-scaled
// gradient of the attention output.  The dH2 and dH1pre tiles go to global memory for the weight-gradient kernel.
// ==========================================================================================
struct PffBwdArgs {
  const float* dH2; const float* H1d; const int64_t* x;
  const uint8_t* w1mn; const uint8_t* w0mn;
  DropCfg dpff, dattn;
  float* dd;                      // out [T, 64]
  uint8_t* dh2t; uint8_t* dh1t;   // out tiles (32 KB each)
  int64_t T;
};

template <int L>
__global__ void __launch_bounds__(kC2Threads, 2) chain_pff_bwd_kernel(const PffBwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sW1 = smem + kCTile;
  uint8_t* sW0 = smem + kCTile + kCW;
  float* sStage = reinterpret_cast<float*>(smem + kCTile + 2 * kCW);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  constexpr int RPW = (32 / L) * L;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r = tid & 127, h = tid >> 7, wq = warp & 3, c0 = h * 32;
  const int64_t ntiles = (a.T + 4 * RPW - 1) / (4 * RPW);
  if (warp == 0) tmem_alloc(&tmem_base_s, 64);
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  for (int i = tid; i < kCW / 16; i += kC2Threads) {
    reinterpret_cast<uint4*>(sW1)[i] = __ldg(reinterpret_cast<const uint4*>(a.w1mn) + i);
    reinterpret_cast<uint4*>(sW0)[i] = __ldg(reinterpret_cast<const uint4*>(a.w0mn) + i);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  ChainCtx c{sA, &bar, tmem_base_s, 0u, r};
  constexpr uint32_t idesc = make_idesc(128, 64, false, true);      // B = W read as [N = column, K = row], MN-major
  const uint32_t w1_hi = smem_u32(sW1), w0_hi = smem_u32(sW0);
  float* stage = sStage + warp * (32 * kHalfStageRow);
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t t0 = (tile * 4 + wq) * RPW, t = t0 + lane;
    const int nrows = warp_rows(t0, RPW, a.T);
    const bool live = lane < nrows;
    float g[32];
    half_rows_load(a.dH2 + t0 * 64 + c0, nrows, stage, lane, g);
    put_half_row(sA, a.dh2t + tile * (int64_t)kCTile, h * 4, r, g);
    float o[32];
    run_stage_half(c, w1_hi, idesc, h, o);
    half_rows_load(a.H1d + t0 * 64 + c0, nrows, stage, lane, g);
    // gradient through H1d = dropout(tanh(.)): dy * f * (1 - (y / f)^2), f = keep * scale
#pragma unroll
    for (int cc = 0; cc < 32; cc += 4) {
      const float4 f = drop_factor4(a.dpff, (uint64_t)t, (uint32_t)(c0 + cc));
      const float ff[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float hh = ff[j] > 0.f ? g[cc + j] / ff[j] : 0.f;
        o[cc + j] = live ? o[cc + j] * ff[j] * (1.f - hh * hh) : 0.f;
      }
    }
    put_half_row(sA, a.dh1t + tile * (int64_t)kCTile, h * 4, r, o);
    run_stage_half(c, w0_hi, idesc, h, o);
    half_rows_load(a.dH2 + t0 * 64 + c0, nrows, stage, lane, g);      // residual branch (second read: L2 hit)
    const float m = (live && a.x[t] != 0) ? 1.f : 0.f;
#pragma unroll
    for (int cc = 0; cc < 32; cc += 4) {
      const float4 f = drop_factor4(a.dattn, (uint64_t)t, (uint32_t)(c0 + cc));
      o[cc] = (o[cc] + g[cc]) * (f.x * m); o[cc + 1] = (o[cc + 1] + g[cc + 1]) * (f.y * m);
      o[cc + 2] = (o[cc + 2] + g[cc + 2]) * (f.z * m); o[cc + 3] = (o[cc + 3] + g[cc + 3]) * (f.w * m);
    }
    half_rows_store(a.dd + t0 * 64 + c0, nrows, stage, lane, o);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base_s, 64);
}

// ==========================================================================================
// B3: LayerNorm / tanh / next_w backward.  dP = (LNbwd(sum of dxhat partials) + dXs) * (1 - X^2);  dV0 = dP next_w;
// dE = dV0 + beta * dtE * (1 - tanh(E)^2).  dP and dV0 tiles go to global memory for the weight-gradient kernel.
// ==========================================================================================
struct MixBwdArgs {
  const float* dxhat; int nparts; int64_t part_stride;
  const float* dXs; const float* xhat; const float* rstd; const float* X;
  const float* dtE; const float* E; float beta;     // dtE may be NULL
  const uint8_t* wnmn;
  float* dE;                       // out [T, 64]
  uint8_t* dpt; uint8_t* dv0t;     // out tiles
  int64_t T;
};

template <int L>
__global__ void __launch_bounds__(kCThreads) chain_mix_bwd_kernel(const MixBwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sW = smem + kCTile;
  float* sStage = reinterpret_cast<float*>(smem + kCTile + kCW);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  constexpr int RPW = (32 / L) * L;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = (a.T + 4 * RPW - 1) / (4 * RPW);
  if (warp == 0) tmem_alloc(&tmem_base_s, 64);
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  load_weight(sW, a.wnmn);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  ChainCtx c{sA, &bar, tmem_base_s, 0u, tid};
  constexpr uint32_t idesc = make_idesc(128, 64, false, true);
  const uint32_t w_hi = smem_u32(sW);
  float* stage = sStage + warp * kStageWarp;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t t0 = (tile * 4 + warp) * RPW, t = t0 + lane;
    const int nrows = warp_rows(t0, RPW, a.T);
    const bool live = lane < nrows;
    float g[64], r[64];
    warp_load_rows(a.dxhat + t0 * 64, nrows, stage, lane, g);
    for (int p = 1; p < a.nparts; ++p) warp_add_rows(a.dxhat + p * a.part_stride + t0 * 64, nrows, stage, lane, g);
    warp_load_rows(a.xhat + t0 * 64, nrows, stage, lane, r);
    {
      float m1 = 0.f, m2 = 0.f;
#pragma unroll
      for (int cc = 0; cc < 64; ++cc) { m1 += g[cc]; m2 = fmaf(g[cc], r[cc], m2); }
      m1 *= (1.0f / 64); m2 *= (1.0f / 64);
      const float rs = live ? __ldg(a.rstd + t) : 0.f;
#pragma unroll
      for (int cc = 0; cc < 64; ++cc) g[cc] = (g[cc] - m1 - r[cc] * m2) * rs;          // LayerNorm backward
    }
    warp_add_rows(a.dXs + t0 * 64, nrows, stage, lane, g);                              // + static-branch gradient
    warp_load_rows(a.X + t0 * 64, nrows, stage, lane, r);
#pragma unroll
    for (int cc = 0; cc < 64; ++cc) g[cc] = live ? g[cc] * (1.f - r[cc] * r[cc]) : 0.f; // X = tanh(P)  (Modules.py:270)
    put_row(sA, a.dpt + tile * (int64_t)kCTile, tid, g);
    float o[64];
    run_stage(c, w_hi, idesc, o);
    if (!live) {
#pragma unroll
      for (int cc = 0; cc < 64; ++cc) o[cc] = 0.f;
    }
    put_row_global(a.dv0t + tile * (int64_t)kCTile, 16384, tid, o);
    if (a.dtE) {
      warp_load_rows(a.E + t0 * 64, nrows, stage, lane, r);
      warp_load_rows(a.dtE + t0 * 64, nrows, stage, lane, g);
#pragma unroll
      for (int cc = 0; cc < 64; ++cc) {
        const float te = tanhf(r[cc]);
        o[cc] = fmaf(g[cc] * (1.f - te * te), a.beta, o[cc]);
      }
    }
    warp_store_rows(a.dE + t0 * 64, nrows, stage, lane, o);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base_s, 64);
}

// ==========================================================================================
// W: two stacked 64-wide weight gradients per launch.  D[128, N] += [A1 | A2]^T . [B1 | B2 | ones]  over the tokens of
// this CTA's tiles (TMEM-resident accumulator): rows 0..63 x columns of B1 = A1^T B1, rows 64..127 x columns of B2 =
// A2^T B2, the ones column = the two bias gradients.  Pure bulk copies + MMAs: every operand arrives pre-split.
// ==========================================================================================
constexpr int kWThreadsP = 192;      // producer warp, MMA warp, 4 readout warps
struct WgradPairArgs {
  const uint8_t* a1; const uint8_t* a2;      // 32 KB tiles
  const uint8_t* b1;                          // 32 KB tiles
  const uint8_t* b2; int b2_planes; int b2_tile_bytes;     // 8 planes (32 KB tiles) or 4 planes (16 KB tiles)
  int64_t ntiles;
  float* part;                                // [grid][128][N]
};

__global__ void __launch_bounds__(kWThreadsP, 1) wgrad_pair_kernel(const WgradPairArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int nbp = 8 + a.b2_planes + 2;                 // B planes per half: B1 | B2 | ones | zeros
  const int b_half = nbp * 2048;
  uint8_t* sAp = smem;                                  // two buffers of: hi A1 planes 0-7, A2 planes 8-15 (32 KB) | lo (32 KB)
  uint8_t* sB = smem + 2 * 65536;                       // hi [nbp planes] | lo [nbp planes]
  // the A tiles are double buffered (the next tile's 64 KB land under this tile's MMAs); the B tiles do not fit twice
  __shared__ uint64_t a_full[2], a_empty[2], full, empty, done;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = nbp * 8;
  if (warp == 0) tmem_alloc(&tmem_base_s, 256);
  if (tid == 32) {
    mbar_init(&full, 1); mbar_init(&empty, 1); mbar_init(&done, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // constant planes: ones column (bf16 1.0 in column 0 of the plane, hi half only) and zeros
  for (int i = tid; i < 2 * 128; i += kWThreadsP) {
    const int pl = i >> 7, row = i & 127;
    *reinterpret_cast<uint4*>(sB + (8 + a.b2_planes + pl) * 2048 + row * 16) = make_uint4(pl == 0 ? 0x00003F80u : 0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(sB + b_half + (8 + a.b2_planes + pl) * 2048 + row * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int64_t my_tiles = (a.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  const int b2_half = a.b2_planes * 2048;

  if (warp == 0) {
    if (lane == 0) {
      for (int64_t k = 0; k < my_tiles; ++k) {
        const int64_t tile = blockIdx.x + k * gridDim.x;
        const int ab = (int)(k & 1);
        uint8_t* sa = sAp + ab * 65536;
        mbar_wait_backoff(&a_empty[ab], (uint32_t)((k >> 1) & 1) ^ 1u);
        mbar_expect_tx(&a_full[ab], 4 * 16384);
        const uint8_t* s1 = a.a1 + tile * (int64_t)kCTile;
        const uint8_t* s2 = a.a2 + tile * (int64_t)kCTile;
        bulk_g2s(sa, s1, 16384, &a_full[ab]);
        bulk_g2s(sa + 16384, s2, 16384, &a_full[ab]);
        bulk_g2s(sa + 32768, s1 + 16384, 16384, &a_full[ab]);
        bulk_g2s(sa + 49152, s2 + 16384, 16384, &a_full[ab]);
        mbar_wait_backoff(&empty, (uint32_t)(k & 1) ^ 1u);
        mbar_expect_tx(&full, 2 * 16384 + 2 * b2_half);
        const uint8_t* t1 = a.b1 + tile * (int64_t)kCTile;
        const uint8_t* t2 = a.b2 + tile * (int64_t)a.b2_tile_bytes;
        bulk_g2s(sB, t1, 16384, &full);
        bulk_g2s(sB + b_half, t1 + 16384, 16384, &full);
        bulk_g2s(sB + 16384, t2, b2_half, &full);
        bulk_g2s(sB + b_half + 16384, t2 + b2_half, b2_half, &full);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(128, N, true, true);
      const uint32_t bh = smem_u32(sB), bl = bh + b_half;
      for (int64_t k = 0; k < my_tiles; ++k) {
        const int ab = (int)(k & 1);
        const uint32_t ah = smem_u32(sAp + ab * 65536), al = ah + 32768;
        mbar_wait_backoff(&a_full[ab], (uint32_t)((k >> 1) & 1));
        mbar_wait_backoff(&full, (uint32_t)(k & 1));
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)      // K = 16 tokens per step; both operands MN-major (LBO 128: next 8 tokens, SBO 2048: next plane)
          umma_x3s(tmem_base, ah + ks * 256, al + ks * 256, bh + ks * 256, bl + ks * 256, 128, 2048, 128, 2048, idesc,
                   k == 0 && ks == 0);
        umma_commit(&a_empty[ab]);
        umma_commit(&empty);
      }
      umma_commit(&done);
    }
  } else {
    const int q = warp & 3;
    if (my_tiles > 0) mbar_wait(&done, 0);
    tc_fence_after();
    const int row = q * 32 + lane;
    float* dst = a.part + ((int64_t)blockIdx.x * 128 + row) * N;
    for (int c0 = 0; c0 < N; c0 += 16) {
      float v[16];
      if (my_tiles > 0) {
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + c0, v);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(dst + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// sum the per-CTA partials and add the four blocks into the gradient buffers
struct WgradPairOut {
  float* w1; int ld1; int n1; float* bias1;       // rows 0..63,  columns [0, n1)
  float* w2; int ld2; int n2; float* bias2;       // rows 64..127, columns [64, 64 + n2)
};
// 8 lanes per output element: each sums every 8th split-K partial (8 independent load streams instead of one serial
// chain of ~148 dependent adds), then a fixed xor-shuffle tree -> deterministic
__global__ void wgrad_pair_reduce_kernel(const float* __restrict__ part, int nparts, int N, int ones_col, const WgradPairOut o) {
  const int total = 128 * N;
  const int sub = threadIdx.x & 7;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 3; i < total; i += (gridDim.x * blockDim.x) >> 3) {   // warp-uniform trip count
    const int row = i / N, col = i - row * N;
    float* dst = nullptr;
    if (row < 64) {
      if (col < o.n1) dst = o.w1 + row * o.ld1 + col;
      else if (col == ones_col) dst = o.bias1 + row;
    } else {
      if (col >= 64 && col < 64 + o.n2) dst = o.w2 + (row - 64) * o.ld2 + (col - 64);
      else if (col == ones_col) dst = o.bias2 + (row - 64);
    }
    float s = 0.f;
    if (dst)
      for (int p = sub; p < nparts; p += 8) s += part[(int64_t)p * total + i];
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    if (dst && sub == 0) *dst += s;
  }
}

template <typename K>
int set_smem_attr_c(K kernel, int bytes) {
  return check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes), "cudaFuncSetAttribute");
}
// one tile per CTA up to 4 CTAs per SM's worth: queued CTAs start as soon as a resident one drains, which hides the
// load -> MMA -> store latency chain of a tile better than a short persistent loop (measured)
template <typename K>
unsigned chain_grid(K, int, int64_t ntiles) {
  const int64_t cap = (int64_t)kSMs * 8;
  return (unsigned)(ntiles < cap ? ntiles : cap);
}

template <int L>
int launch_mix_L(const MixArgs& a, cudaStream_t s) {
  const int smem = kCTile + kCW + kHalfStageBytes + a.attr_dim * 64 * 4;
  static int set_for = 0;
  if (set_for < smem) { if (int rc = set_smem_attr_c(chain_mix_fwd_kernel<L>, smem)) return rc; set_for = smem; }
  chain_mix_fwd_kernel<L><<<chain_grid(chain_mix_fwd_kernel<L>, smem, num_atiles(a.T, L)), kC2Threads, smem, s>>>(a);
  MATCHA_CHECK_LAUNCH("chain_mix_fwd");
  return MATCHA_OK;
}
template <int L>
int launch_pff_L(const PffArgs& a, cudaStream_t s) {
  constexpr int smem = kCTile + 2 * kCW + kHalfStageBytes;
  static bool once = false;
  if (!once) { if (int rc = set_smem_attr_c(chain_pff_fwd_kernel<L>, smem)) return rc; once = true; }
  chain_pff_fwd_kernel<L><<<chain_grid(chain_pff_fwd_kernel<L>, smem, num_atiles(a.T, L)), kC2Threads, smem, s>>>(a);
  MATCHA_CHECK_LAUNCH("chain_pff_fwd");
  return MATCHA_OK;
}


template <int L>
int launch_pff_bwd_L(const PffBwdArgs& a, cudaStream_t s) {
  constexpr int smem = kCTile + 2 * kCW + kHalfStageBytes;
  static bool once = false;
  if (!once) { if (int rc = set_smem_attr_c(chain_pff_bwd_kernel<L>, smem)) return rc; once = true; }
  chain_pff_bwd_kernel<L><<<chain_grid(chain_pff_bwd_kernel<L>, smem, num_atiles(a.T, L)), kC2Threads, smem, s>>>(a);
  MATCHA_CHECK_LAUNCH("chain_pff_bwd");
  return MATCHA_OK;
}
template <int L>
int launch_mix_bwd_L(const MixBwdArgs& a, cudaStream_t s) {
  constexpr int smem = kCTile + kCW + kStageBytes;
  static bool once = false;
  if (!once) { if (int rc = set_smem_attr_c(chain_mix_bwd_kernel<L>, smem)) return rc; once = true; }
  chain_mix_bwd_kernel<L><<<chain_grid(chain_mix_bwd_kernel<L>, smem, num_atiles(a.T, L)), kCThreads, smem, s>>>(a);
  MATCHA_CHECK_LAUNCH("chain_mix_bwd");
  return MATCHA_OK;
}

}  // namespace

int64_t wgrad_pair_scratch_floats() { return (int64_t)kSMs * 128 * 144; }

// [A1 | A2]^T [B1 | B2 | 1]: see wgrad_pair_kernel.  b2_planes = 8 (64 columns) or 4 (32 columns, attribute rows)
int launch_wgrad_pair(const uint8_t* a1, const uint8_t* a2, const uint8_t* b1, const uint8_t* b2, int b2_planes, int64_t ntiles,
                      float* scratch, float* w1, int ld1, int n1, float* bias1, float* w2, int ld2, int n2, float* bias2,
                      cudaStream_t s) {
  if (ntiles <= 0) return MATCHA_OK;
  const int nbp = 8 + b2_planes + 2, N = nbp * 8;
  const int smem = 2 * 65536 + 2 * nbp * 2048;
  static int set_for = 0;
  if (set_for < smem) { if (int rc = set_smem_attr_c(wgrad_pair_kernel, smem)) return rc; set_for = smem; }
  const unsigned grid = (unsigned)(ntiles < kSMs ? ntiles : kSMs);
  WgradPairArgs a{a1, a2, b1, b2, b2_planes, b2_planes * 4096, ntiles, scratch};
  wgrad_pair_kernel<<<grid, kWThreadsP, smem, s>>>(a);
  MATCHA_CHECK_LAUNCH("wgrad_pair");
  WgradPairOut o{w1, ld1, n1, bias1, w2, ld2, n2, bias2};
  wgrad_pair_reduce_kernel<<<(128 * N * 8 + 255) / 256, 256, 0, s>>>(scratch, (int)grid, N, (8 + b2_planes) * 8, o);
  MATCHA_CHECK_LAUNCH("wgrad_pair_reduce");
  return MATCHA_OK;
}

int launch_chain_pff_bwd(const float* dH2, const float* H1d, const int64_t* x, const void* w1_mn, const void* w0_mn,
                         DropCfg dpff, DropCfg dattn, float* dd, uint8_t* dh2_tiles, uint8_t* dh1_tiles, int64_t B, int L,
                         cudaStream_t s) {
  if (B <= 0) return MATCHA_OK;
  PffBwdArgs a{dH2, H1d, x, reinterpret_cast<const uint8_t*>(w1_mn), reinterpret_cast<const uint8_t*>(w0_mn), dpff, dattn, dd,
               dh2_tiles, dh1_tiles, B * L};
  switch (L) {
    case 2: return launch_pff_bwd_L<2>(a, s);
    case 3: return launch_pff_bwd_L<3>(a, s);
    case 4: return launch_pff_bwd_L<4>(a, s);
    case 5: return launch_pff_bwd_L<5>(a, s);
    case 6: return launch_pff_bwd_L<6>(a, s);
    default: set_error("chain_pff_bwd: padded width L=%d unsupported (2..6)", L); return MATCHA_ERR_ARG;
  }
}

int launch_chain_mix_bwd(const float* dxhat, int nparts, int64_t part_stride, const float* dXs, const float* xhat,
                         const float* rstd, const float* X, const float* dtE, const float* E, float beta, const void* wn_mn,
                         float* dE, uint8_t* dp_tiles, uint8_t* dv0_tiles, int64_t B, int L, cudaStream_t s) {
  if (B <= 0) return MATCHA_OK;
  MixBwdArgs a{dxhat, nparts, part_stride, dXs, xhat, rstd, X, dtE, E, beta, reinterpret_cast<const uint8_t*>(wn_mn), dE,
               dp_tiles, dv0_tiles, B * L};
  switch (L) {
    case 2: return launch_mix_bwd_L<2>(a, s);
    case 3: return launch_mix_bwd_L<3>(a, s);
    case 4: return launch_mix_bwd_L<4>(a, s);
    case 5: return launch_mix_bwd_L<5>(a, s);
    case 6: return launch_mix_bwd_L<6>(a, s);
    default: set_error("chain_mix_bwd: padded width L=%d unsupported (2..6)", L); return MATCHA_ERR_ARG;
  }
}

int launch_split_w64x3(const float* W0, const float* W1, const float* W2, void* k0, void* mn0, void* k1, void* mn1, void* k2,
                       void* mn2, cudaStream_t s) {
  SplitW3 a;
  a.W[0] = W0; a.W[1] = W1; a.W[2] = W2;
  a.out_k[0] = reinterpret_cast<uint8_t*>(k0); a.out_k[1] = reinterpret_cast<uint8_t*>(k1); a.out_k[2] = reinterpret_cast<uint8_t*>(k2);
  a.out_mn[0] = reinterpret_cast<uint8_t*>(mn0); a.out_mn[1] = reinterpret_cast<uint8_t*>(mn1); a.out_mn[2] = reinterpret_cast<uint8_t*>(mn2);
  split_w64_kernel<<<6, 256, 0, s>>>(a);
  MATCHA_CHECK_LAUNCH("split_w64");
  return MATCHA_OK;
}

int launch_chain_mix_fwd(const float* E, const int64_t* x, const float* attr_table, int attr_dim, const float* attr_w,
                         const float* attr_b, const void* w_next_k, const float* next_b, float* V0, float* X, float* xhat,
                         float* rstd, uint8_t* xhat_tiles, uint8_t* v0_tiles, uint8_t* attr_tiles, int64_t B, int L,
                         cudaStream_t s) {
  if (B <= 0) return MATCHA_OK;
  if (attr_dim > 32) { set_error("chain_mix_fwd: attribute width %d > 32", attr_dim); return MATCHA_ERR_UNSUPPORTED; }
  MixArgs a{E, x, attr_table, attr_dim, attr_w, attr_b, reinterpret_cast<const uint8_t*>(w_next_k), next_b, V0, X, xhat, rstd,
            xhat_tiles, v0_tiles, attr_tiles, B * L};
  switch (L) {
    case 2: return launch_mix_L<2>(a, s);
    case 3: return launch_mix_L<3>(a, s);
    case 4: return launch_mix_L<4>(a, s);
    case 5: return launch_mix_L<5>(a, s);
    case 6: return launch_mix_L<6>(a, s);
    default: set_error("chain_mix_fwd: padded width L=%d unsupported (2..6)", L); return MATCHA_ERR_ARG;
  }
}

int launch_chain_pff_fwd(const float* U, const float* xhat, const int64_t* x, const void* w0_k, const void* w1_k,
                         const float* b0, const float* b1, ScoreParams p, DropCfg drop, float* H1d, float* H2, float* logits,
                         uint8_t* u_tiles, uint8_t* h1_tiles, int64_t B, int L, cudaStream_t s) {
  if (B <= 0) return MATCHA_OK;
  PffArgs a{U, xhat, x, reinterpret_cast<const uint8_t*>(w0_k), reinterpret_cast<const uint8_t*>(w1_k), b0, b1, p, drop,
            H1d, H2, logits, u_tiles, h1_tiles, B * L};
  switch (L) {
    case 2: return launch_pff_L<2>(a, s);
    case 3: return launch_pff_L<3>(a, s);
    case 4: return launch_pff_L<4>(a, s);
    case 5: return launch_pff_L<5>(a, s);
    case 6: return launch_pff_L<6>(a, s);
    default: set_error("chain_pff_fwd: padded width L=%d unsupported (2..6)", L); return MATCHA_ERR_ARG;
  }
}

}  // namespace matcha
