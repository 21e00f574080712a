// tcgen05 (5th-gen tensor core) contraction kernels for the three large products of the hyperedge
// pipeline -- the QKG projection, its data gradient and its weight gradient (B*L tokens x 64 x 1536).
//
// Precision: fp32 operands are split on the fly into bf16 hi + bf16 lo and three MMAs are issued
// (hi*hi + hi*lo + lo*hi) into an fp32 TMEM accumulator ("bf16x3"): ~2^-16 relative error per product,
// which keeps the whole pipeline inside the rtol 1e-4 contract, at one third of the bf16 MMA rate.
//
// Operand staging is done by the CTA's own threads (global fp32 -> registers -> split -> st.shared in
// the canonical no-swizzle UMMA layouts), because the fp32 -> 2 x bf16 split has to touch every element
// anyway; `fence.proxy.async` publishes the tiles to the tensor-core proxy.  One elected thread issues
// tcgen05.mma; completion is tracked with tcgen05.commit -> mbarrier; accumulators are read back with
// tcgen05.ld (32 lanes x 32 columns per warp).  Canonical layouts (cute/arch/mma_sm100_desc.hpp):
//   K-major  (k contiguous):  [k/8][row][8 elems]      LBO = rows * 16 B (next 8 k),  SBO = 128 B (next 8 rows)
//   MN-major (mn contiguous): [k/8][mn/8][k%8][8 elems] LBO = mn * 16 B  (next 8 k),  SBO = 128 B (next 8 mn)
#include "tc_common.cuh"

namespace matcha {
namespace {

// ==========================================================================================
// NT, K = 64:  C[M, N] = A[M, 64] . B[N, 64]^T (+ bias)        (QKG projection; N % BN == 0)
// ==========================================================================================
template <int BN>
__global__ void __launch_bounds__(kThreads) tc_nt_k64_kernel(const float* __restrict__ A, int64_t lda,
                                                             const float* __restrict__ B, int64_t ldb,
                                                             float* __restrict__ C, int64_t ldc,
                                                             const float* __restrict__ bias, int64_t M, int64_t N) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sAh = smem;                 // [8][128][16 B]
  uint8_t* sAl = smem + 16384;
  uint8_t* sBh = smem + 32768;         // [8][BN][16 B]
  uint8_t* sBl = sBh + BN * 128;
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t m0 = (int64_t)blockIdx.x * 128;
  if (warp == 0) tmem_alloc(&tmem_base_s, BN);
  if (tid == 0) {
    mbar_init(&mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {  // A tile: thread = row
    const int64_t r = m0 + tid;
    const bool ok = r < M;
    const float* src = A + (ok ? r : 0) * lda;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      uint4 hi, lo;
      split8(ldg4z(src + q * 8, ok), ldg4z(src + q * 8 + 4, ok), hi, lo);
      sts16(sAh + q * 2048 + tid * 16, hi);
      sts16(sAl + q * 2048 + tid * 16, lo);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  constexpr uint32_t idesc = make_idesc(128, BN, false, false);
  uint32_t phase = 0;
  for (int64_t n0 = 0; n0 < N; n0 += BN) {
    // B chunk: BN rows x 8 k-chunks
#pragma unroll 4
    for (int u = tid; u < BN * 8; u += kThreads) {
      const int row = u % BN, q = u / BN;
      const float* src = B + (n0 + row) * ldb + q * 8;
      uint4 hi, lo;
      split8(ldg4z(src, true), ldg4z(src + 4, true), hi, lo);
      sts16(sBh + q * (BN * 16) + row * 16, hi);
      sts16(sBl + q * (BN * 16) + row * 16, lo);
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        umma_x3(tmem_base, smem_u32(sAh) + ks * 4096, smem_u32(sAl) + ks * 4096, smem_u32(sBh) + ks * (BN * 32),
                smem_u32(sBl) + ks * (BN * 32), 2048, BN * 16, idesc, ks == 0);
      umma_commit(&mbar);
    }
    mbar_wait(&mbar, phase);
    phase ^= 1;
    tc_fence_after();
    const int64_t r = m0 + warp * 32 + lane;
#pragma unroll 1
    for (int cc = 0; cc < BN / 32; ++cc) {
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + cc * 32, v);
      if (r < M) {
        float* dst = C + r * ldc + n0 + cc * 32;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          if (bias) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + n0 + cc * 32 + j));
            o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
          }
          *reinterpret_cast<float4*>(dst + j) = o;
        }
      }
    }
    tc_fence_before();
    __syncthreads();
  }
  if (warp == 0) tmem_dealloc(tmem_base, BN);
}

// ==========================================================================================
// NN, N = 64:  C[M, 64] = A[M, K] . B[K, 64]        (data gradient of the QKG projection; K % 64 == 0)
// ==========================================================================================
__global__ void __launch_bounds__(kThreads) tc_nn_n64_kernel(const float* __restrict__ A, int64_t lda,
                                                             const float* __restrict__ B, int64_t ldb,
                                                             float* __restrict__ C, int64_t ldc, int64_t M, int64_t K) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int kBuf = 16 * 2064 + 16384;   // per buffer: A hi | A lo (8 padded planes each) | B hi 8 KB | B lo 8 KB
  __shared__ uint64_t mbar[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t m0 = (int64_t)blockIdx.x * 128;
  if (warp == 0) tmem_alloc(&tmem_base_s, 64);
  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  constexpr uint32_t idesc = make_idesc(128, 64, false, true);
  constexpr int kPlane = 2064;                     // padded K-major plane of the A tile (128 rows x 16 B + 16)
  uint32_t phase[2] = {0, 0};
  const int nk = (int)(K / 64);
  const int aq = tid & 7, ar0 = tid >> 3;          // A units: row = ar0 + 16 i, k-chunk aq
  float4 ra[16], rb[8], na[16], nb[8];
  auto load_chunk = [&](int kc, float4 (&xa)[16], float4 (&xb)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t r = m0 + ar0 + 16 * i;
      const bool ok = r < M;
      const float* src = A + (ok ? r : 0) * lda + kc * 64 + aq * 8;
      xa[2 * i] = ldg4z(src, ok);
      xa[2 * i + 1] = ldg4z(src + 4, ok);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {   // B chunk [64 k][64 n]: unit = (k, n-group of 8)
      const int u = i * kThreads + tid, g = u & 7, k = u >> 3;
      const float* src = B + (int64_t)(kc * 64 + k) * ldb + g * 8;
      xb[2 * i] = ldg4z(src, true);
      xb[2 * i + 1] = ldg4z(src + 4, true);
    }
  };
  load_chunk(0, ra, rb);
  for (int kc = 0; kc < nk; ++kc) {
    const int buf = kc & 1;
    uint8_t* sAh = smem + buf * kBuf;
    uint8_t* sAl = sAh + 8 * kPlane;
    uint8_t* sBh = sAh + 16 * kPlane;
    uint8_t* sBl = sBh + 8192;
    if (kc + 1 < nk) load_chunk(kc + 1, na, nb);   // in flight while this chunk is split and multiplied
    if (kc >= 2) { mbar_wait(&mbar[buf], phase[buf]); phase[buf] ^= 1; }   // MMAs of chunk kc-2 released this buffer
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint4 hi, lo;
      split8(ra[2 * i], ra[2 * i + 1], hi, lo);
      const int off = aq * kPlane + (ar0 + 16 * i) * 16;
      sts16(sAh + off, hi);
      sts16(sAl + off, lo);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int u = i * kThreads + tid, g = u & 7, k = u >> 3;
      uint4 hi, lo;
      split8(rb[2 * i], rb[2 * i + 1], hi, lo);
      const int off = (k >> 3) * 1024 + g * 128 + (k & 7) * 16;     // MN-major: [k/8][n/8][k%8][8]
      sts16(sBh + off, hi);
      sts16(sBl + off, lo);
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        umma_x3(tmem_base, smem_u32(sAh) + ks * 2 * kPlane, smem_u32(sAl) + ks * 2 * kPlane, smem_u32(sBh) + ks * 2048,
                smem_u32(sBl) + ks * 2048, kPlane, 1024, idesc, kc == 0 && ks == 0);
      umma_commit(&mbar[buf]);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) ra[i] = na[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) rb[i] = nb[i];
  }
  {  // tcgen05 ops of one thread complete in order: the last commit covers everything
    const int buf = (nk - 1) & 1;
    mbar_wait(&mbar[buf], phase[buf]);
  }
  tc_fence_after();
  const int64_t ro = m0 + warp * 32 + lane;
#pragma unroll 1
  for (int cc = 0; cc < 2; ++cc) {
    float v[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + cc * 32, v);
    if (ro < M) {
      float* dst = C + ro * ldc + cc * 32;
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 64);
}

// ==========================================================================================
// TN, N = 64:  part[s][M, 64] = A[Ks, M]^T . B[Ks, 64] over the token range of split s
//              (weight gradient of the QKG projection; M % 512 == 0; also column sums of A)
// ==========================================================================================
__global__ void __launch_bounds__(kThreads) tc_tn_n64_kernel(const float* __restrict__ A, int64_t lda,
                                                             const float* __restrict__ B, int64_t ldb, int64_t M,
                                                             int64_t K, int64_t k_per_split, float* __restrict__ part,
                                                             float* __restrict__ part_colsum) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int kBuf = 36864;   // per buffer: A hi 16 KB | A lo 16 KB | B hi 2 KB | B lo 2 KB   (16 tokens x 512 / 64 columns)
  __shared__ uint64_t mbar[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_cs[kThreads][8];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mg = blockIdx.x;                       // group of 512 output rows (columns of A)
  const int split = blockIdx.y;
  const int64_t k_begin = (int64_t)split * k_per_split;
  const int64_t k_end = (k_begin + k_per_split < K) ? k_begin + k_per_split : K;
  if (warp == 0) tmem_alloc(&tmem_base_s, 256);
  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  constexpr uint32_t idesc = make_idesc(128, 64, true, true);
  const int ga = tid & 63, ta = tid >> 6;          // A units: column group ga (8 columns), tokens ta + 2 i
  const int gb = tid & 7, tb = tid >> 3;           // B unit : column group gb, token tb
  const float* acol = A + (int64_t)mg * 512 + ga * 8;
  float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  uint32_t phase[2] = {0, 0};
  const int nchunk = (int)((k_end - k_begin + 15) / 16);
  float4 ra[16], rb[2], na[16], nb[2];
  auto load_chunk = [&](int ch, float4 (&xa)[16], float4 (&xb)[2]) {
    const int64_t kb = k_begin + (int64_t)ch * 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t k = kb + ta + 2 * i;
      const bool ok = k < k_end;
      const float* src = acol + (ok ? k : 0) * lda;
      xa[2 * i] = ldg4z(src, ok);
      xa[2 * i + 1] = ldg4z(src + 4, ok);
    }
    const int64_t k = kb + tb;
    const bool ok = k < k_end;
    const float* src = B + (ok ? k : 0) * ldb + gb * 8;
    xb[0] = ldg4z(src, ok);
    xb[1] = ldg4z(src + 4, ok);
  };
  if (nchunk > 0) load_chunk(0, ra, rb);
  for (int ch = 0; ch < nchunk; ++ch) {
    const int buf = ch & 1;
    uint8_t* sAh = smem + buf * kBuf;
    uint8_t* sAl = sAh + 16384;
    uint8_t* sBh = sAh + 32768;
    uint8_t* sBl = sBh + 2048;
    if (ch + 1 < nchunk) load_chunk(ch + 1, na, nb);   // next chunk's loads fly while this one is split and multiplied
    if (ch >= 2) { mbar_wait(&mbar[buf], phase[buf]); phase[buf] ^= 1; }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = ta + 2 * i;                                       // token within the chunk
      uint4 hi, lo;
      split8(ra[2 * i], ra[2 * i + 1], hi, lo);
      const int off = (k >> 3) * 8192 + ga * 128 + (k & 7) * 16;      // MN-major: [k/8][m/8][k%8][8], 512 columns
      sts16(sAh + off, hi);
      sts16(sAl + off, lo);
      cs[0] += ra[2 * i].x; cs[1] += ra[2 * i].y; cs[2] += ra[2 * i].z; cs[3] += ra[2 * i].w;
      cs[4] += ra[2 * i + 1].x; cs[5] += ra[2 * i + 1].y; cs[6] += ra[2 * i + 1].z; cs[7] += ra[2 * i + 1].w;
    }
    {
      uint4 hi, lo;
      split8(rb[0], rb[1], hi, lo);
      const int off = (tb >> 3) * 1024 + gb * 128 + (tb & 7) * 16;
      sts16(sBh + off, hi);
      sts16(sBl + off, lo);
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int j = 0; j < 4; ++j)     // four 128-row output tiles share the B operand
        umma_x3(tmem_base + j * 64, smem_u32(sAh) + j * 2048, smem_u32(sAl) + j * 2048, smem_u32(sBh), smem_u32(sBl), 8192,
                1024, idesc, ch == 0);
      umma_commit(&mbar[buf]);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) ra[i] = na[i];
    rb[0] = nb[0]; rb[1] = nb[1];
  }
  if (nchunk > 0) {
    const int buf = (nchunk - 1) & 1;
    mbar_wait(&mbar[buf], phase[buf]);
  }
  tc_fence_after();
  float* out = part + ((int64_t)split * M + (int64_t)mg * 512) * 64;
#pragma unroll 1
  for (int j = 0; j < 4; ++j) {
#pragma unroll 1
    for (int cc = 0; cc < 2; ++cc) {
      float v[32];
      if (nchunk > 0) {
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + j * 64 + cc * 32, v);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0.f;
      }
      float* dst = out + (int64_t)(j * 128 + warp * 32 + lane) * 64 + cc * 32;
#pragma unroll
      for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    }
  }
  if (part_colsum) {
#pragma unroll
    for (int e = 0; e < 8; ++e) s_cs[tid][e] = cs[e];
    __syncthreads();
    if (tid < 64) {
#pragma unroll
      for (int e = 0; e < 8; ++e)
        part_colsum[(int64_t)split * M + (int64_t)mg * 512 + tid * 8 + e] = s_cs[tid][e] + s_cs[tid + 64][e];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

__global__ void tn_reduce_kernel(const float* __restrict__ part, const float* __restrict__ part_colsum, int splits,
                                 int64_t M, float* __restrict__ C, int64_t ldc, float* __restrict__ colsum,
                                 int64_t colsum_n, float scale) {
  const int64_t total = M * 64;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total + M; i += (int64_t)gridDim.x * blockDim.x) {
    if (i < total) {
      float s = 0.f;
      for (int sp = 0; sp < splits; ++sp) s += part[(int64_t)sp * total + i];
      C[(i >> 6) * ldc + (i & 63)] += s * scale;
    } else if (colsum && part_colsum) {
      const int64_t m = i - total;
      if (m < colsum_n) {
        float s = 0.f;
        for (int sp = 0; sp < splits; ++sp) s += part_colsum[(int64_t)sp * M + m];
        colsum[m] += s * scale;
      }
    }
  }
}

// ==========================================================================================
// v2 of the NT, K = 64 product (the QKG projection), warp-specialised and pipelined.
//   Transposed tile: D[128 features, 128 tokens] = Wsplit[chunk](128 x 64) . xhat_tile(128 x 64)^T, so a TMEM
//   lane is an output FEATURE: the bias is one register per thread and every STG.32 of a warp writes 32
//   consecutive floats of one token row (one full 128-byte line).
//   warp 0: bulk-copies pre-split weight chunks (bf16 hi | lo, canonical K-major, 32 KB) into a 3-stage ring
//   warp 1: issues tcgen05.mma (3 bf16 passes x 4 K-steps per chunk) into a 4-deep ring of TMEM accumulators
//   warps 2-9: split + stage the token tile (double buffered), then drain accumulators (LDTM -> +bias -> STG)
// ==========================================================================================
constexpr int kV2Threads = 320;
constexpr int kV2WStages = 3, kV2AccStages = 4;
constexpr int kV2WBytes = 32768;                 // one weight chunk: hi 16 KB | lo 16 KB
constexpr int kV2XPlane = 2064;                  // 128 rows x 16 B + 16 B pad: conflict-free staging stores
constexpr int kV2XHalf = 8 * kV2XPlane;          // hi (or lo) part of one token tile
constexpr int kV2XBytes = 2 * kV2XHalf;          // one token tile: hi | lo
constexpr int kV2Smem = kV2WStages * kV2WBytes + 2 * kV2XBytes;   // 160 KB

// fp32 [N, 64] row-major  ->  per 128-row chunk: bf16 hi [8][128][8] | bf16 lo [8][128][8]   (N % 128 == 0)
__global__ void split_weights_k64_kernel(const float* __restrict__ W, int64_t ldw, int64_t N, uint8_t* __restrict__ out) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // unit = (row, k-chunk of 8)
  if (u >= N * 8) return;
  const int64_t row = u >> 3;
  const int q = (int)(u & 7);
  const float* src = W + row * ldw + q * 8;
  uint4 hi, lo;
  split8(__ldg(reinterpret_cast<const float4*>(src)), __ldg(reinterpret_cast<const float4*>(src + 4)), hi, lo);
  uint8_t* chunk = out + (row >> 7) * kV2WBytes;
  const int r = (int)(row & 127);
  *reinterpret_cast<uint4*>(chunk + q * 2048 + r * 16) = hi;
  *reinterpret_cast<uint4*>(chunk + 16384 + q * 2048 + r * 16) = lo;
}

__global__ void __launch_bounds__(kV2Threads, 1) tc_qkg_fwd_v2_kernel(const float* __restrict__ A, int64_t lda,
                                                                      const uint8_t* __restrict__ Wsplit,
                                                                      const float* __restrict__ bias, float* __restrict__ C,
                                                                      int64_t ldc, int64_t M, int N) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sW = smem;                                   // [kV2WStages][32 KB]
  uint8_t* sX = smem + kV2WStages * kV2WBytes;          // [2][32 KB]
  __shared__ uint64_t w_full[kV2WStages], w_empty[kV2WStages], x_full[2], x_empty[2], acc_full[kV2AccStages],
      acc_empty[kV2AccStages];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nchunk = N / 128;
  const int64_t ntiles = (M + 127) / 128;
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 32) {
    for (int i = 0; i < kV2WStages; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&x_full[i], 256); mbar_init(&x_empty[i], 1); }
    for (int i = 0; i < kV2AccStages; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ---------------- weight-chunk producer ----------------
    if (lane == 0) {
      int ws = 0; uint32_t wp = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int fc = 0; fc < nchunk; ++fc) {
          mbar_wait(&w_empty[ws], wp ^ 1);
          mbar_expect_tx(&w_full[ws], kV2WBytes);
          bulk_g2s(sW + ws * kV2WBytes, Wsplit + (int64_t)fc * kV2WBytes, kV2WBytes, &w_full[ws]);
          if (++ws == kV2WStages) { ws = 0; wp ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(128, 128, false, false);
      int ws = 0, as = 0, xs = 0; uint32_t wp = 0, ap = 0, xp = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        mbar_wait(&x_full[xs], xp);
        const uint32_t xh = smem_u32(sX + xs * kV2XBytes), xl = xh + kV2XHalf;
        for (int fc = 0; fc < nchunk; ++fc) {
          mbar_wait(&w_full[ws], wp);
          mbar_wait(&acc_empty[as], ap ^ 1);
          tc_fence_after();
          const uint32_t wh = smem_u32(sW + ws * kV2WBytes), wl = wh + 16384;
          const uint32_t d = tmem_base + as * 128;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_x3(d, wh + ks * 4096, wl + ks * 4096, xh + ks * 2 * kV2XPlane, xl + ks * 2 * kV2XPlane, 2048, kV2XPlane, idesc, ks == 0);
          umma_commit(&w_empty[ws]);       // weight stage may be refilled once these MMAs retire
          umma_commit(&acc_full[as]);      // accumulator ready for the epilogue warps
          if (++ws == kV2WStages) { ws = 0; wp ^= 1; }
          if (++as == kV2AccStages) { as = 0; ap ^= 1; }
        }
        umma_commit(&x_empty[xs]);         // token tile buffer may be overwritten
        if (++xs == 2) { xs = 0; xp ^= 1; }
      }
    }
  } else {
    // ---------------- token-tile staging + epilogue (warps 2..9) ----------------
    const int et = tid - 64;               // 0..255
    const int ew = et >> 5;                // 0..7
    const int lane_grp = warp & 3;         // TMEM lane quarter this warp may read (warp id % 4)
    const int tok_half = ew >> 2;          // which 64 tokens of the tile
    int as = 0; uint32_t ap = 0;
    auto stage_tile = [&](int64_t tile, int buf) {
      // 128 rows x 8 k-chunks = 1024 units; thread -> (row = u >> 3, q = u & 7): 8 lanes cover one 256-byte row
      uint8_t* xh = sX + buf * kV2XBytes;
      uint8_t* xl = xh + kV2XHalf;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int u = i * 256 + et, row = u >> 3, q = u & 7;
        const int64_t r = tile * 128 + row;
        const bool ok = r < M;
        const float* src = A + (ok ? r : 0) * lda + q * 8;
        uint4 hi, lo;
        split8(ldg4z(src, ok), ldg4z(src + 4, ok), hi, lo);
        sts16(xh + q * kV2XPlane + row * 16, hi);
        sts16(xl + q * kV2XPlane + row * 16, lo);
      }
      fence_async_smem();
      mbar_arrive(&x_full[buf]);
    };
    int64_t tile = blockIdx.x;
    if (tile < ntiles) stage_tile(tile, 0);
    int stage_buf = 1; uint32_t stage_phase = 0;   // phase of x_empty for the NEXT buffer to fill
    for (; tile < ntiles; tile += gridDim.x) {
      const int64_t next = tile + gridDim.x;
      if (next < ntiles) {
        mbar_wait(&x_empty[stage_buf], stage_phase ^ 1);
        stage_tile(next, stage_buf);
        if (++stage_buf == 2) { stage_buf = 0; stage_phase ^= 1; }
      }
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
        float* dst = C + t0 * ldc + feat;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (t0 + j < M) dst[(int64_t)j * ldc] = v0[j] + b;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (t0 + 32 + j < M) dst[(int64_t)(32 + j) * ldc] = v1[j] + b;
        if (++as == kV2AccStages) { as = 0; ap ^= 1; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
bool plain(const GemmDesc& d) {
  return d.ngroups == 0 && !d.perm && !d.a_ids && !d.b_ids && !d.a_act && !d.b_act && !d.drop_on && !d.addend &&
         !d.epi_act && !d.epi_drop && (d.out_scale == 0.f || d.out_scale == 1.f);
}
template <typename K>
int set_smem(K kernel, int bytes) {
  return check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes), "cudaFuncSetAttribute");
}
}  // namespace

int launch_split_weights_k64(const float* W, int64_t ldw, int64_t N, void* out, cudaStream_t stream) {
  if (N % 128 != 0) { set_error("split_weights: N must be a multiple of 128"); return MATCHA_ERR_ARG; }
  split_weights_k64_kernel<<<(unsigned)((N * 8 + 255) / 256), 256, 0, stream>>>(W, ldw, N, reinterpret_cast<uint8_t*>(out));
  MATCHA_CHECK_LAUNCH("split_weights_k64");
  return MATCHA_OK;
}

int64_t gemm_tc_scratch_floats(int64_t M) { return (int64_t)kTcMaxSplits * (M * 64 + M); }

int launch_gemm_tc(const GemmDesc& d, cudaStream_t stream, bool* handled) {
  *handled = false;
  if (d.M <= 0 || d.N <= 0 || d.K <= 0) return MATCHA_OK;
  const bool al = aligned16(d.A) && aligned16(d.B) && aligned16(d.C) && d.lda % 4 == 0 && d.ldb % 4 == 0 && d.ldc % 4 == 0;
  if (!al) return MATCHA_OK;
  if (d.form == FORM_NT && plain(d) && d.K == 64 && d.N % 128 == 0 && d.b_split && d.M >= 128) {
    static bool once = false;
    if (!once) { if (int rc = set_smem(tc_qkg_fwd_v2_kernel, kV2Smem)) return rc; once = true; }
    const int64_t ntiles = (d.M + 127) / 128;
    const unsigned grid = (unsigned)(ntiles < kSMs ? ntiles : kSMs);
    tc_qkg_fwd_v2_kernel<<<grid, kV2Threads, kV2Smem, stream>>>(d.A, d.lda, d.b_split, d.bias, d.C, d.ldc, d.M, (int)d.N);
    MATCHA_CHECK_LAUNCH("tc_qkg_fwd_v2");
    *handled = true;
  } else if (d.form == FORM_NT && plain(d) && d.K == 64 && d.N % 256 == 0 && (!d.bias || aligned16(d.bias))) {
    constexpr int smem = 32768 + 2 * 256 * 128;   // 96 KB -> two CTAs (2 x 256 TMEM columns) per SM
    static bool once = false;
    if (!once) { if (int rc = set_smem(tc_nt_k64_kernel<256>, smem)) return rc; once = true; }
    tc_nt_k64_kernel<256><<<(unsigned)((d.M + 127) / 128), kThreads, smem, stream>>>(d.A, d.lda, d.B, d.ldb, d.C, d.ldc,
                                                                                     d.bias, d.M, d.N);
    MATCHA_CHECK_LAUNCH("tc_nt_k64");
    *handled = true;
  } else if (d.form == FORM_NN && plain(d) && !d.bias && d.N == 64 && d.K % 64 == 0) {
    constexpr int smem = 2 * (16 * 2064 + 16384);   // ~97 KB -> two CTAs per SM
    static bool once = false;
    if (!once) { if (int rc = set_smem(tc_nn_n64_kernel, smem)) return rc; once = true; }
    tc_nn_n64_kernel<<<(unsigned)((d.M + 127) / 128), kThreads, smem, stream>>>(d.A, d.lda, d.B, d.ldb, d.C, d.ldc, d.M, d.K);
    MATCHA_CHECK_LAUNCH("tc_nn_n64");
    *handled = true;
  } else if (d.form == FORM_TN && d.ngroups == 0 && !d.perm && !d.a_ids && !d.b_ids && !d.a_act && !d.b_act && !d.drop_on &&
             d.N == 64 && d.ldb == 64 && d.M % 512 == 0 && d.scratch && d.scratch_floats >= gemm_tc_scratch_floats(d.M) &&
             d.K >= 2048) {
    constexpr int smem = 81920;                   // 2 x 36 KB used; 80 KB requested so at most two CTAs (2 x 256 TMEM columns) share an SM
    static bool once = false;
    if (!once) { if (int rc = set_smem(tc_tn_n64_kernel, smem)) return rc; once = true; }
    const int mgroups = (int)(d.M / 512);
    int splits = (2 * kSMs + mgroups - 1) / mgroups;
    if (splits > kTcMaxSplits) splits = kTcMaxSplits;
    int64_t kps = (d.K + splits - 1) / splits;
    kps = (kps + 15) / 16 * 16;
    splits = (int)((d.K + kps - 1) / kps);
    float* part = d.scratch;
    float* part_cs = d.scratch + (int64_t)kTcMaxSplits * d.M * 64;
    dim3 grid((unsigned)mgroups, (unsigned)splits);
    tc_tn_n64_kernel<<<grid, kThreads, smem, stream>>>(d.A, d.lda, d.B, d.ldb, d.M, d.K, kps, part, d.colsum ? part_cs : nullptr);
    MATCHA_CHECK_LAUNCH("tc_tn_n64");
    const float scale = d.out_scale == 0.f ? 1.f : d.out_scale;
    tn_reduce_kernel<<<kSMs * 2, 256, 0, stream>>>(part, d.colsum ? part_cs : nullptr, splits, d.M, d.C, d.ldc, d.colsum,
                                                   d.colsum_n, scale);
    MATCHA_CHECK_LAUNCH("tn_reduce");
    *handled = true;
  }
  return MATCHA_OK;
}

}  // namespace matcha
