// Node encoder forward on tensor cores (Modules.py:104-122 via :176-188; embed_dim 64, dense feature rows):
//   H0 = tanh( dropout(F_c[id - start_c]) . W0_c^T ),   E = H0 . W1_c^T        for the tokens of chromosome c
// as ONE kernel over the chromosome-bucketed token list: a CTA owns 128 tokens of one chromosome (thread r = token
// row r = TMEM lane r), walks the feature row in chunks of 64 columns { gather rows (coalesced, through a staging
// area) -> dropout from the counter RNG -> bf16 hi | lo A tile -> tcgen05 against the pre-split W0 chunk }, applies
// tanh to the accumulator row, stores H0 (the backward pass needs it), and chains the 64 x 64 contraction with W1_c.
// Replaces two grouped SIMT launches and the H0 re-read between them.
#include <string.h>

#include "rowwise.cuh"
#include "tc_common.cuh"

namespace matcha {
namespace {

constexpr int kEThreads = 128;
constexpr int kEChunk = 16384;                 // one 64 x 64 weight chunk: K-major [k/8][row][8], hi 8 KB | lo 8 KB
constexpr int kEA = 32768;                     // A tile: 128 tokens x 64 k, hi 16 KB | lo 16 KB
constexpr int kEStageRow = 68;
constexpr int kEStage = 4 * 32 * kEStageRow * 4;     // 34 816
constexpr int kESmem = kEA + 2 * kEChunk + kEStage;  // 100 352 -> two CTAs per SM

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

// gather 64 consecutive floats of 32 scattered rows (row pointer held by the lane that owns the row, NULL = zero row):
// two rows per instruction, 16 lanes x float4 each (256 B contiguous per row), transposed through the staging area
__device__ __forceinline__ void gather_rows(const float* rowp, int64_t k0, int64_t klimit, float* stage, int lane, float (&v)[64]) {
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int row = 2 * k + (lane >> 4), c4 = lane & 15;
    const float* p = shfl_ptr(rowp, row);
    float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p != nullptr && k0 + c4 * 4 < klimit) e = __ldg(reinterpret_cast<const float4*>(p + k0) + c4);
    *reinterpret_cast<float4*>(stage + row * kEStageRow + c4 * 4) = e;
  }
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const float4 e = *reinterpret_cast<const float4*>(stage + lane * kEStageRow + k * 4);
    v[4 * k] = e.x; v[4 * k + 1] = e.y; v[4 * k + 2] = e.z; v[4 * k + 3] = e.w;
  }
  __syncwarp();
}
// scatter 32 rows of 64 floats to their (scattered) destinations, same access pattern
__device__ __forceinline__ void scatter_rows(float* rowp, float* stage, int lane, const float (&v)[64]) {
#pragma unroll
  for (int k = 0; k < 16; ++k)
    *reinterpret_cast<float4*>(stage + lane * kEStageRow + k * 4) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int row = 2 * k + (lane >> 4), c4 = lane & 15;
    float* p = shfl_ptr(rowp, row);
    if (p != nullptr) reinterpret_cast<float4*>(p)[c4] = *reinterpret_cast<const float4*>(stage + row * kEStageRow + c4 * 4);
  }
  __syncwarp();
}
__device__ __forceinline__ void put_tile(uint8_t* sA, int r, const float (&v)[64]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    uint4 hi, lo;
    split8(make_float4(v[8 * j], v[8 * j + 1], v[8 * j + 2], v[8 * j + 3]),
           make_float4(v[8 * j + 4], v[8 * j + 5], v[8 * j + 6], v[8 * j + 7]), hi, lo);
    sts16(sA + j * 2048 + r * 16, hi);
    sts16(sA + 16384 + j * 2048 + r * 16, lo);
  }
}
__device__ __forceinline__ void load_chunk(uint8_t* dst, const uint8_t* src) {   // 16 KB, all 128 threads
  const uint4* s = reinterpret_cast<const uint4*>(src);
  uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int i = 0; i < kEChunk / 16 / kEThreads; ++i) d[threadIdx.x + i * kEThreads] = __ldg(s + threadIdx.x + i * kEThreads);
}

__global__ void __launch_bounds__(kEThreads, 2)
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
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, r = tid;
  if (warp == 0) tmem_alloc(&tmem_base_s, 64);
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t tlane = tmem_base + ((uint32_t)(warp * 32) << 16);
  constexpr uint32_t idesc = make_idesc(128, 64, false, false);
  const uint32_t ah = smem_u32(sA), al = ah + 16384;
  const uint32_t wh = smem_u32(sW), wl = wh + 8192;
  const uint32_t vh = smem_u32(sW1), vl = vh + 8192;
  float* stage = sStage + warp * (32 * kEStageRow);
  uint32_t phase = 0;
  int cur_c = -1;

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
    const int nchunk = (em.nc[c] + 63) / 64;
    const uint8_t* wbase = wsplit + em.woff[c];
    if (c != cur_c) {                       // W1_c stays resident while the CTA works on this chromosome
      __syncthreads();
      load_chunk(sW1, wbase + (int64_t)nchunk * kEChunk);
      cur_c = c;
    }
    for (int kc = 0; kc < nchunk; ++kc) {
      float v[64];
      gather_rows(frow, (int64_t)kc * 64, em.ld[c], stage, lane, v);
      if (drop.thr != 0u && live) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float4 d = drop_apply4(drop, (uint64_t)t, (uint32_t)(kc * 64 + 4 * j), make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
          v[4 * j] = d.x; v[4 * j + 1] = d.y; v[4 * j + 2] = d.z; v[4 * j + 3] = d.w;
        }
      }
      put_tile(sA, r, v);
      load_chunk(sW, wbase + (int64_t)kc * kEChunk);
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
      mbar_wait(&bar, phase);               // the A tile and the weight chunk are rewritten by the next chunk
      phase ^= 1;
      tc_fence_after();
    }
    // ---- H0 = tanh(acc): kept for the backward pass, and the A operand of the second contraction ----
    float h[64];
    {
      uint32_t d0[32], d1[32];
      tmem_ld32_issue(tlane, d0);
      tmem_ld32_issue(tlane + 32, d1);
      tmem_ld_wait(d0);
      tmem_ld_wait(d1);
#pragma unroll
      for (int i = 0; i < 32; ++i) { h[i] = live ? tanhf(__uint_as_float(d0[i])) : 0.f; h[32 + i] = live ? tanhf(__uint_as_float(d1[i])) : 0.f; }
    }
    tc_fence_before();
    scatter_rows(live ? H0 + t * 64 : nullptr, stage, lane, h);
    put_tile(sA, r, h);
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        umma_x3s(tmem_base, ah + ks * 4096, al + ks * 4096, vh + ks * 2048, vl + ks * 2048, 2048, 128, 1024, 128, idesc, ks == 0);
      umma_commit(&bar);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    tc_fence_after();
    {
      uint32_t d0[32], d1[32];
      tmem_ld32_issue(tlane, d0);
      tmem_ld32_issue(tlane + 32, d1);
      tmem_ld_wait(d0);
      tmem_ld_wait(d1);
#pragma unroll
      for (int i = 0; i < 32; ++i) { h[i] = __uint_as_float(d0[i]); h[32 + i] = __uint_as_float(d1[i]); }
    }
    tc_fence_before();
    scatter_rows(live ? E + t * 64 : nullptr, stage, lane, h);
    __syncthreads();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 64);
}

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
  const unsigned grid = (unsigned)(tiles < 2 * kSMs ? tiles : 2 * kSMs);
  enc_tc_fwd_kernel<<<grid, kEThreads, kESmem, s>>>(em, reinterpret_cast<const uint8_t*>(m->derived + split_base), x, perm,
                                                    group_off, H0, E, drop);
  MATCHA_CHECK_LAUNCH("enc_tc_fwd");
  return MATCHA_OK;
}

}  // namespace matcha
