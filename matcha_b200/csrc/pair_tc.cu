// Tensor-core all-pairs scorer (denoise_contact.py:67-88) for the k = 2 closed form of scorer.cu:
//   logit(i, j) = 0.5 * sum_c w_c [ (D[j,c] - S[i,c])^2 + (D[i,c] - S[j,c])^2 ] + b
//               = u_i + u_j - [ (w.S)_i . D_j + D_i . (w.S)_j ] + b,      u_n = 0.5 * sum_c w_c (D[n,c]^2 + S[n,c]^2)
// The bracket is one K = 128 contraction  PA_i . PB_j  with  PA = [w.S | D],  PB = [D | w.S]: a 128 x 128 output tile is
// 24 tcgen05.mma instructions (bf16x3 split, fp32 accumulate in TMEM) instead of 16 384 x 64 x 4 fp32 FMAs, which moves
// the kernel from the FMA pipe to the HBM write roofline (4 bytes per pair).
//   prepare : D, S, w -> PA / PB as pre-split bf16 hi | lo blocks of 128 nodes in the canonical UMMA layout, and u
//   score   : persistent CTAs over the upper-triangular tile list; warp 0 bulk-copies operand blocks, warp 1 issues the
//             MMAs of the TRANSPOSED tile (TMEM lane = column node), two epilogue warpgroups drain alternate TMEM stages
//             and store straight from registers: one store instruction = 128 contiguous bytes of a packed output row
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace matcha {
namespace {

constexpr int kPD = 64;
// one block = 128 nodes x 144 k: hi [18 planes][128][8] (36 KB) | lo [16 planes] (32 KB).  k 0..127 carry the contraction
// operands; k 128..143 (hi half only, one extra bf16 pass) carry -log2(e) * u as three exact bf16 pieces against ones, so
// the accumulator already holds t = -log2(e) * logit and the epilogue is ex2 / add / rcp / store.
constexpr int kPHi = 18 * 2048;
constexpr int kPBlk = kPHi + 16 * 2048;      // 69632
constexpr int kPStages = 4;                  // TMEM accumulator stages = epilogue warpgroups
constexpr int kPThreads = 64 + kPStages * 128;
constexpr int kPSmem = 3 * kPBlk;            // A + 2 x B = 208 896
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__host__ __device__ __forceinline__ int64_t tri_prefix(int64_t r, int64_t n, int64_t md) {
  const int64_t full = n - md;               // items in row 0 when row q holds max(0, n - md - q)
  if (full <= 0) return 0;
  if (r > full) r = full;
  return r * full - r * (r - 1) / 2;
}
__host__ __device__ __forceinline__ int64_t tri_row_of(int64_t p, int64_t n, int64_t md) {   // largest r with prefix(r) <= p
  int64_t a = 0, b = n;
  while (b - a > 1) { const int64_t mid = (a + b) / 2; if (tri_prefix(mid, n, md) <= p) a = mid; else b = mid; }
  return a;
}

// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t bf16_bits(float x) { return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(x)); }
__device__ __forceinline__ float bf16_val(uint32_t b) { return __uint_as_float(b << 16); }

__global__ void pair_prep_kernel(const float* __restrict__ D, const float* __restrict__ S, const float* __restrict__ w,
                                 const float* __restrict__ cls_b, int64_t lo, int64_t n, int64_t nblocks,
                                 uint8_t* __restrict__ PA, uint8_t* __restrict__ PB) {
  const int64_t unit = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // (relative node r, plane g of 18)
  if (unit >= nblocks * 128 * 18) return;
  const int g = (int)(unit % 18);
  const int64_t r = unit / 18;
  const int64_t off = (r >> 7) * kPBlk + g * 2048 + (r & 127) * 16;
  if (g < 16) {
    float a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = 0.f; b[i] = 0.f; }
    if (r < n) {
      const int c0 = (g & 7) * 8;
      const float* dr = D + (lo + r) * kPD + c0;
      const float* sr = S + (lo + r) * kPD + c0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float ws = __ldg(w + c0 + i) * __ldg(sr + i), dv = __ldg(dr + i);
        a[i] = g < 8 ? ws : dv;                       // PA = [w.S | D]
        b[i] = (g < 8 ? dv : ws) * kLog2e;            // PB = log2(e) * [D | w.S]
      }
    }
    uint4 hi, lo4;
    split8(make_float4(a[0], a[1], a[2], a[3]), make_float4(a[4], a[5], a[6], a[7]), hi, lo4);
    *reinterpret_cast<uint4*>(PA + off) = hi;
    *reinterpret_cast<uint4*>(PA + off + kPHi) = lo4;
    split8(make_float4(b[0], b[1], b[2], b[3]), make_float4(b[4], b[5], b[6], b[7]), hi, lo4);
    *reinterpret_cast<uint4*>(PB + off) = hi;
    *reinterpret_cast<uint4*>(PB + off + kPHi) = lo4;
  } else if (g == 16) {
    // k 128..130: three bf16 pieces of -log2(e) u (A side) against ones (B side); k 131..133: ones against the pieces of
    // -log2(e) (u + b); the sum of the pieces is exact to 2^-24
    uint32_t pa[4] = {0u, 0u, 0u, 0u}, pb[4] = {0u, 0u, 0u, 0u};
    if (r < n) {
      float acc = 0.f;
      for (int c = 0; c < kPD; ++c) {
        const float dd = __ldg(D + (lo + r) * kPD + c), ss = __ldg(S + (lo + r) * kPD + c);
        acc = fmaf(__ldg(w + c), fmaf(dd, dd, ss * ss), acc);
      }
      const float u = 0.5f * acc;
      const float xa = -kLog2e * u, xb = -kLog2e * (u + __ldg(cls_b));
      const uint32_t ah = bf16_bits(xa), am = bf16_bits(xa - bf16_val(ah)), al = bf16_bits(xa - bf16_val(ah) - bf16_val(am));
      const uint32_t bh = bf16_bits(xb), bm = bf16_bits(xb - bf16_val(bh)), bl = bf16_bits(xb - bf16_val(bh) - bf16_val(bm));
      const uint32_t one = 0x3F80u;
      pa[0] = ah | (am << 16); pa[1] = al | (one << 16); pa[2] = one | (one << 16);
      pb[0] = one | (one << 16); pb[1] = one | (bh << 16); pb[2] = bm | (bl << 16);
    }
    *reinterpret_cast<uint4*>(PA + off) = make_uint4(pa[0], pa[1], pa[2], pa[3]);
    *reinterpret_cast<uint4*>(PB + off) = make_uint4(pb[0], pb[1], pb[2], pb[3]);
  } else {
    *reinterpret_cast<uint4*>(PA + off) = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(PB + off) = make_uint4(0u, 0u, 0u, 0u);
  }
}

// ------------------------------------------------------------------------------------------
struct PairArgs {
  const uint8_t* PA; const uint8_t* PB;
  int64_t n; int md; int64_t nb; int bmd;          // nodes, min distance, 128-node blocks, md / 128
  int64_t f_begin, f_end;                          // flat tile range of this launch
  int64_t p_begin, p_end;
  int apply_sigmoid;
  float* out;
  int chunk;                                       // consecutive tiles per CTA visit (see pair_tc_kernel)
};

__device__ __forceinline__ float fast_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void flat_to_tile(int64_t f, int64_t nb, int bmd, int64_t& ti, int64_t& tj) {
  ti = tri_row_of(f, nb, bmd);
  tj = ti + bmd + (f - tri_prefix(ti, nb, bmd));
}

__global__ void __launch_bounds__(kPThreads, 1) pair_tc_kernel(const PairArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sAop = smem;
  uint8_t* sBop = smem + kPBlk;
  __shared__ uint64_t a_full, a_empty, b_full[2], b_empty[2], acc_full[kPStages], acc_empty[kPStages];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) tmem_alloc(&tmem_base_s, kPStages * 128);
  if (tid == 32) {
    mbar_init(&a_full, 1); mbar_init(&a_empty, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < kPStages; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  // Work distribution: the flat (row-major) tile list is cut into chunks of a.chunk consecutive tiles, dealt round-robin to
  // the CTAs.  At any moment the whole grid therefore writes inside a window of gridDim.x * chunk tiles = a few tile rows
  // (a few hundred output rows, each receiving long contiguous runs) instead of 148 x 128 unrelated output rows: the
  // HBM write stream stays page-local.  The row block (A) is reloaded per chunk; the operand tables live in L2.
  const int64_t G = a.chunk;
  const int64_t nchunks = (a.f_end - a.f_begin + G - 1) / G;
  const int64_t my_chunks = ((int64_t)blockIdx.x < nchunks) ? (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int64_t f0 = 0, f1 = my_chunks;          // (kept for the role guards below)
  auto chunk_begin = [&](int64_t c) { return a.f_begin + ((int64_t)blockIdx.x + c * gridDim.x) * G; };
  auto chunk_end = [&](int64_t c) { const int64_t e = chunk_begin(c) + G; return e < a.f_end ? e : a.f_end; };

  if (warp == 0) {
    if (f0 < f1 && elect_one()) {
      int64_t ti, tj, cur_ti = -1, na = 0, k = 0;
      for (int64_t c = 0; c < my_chunks; ++c) {
        flat_to_tile(chunk_begin(c), a.nb, a.bmd, ti, tj);
        cur_ti = -1;                                  // the row block is (re)loaded at every chunk start
        for (int64_t f = chunk_begin(c); f < chunk_end(c); ++f, ++k) {
          if (ti != cur_ti) {
            mbar_wait_backoff(&a_empty, (uint32_t)(na & 1) ^ 1u);
            mbar_expect_tx(&a_full, kPBlk);
            bulk_g2s(sAop, a.PA + ti * (int64_t)kPBlk, kPBlk, &a_full);
            cur_ti = ti; ++na;
          }
          const int s = (int)(k & 1);
          mbar_wait_backoff(&b_empty[s], (uint32_t)((k >> 1) & 1) ^ 1u);
          mbar_expect_tx(&b_full[s], kPBlk);
          bulk_g2s(sBop + s * kPBlk, a.PB + tj * (int64_t)kPBlk, kPBlk, &b_full[s]);
          if (++tj >= a.nb) { ++ti; tj = ti + a.bmd; }
        }
      }
    }
  } else if (warp == 1) {
    if (f0 < f1 && elect_one()) {
      constexpr uint32_t idesc = make_idesc(128, 128, false, false);
      int64_t ti, tj, cur_ti = -1, na = 0, k = 0;
      // descriptors are built once: advancing an operand by one k-step (4096 B) adds 256 to the 16-byte address field
      const uint64_t dah = make_smem_desc(smem_u32(sAop), 2048, 128), dal = make_smem_desc(smem_u32(sAop) + kPHi, 2048, 128);
      for (int64_t c = 0; c < my_chunks; ++c) {
        flat_to_tile(chunk_begin(c), a.nb, a.bmd, ti, tj);
        const int64_t fe = chunk_end(c);
        for (int64_t f = chunk_begin(c); f < fe; ++f, ++k) {
          if (ti != cur_ti) { mbar_wait(&a_full, (uint32_t)(na & 1)); cur_ti = ti; ++na; }
          const int s = (int)(k & 1), as = (int)(k & (kPStages - 1));
          mbar_wait(&b_full[s], (uint32_t)((k >> 1) & 1));
          mbar_wait(&acc_empty[as], (uint32_t)((k / kPStages) & 1) ^ 1u);
          tc_fence_after();
          const uint64_t dbh = make_smem_desc(smem_u32(sBop + s * kPBlk), 2048, 128);
          const uint64_t dbl = make_smem_desc(smem_u32(sBop + s * kPBlk) + kPHi, 2048, 128);
          const uint32_t acc = tmem_base + as * 128;
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {      // D[j, i]: the column-node block is the M-side operand; lo*hi + hi*lo + hi*hi
            const uint64_t o = (uint64_t)(ks * 256);
            umma_bf16(acc, dbl + o, dah + o, idesc, ks == 0 ? 0u : 1u);
            umma_bf16(acc, dbh + o, dal + o, idesc, 1u);
            umma_bf16(acc, dbh + o, dah + o, idesc, 1u);
          }
          // k 128..143: the u / bias pieces (exact in bf16), hi * hi pass only
          umma_bf16(acc, dbh + 8 * 256, dah + 8 * 256, idesc, 1u);
          umma_commit(&b_empty[s]);
          umma_commit(&acc_full[as]);
          int64_t nti = ti, ntj = tj + 1;
          if (ntj >= a.nb) { ++nti; ntj = nti + a.bmd; }
          if (nti != ti || f + 1 >= fe) { umma_commit(&a_empty); cur_ti = -1; }      // last tile of the chunk that reads this A block
          ti = nti; tj = ntj;
        }
      }
    }
  } else if (f0 < f1) {
    // The accumulator is the TRANSPOSED tile: TMEM lane = column node j, TMEM column = row node i.  The thread that owns
    // lane j therefore holds 32 consecutive rows i in its registers, and one store instruction of the warp (fixed register,
    // 32 lanes) writes 32 consecutive j of one packed output row: a contiguous 128-byte run, straight from registers.
    // Four epilogue warpgroups <-> four TMEM stages: while one group is held back by the HBM-bound store queue, another is
    // reading TMEM and the tensor pipe runs ahead on the remaining stages.
    const int grp = (warp - 2) >> 2, q = warp & 3;          // epilogue warpgroup <-> TMEM stage; lane quarter
    int64_t ti, tj, k = 0;
    const int64_t full = a.n - a.md;
    const int64_t plen = a.p_end - a.p_begin;
    float* const out = a.out;
    const bool sig = a.apply_sigmoid != 0;
    for (int64_t c = 0; c < my_chunks; ++c) {
    flat_to_tile(chunk_begin(c), a.nb, a.bmd, ti, tj);
    for (int64_t f = chunk_begin(c); f < chunk_end(c); ++f, ++k) {
      if ((int)(k & (kPStages - 1)) == grp) {
        const int64_t i_first = ti * 128, i_last = i_first + 127;
        const int64_t j_first = tj * 128 + q * 32, j_last = j_first + 31, j = j_first + lane;
        // p(i, j) = rowbase(i) + j,  rowbase(i + 1) - rowbase(i) = full - i - 1 for every row that holds pairs
        const int64_t p00 = tri_prefix(i_first, a.n, a.md) - (i_first + a.md) - a.p_begin;
        // interior slab: every (i, j) is a valid pair inside [p_begin, p_end) -> no per-element checks
        const bool interior = i_last < a.n && j_last < a.n && j_first >= i_last + a.md && p00 + j_first >= 0 &&
                              tri_prefix(i_last, a.n, a.md) - (i_last + a.md) + j_last < a.p_end;
        int64_t p = p00 + j;
        int64_t stride = full - i_first - 1;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + grp * 128;
        mbar_wait(&acc_full[grp], (uint32_t)((k / kPStages) & 1));
        tc_fence_after();
        auto emit = [&](const uint32_t (&cur)[32]) {
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const float t = __uint_as_float(cur[c]);                      // -log2(e) * logit
            out[p] = sig ? fast_rcp(1.0f + fast_ex2(t)) : -kLn2 * t;
            p += stride;
            --stride;
          }
        };
        if (interior) {
          uint32_t va[32], vb[32];
          tmem_ld32_issue(taddr, va);
#pragma unroll 1
          for (int half = 0; half < 2; ++half) {     // not unrolled: the fully unrolled epilogue overflowed the instruction cache
            tmem_ld_wait(va);
            tmem_ld32_issue(taddr + half * 64 + 32, vb);
            emit(va);
            tmem_ld_wait(vb);
            if (half == 0) {
              tmem_ld32_issue(taddr + 64, va);
            } else {                    // the whole slab is in registers: hand the TMEM stage back before the last stores
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&acc_empty[grp]);
            }
            emit(vb);
          }
        } else {
          // diagonal / range-boundary slab: per-element checks, compact code (8 columns at a time)
          const bool j_ok = j < a.n;
#pragma unroll 1
          for (int ch = 0; ch < 16; ++ch) {
            uint32_t v[8];
            tmem_ld8_issue(taddr + ch * 8, v);
            tmem_ld_wait(v);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const int64_t i = i_first + ch * 8 + c;
              const float t = __uint_as_float(v[c]);
              if (j_ok && j >= i + a.md && (uint64_t)p < (uint64_t)plen) out[p] = sig ? fast_rcp(1.0f + fast_ex2(t)) : -kLn2 * t;
              p += stride;
              --stride;
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[grp]);
        }
      }
      if (++tj >= a.nb) { ++ti; tj = ti + a.bmd; }
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kPStages * 128);
}

}  // namespace
}  // namespace matcha

using namespace matcha;

extern "C" {

// workspace: PA | PB blocks (68 KB per 128 nodes each)
int64_t matcha_pair_tc_workspace_bytes(int64_t lo, int64_t hi) {
  const int64_t n = hi - lo;
  if (n <= 0) return 0;
  const int64_t nb = (n + 127) / 128;
  return 2 * nb * (int64_t)kPBlk + 256;
}

int matcha_pair_tc_prepare(const float* D, const float* S, const float* cls_w, const float* cls_b, int32_t d, int64_t lo,
                           int64_t hi, void* workspace, int64_t workspace_bytes, void* stream) {
  MATCHA_REQUIRE(D && S && cls_w && cls_b && workspace, "pair_tc_prepare: NULL argument");
  if (d != kPD) { set_error("pair_tc_prepare: embed_dim %d unsupported (64)", d); return MATCHA_ERR_UNSUPPORTED; }
  MATCHA_REQUIRE(((uintptr_t)workspace & 255) == 0, "pair_tc_prepare: workspace must be 256-byte aligned");
  const int64_t n = hi - lo;
  MATCHA_REQUIRE(n > 0 && workspace_bytes >= matcha_pair_tc_workspace_bytes(lo, hi), "pair_tc_prepare: workspace too small");
  const int64_t nb = (n + 127) / 128;
  uint8_t* PA = reinterpret_cast<uint8_t*>(workspace);
  uint8_t* PB = PA + nb * (int64_t)kPBlk;
  const int64_t units = nb * 128 * 18;
  pair_prep_kernel<<<(unsigned)((units + 255) / 256), 256, 0, (cudaStream_t)stream>>>(D, S, cls_w, cls_b, lo, n, nb, PA, PB);
  MATCHA_CHECK_LAUNCH("pair_prep");
  return MATCHA_OK;
}

int matcha_pair_tc_score_range(const void* workspace, int64_t lo, int64_t hi, int32_t min_dis, int64_t p_begin, int64_t p_end,
                               int32_t apply_sigmoid, float* out, void* stream) {
  MATCHA_REQUIRE(workspace && out, "pair_tc_score_range: NULL argument");
  MATCHA_REQUIRE(min_dis >= 0, "pair_tc_score_range: negative min_distance");
  const int64_t n = hi - lo;
  const int64_t total = tri_prefix(n, n, min_dis);
  MATCHA_REQUIRE(p_begin >= 0 && p_end <= total && p_begin <= p_end, "pair_tc_score_range: pair range [%lld, %lld) outside [0, %lld)",
                 (long long)p_begin, (long long)p_end, (long long)total);
  if (p_begin == p_end) return MATCHA_OK;
  const int64_t nb = (n + 127) / 128;
  const int bmd = min_dis / 128;
  const int64_t r0 = tri_row_of(p_begin, n, min_dis), r1 = tri_row_of(p_end - 1, n, min_dis);
  PairArgs a;
  a.PA = reinterpret_cast<const uint8_t*>(workspace);
  a.PB = a.PA + nb * (int64_t)kPBlk;
  a.n = n; a.md = min_dis; a.nb = nb; a.bmd = bmd;
  a.f_begin = tri_prefix(r0 / 128, nb, bmd);
  a.f_end = tri_prefix(r1 / 128 + 1, nb, bmd);
  a.p_begin = p_begin; a.p_end = p_end; a.apply_sigmoid = apply_sigmoid; a.out = out;
  { const char* e = getenv("MATCHA_PAIR_CHUNK"); a.chunk = e ? atoi(e) : 4; if (a.chunk < 1) a.chunk = 1; }   // 4 measured best on B200 (sweep 1..130)
  if (a.f_end <= a.f_begin) return MATCHA_OK;
  static bool once = false;
  if (!once) {
    if (int rc = check_cuda(cudaFuncSetAttribute(pair_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPSmem), "cudaFuncSetAttribute"))
      return rc;
    once = true;
  }
  const int64_t tiles = a.f_end - a.f_begin;
  const unsigned grid = (unsigned)(tiles < kSMs ? tiles : kSMs);
  prof_begin(P_PAIR_SCORE, (cudaStream_t)stream);
  pair_tc_kernel<<<grid, kPThreads, kPSmem, (cudaStream_t)stream>>>(a);
  MATCHA_CHECK_LAUNCH("pair_tc");
  prof_end(P_PAIR_SCORE, 1, (cudaStream_t)stream);
  return MATCHA_OK;
}

}  // extern "C"
