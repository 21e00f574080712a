// Tensor-core all-pairs scorer (denoise_contact.py:67-88) for the k = 2 closed form of scorer.cu:
//   logit(i, j) = 0.5 * sum_c w_c [ (D[j,c] - S[i,c])^2 + (D[i,c] - S[j,c])^2 ] + b
//               = u_i + u_j - [ (w.S)_i . D_j + D_i . (w.S)_j ] + b,      u_n = 0.5 * sum_c w_c (D[n,c]^2 + S[n,c]^2)
// The bracket is one K = 128 contraction  PA_i . PB_j  with  PA = [w.S | D],  PB = [D | w.S]: a 128 x 128 output tile is
// 24 tcgen05.mma instructions (bf16x3 split, fp32 accumulate in TMEM) instead of 16 384 x 64 x 4 fp32 FMAs, which moves
// the kernel from the FMA pipe to the HBM write roofline (4 bytes per pair).
//   prepare : D, S, w -> PA / PB as pre-split bf16 hi | lo blocks of 128 nodes in the canonical UMMA layout, and u
//   score   : persistent CTAs over the upper-triangular tile list; warp 0 bulk-copies operand blocks, warp 1 issues the
//             MMAs, two epilogue warpgroups drain alternate TMEM stages, transpose 16-column chunks through shared
//             memory so that every store instruction writes contiguous runs of the packed row-major pair order
#include "common.cuh"
#include "tc_common.cuh"

namespace matcha {
namespace {

constexpr int kPD = 64;
constexpr int kPBlk = 65536;                 // one block: 128 nodes x 128 k, hi 32 KB ([16 planes][128][8]) | lo 32 KB
constexpr int kPThreads = 320;
constexpr int kPStageRow = 20;               // floats per staged row (16 + 4 pad)
constexpr int kPStageWarp = 32 * kPStageRow;
constexpr int kPSmem = 3 * kPBlk + 8 * kPStageWarp * 4;      // A + 2 x B + staging = 217 088

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
__global__ void pair_prep_kernel(const float* __restrict__ D, const float* __restrict__ S, const float* __restrict__ w,
                                 int64_t lo, int64_t n, int64_t nblocks, uint8_t* __restrict__ PA, uint8_t* __restrict__ PB,
                                 float* __restrict__ u) {
  const int64_t unit = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // (relative node r, group g of 8 of the 128 k)
  if (unit >= nblocks * 128 * 16) return;
  const int g = (int)(unit & 15);
  const int64_t r = unit >> 4;
  float a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = 0.f; b[i] = 0.f; }
  if (r < n) {
    const int c0 = (g & 7) * 8;
    const float* dr = D + (lo + r) * kPD + c0;
    const float* sr = S + (lo + r) * kPD + c0;
    float ws[8], dv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { ws[i] = __ldg(w + c0 + i) * __ldg(sr + i); dv[i] = __ldg(dr + i); }
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = g < 8 ? ws[i] : dv[i]; b[i] = g < 8 ? dv[i] : ws[i]; }
    if (g == 0) {
      float acc = 0.f;
      for (int c = 0; c < kPD; ++c) {
        const float dd = __ldg(D + (lo + r) * kPD + c), ss = __ldg(S + (lo + r) * kPD + c);
        acc = fmaf(__ldg(w + c), fmaf(dd, dd, ss * ss), acc);
      }
      u[r] = 0.5f * acc;
    }
  } else if (g == 0) {
    u[r] = 0.f;
  }
  uint4 hi, lo4;
  const int64_t off = (r >> 7) * kPBlk + g * 2048 + (r & 127) * 16;
  split8(make_float4(a[0], a[1], a[2], a[3]), make_float4(a[4], a[5], a[6], a[7]), hi, lo4);
  *reinterpret_cast<uint4*>(PA + off) = hi;
  *reinterpret_cast<uint4*>(PA + off + 32768) = lo4;
  split8(make_float4(b[0], b[1], b[2], b[3]), make_float4(b[4], b[5], b[6], b[7]), hi, lo4);
  *reinterpret_cast<uint4*>(PB + off) = hi;
  *reinterpret_cast<uint4*>(PB + off + 32768) = lo4;
}

// ------------------------------------------------------------------------------------------
struct PairArgs {
  const uint8_t* PA; const uint8_t* PB; const float* u; const float* cls_b;
  int64_t n; int md; int64_t nb; int bmd;          // nodes, min distance, 128-node blocks, md / 128
  int64_t f_begin, f_end;                          // flat tile range of this launch
  int64_t p_begin, p_end;
  int apply_sigmoid;
  float* out;
};

__device__ __forceinline__ void flat_to_tile(int64_t f, int64_t nb, int bmd, int64_t& ti, int64_t& tj) {
  ti = tri_row_of(f, nb, bmd);
  tj = ti + bmd + (f - tri_prefix(ti, nb, bmd));
}

__global__ void __launch_bounds__(kPThreads, 1) pair_tc_kernel(const PairArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sAop = smem;
  uint8_t* sBop = smem + kPBlk;
  float* sStage = reinterpret_cast<float*>(smem + 3 * kPBlk);
  __shared__ uint64_t a_full, a_empty, b_full[2], b_empty[2], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) tmem_alloc(&tmem_base_s, 256);
  if (tid == 32) {
    mbar_init(&a_full, 1); mbar_init(&a_empty, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  // contiguous share of the flat tile list (row-major, so the A block changes rarely)
  const int64_t total = a.f_end - a.f_begin;
  const int64_t per = (total + gridDim.x - 1) / gridDim.x;
  const int64_t f0 = a.f_begin + (int64_t)blockIdx.x * per;
  const int64_t f1 = (f0 + per < a.f_end) ? f0 + per : a.f_end;

  if (warp == 0) {
    if (lane == 0 && f0 < f1) {
      int64_t ti, tj, cur_ti = -1, na = 0;
      flat_to_tile(f0, a.nb, a.bmd, ti, tj);
      for (int64_t f = f0, k = 0; f < f1; ++f, ++k) {
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
  } else if (warp == 1) {
    if (lane == 0 && f0 < f1) {
      constexpr uint32_t idesc = make_idesc(128, 128, false, false);
      int64_t ti, tj, cur_ti = -1, na = 0;
      flat_to_tile(f0, a.nb, a.bmd, ti, tj);
      const uint32_t ah = smem_u32(sAop), al = ah + 32768;
      for (int64_t f = f0, k = 0; f < f1; ++f, ++k) {
        if (ti != cur_ti) { mbar_wait_backoff(&a_full, (uint32_t)(na & 1)); cur_ti = ti; ++na; }
        const int s = (int)(k & 1);
        mbar_wait_backoff(&b_full[s], (uint32_t)((k >> 1) & 1));
        mbar_wait_backoff(&acc_empty[s], (uint32_t)((k >> 1) & 1) ^ 1u);
        tc_fence_after();
        const uint32_t bh = smem_u32(sBop + s * kPBlk), bl = bh + 32768;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          umma_x3s(tmem_base + s * 128, ah + ks * 4096, al + ks * 4096, bh + ks * 4096, bl + ks * 4096, 2048, 128, 2048, 128,
                   idesc, ks == 0);
        umma_commit(&b_empty[s]);
        umma_commit(&acc_full[s]);
        int64_t nti = ti, ntj = tj + 1;
        if (ntj >= a.nb) { ++nti; ntj = nti + a.bmd; }
        if (nti != ti || f + 1 >= f1) umma_commit(&a_empty);      // last tile that reads this A block
        ti = nti; tj = ntj;
      }
    }
  } else if (f0 < f1) {
    const int grp = (warp - 2) >> 2, q = warp & 3;          // epilogue warpgroup <-> TMEM stage; lane quarter
    float* stage = sStage + (warp - 2) * kPStageWarp;
    const float bias = __ldg(a.cls_b);
    int64_t ti, tj;
    flat_to_tile(f0, a.nb, a.bmd, ti, tj);
    for (int64_t f = f0, k = 0; f < f1; ++f, ++k) {
      if ((int)(k & 1) == grp) {
        const int64_t i_own = ti * 128 + q * 32 + lane;
        const float ui = i_own < a.n ? __ldg(a.u + i_own) : 0.f;
        mbar_wait(&acc_full[grp], (uint32_t)((k >> 1) & 1));
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + grp * 128;
        const int cc = lane & 15, half = lane >> 4;
        const int64_t full = a.n - a.md;
        const int64_t i_first = ti * 128 + q * 32, j_first = tj * 128;
        // interior tile: every (i, j) of this warp's 32 x 128 slab is a valid pair inside [p_begin, p_end) -> no checks
        const int64_t i_last = i_first + 31, j_last = j_first + 127;
        const bool interior = i_last < a.n && j_last < a.n && j_first >= i_last + a.md &&
                              tri_prefix(i_first, a.n, a.md) - (i_first + a.md) + j_first >= a.p_begin &&
                              tri_prefix(i_last, a.n, a.md) - (i_last + a.md) + j_last < a.p_end;
#pragma unroll 1
        for (int ch = 0; ch < 8; ++ch) {
          float v[16];
          tmem_ld16(taddr + ch * 16, v);
#pragma unroll
          for (int c4 = 0; c4 < 16; c4 += 4)
            *reinterpret_cast<float4*>(stage + lane * kPStageRow + c4) = make_float4(v[c4], v[c4 + 1], v[c4 + 2], v[c4 + 3]);
          __syncwarp();
          const int64_t j = j_first + ch * 16 + cc;
          const float ujb = (j < a.n ? __ldg(a.u + j) : 0.f) + bias;
          if (interior) {
            // p(i, j) = rowbase(i) + j with rowbase(i + 1) - rowbase(i) = full - i - 1: two rows per iteration
            int64_t i = i_first + half;
            float* dst = a.out + (tri_prefix(i, a.n, a.md) - (i + a.md) + j - a.p_begin);
#pragma unroll
            for (int it = 0; it < 16; ++it) {
              const int rr = it * 2 + half;
              const float acc = stage[rr * kPStageRow + cc];
              float val = __shfl_sync(0xffffffffu, ui, rr) + ujb - acc;
              if (a.apply_sigmoid) val = __fdividef(1.0f, 1.0f + __expf(-val));
              *dst = val;
              dst += 2 * (full - i) - 3;           // rowbase(i + 2) - rowbase(i)
              i += 2;
            }
          } else {
#pragma unroll 4
            for (int it = 0; it < 16; ++it) {
              const int rr = it * 2 + half;
              const float acc = stage[rr * kPStageRow + cc];
              const float ur = __shfl_sync(0xffffffffu, ui, rr);
              const int64_t i = i_first + rr;
              if (i < a.n && j < a.n && j >= i + a.md) {
                const int64_t p = tri_prefix(i, a.n, a.md) - (i + a.md) + j;
                if (p >= a.p_begin && p < a.p_end) {
                  float val = ur + ujb - acc;
                  if (a.apply_sigmoid) val = __fdividef(1.0f, 1.0f + __expf(-val));
                  a.out[p - a.p_begin] = val;
                }
              }
            }
          }
          __syncwarp();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[grp]);
      }
      if (++tj >= a.nb) { ++ti; tj = ti + a.bmd; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

}  // namespace
}  // namespace matcha

using namespace matcha;

extern "C" {

// workspace: PA | PB blocks (64 KB per 128 nodes each) | u
int64_t matcha_pair_tc_workspace_bytes(int64_t lo, int64_t hi) {
  const int64_t n = hi - lo;
  if (n <= 0) return 0;
  const int64_t nb = (n + 127) / 128;
  return 2 * nb * (int64_t)kPBlk + nb * 128 * 4 + 256;
}

int matcha_pair_tc_prepare(const float* D, const float* S, const float* cls_w, int32_t d, int64_t lo, int64_t hi,
                           void* workspace, int64_t workspace_bytes, void* stream) {
  MATCHA_REQUIRE(D && S && cls_w && workspace, "pair_tc_prepare: NULL argument");
  if (d != kPD) { set_error("pair_tc_prepare: embed_dim %d unsupported (64)", d); return MATCHA_ERR_UNSUPPORTED; }
  MATCHA_REQUIRE(((uintptr_t)workspace & 255) == 0, "pair_tc_prepare: workspace must be 256-byte aligned");
  const int64_t n = hi - lo;
  MATCHA_REQUIRE(n > 0 && workspace_bytes >= matcha_pair_tc_workspace_bytes(lo, hi), "pair_tc_prepare: workspace too small");
  const int64_t nb = (n + 127) / 128;
  uint8_t* PA = reinterpret_cast<uint8_t*>(workspace);
  uint8_t* PB = PA + nb * (int64_t)kPBlk;
  float* u = reinterpret_cast<float*>(PB + nb * (int64_t)kPBlk);
  const int64_t units = nb * 128 * 16;
  pair_prep_kernel<<<(unsigned)((units + 255) / 256), 256, 0, (cudaStream_t)stream>>>(D, S, cls_w, lo, n, nb, PA, PB, u);
  MATCHA_CHECK_LAUNCH("pair_prep");
  return MATCHA_OK;
}

int matcha_pair_tc_score_range(const void* workspace, const float* cls_b, int64_t lo, int64_t hi, int32_t min_dis,
                               int64_t p_begin, int64_t p_end, int32_t apply_sigmoid, float* out, void* stream) {
  MATCHA_REQUIRE(workspace && cls_b && out, "pair_tc_score_range: NULL argument");
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
  a.u = reinterpret_cast<const float*>(a.PB + nb * (int64_t)kPBlk);
  a.cls_b = cls_b; a.n = n; a.md = min_dis; a.nb = nb; a.bmd = bmd;
  a.f_begin = tri_prefix(r0 / 128, nb, bmd);
  a.f_end = tri_prefix(r1 / 128 + 1, nb, bmd);
  a.p_begin = p_begin; a.p_end = p_end; a.apply_sigmoid = apply_sigmoid; a.out = out;
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
