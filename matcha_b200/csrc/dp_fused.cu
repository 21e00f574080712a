// Data-parallel step boundary in ONE kernel over NVLink peer memory (the reference is single-process):
//   cross-GPU barrier  ->  gradient all-reduce (every rank reads all peers' flat gradient buffers through P2P loads and adds
//   them in rank order, so all replicas compute bit-identical sums)  ->  mean  ->  AdamW on the rank's own replica
//   ->  cross-GPU barrier  ->  zero the own gradient buffer for the next step.
// It replaces   copy flags -> ncclAllReduce(gflat) -> copy flags -> 3 AdamW launches -> gflat.zero_()   (7 launches, the
// collective exposed on the main stream) with one launch whose transfer overlaps its arithmetic chunk by chunk.
// Two forms.  "One-shot" (2 ranks): each element is read once from every peer (world x 4 B over NVLink per element).
// "Two-shot" (4 and 8 ranks; MATCHA_DP_TWO_SHOT=0 / 1 forces one): the gradient buffer is 2.5 MB (cfg2) .. 66 MB (cfg5), so a
// one-shot rank would pull 17 .. 460 MB per step at 8 GPUs; instead rank r reduces the r-th part of every block's slice
// (reading it from all peers), writes the mean into ALL replicas' gradient buffers in place (only the owner of a part ever
// reads it, so the overwrite races with nothing), and after the second barrier every rank runs AdamW on its own buffer and
// clears it in the same pass: 2 (world - 1) / world x 4 B per element over NVLink -- a quarter of the one-shot traffic at 8.
//
// Barriers are per block: block b of every rank signals block b of every peer (a monotonically increasing epoch written
// into the peer's flag array with release semantics) and waits for theirs; block b of every rank works on the same slice
// of the buffers, so the second barrier tells a rank that its slice is no longer being read before it zeroes it.
// Spins are bounded: a dead peer traps this context instead of hanging the GPU.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace matcha {
namespace {

constexpr int kDpMaxWorld = 8;
constexpr int kDpThreads = 512;
constexpr int kDpMaxSeg = 4 * MATCHA_MAX_CHROM;

struct DpArgs {
  int world, rank;
  const float* g[kDpMaxWorld];          // flat gradient buffers of all ranks (peer pointers), g[rank] = own
  const int32_t* act[kDpMaxWorld];      // activity flags of all ranks [n_flags]
  uint32_t* bar[kDpMaxWorld];           // barrier flag arrays of all ranks: [2][gridDim.x][kDpMaxWorld]
  float *p, *m1, *m2, *g_own;
  int32_t* act_red;                     // own: OR over ranks [n_flags]
  int64_t n_always, n_flat;
  int n_seg, n_flags;
  const int64_t *seg_begin, *seg_end;
  const int32_t* seg_flag;
  const int32_t* seg_step;
  float lr, b1, b2, eps, wd;
  int step;
  uint32_t epoch;
  int two_shot;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void peer_barrier(const DpArgs& a, int phase, uint32_t epoch) {
  if (a.world == 1) { __syncthreads(); return; }      // single replica: the same kernel is the fused AdamW + clear
  __syncthreads();
  if ((int)threadIdx.x < a.world) {
    const int peer = threadIdx.x;
    const int64_t slot = ((int64_t)phase * gridDim.x + blockIdx.x) * kDpMaxWorld;
    __threadfence_system();
    st_release_sys(a.bar[peer] + slot + a.rank, epoch);
    const uint32_t* mine = a.bar[a.rank] + slot + peer;
    uint32_t spins = 0;
    while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
      if (++spins > (1u << 27)) __trap();          // a peer never arrived (several seconds): fail loudly, do not hang
      __nanosleep(64);
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kDpThreads) dp_reduce_adamw_kernel(const DpArgs a) {
  __shared__ float s_bc1[kDpMaxSeg], s_bc2s[kDpMaxSeg];
  __shared__ int64_t s_beg[kDpMaxSeg], s_end[kDpMaxSeg];
  __shared__ uint8_t s_on[kDpMaxSeg];
  // per-segment activity (OR over ranks) and bias corrections; reads of the peers' flags wait for the first barrier
  peer_barrier(a, 0, a.epoch);
  for (int s = threadIdx.x; s < a.n_seg; s += kDpThreads) {
    const int f = a.seg_flag[s];
    int on = 0;
    for (int r = 0; r < a.world; ++r) on |= a.act[r][f];
    const int step = a.seg_step[s] + 1;
    s_on[s] = on != 0;
    s_bc1[s] = 1.f - powf(a.b1, (float)step);
    s_bc2s[s] = sqrtf(1.f - powf(a.b2, (float)step));
    s_beg[s] = a.seg_begin[s];
    s_end[s] = a.seg_end[s];
  }
  if (blockIdx.x == 0)
    for (int f = threadIdx.x; f < a.n_flags; f += kDpThreads) {
      int on = 0;
      for (int r = 0; r < a.world; ++r) on |= a.act[r][f];
      a.act_red[f] = on;
    }
  __syncthreads();
  const float bc1a = 1.f - powf(a.b1, (float)a.step), bc2sa = sqrtf(1.f - powf(a.b2, (float)a.step));
  const float inv_world = 1.0f / (float)a.world;
  // this block's slice (the same on every rank), in float4 units
  const int64_t n4 = (a.n_flat + 3) / 4;
  const int64_t per = (n4 + gridDim.x - 1) / gridDim.x;
  const int64_t q0 = (int64_t)blockIdx.x * per, q1 = (q0 + per < n4) ? q0 + per : n4;
  auto segment = [&](int64_t i, float& bc1, float& bc2s) -> bool {
    if (i < a.n_always) { bc1 = bc1a; bc2s = bc2sa; return true; }
    int lo = 0, hi = a.n_seg;                      // last segment with begin <= i (segments are sorted, 64-element aligned)
    while (hi - lo > 1) { const int m = (lo + hi) >> 1; if (s_beg[m] <= i) lo = m; else hi = m; }
    bc1 = s_bc1[lo]; bc2s = s_bc2s[lo];
    return a.n_seg > 0 && i >= s_beg[lo] && i < s_end[lo] && s_on[lo];
  };
  auto adamw4 = [&](int64_t i, const float4 g, float scale, float bc1, float bc2s) {
    float4 p = *reinterpret_cast<float4*>(a.p + i), m1 = *reinterpret_cast<float4*>(a.m1 + i), m2 = *reinterpret_cast<float4*>(a.m2 + i);
    float* pp = &p.x; float* pm1 = &m1.x; float* pm2 = &m2.x; const float* pg = &g.x;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float gg = pg[e] * scale;
      pp[e] *= (1.f - a.lr * a.wd);
      pm1[e] = a.b1 * pm1[e] + (1.f - a.b1) * gg;
      pm2[e] = a.b2 * pm2[e] + (1.f - a.b2) * gg * gg;
      const float denom = sqrtf(pm2[e]) / bc2s + a.eps;
      pp[e] -= (a.lr / bc1) * (pm1[e] / denom);
    }
    *reinterpret_cast<float4*>(a.p + i) = p;
    *reinterpret_cast<float4*>(a.m1 + i) = m1;
    *reinterpret_cast<float4*>(a.m2 + i) = m2;
  };
  if (a.two_shot && a.world > 1) {
    // reduce-scatter + all-gather through the peers' buffers: this rank owns part `rank` of the block's slice
    const int64_t len = q1 > q0 ? q1 - q0 : 0;
    const int64_t s0 = q0 + len * a.rank / a.world, s1 = q0 + len * (a.rank + 1) / a.world;
    for (int64_t q = s0 + threadIdx.x; q < s1; q += kDpThreads) {
      const int64_t i = q * 4;
      float bc1, bc2s;
      if (!segment(i, bc1, bc2s)) continue;
      float4 g = *reinterpret_cast<const float4*>(a.g[0] + i);
      for (int r = 1; r < a.world; ++r) {          // rank order: one sum, written to every replica
        const float4 v = *reinterpret_cast<const float4*>(a.g[r] + i);
        g.x += v.x; g.y += v.y; g.z += v.z; g.w += v.w;
      }
      g.x *= inv_world; g.y *= inv_world; g.z *= inv_world; g.w *= inv_world;
      for (int r = 0; r < a.world; ++r) *reinterpret_cast<float4*>(const_cast<float*>(a.g[r]) + i) = g;
    }
    peer_barrier(a, 1, a.epoch);                   // every part of this slice has landed in every replica's buffer
    for (int64_t q = q0 + threadIdx.x; q < q1; q += kDpThreads) {
      const int64_t i = q * 4;
      float bc1, bc2s;
      if (segment(i, bc1, bc2s)) adamw4(i, *reinterpret_cast<const float4*>(a.g_own + i), 1.f, bc1, bc2s);
      *reinterpret_cast<float4*>(a.g_own + i) = make_float4(0.f, 0.f, 0.f, 0.f);       // nobody reads it any more this step
    }
    return;
  }
  for (int64_t q = q0 + threadIdx.x; q < q1; q += kDpThreads) {
    const int64_t i = q * 4;
    float bc1, bc2s;
    if (!segment(i, bc1, bc2s)) continue;
    float4 g = *reinterpret_cast<const float4*>(a.g[0] + i);
    for (int r = 1; r < a.world; ++r) {            // rank order: identical sums on every replica
      const float4 v = *reinterpret_cast<const float4*>(a.g[r] + i);
      g.x += v.x; g.y += v.y; g.z += v.z; g.w += v.w;
    }
    adamw4(i, g, inv_world, bc1, bc2s);
  }
  // every rank has finished reading this slice of every buffer: clear the own one for the next step's accumulation
  peer_barrier(a, 1, a.epoch);
  for (int64_t q = q0 + threadIdx.x; q < q1; q += kDpThreads) *reinterpret_cast<float4*>(a.g_own + q * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
}

__global__ void dp_seg_step_kernel(int n_seg, const int32_t* seg_flag, int32_t* seg_step, const int32_t* act_red) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n_seg && act_red[seg_flag[s]]) seg_step[s] += 1;
}

}  // namespace
}  // namespace matcha

using namespace matcha;

extern "C" {

static int g_two_shot = -1;
/* form of the data-parallel step boundary: 0 one-shot, 1 two-shot, 2 by world size (default) */
void matcha_set_dp_two_shot(int32_t mode) { g_two_shot = mode < 0 ? -1 : (mode > 2 ? 2 : mode); }
int32_t matcha_dp_blocks(void) { return kSMs; }
/* kernels of the current device may dereference memory of `peer_device` afterwards (idempotent) */
int matcha_enable_peer_access(int32_t peer_device) {
  int dev = 0;
  if (int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return rc;
  if (dev == peer_device) return MATCHA_OK;
  int can = 0;
  if (int rc = check_cuda(cudaDeviceCanAccessPeer(&can, dev, peer_device), "cudaDeviceCanAccessPeer")) return rc;
  if (!can) { set_error("device %d cannot map memory of device %d", dev, (int)peer_device); return MATCHA_ERR_UNSUPPORTED; }
  const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return MATCHA_OK; }
  return check_cuda(e, "cudaDeviceEnablePeerAccess");
}
/* CUDA IPC plumbing for the peer buffers: export the handle of a cudaMalloc allocation (its BASE pointer), map a peer's
 * allocation into this process for kernels of the CURRENT device (lazy peer access over NVLink) */
int matcha_ipc_get_handle(const void* base_ptr, uint8_t* handle64) {
  MATCHA_REQUIRE(base_ptr && handle64, "matcha_ipc_get_handle: NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  cudaIpcMemHandle_t h;
  if (int rc = check_cuda(cudaIpcGetMemHandle(&h, const_cast<void*>(base_ptr)), "cudaIpcGetMemHandle")) return rc;
  memcpy(handle64, &h, 64);
  return MATCHA_OK;
}
int matcha_ipc_open(const uint8_t* handle64, void** mapped_base) {
  MATCHA_REQUIRE(handle64 && mapped_base, "matcha_ipc_open: NULL argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  return check_cuda(cudaIpcOpenMemHandle(mapped_base, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
}
int matcha_ipc_close(void* mapped_base) { return check_cuda(cudaIpcCloseMemHandle(mapped_base), "cudaIpcCloseMemHandle"); }
int64_t matcha_dp_barrier_bytes(void) { return (int64_t)2 * kSMs * kDpMaxWorld * sizeof(uint32_t); }

int matcha_dp_reduce_adamw(int32_t world, int32_t rank, const void* const* grad_ptrs, const void* const* active_ptrs,
                           void* const* barrier_ptrs, float* params, float* exp_avg, float* exp_avg_sq, int32_t* active_reduced,
                           int64_t n_always, int64_t n_flat, int32_t n_seg, int32_t n_flags, const int64_t* seg_begin,
                           const int64_t* seg_end, const int32_t* seg_flag, int32_t* seg_step, int32_t step, uint32_t epoch, float lr,
                           float beta1, float beta2, float eps, float weight_decay, void* stream) {
  MATCHA_REQUIRE(world >= 1 && world <= kDpMaxWorld && rank >= 0 && rank < world, "matcha_dp_reduce_adamw: world=%d rank=%d (1..%d ranks)",
                 (int)world, (int)rank, kDpMaxWorld);
  MATCHA_REQUIRE(grad_ptrs && active_ptrs && (world == 1 || barrier_ptrs) && params && exp_avg && exp_avg_sq && active_reduced && step >= 1 && epoch >= 1,
                 "matcha_dp_reduce_adamw: NULL argument");
  MATCHA_REQUIRE(n_seg >= 0 && n_seg <= kDpMaxSeg && n_flat % 4 == 0 && n_always % 4 == 0, "matcha_dp_reduce_adamw: n_seg=%d (<= %d), buffers 4-float aligned",
                 (int)n_seg, kDpMaxSeg);
  DpArgs a;
  a.world = world; a.rank = rank;
  for (int r = 0; r < world; ++r) {
    MATCHA_REQUIRE(grad_ptrs[r] && active_ptrs[r] && (world == 1 || barrier_ptrs[r]), "matcha_dp_reduce_adamw: peer %d pointers missing", r);
    a.g[r] = (const float*)grad_ptrs[r]; a.act[r] = (const int32_t*)active_ptrs[r]; a.bar[r] = world == 1 ? nullptr : (uint32_t*)barrier_ptrs[r];
  }
  a.p = params; a.m1 = exp_avg; a.m2 = exp_avg_sq; a.g_own = const_cast<float*>(a.g[rank]);
  a.act_red = active_reduced; a.n_always = n_always; a.n_flat = n_flat; a.n_seg = n_seg; a.n_flags = n_flags;
  a.seg_begin = seg_begin; a.seg_end = seg_end; a.seg_flag = seg_flag; a.seg_step = seg_step;
  a.lr = lr; a.b1 = beta1; a.b2 = beta2; a.eps = eps; a.wd = weight_decay; a.step = step; a.epoch = epoch;
  if (g_two_shot < 0) {            // by world size (two-shot from 4 ranks on) unless MATCHA_DP_TWO_SHOT=0 / 1 forces a form
    const char* e = getenv("MATCHA_DP_TWO_SHOT");
    g_two_shot = !e ? 2 : (e[0] == '0' ? 0 : 1);
  }
  a.two_shot = g_two_shot == 2 ? (world >= 4) : g_two_shot;
  cudaStream_t s = (cudaStream_t)stream;
  prof_begin(P_ADAMW, s);
  dp_reduce_adamw_kernel<<<kSMs, kDpThreads, 0, s>>>(a);
  MATCHA_CHECK_LAUNCH("dp_reduce_adamw");
  if (n_seg > 0) {
    dp_seg_step_kernel<<<(n_seg + 127) / 128, 128, 0, s>>>(n_seg, seg_flag, seg_step, active_reduced);
    MATCHA_CHECK_LAUNCH("dp_seg_step");
  }
  prof_end(P_ADAMW, 2, s);
  return MATCHA_OK;
}

}  // extern "C"
