// Fused AdamW over the flat parameter buffer (torch.optim.AdamW defaults, main.py:630,671).
// Parameters whose gradient would be None in the reference on a given step (encoder weights of a
// chromosome absent from the batch, reconstruction heads of the chromosomes that were not drawn)
// are skipped entirely -- no decay, no moment update, their own step counter -- as torch does.
#include "common.cuh"

namespace matcha {
namespace {

struct AdamCfg {
  float lr, b1, b2, eps, wd, gscale;
};

__device__ __forceinline__ void adamw_elem(float& p, float g, float& m1, float& m2, const AdamCfg& c, float bc1, float bc2s) {
  g *= c.gscale;
  p *= (1.f - c.lr * c.wd);
  m1 = c.b1 * m1 + (1.f - c.b1) * g;
  m2 = c.b2 * m2 + (1.f - c.b2) * g * g;
  const float denom = sqrtf(m2) / bc2s + c.eps;
  p -= (c.lr / bc1) * (m1 / denom);
}

__global__ void adamw_always_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m1,
                                    float* __restrict__ m2, int64_t n, AdamCfg c, float bc1, float bc2s) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    adamw_elem(p[i], g[i], m1[i], m2[i], c, bc1, bc2s);
}

// one block-row per segment: blockIdx.y = segment
__global__ void adamw_seg_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m1,
                                 float* __restrict__ m2, const int64_t* __restrict__ seg_begin,
                                 const int64_t* __restrict__ seg_end, const int32_t* __restrict__ seg_flag,
                                 const int32_t* __restrict__ seg_step, const int32_t* __restrict__ active, AdamCfg c) {
  const int s = blockIdx.y;
  if (!active[seg_flag[s]]) return;
  const int step = seg_step[s] + 1;  // incremented afterwards by adamw_seg_step_kernel
  const float bc1 = 1.f - powf(c.b1, (float)step), bc2s = sqrtf(1.f - powf(c.b2, (float)step));
  const int64_t b = seg_begin[s], e = seg_end[s];
  for (int64_t i = b + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < e; i += (int64_t)gridDim.x * blockDim.x)
    adamw_elem(p[i], g[i], m1[i], m2[i], c, bc1, bc2s);
}
__global__ void adamw_seg_step_kernel(int n_seg, const int32_t* seg_flag, int32_t* seg_step, const int32_t* active) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n_seg && active[seg_flag[s]]) seg_step[s] += 1;
}
}  // namespace
}  // namespace matcha

using namespace matcha;

extern "C" int matcha_adamw(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n_always,
                            int32_t step, int32_t n_seg, const int64_t* seg_begin, const int64_t* seg_end,
                            const int32_t* seg_flag, int32_t* seg_step, const int32_t* active, float lr, float beta1,
                            float beta2, float eps, float weight_decay, float grad_scale, void* stream) {
  MATCHA_REQUIRE(params && grads && exp_avg && exp_avg_sq && step >= 1, "matcha_adamw: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  AdamCfg c{lr, beta1, beta2, eps, weight_decay, grad_scale};
  prof_begin(P_ADAMW, s);
  if (n_always > 0) {
    const float bc1 = 1.f - powf(beta1, (float)step), bc2s = sqrtf(1.f - powf(beta2, (float)step));
    int64_t blocks = (n_always + 255) / 256;
    if (blocks > kSMs * 8) blocks = kSMs * 8;
    adamw_always_kernel<<<(int)blocks, 256, 0, s>>>(params, grads, exp_avg, exp_avg_sq, n_always, c, bc1, bc2s);
    MATCHA_CHECK_LAUNCH("adamw_always");
  }
  if (n_seg > 0) {
    MATCHA_REQUIRE(seg_begin && seg_end && seg_flag && seg_step && active, "matcha_adamw: segment tables missing");
    dim3 grid(32, (unsigned)n_seg);
    adamw_seg_kernel<<<grid, 256, 0, s>>>(params, grads, exp_avg, exp_avg_sq, seg_begin, seg_end, seg_flag, seg_step,
                                          active, c);
    MATCHA_CHECK_LAUNCH("adamw_seg");
    adamw_seg_step_kernel<<<(n_seg + 127) / 128, 128, 0, s>>>(n_seg, seg_flag, seg_step, active);
    MATCHA_CHECK_LAUNCH("adamw_seg_step");
  }
  prof_end(P_ADAMW, n_seg > 0 ? 3 : 1, s);
  return MATCHA_OK;
}
