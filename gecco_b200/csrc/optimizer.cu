// Training-step tail (config 5): Adam (the reference's optimiser, diffusion.py:207-208 -> torch.optim.Adam defaults) and
// the exponential moving average of the weights (ema.py:187-194: ema = decay * ema + (1 - decay) * p after the
// optimiser step) as ONE pass over flat fp32 buffers -- parameters, gradients, both moments and the EMA copy are each one
// contiguous allocation (gecco_b200/training.py lays them out), so a step is a single launch that streams 5 arrays in
// and 4 out with 16-byte accesses (HBM-bound: 36 B per parameter).  The gradient scale (1 / world size after the
// summing all-reduce, or a loss-scale inverse) is applied on the fly.
#include "common.cuh"
#include "kernels.cuh"

namespace gecco {
namespace {

struct AdamParams {
  float lr_over_c1;    // lr / (1 - beta1^t)
  float inv_sqrt_c2;   // 1 / sqrt(1 - beta2^t)
  float beta1, beta2, omb1, omb2;  // 1 - beta computed in double on the host, like torch's python scalars
  float eps, grad_scale, ema_decay, omd;
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float* e, const AdamParams& a) {
  g *= a.grad_scale;
  m = fmaf(a.beta1, m, a.omb1 * g);          // exp_avg.lerp_(grad, 1 - beta1)
  v = fmaf(a.beta2, v, a.omb2 * g * g);      // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  const float denom = sqrtf(v) * a.inv_sqrt_c2 + a.eps;
  p -= a.lr_over_c1 * (m / denom);
  if (e != nullptr) *e = fmaf(a.ema_decay, *e, a.omd * p);
}

__global__ void __launch_bounds__(256) adam_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                       float* __restrict__ v, float* __restrict__ ema, long long n,
                                                       const AdamParams a) {
  const long long quads = n >> 2;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += (long long)gridDim.x * blockDim.x) {
    float4 p4 = reinterpret_cast<float4*>(p)[q];
    const float4 g4 = __ldcs(reinterpret_cast<const float4*>(g) + q);  // gradients are dead after the step: stream them
    float4 m4 = reinterpret_cast<float4*>(m)[q];
    float4 v4 = reinterpret_cast<float4*>(v)[q];
    float4 e4 = ema ? reinterpret_cast<float4*>(ema)[q] : make_float4(0.f, 0.f, 0.f, 0.f);
    adam_one(p4.x, g4.x, m4.x, v4.x, ema ? &e4.x : nullptr, a);
    adam_one(p4.y, g4.y, m4.y, v4.y, ema ? &e4.y : nullptr, a);
    adam_one(p4.z, g4.z, m4.z, v4.z, ema ? &e4.z : nullptr, a);
    adam_one(p4.w, g4.w, m4.w, v4.w, ema ? &e4.w : nullptr, a);
    reinterpret_cast<float4*>(p)[q] = p4;
    reinterpret_cast<float4*>(m)[q] = m4;
    reinterpret_cast<float4*>(v)[q] = v4;
    if (ema) reinterpret_cast<float4*>(ema)[q] = e4;
  }
  // tail (n not a multiple of 4)
  const long long t = (quads << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) adam_one(p[t], g[t], m[t], v[t], ema ? ema + t : nullptr, a);
}

}  // namespace
}  // namespace gecco

extern "C" int gecco_adam_ema_step(float* p, const float* g, float* m, float* v, float* ema, int64_t n, int64_t step, double lr,
                                   double beta1, double beta2, double eps, double grad_scale, double ema_decay, void* stream) {
  using namespace gecco;
  GECCO_REQUIRE(p && g && m && v && n >= 0, "adam_ema_step: null argument");
  GECCO_REQUIRE(step >= 1, "adam_ema_step: step counts from 1 (got %lld)", (long long)step);
  GECCO_REQUIRE(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0, "adam_ema_step: betas must be in [0, 1)");
  GECCO_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                  reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(ema)) & 15) == 0,
                "adam_ema_step: buffers must be 16-byte aligned");
  if (n == 0) return GECCO_OK;
  AdamParams a;
  // bias corrections in double like torch's scalar path (torch/optim/adam.py, _single_tensor_adam)
  const double c1 = 1.0 - pow(beta1, (double)step), c2 = 1.0 - pow(beta2, (double)step);
  a.lr_over_c1 = static_cast<float>(lr / c1);
  a.inv_sqrt_c2 = static_cast<float>(1.0 / sqrt(c2));
  a.beta1 = (float)beta1; a.beta2 = (float)beta2; a.eps = (float)eps; a.grad_scale = (float)grad_scale; a.ema_decay = (float)ema_decay;
  a.omb1 = static_cast<float>(1.0 - beta1); a.omb2 = static_cast<float>(1.0 - beta2);
  a.omd = static_cast<float>(1.0 - ema_decay);
  const long long quads = (n + 3) / 4;
  long long blocks = (quads + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  adam_ema_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(p, g, m, v, ema, n, a);
  GECCO_CHECK_LAUNCH("adam_ema_kernel");
  return GECCO_OK;
}
