// Element-wise kernels of the training step's autograd functions (gecco_b200/training.py, BASELINE config 5): the Gaussian
// activation (models/activation.py:17-24) and the t-conditioned group normalisation AdaGN (models/normalization.py:36-44),
// forward and backward, each as ONE pass over the [clouds, rows, C] fp32 activations.  The autograd restatement through
// torch ops costs 6 + 10 passes over the 768-wide hidden tensor per activation and, for every AdaGN, two transposing copies
// around nn.functional.group_norm plus its own passes; the step was bound by exactly these copies and element-wise kernels
// (profiles/r4_train_step_torch_profiler.txt).  Everything here is HBM-bound streaming with 16-byte accesses.
//
//   gecco_train_gauss_act_fwd :  y = (exp(-x^2 / (2 alpha^2)) - 0.7) / 0.28           (alpha: DEVICE scalar, graph-safe)
//   gecco_train_gauss_act_bwd :  dx = dy * dy/dx,  dalpha partial per block (sum dy * dy/dalpha)
//   gecco_train_affine        :  out[b,n,c] = p[b,c] u[b,n,c] (+ q[b,c] w[b,n,c]) + r[b,c]
//                                 forward of the normalisation (u = x) and its input gradient (u = dy, w = x)
//   gecco_train_colsum2       :  partial sums over the rows of dy and of dy * x (the reductions of the backward); partials are
//                                 written per block and summed by the caller in a fixed order: no atomics, deterministic
#include "common.cuh"
#include "kernels.cuh"

namespace gecco {
namespace {

__device__ __forceinline__ float gauss_e(float x, float k) { return exp2f(x * x * k); }  // k = -log2(e) / (2 alpha^2)

__global__ void __launch_bounds__(256) gauss_fwd_kernel(const float* __restrict__ x, const float* __restrict__ alpha,
                                                        float* __restrict__ y, long long n, int normalized) {
  const float a = __ldg(alpha);
  const float k = -1.4426950408889634f / (2.f * a * a);
  const float m = normalized ? 1.0f / 0.28f : 1.f, c = normalized ? -0.7f / 0.28f : 0.f;
  const long long quads = n >> 2;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[q];
    float4 o;
    o.x = fmaf(gauss_e(v.x, k), m, c); o.y = fmaf(gauss_e(v.y, k), m, c);
    o.z = fmaf(gauss_e(v.z, k), m, c); o.w = fmaf(gauss_e(v.w, k), m, c);
    reinterpret_cast<float4*>(y)[q] = o;
  }
  const long long t = (quads << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) y[t] = fmaf(gauss_e(x[t], k), m, c);
}

__global__ void __launch_bounds__(256) gauss_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                        const float* __restrict__ alpha, float* __restrict__ dx,
                                                        float* __restrict__ dalpha, long long n, int normalized) {
  const float a = __ldg(alpha);
  const float k = -1.4426950408889634f / (2.f * a * a);
  const float m = normalized ? 1.0f / 0.28f : 1.f;
  const float cx = -m / (a * a), ca = m / (a * a * a);  // dy/dx = e * (-x / alpha^2) * m,  dy/dalpha = e * x^2 / alpha^3 * m
  float acc = 0.f;
  auto one = [&](float xv, float g) {
    const float e = gauss_e(xv, k) * g;
    acc = fmaf(e * xv, xv, acc);
    return e * xv * cx;
  };
  const long long quads = n >> 2;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[q];
    const float4 g = __ldcs(reinterpret_cast<const float4*>(dy) + q);
    float4 o;
    o.x = one(v.x, g.x); o.y = one(v.y, g.y); o.z = one(v.z, g.z); o.w = one(v.w, g.w);
    reinterpret_cast<float4*>(dx)[q] = o;
  }
  const long long t = (quads << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) dx[t] = one(x[t], dy[t]);
  // block sum of the alpha gradient, one atomic per block
  __shared__ float part[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += part[i];
    dalpha[blockIdx.x] = s * ca;  // one partial per block, summed by the caller in a fixed order (deterministic)
  }
}

// One thread owns four channels of ROWS_PER_BLOCK rows of a cloud: the per-(cloud, channel) coefficients stay in registers.
constexpr int AFF_ROWS = 32;

template <bool kTwo>
__global__ void __launch_bounds__(256) affine_kernel(const float* __restrict__ u, const float* __restrict__ w,
                                                     const float* __restrict__ p, const float* __restrict__ q,
                                                     const float* __restrict__ r, float* __restrict__ out, int rows_per_cloud,
                                                     int c4) {
  const int cloud = blockIdx.y;
  const int lane_c = threadIdx.x % c4, sub = threadIdx.x / c4, subs = blockDim.x / c4;
  if (sub >= subs) return;
  const long long cb = (long long)cloud * c4 + lane_c;
  const float4 pv = __ldg(reinterpret_cast<const float4*>(p) + cb);
  const float4 rv = __ldg(reinterpret_cast<const float4*>(r) + cb);
  float4 qv = make_float4(0.f, 0.f, 0.f, 0.f);
  if (kTwo) qv = __ldg(reinterpret_cast<const float4*>(q) + cb);
  const int row0 = blockIdx.x * AFF_ROWS;
  const int row1 = min(row0 + AFF_ROWS, rows_per_cloud);
  for (int row = row0 + sub; row < row1; row += subs) {
    const long long idx = ((long long)cloud * rows_per_cloud + row) * c4 + lane_c;
    const float4 a = __ldcs(reinterpret_cast<const float4*>(u) + idx);
    float4 o;
    o.x = fmaf(pv.x, a.x, rv.x); o.y = fmaf(pv.y, a.y, rv.y); o.z = fmaf(pv.z, a.z, rv.z); o.w = fmaf(pv.w, a.w, rv.w);
    if (kTwo) {
      const float4 b = __ldcs(reinterpret_cast<const float4*>(w) + idx);
      o.x = fmaf(qv.x, b.x, o.x); o.y = fmaf(qv.y, b.y, o.y); o.z = fmaf(qv.z, b.z, o.z); o.w = fmaf(qv.w, b.w, o.w);
    }
    reinterpret_cast<float4*>(out)[idx] = o;
  }
}

constexpr int COL_ROWS = 64;

__global__ void __launch_bounds__(256) colsum2_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                      float* __restrict__ out, int rows_per_cloud, int c4) {
  const int cloud = blockIdx.y;
  const int lane_c = threadIdx.x % c4, sub = threadIdx.x / c4, subs = blockDim.x / c4;
  if (sub >= subs) return;
  const int row0 = blockIdx.x * COL_ROWS;
  const int row1 = min(row0 + COL_ROWS, rows_per_cloud);
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
  for (int row = row0 + sub; row < row1; row += subs) {
    const long long idx = ((long long)cloud * rows_per_cloud + row) * c4 + lane_c;
    const float4 g = reinterpret_cast<const float4*>(dy)[idx];
    const float4 v = reinterpret_cast<const float4*>(x)[idx];
    s1.x += g.x; s1.y += g.y; s1.z += g.z; s1.w += g.w;
    s2.x = fmaf(g.x, v.x, s2.x); s2.y = fmaf(g.y, v.y, s2.y); s2.z = fmaf(g.z, v.z, s2.z); s2.w = fmaf(g.w, v.w, s2.w);
  }
  // one partial per (row chunk, row phase), summed by the caller in a fixed order (deterministic): [cloud][part][channel][2]
  const long long part = (long long)blockIdx.x * subs + sub, parts = (long long)gridDim.x * subs;
  float4* o = reinterpret_cast<float4*>(out + (((long long)cloud * parts + part) * c4 + lane_c) * 8);
  o[0] = make_float4(s1.x, s2.x, s1.y, s2.y);
  o[1] = make_float4(s1.z, s2.z, s1.w, s2.w);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Threads per block for a row of c4 float4 columns: as many whole rows as fit into 256 threads.
int block_for(int c4) { return (256 / c4) * c4; }

}  // namespace
}  // namespace gecco

extern "C" int gecco_train_gauss_act_fwd(const float* x, const float* alpha, float* y, int64_t n, int32_t normalized, void* stream) {
  using namespace gecco;
  GECCO_REQUIRE(x && alpha && y && n >= 0, "train_gauss_act_fwd: null argument");
  GECCO_REQUIRE(aligned16(x) && aligned16(y), "train_gauss_act_fwd: buffers must be 16-byte aligned");
  if (n == 0) return GECCO_OK;
  long long blocks = ((n + 3) / 4 + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  gauss_fwd_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, alpha, y, n, normalized);
  GECCO_CHECK_LAUNCH("gauss_fwd_kernel");
  return GECCO_OK;
}

static long long gauss_bwd_blocks(long long n) {
  long long blocks = ((n + 3) / 4 + 255) / 256;
  const long long cap = (long long)gecco::sm_count() * 16;
  return blocks > cap ? cap : blocks;
}
extern "C" int64_t gecco_train_gauss_act_bwd_parts(int64_t n) { return n <= 0 ? 0 : gauss_bwd_blocks(n); }

extern "C" int gecco_train_gauss_act_bwd(const float* x, const float* dy, const float* alpha, float* dx, float* dalpha, int64_t n,
                                         int32_t normalized, void* stream) {
  using namespace gecco;
  GECCO_REQUIRE(x && dy && alpha && dx && dalpha && n >= 0, "train_gauss_act_bwd: null argument");
  GECCO_REQUIRE(aligned16(x) && aligned16(dy) && aligned16(dx), "train_gauss_act_bwd: buffers must be 16-byte aligned");
  if (n == 0) return GECCO_OK;
  const long long blocks = gauss_bwd_blocks(n);
  gauss_bwd_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, dy, alpha, dx, dalpha, n, normalized);
  GECCO_CHECK_LAUNCH("gauss_bwd_kernel");
  return GECCO_OK;
}

extern "C" int gecco_train_affine(const float* u, const float* w, const float* p, const float* q, const float* r, float* out,
                                  int32_t clouds, int32_t rows_per_cloud, int32_t c, void* stream) {
  using namespace gecco;
  GECCO_REQUIRE(u && p && r && out, "train_affine: null argument");
  GECCO_REQUIRE((w == nullptr) == (q == nullptr), "train_affine: the second input and its coefficients come together");
  GECCO_REQUIRE(c > 0 && c % 4 == 0 && c <= 1024, "train_affine: c must be a multiple of 4, at most 1024 (got %d)", c);
  GECCO_REQUIRE(aligned16(u) && aligned16(w) && aligned16(p) && aligned16(q) && aligned16(r) && aligned16(out),
                "train_affine: buffers must be 16-byte aligned");
  if (clouds <= 0 || rows_per_cloud <= 0) return GECCO_OK;
  GECCO_REQUIRE(clouds <= 65535, "train_affine: at most 65535 clouds");
  const int c4 = c / 4;
  const dim3 grid((unsigned)ceil_div(rows_per_cloud, AFF_ROWS), (unsigned)clouds);
  if (w != nullptr)
    affine_kernel<true><<<grid, block_for(c4), 0, static_cast<cudaStream_t>(stream)>>>(u, w, p, q, r, out, rows_per_cloud, c4);
  else
    affine_kernel<false><<<grid, block_for(c4), 0, static_cast<cudaStream_t>(stream)>>>(u, w, p, q, r, out, rows_per_cloud, c4);
  GECCO_CHECK_LAUNCH("affine_kernel");
  return GECCO_OK;
}

extern "C" int32_t gecco_train_colsum2_parts(int32_t rows_per_cloud, int32_t c) {
  using namespace gecco;
  if (rows_per_cloud <= 0 || c <= 0 || c % 4 != 0 || c > 1024) return 0;
  return ceil_div(rows_per_cloud, COL_ROWS) * (256 / (c / 4));
}

extern "C" int gecco_train_colsum2(const float* dy, const float* x, float* out, int32_t clouds, int32_t rows_per_cloud, int32_t c,
                                   void* stream) {
  using namespace gecco;
  GECCO_REQUIRE(dy && x && out, "train_colsum2: null argument");
  GECCO_REQUIRE(aligned16(out), "train_colsum2: buffers must be 16-byte aligned");
  GECCO_REQUIRE(c > 0 && c % 4 == 0 && c <= 1024, "train_colsum2: c must be a multiple of 4, at most 1024 (got %d)", c);
  GECCO_REQUIRE(aligned16(dy) && aligned16(x), "train_colsum2: buffers must be 16-byte aligned");
  if (clouds <= 0 || rows_per_cloud <= 0) return GECCO_OK;
  GECCO_REQUIRE(clouds <= 65535, "train_colsum2: at most 65535 clouds");
  const int c4 = c / 4;
  const dim3 grid((unsigned)ceil_div(rows_per_cloud, COL_ROWS), (unsigned)clouds);
  colsum2_kernel<<<grid, block_for(c4), 0, static_cast<cudaStream_t>(stream)>>>(dy, x, out, rows_per_cloud, c4);
  GECCO_CHECK_LAUNCH("colsum2_kernel");
  return GECCO_OK;
}
