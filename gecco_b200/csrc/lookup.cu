// Projective feature lookup (models/ray.py:64-87): reparametrise the EDM-scaled points to data
// space, project them through the camera, bilinearly gather every level of the channels-last
// feature pyramid.  One warp per point; a lane owns fixed 8-channel chunks (one 16 B load per tap),
// so the [points, sum C] output row is written with fully coalesced 16 B stores and the GroupNorm
// statistics of models/ray.py:53 are accumulated per lane and reduced once per CTA.
// Also: folding of that GroupNorm into per-cloud projection weights, and the NCHW fp32 ->
// NHWC bf16 repack of the pyramid.
#include "common.cuh"
#include "ptx.cuh"

namespace gecco {

namespace {

constexpr int LK_WARPS = 8;
constexpr int LK_MAXCH = 4;          // 8-channel chunks per lane -> up to 1024 channels
constexpr int LK_POINTS_PER_WARP = 16;

struct LookupP {
  const float* xin;
  const float* sigma;
  int sigma_stride;
  float sigma_data;
  int reparam;
  float mean[3], rsig[3], logit_scale;
  const float* K;
  int n_levels;
  const __nv_bfloat16* lvl_ptr[GECCO_MAX_LEVELS];
  int lvl_h[GECCO_MAX_LEVELS], lvl_w[GECCO_MAX_LEVELS], lvl_c[GECCO_MAX_LEVELS];
  int points, rows_per_cloud, ctot;
  __nv_bfloat16* out16;
  long long ldo16;
  float* out32;
  long long ldo32;
  double* stats;
  int stat_groups;
};

__device__ __forceinline__ void bf16x8_to_float(const uint4& q, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

__global__ void __launch_bounds__(LK_WARPS * 32) lookup_kernel(const LookupP p) {
  extern __shared__ float sgrp[];  // [stat_groups][2]
  const int cloud = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (p.stats != nullptr) {
    for (int i = threadIdx.x; i < p.stat_groups * 2; i += blockDim.x) sgrp[i] = 0.f;
    __syncthreads();
  }
  // fixed chunk -> (level, channel) assignment of this lane
  const __nv_bfloat16* ch_ptr[LK_MAXCH];
  int ch_h[LK_MAXCH], ch_w[LK_MAXCH], ch_c[LK_MAXCH], ch_col[LK_MAXCH];
#pragma unroll
  for (int k = 0; k < LK_MAXCH; ++k) {
    const int col = (lane + 32 * k) * 8;
    ch_col[k] = col;
    ch_ptr[k] = nullptr;
    ch_h[k] = ch_w[k] = ch_c[k] = 0;
    int off = 0;
    for (int l = 0; l < p.n_levels; ++l) {
      if (col >= off && col < off + p.lvl_c[l]) {
        ch_h[k] = p.lvl_h[l];
        ch_w[k] = p.lvl_w[l];
        ch_c[k] = p.lvl_c[l];
        ch_ptr[k] = p.lvl_ptr[l] + (long long)cloud * p.lvl_h[l] * p.lvl_w[l] * p.lvl_c[l] + (col - off);
      }
      off += p.lvl_c[l];
    }
  }
  float s1[LK_MAXCH][8], s2[LK_MAXCH][8];
#pragma unroll
  for (int k = 0; k < LK_MAXCH; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) s1[k][j] = s2[k][j] = 0.f;

  float c_in = 1.f;
  if (p.sigma != nullptr) {
    const float sg = __ldg(p.sigma + (long long)cloud * p.sigma_stride);
    c_in = 1.0f / sqrtf(p.sigma_data * p.sigma_data + sg * sg);
  }
  const float* Kc = p.K + (long long)cloud * 9;
  const float fx = __ldg(Kc + 0), cx = __ldg(Kc + 2), fy = __ldg(Kc + 4), cy = __ldg(Kc + 5);

  for (int pt = blockIdx.x * LK_WARPS + warp; pt < p.points; pt += gridDim.x * LK_WARPS) {
    const float* xp = p.xin + ((long long)cloud * p.points + pt) * 3;
    float g[3] = {c_in * __ldg(xp), c_in * __ldg(xp + 1), c_in * __ldg(xp + 2)};
    float d[3];
    if (p.reparam == 1) {  // GaussianReparam.diffusion_to_data (reparam.py:62-64)
      for (int j = 0; j < 3; ++j) d[j] = g[j] * p.rsig[j] + p.mean[j];
    } else if (p.reparam == 2) {  // UVLReparam.diffusion_to_data (reparam.py:166-201)
      const float u0 = g[0] * p.rsig[0] + p.mean[0];
      const float v0 = g[1] * p.rsig[1] + p.mean[1];
      const float l0 = g[2] * p.rsig[2] + p.mean[2];
      const float h = (tanhf(u0) * p.logit_scale + 1.0f) / 2.0f;
      const float w = (tanhf(v0) * p.logit_scale + 1.0f) / 2.0f;
      const float dep = expf(l0);
      const float x = (h - cx) / fx, y = (w - cy) / fy;
      float nrm = sqrtf(x * x + y * y + 1.0f);
      nrm = fmaxf(nrm, 1e-12f);
      d[0] = x / nrm * dep;
      d[1] = y / nrm * dep;
      d[2] = 1.0f / nrm * dep;
    } else {
      for (int j = 0; j < 3; ++j) d[j] = g[j];
    }
    // kornia project_points (models/ray.py:74)
    const float z = d[2];
    const float sc = (fabsf(z) > 1e-8f) ? 1.0f / (z + 1e-8f) : 1.0f;
    const float u = d[0] * sc * fx + cx;
    const float v = d[1] * sc * fy + cy;
    // grid_sample(align_corners=False) on grid = 2*uv - 1 (models/ray.py:80-82)
    const float gx = u * 2.0f - 1.0f, gy = v * 2.0f - 1.0f;

    const long long orow = (long long)cloud * p.rows_per_cloud + pt;
#pragma unroll
    for (int k = 0; k < LK_MAXCH; ++k) {
      if (ch_ptr[k] == nullptr) continue;
      const int H = ch_h[k], W = ch_w[k], C = ch_c[k];
      const float ix = ((gx + 1.0f) * W - 1.0f) / 2.0f;
      const float iy = ((gy + 1.0f) * H - 1.0f) / 2.0f;
      const float x0 = floorf(ix), y0 = floorf(iy);
      const float x1 = x0 + 1.0f, y1 = y0 + 1.0f;
      const bool vx0 = x0 >= 0.f && x0 <= (float)(W - 1), vx1 = x1 >= 0.f && x1 <= (float)(W - 1);
      const bool vy0 = y0 >= 0.f && y0 <= (float)(H - 1), vy1 = y1 >= 0.f && y1 <= (float)(H - 1);
      const float w_nw = (x1 - ix) * (y1 - iy), w_ne = (ix - x0) * (y1 - iy);
      const float w_sw = (x1 - ix) * (iy - y0), w_se = (ix - x0) * (iy - y0);
      const int xi0 = vx0 ? (int)x0 : 0, xi1 = vx1 ? (int)x1 : 0, yi0 = vy0 ? (int)y0 : 0, yi1 = vy1 ? (int)y1 : 0;
      uint4 q[4];
      const bool tv[4] = {vx0 && vy0, vx1 && vy0, vx0 && vy1, vx1 && vy1};
      const int to[4] = {(yi0 * W + xi0) * C, (yi0 * W + xi1) * C, (yi1 * W + xi0) * C, (yi1 * W + xi1) * C};
#pragma unroll
      for (int t = 0; t < 4; ++t)
        q[t] = tv[t] ? __ldg(reinterpret_cast<const uint4*>(ch_ptr[k] + to[t])) : make_uint4(0, 0, 0, 0);
      const float tw[4] = {w_nw, w_ne, w_sw, w_se};
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        if (tv[t]) {
          float f[8];
          bf16x8_to_float(q[t], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += f[j] * tw[t];
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s1[k][j] += acc[j];
        s2[k][j] += acc[j] * acc[j];
      }
      if (p.out16 != nullptr) {
        *reinterpret_cast<uint4*>(p.out16 + orow * p.ldo16 + ch_col[k]) =
            make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]),
                       pack_bf16x2(acc[6], acc[7]));
      }
      if (p.out32 != nullptr) {
        float4* o = reinterpret_cast<float4*>(p.out32 + orow * p.ldo32 + ch_col[k]);
        o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
      }
    }
  }
  if (p.stats != nullptr) {
    const int gsz = p.ctot / p.stat_groups;
#pragma unroll
    for (int k = 0; k < LK_MAXCH; ++k) {
      if (ch_ptr[k] == nullptr) continue;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int g = (ch_col[k] + j) / gsz;
        atomicAdd(&sgrp[g * 2], s1[k][j]);
        atomicAdd(&sgrp[g * 2 + 1], s2[k][j]);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < p.stat_groups * 2; i += blockDim.x)
      atomicAdd(p.stats + (long long)cloud * p.stat_groups * 2 + i, static_cast<double>(sgrp[i]));
  }
}

// GroupNorm (no affine) followed by Linear, folded per cloud (models/ray.py:52-55):
//   Linear(GN(z))[o] = sum_c (W[o,c] rstd_g(c)) z[c] + (b[o] - sum_c W[o,c] mean_g(c) rstd_g(c))
__global__ void fold_gn_kernel(const float* __restrict__ W, const float* __restrict__ bias,
                               const double* __restrict__ stats, double count, float eps, int groups, int c_in,
                               int c_out, __nv_bfloat16* __restrict__ wb, long long ldwb, float* __restrict__ bb) {
  __shared__ float red[32];
  const int o = blockIdx.x, cloud = blockIdx.y;
  const int gs = c_in / groups;
  const double* cs = stats + (long long)cloud * groups * 2;
  float acc = 0.f;
  for (int c = threadIdx.x; c < c_in; c += blockDim.x) {
    const int g = c / gs;
    const double m = cs[g * 2] / count;
    double var = cs[g * 2 + 1] / count - m * m;
    if (var < 0.0) var = 0.0;
    const float rstd = static_cast<float>(1.0 / sqrt(var + (double)eps));
    const float w = __ldg(W + (long long)o * c_in + c) * rstd;
    wb[((long long)cloud * c_out + o) * ldwb + c] = __float2bfloat16(w);
    acc += w * static_cast<float>(m);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) bb[(long long)cloud * c_out + o] = __ldg(bias + o) - v;
  }
}

// fp32 NCHW -> bf16 NHWC (per image), via a 32x32 smem transpose tile over (C, H*W).
__global__ void nchw_to_nhwc_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int C, int HW) {
  __shared__ float tile[32][33];
  const int img = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const float* s = src + (long long)img * C * HW;
  __nv_bfloat16* d = dst + (long long)img * C * HW;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, pix = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && pix < HW) ? s[(long long)c * HW + pix] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int pix = p0 + i, c = c0 + threadIdx.x;
    if (c < C && pix < HW) d[(long long)pix * C + c] = __float2bfloat16(tile[threadIdx.x][i]);
  }
}

}  // namespace

int launch_lookup(const gecco_lookup_args& a, cudaStream_t s) {
  GECCO_REQUIRE(a.n_levels >= 1 && a.n_levels <= GECCO_MAX_LEVELS, "lookup: 1..%d pyramid levels supported", GECCO_MAX_LEVELS);
  GECCO_REQUIRE(a.reparam >= 0 && a.reparam <= 2, "lookup: unknown reparam %d", a.reparam);
  GECCO_REQUIRE(a.xin && a.K, "lookup: xin and K are required");
  GECCO_REQUIRE(a.out_bf16 || a.out_f32, "lookup: no output");
  LookupP p;
  p.xin = a.xin; p.sigma = a.sigma; p.sigma_stride = a.sigma_stride; p.sigma_data = a.sigma_data;
  p.reparam = a.reparam;
  for (int j = 0; j < 3; ++j) { p.mean[j] = a.mean[j]; p.rsig[j] = a.sigma_r[j]; }
  p.logit_scale = a.logit_scale;
  p.K = a.K;
  p.n_levels = a.n_levels;
  int ctot = 0;
  for (int l = 0; l < GECCO_MAX_LEVELS; ++l) {
    p.lvl_ptr[l] = nullptr; p.lvl_h[l] = p.lvl_w[l] = p.lvl_c[l] = 0;
  }
  for (int l = 0; l < a.n_levels; ++l) {
    GECCO_REQUIRE(a.level_ptr[l] && a.level_c[l] % 8 == 0 && a.level_h[l] > 0 && a.level_w[l] > 0,
                  "lookup: level %d must be non-empty with a channel count that is a multiple of 8", l);
    p.lvl_ptr[l] = static_cast<const __nv_bfloat16*>(a.level_ptr[l]);
    p.lvl_h[l] = a.level_h[l]; p.lvl_w[l] = a.level_w[l]; p.lvl_c[l] = a.level_c[l];
    ctot += a.level_c[l];
  }
  GECCO_REQUIRE(ctot <= 32 * 8 * LK_MAXCH, "lookup: at most %d channels supported", 32 * 8 * LK_MAXCH);
  GECCO_REQUIRE(!a.out_bf16 || a.ldo16 % 8 == 0, "lookup: bf16 leading dimension must be a multiple of 8");
  GECCO_REQUIRE(!a.out_f32 || a.ldo32 % 4 == 0, "lookup: fp32 leading dimension must be a multiple of 4");
  GECCO_REQUIRE(!a.stats || (a.stat_groups > 0 && ctot % a.stat_groups == 0), "lookup: bad statistics groups");
  p.points = a.points; p.rows_per_cloud = a.rows_per_cloud; p.ctot = ctot;
  p.out16 = static_cast<__nv_bfloat16*>(a.out_bf16); p.ldo16 = a.ldo16;
  p.out32 = a.out_f32; p.ldo32 = a.ldo32;
  p.stats = a.stats; p.stat_groups = a.stats ? a.stat_groups : 0;
  if (a.points == 0 || a.clouds == 0) return GECCO_OK;
  dim3 grid(ceil_div(a.points, LK_WARPS * LK_POINTS_PER_WARP), a.clouds);
  lookup_kernel<<<grid, LK_WARPS * 32, p.stat_groups * 2 * sizeof(float), s>>>(p);
  GECCO_CHECK_LAUNCH("lookup_kernel");
  return GECCO_OK;
}

int launch_fold_gn(const float* W, const float* bias, const double* stats, double count, float eps, int groups,
                   int c_in, int c_out, int clouds, void* wb, long long ldwb, float* bb, cudaStream_t s) {
  GECCO_REQUIRE(groups > 0 && c_in % groups == 0, "fold_gn: bad groups");
  dim3 grid(c_out, clouds);
  fold_gn_kernel<<<grid, 128, 0, s>>>(W, bias, stats, count, eps, groups, c_in, c_out,
                                      static_cast<__nv_bfloat16*>(wb), ldwb, bb);
  GECCO_CHECK_LAUNCH("fold_gn_kernel");
  return GECCO_OK;
}

}  // namespace gecco

extern "C" int gecco_lookup(const gecco_lookup_args* a, void* stream) {
  if (!a) { gecco::set_error("gecco_lookup: null args"); return GECCO_ERR_INVALID; }
  return gecco::launch_lookup(*a, static_cast<cudaStream_t>(stream));
}

extern "C" int gecco_fold_group_norm(const float* w, const float* bias, const double* stats, double count, float eps,
                                     int32_t groups, int32_t c_in, int32_t c_out, int32_t clouds, void* w_folded_bf16,
                                     int64_t ldw, float* bias_folded, void* stream) {
  return gecco::launch_fold_gn(w, bias, stats, count, eps, groups, c_in, c_out, clouds, w_folded_bf16, ldw, bias_folded,
                               static_cast<cudaStream_t>(stream));
}

extern "C" int gecco_pack_features(const float* nchw, void* nhwc_bf16, int32_t images, int32_t c, int32_t h, int32_t w,
                                   void* stream) {
  using namespace gecco;
  GECCO_REQUIRE(nchw && nhwc_bf16 && images > 0 && c > 0 && h > 0 && w > 0, "pack_features: bad arguments");
  dim3 grid(ceil_div(h * w, 32), ceil_div(c, 32), images);
  nchw_to_nhwc_bf16_kernel<<<grid, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(
      nchw, static_cast<__nv_bfloat16*>(nhwc_bf16), c, h * w);
  GECCO_CHECK_LAUNCH("nchw_to_nhwc_bf16_kernel");
  return GECCO_OK;
}
