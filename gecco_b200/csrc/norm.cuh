// Group statistics -> mean / rstd, shared by the AdaGN kernels (elementwise.cu) and the A-operand normalisation of the
// CTA-pair GEMM (gemm_pair.cu).
#pragma once
#include <cuda_runtime.h>

namespace gecco {

// mean / rstd of normalisation group `g` (of `gs` channels) from statistics kept at `sgs`-channel
// granularity (gs % sgs == 0): sums of gs/sgs consecutive fine groups.
__device__ __forceinline__ void group_mean_rstd(const double* __restrict__ cstats, int g, int gs, int sgs, double count,
                                                float eps, float& mean, float& rstd) {
  const int per = gs / sgs;
  double s1 = 0.0, s2 = 0.0;
  for (int i = 0; i < per; ++i) {
    s1 += cstats[(g * per + i) * 2];
    s2 += cstats[(g * per + i) * 2 + 1];
  }
  const double m = s1 / count;
  double var = s2 / count - m * m;
  if (var < 0.0) var = 0.0;
  mean = static_cast<float>(m);
  rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
}

}  // namespace gecco
