// Internal launcher declarations shared between the per-kernel translation units and engine.cu.
#pragma once
#include "common.cuh"

namespace gecco {

int launch_gemm(const gecco_gemm_args& a, cudaStream_t s);
int launch_group_stats(const float* x, long long ldx, int clouds, int rows_per_cloud, int valid_rows, int C, int gs,
                       double* stats, cudaStream_t s);
int launch_adagn(const gecco_adagn_args& a, cudaStream_t s);
int launch_fold_adagn(const gecco_fold_adagn_args& a, cudaStream_t s);
int launch_lift(const gecco_lift_args& a, cudaStream_t s);
int launch_head(const gecco_head_args& a, cudaStream_t s);
int launch_sampler_init(const float* latents, const float* noise, double t0, double churn, long long n, double* x_hat,
                        float* xin, cudaStream_t s);
int launch_lookup(const gecco_lookup_args& a, cudaStream_t s);
int launch_fold_gn(const float* W, const float* bias, const double* stats, double count, float eps, int groups,
                   int c_in, int c_out, int clouds, void* wb, long long ldwb, float* bb, cudaStream_t s);
int launch_pool_attention(const gecco_pool_args& a, cudaStream_t s);
int launch_unpool_attention(const gecco_unpool_args& a, cudaStream_t s);

}  // namespace gecco
