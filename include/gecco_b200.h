/*
 * gecco_b200 — C ABI of the B200 (sm_100a) implementation of gecco-torch's
 * reverse-diffusion sampling path.
 *
 * The reference (cvlab-epfl/gecco, gecco-torch) has no FFI of its own: the path
 * sits behind plain Python nn.Module calls.  This header is the
 * boundary a maintainer binds instead of the ATen call sites listed in
 * SURVEY.md §2.2; each entry point cites the reference code it replaces
 * (paths relative to gecco-torch/src/gecco_torch/).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless named host_*; the caller (torch)
 *     owns all memory, the library never allocates or frees user-visible buffers;
 *   - every launch function takes the CUDA stream as a void* (cudaStream_t), is
 *     asynchronous, does no host synchronisation and is CUDA-graph capturable;
 *   - every export returns 0 on success or a negative GECCO_ERR_* code;
 *     gecco_last_error() returns the text of the last failure on this thread;
 *   - there is no CPU fallback: a device that is not sm_100 is an error.
 */
#ifndef GECCO_B200_H
#define GECCO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GECCO_OK 0
#define GECCO_ERR_INVALID (-1) /* bad argument / unsupported shape */
#define GECCO_ERR_CUDA (-2)    /* CUDA runtime or driver error      */
#define GECCO_ERR_DEVICE (-3)  /* not an sm_100 device              */

#define GECCO_ABI_VERSION 1

int gecco_abi_version(void);
const char* gecco_last_error(void);
/* Checks that `device` is sm_100 and resolves the driver entry points. */
int gecco_init(int device);

/* ------------------------------------------------------------------------
 * Dense projection on the tcgen05 tensor cores.
 *   out[m, n] = epilogue( sum_k a[m, k] * w[wrow(m) + n, k] )
 * a: bf16 [M, K] row-major (lda elements), w: bf16 [*, K] row-major (ldw).
 * Replaces every nn.Linear / F.linear on the path: kv_proj
 * (models/set_transformer.py:49), the in/out projections of nn.MultiheadAttention
 * (:90,112), MLP (models/mlp.py:5-39) and img_feature_proj (models/ray.py:52-55).
 * Epilogue, applied in this order on the fp32 accumulator:
 *   + bias[cloud*bias_stride + n]                     (bias may be NULL)
 *   + xyz embed: sum_j c_in(sigma)*geom[m,j]*wx[n,j]  (geom may be NULL; models/ray.py:99,113)
 *   Gaussian activation (exp(-z^2/(2 alpha^2)) - 0.7)/0.28 when act != 0 (models/activation.py:17-24)
 *   + res[m, n]                                       (res may be NULL; residual adds set_transformer.py:164,166)
 *   per (cloud, 12-channel group) sum / sum of squares of the result added into
 *   stats (double [clouds][n_out/12][2]) for the next AdaGN (models/normalization.py:36-44)
 * and written as fp32 (out_f32) and/or bf16 (out_bf16).
 * Rows are grouped in clouds of rows_per_cloud rows of which the first valid_rows
 * are real points (the rest is padding, excluded from stats).
 * ------------------------------------------------------------------------ */
typedef struct gecco_gemm_args {
  const void* a;   int64_t lda;
  const void* w;   int64_t ldw;
  int32_t m, n_out, k;
  int32_t rows_per_cloud;   /* > 0, multiple of 32 */
  int32_t valid_rows;       /* <= rows_per_cloud */
  int32_t w_rows_per_cloud; /* 0: one weight for all clouds; else row offset between per-cloud weights */
  const float* bias; int32_t bias_stride;
  int32_t act; float act_alpha;
  const float* res; int64_t ldr;
  float* out_f32;  int64_t ldo32;
  void* out_bf16;  int64_t ldo16;
  double* stats;
  const float* geom;        /* [M, 3] fp32 raw sampler state x (not yet scaled by c_in) */
  const float* sigma; int32_t sigma_stride; /* sigma[cloud*sigma_stride] */
  const float* wx;          /* [n_out, 3] fp32 */
} gecco_gemm_args;

int gecco_gemm(const gecco_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GECCO_B200_H */
