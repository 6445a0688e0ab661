/* Development-only exports of libgecco_b200.so (cycle counters, tcgen05.mma micro-benchmark).  NOT part of the product
 * ABI in include/gecco_b200.h; used by tools/*.py only. */
#pragma once
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
/* device buffer ([148][32] int64) receiving per-CTA cycle counters of the tcgen05 kernels (builds with
 * GECCO_DEBUG_COUNTERS=1); NULL disables. */
int gecco_set_debug_buffer(void* buf);
/* cycles for `batch` back-to-back M128 x n x K16 bf16 tcgen05.mma on one SM (tools/umma_bench.py).
 * mode 0: A, B from shared memory (K-major); 1: B MN-major; 2: A from TMEM, B K-major; 3: A from TMEM, B MN-major.
 * out: device buffer of two int64 (best, mean over reps - 1 repetitions). */
int gecco_debug_umma_bench(int32_t mode, int32_t n, int32_t batch, int32_t reps, long long* out, void* stream);
#ifdef __cplusplus
}
#endif
