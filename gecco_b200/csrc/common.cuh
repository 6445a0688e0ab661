// Shared host-side helpers for the C-ABI library: error reporting and launch checks.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include "../../include/gecco_b200.h"

namespace gecco {

// Thread-local text for gecco_last_error(); every failing export sets it.
void set_error(const char* fmt, ...);

inline int fail_cuda(cudaError_t e, const char* what) {
  set_error("%s: %s", what, cudaGetErrorString(e));
  return GECCO_ERR_CUDA;
}

// Counts kernel launches issued through the library on this thread (gecco_launch_count).
extern thread_local long long g_launches;

#define GECCO_CHECK_LAUNCH(what)                              \
  do {                                                        \
    cudaError_t e__ = cudaGetLastError();                     \
    if (e__ != cudaSuccess) return gecco::fail_cuda(e__, what); \
    ++gecco::g_launches;                                      \
  } while (0)

#define GECCO_REQUIRE(cond, ...)        \
  do {                                  \
    if (!(cond)) {                      \
      gecco::set_error(__VA_ARGS__);    \
      return GECCO_ERR_INVALID;         \
    }                                   \
  } while (0)

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

int sm_count();

}  // namespace gecco
