// Development aid (tools/umma_bench.py): cycles per tcgen05.mma for the small shapes of the attention kernels, measured on
// one SM.  One thread issues a batch of M128 x N x K16 bf16 MMAs back to back (accumulating into one TMEM tile), commits,
// and waits; the operands are whatever shared memory / TMEM holds (timing only).  Variants:
//   mode 0: A and B from shared memory, both K-major (SWIZZLE_128B)           -- the GEMMs, S = Q K^T
//   mode 1: A from shared memory, B MN-major (second 64-element atom via LBO)  -- O = P V with V as TMA delivers it
//   mode 2: A from TMEM, B K-major                                            -- FA4-style P in TMEM
//   mode 3: A from TMEM, B MN-major
// Not part of the product path.
#include "common.cuh"
#include "debug_api.h"
#include "ptx.cuh"

namespace gecco {
namespace {

__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

constexpr int UB_SMEM = 1024 + 16384 /*A*/ + 65536 /*B*/ + 64;

__global__ void __launch_bounds__(128, 1) umma_bench_kernel(int mode, int n, int batch, int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;            // [128 rows x 128 B]
  uint8_t* sB = sA + 16384;      // K-major: [256 rows x 128 B]; MN-major: atoms of [64 k x 128 B], 8 KB apart
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 65536);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < (16384 + 65536) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  if (warp == 0) {
    const uint32_t sA_u = uniform_u32(smem_u32(sA)), sB_u = uniform_u32(smem_u32(sB));
    const bool mn = (mode & 1) != 0, ts = mode >= 2;
    const uint32_t idesc = umma_idesc_bf16(128, n) | (mn ? (1u << 16) : 0u);
    long long best = 1LL << 60, total = 0;
    for (int r = 0; r < reps; ++r) {
      long long t0 = 0;
      if (elect_one()) {
        t0 = clock64();
        for (int i = 0; i < batch; ++i) {
          const int kk = i & 3;
          const uint64_t db = mn ? desc_sw128(sB_u + kk * 2048, 8192) : desc_sw128(sB_u, 16) + 2 * kk;
          if (ts) umma_bf16_ts(tmem, tmem + 256 + kk * 8, db, idesc, i ? 1u : 0u);
          else umma_bf16_ss(tmem, desc_sw128(sA_u, 16) + 2 * kk, db, idesc, i ? 1u : 0u);
        }
        umma_commit(bar);
      }
      __syncwarp();
      mbar_wait(bar, r & 1u);
      // the elected lane holds t0; take the time on every lane and reduce with the elected lane's start
      const long long t1 = clock64();
      long long t0_all = t0;
      for (int o = 16; o > 0; o >>= 1) {
        const long long other = __shfl_xor_sync(0xffffffffu, t0_all, o);
        t0_all = t0_all > other ? t0_all : other;  // non-elected lanes hold 0
      }
      const long long dt = t1 - t0_all;
      if (r > 0) {  // the first repetition warms the descriptors / instruction cache
        best = dt < best ? dt : best;
        total += dt;
      }
    }
    if (lane == 0) {
      out[0] = best;
      out[1] = reps > 1 ? total / (reps - 1) : total;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem);
  }
}

}  // namespace
}  // namespace gecco

// out[0] = best, out[1] = mean cycles for `batch` MMAs + commit + wake-up (device buffer of 2 int64).
extern "C" int gecco_debug_umma_bench(int32_t mode, int32_t n, int32_t batch, int32_t reps, long long* out, void* stream) {
  using namespace gecco;
  GECCO_REQUIRE(mode >= 0 && mode <= 3 && n >= 16 && n <= 256 && n % 16 == 0 && batch >= 1 && reps >= 2 && out != nullptr,
                "umma_bench: bad arguments");
  GECCO_REQUIRE(!(mode & 1) || n <= 128, "umma_bench: MN-major B is laid out for N <= 128");
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(umma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, UB_SMEM);
    if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(umma_bench_kernel)");
    attr_set = true;
  }
  umma_bench_kernel<<<1, 128, UB_SMEM, static_cast<cudaStream_t>(stream)>>>(mode, n, batch, reps, out);
  GECCO_CHECK_LAUNCH("umma_bench_kernel");
  return GECCO_OK;
}
