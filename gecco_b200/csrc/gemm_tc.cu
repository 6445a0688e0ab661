// Persistent, warp-specialised tcgen05 GEMM with a fused epilogue (gecco_gemm).
//
//   warp 0 : TMA producer  (A tile 128x64 + W tile 192x64 bf16 per stage, SWIZZLE_128B)
//   warp 1 : MMA issuer    (one thread; 4 x tcgen05.mma M=128 N=192 K=16 per stage)
//   warp 2 : TMEM allocator (512 columns = two 192-wide fp32 accumulator slots)
//   warps 4-7 : epilogue   (tcgen05.ld -> bias / xyz-embed / activation / residual /
//                           AdaGN statistics -> fp32 and/or bf16 stores)
// The two TMEM slots let the epilogue of tile i overlap the main loop of tile i+1.
#include "common.cuh"
#include "ptx.cuh"

#include <cudaTypedefs.h>
#include <math.h>

namespace gecco {

namespace {

constexpr int BM = 128;
constexpr int BN = 192;
constexpr int BK = 64;
constexpr int STAGES = 5;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KiB
constexpr int B_STAGE_BYTES = BN * BK * 2;  // 24 KiB
constexpr int ACC_COLS = 256;               // TMEM columns reserved per accumulator slot
constexpr int TMEM_COLS = 512;
constexpr int GEMM_THREADS = 256;
constexpr int CHUNK = 48;  // epilogue column chunk: 4 AdaGN groups of 12 channels
constexpr int SMEM_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 1024 /*align*/ + 256 /*barriers*/;

struct KParams {
  int M, n_out, K;
  int rows_per_cloud, valid_rows, w_rows_per_cloud;
  const float* bias;
  int bias_stride;
  int act;
  float act_k;  // -log2(e) / (2 alpha^2)
  const float* res;
  long long ldr;
  float* out_f32;
  long long ldo32;
  __nv_bfloat16* out_bf16;
  long long ldo16;
  double* stats;
  const float* geom;
  const float* sigma;
  int sigma_stride;
  float sigma_data;
  const float* wx;
  int num_m_blocks, num_n_blocks;
};

__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_w,
               const KParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES));
  uint64_t* full_bar = bars;                 // [STAGES]
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  uint64_t* acc_full = bars + 2 * STAGES;    // [2]
  uint64_t* acc_empty = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_blocks * p.num_n_blocks;
  const int num_kb = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_w);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int m0 = (t / p.num_n_blocks) * BM;
      const int n0 = (t % p.num_n_blocks) * BN;
      const int wrow = (p.w_rows_per_cloud ? (m0 / p.rows_per_cloud) * p.w_rows_per_cloud : 0) + n0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full_bar[stage], A_STAGE_BYTES + B_STAGE_BYTES);
        tma_load_2d(sA + stage * A_STAGE_BYTES, &tma_a, &full_bar[stage], kb * BK, m0);
        tma_load_2d(sB + stage * B_STAGE_BYTES, &tma_w, &full_bar[stage], kb * BK, wrow);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int slot = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&acc_empty[slot], acc_phase ^ 1);
      tc_fence_after_sync();
      const uint32_t tmem_d = tmem_base + slot * ACC_COLS;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after_sync();
        const uint64_t da = umma_desc_k_sw128(smem_u32(sA + stage * A_STAGE_BYTES));
        const uint64_t db = umma_desc_k_sw128(smem_u32(sB + stage * B_STAGE_BYTES));
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // +32 B per K=16 step inside the 128 B swizzle row (address field is in 16 B units)
          umma_bf16_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(&acc_full[slot]);
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue
    const int q = warp & 3;  // TMEM lane quadrant of this warp
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int slot = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int m0 = (t / p.num_n_blocks) * BM;
      const int n0 = (t % p.num_n_blocks) * BN;
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < p.M;
      const int cloud = (m0 + q * 32) / p.rows_per_cloud;  // warp-uniform (rows_per_cloud % 32 == 0)
      const bool row_valid = row_ok && (row - cloud * p.rows_per_cloud) < p.valid_rows;

      float g0 = 0.f, g1 = 0.f, g2 = 0.f;
      if (p.geom != nullptr && row_valid) {
        // geom is compact [clouds, valid_rows, 3]; output rows are padded to rows_per_cloud
        const float s = __ldg(p.sigma + (long long)cloud * p.sigma_stride);
        const float c_in = 1.0f / sqrtf(p.sigma_data * p.sigma_data + s * s);
        const float* gp = p.geom + ((long long)cloud * p.valid_rows + (row - cloud * p.rows_per_cloud)) * 3;
        g0 = c_in * __ldg(gp + 0);
        g1 = c_in * __ldg(gp + 1);
        g2 = c_in * __ldg(gp + 2);
      }

      mbar_wait(&acc_full[slot], acc_phase);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + slot * ACC_COLS;

#pragma unroll 1
      for (int c = 0; c < BN / CHUNK; ++c) {
        const int col0 = n0 + c * CHUNK;
        if (col0 >= p.n_out) break;  // warp-uniform
        uint32_t r[CHUNK];
        tmem_ld16(taddr + c * CHUNK, r);
        tmem_ld16(taddr + c * CHUNK + 16, r + 16);
        tmem_ld16(taddr + c * CHUNK + 32, r + 32);
        tmem_ld_wait();
        float v[CHUNK];
#pragma unroll
        for (int j = 0; j < CHUNK; ++j) v[j] = __uint_as_float(r[j]);

        if (p.bias != nullptr) {
          const float4* bp = reinterpret_cast<const float4*>(p.bias + (long long)cloud * p.bias_stride + col0);
#pragma unroll
          for (int j = 0; j < CHUNK / 4; ++j) {
            if (col0 + 4 * j < p.n_out) {
              const float4 b = __ldg(bp + j);
              v[4 * j + 0] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
            }
          }
        }
        if (p.geom != nullptr) {
          const float4* wp = reinterpret_cast<const float4*>(p.wx + (long long)col0 * 3);
#pragma unroll
          for (int j = 0; j < CHUNK / 4; ++j) {
            if (col0 + 4 * j < p.n_out) {
              const float4 w0 = __ldg(wp + 3 * j), w1 = __ldg(wp + 3 * j + 1), w2 = __ldg(wp + 3 * j + 2);
              v[4 * j + 0] += g0 * w0.x + g1 * w0.y + g2 * w0.z;
              v[4 * j + 1] += g0 * w0.w + g1 * w1.x + g2 * w1.y;
              v[4 * j + 2] += g0 * w1.z + g1 * w1.w + g2 * w2.x;
              v[4 * j + 3] += g0 * w2.y + g1 * w2.z + g2 * w2.w;
            }
          }
        }
        if (p.act) {
#pragma unroll
          for (int j = 0; j < CHUNK; ++j) v[j] = (exp2f(v[j] * v[j] * p.act_k) - 0.7f) * (1.0f / 0.28f);
        }
        if (p.res != nullptr && row_ok) {
          const float4* rp = reinterpret_cast<const float4*>(p.res + (long long)row * p.ldr + col0);
#pragma unroll
          for (int j = 0; j < CHUNK / 4; ++j) {
            if (col0 + 4 * j < p.n_out) {
              const float4 x = rp[j];
              v[4 * j + 0] += x.x; v[4 * j + 1] += x.y; v[4 * j + 2] += x.z; v[4 * j + 3] += x.w;
            }
          }
        }
        if (p.stats != nullptr) {
          float s[8];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int j = 0; j < 12; ++j) {
              const float x = row_valid ? v[12 * g + j] : 0.f;
              s1 += x;
              s2 += x * x;
            }
            s[2 * g] = s1;
            s[2 * g + 1] = s2;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) s[i] = warp_sum(s[i]);
          float mine = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) mine = (lane == i) ? s[i] : mine;
          if (lane < 8) {
            // stats[cloud][col/12][{sum, sumsq}]
            double* dst = p.stats + ((long long)cloud * (p.n_out / 12) + col0 / 12) * 2 + lane;
            atomicAdd(dst, static_cast<double>(mine));
          }
        }
        if (row_ok) {
          if (p.out_f32 != nullptr) {
            float4* op = reinterpret_cast<float4*>(p.out_f32 + (long long)row * p.ldo32 + col0);
#pragma unroll
            for (int j = 0; j < CHUNK / 4; ++j)
              if (col0 + 4 * j < p.n_out) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
          if (p.out_bf16 != nullptr) {
            uint4* op = reinterpret_cast<uint4*>(p.out_bf16 + (long long)row * p.ldo16 + col0);
#pragma unroll
            for (int j = 0; j < CHUNK / 8; ++j)
              if (col0 + 8 * j < p.n_out)
                op[j] = make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                   pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
          }
        }
      }
      tc_fence_before_sync();
      mbar_arrive(&acc_empty[slot]);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

PFN_cuTensorMapEncodeTiled g_encode = nullptr;

}  // namespace

int resolve_driver() {
  if (g_encode != nullptr) return GECCO_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return GECCO_ERR_CUDA;
  }
  g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
  return GECCO_OK;
}

// bf16 row-major [rows, cols] (ld elements) viewed as a 2D tensor {cols, rows}; box {64, box_rows}.
int make_tmap_bf16(CUtensorMap* m, const void* ptr, uint64_t cols, uint64_t rows, uint64_t ld, uint32_t box_rows) {
  if (int rc = resolve_driver()) return rc;
  GECCO_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "operand pointer must be 16-byte aligned");
  GECCO_REQUIRE(ld % 8 == 0, "operand leading dimension must be a multiple of 8 elements (got %llu)",
                (unsigned long long)ld);
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (cols=%llu rows=%llu ld=%llu)", (int)r,
              (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)ld);
    return GECCO_ERR_CUDA;
  }
  return GECCO_OK;
}

int launch_gemm(const gecco_gemm_args& a, cudaStream_t stream) {
  GECCO_REQUIRE(a.a && a.w, "gemm: null operand");
  GECCO_REQUIRE(a.m > 0 && a.n_out > 0 && a.k > 0, "gemm: empty problem m=%d n=%d k=%d", a.m, a.n_out, a.k);
  GECCO_REQUIRE(a.n_out % 8 == 0, "gemm: n_out (%d) must be a multiple of 8", a.n_out);
  GECCO_REQUIRE(a.rows_per_cloud > 0 && a.rows_per_cloud % 32 == 0, "gemm: rows_per_cloud (%d) must be a positive multiple of 32",
                a.rows_per_cloud);
  GECCO_REQUIRE(a.valid_rows > 0 && a.valid_rows <= a.rows_per_cloud, "gemm: valid_rows out of range");
  GECCO_REQUIRE(a.out_f32 || a.out_bf16, "gemm: no output buffer");
  GECCO_REQUIRE(!a.stats || a.n_out % 48 == 0, "gemm: statistics epilogue needs n_out %% 48 == 0");
  GECCO_REQUIRE(!a.geom || (a.sigma && a.wx), "gemm: xyz embed needs sigma and wx");
  GECCO_REQUIRE(!a.w_rows_per_cloud || a.rows_per_cloud % BM == 0,
                "gemm: per-cloud weights need rows_per_cloud %% 128 == 0");
  GECCO_REQUIRE(!a.res || a.ldr % 4 == 0, "gemm: residual leading dimension must be a multiple of 4");
  GECCO_REQUIRE(!a.out_f32 || a.ldo32 % 4 == 0, "gemm: fp32 output leading dimension must be a multiple of 4");
  GECCO_REQUIRE(!a.out_bf16 || a.ldo16 % 8 == 0, "gemm: bf16 output leading dimension must be a multiple of 8");

  const int clouds = ceil_div(a.m, a.rows_per_cloud);
  const uint64_t w_rows = a.w_rows_per_cloud ? (uint64_t)a.w_rows_per_cloud * clouds : (uint64_t)a.n_out;
  CUtensorMap ta, tw;
  if (int rc = make_tmap_bf16(&ta, a.a, a.k, a.m, a.lda, BM)) return rc;
  if (int rc = make_tmap_bf16(&tw, a.w, a.k, w_rows, a.ldw, BN)) return rc;

  KParams p;
  p.M = a.m; p.n_out = a.n_out; p.K = a.k;
  p.rows_per_cloud = a.rows_per_cloud; p.valid_rows = a.valid_rows; p.w_rows_per_cloud = a.w_rows_per_cloud;
  p.bias = a.bias; p.bias_stride = a.bias_stride;
  p.act = a.act;
  p.act_k = a.act ? static_cast<float>(-1.4426950408889634 / (2.0 * (double)a.act_alpha * (double)a.act_alpha)) : 0.f;
  p.res = a.res; p.ldr = a.ldr;
  p.out_f32 = a.out_f32; p.ldo32 = a.ldo32;
  p.out_bf16 = static_cast<__nv_bfloat16*>(a.out_bf16); p.ldo16 = a.ldo16;
  p.stats = a.stats;
  p.geom = a.geom; p.sigma = a.sigma; p.sigma_stride = a.sigma_stride; p.sigma_data = a.sigma_data; p.wx = a.wx;
  p.num_m_blocks = ceil_div(a.m, BM);
  p.num_n_blocks = ceil_div(a.n_out, BN);

  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(gemm_tc_kernel)");
    attr_set = true;
  }
  const int tiles = p.num_m_blocks * p.num_n_blocks;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  gemm_tc_kernel<<<grid, GEMM_THREADS, SMEM_BYTES, stream>>>(ta, tw, p);
  GECCO_CHECK_LAUNCH("gemm_tc_kernel launch");
  return GECCO_OK;
}

}  // namespace gecco

extern "C" int gecco_gemm(const gecco_gemm_args* args, void* stream) {
  if (args == nullptr) {
    gecco::set_error("gecco_gemm: null args");
    return GECCO_ERR_INVALID;
  }
  return gecco::launch_gemm(*args, static_cast<cudaStream_t>(stream));
}
