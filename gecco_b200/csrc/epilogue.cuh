// Row-tile epilogue shared by the tcgen05 kernels: drains a 128-row x 192-column fp32 accumulator panel from
// TMEM in 32-column chunks and applies, in this order,
//   + bias[cloud][n]  + xyz embed  -> Gaussian activation -> + residual -> AdaGN statistics -> fp32 / bf16 stores.
// All global traffic of the epilogue goes through TMA: the residual chunk is prefetched into swizzled shared
// memory by a loader thread (res_full / res_empty mbarriers), results are staged in swizzled shared memory and
// written with bulk tensor stores, so every global access is a full 128 B (fp32) / 64 B (bf16) row segment
// regardless of the one-thread-per-row TMEM layout.  Two epilogue groups of 128 threads (4 warps each, warp q <->
// TMEM lanes 32q..) work on alternate chunks with their own staging buffers, barriers and bulk-store thread.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace gecco {

constexpr int EPI_THREADS = 128;                      // per group
constexpr int EPI_GROUPS = 2;
constexpr int EPI_CHUNK = 32;                         // columns per chunk
constexpr int EPI_PANEL = 192;                        // columns per panel (16 AdaGN groups of 12)
constexpr int EPI_RES_BYTES = 128 * EPI_CHUNK * 4;    // one fp32 chunk (TMA box {32, 128}, SWIZZLE_128B)
constexpr int EPI_O16_BYTES = 128 * EPI_CHUNK * 2;    // one bf16 chunk (TMA box {32, 128}, SWIZZLE_64B)
constexpr int EPI_SMEM_BYTES = 4 * EPI_RES_BYTES + 2 * EPI_O16_BYTES;  // res[2] | o32[2] | o16[2]
constexpr int EPI_BAR_STAGE = 1, EPI_BAR_FREE = 2;    // named barrier ids (+ 2 * group)

struct EpiParams {
  int M, n_out;
  int rows_per_cloud, valid_rows;
  const float* bias;  // [clouds][bias_stride] or nullptr
  int bias_stride;
  int act;
  float act_k;  // -log2(e) / (2 alpha^2)
  int has_res, has_o32, has_o16;
  double* stats;  // [clouds][n_out / 12][2] or nullptr
  const float* geom;  // [clouds, valid_rows, 3] or nullptr
  const float* sigma;
  int sigma_stride;
  float sigma_data;
  const float* wx;  // [n_out, 3]
};

struct EpiSmem {
  uint8_t* res;  // [group] x EPI_RES_BYTES, 1024 B aligned
  uint8_t* o32;  // [group] x EPI_RES_BYTES
  uint8_t* o16;  // [group] x EPI_O16_BYTES
  uint64_t* res_full;   // [group], count 1 + tx
  uint64_t* res_empty;  // [group], count EPI_THREADS
};

// Loader side (one thread): prefetches the residual chunks of one panel in the order the epilogue consumes them.
__device__ __forceinline__ void epi_load_residual_panel(const EpiParams& p, const EpiSmem& sm, const CUtensorMap* tma_res,
                                                        int m0, int n0, uint32_t (&cnt)[EPI_GROUPS]) {
#pragma unroll 1
  for (int c = 0; c < EPI_PANEL / EPI_CHUNK; ++c) {
    const int col0 = n0 + c * EPI_CHUNK;
    if (col0 >= p.n_out) break;
    const uint32_t buf = c & 1u, phase = cnt[buf] & 1u;  // chunk c belongs to group c & 1
    mbar_wait(&sm.res_empty[buf], phase ^ 1u);
    mbar_arrive_expect_tx(&sm.res_full[buf], EPI_RES_BYTES);
    tma_load_2d(sm.res + buf * EPI_RES_BYTES, tma_res, &sm.res_full[buf], col0, m0);
    ++cnt[buf];
  }
}

// Sum over the 32 lanes of v[i] lands in lane i (31 shuffles instead of 32 x 5).
__device__ __forceinline__ float warp_reduce_scatter32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = upper ? v[i] : v[i + off];
      const float keep = upper ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// Epilogue side (all threads of both groups).  taddr: TMEM address of (lane quadrant q, panel column 0).
// grp: epilogue group of this thread, tid: 0..127 within the group; `cnt` counts the chunks this group has processed
// (the loader keeps the same count per group).
__device__ __forceinline__ void epi_panel(const EpiParams& p, const EpiSmem& sm, const CUtensorMap* tma_o32,
                                          const CUtensorMap* tma_o16, uint32_t taddr, int m0, int n0, int grp, int tid,
                                          uint32_t& cnt) {
  const int q = tid >> 5, lane = tid & 31;
  const int r = q * 32 + lane;  // row inside the tile == TMEM lane
  const int row = m0 + r;
  const int cloud = (m0 + q * 32) / p.rows_per_cloud;  // warp-uniform (rows_per_cloud % 32 == 0)
  const bool row_valid = row < p.M && (row - cloud * p.rows_per_cloud) < p.valid_rows;

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
  float st[32];  // [group 0..15][{sum, sumsq}] of this row over the panel
#pragma unroll
  for (int i = 0; i < 32; ++i) st[i] = 0.f;

#pragma unroll
  for (int c = 0; c < EPI_PANEL / EPI_CHUNK; ++c) {
    const int col0 = n0 + c * EPI_CHUNK;
    if ((c & 1) == grp && col0 < p.n_out) {  // uniform over the epilogue group
      uint32_t rr[EPI_CHUNK];
      tmem_ld16(taddr + c * EPI_CHUNK, rr);
      tmem_ld16(taddr + c * EPI_CHUNK + 16, rr + 16);
      tmem_ld_wait();
      float v[EPI_CHUNK];
#pragma unroll
      for (int j = 0; j < EPI_CHUNK; ++j) v[j] = __uint_as_float(rr[j]);

      if (p.bias != nullptr) {
        const float4* bp = reinterpret_cast<const float4*>(p.bias + (long long)cloud * p.bias_stride + col0);
#pragma unroll
        for (int j = 0; j < EPI_CHUNK / 4; ++j) {
          if (col0 + 4 * j < p.n_out) {
            const float4 b = __ldg(bp + j);
            v[4 * j + 0] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
          }
        }
      }
      if (p.geom != nullptr) {
        const float4* wp = reinterpret_cast<const float4*>(p.wx + (long long)col0 * 3);
#pragma unroll
        for (int j = 0; j < EPI_CHUNK / 4; ++j) {
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
        for (int j = 0; j < EPI_CHUNK; ++j) v[j] = (exp2f(v[j] * v[j] * p.act_k) - 0.7f) * (1.0f / 0.28f);
      }
      const uint32_t buf = grp, phase = cnt & 1u;
      if (p.has_res) {
        mbar_wait(&sm.res_full[buf], phase);
        const uint8_t* rb = sm.res + buf * EPI_RES_BYTES;
#pragma unroll
        for (int j = 0; j < EPI_CHUNK / 4; ++j) {
          const float4 x = *reinterpret_cast<const float4*>(rb + swz128(r * 128 + j * 16));
          v[4 * j + 0] += x.x; v[4 * j + 1] += x.y; v[4 * j + 2] += x.z; v[4 * j + 3] += x.w;
        }
        mbar_arrive(&sm.res_empty[buf]);
      }
      if (!row_valid) {
#pragma unroll
        for (int j = 0; j < EPI_CHUNK; ++j) v[j] = 0.f;  // padding rows stay exactly zero
      }
      if (p.stats != nullptr) {
#pragma unroll
        for (int j = 0; j < EPI_CHUNK; ++j) {
          const int g = (c * EPI_CHUNK + j) / 12;  // static: n0 is a multiple of 192
          st[2 * g] += v[j];
          st[2 * g + 1] += v[j] * v[j];
        }
      }
      // the staging buffers of this group were last read by the bulk store of its previous chunk
      if (tid == 0) tma_store_wait_read<0>();
      named_bar_sync(EPI_BAR_FREE + 2 * grp, EPI_THREADS);
      if (p.has_o32) {
        uint8_t* ob = sm.o32 + buf * EPI_RES_BYTES;
#pragma unroll
        for (int j = 0; j < EPI_CHUNK / 4; ++j)
          *reinterpret_cast<float4*>(ob + swz128(r * 128 + j * 16)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
      if (p.has_o16) {
        uint8_t* ob = sm.o16 + buf * EPI_O16_BYTES;
#pragma unroll
        for (int j = 0; j < EPI_CHUNK / 8; ++j)
          *reinterpret_cast<uint4*>(ob + swz64(r * 64 + j * 16)) =
              make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                         pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
      }
      fence_proxy_async_smem();
      named_bar_sync(EPI_BAR_STAGE + 2 * grp, EPI_THREADS);
      if (tid == 0) {
        if (p.has_o32) tma_store_2d(tma_o32, sm.o32 + buf * EPI_RES_BYTES, col0, m0);
        if (p.has_o16) tma_store_2d(tma_o16, sm.o16 + buf * EPI_O16_BYTES, col0, m0);
        tma_store_commit();
      }
      ++cnt;
    }
  }
  if (p.stats != nullptr) {
    const float mine = warp_reduce_scatter32(st, lane);
    const int gidx = (n0 / 12) * 2 + lane;  // [group][{sum, sumsq}]
    if (gidx < (p.n_out / 12) * 2 && m0 + q * 32 < p.M)
      atomicAdd(p.stats + (long long)cloud * (p.n_out / 12) * 2 + gidx, static_cast<double>(mine));
  }
}

}  // namespace gecco
