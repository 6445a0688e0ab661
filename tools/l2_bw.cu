// Measured L2 -> SM read throughput on this GPU (the roofline DESIGN.md 4.10 holds the pair kernels against): every SM
// streams an L2-resident buffer (default 32 MB, far below the 126 MB L2) with 16-byte ld.global.cg loads, many passes.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(1024) rd(const uint4* __restrict__ p, long long n16, int passes, unsigned* sink) {
  unsigned acc = 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (int it = 0; it < passes; ++it)
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
      uint4 v;
      asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p + i));
      acc += v.x ^ v.y ^ v.z ^ v.w;
    }
  if (acc == 0x12345678u) *sink = acc;
}
int main(int argc, char** argv) {
  const long long mb = argc > 1 ? atoll(argv[1]) : 32;
  const long long bytes = mb << 20, n16 = bytes / 16;
  uint4* buf; unsigned* sink;
  cudaMalloc(&buf, bytes); cudaMalloc(&sink, 4); cudaMemset(buf, 1, bytes);
  int sms = 0, clk = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0); cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int bpsm : {1, 2}) for (int threads : {512, 1024}) {
    const int passes = 200;
    rd<<<sms * bpsm, threads>>>(buf, n16, 5, sink);
    cudaEventRecord(e0);
    rd<<<sms * bpsm, threads>>>(buf, n16, passes, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double gbs = (double)bytes * passes / (ms * 1e-3) / 1e9;
    printf("buffer %lld MB, %d CTAs/SM x %d threads: %.0f GB/s L2->SM (%.1f B/clk/SM at the %d MHz max clock)\n", mb, bpsm, threads, gbs,
           gbs * 1e9 / sms / (clk * 1e3), clk / 1000);
  }
  printf("error state: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
