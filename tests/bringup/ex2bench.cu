// MUFU ex2 rate by operand type: f32, f16x2, bf16x2 (two exps per instruction?) -- 256 threads per SM as in the attention kernels' turns.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o variants_tmp/ex2bench tests/bringup/ex2bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256, 1) k(int iters, unsigned long long* cyc, uint32_t* sink) {
  uint32_t x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = 0xBC00BC00u + threadIdx.x + i;     // small negative halves / floats
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(x[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(x[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(x[i]));
      if (MODE == 3) { asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(x[i])); asm volatile("fma.rn.bf16x2 %0, %0, %0, %0;" : "+r"(x[i])); }
      if (MODE == 4) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(x[i])); asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+r"(x[i])); }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s ^= x[i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name) {
  unsigned long long* cyc; uint32_t* sink;
  cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 148 * 256 * 4);
  const int iters = 4000;
  k<MODE><<<148, 256>>>(10, cyc, sink);
  k<MODE><<<148, 256>>>(iters, cyc, sink);
  cudaDeviceSynchronize();
  unsigned long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  printf("%-44s %.2f instr-lanes/clk/SM\n", name, 256.0 * iters * 16 / c);
}
int main() { run<0>("ex2.approx.ftz.f32"); run<1>("ex2.approx.f16x2"); run<2>("ex2.approx.ftz.bf16x2"); run<3>("ex2 bf16x2 + fma bf16x2"); run<4>("ex2 f32 + fma f32"); return 0; }
