// Micro-benchmarks of the per-SM rates the attention kernels are designed around (run on the B200 box):
//   tcgen05.ld / tcgen05.st bandwidth per SM, MUFU ex2 rate per SM, f32x2 FMA rate.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I hybrid_vit_cascade_b200/csrc -o gpurun_out/microbench tests/bringup/microbench.cu
#include <cstdio>
#include <cstdlib>
#include "hvc_common.cuh"
using namespace hvc;

__global__ void __launch_bounds__(512, 1) k_tmem(int mode, int iters, unsigned long long* cyc, unsigned* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  uint32_t v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = threadIdx.x + i;
  __syncthreads();
  const long long t0 = clock64();
  if (mode == 0) {            // ld x32, wait every 4 loads
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(base + ((it * 4 + c) & 7) * 32 + (warp >> 2) * 0, r);
        tmem_ld_wait();
        acc ^= r[0] ^ r[31];
      }
    }
  } else if (mode == 1) {     // 4 loads in flight then one wait
    for (int it = 0; it < iters; ++it) {
      uint32_t r0[32], r1[32], r2[32], r3[32];
      tmem_ld_32x32(base + 0, r0); tmem_ld_32x32(base + 32, r1); tmem_ld_32x32(base + 64, r2); tmem_ld_32x32(base + 96, r3);
      tmem_ld_wait();
      acc ^= r0[0] ^ r1[5] ^ r2[9] ^ r3[31];
    }
  } else {                    // st x32
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_st_32x32(base + c * 32, v);
      tmem_st_wait();
      v[0] += it;
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc + v[0];
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

__global__ void __launch_bounds__(1024, 1) k_mufu(int mode, int iters, unsigned long long* cyc, float* sink) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = -0.001f * (threadIdx.x + i);
  float2 y[4] = {{1.f, 1.f}, {1.f, 1.f}, {1.f, 1.f}, {1.f, 1.f}};
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (mode == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = ex2_approx(x[i]) - 1.0001f;
    } else if (mode == 1) {
#pragma unroll
      for (int i = 0; i < 4; ++i) y[i] = ffma2(y[i], make_float2(0.999f, 1.001f), make_float2(x[i], x[i + 4]));
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], 0.999f, 0.001f);
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s + y[0].x + y[1].y + y[2].x + y[3].y;
}

int main() {
  unsigned long long* cyc; unsigned* sink; float* fsink;
  cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 148 * 1024 * 4); cudaMalloc(&fsink, 148 * 1024 * 4);
  unsigned long long h[148];
  const int iters = 2000;
  for (int mode = 0; mode < 3; ++mode)
    for (int threads : {128, 256, 512}) {
      k_tmem<<<148, threads>>>(mode, iters, cyc, sink);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("tmem mode %d failed: %s\n", mode, cudaGetErrorString(cudaGetLastError())); return 1; }
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double bytes = (double)iters * 4 * 32 * 4 * threads;   // per SM
      printf("tmem %s threads=%d: %.1f B/clk/SM (%llu clk)\n", mode == 0 ? "ld.x32 wait-each" : mode == 1 ? "ld.x32 4-in-flight" : "st.x32", threads, bytes / h[0], h[0]);
    }
  for (int mode = 0; mode < 3; ++mode)
    for (int threads : {128, 256, 512, 1024}) {
      k_mufu<<<148, threads>>>(mode, iters, cyc, fsink);
      cudaDeviceSynchronize();
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double ops = (double)iters * 8 * threads;
      printf("%s threads=%d: %.2f lane-ops/clk/SM\n", mode == 0 ? "ex2+fadd" : mode == 1 ? "ffma2 (2 fma per op)" : "ffma", threads, ops / h[0]);
    }
  return 0;
}
