// Issue-rate micro-benchmark of candidate dropout-mask sequences (per element: hash -> keep decision -> applied to a float).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/maskbench tests/bringup/maskbench.cu && gpurun_out/maskbench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256, 1) k(int iters, unsigned long long* cyc, float* sink, uint32_t key0, uint32_t thr) {
  float e[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) e[i] = 1.0f + 0.001f * (threadIdx.x + i);
  uint32_t key = key0 ^ threadIdx.x * 0x9E3779B1u;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    key = key * 0x85EBCA77u + 1u;
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      if (MODE == 0) {          // current: xor const, IMAD, ISETP, SEL
        const uint32_t h = (key ^ (uint32_t)(c * 0x9E3779B1u)) * 0x2545F491u;
        e[c] = h >= thr ? e[c] : 0.f;
      } else if (MODE == 1) {   // IMAD.HI with addend -> sign bit -> LOP3 xor
        uint32_t d;
        const uint32_t m = (0x2545F491u ^ (uint32_t)(c * 0x9E3779B1u)) >> 1 | 1u;   // < 2^31
        asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(key), "r"(m), "r"(0u - thr));
        e[c] = __uint_as_float(__float_as_uint(e[c]) ^ (d & 0x80000000u));
      } else if (MODE == 2) {   // per-column multiplier: IMAD, ISETP, SEL
        const uint32_t h = key * ((uint32_t)(c * 0x9E3779B1u + 0x2545F491u) | 1u);
        e[c] = h >= thr ? e[c] : 0.f;
      } else if (MODE == 3) {   // per-column multiplier lo product, sign via (h>>1)-thr : IMAD, SHF, IADD, LOP3  (reference point)
        const uint32_t h = key * ((uint32_t)(c * 0x9E3779B1u + 0x2545F491u) | 1u);
        const uint32_t d = (h >> 1) - thr;
        e[c] = __uint_as_float(__float_as_uint(e[c]) ^ (d & 0x80000000u));
      } else if (MODE == 4) {   // IMAD lo only (rate of IMAD)
        const uint32_t h = key * ((uint32_t)(c * 0x9E3779B1u + 0x2545F491u) | 1u);
        e[c] = __uint_as_float(__float_as_uint(e[c]) + h);
      } else if (MODE == 5) {   // IMAD.HI only + IADD
        uint32_t d;
        const uint32_t m = (0x2545F491u ^ (uint32_t)(c * 0x9E3779B1u)) >> 1 | 1u;
        asm("mul.hi.u32 %0, %1, %2;" : "=r"(d) : "r"(key), "r"(m));
        e[c] = __uint_as_float(__float_as_uint(e[c]) + d);
      } else if (MODE == 6) {   // per-column multiplier, compare, predicated negate (FMA pipe?)
        const uint32_t h = key * ((uint32_t)(c * 0x9E3779B1u + 0x2545F491u) | 1u);
        asm("{.reg .pred p; setp.lo.u32 p, %1, %2; @p neg.f32 %0, %0;}" : "+f"(e[c]) : "r"(h), "r"(thr));
      } else if (MODE == 7) {   // IMAD.HI+addend, sign-xor, plus an ex2 per element (interaction with MUFU)
        uint32_t d;
        const uint32_t m = (0x2545F491u ^ (uint32_t)(c * 0x9E3779B1u)) >> 1 | 1u;
        asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(key), "r"(m), "r"(0u - thr));
        float y;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(e[c]));
        e[c] = __uint_as_float(__float_as_uint(y) ^ (d & 0x80000000u));
      } else if (MODE == 8) {   // current + ex2
        const uint32_t h = (key ^ (uint32_t)(c * 0x9E3779B1u)) * 0x2545F491u;
        float y;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(e[c]));
        e[c] = h >= thr ? y : 0.f;
      } else if (MODE == 9) {   // ex2 only
        float y;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(e[c]));
        e[c] = y;
      }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += e[i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name) {
  unsigned long long* cyc; float* sink;
  cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 148 * 256 * 4);
  const int iters = 4000;
  k<MODE><<<148, 256>>>(10, cyc, sink, 12345u, 429496729u);
  k<MODE><<<148, 256>>>(iters, cyc, sink, 12345u, 429496729u);
  cudaDeviceSynchronize();
  unsigned long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  // per SMSP: 2 warps; elements per warp = iters*32
  printf("%-70s %.2f clk per element per warp-pair (2 warps/SMSP) -> %.2f clk/elem/SMSP-warp\n", name, c / (iters * 32.0), c / (iters * 32.0) / 2);
  cudaFree(cyc); cudaFree(sink);
}
int main() {
  run<0>("0 current: xor, IMAD, ISETP, SEL");
  run<1>("1 IMAD.HI+addend, LOP3 sign-xor");
  run<2>("2 per-column multiplier: IMAD, ISETP, SEL");
  run<3>("3 IMAD, SHF, IADD, LOP3");
  run<4>("4 IMAD + IADD");
  run<5>("5 IMAD.HI + IADD");
  run<6>("6 IMAD, ISETP, predicated neg");
  run<7>("7 ex2 + IMAD.HI+addend + LOP3");
  run<8>("8 ex2 + current");
  run<9>("9 ex2 only");
  return 0;
}
