// tcgen05.mma throughput by shape and operand major-ness, as the attention kernels issue them (run on the B200 box):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I hybrid_vit_cascade_b200/csrc -o gpurun_out/umma_bench tests/bringup/umma_bench.cu
// One warp per CTA issues `reps` groups of 8 K=16 steps, commits, waits; reports clocks per MMA instruction at issue and
// at completion, and the implied dense-bf16 rate per SM.
#include <cstdio>
#include <cstdlib>
#include "hvc_common.cuh"
using namespace hvc;

struct Cfg {
  int N;          // MMA N (M is 128)
  int a_src;      // 0 smem K-major, 1 smem MN-major, 2 TMEM
  int b_major;    // 0 K-major, 1 MN-major
  int rowb;       // smem row bytes of the operand tiles: 128 (SW128) or 64 (SW64)
};

__global__ void __launch_bounds__(128, 1) k_umma(Cfg c, int reps, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&slot, 512);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = slot;
  if (warp == 0) {
    const bool leader = elect_one();
    const uint32_t sA = smem_u32(smem), sB = smem_u32(smem + 48 * 1024);
    const uint32_t idesc = make_idesc_bf16(128, c.N, c.a_src == 1 ? kMajorMN : kMajorK, c.b_major ? kMajorMN : kMajorK);
    long long t0 = 0, t1 = 0, t2 = 0;
    t0 = clock64();
    if (leader) {
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int k16 = 0; k16 < 8; ++k16) {
          uint64_t bd, ad = 0;
          if (c.rowb == 128) {
            bd = c.b_major ? Swz<128>::desc(sB + k16 * Swz<128>::kMnStep, 8192) : Swz<128>::desc(sB + (k16 & 3) * 32 + (k16 >> 2) * 16384);
            ad = c.a_src == 1 ? make_sdesc_sw128(sA + k16 * 2048, 16384, 1024) : Swz<128>::desc(sA + (k16 & 3) * 32 + (k16 >> 2) * 16384);
          } else {
            bd = c.b_major ? Swz<64>::desc(sB + k16 * Swz<64>::kMnStep, 8192) : Swz<64>::desc(sB + (k16 & 1) * 32 + (k16 >> 1) * 8192);
            ad = c.a_src == 1 ? make_sdesc<4>(sA + k16 * 1024, 8192, 512) : Swz<64>::desc(sA + (k16 & 1) * 32 + (k16 >> 1) * 8192);
          }
          if (c.a_src == 2) umma_ts(tb, tb + 256 + k16 * 8, bd, idesc, 1u);
          else umma_ss(tb, ad, bd, idesc, 1u);
        }
      }
      t1 = clock64();
      tc_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0, 1);
    t2 = clock64();
    if (leader) { out[blockIdx.x * 2] = t1 - t0; }
    if (threadIdx.x == 0) out[blockIdx.x * 2 + 1] = t2 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

int main() {
  unsigned long long* d;
  const int grid = 148, reps = 64;
  cudaMalloc(&d, grid * 2 * sizeof(unsigned long long));
  cudaFuncSetAttribute(k_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  struct { const char* name; Cfg c; } tests[] = {
      {"SS N=128 A:K   B:K   sw128 (S^T, dP^T d64)", {128, 0, 0, 128}},
      {"SS N=256 A:K   B:K   sw128", {256, 0, 0, 128}},
      {"SS N=64  A:K   B:K   sw128", {64, 0, 0, 128}},
      {"SS N=64  A:K   B:MN  sw128 (dK d64)", {64, 0, 1, 128}},
      {"SS N=64  A:MN  B:MN  sw128 (dQ d64)", {64, 1, 1, 128}},
      {"TS N=64  A:TMEM B:MN sw128 (dV, PV d64)", {64, 2, 1, 128}},
      {"SS N=128 A:MN  B:K   sw128", {128, 1, 0, 128}},
      {"SS N=128 A:K   B:K   sw64  (S^T d32: only 2 k-steps real)", {128, 0, 0, 64}},
      {"SS N=32  A:K   B:MN  sw64  (dK d32)", {32, 0, 1, 64}},
      {"SS N=32  A:MN  B:MN  sw64  (dQ d32)", {32, 1, 1, 64}},
      {"TS N=32  A:TMEM B:MN sw64  (dV, PV d32)", {32, 2, 1, 64}},
  };
  for (auto& t : tests) {
    for (int rep = 0; rep < 2; ++rep) {
      k_umma<<<grid, 128, 100 * 1024>>>(t.c, reps, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", t.name, cudaGetErrorString(e)); return 1; }
    }
    unsigned long long h[2 * 148];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double iss = 0, tot = 0;
    for (int i = 0; i < grid; ++i) { iss += h[2 * i]; tot += h[2 * i + 1]; }
    const int n = reps * 8;
    iss /= grid * (double)n; tot /= grid * (double)n;
    printf("%-58s issue %6.1f clk/mma  complete %6.1f clk/mma  -> %7.0f flop/clk/SM (%.0f%% of 8192)\n", t.name, iss, tot,
           2.0 * 128 * t.c.N * 16 / tot, 100.0 * 2.0 * 128 * t.c.N * 16 / tot / 8192);
  }
  return 0;
}
