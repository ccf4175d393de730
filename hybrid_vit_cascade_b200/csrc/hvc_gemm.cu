// hvc_gemm.cu -- persistent warp-specialised bf16 GEMM for sm_100a.
//
//   D[M,N] = alpha * sum_k A(m,k) B(n,k)         fp32 accumulation in TMEM
//
// One CTA per SM, 192 threads:
//   warp 0      TMA producer   : cp.async.bulk.tensor -> 128B-swizzled smem ring (5 stages x 32 KB)
//   warp 1      MMA issuer     : one elected lane issues tcgen05.mma (128x128x16), commits to mbarriers;
//                                owns the TMEM allocation (2 accumulator stages x 128 columns)
//   warps 2..5  epilogue       : tcgen05.ld accumulator -> registers -> fused epilogue -> global
// The epilogue of tile i overlaps the main loop of tile i+1 through the two TMEM stages.
//
// Operands may be K-major (stored [rows, K]) or MN-major (stored [K, rows]); the latter is what the
// backward GEMMs need (dgrad reads W as stored, wgrad reads dy and x as stored) so no transposed copies
// of activations or weights are ever materialised.
//
// Roofline: tensor pipe for K >= ~512; at the backbone's C=256 projections the kernel sits on the
// HBM/tensor ridge (A read + D write vs 2*M*N*K flops), so the fused epilogues are what matters.
#include "hvc_common.cuh"
#include "hvc_host.h"

namespace hvc {

constexpr int BM = 128, BN = 128, BK = 64;
constexpr int kStages = 5;
constexpr int kTileBytes = BM * BK * 2;  // 16 KB, same for A and B
constexpr int kStageBytes = 2 * kTileBytes;
constexpr int kGemmThreads = 192;
constexpr int kAccStages = 2;
constexpr int kGemmSmem = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;

struct GemmKArgs {
  int M, N, K;
  int m_blocks, n_blocks, k_blocks, k_splits, kb_per_split;
  int epilogue, activation;
  void* out; long long ldo;
  void* out2; long long ldo2;
  const float* bias;
  const float* resid; long long ldr;
  const float* gate; long long gate_ld;
  int rows_per_batch;
  const bf16* aux; long long ldaux;
  float alpha;
  DropArg drop;   // fused nn.Dropout on the epilogue value (seed == nullptr: off)
};

struct WorkItem {
  int m0, n0, kb0, kb1;
};
__device__ __forceinline__ WorkItem decode_work(const GemmKArgs& p, int w) {
  const int tile = w / p.k_splits, split = w - tile * p.k_splits;
  const int mb = tile / p.n_blocks, nb = tile - mb * p.n_blocks;
  WorkItem it;
  it.m0 = mb * BM;
  it.n0 = nb * BN;
  it.kb0 = split * p.kb_per_split;
  it.kb1 = min(it.kb0 + p.kb_per_split, p.k_blocks);
  return it;
}

// nn.Dropout on the 32 values of one row chunk: element (row, col) of this site, see hvc_common.cuh
__device__ __forceinline__ void epilogue_dropout(const GemmKArgs& p, float (&v)[32], int row, int col0) {
  const DropCfg dc = drop_load(p.drop);
  const uint32_t rk = drop_rowkey(dc, static_cast<uint32_t>(row));
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = drop_keep(rk, static_cast<uint32_t>(col0 + j), dc.thr) ? v[j] * dc.inv_keep : 0.f;
}

// ---------------------------------------------------------------- epilogue for one 32-column chunk
__device__ __forceinline__ void epilogue_chunk(const GemmKArgs& p, const uint32_t (&acc)[32], int row, int col0) {
  if (row >= p.M || col0 >= p.N) return;
  const int ncols = min(32, p.N - col0);
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]) * p.alpha;
  if (p.bias != nullptr) {
    if (ncols == 32) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
      }
    } else {
      _Pragma("unroll") for (int j = 0; j < 32; ++j) if (j < ncols) v[j] += __ldg(p.bias + col0 + j);
    }
  }
  const bool vec = (ncols == 32);
  const bool drop = p.drop.seed != nullptr;
  if (drop && p.epilogue != HVC_EPI_BF16) epilogue_dropout(p, v, row, col0);   // residual / f32: the value before gate + residual

  if (p.epilogue == HVC_EPI_BF16) {
    if (p.out2 != nullptr) {
      bf16* o2 = reinterpret_cast<bf16*>(p.out2) + (long long)row * p.ldo2 + col0;
      if (vec && (p.ldo2 & 7) == 0) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint4 u = make_uint4(pack_bf16(v[j], v[j + 1]), pack_bf16(v[j + 2], v[j + 3]), pack_bf16(v[j + 4], v[j + 5]),
                               pack_bf16(v[j + 6], v[j + 7]));
          *reinterpret_cast<uint4*>(o2 + j) = u;
        }
      } else {
        _Pragma("unroll") for (int j = 0; j < 32; ++j) if (j < ncols) o2[j] = __float2bfloat16(v[j]);
      }
    }
    if (p.activation == HVC_ACT_GELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
    } else if (p.activation == HVC_ACT_GELU_GRAD) {
      const bf16* ax = p.aux + (long long)row * p.ldaux + col0;
      if (vec && (p.ldaux & 7) == 0) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          const uint4 u = __ldg(reinterpret_cast<const uint4*>(ax + j));
          const float2 a0 = unpack_bf16(u.x), a1 = unpack_bf16(u.y), a2 = unpack_bf16(u.z), a3 = unpack_bf16(u.w);
          v[j] *= gelu_erf_grad(a0.x); v[j + 1] *= gelu_erf_grad(a0.y);
          v[j + 2] *= gelu_erf_grad(a1.x); v[j + 3] *= gelu_erf_grad(a1.y);
          v[j + 4] *= gelu_erf_grad(a2.x); v[j + 5] *= gelu_erf_grad(a2.y);
          v[j + 6] *= gelu_erf_grad(a3.x); v[j + 7] *= gelu_erf_grad(a3.y);
        }
      } else {
        _Pragma("unroll") for (int j = 0; j < 32; ++j) if (j < ncols) v[j] *= gelu_erf_grad(__bfloat162float(ax[j]));
      }
    }
    if (drop) epilogue_dropout(p, v, row, col0);   // after the activation (mlp: Linear -> GELU -> Dropout); out2 stays pre-activation
    bf16* o = reinterpret_cast<bf16*>(p.out) + (long long)row * p.ldo + col0;
    if (vec && (p.ldo & 7) == 0) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 u = make_uint4(pack_bf16(v[j], v[j + 1]), pack_bf16(v[j + 2], v[j + 3]), pack_bf16(v[j + 4], v[j + 5]),
                             pack_bf16(v[j + 6], v[j + 7]));
        *reinterpret_cast<uint4*>(o + j) = u;
      }
    } else {
      _Pragma("unroll") for (int j = 0; j < 32; ++j) if (j < ncols) o[j] = __float2bfloat16(v[j]);
    }
  } else if (p.epilogue == HVC_EPI_RESIDUAL) {
    if (p.out2 != nullptr) {
      bf16* o2 = reinterpret_cast<bf16*>(p.out2) + (long long)row * p.ldo2 + col0;
      if (vec && (p.ldo2 & 7) == 0) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint4 u = make_uint4(pack_bf16(v[j], v[j + 1]), pack_bf16(v[j + 2], v[j + 3]), pack_bf16(v[j + 4], v[j + 5]),
                               pack_bf16(v[j + 6], v[j + 7]));
          *reinterpret_cast<uint4*>(o2 + j) = u;
        }
      } else {
        _Pragma("unroll") for (int j = 0; j < 32; ++j) if (j < ncols) o2[j] = __float2bfloat16(v[j]);
      }
    }
    const float* r = p.resid + (long long)row * p.ldr + col0;
    float* o = reinterpret_cast<float*>(p.out) + (long long)row * p.ldo + col0;
    const float* g = p.gate ? p.gate + (long long)(row / p.rows_per_batch) * p.gate_ld + col0 : nullptr;
    if (vec && (p.ldr & 3) == 0 && (p.ldo & 3) == 0 && (g == nullptr || (p.gate_ld & 3) == 0)) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 rr = __ldg(reinterpret_cast<const float4*>(r + j));
        float4 gg = make_float4(1.f, 1.f, 1.f, 1.f);
        if (g) gg = __ldg(reinterpret_cast<const float4*>(g + j));
        float4 oo;
        oo.x = fmaf(gg.x, v[j], rr.x); oo.y = fmaf(gg.y, v[j + 1], rr.y);
        oo.z = fmaf(gg.z, v[j + 2], rr.z); oo.w = fmaf(gg.w, v[j + 3], rr.w);
        *reinterpret_cast<float4*>(o + j) = oo;
      }
    } else {
      _Pragma("unroll") for (int j = 0; j < 32; ++j) if (j < ncols) o[j] = fmaf(g ? g[j] : 1.f, v[j], r[j]);
    }
  } else if (p.epilogue == HVC_EPI_F32_ATOMIC) {
    float* o = reinterpret_cast<float*>(p.out) + (long long)row * p.ldo + col0;
    if (vec && (p.ldo & 3) == 0) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) atomicAdd(reinterpret_cast<float4*>(o + j), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
    } else {
      _Pragma("unroll") for (int j = 0; j < 32; ++j) if (j < ncols) atomicAdd(o + j, v[j]);
    }
  } else {  // HVC_EPI_F32
    if (p.activation == HVC_ACT_GELU) {   // fp32 verification mode: the MLP hidden activation stays fp32
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
    }
    float* o = reinterpret_cast<float*>(p.out) + (long long)row * p.ldo + col0;
    if (vec && (p.ldo & 3) == 0) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
      _Pragma("unroll") for (int j = 0; j < 32; ++j) if (j < ncols) o[j] = v[j];
    }
  }
}

// ---------------------------------------------------------------- kernel
template <int A_MAJOR, int B_MAJOR>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmKArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;
  uint64_t* tempty_bar = tfull_bar + kAccStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + kAccStages);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_work = p.m_blocks * p.n_blocks * p.k_splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < kAccStages; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kAccStages * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    uint32_t stage = 0, phase = 0;
    for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
      const WorkItem it = decode_work(p, w);
      for (int kb = it.kb0; kb < it.kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1, 1);
        if (elect_one()) {
          uint8_t* sA = smem + stage * kStageBytes;
          uint8_t* sB = sA + kTileBytes;
          mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
          if (A_MAJOR == kMajorK) {
            tma_load_2d(sA, &tmA, &full_bar[stage], kb * BK, it.m0);
          } else {
            tma_load_2d(sA, &tmA, &full_bar[stage], it.m0, kb * BK);
            tma_load_2d(sA + kTileBytes / 2, &tmA, &full_bar[stage], it.m0 + 64, kb * BK);
          }
          if (B_MAJOR == kMajorK) {
            tma_load_2d(sB, &tmB, &full_bar[stage], kb * BK, it.n0);
          } else {
            tma_load_2d(sB, &tmB, &full_bar[stage], it.n0, kb * BK);
            tma_load_2d(sB + kTileBytes / 2, &tmB, &full_bar[stage], it.n0 + 64, kb * BK);
          }
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MAJOR, B_MAJOR);
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
    for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
      const WorkItem it = decode_work(p, w);
      if (it.kb0 >= it.kb1) continue;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1, 2);
      tc_fence_after();
      for (int kb = it.kb0; kb < it.kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase, 3);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a0 = smem_u32(smem + stage * kStageBytes);
          const uint32_t b0 = a0 + kTileBytes;
#pragma unroll
          for (int k16 = 0; k16 < BK / 16; ++k16) {
            const uint64_t ad = (A_MAJOR == kMajorK) ? make_sdesc_sw128(a0 + k16 * 32, 16, 1024)
                                                     : make_sdesc_sw128(a0 + k16 * 2048, kTileBytes / 2, 1024);
            const uint64_t bd = (B_MAJOR == kMajorK) ? make_sdesc_sw128(b0 + k16 * 32, 16, 1024)
                                                     : make_sdesc_sw128(b0 + k16 * 2048, kTileBytes / 2, 1024);
            umma_ss(tmem_base + acc * BN, ad, bd, idesc, (kb > it.kb0 || k16 > 0) ? 1u : 0u);
          }
          tc_commit(&empty_bar[stage]);
          if (kb == it.kb1 - 1) tc_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    uint32_t acc = 0, acc_phase = 0;
    for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
      const WorkItem it = decode_work(p, w);
      if (it.kb0 >= it.kb1) continue;
      mbar_wait(&tfull_bar[acc], acc_phase, 4);
      tc_fence_after();
      const int row = it.m0 + quarter * 32 + lane;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + c * 32, v);
        tmem_ld_wait();
        epilogue_chunk(p, v, row, it.n0 + c * 32);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kAccStages * BN);
  }
}

template <int A_MAJOR, int B_MAJOR>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmKArgs& ka, int grid, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    HVC_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<A_MAJOR, B_MAJOR>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
    configured = true;
  }
  gemm_bf16_kernel<A_MAJOR, B_MAJOR><<<grid, kGemmThreads, kGemmSmem, st>>>(tmA, tmB, ka);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

}  // namespace hvc

extern "C" int hvc_gemm(const hvc_gemm_args* a, void* stream) {
  using namespace hvc;
  HVC_CHECK_ARG(a != nullptr && a->size == sizeof(hvc_gemm_args), "hvc_gemm: bad args struct (size %u, expected %zu)",
                a ? a->size : 0u, sizeof(hvc_gemm_args));
  HVC_CHECK_ARG(a->M > 0 && a->N > 0 && a->K > 0, "hvc_gemm: empty problem %dx%dx%d", a->M, a->N, a->K);
  HVC_CHECK_ARG(a->A && a->B && a->out, "hvc_gemm: null operand");
  HVC_CHECK_ARG(a->epilogue >= HVC_EPI_BF16 && a->epilogue <= HVC_EPI_F32, "hvc_gemm: unknown epilogue %d", a->epilogue);
  HVC_CHECK_ARG(a->epilogue != HVC_EPI_RESIDUAL || a->resid != nullptr, "hvc_gemm: residual epilogue without resid");
  HVC_CHECK_ARG(a->activation != HVC_ACT_GELU_GRAD || a->aux != nullptr, "hvc_gemm: GELU_GRAD without aux");
  HVC_CHECK_ARG(a->gate == nullptr || a->rows_per_batch > 0, "hvc_gemm: gate without rows_per_batch");
  const int k_splits_req = a->k_splits < 1 ? 1 : a->k_splits;
  HVC_CHECK_ARG(k_splits_req == 1 || a->epilogue == HVC_EPI_F32_ATOMIC, "hvc_gemm: k_splits>1 needs the atomic epilogue");

  CUtensorMap tmA, tmB;
  int rc;
  if (a->a_major == 0) rc = make_tmap_2d(&tmA, a->A, 2, a->M, a->K, a->lda, BK, BM, true);
  else                 rc = make_tmap_2d(&tmA, a->A, 2, a->K, a->M, a->lda, 64, BK, true);
  if (rc) return rc;
  if (a->b_major == 0) rc = make_tmap_2d(&tmB, a->B, 2, a->N, a->K, a->ldb, BK, BN, true);
  else                 rc = make_tmap_2d(&tmB, a->B, 2, a->K, a->N, a->ldb, 64, BK, true);
  if (rc) return rc;

  GemmKArgs ka;
  ka.M = a->M; ka.N = a->N; ka.K = a->K;
  ka.m_blocks = (a->M + BM - 1) / BM;
  ka.n_blocks = (a->N + BN - 1) / BN;
  ka.k_blocks = (a->K + BK - 1) / BK;
  int ks = k_splits_req < ka.k_blocks ? k_splits_req : ka.k_blocks;
  ka.kb_per_split = (ka.k_blocks + ks - 1) / ks;
  ka.k_splits = (ka.k_blocks + ka.kb_per_split - 1) / ka.kb_per_split;  // no empty splits
  ka.epilogue = a->epilogue; ka.activation = a->activation;
  ka.out = a->out; ka.ldo = a->ldo; ka.out2 = a->out2; ka.ldo2 = a->ldo2;
  ka.bias = a->bias; ka.resid = a->resid; ka.ldr = a->ldr;
  ka.gate = a->gate; ka.gate_ld = a->gate_ld; ka.rows_per_batch = a->rows_per_batch > 0 ? a->rows_per_batch : 1;
  ka.aux = reinterpret_cast<const bf16*>(a->aux); ka.ldaux = a->ldaux;
  ka.alpha = a->alpha;
  ka.drop = make_drop(a->drop);
  HVC_CHECK_ARG(ka.drop.seed == nullptr || (a->epilogue != HVC_EPI_F32_ATOMIC && a->drop.p < 1.f), "hvc_gemm: dropout needs p < 1 and a non-atomic epilogue");

  const long long num_work = (long long)ka.m_blocks * ka.n_blocks * ka.k_splits;
  const int sms = device_sm_count();
  const int grid = (int)(num_work < sms ? num_work : sms);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (a->a_major == 0 && a->b_major == 0) return launch_gemm<kMajorK, kMajorK>(tmA, tmB, ka, grid, st);
  if (a->a_major == 0 && a->b_major == 1) return launch_gemm<kMajorK, kMajorMN>(tmA, tmB, ka, grid, st);
  if (a->a_major == 1 && a->b_major == 0) return launch_gemm<kMajorMN, kMajorK>(tmA, tmB, ka, grid, st);
  return launch_gemm<kMajorMN, kMajorMN>(tmA, tmB, ka, grid, st);
}
