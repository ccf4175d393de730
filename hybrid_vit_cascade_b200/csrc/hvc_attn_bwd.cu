// hvc_attn_bwd.cu -- fused flash-style attention backward for sm_100a.
//
// Autograd backward of  out = softmax(q k^T * scale) v  (vit_components.py:46-51, :103-113) with the
// probabilities recomputed from the saved log-sum-exp; nothing of size (N x M) touches HBM.
//
// Kernels:
//   attn_delta_kernel   delta[b,h,q] = sum_d dO * O                       (HBM-bound pre-pass)
//   attn_bwd_kernel     one CTA = one (batch, head, 128-key tile), loops over 128-query tiles:
//        S^T  = K Q^T                (SS)            P^T = exp2(S^T*scale2 - lse2)      -> TMEM (bf16, over S^T)
//        dP^T = V dO^T               (SS)            dS^T = P^T * (dP^T - delta) * scale -> smem (bf16, swizzled)
//        dV  += P^T  dO              (TS, dO MN-major)
//        dK  += dS^T Q               (SS, dS^T K-major, Q MN-major)
//        dQ_i = dS   K               (SS, dS^T read MN-major as A, K MN-major)  -> fp32 atomics into dq_accum
//      thread == key row == TMEM lane, so the elementwise stage needs no shuffles; the single swizzled
//      smem copy of dS^T serves both the dK (K-major) and the dQ (MN-major) products.
//   attn_dq_convert_kernel   dq_accum (f32, per head) -> dq (bf16, packed token-major)
#include "hvc_common.cuh"
#include "hvc_host.h"

namespace hvc {

constexpr int kBwdThreads = 192;
constexpr int kBT = 128;  // tile edge (queries and keys)

struct AttnBwdKArgs {
  int batch, heads, nq, nk, nq_pad, n_q_tiles;
  const float* lse2; const float* delta;   // [B, H, nq_pad]
  float* dq_accum;                          // [B, H, nq_pad, HD]
  bf16* dk; long long lddk;
  bf16* dv; long long lddv;
  float scale, scale2;
};

template <int HD>
struct BwdSmem {
  static constexpr int kTile = kBT * HD * 2;                 // 16 KB
  static constexpr int kK = 0;
  static constexpr int kV = kTile;
  static constexpr int kQStage = 2 * kTile + 2048;           // Q, dO, lse2[128], delta[128] (+pad to 1 KB multiple)
  static constexpr int kQ = 2 * kTile;
  static constexpr int kDS = kQ + 2 * kQStage;               // 2 x [128 x 128] bf16
  static constexpr int kBar = kDS + 2 * (kBT * kBT * 2);
  static constexpr int kTotal = kBar + 256 + 1024;
};

enum { BB_KV = 0, BB_QF = 1, BB_QE = 3, BB_ST = 5, BB_DS = 6, BB_DQF = 7, BB_DQE = 8, BB_N = 9 };

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int HD>
__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO, const AttnBwdKArgs p) {
  static_assert(HD == 64, "head_dim 64 only for now");
  using L = BwdSmem<HD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L::kBar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + BB_N);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int b = bh / p.heads, h = bh - b * p.heads;
  const int j = blockIdx.x;           // key tile
  const int nQ = p.n_q_tiles;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO);
    mbar_init(&bar[BB_KV], 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&bar[BB_QF + s], 1); mbar_init(&bar[BB_QE + s], 1); }
    mbar_init(&bar[BB_ST], 1);
    mbar_init(&bar[BB_DS], 128);
    mbar_init(&bar[BB_DQF], 1);
    mbar_init(&bar[BB_DQE], 128);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t kColSt = 0, kColDPt = 128, kColDV = 256, kColDK = 320, kColDQ = 384;

  if (warp == 4) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(&bar[BB_KV], 2 * L::kTile);
      tma_load_2d(smem + L::kK, &tmK, &bar[BB_KV], h * HD, b * p.nk + j * kBT, kEvictFirst);
      tma_load_2d(smem + L::kV, &tmV, &bar[BB_KV], h * HD, b * p.nk + j * kBT, kEvictFirst);
      for (int i = 0; i < nQ; ++i) {
        const int st = i & 1;
        const uint32_t ph = (i >> 1) & 1;
        uint8_t* base = smem + L::kQ + st * L::kQStage;
        mbar_wait(&bar[BB_QE + st], ph ^ 1, 10);
        mbar_arrive_expect_tx(&bar[BB_QF + st], 2 * L::kTile + 2 * kBT * 4);
        tma_load_2d(base, &tmQ, &bar[BB_QF + st], h * HD, b * p.nq + i * kBT, kEvictLast);
        tma_load_2d(base + L::kTile, &tmDO, &bar[BB_QF + st], h * HD, b * p.nq + i * kBT, kEvictLast);
        bulk_load_1d(base + 2 * L::kTile, p.lse2 + (long long)bh * p.nq_pad + i * kBT, kBT * 4, &bar[BB_QF + st]);
        bulk_load_1d(base + 2 * L::kTile + kBT * 4, p.delta + (long long)bh * p.nq_pad + i * kBT, kBT * 4, &bar[BB_QF + st]);
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t id_s = make_idesc_bf16(kBT, kBT, kMajorK, kMajorK);      // S^T, dP^T
      constexpr uint32_t id_dv = make_idesc_bf16(kBT, HD, kMajorK, kMajorMN);     // dV (A in TMEM), dK (A K-major smem)
      constexpr uint32_t id_dq = make_idesc_bf16(kBT, HD, kMajorMN, kMajorMN);    // dQ
      const uint32_t sK = smem_u32(smem + L::kK), sV = smem_u32(smem + L::kV);
      const uint32_t sQ0 = smem_u32(smem + L::kQ), sDS0 = smem_u32(smem + L::kDS);
      auto issue_s_dp = [&](int st) {
        const uint32_t sQ = sQ0 + st * L::kQStage, sDO = sQ + L::kTile;
#pragma unroll
        for (int k16 = 0; k16 < HD / 16; ++k16)
          umma_ss(tmem_base + kColSt, make_sdesc_sw128(sK + k16 * 32, 16, 1024), make_sdesc_sw128(sQ + k16 * 32, 16, 1024), id_s,
                  k16 > 0 ? 1u : 0u);
#pragma unroll
        for (int k16 = 0; k16 < HD / 16; ++k16)
          umma_ss(tmem_base + kColDPt, make_sdesc_sw128(sV + k16 * 32, 16, 1024), make_sdesc_sw128(sDO + k16 * 32, 16, 1024), id_s,
                  k16 > 0 ? 1u : 0u);
      };
      mbar_wait(&bar[BB_KV], 0, 20);
      mbar_wait(&bar[BB_QF + 0], 0, 21);
      tc_fence_after();
      issue_s_dp(0);
      tc_commit(&bar[BB_ST]);
      for (int i = 0; i < nQ; ++i) {
        const int st = i & 1;
        const uint32_t sQ = sQ0 + st * L::kQStage, sDO = sQ + L::kTile;
        const uint32_t sDS = sDS0 + (i & 1) * (kBT * kBT * 2);
        mbar_wait(&bar[BB_DS], i & 1, 22);
        tc_fence_after();
        // dV += P^T dO     (A = P^T in TMEM over S^T, 8 columns per K=16 step; B = dO MN-major)
#pragma unroll
        for (int k16 = 0; k16 < kBT / 16; ++k16)
          umma_ts(tmem_base + kColDV, tmem_base + kColSt + k16 * 8, make_sdesc_sw128(sDO + k16 * 2048, 8192, 1024), id_dv,
                  (i > 0 || k16 > 0) ? 1u : 0u);
        if (i + 1 < nQ) {
          mbar_wait(&bar[BB_QF + (st ^ 1)], ((i + 1) >> 1) & 1, 23);
          tc_fence_after();
          issue_s_dp(st ^ 1);
          tc_commit(&bar[BB_ST]);
        }
        // dK += dS^T Q     (A = dS^T K-major: two 64-query sub-tiles; B = Q MN-major)
#pragma unroll
        for (int k16 = 0; k16 < kBT / 16; ++k16)
          umma_ss(tmem_base + kColDK, make_sdesc_sw128(sDS + (k16 >> 2) * 16384 + (k16 & 3) * 32, 16, 1024),
                  make_sdesc_sw128(sQ + k16 * 2048, 8192, 1024), id_dv, (i > 0 || k16 > 0) ? 1u : 0u);
        if (i > 0) { mbar_wait(&bar[BB_DQE], (i - 1) & 1, 24); tc_fence_after(); }
        // dQ_i = dS K      (A = dS^T read MN-major: M = queries contiguous, K = key rows; B = K MN-major)
#pragma unroll
        for (int k16 = 0; k16 < kBT / 16; ++k16)
          umma_ss(tmem_base + kColDQ, make_sdesc_sw128(sDS + k16 * 2048, 16384, 1024),
                  make_sdesc_sw128(sK + k16 * 2048, 8192, 1024), id_dq, k16 > 0 ? 1u : 0u);
        tc_commit(&bar[BB_DQF]);
        tc_commit(&bar[BB_QE + st]);
      }
    }
  } else if (warp < 4) {
    // ===================== elementwise warpgroup: thread == key row =====================
    const int quarter = warp;
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tSt = tmem_base + lane_base + kColSt, tDPt = tmem_base + lane_base + kColDPt;
    const bool key_ok = (j * kBT + r) < p.nk;
    const float scale = p.scale, scale2 = p.scale2;

    auto drain_dq = [&](int i) {
      mbar_wait(&bar[BB_DQF], i & 1, 31);
      tc_fence_after();
      const int q = i * kBT + r;     // here the lane is a query row of tile i
      float* dst = p.dq_accum + ((long long)bh * p.nq_pad + q) * HD;
#pragma unroll
      for (int c = 0; c < HD; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + lane_base + kColDQ + c, v);
        tmem_ld_wait();
        if (q < p.nq) {
#pragma unroll
          for (int t = 0; t < 32; t += 4)
            atomicAdd(reinterpret_cast<float4*>(dst + c + t),
                      make_float4(__uint_as_float(v[t]), __uint_as_float(v[t + 1]), __uint_as_float(v[t + 2]), __uint_as_float(v[t + 3])));
        }
      }
      tc_fence_before();
      mbar_arrive(&bar[BB_DQE]);
    };

    for (int i = 0; i < nQ; ++i) {
      const int st = i & 1;
      const uint8_t* stage = smem + L::kQ + st * L::kQStage;
      const float* s_lse = reinterpret_cast<const float*>(stage + 2 * L::kTile);
      const float* s_delta = s_lse + kBT;
      uint8_t* dsbuf = smem + L::kDS + (i & 1) * (kBT * kBT * 2);
      mbar_wait(&bar[BB_QF + st], (i >> 1) & 1, 32);   // lse/delta for this query tile are in smem
      mbar_wait(&bar[BB_ST], i & 1, 33);
      tc_fence_after();
      const int q_valid = p.nq - i * kBT;               // columns >= q_valid are padding
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t sv[32], dv[32];
        tmem_ld_32x32(tSt + c * 32, sv);
        tmem_ld_32x32(tDPt + c * 32, dv);
        tmem_ld_wait();
        uint32_t ppk[16], dpk[16];
#pragma unroll
        for (int t = 0; t < 32; t += 4) {
          const float4 l4 = *reinterpret_cast<const float4*>(s_lse + c * 32 + t);
          const float4 d4 = *reinterpret_cast<const float4*>(s_delta + c * 32 + t);
          const float lse[4] = {l4.x, l4.y, l4.z, l4.w};
          const float dl[4] = {d4.x, d4.y, d4.z, d4.w};
          float pv[4], ds[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const bool ok = key_ok && (c * 32 + t + e) < q_valid;
            const float pe = ex2_approx(fmaf(__uint_as_float(sv[t + e]), scale2, -lse[e]));
            pv[e] = ok ? pe : 0.f;
            ds[e] = ok ? pe * (__uint_as_float(dv[t + e]) - dl[e]) * scale : 0.f;
          }
          ppk[(t >> 1)] = pack_bf16(pv[0], pv[1]);
          ppk[(t >> 1) + 1] = pack_bf16(pv[2], pv[3]);
          dpk[(t >> 1)] = pack_bf16(ds[0], ds[1]);
          dpk[(t >> 1) + 1] = pack_bf16(ds[2], ds[3]);
        }
        tmem_st_32x16(tSt + c * 16, ppk);   // P^T over the S^T columns already consumed
        uint8_t* sub = dsbuf + (c >> 1) * 16384;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t off = sw128_offset(r, (c & 1) * 4 + k);
          *reinterpret_cast<uint4*>(sub + off) = make_uint4(dpk[4 * k], dpk[4 * k + 1], dpk[4 * k + 2], dpk[4 * k + 3]);
        }
      }
      tmem_st_wait();
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&bar[BB_DS]);
      if (i > 0) drain_dq(i - 1);
    }
    drain_dq(nQ - 1);   // also implies dV and dK are complete (commit covers all earlier MMAs)

    // ---- epilogue: dV, dK -> bf16 -> global
    const int key = j * kBT + r;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      bf16* dst = (which == 0 ? p.dv + (long long)(b * p.nk + key) * p.lddv : p.dk + (long long)(b * p.nk + key) * p.lddk) + h * HD;
      const uint32_t tcol = tmem_base + lane_base + (which == 0 ? kColDV : kColDK);
#pragma unroll
      for (int c = 0; c < HD; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tcol + c, v);
        tmem_ld_wait();
        if (key_ok) {
#pragma unroll
          for (int t = 0; t < 32; t += 8) {
            uint4 u;
            u.x = pack_bf16(__uint_as_float(v[t]), __uint_as_float(v[t + 1]));
            u.y = pack_bf16(__uint_as_float(v[t + 2]), __uint_as_float(v[t + 3]));
            u.z = pack_bf16(__uint_as_float(v[t + 4]), __uint_as_float(v[t + 5]));
            u.w = pack_bf16(__uint_as_float(v[t + 6]), __uint_as_float(v[t + 7]));
            *reinterpret_cast<uint4*>(dst + c + t) = u;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// delta[b,h,q] = sum_d dO[b,q,h,d] * O[b,q,h,d]; one warp per (token, head), HD = 64 -> 2 elements per lane
__global__ void __launch_bounds__(256) attn_delta_kernel(const bf16* __restrict__ o, long long ldo, const bf16* __restrict__ d_o,
                                                         long long lddo, float* __restrict__ delta, int batch, int heads, int nq,
                                                         int nq_pad) {
  const long long gw = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const long long total = (long long)batch * nq * heads;
  if (gw >= total) return;
  const int h = (int)(gw % heads);
  const long long tok = gw / heads;
  const int b = (int)(tok / nq), q = (int)(tok - (long long)b * nq);
  const float2 a = unpack_bf16(__ldg(reinterpret_cast<const uint32_t*>(o + tok * ldo + h * 64 + 2 * lane)));
  const float2 g = unpack_bf16(__ldg(reinterpret_cast<const uint32_t*>(d_o + tok * lddo + h * 64 + 2 * lane)));
  const float s = warp_sum(a.x * g.x + a.y * g.y);
  if (lane == 0) delta[((long long)b * heads + h) * nq_pad + q] = s;
}

// dq_accum f32 [B,H,nq_pad,64] -> dq bf16 [B*nq, lddq] (column h*64 + d); 8 elements per thread
__global__ void __launch_bounds__(256) attn_dq_convert_kernel(const float* __restrict__ acc, bf16* __restrict__ dq, long long lddq,
                                                              int batch, int heads, int nq, int nq_pad) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;   // over B*nq*H*8
  const long long total = (long long)batch * nq * heads * 8;
  if (idx >= total) return;
  const int part = (int)(idx & 7);
  const long long t = idx >> 3;
  const int h = (int)(t % heads);
  const long long tok = t / heads;
  const int b = (int)(tok / nq), q = (int)(tok - (long long)b * nq);
  const float* src = acc + (((long long)b * heads + h) * nq_pad + q) * 64 + part * 8;
  const float4 x = __ldg(reinterpret_cast<const float4*>(src)), y = __ldg(reinterpret_cast<const float4*>(src + 4));
  uint4 u = make_uint4(pack_bf16(x.x, x.y), pack_bf16(x.z, x.w), pack_bf16(y.x, y.y), pack_bf16(y.z, y.w));
  *reinterpret_cast<uint4*>(dq + tok * lddq + h * 64 + part * 8) = u;
}

}  // namespace hvc

extern "C" int hvc_attn_bwd(const hvc_attn_args* a, void* stream) {
  using namespace hvc;
  HVC_CHECK_ARG(a != nullptr && a->size == sizeof(hvc_attn_args), "hvc_attn_bwd: bad args struct");
  HVC_CHECK_ARG(a->batch > 0 && a->heads > 0 && a->nq > 0 && a->nk > 0, "hvc_attn_bwd: empty problem");
  HVC_CHECK_ARG(a->head_dim == 64, "hvc_attn_bwd: head_dim %d not supported (64 only)", a->head_dim);
  HVC_CHECK_ARG(a->q && a->k && a->v && a->o && a->d_o && a->lse && a->delta && a->dq_accum && a->dq && a->dk && a->dv,
                "hvc_attn_bwd: null operand");
  HVC_CHECK_ARG(((a->lddq | a->lddk | a->lddv) & 7) == 0, "hvc_attn_bwd: gradient row pitches must be multiples of 8");
  constexpr int HD = 64;
  using L = BwdSmem<HD>;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int nq_pad = (a->nq + 127) / 128 * 128;
  const uint64_t width = (uint64_t)a->heads * HD;

  {
    const long long warps = (long long)a->batch * a->nq * a->heads;
    attn_delta_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(reinterpret_cast<const bf16*>(a->o), a->ldo,
                                                                  reinterpret_cast<const bf16*>(a->d_o), a->lddo, a->delta, a->batch,
                                                                  a->heads, a->nq, nq_pad);
    HVC_LAUNCH_CHECK();
  }
  CUtensorMap tmQ, tmK, tmV, tmDO;
  int rc;
  if ((rc = make_tmap_2d(&tmQ, a->q, 2, (uint64_t)a->batch * a->nq, width, a->ldq, HD, kBT, true))) return rc;
  if ((rc = make_tmap_2d(&tmDO, a->d_o, 2, (uint64_t)a->batch * a->nq, width, a->lddo, HD, kBT, true))) return rc;
  if ((rc = make_tmap_2d(&tmK, a->k, 2, (uint64_t)a->batch * a->nk, width, a->ldk, HD, kBT, true))) return rc;
  if ((rc = make_tmap_2d(&tmV, a->v, 2, (uint64_t)a->batch * a->nk, width, a->ldv, HD, kBT, true))) return rc;
  AttnBwdKArgs ka;
  ka.batch = a->batch; ka.heads = a->heads; ka.nq = a->nq; ka.nk = a->nk; ka.nq_pad = nq_pad;
  ka.n_q_tiles = nq_pad / kBT;
  ka.lse2 = a->lse; ka.delta = a->delta; ka.dq_accum = a->dq_accum;
  ka.dk = reinterpret_cast<bf16*>(a->dk); ka.lddk = a->lddk;
  ka.dv = reinterpret_cast<bf16*>(a->dv); ka.lddv = a->lddv;
  ka.scale = a->scale; ka.scale2 = a->scale * 1.4426950408889634f;
  static bool configured = false;
  if (!configured) {
    HVC_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    configured = true;
  }
  dim3 grid((a->nk + kBT - 1) / kBT, a->batch * a->heads);
  attn_bwd_kernel<HD><<<grid, kBwdThreads, L::kTotal, st>>>(tmQ, tmK, tmV, tmDO, ka);
  HVC_LAUNCH_CHECK();
  {
    const long long threads = (long long)a->batch * a->nq * a->heads * 8;
    attn_dq_convert_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(a->dq_accum, reinterpret_cast<bf16*>(a->dq), a->lddq,
                                                                             a->batch, a->heads, a->nq, nq_pad);
    HVC_LAUNCH_CHECK();
  }
  return HVC_OK;
}
