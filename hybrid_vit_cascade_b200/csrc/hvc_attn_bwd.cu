// hvc_attn_bwd.cu -- fused flash-style attention backward for sm_100a.
//
// Autograd backward of  out = softmax(q k^T * scale) v  (vit_components.py:46-51, :103-113) with the
// probabilities recomputed from the saved log-sum-exp; nothing of size (N x M) touches HBM.
//
// Kernels:
//   attn_delta_kernel   delta[b,h,q] = sum_d dO * O                       (HBM-bound pre-pass)
//   attn_bwd_kernel     one CTA = one (batch, head, 128-key tile), loops over 128-query tiles i:
//        S^T  = K Q_i^T   (SS)      phase A: P^T = exp2(S^T*scale2 - lse2)          -> TMEM (bf16, own columns)
//        dP^T = V dO_i^T  (SS)      phase B: dS^T = P^T * (dP^T - delta)            -> smem (bf16, swizzled, 2 buffers)
//        dV  += P^T  dO_i (TS, dO MN-major)
//        dK  += dS^T Q_i  (d=64: TS, dS^T as bf16 pairs in TMEM over the dP^T columns; d=32: SS, dS^T K-major; Q MN-major)   (scale in the epilogue)
//        dQ_i = dS   K    (SS, dS^T read MN-major as A, K^T K-major) -> TMA reduce-add (fp32) into dq_accum
//      320 threads: two elementwise warpgroups (thread == key row == TMEM lane; warpgroup x owns query columns
//      [64x, 64x+64) of every tile), a TMA warp and an MMA warp.  Design points, each from a measurement on B200:
//      * tcgen05.mma with a K-major A operand in shared memory costs ~108 clk per 128x128x16 step, an MN-major A
//        ~65 clk (profiles/r01_umma_bench_by_operand_layout.log).  K and V stay fixed for the whole CTA, so they are
//        transposed ONCE into MN-major K^T / V^T tiles; S^T and dP^T then run at the fast rate, and the same K^T tile
//        is the (K-major) B operand of dQ.
//      * the v3 kernel ran both warpgroups in lockstep on shared barriers: the exp unit was saturated during phase A
//        and idle for the rest of the tile, and the loop-carried chain  dS(i-1) -> dP^T(i), dK(i-1), dQ(i-1) -> drain
//        dQ -> phase B(i) -> dS(i)  left the tensor pipe idle ~1450 of 3800 clk per tile
//        (profiles/r01_attn_bwd_timeline_d64_v3.log).  Now every product except dQ is issued per column half with its
//        own barriers, the warpgroups take turns on the exp unit (named barriers, as in the forward kernel) so that one
//        is in phase A while the other is in phase B / draining, dS^T is double-buffered and dQ(i-1) is drained after
//        phase B(i).
//      * four warpgroups of 32 columns (more warps per scheduler) were tried and lost: the per-thread fixed cost of a
//        tile (waits, fences, address arithmetic, ~150 instructions) then weighs as much as the arithmetic
//        (profiles/r01_attn_bwd_timeline_d64_v4a.log).
//      * dQ goes to HBM through the TMA engine (cp.reduce.async.bulk.tensor, whole 128-byte lines): red.global.add.v4.f32
//        straight from registers needs no staging but hits the L2 atomic rate (600 vs 794 TFLOP/s at d=64, B=8).
//      * dQ tiles are staged for the TMA reduce in the dS^T sub-tile that the same warpgroup wrote two tiles earlier
//        and dK/dQ have released, which is what lets three Q/dO stages and two dS^T buffers fit in shared memory.
//      * round 2: dQ(i) is drained by its OWN warpgroup (TMEM -> registers -> swizzled staging tile -> TMA reduce-add).  In round 1 each
//        elementwise warpgroup drained its half of dQ(i-1) after phase B(i): ~380 clk waiting for the dQ product plus ~470 clk of drain sat
//        on the serial path of every tile (profiles/r01_attn_bwd_timeline_d64_v5.log).  Upper bound measured by dropping the drain
//        altogether: 793 -> 933 TFLOP/s at d = 64 (646 -> 786 with dropout), profiles/r02_attn_bwd_nodrain_bound.log.  The block grows to
//        512 threads; setmaxnreg moves registers from the drain / TMA / MMA warps to the elementwise warpgroups (168 each).
//      * round 2: the kernel is bound by the tensor pipe's in-order execution when dropout is off (dropping the dV products: +12 %, the dQ
//        products: +21 %, replacing the exp by a multiply: 0 %; profiles/r02_attn_bwd_limiter_experiments.log), and a tcgen05.mma step costs
//        (A + B operand bytes) / 128 B/clk.  dP^T is therefore ONE N = 128 product per tile (8 KB per step) instead of two N = 64 ones
//        (2 x 6 KB), issued in the tail of the previous tile, and S^T(i+1) is issued before dV_0(i) so the pipe never waits for a
//        warpgroup: 790 -> 847 TFLOP/s at d = 64, 451 -> 486 at d = 32 (dropout off; with dropout the elementwise path is the bound: +1 %).
//      TMEM is used to the last column (S^T 128 | P^T 64 | dP^T 128 | dV d | dK d | dQ d).
//   attn_dq_convert_kernel   dq_accum (f32, per head) * scale -> dq (bf16, packed token-major)
#define HVC_WAIT_INLINE 1      // no device function calls in this translation unit: see mbar_wait_slow
#include "hvc_common.cuh"
#include "hvc_host.h"

namespace hvc {

constexpr int kBwdWGs = 2;                        // elementwise warpgroups
// warps 0-7: elementwise warpgroups; 8-11: dQ drain warpgroup; 12: TMA; 13: MMA; 14-15: idle (setmaxnreg works on whole warpgroups)
constexpr int kBwdThreads = 512;
// Register budget: the block launches with 65536 / 512 = 128 registers per thread; the drain warpgroup and the TMA / MMA warpgroup give
// registers back (setmaxnreg.dec), the elementwise warpgroups take them (setmaxnreg.inc): 128*64 + 128*104 + 256*168 = 64512.
constexpr int kDQChunk = 16, kDQBufs = 3;          // dQ leaves in 16-column chunks through a ring of three staging buffers
constexpr int kRegsAux = 64, kRegsDrain = 104, kRegsEw = 168;
constexpr int kBT = 128;                          // tile edge (queries and keys)
constexpr int kQStages = 3;
constexpr int kWgCols = kBT / kBwdWGs;            // query columns per warpgroup (64)
#ifndef HVC_BWD_LSE_PRE
#define HVC_BWD_LSE_PRE 2
#endif
constexpr int kLsePre = HVC_BWD_LSE_PRE;          // lse float4 loads issued ahead of their arithmetic in phase A
#ifndef HVC_BWD_DELTA_PRE
#define HVC_BWD_DELTA_PRE 4
#endif
constexpr int kDeltaPre = HVC_BWD_DELTA_PRE;      // delta float4 loads issued ahead of their arithmetic in phase B (even)
#ifndef HVC_BWD_EMU
#define HVC_BWD_EMU 0
#endif
constexpr int kBwdEmu = HVC_BWD_EMU;              // element pairs of every 16 whose exp2 is the FMA-pipe polynomial (0: MUFU has headroom here)

struct AttnBwdKArgs {
  int batch, heads, nq, nk, nq_pad, n_q_tiles;
  const float* nlse2; const float* delta;  // [B, H, nq_pad] each: -lse2 (written by the pre-pass) and rowsum(dO*O)
  bf16* dk; long long lddk;
  bf16* dv; long long lddv;
  float scale, scale2;
  float* dq_accum;           // [B, H, nq_pad, HD] f32 (HVC_BWD_DQ_RED variant only)
  const uint32_t* rowkeys;   // [B, H, nq_pad] dropout row keys (DROP instantiations; written by the pre-pass)
  DropArg drop;
};

template <int HD>
struct BwdSmem {
  static constexpr int kTile = kBT * HD * 2;                 // one 128 x HD bf16 tile (16 KB at HD 64)
  static constexpr int kKT = 0;                              // K^T: HD rows x 128 keys, two 64-key sub-tiles of 128-byte rows
  static constexpr int kVT = kTile;
  static constexpr int kQStage = 2 * kTile + 2048;           // Q, dO, lse2[128], delta[128], dropout row keys[128] (+pad to 1 KB)
  static constexpr int kQ = 2 * kTile;
  static constexpr int kDS = kQ + kQStages * kQStage;        // 2 x [128 keys x 128 queries] bf16 (two 64-query sub-tiles each);
  static constexpr int kDSBuf = kBT * kBT * 2;               //   also: K/V landing zone at start, dQ staging when released
  static constexpr int kDQStage = kDS + 2 * kDSBuf;            // dQ staging ring of the drain warpgroup: kDQBufs x [128 queries x 16 columns] f32
  static constexpr int kDQBufBytes = kBT * kDQChunk * 4;       // 8 KB
  static constexpr int kBar = kDQStage + kDQBufs * kDQBufBytes;
  static constexpr int kTotal = kBar + 256 + 1024;
};

// mbarriers; the ones marked [2] exist once per column half (warpgroup)
enum { BB_KV = 0, BB_KT = 1, BB_QF = 2, BB_QE = 5, BB_ST = 8 /*[2]*/, BB_STFREE = 10 /*[2]*/, BB_PT = 12 /*[2]*/, BB_DPT = 14 /*[2]*/,
       BB_DS = 16 /*[2]*/, BB_DQF = 18, BB_DQFREE = 19, BB_DONE = 20, BB_N = 21 };
// named barriers: 1 = the drain warpgroup, 3 + x = elementwise warpgroup x's turn on the exp unit
enum { NB_DRAIN = 1, NB_TURN = 3 };

// In-kernel timeline: SM-clock stamps of CTA (0,0) at the protocol points of iterations [kTraceI0, kTraceI0+8) for the two
// warpgroups (roles 0-1) and the MMA warp (role 2), written when tracing is switched on (hvc_debug_bwd_trace_enable;
// tests/bringup/bwd_trace.py prints them).  The stamps stay compiled in: they cost a not-taken branch per protocol point,
// and with them ptxas keeps the loop state in registers -- the same source without them spills ~120 bytes per thread in
// the elementwise loop and runs 13 % slower (same box, B200: 713 vs 818 TFLOP/s at d=64).
constexpr int kTraceI0 = 16, kTraceIters = 8, kTracePts = 12, kTraceRoles = 3;
__device__ unsigned long long g_bwd_trace[kTraceRoles * kTraceIters * kTracePts];
__device__ int g_bwd_trace_on = 0;
#define HVC_TR(role, i, pt)                                                                                     \
  do {                                                                                                          \
    if (trace_on && (i) >= kTraceI0 && (i) < kTraceI0 + kTraceIters)                                            \
      g_bwd_trace[((role) * kTraceIters + ((i) - kTraceI0)) * kTracePts + (pt)] = clock64();                    \
  } while (0)

// ---- elementwise phases of one (key tile, query tile) pair; thread == key row, 32 query columns per thread.
// FULL = no padding rows/columns in this pair (the masked variant is a separate code path: selects cost issue slots).
// Phase A: P^T = exp2(S^T*scale2 - lse2); the pre-pass stores -lse2 so the argument is a single FFMA2.
// DROP: attn_drop on P.  The keep decision of element (query, key) is regenerated from the query's row key (smem,
// written by the pre-pass) and this thread's key column; it is carried to phase B in the SIGN of the fp32 P value
// (P >= 0): dropped entries are stored negated, and the bf16 P^T that feeds dV is packed with relu (dropped -> 0).
template <bool FULL, bool DROP>
__device__ __forceinline__ void bwd_phase_a(const uint32_t (&sv)[kWgCols], uint32_t lse_saddr, float2 nss, bool key_ok, int q_valid,
                                            float2 (&pv)[kWgCols / 2], uint32_t (&ppk)[kWgCols / 2], uint32_t rk_saddr, uint32_t colkey,
                                            uint32_t thr) {
  // lse loads are issued in small batches ahead of their arithmetic: the asm statements keep program order, and a load
  // placed next to its first use exposes the shared-memory latency once per 4 columns
#pragma unroll
  for (int g = 0; g < kWgCols; g += 4 * kLsePre) {
    float4 lq[kLsePre];
#pragma unroll
    for (int u = 0; u < kLsePre; ++u) lq[u] = lds_f4(lse_saddr + (g + 4 * u) * 4);
#pragma unroll
    for (int u = 0; u < kLsePre; ++u) {
      const int t = g + 4 * u;
      const float4 l4 = lq[u];
      const float2 a = ffma2(make_float2(__uint_as_float(sv[t]), __uint_as_float(sv[t + 1])), nss, make_float2(l4.x, l4.y));
      const float2 c = ffma2(make_float2(__uint_as_float(sv[t + 2]), __uint_as_float(sv[t + 3])), nss, make_float2(l4.z, l4.w));
      float2 e0, e1;
#ifdef HVC_BWD_EXP_NO_MUFU      // timing experiment only (wrong results): the exp replaced by one FMUL2
      e0 = fmul2(a, a); e1 = fmul2(c, c);
#else
      if (((t >> 1) * kBwdEmu) % 16 < kBwdEmu) e0 = ex2_poly2(a);
      else e0 = make_float2(ex2_approx(a.x), ex2_approx(a.y));
      if ((((t >> 1) + 1) * kBwdEmu) % 16 < kBwdEmu) e1 = ex2_poly2(c);
      else e1 = make_float2(ex2_approx(c.x), ex2_approx(c.y));
#endif
      if (!FULL) {
        e0.x = (key_ok && t + 0 < q_valid) ? e0.x : 0.f;
        e0.y = (key_ok && t + 1 < q_valid) ? e0.y : 0.f;
        e1.x = (key_ok && t + 2 < q_valid) ? e1.x : 0.f;
        e1.y = (key_ok && t + 3 < q_valid) ? e1.y : 0.f;
      }
      if (DROP) {
        const uint4 rk = lds_u4(rk_saddr + t * 4);
        e0.x = drop_negate_if_dropped(e0.x, rk.x * colkey, thr);      // colkey = this thread's column multiplier: one multiply per element
        e0.y = drop_negate_if_dropped(e0.y, rk.y * colkey, thr);
        e1.x = drop_negate_if_dropped(e1.x, rk.z * colkey, thr);
        e1.y = drop_negate_if_dropped(e1.y, rk.w * colkey, thr);
      }
      pv[t >> 1] = e0;
      pv[(t >> 1) + 1] = e1;
      ppk[t >> 1] = DROP ? pack_bf16_relu(e0.x, e0.y) : pack_bf16(e0.x, e0.y);
      ppk[(t >> 1) + 1] = DROP ? pack_bf16_relu(e1.x, e1.y) : pack_bf16(e1.x, e1.y);
    }
  }
}
// x clamped to [0, 1] on the FMA pipe (add.sat): the kept probability P keep from the signed value s = +-P
__device__ __forceinline__ float sat01(float x) {
  float y;
  asm("add.sat.ftz.f32 %0, %1, 0f00000000;" : "=f"(y) : "f"(x));
  return y;
}
// Phase B: dS^T = P^T * (dP^T - delta) -> four 16-byte chunks (chunk0 ..) of this thread's row in a swizzled sub-tile
// DROP: dS = P * (keep ? dP / (1-p) : 0  -  delta), keep = sign of the stored P value.
template <bool FULL, bool DROP, bool TSDK>
__device__ __forceinline__ void bwd_phase_b(const uint32_t (&dv)[kWgCols], uint32_t delta_saddr, const float2 (&pv)[kWgCols / 2],
                                            bool key_ok, int q_valid, uint32_t sub_saddr, int r, int chunk0, float inv_keep, uint32_t t_ds) {
  const float2 neg1 = make_float2(-1.f, -1.f), neg2 = make_float2(-2.f, -2.f);
  const float2 rp2 = make_float2(inv_keep, inv_keep);
#pragma unroll
  for (int g = 0; g < kWgCols; g += 4 * kDeltaPre) {
    uint32_t dsp[2 * kDeltaPre];            // the group's dS^T as bf16 pairs: also written to TMEM (A operand of dK)
    float4 dq4[kDeltaPre];                  // delta of a column group, loaded ahead of the arithmetic (see phase A)
#pragma unroll
    for (int u = 0; u < kDeltaPre; ++u) dq4[u] = lds_f4(delta_saddr + (g + 4 * u) * 4);
#pragma unroll
    for (int kk = 0; kk < kDeltaPre / 2; ++kk) {
      const int k = (g >> 3) + kk;
      uint32_t w4[4];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int t = 8 * k + 4 * u;
        const float4 d4 = dq4[2 * kk + u];
        float2 x0, x1;
        if (DROP) {
          // s = +-P (sign = kept) and P <= 1, so P keep = sat(s) and   dS = P keep dP / (1-p) - P delta = sat(s) (dP / (1-p) - 2 delta) + s delta:
          // one saturating add per element plus packed multiplies and FMAs -- 3 issue slots per element.  (Round 1 formed |s| with operand
          // modifiers, which the packed instructions do not have: abs + add + multiply per element, 4 slots;
          // profiles/r02_attn_bwd_limiter_experiments.log.)
          const float2 s0 = pv[t >> 1], s1 = pv[(t >> 1) + 1];
          const float2 d0 = make_float2(d4.x, d4.y), d1 = make_float2(d4.z, d4.w);
          const float2 h0 = fmul2(make_float2(__uint_as_float(dv[t]), __uint_as_float(dv[t + 1])), rp2);       // rp2 = 1 / (1-p)
          const float2 h1 = fmul2(make_float2(__uint_as_float(dv[t + 2]), __uint_as_float(dv[t + 3])), rp2);
          const float2 g0 = ffma2(d0, neg2, h0), g1 = ffma2(d1, neg2, h1);
          const float2 k0 = make_float2(sat01(s0.x), sat01(s0.y)), k1 = make_float2(sat01(s1.x), sat01(s1.y));
          x0 = ffma2(k0, g0, fmul2(s0, d0));
          x1 = ffma2(k1, g1, fmul2(s1, d1));
        } else {
          x0 = ffma2(make_float2(d4.x, d4.y), neg1, make_float2(__uint_as_float(dv[t]), __uint_as_float(dv[t + 1])));
          x1 = ffma2(make_float2(d4.z, d4.w), neg1, make_float2(__uint_as_float(dv[t + 2]), __uint_as_float(dv[t + 3])));
          x0 = fmul2(x0, pv[t >> 1]);
          x1 = fmul2(x1, pv[(t >> 1) + 1]);
        }
        if (!FULL) {   // padded delta may be garbage: 0 * NaN must not leak
          x0.x = (key_ok && t + 0 < q_valid) ? x0.x : 0.f;
          x0.y = (key_ok && t + 1 < q_valid) ? x0.y : 0.f;
          x1.x = (key_ok && t + 2 < q_valid) ? x1.x : 0.f;
          x1.y = (key_ok && t + 3 < q_valid) ? x1.y : 0.f;
        }
        w4[2 * u] = pack_bf16(x0.x, x0.y);
        w4[2 * u + 1] = pack_bf16(x1.x, x1.y);
      }
      sts_u4(sub_saddr + sw128_offset(r, chunk0 + k), w4[0], w4[1], w4[2], w4[3]);
      dsp[4 * kk] = w4[0]; dsp[4 * kk + 1] = w4[1]; dsp[4 * kk + 2] = w4[2]; dsp[4 * kk + 3] = w4[3];
    }
    static_assert(kDeltaPre == 4, "one 8-column (16-query) TMEM store per group");
    if (TSDK) tmem_st_32x8(t_ds + (g >> 1), dsp);
  }
}

template <int HD, bool DROP>
__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                const __grid_constant__ CUtensorMap tmDQ, const AttnBwdKArgs p) {
  static_assert(HD == 64 || HD == 32, "head_dim 64 (SWIZZLE_128B tiles) or 32 (SWIZZLE_64B tiles)");
  using L = BwdSmem<HD>;
  using SW = Swz<HD * 2>;        // Q / K / V / dO tiles as TMA delivers them: rows of HD*2 bytes
  constexpr int kDCols = HD / kBwdWGs;     // d columns of dQ / dK / dV owned by one warpgroup
  // dK += dS^T Q with dS^T as a TMEM operand (bf16 pairs written over the dP^T columns by phase B) instead of the K-major smem tile:
  // 53 vs 76 clk per MMA step at d = 64 (+1.6 % on the kernel); at d = 32 the smem form is the faster one (65 clk, no extra TMEM store)
  constexpr bool kTsDk = HD == 64;
  // The warpgroups take turns on the exp unit (staggers them: one in phase A while the other is in phase B).  Without the turns, same box:
  // d = 64 dropout off 857 -> 817 TFLOP/s, with dropout 692 -> 691; d = 32 off 485 -> 473, with dropout 393 -> 407 -- so every instantiation
  // keeps them except <32, dropout> (profiles/r02_attn_bwd_limiter_experiments.log).
  constexpr bool kTurns = !(HD == 32 && DROP);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L::kBar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + BB_N);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);   // warp-uniform for the compiler (role dispatch, setmaxnreg)
  const int lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int b = bh / p.heads, h = bh - b * p.heads;
  const int j = blockIdx.x;           // key tile
  const int nQ = p.n_q_tiles;
  const bool trace_on = blockIdx.x == 0 && blockIdx.y == 0 && *reinterpret_cast<volatile int*>(&g_bwd_trace_on) != 0;
  constexpr int kDrainWarp0 = kBwdWGs * 4, kTmaWarp = kDrainWarp0 + 4, kMmaWarp = kTmaWarp + 1;
  constexpr uint32_t kEw = kBwdWGs * 4;     // elementwise warps: one mbarrier arrival per warp on the joint barriers

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO); tma_prefetch_desc(&tmDQ);
    mbar_init(&bar[BB_KV], 1);
    mbar_init(&bar[BB_KT], kEw);
    for (int s = 0; s < kQStages; ++s) { mbar_init(&bar[BB_QF + s], 1); mbar_init(&bar[BB_QE + s], 1); }
    for (int x = 0; x < 2; ++x) {
      mbar_init(&bar[BB_ST + x], 1);
      mbar_init(&bar[BB_STFREE + x], 4);
      mbar_init(&bar[BB_PT + x], 4);
      mbar_init(&bar[BB_DPT + x], 1);
      mbar_init(&bar[BB_DS + x], 4);
    }
    mbar_init(&bar[BB_DQF], 1);
    mbar_init(&bar[BB_DQFREE], 4);          // one arrival per drain warp
    mbar_init(&bar[BB_DONE], 1);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t kColSt = 0, kColPt = 128, kColDPt = 192, kColDV = 320, kColDK = 320 + HD, kColDQ = 320 + 2 * HD;
  constexpr uint32_t kSubKT = HD * 128;     // one 64-key sub-tile of K^T / V^T: HD rows of 128 bytes
  constexpr uint32_t kHalfRows = 64 * HD * 2;   // byte offset of row 64 inside a Q / dO tile (a whole number of swizzle atoms)

  // setmaxnreg sits INSIDE each role's branch: after a control-flow merge ptxas must assume the smallest budget of the merging paths
  if (warp == kTmaWarp) {
    // ===================== TMA producer =====================
    reg_dealloc<kRegsAux>();
    if (elect_one()) {
      mbar_arrive_expect_tx(&bar[BB_KV], 2 * L::kTile);     // K, V land in the (still unused) dS^T buffers
      tma_load_2d(smem + L::kDS, &tmK, &bar[BB_KV], h * HD, b * p.nk + j * kBT, kEvictFirst);
      tma_load_2d(smem + L::kDS + L::kTile, &tmV, &bar[BB_KV], h * HD, b * p.nk + j * kBT, kEvictFirst);
      for (int i = 0; i < nQ; ++i) {
        const int st = i % kQStages;
        const uint32_t ph = (i / kQStages) & 1;
        uint8_t* base = smem + L::kQ + st * L::kQStage;
        mbar_wait(&bar[BB_QE + st], ph ^ 1, 10);
        mbar_arrive_expect_tx(&bar[BB_QF + st], 2 * L::kTile + (DROP ? 3 : 2) * kBT * 4);
        tma_load_2d(base, &tmQ, &bar[BB_QF + st], h * HD, b * p.nq + i * kBT, kEvictLast);
        tma_load_2d(base + L::kTile, &tmDO, &bar[BB_QF + st], h * HD, b * p.nq + i * kBT, kEvictLast);
        bulk_load_1d(base + 2 * L::kTile, p.nlse2 + (long long)bh * p.nq_pad + i * kBT, kBT * 4, &bar[BB_QF + st]);
        bulk_load_1d(base + 2 * L::kTile + kBT * 4, p.delta + (long long)bh * p.nq_pad + i * kBT, kBT * 4, &bar[BB_QF + st]);
        if (DROP) bulk_load_1d(base + 2 * L::kTile + 2 * kBT * 4, p.rowkeys + (long long)bh * p.nq_pad + i * kBT, kBT * 4, &bar[BB_QF + st]);
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer (whole warp runs the loop; only the elected lane issues) =====================
    reg_dealloc<kRegsAux>();
    // Warpgroup 1 runs about half a tile behind warpgroup 0 (they alternate on the exp unit), and the issue order below
    // follows the order in which their results become available:
    //   dS_1(i-1) | P_0(i) | STFREE_0+1(i) | dS_0(i) | P_1(i)
    const bool leader = elect_one();
    constexpr uint32_t id_s = make_idesc_bf16(kBT, kWgCols, kMajorMN, kMajorK);  // dP^T_x: A = V^T (MN-major), B = 64 rows of dO
    constexpr uint32_t id_s2 = make_idesc_bf16(kBT, kBT, kMajorMN, kMajorK);     // S^T (both halves): A = K^T (MN-major), B = Q
    constexpr uint32_t id_dv = make_idesc_bf16(kBT, HD, kMajorK, kMajorMN);      // dV_x (A in TMEM), dK_x (A K-major smem), B MN-major
    constexpr uint32_t id_dq = make_idesc_bf16(kBT, HD, kMajorMN, kMajorK);      // dQ: A = dS^T read MN-major, B = K^T (K-major)
    const uint32_t sKT = smem_u32(smem + L::kKT), sVT = smem_u32(smem + L::kVT);
    const uint32_t sQ0 = smem_u32(smem + L::kQ), sDS0 = smem_u32(smem + L::kDS);
    auto commit = [&](int barrier) { if (leader) tc_commit(&bar[barrier]); };
    auto issue_st = [&](int st) {       // S^T = K Q^T for both column halves in one N=128 product (cheaper per column than two N=64)
      const uint32_t sQ = sQ0 + st * L::kQStage;
      if (leader) {
#pragma unroll
        for (int k16 = 0; k16 < HD / 16; ++k16)
          umma_ss(tmem_base + kColSt, make_sdesc_sw128(sKT + k16 * 2048, kSubKT, 1024), SW::desc(sQ + k16 * 32), id_s2, k16 > 0 ? 1u : 0u);
      }
      commit(BB_ST + 0);
      commit(BB_ST + 1);
    };
    auto issue_dpt = [&](int x, int st) {      // dP^T_x = V dO_x^T
      const uint32_t sDO = sQ0 + st * L::kQStage + L::kTile + x * kHalfRows;
      if (leader) {
#pragma unroll
        for (int k16 = 0; k16 < HD / 16; ++k16)
          umma_ss(tmem_base + kColDPt + x * kWgCols, make_sdesc_sw128(sVT + k16 * 2048, kSubKT, 1024), SW::desc(sDO + k16 * 32), id_s,
                  k16 > 0 ? 1u : 0u);
      }
      commit(BB_DPT + x);
    };
    auto issue_dv = [&](int x, int st, bool first) {   // dV += P^T_x dO_x   (A = P^T_x in TMEM, 8 columns per K=16 step; B = dO MN-major)
      const uint32_t sDO = sQ0 + st * L::kQStage + L::kTile;
#ifdef HVC_BWD_EXP_SKIP_DV
      if (false) {
#else
      if (leader) {
#endif
#pragma unroll
        for (int k16 = 0; k16 < kWgCols / 16; ++k16)
          umma_ts(tmem_base + kColDV, tmem_base + kColPt + x * (kWgCols / 2) + k16 * 8,
                  SW::desc(sDO + (x * (kWgCols / 16) + k16) * SW::kMnStep, 8192), id_dv, (!first || k16 > 0) ? 1u : 0u);
      }
    };
    auto issue_dk = [&](int x, int st, int buf, bool first) {   // dK += dS^T_x Q_x   (B = Q MN-major)
      const uint32_t sQ = sQ0 + st * L::kQStage;
      if (leader) {
#pragma unroll
        for (int k16 = 0; k16 < kWgCols / 16; ++k16) {
          const uint64_t bq = SW::desc(sQ + (x * (kWgCols / 16) + k16) * SW::kMnStep, 8192);
          if constexpr (kTsDk)    // A = dS^T_x as bf16 pairs in TMEM (over the dP^T_x columns)
            umma_ts(tmem_base + kColDK, tmem_base + kColDPt + x * kWgCols + k16 * 8, bq, id_dv, (!first || k16 > 0) ? 1u : 0u);
          else                    // A = sub-tile x of the dS^T smem buffer, K-major
            umma_ss(tmem_base + kColDK, make_sdesc_sw128(sDS0 + buf * L::kDSBuf + x * 16384 + k16 * 32, 16, 1024), bq, id_dv,
                    (!first || k16 > 0) ? 1u : 0u);
        }
      }
    };
    auto issue_dq = [&](int buf) {   // dQ = dS K   (A = both dS^T sub-tiles read MN-major: M = queries; B = K^T K-major: rows = d, two 64-key sub-tiles)
#ifdef HVC_BWD_EXP_SKIP_DQ
      if (false) {
#else
      if (leader) {
#endif
#pragma unroll
        for (int k16 = 0; k16 < kBT / 16; ++k16)
          umma_ss(tmem_base + kColDQ, make_sdesc_sw128(sDS0 + buf * L::kDSBuf + k16 * 2048, 16384, 1024),
                  make_sdesc_sw128(sKT + (k16 >> 2) * kSubKT + (k16 & 3) * 32, 16, 1024), id_dq, k16 > 0 ? 1u : 0u);
      }
      commit(BB_DQF);
    };
#ifndef HVC_BWD_DPT_SPLIT      // (the round-1 schedule with one dP^T product per column half is kept behind HVC_BWD_DPT_SPLIT for A/B runs)
    // dP^T(i) for both column halves in ONE N = 128 product (8 KB of operands per K = 16 step instead of 2 x 6 KB): issued in the tail of
    // tile i-1, after dK_0(i-1) and dK_1(i-1) have read the dS^T that shares its columns
    auto issue_dpt_full = [&](int st) {
      const uint32_t sDO = sQ0 + st * L::kQStage + L::kTile;
      if (leader) {
#pragma unroll
        for (int k16 = 0; k16 < HD / 16; ++k16)
          umma_ss(tmem_base + kColDPt, make_sdesc_sw128(sVT + k16 * 2048, kSubKT, 1024), SW::desc(sDO + k16 * 32), id_s2, k16 > 0 ? 1u : 0u);
      }
      commit(BB_DPT + 0);
      commit(BB_DPT + 1);
    };
    auto finish_tile_m = [&](int t) {
      mbar_wait(&bar[BB_DS + 1], t & 1, 25);
      tc_fence_after();
      issue_dk(1, t % kQStages, t & 1, false);
      if (t + 1 < nQ) issue_dpt_full((t + 1) % kQStages);
      if (t > 0) { mbar_wait(&bar[BB_DQFREE], (t - 1) & 1, 26); tc_fence_after(); }
      issue_dq(t & 1);
      commit(BB_QE + t % kQStages);
    };
#endif
    // tail of tile t (needs dS_1(t)): dP^T_1(t+1), dK_1(t), dQ(t); releases the Q/dO stage of tile t
    auto finish_tile = [&](int t) {
      mbar_wait(&bar[BB_DS + 1], t & 1, 25);
      tc_fence_after();
      if (leader) HVC_TR(2, t + 1, 2);
      if constexpr (kTsDk) {
        issue_dk(1, t % kQStages, t & 1, false);                 // reads dS^T_1(t) from the dP^T_1 columns ...
        if (t + 1 < nQ) issue_dpt(1, (t + 1) % kQStages);        // ... which dP^T_1(t+1) then overwrites (the tensor pipe runs in order)
      } else {
        if (t + 1 < nQ) issue_dpt(1, (t + 1) % kQStages);
        issue_dk(1, t % kQStages, t & 1, false);
      }
      if (t > 0) { mbar_wait(&bar[BB_DQFREE], (t - 1) & 1, 26); tc_fence_after(); }
      issue_dq(t & 1);
      commit(BB_QE + t % kQStages);
      if (leader) HVC_TR(2, t + 1, 3);
    };
    mbar_wait(&bar[BB_KT], 0, 20);
    mbar_wait(&bar[BB_QF + 0], 0, 21);
    tc_fence_after();
    issue_st(0);
#ifndef HVC_BWD_DPT_SPLIT
    issue_dpt_full(0);
    for (int i = 0; i < nQ; ++i) {
      const int st = i % kQStages;
      const int st1 = (i + 1) % kQStages;
      if (i > 0) finish_tile_m(i - 1);
      if (i + 1 < nQ) {                             // S^T(i+1) once both warpgroups have pulled S^T(i) out of TMEM
        mbar_wait(&bar[BB_STFREE + 0], i & 1, 22);
        mbar_wait(&bar[BB_STFREE + 1], i & 1, 27);
        mbar_wait(&bar[BB_QF + st1], ((i + 1) / kQStages) & 1, 23);
        tc_fence_after();
        issue_st(st1);
      }
      mbar_wait(&bar[BB_PT + 0], i & 1, 24);
      tc_fence_after();
      issue_dv(0, st, i == 0);
      mbar_wait(&bar[BB_DS + 0], i & 1, 28);
      tc_fence_after();
      issue_dk(0, st, i & 1, i == 0);
      mbar_wait(&bar[BB_PT + 1], i & 1, 29);
      tc_fence_after();
      issue_dv(1, st, false);
    }
    finish_tile_m(nQ - 1);
    commit(BB_DONE);
#else
    issue_dpt(0, 0);
    issue_dpt(1, 0);
    for (int i = 0; i < nQ; ++i) {
      const int st = i % kQStages;
      const int st1 = (i + 1) % kQStages;
      const bool more = i + 1 < nQ;
      if (leader) HVC_TR(2, i, 0);
      if (leader) HVC_TR(2, i, 1);
      if (i > 0) finish_tile(i - 1);
      mbar_wait(&bar[BB_PT + 0], i & 1, 24);
      tc_fence_after();
      if (leader) HVC_TR(2, i, 4);
      issue_dv(0, st, i == 0);
      if (more) {                                   // S^T(i+1) once both warpgroups have pulled S^T(i) out of TMEM
        mbar_wait(&bar[BB_STFREE + 0], i & 1, 22);
        mbar_wait(&bar[BB_STFREE + 1], i & 1, 27);
        mbar_wait(&bar[BB_QF + st1], ((i + 1) / kQStages) & 1, 23);
        tc_fence_after();
        issue_st(st1);
      }
      if (leader) HVC_TR(2, i, 5);
      mbar_wait(&bar[BB_DS + 0], i & 1, 28);        // dS_0(i) in smem, dP^T_0(i) consumed
      tc_fence_after();
      if (leader) HVC_TR(2, i, 6);
      if constexpr (kTsDk) {
        issue_dk(0, st, i & 1, i == 0);
        if (more) issue_dpt(0, st1);
      } else {
        if (more) issue_dpt(0, st1);
        issue_dk(0, st, i & 1, i == 0);
      }
      mbar_wait(&bar[BB_PT + 1], i & 1, 29);
      tc_fence_after();
      if (leader) HVC_TR(2, i, 7);
      issue_dv(1, st, false);
    }
    finish_tile(nQ - 1);
    commit(BB_DONE);        // every product issued: dV and dK are complete when this fires
#endif
  } else if (warp >= kDrainWarp0 && warp < kTmaWarp) {
    // ===================== dQ drain warpgroup: thread == query row of the tile =====================
    reg_dealloc<kRegsDrain>();
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    const bool lead = threadIdx.x == kDrainWarp0 * 32;
    const uint32_t stage = smem_u32(smem + L::kDQStage);
    for (int i = 0; i < nQ; ++i) {
      mbar_wait(&bar[BB_DQF], i & 1, 31);
      tc_fence_after();
      uint32_t v[HD];
      if constexpr (HD == 64) {
        uint32_t(&v0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[0]);
        uint32_t(&v1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[32]);
        tmem_ld_32x32(tmem_base + lane_base + kColDQ, v0);
        tmem_ld_32x32(tmem_base + lane_base + kColDQ + 32, v1);
      } else {
        tmem_ld_32x32(tmem_base + lane_base + kColDQ, v);
      }
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive_warp(&bar[BB_DQFREE]);              // the next dQ product may overwrite the accumulator
#if defined(HVC_BWD_EXP_DRAIN_LDONLY)
      asm volatile("" ::"r"(v[0]), "r"(v[HD - 1]));      // timing experiment: TMEM load only
#elif defined(HVC_BWD_DQ_RED)
      {   // variant: L2 reductions straight from registers (no staging tile, no TMA)
        float* dst = p.dq_accum + ((long long)bh * p.nq_pad + i * kBT + r) * HD;
#pragma unroll
        for (int k = 0; k < HD / 4; ++k)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * k), "f"(__uint_as_float(v[4 * k])), "f"(__uint_as_float(v[4 * k + 1])),
                       "f"(__uint_as_float(v[4 * k + 2])), "f"(__uint_as_float(v[4 * k + 3])) : "memory");
      }
#else
      // 16-column chunks through a ring of three 8 KB staging buffers: reduce(c) is issued once every thread's rows of chunk c are in
      // shared memory; the buffer that chunk c + 1 will use was read by reduce(c - 2), which the leader waits for BEFORE the barrier, so
      // one barrier per chunk covers both conditions and the TMA engine always has up to two reduces in flight
#pragma unroll
      for (int c = 0; c < HD / kDQChunk; ++c) {
        const uint32_t buf = static_cast<uint32_t>((i * (HD / kDQChunk) + c) % kDQBufs) * L::kDQBufBytes;
#pragma unroll
        for (int k = 0; k < kDQChunk / 4; ++k)
          sts_u4(stage + buf + swz_offset<kDQChunk * 4>(r, k), v[c * kDQChunk + 4 * k], v[c * kDQChunk + 4 * k + 1], v[c * kDQChunk + 4 * k + 2],
                 v[c * kDQChunk + 4 * k + 3]);
        fence_proxy_async_smem();
        if (lead) bulk_wait_read<1>();
        named_bar_sync(NB_DRAIN, 128);
#ifndef HVC_BWD_EXP_DRAIN_NOREDUCE
        if (lead) {
          tma_reduce_add_2d(&tmDQ, smem + L::kDQStage + buf, c * kDQChunk, bh * p.nq_pad + i * kBT);
          bulk_commit();
        }
#endif
      }
#endif
    }
    if (lead) bulk_wait<0>();
  } else if (warp < kDrainWarp0) {
    // ===================== elementwise warpgroups: thread == key row, warpgroup x == query columns [64x, 64x+64) =====================
    reg_alloc<kRegsEw>();
    const int wg = warp >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    const bool key_ok = (j * kBT + r) < p.nk;
    const bool keys_full = (j + 1) * kBT <= p.nk;
    const float scale2 = p.scale2;
    const bool wg_lead = (threadIdx.x & 127) == 0;
    const uint32_t sDS0 = smem_u32(smem + L::kDS);

    // ---- one-time: transpose K and V (landed K-major, rows = keys) into MN-major K^T / V^T (rows = d, 64 keys per 128-byte
    // row, SWIZZLE_128B).  Warpgroup x moves d columns [x*HD/2, (x+1)*HD/2) of this thread's key row.
    mbar_wait(&bar[BB_KV], 0, 30);
    {
      const uint32_t dst_row_base = (r >> 6) * kSubKT + (r & 7) * 2;
      const uint32_t key_chunk = (r & 63) >> 3;
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        const uint32_t land = sDS0 + which * L::kTile;
        const uint32_t dst = smem_u32(smem + (which == 0 ? L::kKT : L::kVT)) + dst_row_base;
#pragma unroll
        for (int c = 0; c < kDCols / 8; ++c) {
          const int chunk = wg * (kDCols / 8) + c;
          const uint4 u = lds_u4(land + SW::offset(r, chunk));
          const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const uint32_t d = chunk * 8 + e;
            sts_u16(dst + d * 128 + ((key_chunk ^ (d & 7)) << 4), static_cast<uint16_t>(w[e >> 1] >> (16 * (e & 1))));
          }
        }
      }
      fence_proxy_async_smem();
      mbar_arrive_warp(&bar[BB_KT]);
    }

    const float2 nss = make_float2(scale2, scale2);
    uint32_t colkey = 0, thr = 0;
    float inv_keep = 1.f;
    if (DROP) {
      colkey = drop_colmul(static_cast<uint32_t>(j * kBT + r));
      thr = p.drop.thr;
      inv_keep = p.drop.inv_keep;
    }
    if (kTurns && wg == 1) named_bar_arrive(NB_TURN + 0, 256);          // warpgroup 0 takes the first turn on the exp unit
    for (int i = 0; i < nQ; ++i) {
      const int st = i % kQStages;
      const uint32_t lse_saddr = smem_u32(smem + L::kQ + st * L::kQStage + 2 * L::kTile) + wg * kWgCols * 4;
      const uint32_t delta_saddr = lse_saddr + kBT * 4;
      const uint32_t rk_saddr = lse_saddr + 2 * kBT * 4;
      const int q_valid = p.nq - i * kBT - wg * kWgCols;      // this warpgroup's columns >= q_valid are padding
      const bool full = keys_full && q_valid >= kWgCols;
      float2 pv[kWgCols / 2];                                 // P^T row slice, fp32, lives across phase A -> B

      // ---------------- phase A: P^T = exp2(S^T * scale2 - lse2)
      if (wg_lead) HVC_TR(wg, i, 0);
      mbar_wait(&bar[BB_QF + st], (i / kQStages) & 1, 32);  // lse/delta of this query tile are in smem
      mbar_wait(&bar[BB_ST + wg], i & 1, 33);
      tc_fence_after();
      {
        uint32_t sv[kWgCols];
        uint32_t(&s0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sv[0]);
        uint32_t(&s1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sv[32]);
        tmem_ld_32x32(tmem_base + lane_base + kColSt + wg * kWgCols, s0);
        tmem_ld_32x32(tmem_base + lane_base + kColSt + wg * kWgCols + 32, s1);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive_warp(&bar[BB_STFREE + wg]);
        if (wg_lead) HVC_TR(wg, i, 1);
        if (kTurns) named_bar_sync(NB_TURN + wg, 256);                    // my turn on the exp unit
        if (wg_lead) HVC_TR(wg, i, 2);
        uint32_t ppk[kWgCols / 2];
        if (full) bwd_phase_a<true, DROP>(sv, lse_saddr, nss, key_ok, q_valid, pv, ppk, rk_saddr, colkey, thr);
        else      bwd_phase_a<false, DROP>(sv, lse_saddr, nss, key_ok, q_valid, pv, ppk, rk_saddr, colkey, thr);
        if (kTurns && !(wg == 1 && i == nQ - 1)) named_bar_arrive(NB_TURN + (wg ^ 1), 256);
        if (wg_lead) HVC_TR(wg, i, 3);
        // dP^T_x(i) was issued after dV_x(i-1): once it has completed, dV_x(i-1) has finished reading P^T_x(i-1) and the
        // P^T columns may be overwritten (the wait is long satisfied by now; it is also phase B's input)
        mbar_wait(&bar[BB_DPT + wg], i & 1, 34);
        tc_fence_after();
        tmem_st_32x32(tmem_base + lane_base + kColPt + wg * (kWgCols / 2), ppk);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive_warp(&bar[BB_PT + wg]);
      }

      // ---------------- phase B: dS^T_x = P^T * (dP^T - delta) -> sub-tile x of dS^T buffer (i & 1)
      if (wg_lead) HVC_TR(wg, i, 4);
      {
        uint32_t dv[kWgCols];
        uint32_t(&d0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&dv[0]);
        uint32_t(&d1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&dv[32]);
        tmem_ld_32x32(tmem_base + lane_base + kColDPt + wg * kWgCols, d0);
        tmem_ld_32x32(tmem_base + lane_base + kColDPt + wg * kWgCols + 32, d1);
        // sub-tile `wg` of dS^T buffer (i & 1) was last read by dK_wg(i-2) and dQ(i-2); the commit behind BB_ST(i) -- waited for in phase A
        // -- was issued after both, and the tensor pipe completes in order, so the sub-tile is free
        tmem_ld_wait();
        if (wg_lead) HVC_TR(wg, i, 5);
        const uint32_t sub_saddr = sDS0 + (i & 1) * L::kDSBuf + wg * 16384;
        const uint32_t t_ds = tmem_base + lane_base + kColDPt + wg * kWgCols;     // dS^T (bf16 pairs) over the dP^T columns just read
        if (full) bwd_phase_b<true, DROP, kTsDk>(dv, delta_saddr, pv, key_ok, q_valid, sub_saddr, r, 0, inv_keep, t_ds);
        else      bwd_phase_b<false, DROP, kTsDk>(dv, delta_saddr, pv, key_ok, q_valid, sub_saddr, r, 0, inv_keep, t_ds);
      }
      if (wg_lead) HVC_TR(wg, i, 6);
      if (kTsDk) tmem_st_wait();
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive_warp(&bar[BB_DS + wg]);
      if (wg_lead) HVC_TR(wg, i, 7);

      if (wg_lead) HVC_TR(wg, i, 9);
    }
    mbar_wait(&bar[BB_DONE], 0, 35);      // every product has completed: dV and dK are final
    tc_fence_after();

    // ---- epilogue: dV, dK (x softmax scale) -> bf16 -> global; warpgroup x writes d columns [x*HD/2, (x+1)*HD/2)
    const int key = j * kBT + r;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      bf16* dst = (which == 0 ? p.dv + (long long)(b * p.nk + key) * p.lddv : p.dk + (long long)(b * p.nk + key) * p.lddk) + h * HD + wg * kDCols;
      const float mul = which == 0 ? inv_keep : p.scale;     // dV = (drop(P))^T dO carries the 1/(1-p) of the kept entries
      uint32_t v[kDCols];
      if constexpr (HD == 64) tmem_ld_32x32(tmem_base + lane_base + (which == 0 ? kColDV : kColDK) + wg * kDCols, v);
      else                    tmem_ld_32x16(tmem_base + lane_base + (which == 0 ? kColDV : kColDK) + wg * kDCols, v);
      tmem_ld_wait();
      if (key_ok) {
#pragma unroll
        for (int t = 0; t < kDCols; t += 8) {
          uint4 u;
          u.x = pack_bf16(__uint_as_float(v[t]) * mul, __uint_as_float(v[t + 1]) * mul);
          u.y = pack_bf16(__uint_as_float(v[t + 2]) * mul, __uint_as_float(v[t + 3]) * mul);
          u.z = pack_bf16(__uint_as_float(v[t + 4]) * mul, __uint_as_float(v[t + 5]) * mul);
          u.w = pack_bf16(__uint_as_float(v[t + 6]) * mul, __uint_as_float(v[t + 7]) * mul);
          *reinterpret_cast<uint4*>(dst + t) = u;
        }
      }
    }
  } else {
    reg_dealloc<kRegsAux>();      // the two idle warps of the TMA / MMA warpgroup (setmaxnreg is a warpgroup-wide instruction)
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// delta[b,h,q] = sum_d dO[b,q,h,d] * O[b,q,h,d]: HD/8 lanes per (token, head), 16 bytes of O and dO each (a token's heads
// are contiguous, so a warp reads whole 128-byte lines), shuffle-reduced; also writes -lse2 and the dropout row keys.
template <int HD>
__global__ void __launch_bounds__(256) attn_delta_kernel(const bf16* __restrict__ o, long long ldo, const bf16* __restrict__ d_o,
                                                         long long lddo, const float* __restrict__ lse2, float* __restrict__ delta,
                                                         float* __restrict__ nlse2, int batch, int heads, int nq, int nq_pad,
                                                         uint32_t* __restrict__ rowkeys, const DropArg drop) {
  constexpr int kLanes = HD / 8;                          // lanes per (token, head): 8 or 4
  const long long gid = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long item = gid / kLanes;                    // (token, head) index, head fastest
  const int part = (int)(gid - item * kLanes);
  const long long total = (long long)batch * nq * heads;
  const bool live = item < total;
  const long long it = live ? item : total - 1;           // keep the whole warp in the shuffles
  const int h = (int)(it % heads);
  const long long tok = it / heads;
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(o + tok * ldo + h * HD + part * 8));
  const uint4 g = __ldg(reinterpret_cast<const uint4*>(d_o + tok * lddo + h * HD + part * 8));
  const float2 a0 = unpack_bf16(a.x), a1 = unpack_bf16(a.y), a2 = unpack_bf16(a.z), a3 = unpack_bf16(a.w);
  const float2 g0 = unpack_bf16(g.x), g1 = unpack_bf16(g.y), g2 = unpack_bf16(g.z), g3 = unpack_bf16(g.w);
  float sum = a0.x * g0.x + a0.y * g0.y + a1.x * g1.x + a1.y * g1.y + a2.x * g2.x + a2.y * g2.y + a3.x * g3.x + a3.y * g3.y;
#pragma unroll
  for (int off = kLanes / 2; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
  if (live && part == 0) {
    const int b = (int)(tok / nq), q = (int)(tok - (long long)b * nq);
    const long long idx = ((long long)b * heads + h) * nq_pad + q;
    delta[idx] = sum;
    nlse2[idx] = -lse2[idx];
    if (drop.seed != nullptr) rowkeys[idx] = drop_rowkey(drop_load(drop), static_cast<uint32_t>((b * heads + h) * nq + q));
  }
}

// dq_accum f32 [B,H,nq_pad,HD] -> dq bf16 [B*nq, lddq] (column h*HD + d); 8 elements per thread
template <int HD>
__global__ void __launch_bounds__(256) attn_dq_convert_kernel(const float* __restrict__ acc, bf16* __restrict__ dq, long long lddq,
                                                              int batch, int heads, int nq, int nq_pad, float scale) {
  constexpr int kParts = HD / 8;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;   // over B*nq*H*kParts
  const long long total = (long long)batch * nq * heads * kParts;
  if (idx >= total) return;
  const int part = (int)(idx % kParts);
  const long long t = idx / kParts;
  const int h = (int)(t % heads);
  const long long tok = t / heads;
  const int b = (int)(tok / nq), q = (int)(tok - (long long)b * nq);
  const float* src = acc + (((long long)b * heads + h) * nq_pad + q) * HD + part * 8;
  const float4 x = __ldg(reinterpret_cast<const float4*>(src)), y = __ldg(reinterpret_cast<const float4*>(src + 4));
  uint4 u = make_uint4(pack_bf16(x.x * scale, x.y * scale), pack_bf16(x.z * scale, x.w * scale), pack_bf16(y.x * scale, y.y * scale),
                       pack_bf16(y.z * scale, y.w * scale));
  *reinterpret_cast<uint4*>(dq + tok * lddq + h * HD + part * 8) = u;
}

}  // namespace hvc

namespace hvc {
template <int HD, bool DROP>
static int launch_attn_bwd(const hvc_attn_args* a, cudaStream_t st) {
  using L = BwdSmem<HD>;
  const int nq_pad = (a->nq + 127) / 128 * 128;
  const uint64_t width = (uint64_t)a->heads * HD;
  const int swz = HD == 64 ? 1 : 2;
  const long long plane = (long long)a->batch * a->heads * nq_pad;
  const DropArg drop = make_drop(a->drop);
  {
    const long long threads = (long long)a->batch * a->nq * a->heads * (HD / 8);
    attn_delta_kernel<HD><<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(reinterpret_cast<const bf16*>(a->o), a->ldo,
                                                                      reinterpret_cast<const bf16*>(a->d_o), a->lddo, a->lse, a->delta,
                                                                      a->delta + plane, a->batch, a->heads, a->nq, nq_pad,
                                                                      reinterpret_cast<uint32_t*>(a->delta + 2 * plane), drop);
    HVC_LAUNCH_CHECK();
  }
  CUtensorMap tmQ, tmK, tmV, tmDO, tmDQ;
  int rc;
  if ((rc = make_tmap_2d(&tmDQ, a->dq_accum, 4, (uint64_t)a->batch * a->heads * nq_pad, HD, HD, kDQChunk, kBT, 2))) return rc;   // 64-byte rows: SWIZZLE_64B
  if ((rc = make_tmap_2d(&tmQ, a->q, 2, (uint64_t)a->batch * a->nq, width, a->ldq, HD, kBT, swz))) return rc;
  if ((rc = make_tmap_2d(&tmDO, a->d_o, 2, (uint64_t)a->batch * a->nq, width, a->lddo, HD, kBT, swz))) return rc;
  if ((rc = make_tmap_2d(&tmK, a->k, 2, (uint64_t)a->batch * a->nk, width, a->ldk, HD, kBT, swz))) return rc;
  if ((rc = make_tmap_2d(&tmV, a->v, 2, (uint64_t)a->batch * a->nk, width, a->ldv, HD, kBT, swz))) return rc;
  AttnBwdKArgs ka;
  ka.batch = a->batch; ka.heads = a->heads; ka.nq = a->nq; ka.nk = a->nk; ka.nq_pad = nq_pad;
  ka.n_q_tiles = nq_pad / kBT;
  ka.nlse2 = a->delta + plane; ka.delta = a->delta; ka.dq_accum = a->dq_accum;
  ka.rowkeys = reinterpret_cast<const uint32_t*>(a->delta + 2 * plane); ka.drop = drop;
  ka.dk = reinterpret_cast<bf16*>(a->dk); ka.lddk = a->lddk;
  ka.dv = reinterpret_cast<bf16*>(a->dv); ka.lddv = a->lddv;
  ka.scale = a->scale; ka.scale2 = a->scale * 1.4426950408889634f;
  HVC_SMEM_OPT_IN((attn_bwd_kernel<HD, DROP>), L::kTotal);
  dim3 grid((a->nk + kBT - 1) / kBT, a->batch * a->heads);
  attn_bwd_kernel<HD, DROP><<<grid, kBwdThreads, L::kTotal, st>>>(tmQ, tmK, tmV, tmDO, tmDQ, ka);
  HVC_LAUNCH_CHECK();
  {
    const long long threads = (long long)a->batch * a->nq * a->heads * (HD / 8);
    attn_dq_convert_kernel<HD><<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(a->dq_accum, reinterpret_cast<bf16*>(a->dq), a->lddq,
                                                                                 a->batch, a->heads, a->nq, nq_pad, a->scale);
    HVC_LAUNCH_CHECK();
  }
  return HVC_OK;
}
}  // namespace hvc

// bring-up hooks (not part of include/hvc.h): switch the in-kernel timeline of attn_bwd_kernel on/off, read it back
extern "C" int hvc_debug_bwd_trace_enable(int on) {
  return (int)cudaMemcpyToSymbol(hvc::g_bwd_trace_on, &on, sizeof(int));
}
extern "C" int hvc_debug_bwd_trace(unsigned long long* dst) {
  return (int)cudaMemcpyFromSymbol(dst, hvc::g_bwd_trace, sizeof(hvc::g_bwd_trace));
}

extern "C" int hvc_attn_bwd(const hvc_attn_args* a, void* stream) {
  using namespace hvc;
  HVC_CHECK_ARG(a != nullptr && a->size == sizeof(hvc_attn_args), "hvc_attn_bwd: bad args struct");
  HVC_CHECK_ARG(a->batch > 0 && a->heads > 0 && a->nq > 0 && a->nk > 0, "hvc_attn_bwd: empty problem");
  HVC_CHECK_ARG(a->head_dim == 64 || a->head_dim == 32, "hvc_attn_bwd: head_dim %d not supported (32 or 64)", a->head_dim);
  HVC_CHECK_ARG(a->q && a->k && a->v && a->o && a->d_o && a->lse && a->delta && a->dq_accum && a->dq && a->dk && a->dv,
                "hvc_attn_bwd: null operand");
  HVC_CHECK_ARG(((a->lddq | a->lddk | a->lddv) & 7) == 0, "hvc_attn_bwd: gradient row pitches must be multiples of 8");
  HVC_CHECK_ARG(((a->ldo | a->lddo) & 7) == 0 && ((reinterpret_cast<uintptr_t>(a->o) | reinterpret_cast<uintptr_t>(a->d_o)) & 15) == 0,
                "hvc_attn_bwd: o and d_o must have 16-byte aligned rows");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool drop = a->drop.seed != nullptr && a->drop.p > 0.f;
  HVC_CHECK_ARG(!drop || a->drop.p < 1.f, "hvc_attn_bwd: dropout p must be < 1");
  return a->head_dim == 64 ? (drop ? launch_attn_bwd<64, true>(a, st) : launch_attn_bwd<64, false>(a, st))
                           : (drop ? launch_attn_bwd<32, true>(a, st) : launch_attn_bwd<32, false>(a, st));
}
