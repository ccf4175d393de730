// hvc_attn_fwd.cu -- fused flash-style attention forward for sm_100a (self- and cross-attention).
//
// Replaces  attn = softmax(q k^T * d^-1/2);  out = attn v   (vit_components.py:46-51, :103-113)
// without ever writing the (B,h,N,M) score matrix: S and P live in TMEM, O accumulates in TMEM.
//
// One CTA = one (batch, head) x 256 queries (two 128-row tiles A and B), 320 threads:
//   warps 0-3   softmax warpgroup A : thread == query row == TMEM lane; online softmax in registers,
//   warps 4-7   softmax warpgroup B   no cross-thread reductions at all; writes P (bf16) back over S
//   warp  8     TMA producer        : Q (once), then K/V tiles through two 3-stage rings
//   warp  9     MMA issuer          : S = Q K^T (SS, 128x128xd), O += P V (TS: P from TMEM, V MN-major)
// While one warpgroup runs exp2 on tile j the tensor core computes S for the other one, so the MUFU
// pipe (the real bound at d<=64: 16 ex2/clk/SM vs 4*d flop per score) stays busy.
// Part of the exp2 work (kEmu pairs of every 16) runs as a degree-3 polynomial on the FMA pipe (ex2_poly2): +6 % at d=64,
// +8 % at d=32 (profiles/r01_attn_fwd_exp2_poly_sweep.log).  Tried and rejected on B200, both slower: no turn-taking
// (633 vs 765 TFLOP/s) and two warpgroups per query tile splitting the key columns (760 vs 812): per tile the chain
// exp -> P V -> S(next) -> row max leaves ~1300 clk outside the exp section, so shorter turns do not shorten the period.
// Timeline of this version (profiles/r01_attn_fwd_timeline_d64_v2.log): period ~2830 clk per key tile pair = exp turn (~1260)
// + the chain P -> P V -> S(next) -> row max (~1590).  Tried on top of it, all within +-2 % of this version in the training step:
// Q^T as an MN-major A operand + P in its own TMEM columns + S(j+1) issued before P(j) V (the chain shrinks to ~1000 clk but the
// row-max pass, now overlapping the other tile's MMAs, grows from ~420 to ~690 clk: TMEM reads slow down while TS-form MMAs run);
// the whole S row held in registers (one TMEM read per tile; 752 vs 800 TFLOP/s).  The bound is each warpgroup's own serial work.
// O is rescaled lazily: only when a row maximum grows by more than 2^8 (then the owning warp fixes O
// in TMEM); otherwise stale maxima are carried and cancel in the final 1/l normalisation.
//
// Q, K, V are read in place from the packed projection outputs (row = token, column = head*d + j),
// O is written token-major [T, h*d] ready for the output projection -- no head split/merge copies.
#include "hvc_common.cuh"
#include "hvc_host.h"

namespace hvc {

constexpr int kFwdThreads = 320;
constexpr int kQTile = 128;
constexpr int kKTile = 128;
constexpr int kKvStages = 3;
constexpr int kEmu64 = 6, kEmu32 = 6;   // element pairs (of 16) whose exp2 runs as a polynomial on the FMA pipe
#ifndef HVC_FWD_EMU_DROP
#define HVC_FWD_EMU_DROP 0
#endif
// the dropout instantiations keep every exp2 on the MUFU at d = 64: their mask work already fills the FMA pipe -- round 1 (xor hash), in the
// 128^3 training step: 6 -> 531, 3 -> 555, 0 -> 594 TFLOP/s; round 2 (multiplicative hash), kernel level: 2 -> 653, 4 -> 646, 6 -> 587, 0 -> 670.
// At d = 32 (half the tensor work per exp) two pairs of 16 help a little: 335 -> 344 TFLOP/s (profiles/r02_attn_bwd_nodrain_bound.log).
constexpr int kEmuDrop = HVC_FWD_EMU_DROP, kEmuDrop32 = HVC_FWD_EMU_DROP ? HVC_FWD_EMU_DROP : 2;

struct AttnFwdKArgs {
  int batch, heads, nq, nk, n_kv_tiles;
  bf16* o; long long ldo;
  float* lse2;      // [B, H, nq_pad]  log2-domain logsumexp of the scaled scores
  int nq_pad;
  float scale2;     // softmax scale * log2(e)
  DropArg drop;     // attn_drop on P (vit_components.py:49,110); used by the DROP instantiations only
};

template <int HD>
struct FwdSmem {
  static constexpr int kTile = kQTile * HD * 2;       // bytes of one 128 x HD bf16 tile
  static constexpr int kQ = 0;
  static constexpr int kK = 2 * kTile;
  static constexpr int kV = kK + kKvStages * kTile;
  static constexpr int kBar = kV + kKvStages * kTile;
  static constexpr int kTotal = kBar + 256 + 1024;
};

// In-kernel timeline (bring-up): SM-clock stamps of CTA (0,0) for iterations [kFwdTraceI0, +8) of softmax warpgroup A (role 0),
// B (role 1) and the MMA warp (role 2); switched on with hvc_debug_fwd_trace_enable, read with hvc_debug_fwd_trace
// (tests/bringup/fwd_trace.py).  Compiled only with -DHVC_TRACE_FWD (HVC_EXTRA_NVCC_FLAGS): unlike in the backward kernel the
// stamps are not free here -- left in, the forward runs ~5 % slower in the training step (729 vs 770 TFLOP/s).
#ifdef HVC_TRACE_FWD
constexpr int kFwdTraceI0 = 16, kFwdTraceIters = 8, kFwdTracePts = 8;
__device__ unsigned long long g_fwd_trace[3 * kFwdTraceIters * kFwdTracePts];
__device__ int g_fwd_trace_on = 0;
#define HVC_FTR(role, j, pt)                                                                                    \
  do {                                                                                                          \
    if (trace_on && (j) >= kFwdTraceI0 && (j) < kFwdTraceI0 + kFwdTraceIters)                                   \
      g_fwd_trace[((role) * kFwdTraceIters + ((j) - kFwdTraceI0)) * kFwdTracePts + (pt)] = clock64();           \
  } while (0)
#define HVC_FTR_ON(expr) (expr)
#else
#define HVC_FTR(role, j, pt) do {} while (0)
#define HVC_FTR_ON(expr) false
#endif

enum { BAR_Q = 0, BAR_KF = 1, BAR_KE = 4, BAR_VF = 7, BAR_VE = 10, BAR_SF = 13, BAR_PF = 15, BAR_OD = 17, BAR_N = 19 };

// ---- softmax passes over one 128-column S tile held in TMEM (thread == row).  MASKED = last, partial key tile.
template <bool MASKED>
__device__ __forceinline__ float softmax_rowmax(uint32_t tS, int tail) {
  float a0 = -INFINITY, a1 = -INFINITY;
  uint32_t u0[32], u1[32];
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    tmem_ld_32x32(tS + half * 64, u0);
    tmem_ld_32x32(tS + half * 64 + 32, u1);
    tmem_ld_wait();
    if (MASKED) {
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        a0 = fmaxf(a0, half * 64 + c < tail ? __uint_as_float(u0[c]) : -INFINITY);
        a1 = fmaxf(a1, half * 64 + 32 + c < tail ? __uint_as_float(u1[c]) : -INFINITY);
      }
    } else {
#pragma unroll
      for (int c = 0; c < 32; c += 2) {
        a0 = fmax3(a0, __uint_as_float(u0[c]), __uint_as_float(u0[c + 1]));
        a1 = fmax3(a1, __uint_as_float(u1[c]), __uint_as_float(u1[c + 1]));
      }
    }
  }
  return fmaxf(a0, a1);
}

// rowkey/col0/thr: dropout on P (DROP): the row sum uses the undropped probabilities (softmax normalisation comes
// before nn.Dropout in the reference), the P that feeds P V has the dropped entries zeroed; 1/(1-p) is applied to O.
// EMU of every 16 element pairs take the polynomial path (spread evenly so MUFU and FMA work interleave).
template <bool MASKED, bool DROP, int EMU>
__device__ __forceinline__ void softmax_exp_chunk(const uint32_t (&v)[32], uint32_t tP, int c0, int tail, float2 scale2v, float2 neg_m,
                                                  float2& sum, uint32_t rowkey, uint32_t col0, uint32_t thr) {
  uint32_t pk[16];
  // the mask hashes do not depend on the scores, and ptxas hoisted a whole tile of them ahead of the TMEM loads, parking ~40 in local
  // memory (160 B of spills per thread): the empty asm makes this chunk's block key formally depend on its first score, so the hashes
  // are computed chunk by chunk (24 B of spills; forward 594 -> 620 TFLOP/s in the train-mode step, same box)
  if (DROP) asm volatile("" : "+r"(rowkey) : "r"(v[0]));
#pragma unroll
  for (int c = 0; c < 32; c += 2) {
    const float2 a = ffma2(make_float2(__uint_as_float(v[c]), __uint_as_float(v[c + 1])), scale2v, neg_m);
    float2 e;
#ifdef HVC_FWD_EXP_NO_EXP          // timing experiment only (wrong results): the exp replaced by one FMUL2
    e = fmul2(a, a);
#else
    if (((c >> 1) * EMU) % 16 < EMU) e = ex2_poly2(a);
    else e = make_float2(ex2_approx(a.x), ex2_approx(a.y));
#endif
    if (MASKED) {
      e.x = (c0 + c < tail) ? e.x : 0.f;
      e.y = (c0 + c + 1 < tail) ? e.y : 0.f;
    }
    sum = fadd2(sum, e);
    if (DROP) {   // rowkey here is the block key of this 128-key tile (drop_blockkey); c0 + c is the column inside it (a constant multiplier)
      e.x = drop_negate_if_dropped(e.x, drop_hash_in_block(rowkey, c0 + c), thr);
      e.y = drop_negate_if_dropped(e.y, drop_hash_in_block(rowkey, c0 + c + 1), thr);
    }
    pk[c >> 1] = DROP ? pack_bf16_relu(e.x, e.y) : pack_bf16(e.x, e.y);      // relu: the negated (dropped) entries become 0
  }
  tmem_st_32x16(tP + (c0 >> 1), pk);   // P (bf16 pairs) over S columns that were already consumed
}

// P = exp2(S*scale2 - m): chunk c+1 is fetched from TMEM while chunk c is exponentiated.  The section between the
// named-barrier sync and arrive is the warpgroup's turn on the MUFU pipe.
template <bool MASKED, bool DROP, int EMU>
__device__ __forceinline__ float softmax_exp(uint32_t tS, int tail, float scale2, float m, int turn_bar, int next_bar, bool hand_over,
                                             uint32_t rowkey, uint32_t col0, uint32_t thr, bool trace_on, int role, int j) {
  const float2 scale2v = make_float2(scale2, scale2), neg_m = make_float2(-m, -m);
  float2 sum = make_float2(0.f, 0.f);
  if (DROP) rowkey = drop_blockkey(rowkey, col0);   // col0 is a multiple of the 128-key tile
  uint32_t bufa[32], bufb[32];
  tmem_ld_32x32(tS, bufa);
  named_bar_sync(turn_bar, 256);
  HVC_FTR(role, j, 3);
  tmem_ld_wait();
  tmem_ld_32x32(tS + 32, bufb);
  softmax_exp_chunk<MASKED, DROP, EMU>(bufa, tS, 0, tail, scale2v, neg_m, sum, rowkey, col0, thr);
  tmem_ld_wait();
  tmem_ld_32x32(tS + 64, bufa);
  softmax_exp_chunk<MASKED, DROP, EMU>(bufb, tS, 32, tail, scale2v, neg_m, sum, rowkey, col0, thr);
  tmem_ld_wait();
  tmem_ld_32x32(tS + 96, bufb);
  softmax_exp_chunk<MASKED, DROP, EMU>(bufa, tS, 64, tail, scale2v, neg_m, sum, rowkey, col0, thr);
  tmem_ld_wait();
  softmax_exp_chunk<MASKED, DROP, EMU>(bufb, tS, 96, tail, scale2v, neg_m, sum, rowkey, col0, thr);
  if (hand_over) named_bar_arrive(next_bar, 256);
  HVC_FTR(role, j, 4);
  return sum.x + sum.y;
}

template <int HD, bool DROP, int EMU>
__global__ void __launch_bounds__(kFwdThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnFwdKArgs p) {
  static_assert(HD == 64 || HD == 32, "head_dim 64 (128-byte rows, SWIZZLE_128B) or 32 (64-byte rows, SWIZZLE_64B)");
  using L = FwdSmem<HD>;
  using SW = Swz<HD * 2>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L::kBar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + BAR_N);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int b = bh / p.heads, h = bh - b * p.heads;
  const int q0 = blockIdx.x * (2 * kQTile);
  const int n_tiles = p.n_kv_tiles;
  const bool trace_cta = HVC_FTR_ON(blockIdx.x == 0 && blockIdx.y == 0 && *reinterpret_cast<volatile int*>(&g_fwd_trace_on) != 0);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(&bar[BAR_Q], 1);
    for (int s = 0; s < kKvStages; ++s) {
      mbar_init(&bar[BAR_KF + s], 1);
      mbar_init(&bar[BAR_KE + s], 1);
      mbar_init(&bar[BAR_VF + s], 1);
      mbar_init(&bar[BAR_VE + s], 1);
    }
    for (int x = 0; x < 2; ++x) {
      mbar_init(&bar[BAR_SF + x], 1);
      mbar_init(&bar[BAR_PF + x], 4);      // one arrival per softmax warp
      mbar_init(&bar[BAR_OD + x], 1);
    }
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t kColS = 0, kColO = 256;

  if (warp == 8) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(&bar[BAR_Q], 2 * L::kTile);
      tma_load_2d(smem + L::kQ, &tmQ, &bar[BAR_Q], h * HD, b * p.nq + q0, kEvictFirst);
      tma_load_2d(smem + L::kQ + L::kTile, &tmQ, &bar[BAR_Q], h * HD, b * p.nq + q0 + kQTile, kEvictFirst);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % kKvStages;
        const uint32_t ph = (j / kKvStages) & 1;
        mbar_wait(&bar[BAR_KE + st], ph ^ 1, 10);
        mbar_arrive_expect_tx(&bar[BAR_KF + st], L::kTile);
        tma_load_2d(smem + L::kK + st * L::kTile, &tmK, &bar[BAR_KF + st], h * HD, b * p.nk + j * kKTile, kEvictLast);
        mbar_wait(&bar[BAR_VE + st], ph ^ 1, 11);
        mbar_arrive_expect_tx(&bar[BAR_VF + st], L::kTile);
        tma_load_2d(smem + L::kV + st * L::kTile, &tmV, &bar[BAR_VF + st], h * HD, b * p.nk + j * kKTile, kEvictLast);
      }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer (whole warp runs the loop so address math stays on the uniform datapath;
    // only the elected lane issues tcgen05.mma / commit) =====================
    const bool leader = elect_one();
    constexpr uint32_t idesc_s = make_idesc_bf16(kQTile, kKTile, kMajorK, kMajorK);
    constexpr uint32_t idesc_o = make_idesc_bf16(kQTile, HD, kMajorK, kMajorMN);
    const uint32_t sQ = smem_u32(smem + L::kQ), sK = smem_u32(smem + L::kK), sV = smem_u32(smem + L::kV);
    auto issue_s = [&](int x, int st) {   // S_x = Q_x K^T
#ifdef HVC_FWD_EXP_SKIP_S
      if (false) {
#else
      if (leader) {
#endif
#pragma unroll
        for (int k16 = 0; k16 < HD / 16; ++k16)
          umma_ss(tmem_base + kColS + x * kKTile, SW::desc(sQ + x * L::kTile + k16 * 32),
                  SW::desc(sK + st * L::kTile + k16 * 32), idesc_s, k16 > 0 ? 1u : 0u);
      }
    };
    auto issue_pv = [&](int x, int st, bool acc) {   // O_x (+)= P_x V   (P: TMEM, 8 columns per K=16 step)
#ifdef HVC_FWD_EXP_SKIP_PV
      if (false) {
#else
      if (leader) {
#endif
#pragma unroll
        for (int k16 = 0; k16 < kKTile / 16; ++k16)
          umma_ts(tmem_base + kColO + x * HD, tmem_base + kColS + x * kKTile + k16 * 8,
                  SW::desc(sV + st * L::kTile + k16 * SW::kMnStep, 8192), idesc_o, (acc || k16 > 0) ? 1u : 0u);
      }
    };
    auto commit = [&](int barrier) { if (leader) tc_commit(&bar[barrier]); };
    mbar_wait(&bar[BAR_Q], 0, 20);
    mbar_wait(&bar[BAR_KF + 0], 0, 21);
    tc_fence_after();
    issue_s(0, 0); commit(BAR_SF + 0);
    issue_s(1, 0); commit(BAR_SF + 1);
    commit(BAR_KE + 0);
    for (int j = 0; j < n_tiles; ++j) {
      const int st = j % kKvStages;
      const uint32_t ph = (j / kKvStages) & 1;
      const int st1 = (j + 1) % kKvStages;
      const uint32_t ph1 = ((j + 1) / kKvStages) & 1;
      const bool more = (j + 1 < n_tiles);
      const bool trace_on = trace_cta && leader;
      HVC_FTR(2, j, 0);
      mbar_wait(&bar[BAR_VF + st], ph, 22);
      HVC_FTR(2, j, 1);
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        mbar_wait(&bar[BAR_PF + x], j & 1, 23);
        tc_fence_after();
        HVC_FTR(2, j, 2 + 2 * x);
        issue_pv(x, st, j > 0);
        commit(BAR_OD + x);
        if (x == 1) commit(BAR_VE + st);
        if (more) {
          if (x == 0) { mbar_wait(&bar[BAR_KF + st1], ph1, 24); tc_fence_after(); }
          issue_s(x, st1);
          commit(BAR_SF + x);
          if (x == 1) commit(BAR_KE + st1);
        }
        HVC_FTR(2, j, 3 + 2 * x);
      }
    }
  } else {
    // ===================== softmax warpgroups =====================
    const int x = warp >> 2;                 // 0 = tile A, 1 = tile B
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;     // row inside the 128-query tile
    const uint32_t tS = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + kColS + x * kKTile;
    const uint32_t tO = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + kColO + x * HD;
    const float scale2 = p.scale2;
    float m = -INFINITY, l = 0.f;
    const int tail = p.nk - (n_tiles - 1) * kKTile;   // valid keys in the last tile (1..128)
    uint32_t rowkey = 0, thr = 0;
    float inv_keep = 1.f;
    if (DROP) {
      const DropCfg dc = drop_load(p.drop);
      rowkey = drop_rowkey(dc, static_cast<uint32_t>(bh * p.nq + q0 + x * kQTile + row));
      thr = dc.thr;
      inv_keep = dc.inv_keep;
    }
    // The two warpgroups take turns in the MUFU-bound exp section (named barriers 1 and 2): while one runs
    // exp2 the other waits for its next S tile, loads it and finds the row maximum.
    if (x == 1) named_bar_arrive(1, 256);             // warpgroup A goes first

    const bool trace_on = trace_cta && (threadIdx.x & 127) == 0;
    for (int j = 0; j < n_tiles; ++j) {
      const bool masked = (j == n_tiles - 1) && (tail < kKTile);
      HVC_FTR(x, j, 0);
      mbar_wait(&bar[BAR_SF + x], j & 1, 30);
      tc_fence_after();
      HVC_FTR(x, j, 1);
      // ---- pass 1: row maximum (S stays in TMEM; TMEM reads are cheap)
      const float mx = masked ? softmax_rowmax<true>(tS, tail) : softmax_rowmax<false>(tS, tail);
      const float m_new = fmaxf(m, mx * scale2);
      if (j == 0) {
        m = m_new;
      } else {
        const bool need = m_new > m + 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          // rare: fix up the O accumulator of this warp's 32 rows in TMEM
          mbar_wait(&bar[BAR_OD + x], (j - 1) & 1, 31);
          tc_fence_after();
          const float f = need ? ex2_approx(m - m_new) : 1.0f;
#pragma unroll
          for (int c = 0; c < HD; c += 32) {
            uint32_t o[32];
            tmem_ld_32x32(tO + c, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
            tmem_st_32x32(tO + c, o);
          }
          tmem_st_wait();
          if (need) { l *= f; m = m_new; }
        }
      }
      HVC_FTR(x, j, 2);
      // ---- pass 2 (my turn on the MUFU pipe): P = exp2(S*scale2 - m) -> TMEM, row sum
      const bool hand_over = !(x == 1 && j == n_tiles - 1);
      l += masked ? softmax_exp<true, DROP, EMU>(tS, tail, scale2, m, 1 + x, 1 + (x ^ 1), hand_over, rowkey, j * kKTile, thr, trace_on, x, j)
                  : softmax_exp<false, DROP, EMU>(tS, tail, scale2, m, 1 + x, 1 + (x ^ 1), hand_over, rowkey, j * kKTile, thr, trace_on, x, j);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive_warp(&bar[BAR_PF + x]);
      HVC_FTR(x, j, 5);
    }

    // ---- epilogue: O / l -> bf16 -> global, logsumexp
    mbar_wait(&bar[BAR_OD + x], (n_tiles - 1) & 1, 32);
    tc_fence_after();
    const int q = q0 + x * kQTile + row;
    const float inv = inv_keep / l;
    bf16* optr = p.o + (long long)(b * p.nq + q) * p.ldo + h * HD;
#pragma unroll
    for (int c = 0; c < HD; c += 32) {
      uint32_t o[32];
      tmem_ld_32x32(tO + c, o);
      tmem_ld_wait();
      if (q < p.nq) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 u;
          u.x = pack_bf16(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv);
          u.y = pack_bf16(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
          u.z = pack_bf16(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv);
          u.w = pack_bf16(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv);
          *reinterpret_cast<uint4*>(optr + c + i) = u;
        }
      }
    }
    if (q < p.nq && p.lse2 != nullptr) p.lse2[(long long)bh * p.nq_pad + q] = m + log2f(l);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace hvc

namespace hvc {
template <int HD, bool DROP, int EMU>
static int launch_attn_fwd(const hvc_attn_args* a, cudaStream_t st) {
  using L = FwdSmem<HD>;
  const uint64_t width = (uint64_t)a->heads * HD;
  const int swz = HD == 64 ? 1 : 2;
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  if ((rc = make_tmap_2d(&tmQ, a->q, 2, (uint64_t)a->batch * a->nq, width, a->ldq, HD, kQTile, swz))) return rc;
  if ((rc = make_tmap_2d(&tmK, a->k, 2, (uint64_t)a->batch * a->nk, width, a->ldk, HD, kKTile, swz))) return rc;
  if ((rc = make_tmap_2d(&tmV, a->v, 2, (uint64_t)a->batch * a->nk, width, a->ldv, HD, kKTile, swz))) return rc;
  AttnFwdKArgs ka;
  ka.batch = a->batch; ka.heads = a->heads; ka.nq = a->nq; ka.nk = a->nk;
  ka.n_kv_tiles = (a->nk + kKTile - 1) / kKTile;
  ka.o = reinterpret_cast<bf16*>(a->o); ka.ldo = a->ldo;
  ka.lse2 = reinterpret_cast<float*>(a->lse);
  ka.nq_pad = (a->nq + 127) / 128 * 128;
  ka.scale2 = a->scale * 1.4426950408889634f;
  ka.drop = make_drop(a->drop);
  HVC_SMEM_OPT_IN((attn_fwd_kernel<HD, DROP, EMU>), L::kTotal);
  dim3 grid((a->nq + 2 * kQTile - 1) / (2 * kQTile), a->batch * a->heads);
  attn_fwd_kernel<HD, DROP, EMU><<<grid, kFwdThreads, L::kTotal, st>>>(tmQ, tmK, tmV, ka);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

// store_attention slow path (vit_components.py:106-108): probs[b,h,i,j] = exp2(s2[i,j] - lse2[b,h,i]) in place, where the
// buffer holds s2 = q k^T * scale * log2(e) written by hvc_gemm (one GEMM per (b, h) on the head's column slice).
__global__ void __launch_bounds__(256) attn_probs_finalize_kernel(float* __restrict__ probs, const float* __restrict__ lse2, long long rows,
                                                                  int nq, int nk, int nq_pad) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);     // one warp per (b, h, i) row
  if (row >= rows) return;
  const long long bh = row / nq;
  const float l = lse2[bh * nq_pad + (row - bh * nq)];
  float* r = probs + row * nk;
  for (int j = threadIdx.x & 31; j < nk; j += 32) r[j] = ex2_approx(r[j] - l);
}

static int attn_store_probs(const hvc_attn_args* a, cudaStream_t st) {
  const int d = a->head_dim;
  const int nq_pad = (a->nq + 127) / 128 * 128;
  float* probs = reinterpret_cast<float*>(a->probs);
  for (int b = 0; b < a->batch; ++b) {
    for (int h = 0; h < a->heads; ++h) {
      hvc_gemm_args g;
      memset(&g, 0, sizeof(g));
      g.size = sizeof(g);
      g.M = a->nq; g.N = a->nk; g.K = d;
      g.A = reinterpret_cast<const bf16*>(a->q) + (long long)b * a->nq * a->ldq + h * d; g.lda = a->ldq; g.a_major = 0;
      g.B = reinterpret_cast<const bf16*>(a->k) + (long long)b * a->nk * a->ldk + h * d; g.ldb = a->ldk; g.b_major = 0;
      g.epilogue = HVC_EPI_F32; g.activation = HVC_ACT_NONE;
      g.out = probs + ((long long)b * a->heads + h) * a->nq * a->nk; g.ldo = a->nk;
      g.alpha = a->scale * 1.4426950408889634f;
      g.k_splits = 1;
      int rc = hvc_gemm(&g, st);
      if (rc) return rc;
    }
  }
  const long long rows = (long long)a->batch * a->heads * a->nq;
  attn_probs_finalize_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(probs, a->lse, rows, a->nq, a->nk, nq_pad);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}
}  // namespace hvc

#ifdef HVC_TRACE_FWD   // bring-up hooks (not part of include/hvc.h)
extern "C" int hvc_debug_fwd_trace_enable(int on) { return (int)cudaMemcpyToSymbol(hvc::g_fwd_trace_on, &on, sizeof(int)); }
extern "C" int hvc_debug_fwd_trace(unsigned long long* dst) {
  return (int)cudaMemcpyFromSymbol(dst, hvc::g_fwd_trace, sizeof(hvc::g_fwd_trace));
}
#endif

extern "C" int hvc_attn_fwd(const hvc_attn_args* a, void* stream) {
  using namespace hvc;
  HVC_CHECK_ARG(a != nullptr && a->size == sizeof(hvc_attn_args), "hvc_attn_fwd: bad args struct");
  HVC_CHECK_ARG(a->batch > 0 && a->heads > 0 && a->nq > 0 && a->nk > 0, "hvc_attn_fwd: empty problem");
  HVC_CHECK_ARG(a->head_dim == 64 || a->head_dim == 32, "hvc_attn_fwd: head_dim %d not supported (32 or 64)", a->head_dim);
  HVC_CHECK_ARG(a->q && a->k && a->v && a->o, "hvc_attn_fwd: null operand");
  HVC_CHECK_ARG((a->ldo & 7) == 0 && (reinterpret_cast<uintptr_t>(a->o) & 15) == 0, "hvc_attn_fwd: o must be 16-byte aligned rows");
  HVC_CHECK_ARG(a->probs == nullptr || a->lse != nullptr, "hvc_attn_fwd: probs needs lse");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool drop = a->drop.seed != nullptr && a->drop.p > 0.f;
  HVC_CHECK_ARG(!drop || a->drop.p < 1.f, "hvc_attn_fwd: dropout p must be < 1");
  const int rc = a->head_dim == 64 ? (drop ? launch_attn_fwd<64, true, kEmuDrop>(a, st) : launch_attn_fwd<64, false, kEmu64>(a, st))
                         : (drop ? launch_attn_fwd<32, true, kEmuDrop32>(a, st) : launch_attn_fwd<32, false, kEmu32>(a, st));
  if (rc != HVC_OK || a->probs == nullptr) return rc;
  return attn_store_probs(a, st);
}
