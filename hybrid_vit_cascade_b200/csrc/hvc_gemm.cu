// hvc_gemm.cu -- persistent warp-specialised bf16 GEMM for sm_100a.
//
//   D[M,N] = alpha * sum_k A(m,k) B(n,k)         fp32 accumulation in TMEM
//
// One CTA per SM, 320 threads:
//   warp 0      TMA producer   : cp.async.bulk.tensor -> 128B-swizzled smem ring (5 stages x 32 KB)
//   warp 1      MMA issuer     : one elected lane issues tcgen05.mma (128x128x16), commits to mbarriers;
//                                owns the TMEM allocation (2 accumulator stages x 128 columns)
//   warps 2..9  epilogue       : tcgen05.ld accumulator -> registers -> fused epilogue -> global; two warps per TMEM
//                                lane quarter, each taking 64 of the tile's 128 columns
// The epilogue of tile i overlaps the main loop of tile i+1 through the two TMEM stages.  At the backbone's K = 256 a
// tile's main loop is only ~1700 clk, so the epilogue is the bound: it is specialised per epilogue kind (template), uses
// 32-byte global accesses (each thread owns a row), fetches residual / aux operands while the TMEM load is in flight, and
// evaluates GELU with a 2-MUFU erf (hvc_common.cuh) -- with four epilogue warps and erff() the MLP GEMMs ran at
// 145-216 TFLOP/s (profiles/r01_gemm_time_by_shape.log).
//
// Operands may be K-major (stored [rows, K]) or MN-major (stored [K, rows]); the latter is what the
// backward GEMMs need (dgrad reads W as stored, wgrad reads dy and x as stored) so no transposed copies
// of activations or weights are ever materialised.
//
// Roofline: tensor pipe for K >= ~512; at the backbone's C=256 projections the kernel sits on the
// HBM/tensor ridge (A read + D write vs 2*M*N*K flops), so the fused epilogues are what matters.
#include <stdlib.h>

#include "hvc_common.cuh"
#include "hvc_host.h"

namespace hvc {

constexpr int BM = 128, BN = 128, BK = 64;
constexpr int kStages = 5;
constexpr int kTileBytes = BM * BK * 2;  // 16 KB, same for A and B
constexpr int kStageBytes = 2 * kTileBytes;
constexpr int kGemmThreads = 320;
constexpr int kEpiWarps = 8;
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
  int v256;       // every epilogue operand (out, out2, resid, aux) is 32-byte aligned with a 32-byte-multiple row pitch
  // implicit 3x3x3 convolution over a zero-padded channels-last volume (hvc_gemm_args::taps): the operand on `taps_side` is a
  // [padded voxels, tap_cin] matrix read with a per-tap row shift instead of a materialised patch matrix
  int taps_side, tap_cin, n_taps;
  int tap_off[27];
  int n_mma;      // N of the tcgen05.mma: 128, or 64 / 32 when the whole problem is that narrow (the Cout = 32 / 64 convs of the stage
                  // wrappers).  The B box then has n_mma rows, and an operand half that lies wholly outside the problem is not loaded:
                  // measured on B200, a TMA box that is mostly out-of-bounds zero fill costs far more than the same box of data
                  // (profiles/r01_implicit_conv_gemm.log)
  int a_halves, b_halves;   // MN-major operands: 64-element halves to load per k-block (2, or 1 when M / N <= 64)
  uint32_t stage_tx;        // bytes one stage's loads deliver
  // Resident weight panel (round 2): with K <= 256 the whole B panel of an N tile (<= 4 k-blocks x 16 KB) fits in the B halves of the ring's
  // first stages.  CTA c keeps N tile c % n_blocks for its whole life, loads that panel once and streams only A through the ring; the
  // n_blocks CTAs of a group walk the same M tiles in step, so an A tile is fetched from HBM once and shared through L2.  (A contiguous
  // n-major range per CTA was tried first: the CTAs of different N tiles drift apart and A is re-read from HBM -- qkv 147 -> 197 us.)  The K = 256 projections ran at the L2 -> SM throughput cap of the chip (every 128x128 tile fetched 64 KB of A and
  // the same 64 KB weight panel again: 5800 of ~6300 B/clk, profiles/r02_gemm_time_by_shape.log); this halves that traffic.
  int b_resident, tiles_per_cta;
  uint32_t stage_tx_a, panel_tx;
};

// row shift of tap t (hvc_conv_taps::offsets); a half-tile past the last tap reads tap n_taps-1 again (its columns are never stored)
__device__ __forceinline__ int tap_shift(const GemmKArgs& p, int tap) { return p.tap_off[min(tap, p.n_taps - 1)]; }

struct WorkItem {
  int m0, n0, kb0, kb1;
};
__device__ __forceinline__ WorkItem decode_work(const GemmKArgs& p, int w) {
  if (p.b_resident) {     // this CTA keeps ONE N tile (blockIdx % n_blocks) and walks the M tiles w = group, group + groups, ...
    WorkItem it;
    it.m0 = w * BM; it.n0 = (static_cast<int>(blockIdx.x) % p.n_blocks) * BN; it.kb0 = 0; it.kb1 = p.k_blocks;
    return it;
  }
#ifdef HVC_GEMM_SPLIT_FASTEST
  const int tile = w / p.k_splits, split = w - tile * p.k_splits;
#else
  // tile-fastest: the CTAs in flight at any time cover ALL output tiles of a few K ranges, so each K range of A and B comes from
  // HBM once and is shared through L2 (split-fastest re-read A once per N tile and B once per M tile)
  const int tiles = p.m_blocks * p.n_blocks;
  const int split = w / tiles, tile = w - split * tiles;
#endif
  const int mb = tile / p.n_blocks, nb = tile - mb * p.n_blocks;
  WorkItem it;
  it.m0 = mb * BM;
  it.n0 = nb * BN;
  it.kb0 = split * p.kb_per_split;
  it.kb1 = min(it.kb0 + p.kb_per_split, p.k_blocks);
  return it;
}

// nn.Dropout on the 32 values of one row chunk: element (row, col) of this site, see hvc_common.cuh.  The chunk starts at a multiple of
// 32 inside its 128-column block; one instantiation per start makes the 32 in-block multipliers compile-time constants.
template <int C0>
__device__ __forceinline__ void epilogue_dropout_chunk(float (&v)[32], uint32_t bk, uint32_t thr, float inv_keep) {
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = drop_keep_in_block(bk, C0 + j, thr) ? v[j] * inv_keep : 0.f;
}
__device__ __forceinline__ void epilogue_dropout(const GemmKArgs& p, float (&v)[32], int row, int col0) {
  const DropCfg dc = drop_load(p.drop);
  const uint32_t bk = drop_blockkey(drop_rowkey(dc, static_cast<uint32_t>(row)), static_cast<uint32_t>(col0));   // col0 % 32 == 0
  switch ((static_cast<uint32_t>(col0) & 127u) >> 5) {
    case 0: epilogue_dropout_chunk<0>(v, bk, dc.thr, dc.inv_keep); break;
    case 1: epilogue_dropout_chunk<32>(v, bk, dc.thr, dc.inv_keep); break;
    case 2: epilogue_dropout_chunk<64>(v, bk, dc.thr, dc.inv_keep); break;
    default: epilogue_dropout_chunk<96>(v, bk, dc.thr, dc.inv_keep); break;
  }
}

// ---------------------------------------------------------------- epilogue for one 32-column chunk
// Operands that come from global memory (residual row, aux row) are fetched into `Side` before the TMEM load is waited
// for, on the fast path (full chunk, 32-byte aligned).
struct Side {
  uint32_t r[32];   // residual, f32 bits              (HVC_EPI_RESIDUAL)
  uint32_t a[16];   // aux, bf16 pairs                 (HVC_ACT_GELU_GRAD)
};
template <int EPI>
__device__ __forceinline__ void side_load(const GemmKArgs& p, int row, int col0, bool fast, Side& s) {
  if (!fast) return;
  if (EPI == HVC_EPI_RESIDUAL) {
    const float* r = p.resid + (long long)row * p.ldr + col0;
#pragma unroll
    for (int j = 0; j < 4; ++j) ldg256(r + 8 * j, *reinterpret_cast<uint32_t(*)[8]>(&s.r[8 * j]));
  }
  if (EPI == HVC_EPI_BF16 && p.activation == HVC_ACT_GELU_GRAD) {
    const bf16* ax = p.aux + (long long)row * p.ldaux + col0;
    ldg256(ax, *reinterpret_cast<uint32_t(*)[8]>(&s.a[0]));
    ldg256(ax + 16, *reinterpret_cast<uint32_t(*)[8]>(&s.a[8]));
  }
}
__device__ __forceinline__ void store_bf16_row(bf16* o, const float (&v)[32], int ncols, bool fast) {
  if (fast) {
#pragma unroll
    for (int j = 0; j < 32; j += 16) {
      uint32_t u[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) u[k] = pack_bf16(v[j + 2 * k], v[j + 2 * k + 1]);
      stg256(o + j, u);
    }
  } else {
    _Pragma("unroll") for (int j = 0; j < 32; ++j) if (j < ncols) o[j] = __float2bfloat16(v[j]);
  }
}
__device__ __forceinline__ void store_f32_row(float* o, const float (&v)[32], int ncols, bool fast) {
  if (fast) {
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      uint32_t u[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) u[k] = __float_as_uint(v[j + k]);
      stg256(o + j, u);
    }
  } else {
    _Pragma("unroll") for (int j = 0; j < 32; ++j) if (j < ncols) o[j] = v[j];
  }
}

template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const GemmKArgs& p, const uint32_t (&acc)[32], int row, int col0, bool fast, const Side& s) {
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
  const bool drop = p.drop.seed != nullptr;
  if (drop && EPI != HVC_EPI_BF16) epilogue_dropout(p, v, row, col0);   // residual / f32: the value before gate + residual

  if (EPI == HVC_EPI_BF16) {
    if (p.out2 != nullptr) store_bf16_row(reinterpret_cast<bf16*>(p.out2) + (long long)row * p.ldo2 + col0, v, ncols, fast);
    if (p.activation == HVC_ACT_GELU) {
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float2 g = gelu_fast2(make_float2(v[j], v[j + 1]));
        v[j] = g.x; v[j + 1] = g.y;
      }
    } else if (p.activation == HVC_ACT_GELU_GRAD) {
      if (fast) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float2 g = gelu_grad_fast2(unpack_bf16(s.a[j]));
          v[2 * j] *= g.x;
          v[2 * j + 1] *= g.y;
        }
      } else {
        const bf16* ax = p.aux + (long long)row * p.ldaux + col0;
        _Pragma("unroll") for (int j = 0; j < 32; ++j) if (j < ncols) v[j] *= gelu_grad_fast(__bfloat162float(ax[j]));
      }
    }
    if (drop) epilogue_dropout(p, v, row, col0);   // after the activation (mlp: Linear -> GELU -> Dropout); out2 stays pre-activation
    store_bf16_row(reinterpret_cast<bf16*>(p.out) + (long long)row * p.ldo + col0, v, ncols, fast);
  } else if (EPI == HVC_EPI_RESIDUAL) {
    if (p.out2 != nullptr) store_bf16_row(reinterpret_cast<bf16*>(p.out2) + (long long)row * p.ldo2 + col0, v, ncols, fast);
    float* o = reinterpret_cast<float*>(p.out) + (long long)row * p.ldo + col0;
    const float* g = p.gate ? p.gate + (long long)(row / p.rows_per_batch) * p.gate_ld + col0 : nullptr;
    if (fast && (g == nullptr || (p.gate_ld & 3) == 0)) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 gg = make_float4(1.f, 1.f, 1.f, 1.f);
        if (g) gg = __ldg(reinterpret_cast<const float4*>(g + j));
        v[j] = fmaf(gg.x, v[j], __uint_as_float(s.r[j]));
        v[j + 1] = fmaf(gg.y, v[j + 1], __uint_as_float(s.r[j + 1]));
        v[j + 2] = fmaf(gg.z, v[j + 2], __uint_as_float(s.r[j + 2]));
        v[j + 3] = fmaf(gg.w, v[j + 3], __uint_as_float(s.r[j + 3]));
      }
      store_f32_row(o, v, ncols, true);
    } else {
      const float* r = p.resid + (long long)row * p.ldr + col0;
      _Pragma("unroll") for (int j = 0; j < 32; ++j) if (j < ncols) o[j] = fmaf(g ? g[j] : 1.f, v[j], r[j]);
    }
  } else if (EPI == HVC_EPI_F32_ATOMIC) {
    float* o = reinterpret_cast<float*>(p.out) + (long long)row * p.ldo + col0;
    if (ncols == 32 && (p.ldo & 3) == 0) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) atomicAdd(reinterpret_cast<float4*>(o + j), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
    } else {
      _Pragma("unroll") for (int j = 0; j < 32; ++j) if (j < ncols) atomicAdd(o + j, v[j]);
    }
  } else {  // HVC_EPI_F32
    if (p.activation == HVC_ACT_GELU) {   // (kept exact: this epilogue also serves verification-style callers)
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
    }
    store_f32_row(reinterpret_cast<float*>(p.out) + (long long)row * p.ldo + col0, v, ncols, fast);
  }
}

// ---------------------------------------------------------------- kernel
template <int A_MAJOR, int B_MAJOR, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmKArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;
  uint64_t* tempty_bar = tfull_bar + kAccStages;
  uint64_t* panel_full = tempty_bar + kAccStages;      // resident weight panel: loaded / released (b_resident)
  uint64_t* panel_empty = panel_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(panel_empty + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_work = p.m_blocks * p.n_blocks * p.k_splits;
  // work items of this CTA: strided over the grid, or (resident panel) one contiguous range
  const int w_first = p.b_resident ? static_cast<int>(blockIdx.x) / p.n_blocks : blockIdx.x;
  const int w_step = p.b_resident ? p.tiles_per_cta /* = groups */ : gridDim.x;
  const int w_end = p.b_resident ? p.m_blocks : num_work;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < kAccStages; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], kEpiWarps);
    }
    mbar_init(panel_full, 1);
    mbar_init(panel_empty, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kAccStages * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    uint32_t stage = 0, phase = 0, panel_gen = 0;
    int panel_n0 = -1;
    auto load_b = [&](uint8_t* sB, uint64_t* bar, int kb, int n0) {
      if (B_MAJOR == kMajorK) {
        tma_load_2d(sB, &tmB, bar, kb * BK, n0);
      } else if (p.taps_side == 2) {   // each 64-column half lies inside one tap (tap_cin % 64 == 0): K rows shifted per tap
        const int t0 = n0 / p.tap_cin, t1 = (n0 + 64) / p.tap_cin;
        tma_load_2d(sB, &tmB, bar, n0 - t0 * p.tap_cin, kb * BK + tap_shift(p, t0));
        tma_load_2d(sB + kTileBytes / 2, &tmB, bar, n0 + 64 - t1 * p.tap_cin, kb * BK + tap_shift(p, t1));
      } else {
        tma_load_2d(sB, &tmB, bar, n0, kb * BK);
        if (p.b_halves == 2) tma_load_2d(sB + kTileBytes / 2, &tmB, bar, n0 + 64, kb * BK);
      }
    };
    for (int w = w_first; w < w_end; w += w_step) {
      const WorkItem it = decode_work(p, w);
      if (p.b_resident && it.n0 != panel_n0) {
        // new N tile: once the MMAs of the previous one have read the old panel, load k-block kb's B tile into stage kb's B half
        mbar_wait(panel_empty, (panel_gen & 1) ^ 1, 5);
        if (elect_one()) {
          mbar_arrive_expect_tx(panel_full, p.panel_tx);
          for (int kb = 0; kb < p.k_blocks; ++kb) load_b(smem + kb * kStageBytes + kTileBytes, panel_full, kb, it.n0);
        }
        __syncwarp();
        panel_n0 = it.n0;
        ++panel_gen;
      }
      for (int kb = it.kb0; kb < it.kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1, 1);
        if (elect_one()) {
          uint8_t* sA = smem + stage * kStageBytes;
          uint8_t* sB = sA + kTileBytes;
          mbar_arrive_expect_tx(&full_bar[stage], p.b_resident ? p.stage_tx_a : p.stage_tx);
          if (A_MAJOR == kMajorK) {
            if (p.taps_side == 1) {   // k-block kb lies inside one tap (tap_cin % BK == 0): columns of that tap, rows shifted
              const int k = kb * BK, tap = k / p.tap_cin;
              tma_load_2d(sA, &tmA, &full_bar[stage], k - tap * p.tap_cin, it.m0 + tap_shift(p, tap));
            } else {
              tma_load_2d(sA, &tmA, &full_bar[stage], kb * BK, it.m0);
            }
          } else {
            tma_load_2d(sA, &tmA, &full_bar[stage], it.m0, kb * BK);
            if (p.a_halves == 2) tma_load_2d(sA + kTileBytes / 2, &tmA, &full_bar[stage], it.m0 + 64, kb * BK);
          }
          if (!p.b_resident) load_b(sB, &full_bar[stage], kb, it.n0);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = make_idesc_bf16(BM, static_cast<uint32_t>(p.n_mma), A_MAJOR, B_MAJOR);
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0, panel_gen = 0;
    int panel_n0 = -1;
    for (int w = w_first; w < w_end; w += w_step) {
      const WorkItem it = decode_work(p, w);
      if (it.kb0 >= it.kb1) continue;
      if (p.b_resident && it.n0 != panel_n0) {
        mbar_wait(panel_full, panel_gen & 1, 6);
        panel_n0 = it.n0;
        ++panel_gen;
      }
      // last tile that reads this panel: the next work item of this CTA belongs to another N tile (or there is none)
      const bool panel_last = p.b_resident && w + w_step >= w_end;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1, 2);
      tc_fence_after();
      for (int kb = it.kb0; kb < it.kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase, 3);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a0 = smem_u32(smem + stage * kStageBytes);
          const uint32_t b0 = p.b_resident ? smem_u32(smem + kb * kStageBytes) + kTileBytes : a0 + kTileBytes;
#pragma unroll
          for (int k16 = 0; k16 < BK / 16; ++k16) {
            const uint64_t ad = (A_MAJOR == kMajorK) ? make_sdesc_sw128(a0 + k16 * 32, 16, 1024)
                                                     : make_sdesc_sw128(a0 + k16 * 2048, kTileBytes / 2, 1024);
            const uint64_t bd = (B_MAJOR == kMajorK) ? make_sdesc_sw128(b0 + k16 * 32, 16, 1024)
                                                     : make_sdesc_sw128(b0 + k16 * 2048, kTileBytes / 2, 1024);
            umma_ss(tmem_base + acc * BN, ad, bd, idesc, (kb > it.kb0 || k16 > 0) ? 1u : 0u);
          }
          tc_commit(&empty_bar[stage]);
          if (kb == it.kb1 - 1) {
            tc_commit(&tfull_bar[acc]);
            if (panel_last) tc_commit(panel_empty);
          }
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int quarter = warp & 3;            // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;        // which 64 of the tile's 128 columns
    uint32_t acc = 0, acc_phase = 0;
    for (int w = w_first; w < w_end; w += w_step) {
      const WorkItem it = decode_work(p, w);
      if (it.kb0 >= it.kb1) continue;
      mbar_wait(&tfull_bar[acc], acc_phase, 4);
      tc_fence_after();
      const int row = it.m0 + quarter * 32 + lane;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + half * 64;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int col0 = it.n0 + half * 64 + c * 32;
        if (col0 >= p.N) break;                 // nothing of this chunk is stored (and a narrow MMA never wrote these TMEM columns)
        const bool fast = p.v256 && row < p.M && col0 + 32 <= p.N;
        uint32_t v[32];
        Side side;
        tmem_ld_32x32(taddr + c * 32, v);
        side_load<EPI>(p, row, col0, fast, side);
        tmem_ld_wait();
        epilogue_chunk<EPI>(p, v, row, col0, fast, side);
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

template <int A_MAJOR, int B_MAJOR, int EPI>
static int launch_gemm_epi(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmKArgs& ka, int grid, cudaStream_t st) {
  HVC_SMEM_OPT_IN((gemm_bf16_kernel<A_MAJOR, B_MAJOR, EPI>), kGemmSmem);
  gemm_bf16_kernel<A_MAJOR, B_MAJOR, EPI><<<grid, kGemmThreads, kGemmSmem, st>>>(tmA, tmB, ka);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}
template <int A_MAJOR, int B_MAJOR>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmKArgs& ka, int grid, cudaStream_t st) {
  switch (ka.epilogue) {
    case HVC_EPI_BF16: return launch_gemm_epi<A_MAJOR, B_MAJOR, HVC_EPI_BF16>(tmA, tmB, ka, grid, st);
    case HVC_EPI_RESIDUAL: return launch_gemm_epi<A_MAJOR, B_MAJOR, HVC_EPI_RESIDUAL>(tmA, tmB, ka, grid, st);
    case HVC_EPI_F32_ATOMIC: return launch_gemm_epi<A_MAJOR, B_MAJOR, HVC_EPI_F32_ATOMIC>(tmA, tmB, ka, grid, st);
    default: return launch_gemm_epi<A_MAJOR, B_MAJOR, HVC_EPI_F32>(tmA, tmB, ka, grid, st);
  }
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

  const hvc_conv_taps& tp = a->taps;
  HVC_CHECK_ARG(tp.side >= 0 && tp.side <= 2, "hvc_gemm: taps.side must be 0, 1 or 2");
  if (tp.side != 0) HVC_CHECK_ARG(tp.n_taps >= 1 && tp.n_taps <= 27 && tp.rows >= 0, "hvc_gemm: taps.n_taps must be 1..27");
  if (tp.side == 1) HVC_CHECK_ARG(a->a_major == 0 && tp.cin > 0 && tp.cin % BK == 0 && a->K == tp.n_taps * tp.cin,
                                  "hvc_gemm: taps on A need a K-major A, cin %% 64 == 0 and K == n_taps*cin");
  if (tp.side == 2) HVC_CHECK_ARG(a->b_major == 1 && tp.cin > 0 && tp.cin % 64 == 0 && a->N == tp.n_taps * tp.cin,
                                  "hvc_gemm: taps on B need an MN-major B, cin %% 64 == 0 and N == n_taps*cin");

  const int n_mma = (a->N <= 32 && a->b_major == 0) ? 32 : (a->N <= 64 ? 64 : BN);   // an MN-major B keeps whole 64-element swizzle rows
  CUtensorMap tmA, tmB;
  int rc;
  if (a->a_major == 0) rc = make_tmap_2d(&tmA, a->A, 2, (tp.side == 1 && tp.rows) ? tp.rows : a->M, tp.side == 1 ? tp.cin : a->K, a->lda, BK, BM, true);
  else                 rc = make_tmap_2d(&tmA, a->A, 2, a->K, a->M, a->lda, 64, BK, true);
  if (rc) return rc;
  if (a->b_major == 0) rc = make_tmap_2d(&tmB, a->B, 2, a->N, a->K, a->ldb, BK, n_mma, true);
  else                 rc = make_tmap_2d(&tmB, a->B, 2, (tp.side == 2 && tp.rows) ? tp.rows : a->K, tp.side == 2 ? tp.cin : a->N, a->ldb, 64, BK, true);
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
  ka.taps_side = tp.side; ka.tap_cin = tp.cin; ka.n_taps = tp.side ? tp.n_taps : 1;
  for (int t = 0; t < 27; ++t) ka.tap_off[t] = tp.side ? tp.offsets[t] : 0;
  ka.n_mma = n_mma;
  ka.a_halves = (a->a_major == 1 && a->M <= 64) ? 1 : 2;
  ka.b_halves = (a->b_major == 1 && a->N <= 64) ? 1 : 2;
  ka.stage_tx = (a->a_major == 0 ? kTileBytes : ka.a_halves * (kTileBytes / 2)) +
                (a->b_major == 0 ? n_mma * BK * 2 : ka.b_halves * (kTileBytes / 2));
  {
    auto ok32 = [](const void* ptr, long long ld, int esz) {
      return ptr == nullptr || ((reinterpret_cast<uintptr_t>(ptr) & 31u) == 0 && ((ld * esz) & 31) == 0);
    };
    const int out_esz = a->epilogue == HVC_EPI_BF16 ? 2 : 4;
    ka.v256 = ok32(a->out, a->ldo, out_esz) && ok32(a->out2, a->ldo2, 2) && ok32(a->resid, a->ldr, 4) && ok32(a->aux, a->ldaux, 2);
  }
  HVC_CHECK_ARG(ka.drop.seed == nullptr || (a->epilogue != HVC_EPI_F32_ATOMIC && a->drop.p < 1.f), "hvc_gemm: dropout needs p < 1 and a non-atomic epilogue");

  const long long num_work = (long long)ka.m_blocks * ka.n_blocks * ka.k_splits;
  const int sms = device_sm_count();
  int grid = (int)(num_work < sms ? num_work : sms);
  // resident weight panel: whole-K panels that fit the ring's B slots, no split-K / taps, at most one group's worth of N tiles per SM
  // count and enough M tiles per CTA to amortise the panel load; the grid is a whole number of groups of n_blocks CTAs
  ka.b_resident = (getenv("HVC_GEMM_NO_RESIDENT") == nullptr && ka.k_blocks <= kStages - 1 && ka.k_splits == 1 && tp.side == 0 &&
                   ka.n_blocks <= sms / 4 && num_work >= 8LL * sms) ? 1 : 0;
  ka.tiles_per_cta = 1;
  if (ka.b_resident) {
    const int groups = sms / ka.n_blocks;
    ka.tiles_per_cta = groups;              // the step between the M tiles of one CTA
    grid = groups * ka.n_blocks;
  }
  ka.stage_tx_a = a->a_major == 0 ? kTileBytes : ka.a_halves * (kTileBytes / 2);
  ka.panel_tx = (uint32_t)ka.k_blocks * (ka.stage_tx - ka.stage_tx_a);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (a->a_major == 0 && a->b_major == 0) return launch_gemm<kMajorK, kMajorK>(tmA, tmB, ka, grid, st);
  if (a->a_major == 0 && a->b_major == 1) return launch_gemm<kMajorK, kMajorMN>(tmA, tmB, ka, grid, st);
  if (a->a_major == 1 && a->b_major == 0) return launch_gemm<kMajorMN, kMajorK>(tmA, tmB, ka, grid, st);
  return launch_gemm<kMajorMN, kMajorMN>(tmA, tmB, ka, grid, st);
}
