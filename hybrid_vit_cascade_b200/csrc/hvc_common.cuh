// hvc_common.cuh -- sm_100a building blocks shared by every kernel in libhvc_sm100a.so:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld / st) wrappers,
// UMMA shared-memory + instruction descriptors, small math helpers.
//
// Descriptor bit layouts follow the PTX ISA "tcgen05 matrix/instruction descriptor" tables
// (cross-checked against cute/arch/mma_sm100_desc.hpp field comments):
//   smem desc : [0,14) addr>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) swizzle
//   instr desc: [4,6) D fmt (1=f32) | [7,10) A fmt (1=bf16) | [10,13) B fmt | 15 A MN-major |
//               16 B MN-major | [17,23) N>>3 | [24,29) M>>4
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/hvc.h"

namespace hvc {

typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;

// ------------------------------------------------------------------ misc
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t.reg .b32 R1;\n\t"
      "elect.sync R1|P1, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// generic-proxy writes (st.shared) -> visible to the async proxy (TMA store / tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// One arrival for the whole (converged) warp: every lane's earlier writes / tcgen05 operations are ordered before the
// arrival by the warp barrier.  Per-thread arrivals on one mbarrier serialise in the shared-memory atomic unit -- 512 of
// them cost several hundred clocks per synchronisation point (profiles/r01_attn_bwd_timeline_d64_v4a.log).
__device__ __forceinline__ void mbar_arrive_warp(uint64_t* bar) {
  __syncwarp();
  if ((threadIdx.x & 31u) == 0) mbar_arrive(bar);
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// try_wait parks the warp in hardware until the phase completes or a system time limit expires.  Tried and rejected (round 2): the optional
// suspend-time hint (1 ms) to thin out the retry loops of the mostly-waiting TMA / MMA / drain warps (0.45 TRYWAIT + 0.9 BRA per score in
// attn_bwd_kernel, profiles/r02_ncu_full_attn_d64_b1.csv) -- wake-ups became slower instead: attention backward 797 -> 743, forward
// 797 -> 766 TFLOP/s at d = 64.
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must not hang the GPU (a hung box is a lost round).  After ~4 s of
// polling the kernel traps, which surfaces as a CUDA error at the next sync.
#ifndef HVC_WAIT_LIMIT_NS
#define HVC_WAIT_LIMIT_NS 4000000000ull
#endif
// HVC_WAIT_INLINE (kernels that re-partition registers with setmaxnreg): the slow path is inlined and reports through a device word
// instead of printf.  A real call -- to a __noinline__ function or to vprintf -- makes ptxas allocate the WHOLE kernel inside the smallest
// setmaxnreg budget of any call site (measured: attn_bwd_kernel stayed below R61 with budgets 64 / 104 / 168 until the calls were gone).
static __device__ unsigned int g_hvc_wait_timeout_tag = 0;      // 0 = no timeout; else (tag << 16 | thread) of the first waiter that gave up
#ifdef HVC_WAIT_INLINE
static __device__ __forceinline__ void mbar_wait_slow(uint64_t* bar, uint32_t parity, int tag) {
  const uint64_t t0 = globaltimer_ns();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3ffu) == 0 && globaltimer_ns() - t0 > HVC_WAIT_LIMIT_NS) {
      atomicCAS(&g_hvc_wait_timeout_tag, 0u, (static_cast<unsigned>(tag) << 16) | threadIdx.x);
      __trap();
    }
  }
}
#else
static __device__ __noinline__ void mbar_wait_slow(uint64_t* bar, uint32_t parity, int tag) {
  const uint64_t t0 = globaltimer_ns();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3ffu) == 0 && globaltimer_ns() - t0 > HVC_WAIT_LIMIT_NS) {
      printf("[hvc] mbarrier wait timeout: block (%d,%d,%d) thread %d tag %d parity %u\n", blockIdx.x,
             blockIdx.y, blockIdx.z, threadIdx.x, tag, parity);
      __trap();
    }
  }
}
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag = 0) {
  if (mbar_try_wait(bar, parity)) return;
  mbar_wait_slow(bar, parity, tag);
}

// ------------------------------------------------------------------ TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// L2 eviction-priority policies (createpolicy encodings used by TMA cache hints)
constexpr uint64_t kEvictNormal = 0x1000000000000000ull;
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            uint64_t hint = kEvictNormal) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
      "l"(hint)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2, uint64_t hint = kEvictNormal) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
      "r"(c2), "l"(hint)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
// smem (fp32, contiguous) += into global (fp32, contiguous): L2-side reduction, no return traffic.
__device__ __forceinline__ void bulk_reduce_add_f32(float* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
               ::"l"(reinterpret_cast<uint64_t>(gdst)), "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
// smem tile (fp32) += into a global tensor through its tensor map (L2-side reduction, un-swizzles on the way)
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ------------------------------------------------------------------ tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// Arrive on an mbarrier once every tcgen05.mma issued so far by this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]   (A: lane = row, 32-bit column = two packed 16-bit K elements)
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

constexpr uint32_t kMajorK = 0, kMajorMN = 1;
__host__ __device__ constexpr uint32_t make_idesc_bf16(uint32_t M, uint32_t N, uint32_t a_major = kMajorK,
                                                       uint32_t b_major = kMajorK) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (a_major << 15) | (b_major << 16) | ((N >> 3) << 17) |
         ((M >> 4) << 24);
}
// 128-byte-swizzled operand tile.  K-major: rows of 64 bf16 (128 B), 8-row groups 1024 B apart (SBO);
// a K=16 step advances the start address by 32 B inside the swizzle atom.  MN-major: rows are K
// indices holding 64 contiguous MN elements; SBO = 1024 B between 8-row K groups, LBO = byte distance
// between 64-element MN chunks; a K=16 step advances the start address by 16 rows = 2048 B.
// layout type (bits [61,64)): 2 = SWIZZLE_128B, 4 = SWIZZLE_64B (rows of 64 B = head_dim 32 tiles: 8-row groups are
// 512 B apart, a K=16 step of an MN-major operand advances 16 rows = 1024 B)
template <uint32_t LAYOUT>
__device__ __forceinline__ uint64_t make_sdesc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;  // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(LAYOUT) << 61;
  return d;
}
__device__ __forceinline__ uint64_t make_sdesc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return make_sdesc<2>(smem_addr, lbo_bytes, sbo_bytes);
}
// Swizzled operand tiles whose rows are ROWB bytes (128 -> SWIZZLE_128B, 64 -> SWIZZLE_64B).
template <int ROWB>
struct Swz {
  static_assert(ROWB == 128 || ROWB == 64, "row bytes");
  static constexpr uint32_t kLayout = ROWB == 128 ? 2u : 4u;
  static constexpr uint32_t kGroup = 8u * ROWB;        // SBO: 8-row group pitch
  static constexpr uint32_t kMnStep = 16u * ROWB;      // MN-major operand: one K=16 step = 16 rows
  __device__ static __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo = 16) { return make_sdesc<kLayout>(addr, lbo, kGroup); }
  // byte offset of 16-byte chunk `chunk` of row `row`
  __device__ static __forceinline__ uint32_t offset(uint32_t row, uint32_t chunk) {
    return ROWB == 128 ? row * 128u + ((chunk ^ (row & 7u)) << 4) : row * 64u + ((chunk ^ ((row >> 1) & 3u)) << 4);
  }
};

#define HVC_R4(v, i) "=r"(v[i]), "=r"(v[i + 1]), "=r"(v[i + 2]), "=r"(v[i + 3])
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets lane (taddr.lane + t).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : HVC_R4(v, 0), HVC_R4(v, 4), HVC_R4(v, 8), HVC_R4(v, 12), HVC_R4(v, 16), HVC_R4(v, 20), HVC_R4(v, 24),
        HVC_R4(v, 28)
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : HVC_R4(v, 0), HVC_R4(v, 4), HVC_R4(v, 8), HVC_R4(v, 12)
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : HVC_R4(v, 0), HVC_R4(v, 4)
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
#define HVC_S4(v, i) "r"(v[i]), "r"(v[i + 1]), "r"(v[i + 2]), "r"(v[i + 3])
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), HVC_S4(v, 0), HVC_S4(v, 4), HVC_S4(v, 8), HVC_S4(v, 12)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), HVC_S4(v, 0), HVC_S4(v, 4), HVC_S4(v, 8), HVC_S4(v, 12), HVC_S4(v, 16), HVC_S4(v, 20),
      HVC_S4(v, 24), HVC_S4(v, 28)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), HVC_S4(v, 0), HVC_S4(v, 4)
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// byte offset of 16-byte chunk `chunk` of row `row` in a tile whose rows are ROWB bytes, swizzled the way TMA does for
// that row size (128 -> SWIZZLE_128B, 64 -> SWIZZLE_64B, 32 -> SWIZZLE_32B): chunk index ^= address bits [7, 7+log2(ROWB/16))
template <int ROWB>
__device__ __forceinline__ uint32_t swz_offset(uint32_t row, uint32_t chunk) {
  static_assert(ROWB == 128 || ROWB == 64 || ROWB == 32, "row bytes");
  if (ROWB == 128) return row * 128u + ((chunk ^ (row & 7u)) << 4);
  if (ROWB == 64) return row * 64u + ((chunk ^ ((row >> 1) & 3u)) << 4);
  return row * 32u + ((chunk ^ ((row >> 2) & 1u)) << 4);
}

// Register re-partitioning between warpgroups (all four warps of a warpgroup execute the same instruction): the block is
// launched with a uniform budget; producer / MMA warps give registers back, the elementwise warpgroups take them.
template <int N>
__device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

// ------------------------------------------------------------------ named barriers
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// explicit shared-space vector load (keeps LDS even when the pointer's provenance is lost to the compiler)
__device__ __forceinline__ float4 lds_f4(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts_u16(uint32_t saddr, uint16_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(saddr), "h"(v) : "memory");
}
__device__ __forceinline__ void sts_u4(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ------------------------------------------------------------------ math
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  bf162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// same, negative inputs clamp to +0 (the sign bit carries the dropout decision in the attention backward)
__device__ __forceinline__ uint32_t pack_bf16_relu(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ uint4 lds_u4(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
  bf162 v = *reinterpret_cast<bf162*>(&u);
  return __bfloat1622float2(v);
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^(-x): the negation is folded into the MUFU source modifier by ptxas
__device__ __forceinline__ float ex2_neg_approx(float x) {
  float y;
  asm("{\n\t.reg .f32 t;\n\tneg.ftz.f32 t, %1;\n\tex2.approx.ftz.f32 %0, t;\n\t}" : "=f"(y) : "f"(x));
  return y;
}
// 256-bit global accesses (one full 32-byte sector per lane): the GEMM epilogue's threads each own a row, so a warp-wide
// store touches 32 different lines; 32-byte pieces halve the number of sector transactions of the 16-byte form.
__device__ __forceinline__ void ldg256(const void* p, uint32_t (&v)[8]) {
  asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint32_t (&v)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
               "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
// erf(u) and exp(-u^2) together (Abramowitz & Stegun 7.1.26, |error| <= 1.5e-7): two MUFU ops and ~10 FMA-pipe instructions,
// about half of erff(); the bf16 GEMM epilogues evaluate GELU / GELU' for every element of the MLP hidden activation
// with it.  (The fp32 verification path keeps erff.)
__device__ __forceinline__ void erf_exp_fast(float u, float& erf_u, float& exp_mu2) {
  const float a = fabsf(u);
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, a, 1.0f)));
  float pl = fmaf(1.061405429f, t, -1.453152027f);
  pl = fmaf(pl, t, 1.421413741f);
  pl = fmaf(pl, t, -0.284496736f);
  pl = fmaf(pl, t, 0.254829592f);
  pl *= t;
  exp_mu2 = ex2_approx(-1.4426950408889634f * a * a);
  erf_u = copysignf(fmaf(-pl, exp_mu2, 1.0f), u);
}
__device__ __forceinline__ float gelu_fast(float x) {
  float e, ex;
  erf_exp_fast(x * 0.70710678118654752f, e, ex);
  return 0.5f * x * (1.0f + e);
}
__device__ __forceinline__ float gelu_grad_fast(float x) {     // Phi(x) + x phi(x); exp(-x^2/2) is shared with the erf evaluation
  float e, ex;
  erf_exp_fast(x * 0.70710678118654752f, e, ex);
  return fmaf(0.3989422804014327f * x, ex, fmaf(0.5f, e, 0.5f));
}
// The same GELU / GELU' on a PAIR of values with packed fp32 arithmetic (FFMA2 / FMUL2: one issue slot per two elements) and without the
// sign handling of erf: with h(x) = 1/2 pl(t) exp(-x^2/2), t = 1 / (1 + 0.3275911 |x| / sqrt 2)  (Abramowitz & Stegun 7.1.26),
//     Phi(x) = x >= 0 ? 1 - h : h,        GELU(x) = relu(x) - |x| h,        GELU'(x) = Phi(x) + x exp(-x^2/2) / sqrt(2 pi).
// ~10 issue slots per element instead of ~19: the GEMMs with a GELU epilogue were bound by the epilogue warps' instruction issue
// (profiles/r02_gemm_time_by_shape.log: 429 / 468 us against 160-240 us for the same shapes without the activation).
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void gelu_h_exp2(float2 x, float2& h, float2& e) {
  const float2 x2 = __fmul2_rn(x, x);
  const float2 arg = __fmul2_rn(x2, make_float2(-0.72134752044448170f, -0.72134752044448170f));       // -x^2/2 * log2(e)
  e = make_float2(ex2_approx(arg.x), ex2_approx(arg.y));
  const float2 t = make_float2(rcp_approx(fmaf(0.23164189f, fabsf(x.x), 1.0f)), rcp_approx(fmaf(0.23164189f, fabsf(x.y), 1.0f)));
  // 1/2 * (((((a5 t + a4) t + a3) t + a2) t + a1) t): the 1/2 is folded into the coefficients
  float2 pl = __ffma2_rn(t, make_float2(0.5307027145f, 0.5307027145f), make_float2(-0.7265760135f, -0.7265760135f));
  pl = __ffma2_rn(pl, t, make_float2(0.7107068705f, 0.7107068705f));
  pl = __ffma2_rn(pl, t, make_float2(-0.142248368f, -0.142248368f));
  pl = __ffma2_rn(pl, t, make_float2(0.127414796f, 0.127414796f));
  h = __fmul2_rn(__fmul2_rn(pl, t), e);
}
__device__ __forceinline__ float2 gelu_fast2(float2 x) {
  float2 h, e;
  gelu_h_exp2(x, h, e);
  return make_float2(fmaf(-fabsf(x.x), h.x, fmaxf(x.x, 0.f)), fmaf(-fabsf(x.y), h.y, fmaxf(x.y, 0.f)));
}
__device__ __forceinline__ float2 gelu_grad_fast2(float2 x) {
  float2 h, e;
  gelu_h_exp2(x, h, e);
  // Phi = h + [x >= 0] (1 - 2 h);   + x e / sqrt(2 pi)
  const float2 one_m2h = __ffma2_rn(h, make_float2(-2.f, -2.f), make_float2(1.f, 1.f));
  const float2 xe = __fmul2_rn(__fmul2_rn(x, e), make_float2(0.3989422804014327f, 0.3989422804014327f));
  const float kx = x.x >= 0.f ? 1.f : 0.f, ky = x.y >= 0.f ? 1.f : 0.f;
  return __fadd2_rn(__ffma2_rn(make_float2(kx, ky), one_m2h, h), xe);
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}
// packed fp32 pairs (FFMA2 / FADD2 / FMUL2 on sm_100): half the issue slots of the scalar forms
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
// 2^a for a pair of arguments on the FMA/ALU pipes instead of MUFU (the exp unit runs 16/clk/SM, the bound of the
// attention kernels' softmax stages at head_dim <= 64): round-to-nearest split a = n + f via the 1.5*2^23 magic constant, degree-3 minimax polynomial of
// 2^f on [-0.5, 0.5] (max relative error 7.5e-5, far below the bf16 rounding of P), then n is added into the exponent
// field.  a <= ~8 by construction (stale maxima are bounded by the lazy-rescale threshold); a is clamped at -125.
__device__ __forceinline__ float2 ex2_poly2(float2 a) {
  a.x = fmaxf(a.x, -125.f);
  a.y = fmaxf(a.y, -125.f);
  const float2 t = __fadd2_rn(a, make_float2(12582912.f, 12582912.f));
  const float2 n = __fadd2_rn(t, make_float2(-12582912.f, -12582912.f));
  const float2 f = __ffma2_rn(n, make_float2(-1.f, -1.f), a);
  float2 p = __ffma2_rn(f, make_float2(0.05517132207751274f, 0.05517132207751274f), make_float2(0.24261054396629333f, 0.24261054396629333f));
  p = __ffma2_rn(p, f, make_float2(0.6932609677314758f, 0.6932609677314758f));
  p = __ffma2_rn(p, f, make_float2(0.9999281167984009f, 0.9999281167984009f));
  return make_float2(__uint_as_float(__float_as_uint(p.x) + (__float_as_uint(t.x) << 23)),
                     __uint_as_float(__float_as_uint(p.y) + (__float_as_uint(t.y) << 23)));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------ dropout masks (counter-based, layout-free)
// nn.Dropout on the hot path (vit_components.py:27-29,76-78; hybrid_vit_backbone.py:78,80) is applied inside the
// kernels.  keep(row, col) is a pure function of (seed words, site, row, col), so forward and backward kernels with
// different thread layouts regenerate the same mask and nothing of mask size is stored.  The seed is two 32-bit
// words drawn by the caller from torch's CUDA generator (replayed by torch.utils.checkpoint) and read from device
// memory, so no host synchronisation is needed.  oracle/dropout_mask.py restates it for the parity tests.
//
//   keep(row, col)  <=>  (rowkey(row) * colmul(col)) mod 2^32 >= thr          (thr = p * 2^32; decision in the TOP bits of the product)
//   rowkey(row)  = three rounds of a 32x32->64 multiply folded by xor ("mum") over (row, seed, site), forced odd
//   colmul(col)  = blockmul(col / 128) * inblockmul(col % 128) mod 2^32, both odd outputs of a 32-bit finaliser
//
// Round 2: the per-element work is ONE 32-bit multiply and one compare.  A thread that owns a row folds blockmul into its key once per
// 128-column tile and multiplies by the compile-time constants inblockmul(c) (attention forward, GEMM epilogues); a thread that owns a
// column multiplies the row keys by its own colmul (attention backward).  Round 1 hashed (rowkey ^ colterm) * M: an extra xor per
// element on the ALU pipe, which the mask work saturates (profiles/r02_maskbench_dropout_sequences.log: 6.06 -> 4.13 clk per element
// and SM sub-partition with the keep decision applied as a predicated negate on the FMA pipe instead of a select).  Two columns of one
// row are related by a fixed odd ratio B2/B1; for the 128 constants below (and their products with the block multipliers) the joint
// keep statistics of all pairs are indistinguishable from independent draws (tests/test_dropout_mask_cpu.py: every pair within 5 sigma
// over 2^20 rows, per-row counts binomial).
struct DropCfg {
  uint32_t k0, k1, site, thr;   // drop element iff hash < thr (thr = p * 2^32)
  float inv_keep;               // 1 / (1 - p)
};
__host__ __device__ __forceinline__ uint32_t mum32(uint32_t a, uint32_t b) {
  const uint64_t w = static_cast<uint64_t>(a) * b;
  return static_cast<uint32_t>(w) ^ static_cast<uint32_t>(w >> 32);
}
__device__ __forceinline__ uint32_t drop_rowkey(const DropCfg& c, uint32_t row) {
  uint32_t h = mum32(row ^ c.k0, 0x9E3779B1u);
  h = mum32(h ^ c.site ^ c.k1, 0x85EBCA77u);
  return mum32(h + 0x6A09E667u, 0xC2B2AE3Du) | 1u;      // odd: multiplication by it is a bijection of the 32-bit words
}
__host__ __device__ constexpr uint32_t drop_mix32(uint32_t c) {
  uint32_t x = (c + 1u) * 0x9E3779B1u;
  x ^= x >> 15; x *= 0x85EBCA77u;
  x ^= x >> 13; x *= 0xC2B2AE3Du;
  x ^= x >> 16;
  return x;
}
__host__ __device__ constexpr uint32_t drop_inblock_mul(uint32_t col_in_block) { return drop_mix32(col_in_block) | 1u; }
__host__ __device__ constexpr uint32_t drop_block_mul(uint32_t block) { return drop_mix32(block ^ 0x5BD1E995u) | 1u; }
__host__ __device__ constexpr uint32_t drop_colmul(uint32_t col) { return drop_block_mul(col >> 7) * drop_inblock_mul(col & 127u); }
// rowkey * blockmul once per (row, 128-column block), then per element one multiply by the column's in-block constant
__device__ __forceinline__ uint32_t drop_blockkey(uint32_t rowkey, uint32_t col) { return rowkey * drop_block_mul(col >> 7); }
__device__ __forceinline__ uint32_t drop_hash_in_block(uint32_t blockkey, uint32_t col_in_block) { return blockkey * drop_inblock_mul(col_in_block); }
__device__ __forceinline__ bool drop_keep_in_block(uint32_t blockkey, uint32_t col_in_block, uint32_t thr) {
  return drop_hash_in_block(blockkey, col_in_block) >= thr;
}
__device__ __forceinline__ uint32_t drop_hash(uint32_t rowkey, uint32_t col) { return rowkey * drop_colmul(col); }
__device__ __forceinline__ bool drop_keep(uint32_t rowkey, uint32_t col, uint32_t thr) { return drop_hash(rowkey, col) >= thr; }
// The attention kernels carry the decision in the SIGN of the (non-negative) probability: dropped -> negated.  One compare (ALU pipe) and
// a predicated negate (FMA pipe); the bf16 pack with relu then zeroes the dropped entries for free.
__device__ __forceinline__ float drop_negate_if_dropped(float e, uint32_t hash, uint32_t thr) {
  asm("{\n\t.reg .pred p;\n\tsetp.lo.u32 p, %1, %2;\n\t@p neg.f32 %0, %0;\n\t}" : "+f"(e) : "r"(hash), "r"(thr));
  return e;
}
// host-side description -> kernel config (reads the seed words on the device)
struct DropArg {
  const uint32_t* seed; uint32_t site; uint32_t thr; float inv_keep;   // seed == nullptr: disabled
};
// hvc_dropout (C ABI) -> DropArg; disabled when seed is NULL or p <= 0
inline DropArg make_drop(const hvc_dropout& d) {
  DropArg r;
  const bool on = d.seed != nullptr && d.p > 0.f;
  r.seed = on ? d.seed : nullptr;
  r.site = d.site;
  double t = on ? (double)d.p * 4294967296.0 : 0.0;
  if (t > 4294967295.0) t = 4294967295.0;
  r.thr = (uint32_t)t;
  r.inv_keep = on ? (float)(1.0 / (1.0 - (double)r.thr / 4294967296.0)) : 1.f;
  return r;
}
__device__ __forceinline__ DropCfg drop_load(const DropArg& a) {
  DropCfg c;
  c.k0 = __ldg(a.seed); c.k1 = __ldg(a.seed + 1); c.site = a.site; c.thr = a.thr; c.inv_keep = a.inv_keep;
  return c;
}

// byte offset of 16-byte chunk `chunk` (0..7) of row `row` inside a 128B-swizzled tile whose rows are 128 B
__device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t chunk) {
  return row * 128u + ((chunk ^ (row & 7u)) << 4);
}

}  // namespace hvc
