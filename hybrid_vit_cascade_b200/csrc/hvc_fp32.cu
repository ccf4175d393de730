// hvc_fp32.cu -- fp32 verification mode: fp32-accurate products on the bf16 tensor cores.
//
// The hot path computes in bf16 (what torch.autocast does to the reference).  To check the kernels against the
// reference's plain fp32 modules at the 1e-4 bar (BASELINE.json north_star) the same tcgen05 GEMM
// is fed with operands split into three bf16 terms, x = x0 + x1 + x2 (8 mantissa bits each, 24 together):
//     sum_k a_k b_k  ~=  sum_k (a0 b0 + a1 b0 + a2 b0 + a0 b1 + a1 b1 + a0 b2)_k
// (the dropped terms are below 2^-24 relative).  The six partial products are laid out along K, so the product is ONE
// hvc_gemm call with K' = 6K and fp32 accumulation in TMEM:
//     A' = [A0 | A1 | A2 | A0 | A1 | A0]      B' = [B0 | B0 | B0 | B1 | B1 | B2]
// hvc_split3 writes these patterns (along columns for K-major operands, along rows for MN-major ones such as V in
// P V).  Attention in this mode materialises the score matrix per (batch, head): S' = Q' K'^T (alpha = scale*log2 e),
// hvc_softmax_rows in place, O = P' V'.  It is a verification path, not a fast one: ~6x the tensor work plus the
// (N x M) matrices in HBM.
//
// Measured on B200 (tests/bringup/split_probe.py): the tcgen05 accumulator does not round to nearest -- the error of a
// plain bf16 GEMM against the exact product grows linearly with the length of the accumulation chain (1.3e-7 at K=64,
// 1.3e-6 at K=1024, 3.9e-5 at K=32768), irrelevant next to bf16 operand rounding but not at the 1e-4 fp32 bar.  The
// verification GEMMs therefore run split-K with <= 256 K' elements per TMEM chain; the partial sums meet in fp32
// round-to-nearest atomics (HVC_EPI_F32_ATOMIC), and bias / GELU / gate / residual are applied by hvc_epilogue_f32.
#include "hvc_common.cuh"
#include "hvc_host.h"

namespace hvc {

__device__ __forceinline__ void split3(float x, bf16& p0, bf16& p1, bf16& p2) {
  p0 = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(p0);          // exact: p0 holds the leading bits of x
  p1 = __float2bfloat16_rn(r1);
  p2 = __float2bfloat16_rn(r1 - __bfloat162float(p1));
}

// pattern 0 (A side): parts {0,1,2,0,1,0}; pattern 1 (B side): parts {0,0,0,1,1,2}   (six blocks, ~2^-24)
// pattern 2 (A side): parts {0,1,0};        pattern 3 (B side): parts {0,0,1}           (three blocks: a0 b0 + a1 b0 + a0 b1, ~2^-16;
//                                                                                         the X-ray encoder's forward convolutions)
__device__ __forceinline__ int split_part(int pattern, int s) {
  const uint32_t code = pattern == 0 ? 0x012010u : pattern == 1 ? 0x000112u : pattern == 2 ? 0x010000u : 0x001000u;
  return (code >> (4 * (5 - s))) & 3;
}
__host__ __device__ __forceinline__ int split_blocks(int pattern) { return pattern < 2 ? 6 : 3; }

// x f32 [R, C] (row pitch ldx) -> out bf16; one thread = 8 consecutive columns of one row.
//   concat_rows == 0: out[r, s*C + c]      (out is [R, 6C], row pitch ldo)
//   concat_rows == 1: out[s*R + r, c]      (out is [6R, C], row pitch ldo)
__global__ void __launch_bounds__(256) split3_kernel(const float* __restrict__ x, long long ldx, int R, int C, bf16* __restrict__ out,
                                                     long long ldo, int pattern, int concat_rows) {
  const int cvecs = (C + 7) >> 3;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)R * cvecs) return;
  const int r = (int)(idx / cvecs), c0 = (int)(idx - (long long)r * cvecs) * 8;
  const int n = min(8, C - c0);
  const float* src = x + (long long)r * ldx + c0;
  uint32_t w[3][4];     // packed bf16 pairs of the three terms
#pragma unroll
  for (int j = 0; j < 8; j += 2) {
    bf16 a[3], b[3];
    split3(j < n ? src[j] : 0.f, a[0], a[1], a[2]);
    split3(j + 1 < n ? src[j + 1] : 0.f, b[0], b[1], b[2]);
#pragma unroll
    for (int t = 0; t < 3; ++t) w[t][j >> 1] = (uint32_t)__bfloat16_as_ushort(a[t]) | ((uint32_t)__bfloat16_as_ushort(b[t]) << 16);
  }
  const int nb = split_blocks(pattern);
#pragma unroll
  for (int s = 0; s < 6; ++s) {
    if (s >= nb) break;
    const int which = split_part(pattern, s);
    uint4 u;
    u.x = which == 0 ? w[0][0] : which == 1 ? w[1][0] : w[2][0];
    u.y = which == 0 ? w[0][1] : which == 1 ? w[1][1] : w[2][1];
    u.z = which == 0 ? w[0][2] : which == 1 ? w[1][2] : w[2][2];
    u.w = which == 0 ? w[0][3] : which == 1 ? w[1][3] : w[2][3];
    bf16* dst = concat_rows ? out + ((long long)s * R + r) * ldo + c0 : out + (long long)r * ldo + (long long)s * C + c0;
    if (n == 8 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
      *reinterpret_cast<uint4*>(dst) = u;
    } else {
      const uint32_t ww[4] = {u.x, u.y, u.z, u.w};
      for (int j = 0; j < n; ++j) dst[j] = __ushort_as_bfloat16((unsigned short)(ww[j >> 1] >> (16 * (j & 1))));
    }
  }
}

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, t) : v + t;
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  v = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) v = is_max ? fmaxf(v, red[w]) : v + red[w];
  return v;
}

// In place: s[r, :] <- exp2(s[r, :] - max_r) / sum_r.  s holds log2-domain scaled scores.  One CTA per row.
__global__ void __launch_bounds__(256) softmax_rows_kernel(float* __restrict__ s, long long lds, int M, float* __restrict__ lse2) {
  __shared__ float red[8];
  float* row = s + (long long)blockIdx.x * lds;
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < M; j += 256) mx = fmaxf(mx, row[j]);
  mx = block_reduce(mx, red, true);
  float sum = 0.f;
  for (int j = threadIdx.x; j < M; j += 256) sum += exp2f(row[j] - mx);
  sum = block_reduce(sum, red, false);
  const float inv = 1.0f / sum;
  for (int j = threadIdx.x; j < M; j += 256) row[j] = exp2f(row[j] - mx) * inv;
  if (lse2 != nullptr && threadIdx.x == 0) lse2[blockIdx.x] = mx + log2f(sum);
}

// out[t, n] = resid[t, n] + gate[t / rows_per_batch, n] * act(acc[t, n] + bias[n])     (each operand optional)
// The split-K partial sums of a verification-mode GEMM are reduced with fp32 atomics, so bias / GELU / gate / residual
// cannot ride in the GEMM epilogue: this pass applies them.  One thread = 4 consecutive columns.
__global__ void __launch_bounds__(256) epilogue_f32_kernel(const float* __restrict__ acc, long long lda, int T, int N,
                                                           const float* __restrict__ bias, int gelu, const float* __restrict__ resid,
                                                           long long ldr, const float* __restrict__ gate, long long gate_ld,
                                                           int rows_per_batch, float* __restrict__ out, long long ldo) {
  const int nv = N >> 2;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)T * nv) return;
  const int t = (int)(idx / nv), n = (int)(idx - (long long)t * nv) * 4;
  float4 v = *reinterpret_cast<const float4*>(acc + (long long)t * lda + n);
  if (bias) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(bias + n));
    v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
  }
  if (gelu) { v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w); }
  if (gate) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(gate + (long long)(t / rows_per_batch) * gate_ld + n));
    v.x *= g.x; v.y *= g.y; v.z *= g.z; v.w *= g.w;
  }
  if (resid) {
    const float4 r = __ldg(reinterpret_cast<const float4*>(resid + (long long)t * ldr + n));
    v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
  }
  *reinterpret_cast<float4*>(out + (long long)t * ldo + n) = v;
}

}  // namespace hvc

using namespace hvc;

extern "C" int hvc_epilogue_f32(const float* acc, int64_t lda, int32_t T, int32_t N, const float* bias, int32_t activation,
                                const float* resid, int64_t ldr, const float* gate, int64_t gate_ld, int32_t rows_per_batch,
                                float* out, int64_t ldo, void* stream) {
  HVC_CHECK_ARG(acc && out && T > 0 && N > 0, "hvc_epilogue_f32: empty or null operand");
  HVC_CHECK_ARG((N & 3) == 0 && (lda & 3) == 0 && (ldo & 3) == 0 && (resid == nullptr || (ldr & 3) == 0) &&
                    (gate == nullptr || ((gate_ld & 3) == 0 && rows_per_batch > 0)),
                "hvc_epilogue_f32: N and the row pitches must be multiples of 4");
  HVC_CHECK_ARG(activation == HVC_ACT_NONE || activation == HVC_ACT_GELU, "hvc_epilogue_f32: activation must be NONE or GELU");
  const long long threads = (long long)T * (N / 4);
  epilogue_f32_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      acc, lda, T, N, bias, activation == HVC_ACT_GELU, resid, ldr, gate, gate_ld, rows_per_batch > 0 ? rows_per_batch : 1, out, ldo);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

extern "C" int hvc_split3(const float* x, int64_t ldx, int32_t R, int32_t C, void* out, int64_t ldo, int32_t pattern,
                          int32_t concat_rows, void* stream) {
  HVC_CHECK_ARG(x && out && R > 0 && C > 0, "hvc_split3: empty or null operand");
  HVC_CHECK_ARG(pattern >= 0 && pattern <= 3, "hvc_split3: pattern must be 0/1 (six-block A/B side) or 2/3 (three-block A/B side)");
  HVC_CHECK_ARG(concat_rows ? ldo >= C : ldo >= (long long)split_blocks(pattern) * C, "hvc_split3: output pitch too small");
  const long long threads = (long long)R * ((C + 7) / 8);
  split3_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, ldx, R, C, reinterpret_cast<bf16*>(out), ldo, pattern, concat_rows);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

extern "C" int hvc_softmax_rows(float* s, int64_t lds, int32_t R, int32_t M, float* lse2, void* stream) {
  HVC_CHECK_ARG(s && R > 0 && M > 0, "hvc_softmax_rows: empty or null operand");
  softmax_rows_kernel<<<(unsigned)R, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(s, lds, M, lse2);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}
