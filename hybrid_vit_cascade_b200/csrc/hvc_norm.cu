// hvc_norm.cu -- HBM-bound fused kernels around the GEMMs / attention of one HybridViTBlock3D:
//   * LayerNorm (+AdaLN modulate) forward  -> bf16 GEMM operand     (hybrid_vit_backbone.py:120-121,126,136-137)
//   * its backward (dx + per-batch column sums for dw/db/dshift/dscale)
//   * backward of the gated residual  x += gate * branch            (hybrid_vit_backbone.py:123,128,139)
//   * bias-gradient column sums, fp32 -> bf16 casts (weights, context tokens)
//   * AdaLN modulation linear on the tiny (B, cond) input           (vit_components.py:144)
// All are one-pass over their [T, C] operands with 128-bit accesses; roofline = HBM bandwidth.
// Row layout: one warp per token row, lane l owns columns {4*(l + 32*i) .. +3}, i < VPL.
#include <stdlib.h>

#include "hvc_common.cuh"
#include "hvc_host.h"

namespace hvc {

constexpr float kLnEps = 1e-5f;

template <int VPL>
__device__ __forceinline__ void load_row_f32(const float* __restrict__ p, int C, int lane, float4 (&v)[VPL]) {
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = 4 * (lane + 32 * i);
    v[i] = (c < C) ? __ldg(reinterpret_cast<const float4*>(p + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
template <int VPL>
__device__ __forceinline__ void load_row_bf16(const bf16* __restrict__ p, int C, int lane, float4 (&v)[VPL]) {
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = 4 * (lane + 32 * i);
    if (c < C) {
      const uint2 u = __ldg(reinterpret_cast<const uint2*>(p + c));
      const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
      v[i] = make_float4(a.x, a.y, b.x, b.y);
    } else {
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}
__device__ __forceinline__ float4 ld4(const float* p, int c, int C) {
  return (c < C) ? __ldg(reinterpret_cast<const float4*>(p + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
}

// ------------------------------------------------------------------ LayerNorm (+modulate) forward
struct LnFwdArgs {
  const float* x; long long ldx;
  const float* w; const float* b;
  const float* shift; const float* scale; long long mod_ld; int rows_per_batch;
  void* y; long long ldy; int y_bf16;
  float* mean; float* rstd;
  int T, C;
};

template <int VPL>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const LnFwdArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= a.T) return;
  float4 v[VPL];
  load_row_f32<VPL>(a.x + (long long)row * a.ldx, a.C, lane, v);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) s += v[i].x + v[i].y + v[i].z + v[i].w;
  const float mean = warp_sum(s) / a.C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = 4 * (lane + 32 * i);
    if (c < a.C) {
      const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
      q += dx * dx + dy * dy + dz * dz + dw * dw;
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / a.C + kLnEps);
  if (lane == 0) {
    if (a.mean) a.mean[row] = mean;
    if (a.rstd) a.rstd[row] = rstd;
  }
  const long long mrow = a.scale ? (long long)(row / a.rows_per_batch) * a.mod_ld : 0;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = 4 * (lane + 32 * i);
    if (c >= a.C) continue;
    const float4 w = ld4(a.w, c, a.C), b = ld4(a.b, c, a.C);
    float4 y;
    y.x = (v[i].x - mean) * rstd * w.x + b.x;
    y.y = (v[i].y - mean) * rstd * w.y + b.y;
    y.z = (v[i].z - mean) * rstd * w.z + b.z;
    y.w = (v[i].w - mean) * rstd * w.w + b.w;
    if (a.scale) {
      const float4 sc = ld4(a.scale + mrow, c, a.C), sh = ld4(a.shift + mrow, c, a.C);
      y.x = fmaf(y.x, 1.f + sc.x, sh.x);
      y.y = fmaf(y.y, 1.f + sc.y, sh.y);
      y.z = fmaf(y.z, 1.f + sc.z, sh.z);
      y.w = fmaf(y.w, 1.f + sc.w, sh.w);
    }
    if (a.y_bf16) {
      uint2 u = make_uint2(pack_bf16(y.x, y.y), pack_bf16(y.z, y.w));
      *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(a.y) + (long long)row * a.ldy + c) = u;
    } else {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(a.y) + (long long)row * a.ldy + c) = y;
    }
  }
}

// ------------------------------------------------------------------ LayerNorm backward
// Upstream gradient dz (w.r.t. the modulated output), three sources:
//   dz_bf16 [T,C]  |  dz_f32 [T,C]  |  dz_row [T] (one scalar per row: the output head, dz_c = dv)
// Column multiplier m_c for the LN-output gradient: mult_batch (1 + scale[b,c]) | mult_vec (wo[c]) | 1.
//   dy_c = dz_c * m_c ; dxhat_c = dy_c * w_c ; dx = rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat*xhat))
// Per-batch column sums S1[b,c] = sum_n dz_c, S2[b,c] = sum_n dz_c * xhat_c are accumulated with
// atomics; ln_bwd_finalize turns them into dw/db/dshift/dscale (or dwo/dbo for the head).
struct LnBwdArgs {
  const bf16* dz_bf16; const float* dz_f32; long long lddz;
  const float* dz_row;
  const float* x; long long ldx;
  const float* mean; const float* rstd;
  const float* w;
  const float* scale; long long mod_ld;   // per-batch (1+scale) multiplier, or null
  const float* mult_vec;                  // shared multiplier, or null
  const float* dx_in; long long lddxi;    // optional residual-stream gradient to add
  float* dx; long long lddx;
  float* S1; float* S2;                   // [B, C]
  int rows_per_batch, rows_per_block, C;
};

template <int VPL>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const LnBwdArgs a) {
  __shared__ float red[8][VPL * 128 + 4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int r0 = blockIdx.x * a.rows_per_block;
  const int r1 = min(r0 + a.rows_per_block, a.rows_per_batch);
  float4 s1[VPL], s2[VPL], wv[VPL], mv[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = 4 * (lane + 32 * i);
    s1[i] = s2[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    wv[i] = ld4(a.w, c, a.C);
    if (a.scale) {
      const float4 sc = ld4(a.scale + (long long)b * a.mod_ld, c, a.C);
      mv[i] = make_float4(1.f + sc.x, 1.f + sc.y, 1.f + sc.z, 1.f + sc.w);
    } else if (a.mult_vec) {
      mv[i] = ld4(a.mult_vec, c, a.C);
    } else {
      mv[i] = make_float4(1.f, 1.f, 1.f, 1.f);
    }
  }
  const float invC = 1.0f / a.C;
  // Every operand of a row (x, dz, the residual-stream gradient dx_in, mean, rstd) is requested before any of it is used: round 1 fetched
  // dx_in after the two warp reductions, which put a second exposed memory latency on every row (4.1 TB/s = 0.62 of the copy bandwidth).
  for (int r = r0 + warp; r < r1; r += 8) {
    const long long row = (long long)b * a.rows_per_batch + r;
    float4 xv[VPL], dz[VPL], ev[VPL];
    load_row_f32<VPL>(a.x + row * a.ldx, a.C, lane, xv);
    if (a.dx_in) load_row_f32<VPL>(a.dx_in + row * a.lddxi, a.C, lane, ev);
    if (a.dz_bf16) {
      load_row_bf16<VPL>(a.dz_bf16 + row * a.lddz, a.C, lane, dz);
    } else if (a.dz_f32) {
      load_row_f32<VPL>(a.dz_f32 + row * a.lddz, a.C, lane, dz);
    } else {
      const float d = __ldg(a.dz_row + row);
#pragma unroll
      for (int i = 0; i < VPL; ++i) dz[i] = (4 * (lane + 32 * i) < a.C) ? make_float4(d, d, d, d) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float mean = __ldg(a.mean + row), rstd = __ldg(a.rstd + row);
    float t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const bool in = 4 * (lane + 32 * i) < a.C;
      // xhat in xv, dxhat in dz (after accumulating the column sums)
      xv[i].x = in ? (xv[i].x - mean) * rstd : 0.f;
      xv[i].y = in ? (xv[i].y - mean) * rstd : 0.f;
      xv[i].z = in ? (xv[i].z - mean) * rstd : 0.f;
      xv[i].w = in ? (xv[i].w - mean) * rstd : 0.f;
      s1[i].x += dz[i].x; s1[i].y += dz[i].y; s1[i].z += dz[i].z; s1[i].w += dz[i].w;
      s2[i].x += dz[i].x * xv[i].x; s2[i].y += dz[i].y * xv[i].y;
      s2[i].z += dz[i].z * xv[i].z; s2[i].w += dz[i].w * xv[i].w;
      dz[i].x *= mv[i].x * wv[i].x; dz[i].y *= mv[i].y * wv[i].y;
      dz[i].z *= mv[i].z * wv[i].z; dz[i].w *= mv[i].w * wv[i].w;
      t1 += dz[i].x + dz[i].y + dz[i].z + dz[i].w;
      t2 += dz[i].x * xv[i].x + dz[i].y * xv[i].y + dz[i].z * xv[i].z + dz[i].w * xv[i].w;
    }
    t1 = warp_sum(t1) * invC;
    t2 = warp_sum(t2) * invC;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = 4 * (lane + 32 * i);
      if (c >= a.C) continue;
      float4 g;
      g.x = rstd * (dz[i].x - t1 - xv[i].x * t2);
      g.y = rstd * (dz[i].y - t1 - xv[i].y * t2);
      g.z = rstd * (dz[i].z - t1 - xv[i].z * t2);
      g.w = rstd * (dz[i].w - t1 - xv[i].w * t2);
      if (a.dx_in) { g.x += ev[i].x; g.y += ev[i].y; g.z += ev[i].z; g.w += ev[i].w; }
      *reinterpret_cast<float4*>(a.dx + row * a.lddx + c) = g;
    }
  }
  // cross-warp reduction of the column sums, then one atomic per column per block
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const float4 v = pass == 0 ? s1[i] : s2[i];
      *reinterpret_cast<float4*>(&red[warp][4 * (lane + 32 * i)]) = v;
    }
    __syncthreads();
    float* dst = (pass == 0 ? a.S1 : a.S2) + (long long)b * a.C;
    for (int c = threadIdx.x; c < a.C; c += 256) {
      float t = 0.f;
#pragma unroll
      for (int wq = 0; wq < 8; ++wq) t += red[wq][c];
      atomicAdd(dst + c, t);
    }
  }
}

// dw += sum_b m_b * S2_b ; db += sum_b m_b * S1_b ; dshift_b = S1_b ; dscale_b = w*S2_b + b*S1_b
// head mode (mult_vec = wo): dw = wo*S2, db = wo*S1, dwo = w*S2 + b*S1, dbo = S1[0]
struct LnFinArgs {
  const float* S1; const float* S2;
  const float* w; const float* b;
  const float* scale; long long mod_ld;
  const float* mult_vec;
  float* dw; float* db;
  float* dshift; float* dscale; long long dmod_ld;
  float* dvec; float* dscalar;
  int B, C;
};
__global__ void ln_bwd_finalize_kernel(const LnFinArgs a) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.C) return;
  const float w = a.w[c], bb = a.b[c];
  float dw = 0.f, db = 0.f, v1 = 0.f, v2 = 0.f;
  for (int b = 0; b < a.B; ++b) {
    const float s1 = a.S1[(long long)b * a.C + c], s2 = a.S2[(long long)b * a.C + c];
    float m = 1.f;
    if (a.scale) m = 1.f + a.scale[(long long)b * a.mod_ld + c];
    else if (a.mult_vec) m = a.mult_vec[c];
    dw += m * s2;
    db += m * s1;
    v1 += s1;
    v2 += s2;
    if (a.dshift) a.dshift[(long long)b * a.dmod_ld + c] = s1;
    if (a.dscale) a.dscale[(long long)b * a.dmod_ld + c] = w * s2 + bb * s1;
  }
  if (a.dw) a.dw[c] = dw;
  if (a.db) a.db[c] = db;
  if (a.dvec) a.dvec[c] = w * v2 + bb * v1;
  if (a.dscalar && c == 0) a.dscalar[0] = v1;
}

// ------------------------------------------------------------------ gated residual backward
// out = resid + gate_b * branch  =>  dbranch (bf16) = gate_b * dout ; D1[b,c] = sum_n dout ; D2[b,c] = sum_n dout*branch
struct ResidBwdArgs {
  const float* dout; long long lddo;
  const bf16* branch; long long ldbr;
  const float* gate; long long gate_ld;
  bf16* dbranch; long long lddb;
  float* D1; float* D2;
  int rows_per_batch, rows_per_block, C;
  DropArg drop;   // the dropout applied to branch in the forward epilogue (seed == nullptr: none)
};
template <int VPL>
__global__ void __launch_bounds__(256) resid_bwd_kernel(const ResidBwdArgs a) {
  __shared__ float red[8][VPL * 128 + 4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int r0 = blockIdx.x * a.rows_per_block;
  const int r1 = min(r0 + a.rows_per_block, a.rows_per_batch);
  float4 s1[VPL], s2[VPL], gv[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    s1[i] = s2[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    gv[i] = a.gate ? ld4(a.gate + (long long)b * a.gate_ld, 4 * (lane + 32 * i), a.C) : make_float4(1.f, 1.f, 1.f, 1.f);
  }
  uint4 cm[VPL];      // dropout: the column multipliers of this thread's four columns per vector (the same for every row)
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const uint32_t c = 4u * (lane + 32 * i);
    cm[i] = make_uint4(drop_colmul(c), drop_colmul(c + 1), drop_colmul(c + 2), drop_colmul(c + 3));
  }
  for (int r = r0 + warp; r < r1; r += 8) {
    const long long row = (long long)b * a.rows_per_batch + r;
    float4 d[VPL], br[VPL];
    load_row_f32<VPL>(a.dout + row * a.lddo, a.C, lane, d);
    if (a.branch) load_row_bf16<VPL>(a.branch + row * a.ldbr, a.C, lane, br);
    uint32_t rk = 0;
    if (a.drop.seed != nullptr) rk = drop_rowkey(drop_load(a.drop), static_cast<uint32_t>(row));
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = 4 * (lane + 32 * i);
      if (a.branch) {     // dgate: branch already is the dropped (and rescaled) value, so the raw gradient is used here
        s2[i].x += d[i].x * br[i].x; s2[i].y += d[i].y * br[i].y;
        s2[i].z += d[i].z * br[i].z; s2[i].w += d[i].w * br[i].w;
      }
      if (a.drop.seed != nullptr) {   // d(pre-dropout branch) = mask / (1-p) * gate * dout; cm = this thread's column multipliers
        d[i].x = rk * cm[i].x >= a.drop.thr ? d[i].x * a.drop.inv_keep : 0.f;
        d[i].y = rk * cm[i].y >= a.drop.thr ? d[i].y * a.drop.inv_keep : 0.f;
        d[i].z = rk * cm[i].z >= a.drop.thr ? d[i].z * a.drop.inv_keep : 0.f;
        d[i].w = rk * cm[i].w >= a.drop.thr ? d[i].w * a.drop.inv_keep : 0.f;
      }
      s1[i].x += d[i].x; s1[i].y += d[i].y; s1[i].z += d[i].z; s1[i].w += d[i].w;
      if (c < a.C) {
        uint2 u = make_uint2(pack_bf16(d[i].x * gv[i].x, d[i].y * gv[i].y), pack_bf16(d[i].z * gv[i].z, d[i].w * gv[i].w));
        *reinterpret_cast<uint2*>(a.dbranch + row * a.lddb + c) = u;
      }
    }
  }
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    if (pass == 1 && a.D2 == nullptr) break;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const float4 v = pass == 0 ? s1[i] : s2[i];
      *reinterpret_cast<float4*>(&red[warp][4 * (lane + 32 * i)]) = v;
    }
    __syncthreads();
    float* dst = (pass == 0 ? a.D1 : a.D2) + (long long)b * a.C;
    for (int c = threadIdx.x; c < a.C; c += 256) {
      float t = 0.f;
#pragma unroll
      for (int wq = 0; wq < 8; ++wq) t += red[wq][c];
      atomicAdd(dst + c, t);
    }
  }
}
// dbias[c] = sum_b gate_b[c] * D1[b,c]
__global__ void resid_bwd_finalize_kernel(const float* D1, const float* gate, long long gate_ld, float* dbias, int B, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float t = 0.f;
  for (int b = 0; b < B; ++b) t += (gate ? gate[(long long)b * gate_ld + c] : 1.f) * D1[(long long)b * C + c];
  dbias[c] = t;
}

// ------------------------------------------------------------------ column sums of a bf16 matrix (bias grads)
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const bf16* __restrict__ x, long long ldx, int T, int C,
                                                          int rows_per_block, float* __restrict__ out) {
  // thread owns 2 adjacent columns; block covers 512 columns x rows_per_block rows
  const int c = (blockIdx.y * 256 + threadIdx.x) * 2;
  if (c >= C) return;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(r0 + rows_per_block, T);
  float s0 = 0.f, s1 = 0.f;
  for (int r = r0; r < r1; ++r) {
    const float2 v = unpack_bf16(__ldg(reinterpret_cast<const uint32_t*>(x + (long long)r * ldx + c)));
    s0 += v.x;
    s1 += v.y;
  }
  atomicAdd(out + c, s0);
  atomicAdd(out + c + 1, s1);
}

// ------------------------------------------------------------------ casts
__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ y, long long n) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(x + i)), b = __ldg(reinterpret_cast<const float4*>(x + i + 4));
    uint4 u = make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
    *reinterpret_cast<uint4*>(y + i) = u;
  } else {
    for (long long j = i; j < n; ++j) y[j] = __float2bfloat16(x[j]);
  }
}
// x[b, m, c] with arbitrary element strides (sb, sm, sc) -> y bf16 [B*M, C] contiguous (tile transpose through smem
// so that both the read (along whichever of m/c is contiguous) and the write are coalesced).
template <typename TIn>
__global__ void __launch_bounds__(256) cast_tokens_kernel(const TIn* __restrict__ x, long long sb, long long sm, long long sc,
                                                          bf16* __restrict__ y, int M, int C) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, m0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const TIn* xb = x + (long long)b * sb;
  if (sm <= sc) {  // m is the contiguous-ish axis of the input: read with tx along m
    for (int j = ty; j < 32; j += 8) {
      const int m = m0 + tx, c = c0 + j;
      tile[j][tx] = (m < M && c < C) ? static_cast<float>(xb[(long long)m * sm + (long long)c * sc]) : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
      const int m = m0 + j, c = c0 + tx;
      if (m < M && c < C) y[((long long)b * M + m) * C + c] = __float2bfloat16(tile[tx][j]);
    }
  } else {
    for (int j = ty; j < 32; j += 8) {
      const int m = m0 + j, c = c0 + tx;
      if (m < M && c < C) y[((long long)b * M + m) * C + c] = __float2bfloat16(static_cast<float>(xb[(long long)m * sm + (long long)c * sc]));
    }
  }
}

// ------------------------------------------------------------------ AdaLN modulation linear (tiny)
// out[b, j] = bias[j] + sum_k cond[b,k] * W[j,k]      one warp per output column j, all batch rows at once
__global__ void __launch_bounds__(256) adaln_fwd_kernel(const float* __restrict__ cond, long long ldc, const float* __restrict__ W,
                                                        const float* __restrict__ bias, float* __restrict__ out, int B, int K, int J) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 8 + warp;
  if (j >= J) return;
  for (int b0 = 0; b0 < B; b0 += 8) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k = lane; k < K; k += 32) {
      const float w = __ldg(W + (long long)j * K + k);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (b0 + i < B) acc[i] = fmaf(w, __ldg(cond + (long long)(b0 + i) * ldc + k), acc[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float t = warp_sum(acc[i]);
      if (lane == 0 && b0 + i < B) out[(long long)(b0 + i) * J + j] = t + (bias ? bias[j] : 0.f);
    }
  }
}
// dW[j,k] = sum_b dp[b,j] cond[b,k]; db[j] = sum_b dp[b,j]
__global__ void adaln_wgrad_kernel(const float* __restrict__ dp, const float* __restrict__ cond, long long ldc, float* __restrict__ dW,
                                   float* __restrict__ db, int B, int K, int J) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (k >= K) return;
  float t = 0.f, s = 0.f;
  for (int b = 0; b < B; ++b) {
    const float d = __ldg(dp + (long long)b * J + j);
    t = fmaf(d, __ldg(cond + (long long)b * ldc + k), t);
    s += d;
  }
  dW[(long long)j * K + k] = t;
  if (k == 0 && db) db[j] = s;
}
// dcond[b,k] += sum_{j in this CTA's slab} dp[b,j] W[j,k]: W is read once (each CTA streams a slab of rows for every batch row,
// up to 8 at a time in registers) instead of once per batch row by a handful of CTAs; dcond is zero-filled by the caller.
__global__ void __launch_bounds__(128) adaln_dgrad_kernel(const float* __restrict__ dp, const float* __restrict__ W, float* __restrict__ dcond,
                                                          int B, int K, int J, int rows_per_cta) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int j0 = blockIdx.y * rows_per_cta, j1 = min(j0 + rows_per_cta, J);
  if (k >= K) return;
  for (int b0 = 0; b0 < B; b0 += 8) {
    float t[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int nb = min(8, B - b0);
    for (int j = j0; j < j1; ++j) {
      const float w = __ldg(W + (long long)j * K + k);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (i < nb) t[i] = fmaf(__ldg(dp + (long long)(b0 + i) * J + j), w, t[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < nb) atomicAdd(dcond + (long long)(b0 + i) * K + k, t[i]);
  }
}

// rows per block of the backward kernels before the grid-size rule below shrinks it (HVC_NORM_RPB: tuning only)
static int hvc_rows_per_block() {
  static int v = 0;
  if (v == 0) { const char* e = getenv("HVC_NORM_RPB"); v = e ? atoi(e) : 64; if (v < 8) v = 8; }
  return v;
}
static int pick_vpl(int C) { return C <= 128 ? 1 : C <= 256 ? 2 : C <= 512 ? 4 : C <= 1024 ? 8 : 0; }

}  // namespace hvc

using namespace hvc;

#define HVC_DISPATCH_VPL(vpl, CALL) \
  switch (vpl) {                    \
    case 1: { constexpr int VPL = 1; CALL; } break; \
    case 2: { constexpr int VPL = 2; CALL; } break; \
    case 4: { constexpr int VPL = 4; CALL; } break; \
    default: { constexpr int VPL = 8; CALL; } break; \
  }

extern "C" int hvc_ln_fwd(const hvc_ln_args* a, void* stream) {
  HVC_CHECK_ARG(a && a->size == sizeof(hvc_ln_args), "hvc_ln_fwd: bad args struct");
  HVC_CHECK_ARG(a->T > 0 && a->C > 0 && (a->C & 3) == 0 && a->C <= 1024, "hvc_ln_fwd: C=%d must be a multiple of 4, <= 1024", a->C);
  HVC_CHECK_ARG(a->x && a->w && a->b && a->y, "hvc_ln_fwd: null operand");
  HVC_CHECK_ARG((a->shift == nullptr) == (a->scale == nullptr), "hvc_ln_fwd: shift and scale go together");
  HVC_CHECK_ARG(a->scale == nullptr || a->rows_per_batch > 0, "hvc_ln_fwd: modulation needs rows_per_batch");
  LnFwdArgs k;
  k.x = a->x; k.ldx = a->ldx; k.w = a->w; k.b = a->b; k.shift = a->shift; k.scale = a->scale; k.mod_ld = a->mod_ld;
  k.rows_per_batch = a->rows_per_batch > 0 ? a->rows_per_batch : a->T;
  k.y = a->y; k.ldy = a->ldy; k.y_bf16 = a->y_is_bf16; k.mean = a->mean; k.rstd = a->rstd; k.T = a->T; k.C = a->C;
  const int grid = (a->T + 7) / 8;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  HVC_DISPATCH_VPL(pick_vpl(a->C), (ln_fwd_kernel<VPL><<<grid, 256, 0, st>>>(k)));
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

extern "C" int hvc_ln_bwd(const hvc_ln_bwd_args* a, void* stream) {
  HVC_CHECK_ARG(a && a->size == sizeof(hvc_ln_bwd_args), "hvc_ln_bwd: bad args struct");
  HVC_CHECK_ARG(a->batch > 0 && a->rows_per_batch > 0 && a->C > 0 && (a->C & 3) == 0 && a->C <= 1024, "hvc_ln_bwd: bad shape");
  HVC_CHECK_ARG((a->dz_bf16 != nullptr) + (a->dz_f32 != nullptr) + (a->dz_row != nullptr) == 1, "hvc_ln_bwd: exactly one dz source");
  HVC_CHECK_ARG(a->x && a->mean && a->rstd && a->w && a->b && a->dx && a->S1 && a->S2, "hvc_ln_bwd: null operand");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  HVC_CUDA(cudaMemsetAsync(a->S1, 0, sizeof(float) * a->batch * a->C, st));
  HVC_CUDA(cudaMemsetAsync(a->S2, 0, sizeof(float) * a->batch * a->C, st));
  LnBwdArgs k;
  k.dz_bf16 = reinterpret_cast<const bf16*>(a->dz_bf16); k.dz_f32 = a->dz_f32; k.lddz = a->lddz; k.dz_row = a->dz_row;
  k.x = a->x; k.ldx = a->ldx; k.mean = a->mean; k.rstd = a->rstd; k.w = a->w;
  k.scale = a->scale; k.mod_ld = a->mod_ld; k.mult_vec = a->mult_vec;
  k.dx_in = a->dx_in; k.lddxi = a->lddx_in; k.dx = a->dx; k.lddx = a->lddx; k.S1 = a->S1; k.S2 = a->S2;
  k.rows_per_batch = a->rows_per_batch; k.C = a->C;
  // ~4 blocks per SM worth of work, but at least 8 rows (one per warp) per block
  int rpb = hvc_rows_per_block();
  while (rpb > 8 && (long long)a->batch * ((a->rows_per_batch + rpb - 1) / rpb) < 4LL * device_sm_count()) rpb >>= 1;
  k.rows_per_block = rpb;
  dim3 grid((a->rows_per_batch + rpb - 1) / rpb, a->batch);
  HVC_DISPATCH_VPL(pick_vpl(a->C), (ln_bwd_kernel<VPL><<<grid, 256, 0, st>>>(k)));
  HVC_LAUNCH_CHECK();
  LnFinArgs f;
  f.S1 = a->S1; f.S2 = a->S2; f.w = a->w; f.b = a->b; f.scale = a->scale; f.mod_ld = a->mod_ld; f.mult_vec = a->mult_vec;
  f.dw = a->dw; f.db = a->db; f.dshift = a->dshift; f.dscale = a->dscale; f.dmod_ld = a->dmod_ld;
  f.dvec = a->dvec; f.dscalar = a->dscalar; f.B = a->batch; f.C = a->C;
  ln_bwd_finalize_kernel<<<(a->C + 127) / 128, 128, 0, st>>>(f);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

extern "C" int hvc_resid_bwd(const hvc_resid_bwd_args* a, void* stream) {
  HVC_CHECK_ARG(a && a->size == sizeof(hvc_resid_bwd_args), "hvc_resid_bwd: bad args struct");
  HVC_CHECK_ARG(a->batch > 0 && a->rows_per_batch > 0 && a->C > 0 && (a->C & 3) == 0 && a->C <= 1024, "hvc_resid_bwd: bad shape");
  HVC_CHECK_ARG(a->dout && a->dbranch && a->D1, "hvc_resid_bwd: null operand");
  HVC_CHECK_ARG(a->dgate == nullptr || (a->branch != nullptr && a->gate != nullptr), "hvc_resid_bwd: dgate needs branch and gate");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  HVC_CUDA(cudaMemsetAsync(a->D1, 0, sizeof(float) * a->batch * a->C, st));
  if (a->dgate) HVC_CUDA(cudaMemsetAsync(a->dgate, 0, sizeof(float) * a->batch * a->C, st));
  ResidBwdArgs k;
  k.dout = a->dout; k.lddo = a->lddout; k.branch = a->dgate ? reinterpret_cast<const bf16*>(a->branch) : nullptr; k.ldbr = a->ldbranch;
  k.gate = a->gate; k.gate_ld = a->gate_ld; k.dbranch = reinterpret_cast<bf16*>(a->dbranch); k.lddb = a->lddbranch;
  k.D1 = a->D1; k.D2 = a->dgate; k.rows_per_batch = a->rows_per_batch; k.C = a->C;
  k.drop = make_drop(a->drop);
  HVC_CHECK_ARG(k.drop.seed == nullptr || a->drop.p < 1.f, "hvc_resid_bwd: dropout p must be < 1");
  int rpb = hvc_rows_per_block();
  while (rpb > 8 && (long long)a->batch * ((a->rows_per_batch + rpb - 1) / rpb) < 4LL * device_sm_count()) rpb >>= 1;
  k.rows_per_block = rpb;
  dim3 grid((a->rows_per_batch + rpb - 1) / rpb, a->batch);
  HVC_DISPATCH_VPL(pick_vpl(a->C), (resid_bwd_kernel<VPL><<<grid, 256, 0, st>>>(k)));
  HVC_LAUNCH_CHECK();
  if (a->dbias) {
    resid_bwd_finalize_kernel<<<(a->C + 127) / 128, 128, 0, st>>>(a->D1, a->gate, a->gate_ld, a->dbias, a->batch, a->C);
    HVC_LAUNCH_CHECK();
  }
  return HVC_OK;
}

extern "C" int hvc_colsum_bf16(const void* x, int64_t ldx, int32_t T, int32_t C, float* out, void* stream) {
  HVC_CHECK_ARG(x && out && T > 0 && C > 0 && (C & 1) == 0 && (ldx & 1) == 0, "hvc_colsum_bf16: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  HVC_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * C, st));
  int rpb = 256;
  const int cblocks = (C / 2 + 255) / 256;
  while (rpb > 16 && (long long)cblocks * ((T + rpb - 1) / rpb) < 4LL * device_sm_count()) rpb >>= 1;
  dim3 grid((T + rpb - 1) / rpb, cblocks);
  colsum_bf16_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const bf16*>(x), ldx, T, C, rpb, out);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

extern "C" int hvc_cast_bf16(const float* x, void* y, int64_t n, void* stream) {
  HVC_CHECK_ARG(x && y && n > 0, "hvc_cast_bf16: bad arguments");
  HVC_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0, "hvc_cast_bf16: unaligned");
  const long long threads = (n + 7) / 8;
  cast_f32_bf16_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, reinterpret_cast<bf16*>(y), n);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

extern "C" int hvc_cast_tokens(const void* x, int32_t x_is_bf16, int64_t sb, int64_t sm, int64_t sc, void* y, int32_t B, int32_t M,
                               int32_t C, void* stream) {
  HVC_CHECK_ARG(x && y && B > 0 && M > 0 && C > 0, "hvc_cast_tokens: bad arguments");
  dim3 grid((C + 31) / 32, (M + 31) / 32, B);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (x_is_bf16)
    cast_tokens_kernel<bf16><<<grid, 256, 0, st>>>(reinterpret_cast<const bf16*>(x), sb, sm, sc, reinterpret_cast<bf16*>(y), M, C);
  else
    cast_tokens_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(x), sb, sm, sc, reinterpret_cast<bf16*>(y), M, C);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

extern "C" int hvc_adaln_fwd(const float* cond, int64_t ldc, const float* W, const float* bias, float* out, int32_t B, int32_t K,
                             int32_t J, void* stream) {
  HVC_CHECK_ARG(cond && W && out && B > 0 && K > 0 && J > 0, "hvc_adaln_fwd: bad arguments");
  adaln_fwd_kernel<<<(J + 7) / 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(cond, ldc, W, bias, out, B, K, J);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

extern "C" int hvc_adaln_bwd(const float* dparams, const float* cond, int64_t ldc, const float* W, float* dW, float* dbias,
                             float* dcond, int32_t B, int32_t K, int32_t J, void* stream) {
  HVC_CHECK_ARG(dparams && cond && W && B > 0 && K > 0 && J > 0, "hvc_adaln_bwd: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dW) {
    adaln_wgrad_kernel<<<dim3((K + 127) / 128, J), 128, 0, st>>>(dparams, cond, ldc, dW, dbias, B, K, J);
    HVC_LAUNCH_CHECK();
  }
  if (dcond) {
    HVC_CUDA(cudaMemsetAsync(dcond, 0, sizeof(float) * B * K, st));
    const int rows_per_cta = 32;
    adaln_dgrad_kernel<<<dim3((K + 127) / 128, (J + rows_per_cta - 1) / rows_per_cta), 128, 0, st>>>(dparams, W, dcond, B, K, J, rows_per_cta);
    HVC_LAUNCH_CHECK();
  }
  return HVC_OK;
}
