// hvc_embed.cu -- the voxel-embedding stack of HybridViT3D (hybrid_vit_backbone.py:195-210,252):
//   Conv3d(k3, p1, stride 1|2) [+ GroupNorm(min(8,C)) + SiLU] ... -> token-major activations.
// Each conv runs as  im2col (bf16 patch matrix) -> tcgen05 GEMM (hvc_gemm, bias fused) with
// channels-last activations [B, voxels, C], so the last conv emits the (B, N, C) token layout
// directly (the reference's flatten(2).transpose(1,2), :255, costs nothing).  Backward = dgrad GEMM +
// col2im gather, wgrad GEMM (split-K).  GroupNorm+SiLU forward/backward are HBM-bound passes over the
// channels-last tensor with per-(batch, channel) partial sums reduced through shared memory.
// (Round-2 item: fold im2col into the GEMM's TMA producer -- cuTensorMapEncodeIm2col -- so the patch
// matrix never exists in HBM.)
#include "hvc_common.cuh"
#include "hvc_host.h"

namespace hvc {

struct Conv3dGeom {
  int B, Cin, D, H, W;        // input
  int Do, Ho, Wo, stride;     // output grid (k=3, pad=1)
  long long sb, sc, sd, sh, sw;  // input element strides
  int K, Kp;                  // Cin*27 and its padding to a multiple of 8
};

// cols[m, k] = x[b, cin, od*s-1+kd, oh*s-1+kh, ow*s-1+kw]   (0 outside), k = cin*27 + kd*9 + kh*3 + kw
template <typename TIn, typename TOut = bf16>
__global__ void __launch_bounds__(256) im2col3d_kernel(const TIn* __restrict__ x, TOut* __restrict__ cols, const Conv3dGeom g) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long total = (long long)g.B * g.Do * g.Ho * g.Wo * g.Kp;
  if (idx >= total) return;
  const int k = (int)(idx % g.Kp);
  long long m = idx / g.Kp;
  float v = 0.f;
  if (k < g.K) {
    const int cin = k / 27, tap = k - cin * 27;
    const int kd = tap / 9, kh = (tap - kd * 9) / 3, kw = tap - kd * 9 - kh * 3;
    const int ow = (int)(m % g.Wo); m /= g.Wo;
    const int oh = (int)(m % g.Ho); m /= g.Ho;
    const int od = (int)(m % g.Do);
    const int b = (int)(m / g.Do);
    const int id = od * g.stride - 1 + kd, ih = oh * g.stride - 1 + kh, iw = ow * g.stride - 1 + kw;
    if (id >= 0 && id < g.D && ih >= 0 && ih < g.H && iw >= 0 && iw < g.W)
      v = static_cast<float>(x[b * g.sb + cin * g.sc + id * g.sd + ih * g.sh + iw * g.sw]);
  }
  if constexpr (sizeof(TOut) == 4) cols[idx] = v;
  else cols[idx] = __float2bfloat16(v);
}

// Cin == 1 (the first conv of every stack: K = 27, Kp = 32): one thread builds one whole 64-byte patch row -- 27 reads that are
// coalesced along w across the warp, four 16-byte stores -- instead of one thread (and three divisions) per element.
template <typename TIn>
__global__ void __launch_bounds__(256) im2col3d_c1_kernel(const TIn* __restrict__ x, bf16* __restrict__ cols, const Conv3dGeom g) {
  long long m = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long total = (long long)g.B * g.Do * g.Ho * g.Wo;
  if (m >= total) return;
  uint4* dst = reinterpret_cast<uint4*>(cols + m * 32);
  const int ow = (int)(m % g.Wo); m /= g.Wo;
  const int oh = (int)(m % g.Ho); m /= g.Ho;
  const int od = (int)(m % g.Do);
  const int b = (int)(m / g.Do);
  const TIn* xb = x + b * g.sb;
  float v[28];
  v[27] = 0.f;
#pragma unroll
  for (int kd = 0; kd < 3; ++kd) {
    const int id = od * g.stride - 1 + kd;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int ih = oh * g.stride - 1 + kh;
      const bool ok = id >= 0 && id < g.D && ih >= 0 && ih < g.H;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int iw = ow * g.stride - 1 + kw;
        v[kd * 9 + kh * 3 + kw] = (ok && iw >= 0 && iw < g.W) ? static_cast<float>(xb[id * g.sd + ih * g.sh + iw * g.sw]) : 0.f;
      }
    }
  }
  uint32_t u[16];
#pragma unroll
  for (int j = 0; j < 14; ++j) u[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
  u[14] = 0u; u[15] = 0u;
#pragma unroll
  for (int j = 0; j < 4; ++j) dst[j] = make_uint4(u[4 * j], u[4 * j + 1], u[4 * j + 2], u[4 * j + 3]);
}

// ---- zero-padded channels-last volumes for the implicit-GEMM conv (hvc_conv_taps) -----------------------------------------------
// pad: src (B, D, H, W, Cs) f32|bf16, dense channels-last  ->  dst bf16 (B, D+2, H+2, W+2, Cp), Cs <= Cp, both multiples of 8; the
// border voxels and the channels [Cs, Cp) are written as zeros (no separate memset).  One thread = one 8-channel run.
template <typename TIn>
__global__ void __launch_bounds__(256) pad3d_cl_kernel(const TIn* __restrict__ src, bf16* __restrict__ dst, int B, int D, int H, int W,
                                                       int Cs, int Cp, int hi) {
  const int c8n = Cp >> 3;
  const int Dp = D + 1 + hi, Hp = H + 1 + hi, Wp = W + 1 + hi;       // one zero voxel on the low side, `hi` (0 | 1) on the high side
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long total = (long long)B * Dp * Hp * Wp * c8n;
  if (idx >= total) return;
  const int c8 = (int)(idx % c8n);
  long long t = idx / c8n;
  const int w = (int)(t % Wp) - 1; t /= Wp;
  const int h = (int)(t % Hp) - 1; t /= Hp;
  const int d = (int)(t % Dp) - 1;
  const int b = (int)(t / Dp);
  uint4 o = make_uint4(0u, 0u, 0u, 0u);
  if (w >= 0 && w < W && h >= 0 && h < H && d >= 0 && d < D && c8 * 8 < Cs) {
    const TIn* p = src + ((((long long)b * D + d) * H + h) * W + w) * Cs + c8 * 8;
    if constexpr (sizeof(TIn) == 4) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(p)), c = __ldg(reinterpret_cast<const float4*>(p) + 1);
      o = make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(c.x, c.y), pack_bf16(c.z, c.w));
    } else {
      o = __ldg(reinterpret_cast<const uint4*>(p));
    }
  }
  reinterpret_cast<uint4*>(dst)[idx] = o;
}
// unpad: src f32 (B, D+2, H+2, W+2, C) -> dst f32 (B, D, H, W, C) dense, the interior voxels.  One thread = 4 channels.
__global__ void __launch_bounds__(256) unpad3d_cl_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int D, int H, int W,
                                                         int C, int hi) {
  const int c4n = C >> 2;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long total = (long long)B * D * H * W * c4n;
  if (idx >= total) return;
  const int c4 = (int)(idx % c4n);
  long long t = idx / c4n;
  const int w = (int)(t % W); t /= W;
  const int h = (int)(t % H); t /= H;
  const int d = (int)(t % D);
  const int b = (int)(t / D);
  const long long prow = (((long long)b * (D + 1 + hi) + d + 1) * (H + 1 + hi) + h + 1) * (W + 1 + hi) + w + 1;
  reinterpret_cast<float4*>(dst)[idx] = __ldg(reinterpret_cast<const float4*>(src + prow * C) + c4);
}
// stride-2 implicit conv: split the volume by the parity of (d, h, w) into eight half-resolution volumes, each padded by one zero
// voxel on the low side and stacked: src (B, D, H, W, C) dense (D, H, W even) -> dst bf16 (8, B, D/2+1, H/2+1, W/2+1, C); parity index
// (d&1)*4 + (h&1)*2 + (w&1), voxel (d>>1, h>>1, w>>1) at padded coordinates +1.  One thread = one 8-channel run.
template <typename TIn>
__global__ void __launch_bounds__(256) s2d_pad_cl_kernel(const TIn* __restrict__ src, bf16* __restrict__ dst, int B, int D, int H, int W, int C) {
  const int c8n = C >> 3;
  const int Dp = D / 2 + 1, Hp = H / 2 + 1, Wp = W / 2 + 1;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long total = 8LL * B * Dp * Hp * Wp * c8n;
  if (idx >= total) return;
  const int c8 = (int)(idx % c8n);
  long long t = idx / c8n;
  const int wq = (int)(t % Wp); t /= Wp;
  const int hq = (int)(t % Hp); t /= Hp;
  const int dq = (int)(t % Dp); t /= Dp;
  const int b = (int)(t % B);
  const int par = (int)(t / B);
  uint4 o = make_uint4(0u, 0u, 0u, 0u);
  if (wq > 0 && hq > 0 && dq > 0) {
    const int d = 2 * (dq - 1) + (par >> 2), h = 2 * (hq - 1) + ((par >> 1) & 1), w = 2 * (wq - 1) + (par & 1);
    const TIn* p = src + ((((long long)b * D + d) * H + h) * W + w) * C + c8 * 8;
    if constexpr (sizeof(TIn) == 4) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(p)), c = __ldg(reinterpret_cast<const float4*>(p) + 1);
      o = make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(c.x, c.y), pack_bf16(c.z, c.w));
    } else {
      o = __ldg(reinterpret_cast<const uint4*>(p));
    }
  }
  reinterpret_cast<uint4*>(dst)[idx] = o;
}
// the inverse for gradients: src f32 (8, B, D/2+1, H/2+1, W/2+1, C) -> dst f32 (B, D, H, W, C) dense.  One thread = 4 channels.
__global__ void __launch_bounds__(256) d2s_unpad_cl_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int D, int H, int W,
                                                           int C) {
  const int c4n = C >> 2;
  const int Dp = D / 2 + 1, Hp = H / 2 + 1, Wp = W / 2 + 1;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long total = (long long)B * D * H * W * c4n;
  if (idx >= total) return;
  const int c4 = (int)(idx % c4n);
  long long t = idx / c4n;
  const int w = (int)(t % W); t /= W;
  const int h = (int)(t % H); t /= H;
  const int d = (int)(t % D);
  const int b = (int)(t / D);
  const int par = (d & 1) * 4 + (h & 1) * 2 + (w & 1);
  const long long prow = ((((long long)par * B + b) * Dp + (d >> 1) + 1) * Hp + (h >> 1) + 1) * Wp + (w >> 1) + 1;
  reinterpret_cast<float4*>(dst)[idx] = __ldg(reinterpret_cast<const float4*>(src + prow * C) + c4);
}

// dx[b, c, d, h, w] = sum over taps/outputs that read it of dcols[(b,od,oh,ow), c*27 + tap]
__global__ void __launch_bounds__(256) col2im3d_kernel(const bf16* __restrict__ dcols, float* __restrict__ dx, const Conv3dGeom g) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;   // over B*D*H*W*Cin, c fastest when channels-last
  const long long total = (long long)g.B * g.Cin * g.D * g.H * g.W;
  if (idx >= total) return;
  int c, w, h, d, b;
  long long t = idx;
  if (g.sc == 1) {   // channels-last: keep c fastest so writes coalesce
    c = (int)(t % g.Cin); t /= g.Cin;
    w = (int)(t % g.W); t /= g.W;
    h = (int)(t % g.H); t /= g.H;
    d = (int)(t % g.D); b = (int)(t / g.D);
  } else {
    w = (int)(t % g.W); t /= g.W;
    h = (int)(t % g.H); t /= g.H;
    d = (int)(t % g.D); t /= g.D;
    c = (int)(t % g.Cin); b = (int)(t / g.Cin);
  }
  float acc = 0.f;
#pragma unroll
  for (int kd = 0; kd < 3; ++kd) {
    const int nd = d + 1 - kd;
    if (nd < 0 || nd % g.stride) continue;
    const int od = nd / g.stride;
    if (od >= g.Do) continue;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int nh = h + 1 - kh;
      if (nh < 0 || nh % g.stride) continue;
      const int oh = nh / g.stride;
      if (oh >= g.Ho) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int nw = w + 1 - kw;
        if (nw < 0 || nw % g.stride) continue;
        const int ow = nw / g.stride;
        if (ow >= g.Wo) continue;
        const long long m = (((long long)b * g.Do + od) * g.Ho + oh) * g.Wo + ow;
        acc += __bfloat162float(dcols[m * g.Kp + c * 27 + kd * 9 + kh * 3 + kw]);
      }
    }
  }
  dx[b * g.sb + c * g.sc + d * g.sd + h * g.sh + w * g.sw] = acc;
}

// ---- channels-last, tap-major variants (k = tap*Cin + c; sc == 1, Cin % 8 == 0) ----------------------------------------------
// The cin-major column order above follows weight.view(Cout, Cin*27); on a channels-last input it makes both kernels touch 2 useful
// bytes per 54-byte stride.  With the weight matrix permuted to [Cout, 27, Cin] instead, a patch row is 27 contiguous runs of Cin
// channels: one thread moves 8 channels (a 16-byte store / load), a warp a whole 256-channel run.
template <typename TIn>
__global__ void __launch_bounds__(256) im2col3d_cl_kernel(const TIn* __restrict__ x, bf16* __restrict__ cols, const Conv3dGeom g) {
  const int c8n = g.Cin >> 3;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long total = (long long)g.B * g.Do * g.Ho * g.Wo * 27 * c8n;
  if (idx >= total) return;
  const int c8 = (int)(idx % c8n);
  long long m = idx / c8n;
  const int tap = (int)(m % 27); m /= 27;
  const int kd = tap / 9, kh = (tap - kd * 9) / 3, kw = tap - kd * 9 - kh * 3;
  const int ow = (int)(m % g.Wo); m /= g.Wo;
  const int oh = (int)(m % g.Ho); m /= g.Ho;
  const int od = (int)(m % g.Do);
  const int b = (int)(m / g.Do);
  const int id = od * g.stride - 1 + kd, ih = oh * g.stride - 1 + kh, iw = ow * g.stride - 1 + kw;
  uint4 o = make_uint4(0u, 0u, 0u, 0u);
  if (id >= 0 && id < g.D && ih >= 0 && ih < g.H && iw >= 0 && iw < g.W) {
    const TIn* src = x + b * g.sb + id * g.sd + ih * g.sh + iw * g.sw + c8 * 8;
    if constexpr (sizeof(TIn) == 4) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(src)), c = __ldg(reinterpret_cast<const float4*>(src) + 1);
      o = make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(c.x, c.y), pack_bf16(c.z, c.w));
    } else {
      o = __ldg(reinterpret_cast<const uint4*>(src));
    }
  }
  reinterpret_cast<uint4*>(cols)[idx] = o;
}

__global__ void __launch_bounds__(256) col2im3d_cl_kernel(const bf16* __restrict__ dcols, float* __restrict__ dx, const Conv3dGeom g) {
  const int c8n = g.Cin >> 3;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long total = (long long)g.B * g.D * g.H * g.W * c8n;
  if (idx >= total) return;
  const int c8 = (int)(idx % c8n);
  long long t = idx / c8n;
  const int w = (int)(t % g.W); t /= g.W;
  const int h = (int)(t % g.H); t /= g.H;
  const int d = (int)(t % g.D);
  const int b = (int)(t / g.D);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int kd = 0; kd < 3; ++kd) {
    const int nd = d + 1 - kd;
    if (nd < 0 || nd % g.stride) continue;
    const int od = nd / g.stride;
    if (od >= g.Do) continue;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int nh = h + 1 - kh;
      if (nh < 0 || nh % g.stride) continue;
      const int oh = nh / g.stride;
      if (oh >= g.Ho) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int nw = w + 1 - kw;
        if (nw < 0 || nw % g.stride) continue;
        const int ow = nw / g.stride;
        if (ow >= g.Wo) continue;
        const long long m = (((long long)b * g.Do + od) * g.Ho + oh) * g.Wo + ow;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(dcols + m * g.Kp + (kd * 9 + kh * 3 + kw) * g.Cin + c8 * 8));
        const float2 p0 = unpack_bf16(v.x), p1 = unpack_bf16(v.y), p2 = unpack_bf16(v.z), p3 = unpack_bf16(v.w);
        acc[0] += p0.x; acc[1] += p0.y; acc[2] += p1.x; acc[3] += p1.y;
        acc[4] += p2.x; acc[5] += p2.y; acc[6] += p3.x; acc[7] += p3.y;
      }
    }
  }
  float4* dst = reinterpret_cast<float4*>(dx + b * g.sb + d * g.sd + h * g.sh + w * g.sw + c8 * 8);
  dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
  dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
}

// ------------------------------------------------------------------ per-(batch, channel) sums over voxels
// MODE 0: S1 = sum x,        S2 = sum x^2                          (GroupNorm forward statistics)
// MODE 1: S1 = sum ds,       S2 = sum ds * xhat   with ds = dy * silu'(z), z = xhat*w + b   (backward)
struct GnStatArgs {
  const float* x; const float* dy;      // [B, V, C] channels-last f32
  const float* mean; const float* rstd; // [B, G]
  const float* w; const float* b;       // [C]
  float* S1; float* S2;                 // [B, C]
  int V, C, cpg, rows_per_block;
  int act;                              // 0 = SiLU (GroupNorm+SiLU of the voxel embed), 1 = ReLU (BatchNorm2d+ReLU of the X-ray encoder),
                                        // 2 = GELU (GroupNorm+GELU of the cascade's multi-scale branches)
};
__device__ __forceinline__ float act_fwd(float z, int act) { return act == 0 ? z / (1.f + __expf(-z)) : act == 1 ? fmaxf(z, 0.f) : gelu_erf(z); }
__device__ __forceinline__ float silu_grad(float z) {
  const float s = 1.f / (1.f + __expf(-z));
  return s * (1.f + z * (1.f - s));
}
__device__ __forceinline__ float act_grad(float z, int act) { return act == 0 ? silu_grad(z) : act == 1 ? (z > 0.f ? 1.f : 0.f) : gelu_erf_grad(z); }
template <int MODE>
__global__ void __launch_bounds__(256) gn_stats_kernel(const GnStatArgs a) {
  __shared__ float4 red1[256], red2[256];
  const int tpr = a.C >> 2;                 // threads per row (C/4 <= 256)
  const int rpp = 256 / tpr;                // rows per pass
  const int rin = threadIdx.x / tpr, cv = threadIdx.x - rin * tpr;
  const int b = blockIdx.y;
  const int r0 = blockIdx.x * a.rows_per_block, r1 = min(r0 + a.rows_per_block, a.V);
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
  float mean[4] = {0, 0, 0, 0}, rstd[4] = {1, 1, 1, 1};
  float4 wv = make_float4(1, 1, 1, 1), bv = make_float4(0, 0, 0, 0);
  if (MODE == 1 && rin < rpp) {
    const int G = a.C / a.cpg;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int g = (4 * cv + e) / a.cpg;
      mean[e] = a.mean[b * G + g];
      rstd[e] = a.rstd[b * G + g];
    }
    wv = *reinterpret_cast<const float4*>(a.w + 4 * cv);
    bv = *reinterpret_cast<const float4*>(a.b + 4 * cv);
  }
  if (rin < rpp) {
    for (int r = r0 + rin; r < r1; r += rpp) {
      const long long off = ((long long)b * a.V + r) * a.C + 4 * cv;
      const float4 x = __ldg(reinterpret_cast<const float4*>(a.x + off));
      if (MODE == 0) {
        s1.x += x.x; s1.y += x.y; s1.z += x.z; s1.w += x.w;
        s2.x += x.x * x.x; s2.y += x.y * x.y; s2.z += x.z * x.z; s2.w += x.w * x.w;
      } else {
        const float4 dy = __ldg(reinterpret_cast<const float4*>(a.dy + off));
        const float xh[4] = {(x.x - mean[0]) * rstd[0], (x.y - mean[1]) * rstd[1], (x.z - mean[2]) * rstd[2], (x.w - mean[3]) * rstd[3]};
        const float d0 = dy.x * act_grad(xh[0] * wv.x + bv.x, a.act), d1 = dy.y * act_grad(xh[1] * wv.y + bv.y, a.act);
        const float d2 = dy.z * act_grad(xh[2] * wv.z + bv.z, a.act), d3 = dy.w * act_grad(xh[3] * wv.w + bv.w, a.act);
        s1.x += d0; s1.y += d1; s1.z += d2; s1.w += d3;
        s2.x += d0 * xh[0]; s2.y += d1 * xh[1]; s2.z += d2 * xh[2]; s2.w += d3 * xh[3];
      }
    }
  }
  red1[threadIdx.x] = s1;
  red2[threadIdx.x] = s2;
  __syncthreads();
  if (rin == 0) {
    for (int k = 1; k < rpp; ++k) {
      const float4 u = red1[k * tpr + cv], v = red2[k * tpr + cv];
      s1.x += u.x; s1.y += u.y; s1.z += u.z; s1.w += u.w;
      s2.x += v.x; s2.y += v.y; s2.z += v.z; s2.w += v.w;
    }
    float* d1 = a.S1 + (long long)b * a.C + 4 * cv;
    float* d2 = a.S2 + (long long)b * a.C + 4 * cv;
    atomicAdd(d1, s1.x); atomicAdd(d1 + 1, s1.y); atomicAdd(d1 + 2, s1.z); atomicAdd(d1 + 3, s1.w);
    atomicAdd(d2, s2.x); atomicAdd(d2 + 1, s2.y); atomicAdd(d2 + 2, s2.z); atomicAdd(d2 + 3, s2.w);
  }
}
// forward: channel sums -> group mean / rstd
__global__ void gn_fwd_finalize_kernel(const float* S1, const float* S2, float* mean, float* rstd, int B, int C, int cpg, int V) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int G = C / cpg;
  if (idx >= B * G) return;
  const int b = idx / G, g = idx - b * G;
  double s = 0.0, q = 0.0;
  for (int c = g * cpg; c < (g + 1) * cpg; ++c) { s += S1[b * C + c]; q += S2[b * C + c]; }
  const double n = (double)V * cpg;
  const double m = s / n;
  double var = q / n - m * m;
  if (var < 0.0) var = 0.0;
  mean[idx] = (float)m;
  rstd[idx] = (float)(1.0 / sqrt(var + 1e-5));
}
// backward: dw[c] = sum_b T2, db[c] = sum_b T1, A[b,g] = sum_{c in g} w_c T1 / n, Bq[b,g] = sum_{c in g} w_c T2 / n
__global__ void gn_bwd_finalize_kernel(const float* T1, const float* T2, const float* w, float* dw, float* db, float* A, float* Bq,
                                       int B, int C, int cpg, int V) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int G = C / cpg;
  if (idx < C) {
    float t1 = 0.f, t2 = 0.f;
    for (int b = 0; b < B; ++b) { t1 += T1[b * C + idx]; t2 += T2[b * C + idx]; }
    dw[idx] = t2;
    db[idx] = t1;
  }
  if (idx < B * G) {
    const int b = idx / G, g = idx - b * G;
    double a1 = 0.0, a2 = 0.0;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) { a1 += (double)w[c] * T1[b * C + c]; a2 += (double)w[c] * T2[b * C + c]; }
    const double n = (double)V * cpg;
    A[idx] = (float)(a1 / n);
    Bq[idx] = (float)(a2 / n);
  }
}
// forward apply: y = act(xhat*w + b) (bf16 or f32, channels-last).  backward apply: dx = rstd*(ds*w - A - xhat*Bq).
// A thread keeps one 4-channel group (its statistics and affine terms live in registers) and walks rows_per_block rows of one
// batch element with four independent 16-byte loads in flight.
template <int MODE>
__global__ void __launch_bounds__(256) gn_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ mean,
                                                       const float* __restrict__ rstd, const float* __restrict__ w, const float* __restrict__ b,
                                                       const float* __restrict__ A, const float* __restrict__ Bq, bf16* __restrict__ y,
                                                       float* __restrict__ dx, int V, int C, int cpg, int y_f32, int act, int rows_per_block) {
  const int tpr = C >> 2;                   // threads per row (C/4 <= 256)
  const int rpp = 256 / tpr;                // rows per pass
  const int rin = threadIdx.x / tpr, cv = threadIdx.x - rin * tpr;
  if (rin >= rpp) return;
  const int bb = blockIdx.y, G = C / cpg;
  float mu[4], rs[4], ag[4] = {0, 0, 0, 0}, bq[4] = {0, 0, 0, 0};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int g = bb * G + (4 * cv + e) / cpg;
    mu[e] = mean[g];
    rs[e] = rstd[g];
    if (MODE == 1) { ag[e] = A[g]; bq[e] = Bq[g]; }
  }
  const float4 wv = *reinterpret_cast<const float4*>(w + 4 * cv), bv = *reinterpret_cast<const float4*>(b + 4 * cv);
  const float ws[4] = {wv.x, wv.y, wv.z, wv.w}, bs[4] = {bv.x, bv.y, bv.z, bv.w};
  const int r0 = blockIdx.x * rows_per_block, r1 = min(r0 + rows_per_block, V);
  for (int r = r0 + rin; r < r1; r += 4 * rpp) {
    float4 xv[4], dv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int rr = r + u * rpp;
      if (rr < r1) {
        const long long off = ((long long)bb * V + rr) * C + 4 * cv;
        xv[u] = __ldg(reinterpret_cast<const float4*>(x + off));
        if (MODE == 1) dv[u] = __ldg(reinterpret_cast<const float4*>(dy + off));
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int rr = r + u * rpp;
      if (rr >= r1) break;
      const long long off = ((long long)bb * V + rr) * C + 4 * cv;
      const float xs[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w};
      const float dys[4] = {dv[u].x, dv[u].y, dv[u].z, dv[u].w};
      float out[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float xh = (xs[e] - mu[e]) * rs[e];
        const float z = xh * ws[e] + bs[e];
        if (MODE == 0) out[e] = act_fwd(z, act);
        else out[e] = rs[e] * (dys[e] * act_grad(z, act) * ws[e] - ag[e] - xh * bq[e]);
      }
      if (MODE == 0) {
        if (y_f32) *reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + off) = make_float4(out[0], out[1], out[2], out[3]);
        else *reinterpret_cast<uint2*>(y + off) = make_uint2(pack_bf16(out[0], out[1]), pack_bf16(out[2], out[3]));
      } else {
        *reinterpret_cast<float4*>(dx + off) = make_float4(out[0], out[1], out[2], out[3]);
      }
    }
  }
}

static int fill_geom(Conv3dGeom* g, const hvc_conv3d_geom* a) {
  HVC_CHECK_ARG(a->B > 0 && a->Cin > 0 && a->D > 0 && a->H > 0 && a->W > 0, "conv3d: empty input");
  HVC_CHECK_ARG(a->stride == 1 || a->stride == 2, "conv3d: stride %d not supported", a->stride);
  g->B = a->B; g->Cin = a->Cin; g->D = a->D; g->H = a->H; g->W = a->W; g->stride = a->stride;
  g->Do = (a->D - 1) / a->stride + 1; g->Ho = (a->H - 1) / a->stride + 1; g->Wo = (a->W - 1) / a->stride + 1;
  g->sb = a->sb; g->sc = a->sc; g->sd = a->sd; g->sh = a->sh; g->sw = a->sw;
  g->K = a->Cin * 27; g->Kp = (g->K + 7) / 8 * 8;
  return HVC_OK;
}

}  // namespace hvc

using namespace hvc;

extern "C" int hvc_im2col3d(const void* x, int32_t x_is_bf16, const hvc_conv3d_geom* geom, void* cols, void* stream) {
  HVC_CHECK_ARG(x && geom && cols, "hvc_im2col3d: null operand");
  Conv3dGeom g;
  int rc = fill_geom(&g, geom);
  if (rc) return rc;
  const long long total = (long long)g.B * g.Do * g.Ho * g.Wo * g.Kp;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (g.Cin == 1) {
    const unsigned rb = (unsigned)(((long long)g.B * g.Do * g.Ho * g.Wo + 255) / 256);
    if (x_is_bf16) im2col3d_c1_kernel<bf16><<<rb, 256, 0, st>>>(reinterpret_cast<const bf16*>(x), reinterpret_cast<bf16*>(cols), g);
    else im2col3d_c1_kernel<float><<<rb, 256, 0, st>>>(reinterpret_cast<const float*>(x), reinterpret_cast<bf16*>(cols), g);
  } else if (x_is_bf16) {
    im2col3d_kernel<bf16><<<blocks, 256, 0, st>>>(reinterpret_cast<const bf16*>(x), reinterpret_cast<bf16*>(cols), g);
  } else {
    im2col3d_kernel<float><<<blocks, 256, 0, st>>>(reinterpret_cast<const float*>(x), reinterpret_cast<bf16*>(cols), g);
  }
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

// fp32 verification mode (hvc_fp32.cu): the patch matrix keeps the full fp32 values; hvc_split3 then turns it into
// the three-term bf16 operand of the tensor-core GEMM.
extern "C" int hvc_im2col3d_f32(const float* x, const hvc_conv3d_geom* geom, float* cols, void* stream) {
  HVC_CHECK_ARG(x && geom && cols, "hvc_im2col3d_f32: null operand");
  Conv3dGeom g;
  int rc = fill_geom(&g, geom);
  if (rc) return rc;
  const long long total = (long long)g.B * g.Do * g.Ho * g.Wo * g.Kp;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  im2col3d_kernel<float, float><<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, cols, g);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

extern "C" int hvc_col2im3d(const void* dcols, const hvc_conv3d_geom* geom, float* dx, void* stream) {
  HVC_CHECK_ARG(dcols && geom && dx, "hvc_col2im3d: null operand");
  Conv3dGeom g;
  int rc = fill_geom(&g, geom);
  if (rc) return rc;
  const long long total = (long long)g.B * g.Cin * g.D * g.H * g.W;
  col2im3d_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const bf16*>(dcols), dx, g);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

static int check_cl_geom(const hvc_conv3d_geom* geom, const void* p, int elem_bytes) {
  HVC_CHECK_ARG(geom->sc == 1 && geom->Cin % 8 == 0, "hvc_*3d_cl: needs a channels-last tensor (sc == 1) with Cin % 8 == 0");
  HVC_CHECK_ARG(geom->sb % 8 == 0 && geom->sd % 8 == 0 && geom->sh % 8 == 0 && geom->sw % 8 == 0 &&
                reinterpret_cast<uintptr_t>(p) % (8 * elem_bytes) == 0, "hvc_*3d_cl: strides / base must keep 8-channel runs aligned");
  return HVC_OK;
}

extern "C" int hvc_im2col3d_cl(const void* x, int32_t x_is_bf16, const hvc_conv3d_geom* geom, void* cols, void* stream) {
  HVC_CHECK_ARG(x && geom && cols, "hvc_im2col3d_cl: null operand");
  int rc = check_cl_geom(geom, x, x_is_bf16 ? 2 : 4);
  if (rc) return rc;
  Conv3dGeom g;
  rc = fill_geom(&g, geom);
  if (rc) return rc;
  const long long total = (long long)g.B * g.Do * g.Ho * g.Wo * 27 * (g.Cin / 8);
  const unsigned blocks = (unsigned)((total + 255) / 256);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (x_is_bf16) im2col3d_cl_kernel<bf16><<<blocks, 256, 0, st>>>(reinterpret_cast<const bf16*>(x), reinterpret_cast<bf16*>(cols), g);
  else im2col3d_cl_kernel<float><<<blocks, 256, 0, st>>>(reinterpret_cast<const float*>(x), reinterpret_cast<bf16*>(cols), g);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

extern "C" int hvc_col2im3d_cl(const void* dcols, const hvc_conv3d_geom* geom, float* dx, void* stream) {
  HVC_CHECK_ARG(dcols && geom && dx, "hvc_col2im3d_cl: null operand");
  int rc = check_cl_geom(geom, dx, 4);
  if (rc) return rc;
  Conv3dGeom g;
  rc = fill_geom(&g, geom);
  if (rc) return rc;
  const long long total = (long long)g.B * g.D * g.H * g.W * (g.Cin / 8);
  col2im3d_cl_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const bf16*>(dcols), dx, g);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

extern "C" int hvc_pad3d_cl(const void* src, int32_t src_is_bf16, void* dst, int32_t B, int32_t D, int32_t H, int32_t W, int32_t Cs,
                            int32_t Cp, int32_t pad_hi, void* stream) {
  HVC_CHECK_ARG(src && dst && B > 0 && D > 0 && H > 0 && W > 0 && (pad_hi == 0 || pad_hi == 1), "hvc_pad3d_cl: bad arguments");
  HVC_CHECK_ARG(Cs > 0 && Cs % 8 == 0 && Cp % 8 == 0 && Cs <= Cp, "hvc_pad3d_cl: channel counts must be multiples of 8, Cs <= Cp");
  const long long total = (long long)B * (D + 1 + pad_hi) * (H + 1 + pad_hi) * (W + 1 + pad_hi) * (Cp / 8);
  const unsigned blocks = (unsigned)((total + 255) / 256);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (src_is_bf16) pad3d_cl_kernel<bf16><<<blocks, 256, 0, st>>>(reinterpret_cast<const bf16*>(src), reinterpret_cast<bf16*>(dst), B, D, H, W, Cs, Cp, pad_hi);
  else pad3d_cl_kernel<float><<<blocks, 256, 0, st>>>(reinterpret_cast<const float*>(src), reinterpret_cast<bf16*>(dst), B, D, H, W, Cs, Cp, pad_hi);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

extern "C" int hvc_unpad3d_cl(const float* src, float* dst, int32_t B, int32_t D, int32_t H, int32_t W, int32_t C, int32_t pad_hi,
                              void* stream) {
  HVC_CHECK_ARG(src && dst && B > 0 && D > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0 && (pad_hi == 0 || pad_hi == 1), "hvc_unpad3d_cl: bad arguments");
  const long long total = (long long)B * D * H * W * (C / 4);
  unpad3d_cl_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, dst, B, D, H, W, C, pad_hi);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

extern "C" int hvc_s2d_pad_cl(const void* src, int32_t src_is_bf16, void* dst, int32_t B, int32_t D, int32_t H, int32_t W, int32_t C,
                              void* stream) {
  HVC_CHECK_ARG(src && dst && B > 0 && D > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "hvc_s2d_pad_cl: bad arguments");
  HVC_CHECK_ARG(D % 2 == 0 && H % 2 == 0 && W % 2 == 0, "hvc_s2d_pad_cl: the volume sizes must be even");
  const long long total = 8LL * B * (D / 2 + 1) * (H / 2 + 1) * (W / 2 + 1) * (C / 8);
  const unsigned blocks = (unsigned)((total + 255) / 256);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (src_is_bf16) s2d_pad_cl_kernel<bf16><<<blocks, 256, 0, st>>>(reinterpret_cast<const bf16*>(src), reinterpret_cast<bf16*>(dst), B, D, H, W, C);
  else s2d_pad_cl_kernel<float><<<blocks, 256, 0, st>>>(reinterpret_cast<const float*>(src), reinterpret_cast<bf16*>(dst), B, D, H, W, C);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

extern "C" int hvc_d2s_unpad_cl(const float* src, float* dst, int32_t B, int32_t D, int32_t H, int32_t W, int32_t C, void* stream) {
  HVC_CHECK_ARG(src && dst && B > 0 && D > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, "hvc_d2s_unpad_cl: bad arguments");
  HVC_CHECK_ARG(D % 2 == 0 && H % 2 == 0 && W % 2 == 0, "hvc_d2s_unpad_cl: the volume sizes must be even");
  const long long total = (long long)B * D * H * W * (C / 4);
  d2s_unpad_cl_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, dst, B, D, H, W, C);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

static int gn_rows_per_block(int B, int V) {
  int rpb = 1024;
  while (rpb > 32 && (long long)B * ((V + rpb - 1) / rpb) < 4LL * device_sm_count()) rpb >>= 1;
  return rpb;
}

extern "C" int hvc_norm_act_fwd(const float* x, const float* w, const float* b, int32_t B, int32_t V, int32_t C, int32_t groups,
                                int32_t activation, int32_t stats_given, void* y, int32_t y_is_bf16, float* mean, float* rstd,
                                float* scratch, void* stream) {
  HVC_CHECK_ARG(x && w && b && y && mean && rstd && (scratch || stats_given), "hvc_norm_act_fwd: null operand");
  HVC_CHECK_ARG(B > 0 && V > 0 && C > 0 && (C & 3) == 0 && C <= 1024 && groups > 0 && C % groups == 0, "hvc_norm_act_fwd: bad shape C=%d G=%d", C, groups);
  HVC_CHECK_ARG(activation >= 0 && activation <= 2, "hvc_norm_act_fwd: activation must be 0 (SiLU), 1 (ReLU) or 2 (GELU)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int cpg = C / groups;
  if (!stats_given) {
    HVC_CUDA(cudaMemsetAsync(scratch, 0, sizeof(float) * 2 * B * C, st));
    GnStatArgs a;
    a.x = x; a.dy = nullptr; a.mean = nullptr; a.rstd = nullptr; a.w = w; a.b = b; a.S1 = scratch; a.S2 = scratch + (long long)B * C;
    a.V = V; a.C = C; a.cpg = cpg; a.rows_per_block = gn_rows_per_block(B, V); a.act = activation;
    gn_stats_kernel<0><<<dim3((V + a.rows_per_block - 1) / a.rows_per_block, B), 256, 0, st>>>(a);
    HVC_LAUNCH_CHECK();
    gn_fwd_finalize_kernel<<<(B * groups + 127) / 128, 128, 0, st>>>(a.S1, a.S2, mean, rstd, B, C, cpg, V);
    HVC_LAUNCH_CHECK();
  }
  const int rpb = gn_rows_per_block(B, V);
  gn_apply_kernel<0><<<dim3((V + rpb - 1) / rpb, B), 256, 0, st>>>(x, nullptr, mean, rstd, w, b, nullptr, nullptr, reinterpret_cast<bf16*>(y),
                                                                   nullptr, V, C, cpg, y_is_bf16 ? 0 : 1, activation, rpb);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}
extern "C" int hvc_groupnorm_silu_fwd(const float* x, const float* w, const float* b, int32_t B, int32_t V, int32_t C, int32_t groups,
                                      void* y, int32_t y_is_bf16, float* mean, float* rstd, float* scratch, void* stream) {
  return hvc_norm_act_fwd(x, w, b, B, V, C, groups, 0, 0, y, y_is_bf16, mean, rstd, scratch, stream);
}

extern "C" int hvc_groupnorm_silu_bwd(const float* dy, const float* x, const float* w, const float* b, const float* mean, const float* rstd,
                                      int32_t B, int32_t V, int32_t C, int32_t groups, float* dx, float* dw, float* db, float* scratch,
                                      void* stream) {
  return hvc_norm_act_bwd(dy, x, w, b, mean, rstd, B, V, C, groups, 0, 0, dx, dw, db, scratch, stream);
}
extern "C" int hvc_norm_act_bwd(const float* dy, const float* x, const float* w, const float* b, const float* mean, const float* rstd,
                                int32_t B, int32_t V, int32_t C, int32_t groups, int32_t activation, int32_t stats_frozen, float* dx,
                                float* dw, float* db, float* scratch, void* stream) {
  HVC_CHECK_ARG(dy && x && w && b && mean && rstd && dx && dw && db && scratch, "hvc_norm_act_bwd: null operand");
  HVC_CHECK_ARG(B > 0 && V > 0 && C > 0 && (C & 3) == 0 && C <= 1024 && groups > 0 && C % groups == 0, "hvc_norm_act_bwd: bad shape");
  HVC_CHECK_ARG(activation >= 0 && activation <= 2, "hvc_norm_act_bwd: activation must be 0 (SiLU), 1 (ReLU) or 2 (GELU)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // scratch: T1 [B,C], T2 [B,C], A [B,G], Bq [B,G]
  HVC_CUDA(cudaMemsetAsync(scratch, 0, sizeof(float) * 2 * B * C, st));
  float* T1 = scratch; float* T2 = scratch + (long long)B * C; float* A = T2 + (long long)B * C; float* Bq = A + (long long)B * groups;
  GnStatArgs a;
  a.x = x; a.dy = dy; a.mean = mean; a.rstd = rstd; a.w = w; a.b = b; a.S1 = T1; a.S2 = T2;
  a.V = V; a.C = C; a.cpg = C / groups; a.rows_per_block = gn_rows_per_block(B, V); a.act = activation;
  gn_stats_kernel<1><<<dim3((V + a.rows_per_block - 1) / a.rows_per_block, B), 256, 0, st>>>(a);
  HVC_LAUNCH_CHECK();
  const int n = C > B * groups ? C : B * groups;
  gn_bwd_finalize_kernel<<<(n + 127) / 128, 128, 0, st>>>(T1, T2, w, dw, db, A, Bq, B, C, a.cpg, V);
  HVC_LAUNCH_CHECK();
  if (stats_frozen) HVC_CUDA(cudaMemsetAsync(A, 0, sizeof(float) * 2 * B * groups, st));   // eval-mode BatchNorm: statistics are constants
  gn_apply_kernel<1><<<dim3((V + a.rows_per_block - 1) / a.rows_per_block, B), 256, 0, st>>>(x, dy, mean, rstd, w, b, A, Bq, nullptr, dx, V, C,
                                                                                             a.cpg, 0, activation, a.rows_per_block);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}
