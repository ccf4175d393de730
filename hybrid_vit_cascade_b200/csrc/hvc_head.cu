// hvc_head.cu -- the ends of HybridViT3D.forward around the transformer blocks:
//   tokens + pos_embed                                   (hybrid_vit_backbone.py:258)
//   final LayerNorm -> output_proj (C -> 1) per token     (:265-266), fused: the normalised row never hits HBM
//   reshape to the token grid + trilinear upsample, align_corners=True   (:269-272)
// and their backward passes.  All HBM-bound.
#include "hvc_common.cuh"
#include "hvc_host.h"

namespace hvc {

// v[t] = bo + sum_c (LN(x[t])_c * wo_c) ; one warp per row, 4*VPL columns per lane
template <int VPL>
__global__ void __launch_bounds__(256) head_fwd_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ w,
                                                       const float* __restrict__ b, const float* __restrict__ wo, const float* __restrict__ bo,
                                                       float* __restrict__ v, float* __restrict__ mean_out, float* __restrict__ rstd_out, int T, int C) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= T) return;
  float4 xv[VPL];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = 4 * (lane + 32 * i);
    xv[i] = c < C ? __ldg(reinterpret_cast<const float4*>(x + (long long)row * ldx + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    s += xv[i].x + xv[i].y + xv[i].z + xv[i].w;
  }
  const float mean = warp_sum(s) / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    if (4 * (lane + 32 * i) < C) {
      const float a0 = xv[i].x - mean, a1 = xv[i].y - mean, a2 = xv[i].z - mean, a3 = xv[i].w - mean;
      q += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3;
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / C + 1e-5f);
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = 4 * (lane + 32 * i);
    if (c < C) {
      const float4 ww = __ldg(reinterpret_cast<const float4*>(w + c)), bb = __ldg(reinterpret_cast<const float4*>(b + c));
      const float4 oo = __ldg(reinterpret_cast<const float4*>(wo + c));
      dot += ((xv[i].x - mean) * rstd * ww.x + bb.x) * oo.x + ((xv[i].y - mean) * rstd * ww.y + bb.y) * oo.y +
             ((xv[i].z - mean) * rstd * ww.z + bb.z) * oo.z + ((xv[i].w - mean) * rstd * ww.w + bb.w) * oo.w;
    }
  }
  dot = warp_sum(dot);
  if (lane == 0) {
    v[row] = dot + __ldg(bo);
    mean_out[row] = mean;
    rstd_out[row] = rstd;
  }
}

struct UpGeom { int B, Di, Hi, Wi, Do, Ho, Wo; float sd, sh, sw; float half; };   // half = 0.5 for align_corners=False, 0 for True

// align_corners=True: src = o * (in-1)/(out-1); align_corners=False: src = max((o + 0.5) * in/out - 0.5, 0)  (PyTorch's area_pixel rule)
__device__ __forceinline__ void up_coord(int o, float scale, float half, int in, int& i0, int& i1, float& l1) {
  const float src = fmaxf(scale * (o + half) - half, 0.f);
  i0 = static_cast<int>(src);
  if (i0 > in - 1) i0 = in - 1;
  i1 = i0 + (i0 < in - 1 ? 1 : 0);
  l1 = src - i0;
}
// out[b, d, h, w] = trilinear(v[b]) with align_corners=True
__global__ void __launch_bounds__(256) upsample_fwd_kernel(const float* __restrict__ v, float* __restrict__ out, const UpGeom g) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long total = (long long)g.B * g.Do * g.Ho * g.Wo;
  if (idx >= total) return;
  long long t = idx;
  const int w = (int)(t % g.Wo); t /= g.Wo;
  const int h = (int)(t % g.Ho); t /= g.Ho;
  const int d = (int)(t % g.Do);
  const int b = (int)(t / g.Do);
  int d0, d1, h0, h1, w0, w1;
  float ld, lh, lw;
  up_coord(d, g.sd, g.half, g.Di, d0, d1, ld);
  up_coord(h, g.sh, g.half, g.Hi, h0, h1, lh);
  up_coord(w, g.sw, g.half, g.Wi, w0, w1, lw);
  const float* p = v + (long long)b * g.Di * g.Hi * g.Wi;
  auto at = [&](int dd, int hh, int ww) { return __ldg(p + ((long long)dd * g.Hi + hh) * g.Wi + ww); };
  const float c00 = at(d0, h0, w0) * (1.f - lw) + at(d0, h0, w1) * lw;
  const float c01 = at(d0, h1, w0) * (1.f - lw) + at(d0, h1, w1) * lw;
  const float c10 = at(d1, h0, w0) * (1.f - lw) + at(d1, h0, w1) * lw;
  const float c11 = at(d1, h1, w0) * (1.f - lw) + at(d1, h1, w1) * lw;
  const float c0 = c00 * (1.f - lh) + c01 * lh, c1 = c10 * (1.f - lh) + c11 * lh;
  out[idx] = c0 * (1.f - ld) + c1 * ld;
}
// adjoint: dv[b, coarse] += weights * dout[b, fine]   (dv zero-filled by the caller)
__global__ void __launch_bounds__(256) upsample_bwd_kernel(const float* __restrict__ dout, float* __restrict__ dv, const UpGeom g) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long total = (long long)g.B * g.Do * g.Ho * g.Wo;
  if (idx >= total) return;
  long long t = idx;
  const int w = (int)(t % g.Wo); t /= g.Wo;
  const int h = (int)(t % g.Ho); t /= g.Ho;
  const int d = (int)(t % g.Do);
  const int b = (int)(t / g.Do);
  int d0, d1, h0, h1, w0, w1;
  float ld, lh, lw;
  up_coord(d, g.sd, g.half, g.Di, d0, d1, ld);
  up_coord(h, g.sh, g.half, g.Hi, h0, h1, lh);
  up_coord(w, g.sw, g.half, g.Wi, w0, w1, lw);
  const float go = __ldg(dout + idx);
  float* p = dv + (long long)b * g.Di * g.Hi * g.Wi;
  auto add = [&](int dd, int hh, int ww, float wt) { atomicAdd(p + ((long long)dd * g.Hi + hh) * g.Wi + ww, go * wt); };
  add(d0, h0, w0, (1.f - ld) * (1.f - lh) * (1.f - lw)); add(d0, h0, w1, (1.f - ld) * (1.f - lh) * lw);
  add(d0, h1, w0, (1.f - ld) * lh * (1.f - lw));         add(d0, h1, w1, (1.f - ld) * lh * lw);
  add(d1, h0, w0, ld * (1.f - lh) * (1.f - lw));         add(d1, h0, w1, ld * (1.f - lh) * lw);
  add(d1, h1, w0, ld * lh * (1.f - lw));                 add(d1, h1, w1, ld * lh * lw);
}

// out[b, i] = x[(b mod xB), i] + pos[i]  for i < n (n = N*C, multiple of 4).  xB < B broadcasts the embedding of a
// batch-expanded input (model_direct.py:75 feeds the same learned volume to every sample).
__global__ void __launch_bounds__(256) add_pos_kernel(const float* __restrict__ x, const float* __restrict__ pos, float* __restrict__ out,
                                                      long long n, int B, int xB) {
  const long long i = ((long long)blockIdx.x * 256 + threadIdx.x) * 4;
  if (i >= n) return;
  const float4 p = __ldg(reinterpret_cast<const float4*>(pos + i));
  for (int b = 0; b < B; ++b) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + (long long)(b % xB) * n + i));
    *reinterpret_cast<float4*>(out + (long long)b * n + i) = make_float4(v.x + p.x, v.y + p.y, v.z + p.z, v.w + p.w);
  }
}
// out[i] = sum_b x[b, i]
__global__ void __launch_bounds__(256) batch_sum_kernel(const float* __restrict__ x, float* __restrict__ out, long long n, int B) {
  const long long i = ((long long)blockIdx.x * 256 + threadIdx.x) * 4;
  if (i >= n) return;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int b = 0; b < B; ++b) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + (long long)b * n + i));
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  *reinterpret_cast<float4*>(out + i) = s;
}

static UpGeom make_up(int B, int Di, int Hi, int Wi, int Do, int Ho, int Wo, int align_corners = 1) {
  UpGeom g;
  g.B = B; g.Di = Di; g.Hi = Hi; g.Wi = Wi; g.Do = Do; g.Ho = Ho; g.Wo = Wo;
  if (align_corners) {
    g.sd = Do > 1 ? (float)(Di - 1) / (float)(Do - 1) : 0.f;
    g.sh = Ho > 1 ? (float)(Hi - 1) / (float)(Ho - 1) : 0.f;
    g.sw = Wo > 1 ? (float)(Wi - 1) / (float)(Wo - 1) : 0.f;
    g.half = 0.f;
  } else {
    g.sd = (float)Di / (float)Do; g.sh = (float)Hi / (float)Ho; g.sw = (float)Wi / (float)Wo;
    g.half = 0.5f;
  }
  return g;
}

}  // namespace hvc

using namespace hvc;

extern "C" int hvc_head_fwd(const float* x, int64_t ldx, const float* w, const float* b, const float* wo, const float* bo, float* v,
                            float* mean, float* rstd, int32_t T, int32_t C, void* stream) {
  HVC_CHECK_ARG(x && w && b && wo && bo && v && mean && rstd, "hvc_head_fwd: null operand");
  HVC_CHECK_ARG(T > 0 && C > 0 && (C & 3) == 0 && C <= 1024, "hvc_head_fwd: C=%d must be a multiple of 4, <= 1024", C);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int grid = (T + 7) / 8;
  if (C <= 128) head_fwd_kernel<1><<<grid, 256, 0, st>>>(x, ldx, w, b, wo, bo, v, mean, rstd, T, C);
  else if (C <= 256) head_fwd_kernel<2><<<grid, 256, 0, st>>>(x, ldx, w, b, wo, bo, v, mean, rstd, T, C);
  else if (C <= 512) head_fwd_kernel<4><<<grid, 256, 0, st>>>(x, ldx, w, b, wo, bo, v, mean, rstd, T, C);
  else head_fwd_kernel<8><<<grid, 256, 0, st>>>(x, ldx, w, b, wo, bo, v, mean, rstd, T, C);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

extern "C" int hvc_upsample3d_fwd(const float* v, float* out, int32_t B, int32_t Di, int32_t Hi, int32_t Wi, int32_t Do, int32_t Ho,
                                  int32_t Wo, void* stream) {
  HVC_CHECK_ARG(v && out && B > 0 && Di > 0 && Hi > 0 && Wi > 0 && Do > 0 && Ho > 0 && Wo > 0, "hvc_upsample3d_fwd: bad arguments");
  const UpGeom g = make_up(B, Di, Hi, Wi, Do, Ho, Wo);
  const long long total = (long long)B * Do * Ho * Wo;
  upsample_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(v, out, g);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

extern "C" int hvc_upsample3d_bwd(const float* dout, float* dv, int32_t B, int32_t Di, int32_t Hi, int32_t Wi, int32_t Do, int32_t Ho,
                                  int32_t Wo, void* stream) {
  HVC_CHECK_ARG(dout && dv && B > 0 && Di > 0 && Hi > 0 && Wi > 0 && Do > 0 && Ho > 0 && Wo > 0, "hvc_upsample3d_bwd: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  HVC_CUDA(cudaMemsetAsync(dv, 0, sizeof(float) * (size_t)B * Di * Hi * Wi, st));
  const UpGeom g = make_up(B, Di, Hi, Wi, Do, Ho, Wo);
  const long long total = (long long)B * Do * Ho * Wo;
  upsample_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(dout, dv, g);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

// trilinear resize with either corner convention (align_corners = 0: nn.Upsample / F.interpolate of the cascade's stage wrappers,
// progressive_cascade/model_progressive.py:169,211-212)
extern "C" int hvc_interp3d_fwd(const float* v, float* out, int32_t B, int32_t Di, int32_t Hi, int32_t Wi, int32_t Do, int32_t Ho,
                                int32_t Wo, int32_t align_corners, void* stream) {
  HVC_CHECK_ARG(v && out && B > 0 && Di > 0 && Hi > 0 && Wi > 0 && Do > 0 && Ho > 0 && Wo > 0, "hvc_interp3d_fwd: bad arguments");
  const UpGeom g = make_up(B, Di, Hi, Wi, Do, Ho, Wo, align_corners);
  const long long total = (long long)B * Do * Ho * Wo;
  upsample_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(v, out, g);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}
extern "C" int hvc_interp3d_bwd(const float* dout, float* dv, int32_t B, int32_t Di, int32_t Hi, int32_t Wi, int32_t Do, int32_t Ho,
                                int32_t Wo, int32_t align_corners, void* stream) {
  HVC_CHECK_ARG(dout && dv && B > 0 && Di > 0 && Hi > 0 && Wi > 0 && Do > 0 && Ho > 0 && Wo > 0, "hvc_interp3d_bwd: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  HVC_CUDA(cudaMemsetAsync(dv, 0, sizeof(float) * (size_t)B * Di * Hi * Wi, st));
  const UpGeom g = make_up(B, Di, Hi, Wi, Do, Ho, Wo, align_corners);
  const long long total = (long long)B * Do * Ho * Wo;
  upsample_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(dout, dv, g);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

extern "C" int hvc_add_pos(const float* x, int32_t x_batch, const float* pos, float* out, int32_t B, int64_t n, void* stream) {
  HVC_CHECK_ARG(x && pos && out && B > 0 && x_batch > 0 && n > 0 && (n & 3) == 0, "hvc_add_pos: bad arguments");
  add_pos_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, pos, out, n, B, x_batch);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

extern "C" int hvc_batch_sum(const float* x, float* out, int32_t B, int64_t n, void* stream) {
  HVC_CHECK_ARG(x && out && B > 0 && n > 0 && (n & 3) == 0, "hvc_batch_sum: bad arguments");
  batch_sum_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, out, n, B);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}
