// hvc_encoder.cu -- the X-ray encoder in front of the backbone:
// XrayConditioningModule, models/diagnostic_losses.py:68-138 -- three Conv2d + BatchNorm2d + ReLU stages with two max-pools,
// the mean over the AP / lateral views and the global pooling that feed `context` and `cond` of HybridViT3D.
// Convolutions run as im2col -> tcgen05 GEMM (hvc_gemm) on channels-last activations [images, pixels, C],
// so the last stage emits the (B, H'W', C) context-token layout directly (the reference's flatten(2).transpose(1,2),
// model_direct.py:80, costs nothing); BatchNorm2d + ReLU is hvc_norm_act (hvc_embed.cu) with one group per channel over all
// rows.  The kernels here are the HBM-bound data movers around those: patch gather / scatter, max-pool, view mean, pooling,
// SiLU of the time MLP.
#include "hvc_common.cuh"
#include "hvc_host.h"

namespace hvc {

struct Conv2dGeom {
  int N, Cin, H, W;              // input images
  int Ho, Wo, k, stride, pad;    // square kernel
  long long sn, sc, sh, sw;      // input element strides
  int K, Kp;                     // Cin*k*k and its padding to a multiple of 8
};

// cols[m, kk] = x[n, cin, oh*s-p+kh, ow*s-p+kw] (0 outside), kk = cin*k*k + kh*k + kw  (= weight.view(Cout, Cin*k*k) order)
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) im2col2d_kernel(const TIn* __restrict__ x, TOut* __restrict__ cols, const Conv2dGeom g) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long total = (long long)g.N * g.Ho * g.Wo * g.Kp;
  if (idx >= total) return;
  const int kk = (int)(idx % g.Kp);
  long long m = idx / g.Kp;
  float v = 0.f;
  if (kk < g.K) {
    const int k2 = g.k * g.k;
    const int cin = kk / k2, tap = kk - cin * k2;
    const int kh = tap / g.k, kw = tap - kh * g.k;
    const int ow = (int)(m % g.Wo); m /= g.Wo;
    const int oh = (int)(m % g.Ho);
    const int n = (int)(m / g.Ho);
    const int ih = oh * g.stride - g.pad + kh, iw = ow * g.stride - g.pad + kw;
    if (ih >= 0 && ih < g.H && iw >= 0 && iw < g.W) v = static_cast<float>(x[n * g.sn + cin * g.sc + ih * g.sh + iw * g.sw]);
  }
  if constexpr (sizeof(TOut) == 4) cols[idx] = v;
  else cols[idx] = __float2bfloat16(v);
}

// The same gather with the two-term operand split fused in: out[m, :] = [c0 | c1 | c0] (three blocks of Kp columns), c0 = bf16(v),
// c1 = bf16(v - c0).  Against weights laid out [w0 | w0 | w1] one GEMM with K' = 3 Kp gives c0 w0 + c1 w0 + c0 w1 ~ the fp32
// product to 2^-16; block 0 alone is the plain bf16 patch matrix the backward GEMMs read (a strided view, no second gather).
__global__ void __launch_bounds__(256) im2col2d_split_kernel(const float* __restrict__ x, bf16* __restrict__ out, const Conv2dGeom g) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;      // one thread = 8 consecutive patch columns of one output pixel
  const int kv = g.Kp >> 3;
  const long long total = (long long)g.N * g.Ho * g.Wo * kv;
  if (idx >= total) return;
  const int kk0 = (int)(idx % kv) * 8;
  const long long m0 = idx / kv;
  long long m = m0;
  const int ow = (int)(m % g.Wo); m /= g.Wo;
  const int oh = (int)(m % g.Ho);
  const int n = (int)(m / g.Ho);
  const int k2 = g.k * g.k;
  uint32_t c0[4], c1[4];
#pragma unroll
  for (int e = 0; e < 8; e += 2) {
    float v[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int kk = kk0 + e + u;
      v[u] = 0.f;
      if (kk < g.K) {
        const int cin = kk / k2, tap = kk - cin * k2;
        const int kh = tap / g.k, kw = tap - kh * g.k;
        const int ih = oh * g.stride - g.pad + kh, iw = ow * g.stride - g.pad + kw;
        if (ih >= 0 && ih < g.H && iw >= 0 && iw < g.W) v[u] = __ldg(x + n * g.sn + cin * g.sc + ih * g.sh + iw * g.sw);
      }
    }
    const bf16 a0 = __float2bfloat16_rn(v[0]), b0 = __float2bfloat16_rn(v[1]);
    const bf16 a1 = __float2bfloat16_rn(v[0] - __bfloat162float(a0)), b1 = __float2bfloat16_rn(v[1] - __bfloat162float(b0));
    c0[e >> 1] = (uint32_t)__bfloat16_as_ushort(a0) | ((uint32_t)__bfloat16_as_ushort(b0) << 16);
    c1[e >> 1] = (uint32_t)__bfloat16_as_ushort(a1) | ((uint32_t)__bfloat16_as_ushort(b1) << 16);
  }
  bf16* row = out + m0 * (3LL * g.Kp) + kk0;
  *reinterpret_cast<uint4*>(row) = make_uint4(c0[0], c0[1], c0[2], c0[3]);
  *reinterpret_cast<uint4*>(row + g.Kp) = make_uint4(c1[0], c1[1], c1[2], c1[3]);
  *reinterpret_cast<uint4*>(row + 2 * g.Kp) = make_uint4(c0[0], c0[1], c0[2], c0[3]);
}

// dx[n, c, h, w] = sum over the (output pixel, tap) pairs that read it of dcols[(n,oh,ow), c*k*k + tap]; c fastest (channels-last dx)
__global__ void __launch_bounds__(256) col2im2d_kernel(const bf16* __restrict__ dcols, float* __restrict__ dx, const Conv2dGeom g) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long total = (long long)g.N * g.Cin * g.H * g.W;
  if (idx >= total) return;
  long long t = idx;
  const int c = (int)(t % g.Cin); t /= g.Cin;
  const int w = (int)(t % g.W); t /= g.W;
  const int h = (int)(t % g.H);
  const int n = (int)(t / g.H);
  const int k2 = g.k * g.k;
  float acc = 0.f;
  for (int kh = 0; kh < g.k; ++kh) {
    const int nh = h + g.pad - kh;
    if (nh < 0 || nh % g.stride) continue;
    const int oh = nh / g.stride;
    if (oh >= g.Ho) continue;
    for (int kw = 0; kw < g.k; ++kw) {
      const int nw = w + g.pad - kw;
      if (nw < 0 || nw % g.stride) continue;
      const int ow = nw / g.stride;
      if (ow >= g.Wo) continue;
      const long long m = ((long long)n * g.Ho + oh) * g.Wo + ow;
      acc += __bfloat162float(dcols[m * g.Kp + c * k2 + kh * g.k + kw]);
    }
  }
  dx[n * g.sn + c * g.sc + h * g.sh + w * g.sw] = acc;
}

// MaxPool2d on channels-last f32 [N, H, W, C] -> [N, Ho, Wo, C] (f32: the arg-max has to be taken on unrounded values -- on
// bf16 activations neighbouring pixels tie and the gradient is routed to the wrong one); idx = window-relative position of the first maximum
// (row-major scan, the element nn.MaxPool2d routes the gradient to), 255 for an all-padding window (cannot happen for the
// two pools of the encoder).
__global__ void __launch_bounds__(256) maxpool2d_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, uint8_t* __restrict__ arg, int N,
                                                            int H, int W, int C, int Ho, int Wo, int k, int stride, int pad) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long total = (long long)N * Ho * Wo * C;
  if (idx >= total) return;
  long long t = idx;
  const int c = (int)(t % C); t /= C;
  const int ow = (int)(t % Wo); t /= Wo;
  const int oh = (int)(t % Ho);
  const int n = (int)(t / Ho);
  float best = -INFINITY;
  int bi = 255;
  for (int kh = 0; kh < k; ++kh) {
    const int ih = oh * stride - pad + kh;
    if (ih < 0 || ih >= H) continue;
    for (int kw = 0; kw < k; ++kw) {
      const int iw = ow * stride - pad + kw;
      if (iw < 0 || iw >= W) continue;
      const float v = x[(((long long)n * H + ih) * W + iw) * C + c];
      if (v > best) { best = v; bi = kh * k + kw; }
    }
  }
  y[idx] = best;
  arg[idx] = static_cast<uint8_t>(bi);
}
// dx[n,h,w,c] = sum of dy over the windows whose recorded maximum is this element; four channels per thread (C % 4 == 0)
__global__ void __launch_bounds__(256) maxpool2d_bwd_kernel(const float* __restrict__ dy, const uint8_t* __restrict__ arg, float* __restrict__ dx,
                                                            int N, int H, int W, int C, int Ho, int Wo, int k, int stride, int pad) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  const int cv = C >> 2;
  const long long total = (long long)N * H * W * cv;
  if (idx >= total) return;
  long long t = idx;
  const int c = (int)(t % cv) * 4; t /= cv;
  const int w = (int)(t % W); t /= W;
  const int h = (int)(t % H);
  const int n = (int)(t / H);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int kh = 0; kh < k; ++kh) {
    const int nh = h + pad - kh;
    if (nh < 0 || nh % stride) continue;
    const int oh = nh / stride;
    if (oh >= Ho) continue;
    for (int kw = 0; kw < k; ++kw) {
      const int nw = w + pad - kw;
      if (nw < 0 || nw % stride) continue;
      const int ow = nw / stride;
      if (ow >= Wo) continue;
      const long long o = (((long long)n * Ho + oh) * Wo + ow) * C + c;
      const uchar4 a = *reinterpret_cast<const uchar4*>(arg + o);
      const float4 g = __ldg(reinterpret_cast<const float4*>(dy + o));
      const int me = kh * k + kw;
      acc.x += a.x == me ? g.x : 0.f;
      acc.y += a.y == me ? g.y : 0.f;
      acc.z += a.z == me ? g.z : 0.f;
      acc.w += a.w == me ? g.w : 0.f;
    }
  }
  *reinterpret_cast<float4*>(dx + (((long long)n * H + h) * W + w) * C + c) = acc;
}

// out[b, :] = mean_v x[b*V + v, :]   (n = pixels*C f32 elements per image; diagnostic_losses.py:125)
__global__ void __launch_bounds__(256) view_mean_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int V, long long n) {
  const long long idx = ((long long)blockIdx.x * 256 + threadIdx.x) * 4;
  if (idx >= (long long)B * n) return;
  const long long b = idx / n, e = idx - b * n;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int v = 0; v < V; ++v) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(x + (b * V + v) * n + e));
    acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
  }
  const float inv = 1.f / V;
  *reinterpret_cast<float4*>(out + idx) = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
}
// backward of the view mean fused with the backward of the global pooling:
//   dx[b*V + v, p, c] = (dfeat[b, p, c] + dpool[b, c] / P) / V          (dfeat or dpool may be NULL)
__global__ void __launch_bounds__(256) view_mean_bwd_kernel(const float* __restrict__ dfeat, const float* __restrict__ dpool, float* __restrict__ dx,
                                                            int B, int V, int P, int C) {
  const long long idx = ((long long)blockIdx.x * 256 + threadIdx.x) * 4;
  const long long n = (long long)P * C;
  if (idx >= (long long)B * n) return;
  const long long b = idx / n, e = idx - b * n;
  const int c = (int)(e % C);
  float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
  if (dfeat) g = __ldg(reinterpret_cast<const float4*>(dfeat + idx));
  if (dpool) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(dpool + b * C + c));
    const float ip = 1.f / P;
    g.x += q.x * ip; g.y += q.y * ip; g.z += q.z * ip; g.w += q.w * ip;
  }
  const float inv = 1.f / V;
  g.x *= inv; g.y *= inv; g.z *= inv; g.w *= inv;
  for (int v = 0; v < V; ++v) *reinterpret_cast<float4*>(dx + (b * V + v) * n + e) = g;
}
// out[b, c] = mean_p x[b, p, c]   (diagnostic_losses.py:130); one CTA per (batch, 4-column group slab)
__global__ void __launch_bounds__(256) pool_mean_kernel(const float* __restrict__ x, float* __restrict__ out, int P, int C) {
  __shared__ float4 red[256];
  const int b = blockIdx.y;
  const int tpr = C >> 2;                    // threads per row (C/4 <= 256)
  const int rpp = 256 / tpr;
  const int rin = threadIdx.x / tpr, cv = threadIdx.x - rin * tpr;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (rin < rpp) {
    for (int p = blockIdx.x * rpp + rin; p < P; p += gridDim.x * rpp) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(x + ((long long)b * P + p) * C + 4 * cv));
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  if (rin == 0) {
    for (int k = 1; k < rpp; ++k) {
      const float4 u = red[k * tpr + cv];
      acc.x += u.x; acc.y += u.y; acc.z += u.z; acc.w += u.w;
    }
    const float ip = 1.f / P;
    float* o = out + (long long)b * C + 4 * cv;
    atomicAdd(o, acc.x * ip); atomicAdd(o + 1, acc.y * ip); atomicAdd(o + 2, acc.z * ip); atomicAdd(o + 3, acc.w * ip);
  }
}

// SiLU of the time MLP (diagnostic_losses.py:100): y = x sigmoid(x); backward dx = dy * silu'(x)
__global__ void __launch_bounds__(256) silu_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ out, long long n) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= n) return;
  const float z = x[idx];
  const float s = 1.f / (1.f + __expf(-z));
  out[idx] = dy ? dy[idx] * s * (1.f + z * (1.f - s)) : z * s;
}

static int fill_geom2d(Conv2dGeom* g, const hvc_conv2d_geom* a) {
  HVC_CHECK_ARG(a->N > 0 && a->Cin > 0 && a->H > 0 && a->W > 0, "conv2d: empty input");
  HVC_CHECK_ARG(a->k >= 1 && a->k <= 7 && a->stride >= 1 && a->stride <= 2 && a->pad >= 0 && a->pad <= 3, "conv2d: kernel %d stride %d pad %d not supported",
                a->k, a->stride, a->pad);
  g->N = a->N; g->Cin = a->Cin; g->H = a->H; g->W = a->W; g->k = a->k; g->stride = a->stride; g->pad = a->pad;
  g->Ho = (a->H + 2 * a->pad - a->k) / a->stride + 1;
  g->Wo = (a->W + 2 * a->pad - a->k) / a->stride + 1;
  HVC_CHECK_ARG(g->Ho > 0 && g->Wo > 0, "conv2d: empty output");
  g->sn = a->sn; g->sc = a->sc; g->sh = a->sh; g->sw = a->sw;
  g->K = a->Cin * a->k * a->k; g->Kp = (g->K + 7) / 8 * 8;
  return HVC_OK;
}


// ---- Conv3d(C -> 1, kernel 1): the last layer of Stage3Refiner256.detail_enhancer (model_progressive.py:266) ------------------
// y: f32 [M, C] channels-last (one 128-byte row per voxel at C = 32).  out[m] = bias + sum_c y[m,c] w[c].
template <int C>
__global__ void __launch_bounds__(256) chan_dot_fwd_kernel(const float* __restrict__ y, const float* __restrict__ w,
                                                           const float* __restrict__ bias, float* __restrict__ out, long long M) {
  __shared__ float ws[C];
  if (threadIdx.x < C) ws[threadIdx.x] = w[threadIdx.x];
  __syncthreads();
  const float b = bias ? bias[0] : 0.f;
  for (long long m = (long long)blockIdx.x * 256 + threadIdx.x; m < M; m += (long long)gridDim.x * 256) {
    const float4* row = reinterpret_cast<const float4*>(y + m * C);
    float acc = b;
#pragma unroll
    for (int j = 0; j < C / 4; ++j) {
      const float4 v = row[j];
      acc += v.x * ws[4 * j] + v.y * ws[4 * j + 1] + v.z * ws[4 * j + 2] + v.w * ws[4 * j + 3];
    }
    out[m] = acc;
  }
}

// dy[m,c] = dout[m] w[c];  dw[c] += sum_m dout[m] y[m,c];  db += sum_m dout[m]   (dw, db zero-filled by the caller)
template <int C>
__global__ void __launch_bounds__(256) chan_dot_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ y,
                                                           const float* __restrict__ w, float* __restrict__ dy,
                                                           float* __restrict__ dw, float* __restrict__ db, long long M) {
  __shared__ float ws[C];
  __shared__ float red[C + 1];
  if (threadIdx.x < C) ws[threadIdx.x] = w[threadIdx.x];
  if (threadIdx.x <= C) red[threadIdx.x] = 0.f;
  __syncthreads();
  float acc[C];
  float accb = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = 0.f;
  for (long long m = (long long)blockIdx.x * 256 + threadIdx.x; m < M; m += (long long)gridDim.x * 256) {
    const float g = dout[m];
    accb += g;
    const float4* row = reinterpret_cast<const float4*>(y + m * C);
    float4* drow = reinterpret_cast<float4*>(dy + m * C);
#pragma unroll
    for (int j = 0; j < C / 4; ++j) {
      const float4 v = row[j];
      acc[4 * j] += g * v.x; acc[4 * j + 1] += g * v.y; acc[4 * j + 2] += g * v.z; acc[4 * j + 3] += g * v.w;
      drow[j] = make_float4(g * ws[4 * j], g * ws[4 * j + 1], g * ws[4 * j + 2], g * ws[4 * j + 3]);
    }
  }
#pragma unroll
  for (int c = 0; c < C; ++c) {
    float v = acc[c];
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&red[c], v);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) accb += __shfl_xor_sync(0xffffffffu, accb, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(&red[C], accb);
  __syncthreads();
  if (threadIdx.x < C) atomicAdd(&dw[threadIdx.x], red[threadIdx.x]);
  if (threadIdx.x == C) atomicAdd(db, red[C]);
}

}  // namespace hvc

using namespace hvc;

extern "C" int hvc_im2col2d(const void* x, int32_t x_is_bf16, const hvc_conv2d_geom* geom, void* cols, int32_t cols_is_f32, void* stream) {
  HVC_CHECK_ARG(x && geom && cols, "hvc_im2col2d: null operand");
  HVC_CHECK_ARG(!(x_is_bf16 && cols_is_f32), "hvc_im2col2d: an f32 patch matrix needs an f32 input");
  Conv2dGeom g;
  int rc = fill_geom2d(&g, geom);
  if (rc) return rc;
  const long long total = (long long)g.N * g.Ho * g.Wo * g.Kp;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (x_is_bf16) im2col2d_kernel<bf16, bf16><<<blocks, 256, 0, st>>>(reinterpret_cast<const bf16*>(x), reinterpret_cast<bf16*>(cols), g);
  else if (cols_is_f32) im2col2d_kernel<float, float><<<blocks, 256, 0, st>>>(reinterpret_cast<const float*>(x), reinterpret_cast<float*>(cols), g);
  else im2col2d_kernel<float, bf16><<<blocks, 256, 0, st>>>(reinterpret_cast<const float*>(x), reinterpret_cast<bf16*>(cols), g);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

extern "C" int hvc_im2col2d_split(const float* x, const hvc_conv2d_geom* geom, void* out, void* stream) {
  HVC_CHECK_ARG(x && geom && out, "hvc_im2col2d_split: null operand");
  Conv2dGeom g;
  int rc = fill_geom2d(&g, geom);
  if (rc) return rc;
  const long long total = (long long)g.N * g.Ho * g.Wo * (g.Kp / 8);
  im2col2d_split_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, reinterpret_cast<bf16*>(out), g);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

extern "C" int hvc_col2im2d(const void* dcols, const hvc_conv2d_geom* geom, float* dx, void* stream) {
  HVC_CHECK_ARG(dcols && geom && dx, "hvc_col2im2d: null operand");
  Conv2dGeom g;
  int rc = fill_geom2d(&g, geom);
  if (rc) return rc;
  const long long total = (long long)g.N * g.Cin * g.H * g.W;
  col2im2d_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const bf16*>(dcols), dx, g);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

extern "C" int hvc_maxpool2d_fwd(const void* x, void* y, uint8_t* arg, int32_t N, int32_t H, int32_t W, int32_t C, int32_t k, int32_t stride,
                                 int32_t pad, void* stream) {
  HVC_CHECK_ARG(x && y && arg && N > 0 && H > 0 && W > 0 && C > 0 && k >= 1 && k <= 3 && stride >= 1 && pad >= 0 && pad < k,
                "hvc_maxpool2d_fwd: bad arguments");
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  const long long total = (long long)N * Ho * Wo * C;
  maxpool2d_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float*>(x), reinterpret_cast<float*>(y), arg, N, H, W, C, Ho, Wo, k, stride, pad);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

extern "C" int hvc_maxpool2d_bwd(const float* dy, const uint8_t* arg, float* dx, int32_t N, int32_t H, int32_t W, int32_t C, int32_t k,
                                 int32_t stride, int32_t pad, void* stream) {
  HVC_CHECK_ARG(dy && arg && dx && N > 0 && H > 0 && W > 0 && C > 0 && k >= 1 && k <= 3 && stride >= 1 && pad >= 0 && pad < k,
                "hvc_maxpool2d_bwd: bad arguments");
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  HVC_CHECK_ARG((C & 3) == 0, "hvc_maxpool2d_bwd: C must be a multiple of 4");
  const long long total = (long long)N * H * W * (C / 4);
  maxpool2d_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dy, arg, dx, N, H, W, C, Ho, Wo, k,
                                                                                                            stride, pad);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

extern "C" int hvc_view_mean_fwd(const float* x, float* feat, float* pooled, int32_t B, int32_t V, int32_t P, int32_t C, void* stream) {
  HVC_CHECK_ARG(x && feat && B > 0 && V > 0 && P > 0 && C > 0 && (C & 3) == 0 && C <= 1024, "hvc_view_mean_fwd: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long n = (long long)P * C;
  view_mean_kernel<<<(unsigned)(((long long)B * n / 4 + 255) / 256), 256, 0, st>>>(x, feat, B, V, n);
  HVC_LAUNCH_CHECK();
  if (pooled) {
    HVC_CUDA(cudaMemsetAsync(pooled, 0, sizeof(float) * B * C, st));
    const int rpp = 256 / (C / 4);
    int blocks = (P + rpp - 1) / rpp;
    if (blocks > 64) blocks = 64;
    pool_mean_kernel<<<dim3(blocks, B), 256, 0, st>>>(feat, pooled, P, C);
    HVC_LAUNCH_CHECK();
  }
  return HVC_OK;
}

extern "C" int hvc_view_mean_bwd(const float* dfeat, const float* dpooled, float* dx, int32_t B, int32_t V, int32_t P, int32_t C, void* stream) {
  HVC_CHECK_ARG(dx && (dfeat || dpooled) && B > 0 && V > 0 && P > 0 && C > 0 && (C & 3) == 0, "hvc_view_mean_bwd: bad arguments");
  const long long n = (long long)P * C;
  view_mean_bwd_kernel<<<(unsigned)(((long long)B * n / 4 + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dfeat, dpooled, dx, B,
                                                                                                                           V, P, C);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

extern "C" int hvc_silu(const float* x, const float* dy, float* out, int64_t n, void* stream) {
  HVC_CHECK_ARG(x && out && n > 0, "hvc_silu: bad arguments");
  silu_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, dy, out, n);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

extern "C" int hvc_chan_dot_fwd(const float* y, const float* w, const float* bias, float* out, int64_t M, int32_t C, void* stream) {
  HVC_CHECK_ARG(y && w && out && M > 0, "hvc_chan_dot_fwd: bad arguments");
  HVC_CHECK_ARG(C == 32 || C == 64, "hvc_chan_dot_fwd: C must be 32 or 64");
  const unsigned blocks = (unsigned)std::min<long long>((M + 255) / 256, 148 * 16);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (C == 32) chan_dot_fwd_kernel<32><<<blocks, 256, 0, st>>>(y, w, bias, out, M);
  else chan_dot_fwd_kernel<64><<<blocks, 256, 0, st>>>(y, w, bias, out, M);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

extern "C" int hvc_chan_dot_bwd(const float* dout, const float* y, const float* w, float* dy, float* dw, float* db, int64_t M,
                                int32_t C, void* stream) {
  HVC_CHECK_ARG(dout && y && w && dy && dw && db && M > 0, "hvc_chan_dot_bwd: bad arguments");
  HVC_CHECK_ARG(C == 32 || C == 64, "hvc_chan_dot_bwd: C must be 32 or 64");
  const unsigned blocks = (unsigned)std::min<long long>((M + 255) / 256, 148 * 4);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (C == 32) chan_dot_bwd_kernel<32><<<blocks, 256, 0, st>>>(dout, y, w, dy, dw, db, M);
  else chan_dot_bwd_kernel<64><<<blocks, 256, 0, st>>>(dout, y, w, dy, dw, db, M);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}
