// hvc_loss_multiscale.cu -- the stage 2-3 loss terms of the progressive cascade:
//   TotalVariationLoss     direct_regression/progressive_cascade/loss_multiscale.py:140-188
//   FrequencyLoss          :191-236   (the 3-D FFT itself is cuFFT through torch.fft -- a library transform, like cuBLAS for a plain GEMM;
//                                      everything around it -- magnitudes, the radial mask, the two masked L1 sums, the gradient -- is here)
//   DRRReprojectionLoss    :239-293   (mean-intensity projections along depth and width; the bilinear resize to the X-ray size is
//                                      hvc_interp3d with a unit depth; the L1 against the X-rays is hvc_l1_*)
// All of it is HBM-bound elementwise / reduction work on fp32 volumes; sums are accumulated in double.
#include <algorithm>

#include "hvc_common.cuh"
#include "hvc_host.h"

namespace hvc {

__device__ __forceinline__ void block_add2(double a, double b, double* out0, double* out1) {
  __shared__ double red[2][8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { a += red[0][w]; b += red[1][w]; }
    atomicAdd(out0, a);
    if (out1 != nullptr) atomicAdd(out1, b);
  }
  __syncthreads();
}

// ---- total variation: sums[k] += sum sqrt((x[i + stride_k] - x[i])^2 + eps), k = depth, height, width (loss_multiscale.py:162-170)
__global__ void __launch_bounds__(256) tv_fwd_kernel(const float* __restrict__ x, long long n, int D, int H, int W, float eps, double* __restrict__ sums) {
  double sd = 0.0, sh = 0.0, sw = 0.0;
  const long long HW = (long long)H * W;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const int w = (int)(i % W), h = (int)((i / W) % H), d = (int)((i / HW) % D);
    const float v = __ldg(x + i);
    if (d + 1 < D) { const float u = __ldg(x + i + HW) - v; sd += (double)sqrtf(u * u + eps); }
    if (h + 1 < H) { const float u = __ldg(x + i + W) - v; sh += (double)sqrtf(u * u + eps); }
    if (w + 1 < W) { const float u = __ldg(x + i + 1) - v; sw += (double)sqrtf(u * u + eps); }
  }
  block_add2(sd, sh, sums, sums + 1);
  block_add2(sw, 0.0, sums + 2, nullptr);
}

// value: tv(x) = clamp((sum_d / n_d + sum_h / n_h + sum_w / n_w) / 3, 0, 100); loss = tv_pred (sums_t == NULL) or |tv_pred - tv_target|
// coef (device, f32[1]) = d loss / d tv_pred  (0 outside the clamp)
__global__ void tv_finalize_kernel(const double* __restrict__ sums_p, const double* __restrict__ sums_t, double nd, double nh, double nw,
                                   float* __restrict__ loss, float* __restrict__ coef) {
  const double raw_p = (sums_p[0] / nd + sums_p[1] / nh + sums_p[2] / nw) / 3.0;
  const double tp = fmin(fmax(raw_p, 0.0), 100.0);
  const double inside = (raw_p > 0.0 && raw_p < 100.0) ? 1.0 : 0.0;
  if (sums_t == nullptr) {
    loss[0] = (float)tp;
    coef[0] = (float)inside;
    return;
  }
  const double raw_t = (sums_t[0] / nd + sums_t[1] / nh + sums_t[2] / nw) / 3.0;
  const double tt = fmin(fmax(raw_t, 0.0), 100.0);
  loss[0] = (float)fabs(tp - tt);
  coef[0] = (float)(inside * (tp > tt ? 1.0 : (tp < tt ? -1.0 : 0.0)));
}

// dx[i] = upstream * coef * sum_k (1 / (3 n_k)) * (g(x[i] - x[i - s_k]) - g(x[i + s_k] - x[i])),  g(u) = u / sqrt(u^2 + eps)
__global__ void __launch_bounds__(256) tv_bwd_kernel(const float* __restrict__ x, long long n, int D, int H, int W, float eps, float cd, float ch, float cw,
                                                     const float* __restrict__ coef, const float* __restrict__ upstream, float* __restrict__ dx) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const long long HW = (long long)H * W;
  const int w = (int)(i % W), h = (int)((i / W) % H), d = (int)((i / HW) % D);
  const float v = __ldg(x + i);
  float acc = 0.f;
  auto g = [eps](float u) { return u * rsqrtf(u * u + eps); };
  if (d > 0) acc += cd * g(v - __ldg(x + i - HW));
  if (d + 1 < D) acc -= cd * g(__ldg(x + i + HW) - v);
  if (h > 0) acc += ch * g(v - __ldg(x + i - W));
  if (h + 1 < H) acc -= ch * g(__ldg(x + i + W) - v);
  if (w > 0) acc += cw * g(v - __ldg(x + i - 1));
  if (w + 1 < W) acc -= cw * g(__ldg(x + i + 1) - v);
  const float s = __ldg(coef) * (upstream != nullptr ? __ldg(upstream) : 1.f);
  dx[i] = acc * s;
}

// ---- frequency loss: spectra as interleaved complex64 [B, D, H, W]; mask = dist((d,h,w) - (D/2,H/2,W/2)) > min(D,H,W)/4 on the
// UNSHIFTED spectrum, exactly as loss_multiscale.py:214-229 builds it.  sums[0] += sum_{!mask} | |Fp| - |Ft| |, sums[1] += sum_{mask} ...
__device__ __forceinline__ bool freq_high(long long i, int D, int H, int W, float radius) {
  const int w = (int)(i % W), h = (int)((i / W) % H), d = (int)((i / ((long long)H * W)) % D);
  const float dd = (float)(d - D / 2), hh = (float)(h - H / 2), ww = (float)(w - W / 2);
  return sqrtf(dd * dd + hh * hh + ww * ww) > radius;
}
__global__ void __launch_bounds__(256) freq_fwd_kernel(const float2* __restrict__ fp, const float2* __restrict__ ft, long long n, int D, int H, int W,
                                                       float radius, double* __restrict__ sums) {
  double lo = 0.0, hi = 0.0;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const float2 a = __ldg(fp + i), b = __ldg(ft + i);
    const float diff = fabsf(hypotf(a.x, a.y) - hypotf(b.x, b.y));
    if (freq_high(i, D, H, W, radius)) hi += (double)diff; else lo += (double)diff;
  }
  block_add2(lo, hi, sums, sums + 1);
}
// G = upstream * (c_lo | c_hi) * sign(|Fp| - |Ft|) * Fp / |Fp|   (0 where |Fp| == 0, like torch.abs of a complex zero)
__global__ void __launch_bounds__(256) freq_bwd_kernel(const float2* __restrict__ fp, const float2* __restrict__ ft, long long n, int D, int H, int W,
                                                       float radius, float c_lo, float c_hi, const float* __restrict__ upstream, float2* __restrict__ g) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float2 a = __ldg(fp + i), b = __ldg(ft + i);
  const float ma = hypotf(a.x, a.y), mb = hypotf(b.x, b.y);
  float c = freq_high(i, D, H, W, radius) ? c_hi : c_lo;
  if (upstream != nullptr) c *= __ldg(upstream);
  const float sg = ma > mb ? 1.f : (ma < mb ? -1.f : 0.f);
  const float s = ma > 0.f ? c * sg / ma : 0.f;
  g[i] = make_float2(a.x * s, a.y * s);
}

// ---- DRR projections (loss_multiscale.py:249-271): ap[b,h,w] = mean_d vol[b,d,h,w]; lat[b,d,h] = mean_w vol[b,d,h,w]
__global__ void __launch_bounds__(256) proj_ap_kernel(const float* __restrict__ vol, float* __restrict__ ap, int B, int D, long long HW) {
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t >= (long long)B * HW) return;
  const long long b = t / HW, r = t - b * HW;
  const float* p = vol + b * D * HW + r;
  float acc = 0.f;
  for (int d = 0; d < D; ++d) acc += __ldg(p + d * HW);
  ap[t] = acc / (float)D;
}
__global__ void __launch_bounds__(256) proj_lat_kernel(const float* __restrict__ vol, float* __restrict__ lat, long long rows, int W) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);     // one warp per (b, d, h) row
  if (row >= rows) return;
  const float* p = vol + row * W;
  float acc = 0.f;
  for (int w = threadIdx.x & 31; w < W; w += 32) acc += __ldg(p + w);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) lat[row] = acc / (float)W;
}
// dvol[b,d,h,w] = dap[b,h,w] / D + dlat[b,d,h] / W
__global__ void __launch_bounds__(256) proj_bwd_kernel(const float* __restrict__ dap, const float* __restrict__ dlat, float* __restrict__ dvol, long long n,
                                                       int D, int H, int W) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const long long HW = (long long)H * W;
  const long long b = i / (D * HW);
  const long long r = i - b * D * HW;
  const int d = (int)(r / HW);
  const long long hw = r - d * HW;
  const int h = (int)(hw / W);
  dvol[i] = __ldg(dap + b * HW + hw) / (float)D + __ldg(dlat + (b * D + d) * H + h) / (float)W;
}

// ---- plain L1: sum[0] += sum |a - b|;  da = upstream * c * sign(a - b)
__global__ void __launch_bounds__(256) l1_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, double* __restrict__ sum) {
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) s += (double)fabsf(__ldg(a + i) - __ldg(b + i));
  block_add2(s, 0.0, sum, nullptr);
}
__global__ void __launch_bounds__(256) l1_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, float c,
                                                     const float* __restrict__ upstream, float* __restrict__ da) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float d = __ldg(a + i) - __ldg(b + i);
  if (upstream != nullptr) c *= __ldg(upstream);
  da[i] = d > 0.f ? c : (d < 0.f ? -c : 0.f);
}

static unsigned red_blocks(long long n) {
  const long long want = (n + 255) / 256;
  return (unsigned)std::max<long long>(1, std::min<long long>(want, 8LL * device_sm_count()));
}

}  // namespace hvc

using namespace hvc;

extern "C" int hvc_tv_fwd(const float* x, int32_t B, int32_t D, int32_t H, int32_t W, float eps, double* sums, void* stream) {
  HVC_CHECK_ARG(x && sums && B > 0 && D > 0 && H > 0 && W > 0, "hvc_tv_fwd: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long n = (long long)B * D * H * W;
  HVC_CUDA(cudaMemsetAsync(sums, 0, 3 * sizeof(double), st));
  tv_fwd_kernel<<<red_blocks(n), 256, 0, st>>>(x, n, D, H, W, eps, sums);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

extern "C" int hvc_tv_finalize(const double* sums_pred, const double* sums_target, int32_t B, int32_t D, int32_t H, int32_t W, float* loss, float* coef,
                               void* stream) {
  HVC_CHECK_ARG(sums_pred && loss && coef && B > 0 && D > 1 && H > 1 && W > 1, "hvc_tv_finalize: bad arguments (every axis needs at least 2 samples)");
  const double nd = (double)B * (D - 1) * H * W, nh = (double)B * D * (H - 1) * W, nw = (double)B * D * H * (W - 1);
  tv_finalize_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream)>>>(sums_pred, sums_target, nd, nh, nw, loss, coef);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

extern "C" int hvc_tv_bwd(const float* x, int32_t B, int32_t D, int32_t H, int32_t W, float eps, const float* coef, const float* upstream, float* dx,
                          void* stream) {
  HVC_CHECK_ARG(x && coef && dx && B > 0 && D > 1 && H > 1 && W > 1, "hvc_tv_bwd: bad arguments");
  const long long n = (long long)B * D * H * W;
  const float cd = (float)(1.0 / (3.0 * (double)B * (D - 1) * H * W)), ch = (float)(1.0 / (3.0 * (double)B * D * (H - 1) * W)),
              cw = (float)(1.0 / (3.0 * (double)B * D * H * (W - 1)));
  tv_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, n, D, H, W, eps, cd, ch, cw, coef, upstream, dx);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

extern "C" int hvc_freq_l1_fwd(const float* spec_pred, const float* spec_target, int32_t B, int32_t D, int32_t H, int32_t W, double* sums, void* stream) {
  HVC_CHECK_ARG(spec_pred && spec_target && sums && B > 0 && D > 0 && H > 0 && W > 0, "hvc_freq_l1_fwd: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long n = (long long)B * D * H * W;
  const float radius = (float)(std::min(D, std::min(H, W)) / 4);
  HVC_CUDA(cudaMemsetAsync(sums, 0, 2 * sizeof(double), st));
  freq_fwd_kernel<<<red_blocks(n), 256, 0, st>>>(reinterpret_cast<const float2*>(spec_pred), reinterpret_cast<const float2*>(spec_target), n, D, H, W,
                                                  radius, sums);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

extern "C" int hvc_freq_l1_bwd(const float* spec_pred, const float* spec_target, int32_t B, int32_t D, int32_t H, int32_t W, float c_low, float c_high,
                               const float* upstream, float* dspec, void* stream) {
  HVC_CHECK_ARG(spec_pred && spec_target && dspec && B > 0 && D > 0 && H > 0 && W > 0, "hvc_freq_l1_bwd: bad arguments");
  const long long n = (long long)B * D * H * W;
  const float radius = (float)(std::min(D, std::min(H, W)) / 4);
  freq_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float2*>(spec_pred), reinterpret_cast<const float2*>(spec_target), n, D, H, W, radius, c_low, c_high, upstream,
      reinterpret_cast<float2*>(dspec));
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

extern "C" int hvc_proj_mean_fwd(const float* vol, int32_t B, int32_t D, int32_t H, int32_t W, float* ap, float* lat, void* stream) {
  HVC_CHECK_ARG(vol && ap && lat && B > 0 && D > 0 && H > 0 && W > 0, "hvc_proj_mean_fwd: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long HW = (long long)H * W, rows = (long long)B * D * H;
  proj_ap_kernel<<<(unsigned)((B * HW + 255) / 256), 256, 0, st>>>(vol, ap, B, D, HW);
  HVC_LAUNCH_CHECK();
  proj_lat_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(vol, lat, rows, W);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

extern "C" int hvc_proj_mean_bwd(const float* dap, const float* dlat, int32_t B, int32_t D, int32_t H, int32_t W, float* dvol, void* stream) {
  HVC_CHECK_ARG(dap && dlat && dvol && B > 0 && D > 0 && H > 0 && W > 0, "hvc_proj_mean_bwd: bad arguments");
  const long long n = (long long)B * D * H * W;
  proj_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dap, dlat, dvol, n, D, H, W);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

extern "C" int hvc_l1_fwd(const float* a, const float* b, int64_t n, double* sum, void* stream) {
  HVC_CHECK_ARG(a && b && sum && n > 0, "hvc_l1_fwd: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  HVC_CUDA(cudaMemsetAsync(sum, 0, sizeof(double), st));
  l1_fwd_kernel<<<red_blocks(n), 256, 0, st>>>(a, b, n, sum);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

extern "C" int hvc_l1_bwd(const float* a, const float* b, int64_t n, float c, const float* upstream, float* da, void* stream) {
  HVC_CHECK_ARG(a && b && da && n > 0, "hvc_l1_bwd: bad arguments");
  l1_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a, b, n, c, upstream, da);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}
