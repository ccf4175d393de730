// hvc_loss.cu -- the direct-regression training loss:
//   DirectRegressionLoss = l1_weight * L1 + ssim_weight * (1 - mean SSIM3D), direct_regression/model_direct.py:88-131.
// compute_ssim_loss (:88-107) filters five volumes (pred, target, pred^2, target^2, pred*target) with an 11^3 box
// (F.avg_pool3d, stride 1, zero padding counted in the divisor).  The box is separable and symmetric: three 1-D passes per
// filter, the five (forward) / three (backward) volumes stacked so each pass is one launch, and the backward filter is the
// same operator.  Everything is HBM / L2 bound elementwise work.
#include "hvc_common.cuh"
#include "hvc_host.h"

namespace hvc {

constexpr float kC1 = 0.01f * 0.01f, kC2 = 0.03f * 0.03f;

// out[0:n] = p, [n:2n] = t, [2n:3n] = p^2, [3n:4n] = t^2, [4n:5n] = p t
__global__ void __launch_bounds__(256) ssim_stack_kernel(const float* __restrict__ p, const float* __restrict__ t, float* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float a = p[i], b = t[i];
  out[i] = a; out[n + i] = b; out[2 * n + i] = a * a; out[3 * n + i] = b * b; out[4 * n + i] = a * b;
}

// 1-D box of width k (odd) along one axis of a stack of volumes: out[i] = (1/k) * sum_{|j| <= k/2} x[i + j*stride] inside the axis
__global__ void __launch_bounds__(256) box1d_kernel(const float* __restrict__ x, float* __restrict__ out, long long total, int len, long long stride,
                                                    int k) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const int pos = (int)((i / stride) % len);
  const int r = k >> 1;
  const int lo = max(-r, -pos), hi = min(r, len - 1 - pos);
  float acc = 0.f;
  for (int j = lo; j <= hi; ++j) acc += __ldg(x + i + j * stride);
  out[i] = acc * (1.0f / k);
}

// The same filter with the window kept in registers (bit-identical sums: same ascending order, absent positions add 0.0f).
// Along W (contiguous axis): one thread produces CH consecutive outputs from CH + K - 1 loads.
template <int K, int CH>
__global__ void __launch_bounds__(256) box1d_row_kernel(const float* __restrict__ x, float* __restrict__ out, long long chunks, int len,
                                                        int chunks_per_row) {
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t >= chunks) return;
  const long long row = t / chunks_per_row;
  const int c0 = (int)(t - row * chunks_per_row) * CH;
  const float* px = x + row * len;
  float v[CH + K - 1];
#pragma unroll
  for (int j = 0; j < CH + K - 1; ++j) {
    const int pos = c0 - K / 2 + j;
    v[j] = (pos >= 0 && pos < len) ? __ldg(px + pos) : 0.f;
  }
#pragma unroll
  for (int o = 0; o < CH; ++o) {
    if (c0 + o >= len) break;
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < K; ++j) acc += v[o + j];
    out[row * len + c0 + o] = acc * (1.0f / K);
  }
}
// Along H or D (element stride `inner`): one thread walks one line with a K-deep register window; neighbouring threads are
// neighbouring `inner` positions, so every load and store is coalesced and each input is read once.
template <int K>
__global__ void __launch_bounds__(256) box1d_line_kernel(const float* __restrict__ x, float* __restrict__ out, long long lines, int len,
                                                         long long inner) {
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t >= lines) return;
  const long long o = t / inner, base = o * len * inner + (t - o * inner);
  const float* px = x + base;
  float* po = out + base;
  constexpr int R = K / 2;
  float w[K];
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const int pos = j - R;
    w[j] = (pos >= 0 && pos < len) ? __ldg(px + pos * inner) : 0.f;
  }
  for (int p = 0; p < len; ++p) {
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < K; ++j) acc += w[j];
    po[p * inner] = acc * (1.0f / K);
#pragma unroll
    for (int j = 0; j < K - 1; ++j) w[j] = w[j + 1];
    const int np = p + 1 + R;
    w[K - 1] = np < len ? __ldg(px + np * inner) : 0.f;
  }
}

__device__ __forceinline__ void ssim_terms(float mp, float mt, float epp, float ett, float ept, float& A1, float& A2, float& B1, float& B2) {
  A1 = 2.f * mp * mt + kC1;
  A2 = 2.f * (ept - mp * mt) + kC2;
  B1 = mp * mp + mt * mt + kC1;
  B2 = (epp - mp * mp) + (ett - mt * mt) + kC2;
}

// sums[0] += sum SSIM, sums[1] += sum |p - t|     (double accumulators)
__global__ void __launch_bounds__(256) ssim_point_fwd_kernel(const float* __restrict__ F, const float* __restrict__ p, const float* __restrict__ t,
                                                             long long n, double* __restrict__ sums) {
  __shared__ double red[2][8];
  double s = 0.0, l = 0.0;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    float A1, A2, B1, B2;
    ssim_terms(F[i], F[n + i], F[2 * n + i], F[3 * n + i], F[4 * n + i], A1, A2, B1, B2);
    s += (double)((A1 * A2) / (B1 * B2));
    l += (double)fabsf(p[i] - t[i]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); l += __shfl_xor_sync(0xffffffffu, l, o); }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s; red[1][threadIdx.x >> 5] = l; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { s += red[0][w]; l += red[1][w]; }
    atomicAdd(sums, s);
    atomicAdd(sums + 1, l);
  }
}

// G[0:n] = dS/d mu_p, G[n:2n] = dS/d E[p^2], G[2n:3n] = dS/d E[p t]
__global__ void __launch_bounds__(256) ssim_point_bwd_kernel(const float* __restrict__ F, float* __restrict__ G, long long n) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float mp = F[i], mt = F[n + i];
  float A1, A2, B1, B2;
  ssim_terms(mp, mt, F[2 * n + i], F[3 * n + i], F[4 * n + i], A1, A2, B1, B2);
  const float inv = 1.f / (B1 * B2);
  const float S = A1 * A2 * inv;
  const float dA1 = A2 * inv, dA2 = A1 * inv, dB1 = -S / B1, dB2 = -S / B2;
  G[i] = 2.f * mt * (dA1 - dA2) + 2.f * mp * (dB1 - dB2);
  G[n + i] = dB2;
  G[2 * n + i] = 2.f * dA2;
}

// dp = c_ssim * (FG0 + 2 p FG1 + t FG2) + c_l1 * sign(p - t)      (FG = box-filtered G)
__global__ void __launch_bounds__(256) ssim_combine_kernel(const float* __restrict__ FG, const float* __restrict__ p, const float* __restrict__ t,
                                                           float* __restrict__ dp, long long n, float c_ssim, float c_l1,
                                                           const float* __restrict__ upstream) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  if (upstream != nullptr) {        // gradient of the scalar loss, read on the device (no host synchronisation)
    const float g = __ldg(upstream);
    c_ssim *= g;
    c_l1 *= g;
  }
  const float a = p[i], b = t[i];
  const float d = a - b;
  const float sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
  dp[i] = c_ssim * (FG[i] + 2.f * a * FG[n + i] + b * FG[2 * n + i]) + c_l1 * sg;
}

}  // namespace hvc

using namespace hvc;

static int box3d(const float* in, float* tmp, float* out, int stack, int D, int H, int W, int k, cudaStream_t st) {
  // three passes: W (in -> out), H (out -> tmp), D (tmp -> out)
  const long long total = (long long)stack * D * H * W;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  if (k == 11) {   // the reference window (model_direct.py:88): register-window kernels
    constexpr int CH = 8;
    const int cpr = (W + CH - 1) / CH;
    const long long chunks = (long long)stack * D * H * cpr;
    box1d_row_kernel<11, CH><<<(unsigned)((chunks + 255) / 256), 256, 0, st>>>(in, out, chunks, W, cpr);
    HVC_LAUNCH_CHECK();
    const long long lines_h = (long long)stack * D * W, lines_d = (long long)stack * H * W;
    box1d_line_kernel<11><<<(unsigned)((lines_h + 255) / 256), 256, 0, st>>>(out, tmp, lines_h, H, W);
    HVC_LAUNCH_CHECK();
    box1d_line_kernel<11><<<(unsigned)((lines_d + 255) / 256), 256, 0, st>>>(tmp, out, lines_d, D, (long long)H * W);
    HVC_LAUNCH_CHECK();
    return HVC_OK;
  }
  box1d_kernel<<<blocks, 256, 0, st>>>(in, out, total, W, 1, k);
  HVC_LAUNCH_CHECK();
  box1d_kernel<<<blocks, 256, 0, st>>>(out, tmp, total, H, W, k);
  HVC_LAUNCH_CHECK();
  box1d_kernel<<<blocks, 256, 0, st>>>(tmp, out, total, D, (long long)H * W, k);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

extern "C" int hvc_ssim_l1_fwd(const float* pred, const float* target, int32_t B, int32_t D, int32_t H, int32_t W, int32_t window, float* filtered,
                               float* scratch, double* sums, void* stream) {
  HVC_CHECK_ARG(pred && target && filtered && scratch && sums && B > 0 && D > 0 && H > 0 && W > 0, "hvc_ssim_l1_fwd: bad arguments");
  HVC_CHECK_ARG(window >= 1 && (window & 1) == 1 && window <= 31, "hvc_ssim_l1_fwd: window must be odd and <= 31");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long n = (long long)B * D * H * W;
  float* stack = scratch;                 // [5n]
  float* tmp = scratch + 5 * n;           // [5n]
  ssim_stack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(pred, target, stack, n);
  HVC_LAUNCH_CHECK();
  int rc = box3d(stack, tmp, filtered, 5 * B, D, H, W, window, st);
  if (rc) return rc;
  HVC_CUDA(cudaMemsetAsync(sums, 0, 2 * sizeof(double), st));
  const int sms = device_sm_count();
  ssim_point_fwd_kernel<<<sms * 8, 256, 0, st>>>(filtered, pred, target, n, sums);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

extern "C" int hvc_ssim_l1_bwd(const float* pred, const float* target, const float* filtered, int32_t B, int32_t D, int32_t H, int32_t W,
                               int32_t window, float c_ssim, float c_l1, const float* upstream, float* scratch, float* dpred, void* stream) {
  HVC_CHECK_ARG(pred && target && filtered && scratch && dpred && B > 0 && D > 0 && H > 0 && W > 0, "hvc_ssim_l1_bwd: bad arguments");
  HVC_CHECK_ARG(window >= 1 && (window & 1) == 1 && window <= 31, "hvc_ssim_l1_bwd: window must be odd and <= 31");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long n = (long long)B * D * H * W;
  float* G = scratch;                     // [3n]
  float* tmp = scratch + 3 * n;           // [3n]
  float* FG = scratch + 6 * n;            // [3n]
  ssim_point_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(filtered, G, n);
  HVC_LAUNCH_CHECK();
  int rc = box3d(G, tmp, FG, 3 * B, D, H, W, window, st);
  if (rc) return rc;
  ssim_combine_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(FG, pred, target, dpred, n, c_ssim, c_l1, upstream);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}
