// hvc_optim.cu -- optimizer step on the flat gradient buckets:
//   torch.nn.utils.clip_grad_norm_(params, max_norm) + torch.optim.AdamW.step()   (train_direct_4gpu.py:72-80, config_direct.json:15-21)
// The data-parallel layer already keeps every gradient as a view into a few flat fp32 buckets (dp.GradientBuckets); with the
// parameters and both moments laid out the same way, the whole step is one sum-of-squares pass per bucket and one fused update pass
// per bucket -- HBM-bound, 16 B read + 12 B written per parameter -- instead of a multi-tensor launch per 100 tensors.  The step
// counter lives on the device so the step can be captured in a CUDA graph.
#include "hvc_common.cuh"
#include "hvc_host.h"

namespace hvc {

// accum[0] += sum x^2   (double accumulator: the result feeds a global norm over ~15 M values)
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ x, long long n, double* __restrict__ accum) {
  __shared__ double red[8];
  double s = 0.0;
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    s += (double)(v.x * v.x + v.y * v.y) + (double)(v.z * v.z + v.w * v.w);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) { const float v = x[(n4 << 2) + threadIdx.x]; s += (double)v * v; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) s += red[w];
    atomicAdd(accum, s);
  }
}

// state[0] = step count (float), incremented once per optimizer step before the update kernels
__global__ void adamw_tick_kernel(float* state) { state[0] += 1.f; }

struct AdamWArgs {
  float lr, beta1, beta2, eps, weight_decay, max_norm;
};
// g' = g * min(1, max_norm / (||g||_2 + 1e-6));  p *= 1 - lr*wd;  m = b1 m + (1-b1) g';  v = b2 v + (1-b2) g'^2;
// p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)            (torch.optim.AdamW, decoupled weight decay)
__global__ void __launch_bounds__(256) adamw_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                         float* __restrict__ v, long long n, const AdamWArgs a,
                                                         const double* __restrict__ sumsq, const float* __restrict__ state) {
  const float t = state[0];
  const float bc1 = 1.f - powf(a.beta1, t), bc2 = 1.f - powf(a.beta2, t);
  float clip = 1.f;
  if (a.max_norm > 0.f) clip = fminf(1.f, a.max_norm / ((float)sqrt(sumsq[0]) + 1e-6f));
  const float step_size = a.lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2), decay = 1.f - a.lr * a.weight_decay;
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    float4 pv = reinterpret_cast<float4*>(p)[i], mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    const float4 gv = __ldg(reinterpret_cast<const float4*>(g) + i);
    float* pp = &pv.x; float* mp = &mv.x; float* vp = &vv.x; const float* gp = &gv.x;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float ge = gp[e] * clip;
      mp[e] = a.beta1 * mp[e] + (1.f - a.beta1) * ge;
      vp[e] = a.beta2 * vp[e] + (1.f - a.beta2) * ge * ge;
      pp[e] = pp[e] * decay - step_size * mp[e] / (sqrtf(vp[e]) * inv_sqrt_bc2 + a.eps);
    }
    reinterpret_cast<float4*>(p)[i] = pv; reinterpret_cast<float4*>(m)[i] = mv; reinterpret_cast<float4*>(v)[i] = vv;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = (n4 << 2) + threadIdx.x;
    const float ge = g[i] * clip;
    const float me = a.beta1 * m[i] + (1.f - a.beta1) * ge, ve = a.beta2 * v[i] + (1.f - a.beta2) * ge * ge;
    m[i] = me; v[i] = ve;
    p[i] = p[i] * decay - step_size * me / (sqrtf(ve) * inv_sqrt_bc2 + a.eps);
  }
}

}  // namespace hvc

using namespace hvc;

extern "C" int hvc_sumsq_f32(const float* x, int64_t n, double* accum, void* stream) {
  HVC_CHECK_ARG(x && accum && n > 0, "hvc_sumsq_f32: bad arguments");
  HVC_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15u) == 0, "hvc_sumsq_f32: x must be 16-byte aligned");
  const long long want = (n / 4 + 255) / 256;
  const unsigned blocks = (unsigned)std::max<long long>(1, std::min<long long>(want, 8LL * device_sm_count()));
  sumsq_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, n, accum);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

extern "C" int hvc_adamw_tick(float* state, void* stream) {
  HVC_CHECK_ARG(state, "hvc_adamw_tick: null state");
  adamw_tick_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream)>>>(state);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}

extern "C" int hvc_adamw_flat(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                              float weight_decay, float max_norm, const double* sumsq, const float* state, void* stream) {
  HVC_CHECK_ARG(p && g && m && v && state && n > 0 && (max_norm <= 0.f || sumsq), "hvc_adamw_flat: bad arguments");
  HVC_CHECK_ARG(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                  reinterpret_cast<uintptr_t>(v)) & 15u) == 0, "hvc_adamw_flat: buffers must be 16-byte aligned");
  AdamWArgs a{lr, beta1, beta2, eps, weight_decay, max_norm};
  const long long want = (n / 4 + 255) / 256;
  const unsigned blocks = (unsigned)std::max<long long>(1, std::min<long long>(want, 8LL * device_sm_count()));
  adamw_flat_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p, g, m, v, n, a, sumsq, state);
  HVC_LAUNCH_CHECK();
  return HVC_OK;
}
