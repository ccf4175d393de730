// hvc_host.h -- host-side helpers shared by the translation units of libhvc_sm100a.so
// (error reporting across the C ABI, TMA tensor-map encoding, launch accounting).
#pragma once
#include <atomic>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/hvc.h"

namespace hvc {

// ---- error channel: no exception crosses the ABI; hvc_last_error() returns this thread's message
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define HVC_FAIL(code, ...)      \
  do {                           \
    hvc::set_error(__VA_ARGS__); \
    return (code);               \
  } while (0)

#define HVC_CHECK_ARG(cond, ...) \
  do {                           \
    if (!(cond)) HVC_FAIL(HVC_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define HVC_CUDA(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      HVC_FAIL(HVC_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define HVC_LAUNCH_CHECK()                                                                  \
  do {                                                                                      \
    cudaError_t _e = cudaGetLastError();                                                    \
    if (_e != cudaSuccess)                                                                  \
      HVC_FAIL(HVC_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
    hvc::count_launch();                                                                    \
  } while (0)

// ---- TMA descriptors.  2-D row-major tensor [rows, cols] of `elem_bytes` elements with leading
// dimension `ld` (elements); box = box_cols x box_rows; swizzle: 0 none, 1 (true) 128-byte, 2 64-byte, 3 32-byte.
int make_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t rows, uint64_t cols, uint64_t ld,
                 uint32_t box_cols, uint32_t box_rows, int swizzle);

int device_sm_count();

// ---- opt-in to > 48 KB of dynamic shared memory.  cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute of the
// function: a process that touches a second GPU must set it there too.  `done` is one bit per device ordinal (a function-local
// static std::atomic<uint64_t> at each call site); a racing second thread at worst sets the attribute twice, which is harmless.
int smem_opt_in(const void* func, int bytes, std::atomic<uint64_t>& done);
#define HVC_SMEM_OPT_IN(kernel, bytes)                                                  \
  do {                                                                                  \
    static std::atomic<uint64_t> _done{0};                                              \
    int _r = hvc::smem_opt_in(reinterpret_cast<const void*>(kernel), (bytes), _done);   \
    if (_r != HVC_OK) return _r;                                                        \
  } while (0)

}  // namespace hvc
