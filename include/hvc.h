/* hvc.h -- C ABI of libhvc_sm100a.so, the B200 (sm_100a) kernels behind the Hybrid-ViT-Cascade
 * 3D ViT backbone hot path.
 *
 * The reference (kanadm12/Hybrid-ViT-Cascade) has no FFI: its hot path is torch.nn modules
 * (models/vit_components.py, models/hybrid_vit_backbone.py) dispatching to ATen/cuBLAS/cuDNN.
 * The drop-in boundary is therefore the Python module API + state_dict; this
 * header is the level below it, the calls the package's autograd Functions bind with ctypes
 * (hybrid_vit_cascade_b200/_lib.py).  Each entry point cites the reference lines whose ATen
 * call(s) it replaces.  See INTEGRATION.md for the reference-side binding.
 *
 * Conventions
 *   - plain C: raw device pointers, explicit sizes / leading dimensions (in ELEMENTS), no torch
 *     types, no exceptions.  Every function returns HVC_OK (0) or a negative hvc_status and leaves
 *     a message for hvc_last_error() (thread-local).
 *   - all device memory is owned by the caller (torch); the library allocates nothing on the
 *     device.  Work is enqueued asynchronously on `stream` (a cudaStream_t passed as void*).
 *   - argument structs start with `size` = sizeof(struct) for versioning.
 *   - bf16 = __nv_bfloat16 storage, f32 = float.  "tokens" T = batch * tokens-per-sample; token-major
 *     tensors are [T, C] row-major.
 *   - there is no CPU fallback and no other architecture: on a device that is not sm_100 every
 *     compute entry point fails with HVC_ERR_ARCH.
 */
#ifndef HVC_H_
#define HVC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HVC_VERSION 100

typedef enum hvc_status {
  HVC_OK = 0,
  HVC_ERR_INVALID = -1, /* bad argument / unsupported shape */
  HVC_ERR_CUDA = -2,    /* CUDA runtime / driver error      */
  HVC_ERR_ARCH = -3     /* not running on sm_100            */
} hvc_status;

int hvc_version(void);
const char* hvc_last_error(void);
/* Number of kernels this library has launched in this process (bench.py reports it as gpu_launches). */
uint64_t hvc_launch_count(void);
/* 0 when the current device is sm_100 and the TMA driver entry point resolves. */
int hvc_check_device(void);

/* ------------------------------------------------------------------------------------------------
 * GEMM on tcgen05 tensor cores (TMA -> 128B-swizzled smem -> tcgen05.mma -> TMEM -> epilogue).
 *   D[M,N] = alpha * sum_k A(m,k) * B(n,k)   (bf16 operands, fp32 accumulation)
 * Replaces every nn.Linear on the path and their backward GEMMs:
 *   forward  y = x W^T      : A = x [T,K] (a_major 0), B = W [N,K] (b_major 0)
 *            vit_components.py:41,54,95,98,116; hybrid_vit_backbone.py:76,79
 *   dgrad    dx = dy W      : A = dy [T,N'] (a_major 0), B = W stored [N',K'] read as (k',n') -> b_major 1
 *   wgrad    dW = dy^T x    : A = dy stored [T,N'] -> a_major 1, B = x stored [T,K'] -> b_major 1,
 *            reduction over T split across CTAs (k_splits), fp32 atomic accumulation.
 * a_major/b_major: 0 = operand stored [rows, K] (K contiguous); 1 = stored [K, rows] (rows contiguous).
 * Requirements: lda/ldb multiples of 8 elements, 16-byte aligned bases.
 * ---------------------------------------------------------------------------------------------- */
typedef enum hvc_epilogue {
  HVC_EPI_BF16 = 0,       /* out bf16 = act(alpha*acc + bias) [* f'(aux)]; optional out2 bf16 = pre-activation */
  HVC_EPI_RESIDUAL = 1,   /* out f32 = resid + gate[row/rows_per_batch, col] * (acc + bias); out2 bf16 = acc+bias */
  HVC_EPI_F32_ATOMIC = 2, /* out f32 += alpha*acc  (red.global.add; used with k_splits >= 1)                 */
  HVC_EPI_F32 = 3         /* out f32 = alpha*acc + bias                                                       */
} hvc_epilogue;

typedef enum hvc_activation {
  HVC_ACT_NONE = 0,
  HVC_ACT_GELU = 1,       /* exact erf GELU (nn.GELU(), hybrid_vit_backbone.py:77)                     */
  HVC_ACT_GELU_GRAD = 2   /* out = acc * gelu'(aux)  (backward of the MLP hidden activation)          */
} hvc_activation;

typedef struct hvc_gemm_args {
  uint32_t size;
  int32_t M, N, K;
  const void* A; int64_t lda; int32_t a_major;
  const void* B; int64_t ldb; int32_t b_major;
  int32_t epilogue;             /* hvc_epilogue   */
  int32_t activation;           /* hvc_activation */
  void* out; int64_t ldo;       /* bf16 or f32 by epilogue */
  void* out2; int64_t ldo2;     /* optional secondary bf16 output (NULL to skip) */
  const float* bias;            /* [N] or NULL */
  const float* resid; int64_t ldr;        /* f32 [M,N]  (HVC_EPI_RESIDUAL) */
  const float* gate; int64_t gate_ld;     /* f32, gate[(row / rows_per_batch) * gate_ld + col]; NULL = 1 */
  int32_t rows_per_batch;
  const void* aux; int64_t ldaux;         /* bf16 [M,N] elementwise operand (HVC_ACT_GELU_GRAD) */
  float alpha;
  int32_t k_splits;             /* >= 1; > 1 only with HVC_EPI_F32_ATOMIC */
} hvc_gemm_args;

int hvc_gemm(const hvc_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HVC_H_ */
