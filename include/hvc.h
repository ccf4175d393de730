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
 * Train-mode dropout (nn.Dropout(p) at vit_components.py:27-29,76-78 and hybrid_vit_backbone.py:78,80) is fused into
 * the kernels: element (row, col) of a site is dropped iff hash(seed, site, row, col) < p * 2^32 and kept values are
 * scaled by 1/(1-p).  `seed` points to two 32-bit words in DEVICE memory (drawn by the caller from its generator,
 * e.g. torch's CUDA generator, so checkpoint recomputation replays them); `site` separates the dropout sites that
 * share one seed.  seed == NULL or p <= 0 disables it (eval mode).  Forward and backward calls of one site must
 * pass the same (seed, site, p).  Rows/cols: attention probabilities -> row = (b*heads + h)*nq + q, col = key;
 * GEMM epilogues / residual backward -> row = output row (token), col = output column (feature).
 * ---------------------------------------------------------------------------------------------- */
typedef struct hvc_dropout {
  const uint32_t* seed;
  uint32_t site;
  float p;
} hvc_dropout;

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
  HVC_EPI_F32 = 3         /* out f32 = act(alpha*acc + bias), act = none | GELU (fp32 verification mode)              */
} hvc_epilogue;

typedef enum hvc_activation {
  HVC_ACT_NONE = 0,
  HVC_ACT_GELU = 1,       /* exact erf GELU (nn.GELU(), hybrid_vit_backbone.py:77)                     */
  HVC_ACT_GELU_GRAD = 2   /* out = acc * gelu'(aux)  (backward of the MLP hidden activation)          */
} hvc_activation;

/* Implicit Conv3d(k3, pad 1) through hvc_gemm: the activation is a zero-padded channels-last volume viewed as a matrix
 * [rows = padded voxels, cin] (bf16), laid out so that every filter tap is a ROW SHIFT offsets[t], which the TMA producer adds to the tile
 * coordinate (rows outside the matrix read as zero).  No patch matrix exists in HBM.
 *   stride 1: rows = B*(D+2)*(H+2)*(W+2) (hvc_pad3d_cl); tap t = kd*9 + kh*3 + kw shifts by (kd-1)*(H+2)*(W+2) + (kh-1)*(W+2) + (kw-1).
 *   stride 2: the input is split by the parity of (d, h, w) into eight half-resolution volumes, each padded by one voxel on the low
 *             side and stacked along the rows (hvc_s2d_pad_cl): output voxel o reads input 2o-1+k, i.e. parity (k != 1) at position
 *             o - (k == 0), so tap t shifts by parity_index(t)*rows_per_volume - (kd==0)*Hp*Wp - (kh==0)*Wp - (kw==0).
 * Column/row index of tap t along the tapped dimension: [t*cin, (t+1)*cin), i.e. weight.permute(0,2,3,4,1).reshape(Cout, 27*cin).
 *   side 1: A is the padded volume (a_major 0, K = n_taps*cin): out[r, n] = sum_t sum_c A[r + offsets[t], c] B[n, t*cin + c]  (forward conv;
 *           data gradient with negated shifts and the transposed filter); rows r at padding positions hold don't-care values.
 *   side 2: B is the padded volume (b_major 1, N = n_taps*cin, K = rows of A): out[m, t*cin + c] = sum_r A(r, m) B[r + offsets[t], c]
 *           (weight gradient; A = padded output gradient with zero rows at the padding positions).
 * rows: row count of the tapped matrix when it differs from M (side 1) / K (side 2), e.g. the eight stacked parity volumes; 0 = same.
 * cin % 64 == 0.  side 0: a plain GEMM. */
typedef struct hvc_conv_taps {
  int32_t side, cin, n_taps, rows;
  int32_t offsets[27];
} hvc_conv_taps;

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
  hvc_dropout drop;             /* applied to act(alpha*acc + bias) before the residual/gate (EPI_BF16, EPI_RESIDUAL, EPI_F32); with
                                   HVC_ACT_GELU_GRAD it masks the incoming gradient: out = drop(acc) * gelu'(aux) */
  hvc_conv_taps taps;           /* implicit 3x3x3 convolution (see above); zero-filled = plain GEMM */
} hvc_gemm_args;

int hvc_gemm(const hvc_gemm_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused flash-style attention (self- and cross-), forward and backward.
 *   out = softmax(q k^T * scale) v      per (batch, head); scores never leave TMEM
 * Replaces vit_components.py:46-51 (self) and :103-113 (cross) and their autograd backward.
 * q/k/v/o are read/written IN PLACE in token-major packed buffers: row = b*n + token, the head's
 * columns are [head*head_dim, (head+1)*head_dim) starting at the given base pointer (so q, k, v may
 * all point into one [T, 3C] qkv projection output; ld* are the row pitches in elements).
 * lse: f32 [batch, heads, nq_pad] base-2 logsumexp of the scaled scores (forward output, backward
 * input), nq_pad = nq rounded up to 128.  delta: f32 [3, batch, heads, nq_pad] scratch written by
 * hvc_attn_bwd (plane 0 = rowsum(dO * O), plane 1 = -lse, plane 2 = dropout row keys).  dq_accum: f32 [batch, heads, nq_pad, head_dim] zero-filled
 * scratch for the cross-CTA dQ reduction.
 * head_dim: 64 or 32.
 * ---------------------------------------------------------------------------------------------- */
typedef struct hvc_attn_args {
  uint32_t size;
  int32_t batch, heads, nq, nk, head_dim;
  const void* q; int64_t ldq;
  const void* k; int64_t ldk;
  const void* v; int64_t ldv;
  void* o; int64_t ldo;           /* bf16 [batch*nq, heads*head_dim] (forward: output; backward: input) */
  float* lse;
  const void* d_o; int64_t lddo;  /* backward only from here on */
  void* dq; int64_t lddq;
  void* dk; int64_t lddk;
  void* dv; int64_t lddv;
  float* delta;
  float* dq_accum;
  void* probs;                    /* optional f32 [batch, heads, nq, nk]: materialised softmax (store_attention), before dropout */
  float scale;
  hvc_dropout drop;               /* attn_drop on the probabilities (vit_components.py:49,110) */
} hvc_attn_args;

int hvc_attn_fwd(const hvc_attn_args* args, void* stream);
int hvc_attn_bwd(const hvc_attn_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused LayerNorm (+AdaLN modulate) -> GEMM operand, and its backward.  HBM-bound, one pass.
 *   y = LN(x; w, b, eps=1e-5) [* (1 + scale[batch]) + shift[batch]]
 * Replaces nn.LayerNorm + the modulate elementwise ops, hybrid_vit_backbone.py:120-121,126,136-137,265.
 * x: f32 [T, C]; y: bf16 (y_is_bf16) or f32 [T, C]; mean/rstd: f32 [T] saved for backward (may be NULL);
 * shift/scale: f32, element (b, c) at [b*mod_ld + c] (both or neither); C multiple of 4, <= 1024.
 * ---------------------------------------------------------------------------------------------- */
typedef struct hvc_ln_args {
  uint32_t size;
  int32_t T, C;
  const float* x; int64_t ldx;
  const float* w; const float* b;
  const float* shift; const float* scale; int64_t mod_ld; int32_t rows_per_batch;
  void* y; int64_t ldy; int32_t y_is_bf16;
  float* mean; float* rstd;
} hvc_ln_args;
int hvc_ln_fwd(const hvc_ln_args* args, void* stream);

/* Backward of the above.  dz (gradient w.r.t. y) comes from exactly one of: dz_bf16 [T,C], dz_f32 [T,C],
 * dz_row [T] (the output head: y is contracted with mult_vec = output_proj.weight, hybrid_vit_backbone.py:266,
 * so dz[t,c] = dz_row[t] * mult_vec[c]).  Outputs: dx f32 [T,C] (= dx_in + LN backward when dx_in != NULL),
 * dw/db [C], dshift/dscale [batch, C] at pitch dmod_ld (modulated LN), dvec [C] + dscalar [1] (head:
 * d output_proj.weight / .bias).  S1/S2: f32 [batch, C] scratch. */
typedef struct hvc_ln_bwd_args {
  uint32_t size;
  int32_t batch, rows_per_batch, C;
  const void* dz_bf16; const float* dz_f32; int64_t lddz; const float* dz_row;
  const float* x; int64_t ldx;
  const float* mean; const float* rstd;
  const float* w; const float* b;
  const float* scale; int64_t mod_ld;
  const float* mult_vec;
  const float* dx_in; int64_t lddx_in;
  float* dx; int64_t lddx;
  float* dw; float* db;
  float* dshift; float* dscale; int64_t dmod_ld;
  float* dvec; float* dscalar;
  float* S1; float* S2;
} hvc_ln_bwd_args;
int hvc_ln_bwd(const hvc_ln_bwd_args* args, void* stream);

/* Backward of the gated residual  out = resid + gate[batch] * branch  (hybrid_vit_backbone.py:123,128,139):
 *   dbranch bf16 [T,C] = gate * dout;  dgate f32 [batch,C] = sum_n dout*branch (optional, needs branch = the dropped branch);
 *   dbias f32 [C] = sum_T dbranch (optional; the bias of the projection that produced branch).
 * gate == NULL means 1 (cross-attention residual).  D1: f32 [batch, C] scratch. */
typedef struct hvc_resid_bwd_args {
  uint32_t size;
  int32_t batch, rows_per_batch, C;
  const float* dout; int64_t lddout;
  const void* branch; int64_t ldbranch;
  const float* gate; int64_t gate_ld;
  void* dbranch; int64_t lddbranch;
  float* dgate; float* dbias; float* D1;
  hvc_dropout drop;               /* the dropout that was applied to branch in the forward epilogue: dbranch = gate * drop(dout) */
} hvc_resid_bwd_args;
int hvc_resid_bwd(const hvc_resid_bwd_args* args, void* stream);

/* out[c] = sum_t x[t,c] for bf16 x [T,C] (bias gradient of mlp.0, hybrid_vit_backbone.py:76). */
int hvc_colsum_bf16(const void* x, int64_t ldx, int32_t T, int32_t C, float* out, void* stream);
/* f32 -> bf16, contiguous (weights are cast once per optimizer step). */
int hvc_cast_bf16(const float* x, void* y, int64_t n, void* stream);
/* x[b,m,c] (f32 or bf16, element strides sb/sm/sc -- e.g. the transposed view of the (B,C,H,W) X-ray
 * feature map, model_direct.py:80) -> y bf16 [B*M, C] contiguous. */
int hvc_cast_tokens(const void* x, int32_t x_is_bf16, int64_t sb, int64_t sm, int64_t sc, void* y, int32_t B,
                    int32_t M, int32_t C, void* stream);
/* AdaLN modulation linear, vit_components.py:144: out[B,J] = cond[B,K] W[J,K]^T + bias[J] (fp32 throughout). */
int hvc_adaln_fwd(const float* cond, int64_t ldc, const float* W, const float* bias, float* out, int32_t B,
                  int32_t K, int32_t J, void* stream);
int hvc_adaln_bwd(const float* dparams, const float* cond, int64_t ldc, const float* W, float* dW, float* dbias,
                  float* dcond, int32_t B, int32_t K, int32_t J, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Voxel embedding (hybrid_vit_backbone.py:195-210,252): Conv3d(k=3, pad=1, stride 1|2) as
 * im2col + hvc_gemm, GroupNorm(groups)+SiLU, on channels-last activations [B, voxels, C].
 * hvc_conv3d_geom describes the conv INPUT: sizes and element strides (any layout: NCDHW for the
 * module input, channels-last for intermediates).  Patch matrix: bf16 [B*Do*Ho*Wo, Kp], column
 * k = cin*27 + kd*9 + kh*3 + kw (the order of weight.view(Cout, Cin*27)), Kp = K rounded up to 8.
 * ---------------------------------------------------------------------------------------------- */
typedef struct hvc_conv3d_geom {
  int32_t B, Cin, D, H, W, stride;
  int64_t sb, sc, sd, sh, sw;
} hvc_conv3d_geom;
int hvc_im2col3d(const void* x, int32_t x_is_bf16, const hvc_conv3d_geom* geom, void* cols, void* stream);
/* dx (f32, layout given by geom strides) = adjoint of im2col applied to dcols (bf16 [M, Kp]). */
int hvc_col2im3d(const void* dcols, const hvc_conv3d_geom* geom, float* dx, void* stream);
/* Channels-last variants (geom->sc == 1, Cin % 8 == 0, strides multiples of 8): column k = (kd*9 + kh*3 + kw)*Cin + cin, i.e. the
 * order of weight.permute(0,2,3,4,1).reshape(Cout, 27*Cin); every access is a 16-byte run of 8 channels.  Used by the full-resolution
 * convs of the cascade's detail_enhancer (progressive_cascade/model_progressive.py:263). */
int hvc_im2col3d_cl(const void* x, int32_t x_is_bf16, const hvc_conv3d_geom* geom, void* cols, void* stream);
int hvc_col2im3d_cl(const void* dcols, const hvc_conv3d_geom* geom, float* dx, void* stream);
/* Zero-padded channels-last volumes, the operand layouts of the implicit-GEMM conv (hvc_conv_taps):
 * pad:   src (B, D, H, W, Cs) f32|bf16 dense -> dst bf16 (B, D+1+pad_hi, H+1+pad_hi, W+1+pad_hi, Cp): one zero voxel on the low side of every
 *        axis, pad_hi (0 | 1) on the high side; channels [Cs, Cp) zero-filled.   unpad: the interior of such an f32 volume -> dense.
 * s2d_pad (stride-2 convs): the volume split by the parity of (d, h, w) into eight half-resolution volumes, each padded by one zero
 *        voxel on the low side, stacked: dst bf16 (8, B, D/2+1, H/2+1, W/2+1, C), parity index (d&1)*4 + (h&1)*2 + (w&1).
 * d2s_unpad: the inverse for gradients, src f32 (8, B, D/2+1, H/2+1, W/2+1, C) -> dst f32 (B, D, H, W, C).
 * Cs, Cp, C % 8 == 0 (C % 4 for the f32 sources); D, H, W even for the parity split. */
int hvc_pad3d_cl(const void* src, int32_t src_is_bf16, void* dst, int32_t B, int32_t D, int32_t H, int32_t W, int32_t Cs, int32_t Cp,
                 int32_t pad_hi, void* stream);
int hvc_unpad3d_cl(const float* src, float* dst, int32_t B, int32_t D, int32_t H, int32_t W, int32_t C, int32_t pad_hi, void* stream);
int hvc_s2d_pad_cl(const void* src, int32_t src_is_bf16, void* dst, int32_t B, int32_t D, int32_t H, int32_t W, int32_t C, void* stream);
int hvc_d2s_unpad_cl(const float* src, float* dst, int32_t B, int32_t D, int32_t H, int32_t W, int32_t C, void* stream);
/* y [B,V,C] (bf16, or f32 when it feeds the token stream) = SiLU(GroupNorm(x f32 [B,V,C])); mean/rstd f32
 * [B,groups] saved; scratch f32 [2*B*C]. */
int hvc_groupnorm_silu_fwd(const float* x, const float* w, const float* b, int32_t B, int32_t V, int32_t C,
                           int32_t groups, void* y, int32_t y_is_bf16, float* mean, float* rstd, float* scratch,
                           void* stream);
/* dx f32 [B,V,C], dw/db f32 [C]; scratch f32 [2*B*C + 2*B*groups]. */
int hvc_groupnorm_silu_bwd(const float* dy, const float* x, const float* w, const float* b, const float* mean,
                           const float* rstd, int32_t B, int32_t V, int32_t C, int32_t groups, float* dx,
                           float* dw, float* db, float* scratch, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Ends of HybridViT3D.forward (hybrid_vit_backbone.py:258,265-272).
 * ---------------------------------------------------------------------------------------------- */
/* out[b,:] = x[b mod x_batch,:] + pos[:]  (n = N*C f32 elements per sample). */
int hvc_add_pos(const float* x, int32_t x_batch, const float* pos, float* out, int32_t B, int64_t n, void* stream);
/* out[:] = sum_b x[b,:]  (gradient of pos_embed / of a batch-expanded input). */
int hvc_batch_sum(const float* x, float* out, int32_t B, int64_t n, void* stream);
/* v[t] = output_proj(LayerNorm(x[t]))  (C -> 1), mean/rstd saved; backward is hvc_ln_bwd with dz_row. */
int hvc_head_fwd(const float* x, int64_t ldx, const float* w, const float* b, const float* wo, const float* bo,
                 float* v, float* mean, float* rstd, int32_t T, int32_t C, void* stream);
/* F.interpolate(mode="trilinear", align_corners=True) on [B, Di,Hi,Wi] -> [B, Do,Ho,Wo], and its adjoint. */
int hvc_upsample3d_fwd(const float* v, float* out, int32_t B, int32_t Di, int32_t Hi, int32_t Wi, int32_t Do,
                       int32_t Ho, int32_t Wo, void* stream);
int hvc_upsample3d_bwd(const float* dout, float* dv, int32_t B, int32_t Di, int32_t Hi, int32_t Wi, int32_t Do,
                       int32_t Ho, int32_t Wo, void* stream);
/* The same with either corner convention: align_corners = 0 is nn.Upsample(scale_factor=2, mode="trilinear") / F.interpolate(...,
 * align_corners=False) of the cascade's stage wrappers (progressive_cascade/model_progressive.py:169,211-212). */
int hvc_interp3d_fwd(const float* v, float* out, int32_t B, int32_t Di, int32_t Hi, int32_t Wi, int32_t Do, int32_t Ho,
                     int32_t Wo, int32_t align_corners, void* stream);
int hvc_interp3d_bwd(const float* dout, float* dv, int32_t B, int32_t Di, int32_t Hi, int32_t Wi, int32_t Do, int32_t Ho,
                     int32_t Wo, int32_t align_corners, void* stream);

/* Conv3d(C -> 1, kernel 1) on a channels-last activation y f32 [M, C] (C = 32 | 64): the last layer of
 * Stage3Refiner256.detail_enhancer (progressive_cascade/model_progressive.py:266).
 * fwd: out[m] = bias[0] + sum_c y[m,c] w[c].   bwd: dy[m,c] = dout[m] w[c]; dw[c] += sum_m dout[m] y[m,c]; db[0] += sum_m dout[m]
 * (dw, db are accumulated into: zero-fill them first). */
int hvc_chan_dot_fwd(const float* y, const float* w, const float* bias, float* out, int64_t M, int32_t C, void* stream);
int hvc_chan_dot_bwd(const float* dout, const float* y, const float* w, float* dy, float* dw, float* db, int64_t M, int32_t C,
                     void* stream);


/* ------------------------------------------------------------------------------------------------
 * fp32 verification mode (the 1e-4 fp32 parity bar): fp32-accurate products on the bf16 tensor cores.
 * An fp32 operand is split into three bf16 terms x = x0 + x1 + x2; the six significant partial products are
 * laid out along K so ONE hvc_gemm with K' = 6K gives sum_k a_k b_k to ~fp32 accuracy:
 *   pattern 0 (A side): [x0 | x1 | x2 | x0 | x1 | x0]     pattern 1 (B side): [x0 | x0 | x0 | x1 | x1 | x2]
 *   pattern 2 (A side): [x0 | x1 | x0]                    pattern 3 (B side): [x0 | x0 | x1]      (two terms, ~2^-16)
 * hvc_split3: x f32 [R, C] (pitch ldx) -> out bf16 [R, 6C] (concat_rows = 0; for K-major operands) or
 * [6R, C] (concat_rows = 1; for MN-major operands such as V in P V), pitch ldo.
 * hvc_softmax_rows: in place s[r,:] <- exp2(s[r,:] - max) / sum for f32 s [R, M] holding scores * scale * log2(e)
 * (the materialised softmax of vit_components.py:46-48,103-105); lse2 [R] optional.
 * hvc_im2col3d_f32: hvc_im2col3d with an f32 patch matrix [B*Do*Ho*Wo, Kp].
 * Forward only; used by hybrid_vit_cascade_b200.precision("fp32").
 * ---------------------------------------------------------------------------------------------- */
int hvc_split3(const float* x, int64_t ldx, int32_t R, int32_t C, void* out, int64_t ldo, int32_t pattern,
               int32_t concat_rows, void* stream);
int hvc_softmax_rows(float* s, int64_t lds, int32_t R, int32_t M, float* lse2, void* stream);
int hvc_im2col3d_f32(const float* x, const hvc_conv3d_geom* geom, float* cols, void* stream);
/* out f32 [T,N] = resid + gate[t / rows_per_batch] * act(acc + bias)  (bias, resid, gate optional; act NONE | GELU):
 * the epilogue of a verification-mode GEMM whose split-K partial sums were reduced into acc with fp32 atomics
 * (TMEM accumulation chains are kept <= 256 elements because the tensor-core accumulator truncates). out may alias acc. */
int hvc_epilogue_f32(const float* acc, int64_t lda, int32_t T, int32_t N, const float* bias, int32_t activation,
                     const float* resid, int64_t ldr, const float* gate, int64_t gate_ld, int32_t rows_per_batch,
                     float* out, int64_t ldo, void* stream);

/* ------------------------------------------------------------------------------------------------
 * X-ray encoder in front of the backbone: XrayConditioningModule, models/diagnostic_losses.py:68-138.
 * Conv2d (k x k, stride 1|2) as im2col + hvc_gemm on channels-last activations [images, pixels, C]; BatchNorm2d + ReLU as
 * hvc_norm_act with one group per channel over all rows (train mode: batch statistics; eval mode: stats_given = 1 with
 * mean / rstd derived from the running buffers).  hvc_conv2d_geom describes the conv INPUT: sizes and element strides.
 * Patch matrix: bf16 or f32 [N*Ho*Wo, Kp], column kk = cin*k*k + kh*k + kw (the order of weight.view(Cout, Cin*k*k)), Kp = K rounded up
 * to 8.  The forward convolutions of the encoder use the f32 patch matrix with the three-term operand split (hvc_split3): ReLU and
 * max-pool are discontinuous, and with bf16 products ~0.2 % of the masks flip, which is visible in the gradients of the early layers.
 * ---------------------------------------------------------------------------------------------- */
typedef struct hvc_conv2d_geom {
  int32_t N, Cin, H, W, k, stride, pad;
  int64_t sn, sc, sh, sw;
} hvc_conv2d_geom;
int hvc_im2col2d(const void* x, int32_t x_is_bf16, const hvc_conv2d_geom* geom, void* cols, int32_t cols_is_f32, void* stream);
/* The gather with the two-term operand split fused in: out bf16 [M, 3*Kp] = [c0 | c1 | c0], c0 = bf16(v), c1 = bf16(v - c0); against
 * weights split with hvc_split3 pattern 3 ([w0 | w0 | w1]) one hvc_gemm with K' = 3*Kp gives the product to ~2^-16.  Block 0 is the
 * plain bf16 patch matrix (row pitch 3*Kp) that the backward GEMMs read. */
int hvc_im2col2d_split(const float* x, const hvc_conv2d_geom* geom, void* out, void* stream);
/* dx (f32, layout given by geom strides) = adjoint of im2col2d applied to dcols (bf16 [M, Kp]). */
int hvc_col2im2d(const void* dcols, const hvc_conv2d_geom* geom, float* dx, void* stream);
/* y = act(norm(x) * w + b) on f32 x [B, V, C] channels-last with `groups` groups per sample: activation 0 = SiLU (GroupNorm+SiLU
 * of the voxel embed; hvc_groupnorm_silu_* are the activation-0 forms), 1 = ReLU (BatchNorm2d+ReLU: B = 1, V = all rows,
 * groups = C, diagnostic_losses.py:81-93), 2 = exact-erf GELU (GroupNorm+GELU of MultiScaleXrayEncoder.to_stage1/2,
 * progressive_cascade/model_progressive.py:38-52).  stats_given: mean/rstd [B, groups] are inputs (eval-mode BatchNorm) instead of outputs.
 * Backward: stats_frozen = 1 treats mean/rstd as constants.  Scratch sizes as for hvc_groupnorm_silu_*. */
int hvc_norm_act_fwd(const float* x, const float* w, const float* b, int32_t B, int32_t V, int32_t C, int32_t groups,
                     int32_t activation, int32_t stats_given, void* y, int32_t y_is_bf16, float* mean, float* rstd,
                     float* scratch, void* stream);
int hvc_norm_act_bwd(const float* dy, const float* x, const float* w, const float* b, const float* mean, const float* rstd,
                     int32_t B, int32_t V, int32_t C, int32_t groups, int32_t activation, int32_t stats_frozen, float* dx,
                     float* dw, float* db, float* scratch, void* stream);
/* nn.MaxPool2d(k, stride, pad) on channels-last f32 [N, H, W, C] -> f32 (diagnostic_losses.py:84,89); arg = window-relative index of
 * the maximum (u8 [N, Ho, Wo, C]) for the backward, which gathers dy (f32) into dx (f32 [N, H, W, C], every element written). */
int hvc_maxpool2d_fwd(const void* x, void* y, uint8_t* arg, int32_t N, int32_t H, int32_t W, int32_t C, int32_t k, int32_t stride,
                      int32_t pad, void* stream);
int hvc_maxpool2d_bwd(const float* dy, const uint8_t* arg, float* dx, int32_t N, int32_t H, int32_t W, int32_t C, int32_t k,
                      int32_t stride, int32_t pad, void* stream);
/* feat[b, p, c] = mean_v x[b*V + v, p, c] (f32; the view average, diagnostic_losses.py:125) and, when pooled != NULL,
 * pooled[b, c] = mean_p feat[b, p, c] (:130).  Backward: dx = (dfeat + dpooled / P) / V broadcast over the views. */
int hvc_view_mean_fwd(const float* x, float* feat, float* pooled, int32_t B, int32_t V, int32_t P, int32_t C, void* stream);
int hvc_view_mean_bwd(const float* dfeat, const float* dpooled, float* dx, int32_t B, int32_t V, int32_t P, int32_t C, void* stream);
/* out = silu(x) (dy == NULL) or dy * silu'(x): the nn.SiLU of the time MLP (diagnostic_losses.py:100). */
int hvc_silu(const float* x, const float* dy, float* out, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Direct-regression training loss: DirectRegressionLoss = l1_weight * L1 + ssim_weight * (1 - mean SSIM3D),
 * direct_regression/model_direct.py:88-131 (compute_ssim_loss: five F.avg_pool3d(window, stride 1, zero padding) box filters).
 * pred / target: f32 [B, D, H, W] contiguous.  Forward: filtered f32 [5n] (box-filtered pred, target, pred^2, target^2, pred*target;
 * kept for the backward), scratch f32 [10n], sums f64 [2] = {sum SSIM, sum |pred - target|} (n = B*D*H*W).
 * Backward: dpred = upstream * (c_ssim * d(sum SSIM)/dpred + c_l1 * sign(pred - target)); upstream = one f32 in device memory (the
 * gradient of the scalar loss; NULL = 1); scratch f32 [9n].
 * ---------------------------------------------------------------------------------------------- */
int hvc_ssim_l1_fwd(const float* pred, const float* target, int32_t B, int32_t D, int32_t H, int32_t W, int32_t window, float* filtered,
                    float* scratch, double* sums, void* stream);
int hvc_ssim_l1_bwd(const float* pred, const float* target, const float* filtered, int32_t B, int32_t D, int32_t H, int32_t W,
                    int32_t window, float c_ssim, float c_l1, const float* upstream, float* scratch, float* dpred, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Stage 2-3 loss terms of the progressive cascade (direct_regression/progressive_cascade/loss_multiscale.py).
 * Volumes are f32 [B, D, H, W] contiguous; sums are f64 device accumulators (zeroed by the call); `upstream` = one f32 in device memory
 * (gradient of the scalar loss; NULL = 1), so no value crosses to the host.
 *   TotalVariationLoss (:140-188): hvc_tv_fwd -> sums[3] = sum sqrt(diff^2 + eps) along D, H, W; hvc_tv_finalize -> loss[0] = clamp(tv, 0, 100)
 *     (sums_target == NULL) or |tv_pred - tv_target|, coef[0] = d loss / d tv_pred; hvc_tv_bwd -> dx.
 *   FrequencyLoss (:191-236): spectra = interleaved complex64 [B, D, H, W] from the library FFT (torch.fft.fftn); hvc_freq_l1_fwd ->
 *     sums[2] = sum | |Fp| - |Ft| | over the low / high-frequency sets (mask built on the unshifted spectrum exactly as :214-229);
 *     hvc_freq_l1_bwd -> dspec (complex64) = upstream * (c_low | c_high) * sign(|Fp| - |Ft|) * Fp / |Fp|.
 *   DRRReprojectionLoss (:239-293): hvc_proj_mean_fwd -> ap [B, H, W] = mean over D, lat [B, D, H] = mean over W; hvc_proj_mean_bwd is
 *     the adjoint; the bilinear resize is hvc_interp3d_* with unit depth; hvc_l1_fwd / hvc_l1_bwd = F.l1_loss sums and gradient.
 * ---------------------------------------------------------------------------------------------- */
int hvc_tv_fwd(const float* x, int32_t B, int32_t D, int32_t H, int32_t W, float eps, double* sums, void* stream);
int hvc_tv_finalize(const double* sums_pred, const double* sums_target, int32_t B, int32_t D, int32_t H, int32_t W, float* loss, float* coef,
                    void* stream);
int hvc_tv_bwd(const float* x, int32_t B, int32_t D, int32_t H, int32_t W, float eps, const float* coef, const float* upstream, float* dx,
               void* stream);
int hvc_freq_l1_fwd(const float* spec_pred, const float* spec_target, int32_t B, int32_t D, int32_t H, int32_t W, double* sums, void* stream);
int hvc_freq_l1_bwd(const float* spec_pred, const float* spec_target, int32_t B, int32_t D, int32_t H, int32_t W, float c_low, float c_high,
                    const float* upstream, float* dspec, void* stream);
int hvc_proj_mean_fwd(const float* vol, int32_t B, int32_t D, int32_t H, int32_t W, float* ap, float* lat, void* stream);
int hvc_proj_mean_bwd(const float* dap, const float* dlat, int32_t B, int32_t D, int32_t H, int32_t W, float* dvol, void* stream);
int hvc_l1_fwd(const float* a, const float* b, int64_t n, double* sum, void* stream);
int hvc_l1_bwd(const float* a, const float* b, int64_t n, float c, const float* upstream, float* da, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Optimizer step on flat fp32 buffers: torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW.step() of the
 * reference trainers (train_direct_4gpu.py:72-80) over the data-parallel gradient buckets.
 *   hvc_sumsq_f32:  accum[0] += sum x^2 (double; zero it first, call once per bucket)
 *   hvc_adamw_tick: state[0] += 1 (the step count t, kept on the device so the step can live in a CUDA graph)
 *   hvc_adamw_flat: g' = g * min(1, max_norm / (sqrt(sumsq[0]) + 1e-6)) (max_norm <= 0: no clipping);
 *                   p *= 1 - lr*wd;  m = b1 m + (1-b1) g';  v = b2 v + (1-b2) g'^2;  p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
 * ---------------------------------------------------------------------------------------------- */
int hvc_sumsq_f32(const float* x, int64_t n, double* accum, void* stream);
int hvc_adamw_tick(float* state, void* stream);
int hvc_adamw_flat(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                   float weight_decay, float max_norm, const double* sumsq, const float* state, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HVC_H_ */
