"""ctypes binding of libhvc_sm100a.so (the C ABI declared in include/hvc.h).

There is no fallback: if the shared library is missing, fails to load, or the device is not
sm_100, every compute entry point raises.  ``python -m hybrid_vit_cascade_b200.build`` (or
``__graft_entry__.build()``) produces the library in-tree.
"""
import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HVC_LIB") or os.path.join(_PKG, "libhvc_sm100a.so")     # HVC_LIB: A/B-test another build of the same ABI

c_f32p = C.c_void_p
_lib = None
_device_ok = {}


class HvcError(RuntimeError):
    pass


class Dropout(C.Structure):
    _fields_ = [("seed", C.c_void_p), ("site", C.c_uint32), ("p", C.c_float)]


class ConvTaps(C.Structure):
    _fields_ = [("side", C.c_int32), ("cin", C.c_int32), ("n_taps", C.c_int32), ("rows", C.c_int32), ("offsets", C.c_int32 * 27)]


class GemmArgs(C.Structure):
    _fields_ = [
        ("size", C.c_uint32), ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("A", C.c_void_p), ("lda", C.c_int64), ("a_major", C.c_int32),
        ("B", C.c_void_p), ("ldb", C.c_int64), ("b_major", C.c_int32),
        ("epilogue", C.c_int32), ("activation", C.c_int32),
        ("out", C.c_void_p), ("ldo", C.c_int64),
        ("out2", C.c_void_p), ("ldo2", C.c_int64),
        ("bias", C.c_void_p),
        ("resid", C.c_void_p), ("ldr", C.c_int64),
        ("gate", C.c_void_p), ("gate_ld", C.c_int64),
        ("rows_per_batch", C.c_int32),
        ("aux", C.c_void_p), ("ldaux", C.c_int64),
        ("alpha", C.c_float), ("k_splits", C.c_int32),
        ("drop", Dropout),
        ("taps", ConvTaps),
    ]


class AttnArgs(C.Structure):
    _fields_ = [
        ("size", C.c_uint32),
        ("batch", C.c_int32), ("heads", C.c_int32), ("nq", C.c_int32), ("nk", C.c_int32), ("head_dim", C.c_int32),
        ("q", C.c_void_p), ("ldq", C.c_int64),
        ("k", C.c_void_p), ("ldk", C.c_int64),
        ("v", C.c_void_p), ("ldv", C.c_int64),
        ("o", C.c_void_p), ("ldo", C.c_int64),
        ("lse", C.c_void_p),
        ("d_o", C.c_void_p), ("lddo", C.c_int64),
        ("dq", C.c_void_p), ("lddq", C.c_int64),
        ("dk", C.c_void_p), ("lddk", C.c_int64),
        ("dv", C.c_void_p), ("lddv", C.c_int64),
        ("delta", C.c_void_p),
        ("dq_accum", C.c_void_p),
        ("probs", C.c_void_p),
        ("scale", C.c_float),
        ("drop", Dropout),
    ]


class LnArgs(C.Structure):
    _fields_ = [
        ("size", C.c_uint32), ("T", C.c_int32), ("C", C.c_int32),
        ("x", C.c_void_p), ("ldx", C.c_int64),
        ("w", C.c_void_p), ("b", C.c_void_p),
        ("shift", C.c_void_p), ("scale", C.c_void_p), ("mod_ld", C.c_int64), ("rows_per_batch", C.c_int32),
        ("y", C.c_void_p), ("ldy", C.c_int64), ("y_is_bf16", C.c_int32),
        ("mean", C.c_void_p), ("rstd", C.c_void_p),
    ]


class LnBwdArgs(C.Structure):
    _fields_ = [
        ("size", C.c_uint32), ("batch", C.c_int32), ("rows_per_batch", C.c_int32), ("C", C.c_int32),
        ("dz_bf16", C.c_void_p), ("dz_f32", C.c_void_p), ("lddz", C.c_int64), ("dz_row", C.c_void_p),
        ("x", C.c_void_p), ("ldx", C.c_int64),
        ("mean", C.c_void_p), ("rstd", C.c_void_p),
        ("w", C.c_void_p), ("b", C.c_void_p),
        ("scale", C.c_void_p), ("mod_ld", C.c_int64),
        ("mult_vec", C.c_void_p),
        ("dx_in", C.c_void_p), ("lddx_in", C.c_int64),
        ("dx", C.c_void_p), ("lddx", C.c_int64),
        ("dw", C.c_void_p), ("db", C.c_void_p),
        ("dshift", C.c_void_p), ("dscale", C.c_void_p), ("dmod_ld", C.c_int64),
        ("dvec", C.c_void_p), ("dscalar", C.c_void_p),
        ("S1", C.c_void_p), ("S2", C.c_void_p),
    ]


class ResidBwdArgs(C.Structure):
    _fields_ = [
        ("size", C.c_uint32), ("batch", C.c_int32), ("rows_per_batch", C.c_int32), ("C", C.c_int32),
        ("dout", C.c_void_p), ("lddout", C.c_int64),
        ("branch", C.c_void_p), ("ldbranch", C.c_int64),
        ("gate", C.c_void_p), ("gate_ld", C.c_int64),
        ("dbranch", C.c_void_p), ("lddbranch", C.c_int64),
        ("dgate", C.c_void_p), ("dbias", C.c_void_p), ("D1", C.c_void_p),
        ("drop", Dropout),
    ]


class Conv3dGeom(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("Cin", C.c_int32), ("D", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("stride", C.c_int32),
        ("sb", C.c_int64), ("sc", C.c_int64), ("sd", C.c_int64), ("sh", C.c_int64), ("sw", C.c_int64),
    ]


class Conv2dGeom(C.Structure):
    _fields_ = [
        ("N", C.c_int32), ("Cin", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("k", C.c_int32), ("stride", C.c_int32),
        ("pad", C.c_int32),
        ("sn", C.c_int64), ("sc", C.c_int64), ("sh", C.c_int64), ("sw", C.c_int64),
    ]


def lib():
    """Load the shared library once; raise loudly if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HvcError(
            f"{LIB_PATH} not found: the hybrid_vit_cascade_b200 CUDA extension is not built "
            "(run `python -m hybrid_vit_cascade_b200.build`). There is no CPU or PyTorch fallback.")
    L = C.CDLL(LIB_PATH)
    L.hvc_version.restype = C.c_int
    L.hvc_last_error.restype = C.c_char_p
    L.hvc_launch_count.restype = C.c_uint64
    L.hvc_check_device.restype = C.c_int
    for name in EXPORTS:
        getattr(L, name)  # AttributeError here = header/library mismatch
    _lib = L
    return L


# every symbol include/hvc.h declares (tests/test_abi.py checks header <-> library <-> this list)
EXPORTS = [
    "hvc_version", "hvc_last_error", "hvc_launch_count", "hvc_check_device",
    "hvc_gemm", "hvc_attn_fwd", "hvc_attn_bwd",
    "hvc_ln_fwd", "hvc_ln_bwd", "hvc_resid_bwd", "hvc_colsum_bf16", "hvc_cast_bf16", "hvc_cast_tokens",
    "hvc_adaln_fwd", "hvc_adaln_bwd",
    "hvc_im2col3d", "hvc_col2im3d", "hvc_im2col3d_cl", "hvc_col2im3d_cl", "hvc_pad3d_cl", "hvc_unpad3d_cl", "hvc_s2d_pad_cl", "hvc_d2s_unpad_cl", "hvc_groupnorm_silu_fwd", "hvc_groupnorm_silu_bwd",
    "hvc_add_pos", "hvc_batch_sum", "hvc_head_fwd", "hvc_upsample3d_fwd", "hvc_upsample3d_bwd", "hvc_interp3d_fwd", "hvc_interp3d_bwd",
    "hvc_split3", "hvc_softmax_rows", "hvc_im2col3d_f32", "hvc_epilogue_f32",
    "hvc_im2col2d", "hvc_im2col2d_split", "hvc_col2im2d", "hvc_norm_act_fwd", "hvc_norm_act_bwd", "hvc_maxpool2d_fwd", "hvc_maxpool2d_bwd",
    "hvc_view_mean_fwd", "hvc_view_mean_bwd", "hvc_silu", "hvc_ssim_l1_fwd", "hvc_ssim_l1_bwd", "hvc_chan_dot_fwd", "hvc_chan_dot_bwd",
    "hvc_sumsq_f32", "hvc_adamw_tick", "hvc_adamw_flat",
    "hvc_tv_fwd", "hvc_tv_finalize", "hvc_tv_bwd", "hvc_freq_l1_fwd", "hvc_freq_l1_bwd", "hvc_proj_mean_fwd", "hvc_proj_mean_bwd",
    "hvc_l1_fwd", "hvc_l1_bwd",
]


def check(rc, what=""):
    if rc != 0:
        msg = lib().hvc_last_error().decode(errors="replace")
        raise HvcError(f"{what} failed ({rc}): {msg}")


def require_device(dev_index):
    """Verify once per device that we are on sm_100 with a usable driver.  hvc_check_device() inspects the CURRENT device, so
    `dev_index` must be the current device's ordinal (kernels._need_cuda passes torch.cuda.current_device())."""
    if not _device_ok.get(dev_index):
        check(lib().hvc_check_device(), "hvc_check_device")
        _device_ok[dev_index] = True


def launch_count():
    return int(lib().hvc_launch_count())
