"""ctypes binding of libdcb200.so (include/dcb200.h).  No fallbacks: a missing library is an error."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdcb200.so")

MAX_TAPS = 9
c_int = ctypes.c_int32
c_i64 = ctypes.c_int64
c_vp = ctypes.c_void_p
c_f = ctypes.c_float


class DcbError(RuntimeError):
    pass


class TapGemmArgs(ctypes.Structure):
    _fields_ = [
        ("a", c_vp), ("a_s1", c_i64), ("a_s2", c_i64), ("a_s3", c_i64),
        ("C", c_int), ("D1", c_int), ("D2", c_int), ("D3", c_int),
        ("A1", c_int), ("A2", c_int), ("A3", c_int),
        ("b1", c_int), ("b2", c_int), ("b3", c_int),
        ("ntaps", c_int), ("taps", (c_int * 3) * MAX_TAPS),
        ("w", c_vp), ("w_ld", c_i64), ("N", c_int), ("block_n", c_int),
        ("bias", c_vp), ("group_bias", c_vp), ("rows_per_group", c_int), ("ld_group_bias", c_int),
        ("residual", c_vp), ("ld_residual", c_int), ("geglu", c_int),
        ("out_f32", c_vp), ("out_bf16", c_vp), ("ld_out", c_int), ("out_scale", c_f),
        ("gn_acc", c_vp), ("gn_cg", c_int), ("gn_unit_mode", c_int),
        ("ln_acc", c_vp), ("ln_x_bf16", c_vp), ("ld_ln_x", c_int), ("ln_channels", c_int),
        ("ln_colsum", c_vp), ("ln_eps", c_f), ("reserved0", c_int),
        ("out_s1", c_i64), ("out_s2", c_i64), ("out_s3", c_i64),
    ]


class AttentionArgs(ctypes.Structure):
    _fields_ = [
        ("q", c_vp), ("ldq", c_i64),
        ("k", c_vp * 2), ("v", c_vp * 2), ("ldk", c_i64 * 2), ("ldv", c_i64 * 2),
        ("nk", c_int * 2), ("kv_batch_div", c_int * 2), ("seg_scale", c_f * 2),
        ("nseg", c_int), ("batch", c_int), ("heads", c_int), ("nq", c_int), ("head_dim", c_int),
        ("o", c_vp), ("ldo", c_i64), ("softmax_scale", c_f),
    ]


class DdimCoeffs(ctypes.Structure):
    _fields_ = [
        ("is_v", c_int), ("cfg_scale", c_f), ("guidance_rescale", c_f), ("one_minus_guidance_rescale", c_f),
        ("sqrt_acp_t", c_f), ("sqrt_1m_acp_t", c_f), ("sqrt_a_t", c_f), ("sqrt_1m_a_t", c_f),
        ("sqrt_a_prev", c_f), ("dir_coef", c_f), ("sigma_t", c_f), ("temperature", c_f), ("x0_rescale", c_f),
    ]


_lib = None


def lib():
    """Load (once) and return the CDLL.  Raises if the extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DcbError(
                f"{LIB_PATH} is missing: run `python -m dynamicrafter_b200.build` "
                "(there is no CPU or PyTorch fallback for the dcb200 kernels)")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.dcb_last_error.restype = ctypes.c_char_p
        for name in EXPORTS:
            fn = getattr(_lib, name)       # AttributeError if the symbol is absent
            if name == "dcb_group_norm_workspace_bytes":
                fn.restype = ctypes.c_size_t
            elif name != "dcb_last_error":
                fn.restype = ctypes.c_int
    return _lib


# every symbol include/dcb200.h declares (tests check the .so exports all of them)
EXPORTS = [
    "dcb_last_error", "dcb_abi_version", "dcb_check_device", "dcb_tap_gemm",
    "dcb_group_norm_workspace_bytes", "dcb_group_norm_stats", "dcb_group_norm_apply", "dcb_layer_norm",
    "dcb_group_norm_stats_bf16", "dcb_group_norm_apply_bf16", "dcb_group_norm_stats_from_acc",
    "dcb_group_norm_fused", "dcb_group_norm_fused_bf16",
    "dcb_group_norm_sums", "dcb_group_norm_sums_bf16", "dcb_group_norm_stats_from_sums", "dcb_group_norm_apply_f32out",
    "dcb_flash_attention", "dcb_temporal_attention", "dcb_softmax_rows",
    "dcb_nchw_to_rows", "dcb_rows_to_nchw", "dcb_upsample2x", "dcb_space_to_depth", "dcb_cast_bf16", "dcb_gelu_bf16",
    "dcb_timestep_embedding", "dcb_small_linear", "dcb_ddim_update", "dcb_ddim_update_multicond", "dcb_video_to_uint8",
    "dcb_split_bf16x3", "dcb_geglu_f32", "dcb_attention_f32", "dcb_group_norm_fused_f32out", "dcb_layer_norm_f32out",
    "dcb_nchw_to_rows_f32out", "dcb_upsample2x_f32out", "dcb_space_to_depth_f32out",
    "dcb_q_sample", "dcb_vae_attention", "dcb_group_norm_stream_apply", "dcb_block_transpose", "dcb_group_norm_stream_apply2", "dcb_sds_grad",
]


CALLS = [0]     # number of successful C-ABI launches (bench.py reports it as gpu_launches)


def check(rc, what=""):
    CALLS[0] += 1
    if rc != 0:
        msg = lib().dcb_last_error().decode(errors="replace")
        raise DcbError(f"{what} failed ({rc}): {msg}")


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())
