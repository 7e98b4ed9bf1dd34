"""ctypes binding of libadaface_b200.so -- the C-ABI declared in include/adaface_b200.h.

There is NO fallback: if the CUDA library is missing or a call fails, a RuntimeError is raised."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ADAFACE_B200_LIB") or os.path.join(_HERE, "csrc", "libadaface_b200.so")   # env override: A/B builds

_c = ctypes
_p, _i64, _i32, _f32 = _c.c_void_p, _c.c_int64, _c.c_int, _c.c_float

# name -> argtypes (mirrors include/adaface_b200.h one to one)
SIGNATURES = {
    "adaface_version": [],
    "adaface_last_error": [],
    "adaface_launch_count": [],
    "adaface_set_pdl": [_i32],
    "adaface_proj_lora_fwd": [_p, _i64, _p, _p, _i64, _p, _p, _p, _p, _i64, _i32, _p, _i64, _i32, _i64, _i64, _i64,
                              _i64, _i32, _p],
    "adaface_attn_fwd": [_p, _i64, _i64, _p, _i64, _i64, _p, _i64, _i64, _p, _i64, _i64, _i64, _i64, _i64, _i64, _i64,
                         _p, _i32, _f32, _p, _p],
    "adaface_attn_headmajor_fwd": [_p, _i64, _i64, _i64, _p, _i64, _i64, _i64, _p, _i64, _i64, _i64, _p, _i64, _i64, _i64,
                                   _i64, _i64, _i64, _i64, _i64, _i64, _f32, _p, _p],
    "adaface_proj_lora_heads_fwd": [_p, _i64, _p, _p, _i64, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64,
                                    _p],
    "adaface_attn_cross_capture_fwd": [_p, _i64, _i64, _p, _i64, _i64, _p, _i64, _i64, _p, _i64, _i64, _i64, _i64,
                                       _i64, _i64, _i64, _f32, _p, _p, _p, _p, _i64, _p, _p, _p, _i32, _i32, _p],
    "adaface_qmean": [_p, _i32, _i64, _i64, _i64, _i64, _i64, _p, _p],
    "adaface_capture_chan_major": [_p, _i32, _i64, _i64, _i64, _i64, _i64, _f32, _p, _p],
    "adaface_layernorm_fwd": [_p, _i32, _i64, _p, _p, _p, _i32, _i64, _i64, _i64, _f32, _p],
    "adaface_sbg_head_fwd": [_p, _p, _p, _p, _c.POINTER(_f32), _i32, _i64, _p, _p, _p, _i64, _i64, _i64, _f32, _p],
    "adaface_groupnorm_tokens_fwd": [_p, _i32, _p, _p, _i64, _i64, _i64, _i64, _f32, _p, _p, _p, _p],
    "adaface_tokens_to_nchw_add": [_p, _p, _i32, _p, _i64, _i64, _i64, _p],
    "adaface_conv3x3_fwd": [_p, _i64, _i64, _i64, _i64, _p, _p, _i64, _p, _i64, _p, _p, _p, _p, _i64, _i32, _p, _i64, _i32, _i64,
                            _i32, _i32, _p],
    "adaface_groupnorm_act_tokens_fwd": [_p, _p, _p, _i64, _i64, _i64, _i64, _f32, _i32, _p, _p, _p, _p, _p],
    "adaface_groupnorm_act_tokens_ws_floats": [_i64, _i64, _i64, _i64],
    "adaface_silu_fwd": [_p, _i32, _p, _i64, _p],
    "adaface_upsample2x_tokens": [_p, _p, _i64, _i64, _i64, _i64, _p],
    "adaface_timestep_embedding": [_p, _i64, _i64, _f32, _p, _p],
    "adaface_groupnorm_act_tokens_bwd": [_p, _p, _p, _p, _i64, _i64, _i64, _i64, _f32, _i32, _p, _p, _p],
    "adaface_resample2x_bwd": [_p, _p, _i64, _i64, _i64, _i64, _i32, _p],
    # ---- backward (ABI v2)
    "adaface_attn_bwd": [_p, _i64, _i64, _p, _i64, _i64, _p, _i64, _i64, _p, _i64, _i64, _p, _i64, _i64, _p, _p,
                         _p, _i64, _i64, _p, _i64, _i64, _p, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _p, _i32, _f32, _p],
    "adaface_attn_cross_capture_bwd_chunks": [_i64, _i64, _i64],
    "adaface_attn_cross_capture_bwd": [_p, _i64, _i64, _p, _i64, _i64, _p, _i64, _i64, _p, _i64, _i64, _p, _p, _i64, _i64,
                                       _i64, _i64, _i64, _f32, _p, _p, _p, _i32, _i32, _p, _i64, _i64, _p, _i64, _i64, _p, _i64,
                                       _i64, _i32, _p, _f32, _p, _p, _p, _p],
    "adaface_transpose": [_p, _i32, _i64, _i64, _p, _i32, _i64, _i64, _i64, _i64, _i64, _f32, _p, _p, _p],
    "adaface_colsum": [_p, _i32, _i64, _p, _i32, _i64, _p, _p, _p, _i64, _i64, _p],
    "adaface_layernorm_bwd": [_p, _i32, _i64, _p, _i32, _i64, _p, _p, _i64, _p, _p, _i64, _i64, _f32, _p],
    "adaface_act_fwd": [_p, _i64, _p, _i64, _i64, _i64, _i32, _p],
    "adaface_act_bwd": [_p, _i64, _p, _i64, _p, _i64, _i64, _i64, _i32, _p],
    # ---- capture with fused consumers (ABI v4)
    "adaface_attn_cross_consume_fwd": [_p, _i64, _i64, _p, _i64, _i64, _p, _i64, _i64, _p, _i64, _i64, _i64, _i64, _i64, _i64, _i64,
                                       _f32, _p, _p, _p, _p, _i32, _p, _p, _p, _p, _i64, _p],
    "adaface_attn_cross_consume_bwd": [_p, _i64, _i64, _p, _i64, _i64, _p, _i64, _i64, _p, _i64, _i64, _p, _i64, _i64, _i64, _i64,
                                       _i64, _f32, _p, _p, _p, _i32, _p, _i64, _i64, _p, _i64, _i64, _p, _i64, _i64, _i32, _p, _f32,
                                       _p, _p, _p, _p, _p, _p, _p, _p],
    "adaface_sbg_head_fwd_dev": [_p, _p, _p, _p, _p, _i32, _i64, _p, _p, _p, _i64, _i64, _i64, _f32, _p],
    "adaface_sbg_head_bwd_dev": [_p, _p, _p, _p, _p, _i32, _i64, _p, _p, _i64, _p, _p, _p, _p, _p, _p, _p, _i64, _i64, _f32, _p],
    "adaface_dora_colscale": [_p, _i32, _p, _i64, _f32, _p, _p, _i64, _i64, _p],
    "adaface_im2col3x3_tokens": [_p, _p, _i64, _i64, _i64, _i64, _p],
    # ---- sampler step (ABI v4)
    "adaface_ddim_cfg_step": [_p, _i64, _i64, _i32, _p, _p, _p, _p, _p, _p, _p],
    "adaface_sbg_head_bwd": [_p, _p, _p, _p, _c.POINTER(_f32), _i32, _i64, _p, _p, _i64, _p, _p, _p, _p, _p, _p, _p, _i64,
                             _i64, _f32, _p],
}
# entry points that return a value instead of a status
VALUE_RETURNING = ("adaface_version", "adaface_last_error", "adaface_launch_count", "adaface_set_pdl",
                   "adaface_attn_cross_capture_bwd_chunks", "adaface_groupnorm_act_tokens_ws_floats")

_lib = None


def load():
    """Load the shared library (once).  Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). "
            "adaface_b200 has no CPU / PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.argtypes = argtypes
        fn.restype = _c.c_int
    lib.adaface_last_error.restype = _c.c_char_p
    lib.adaface_launch_count.restype = _c.c_int64
    lib.adaface_groupnorm_act_tokens_ws_floats.restype = _c.c_int64
    _lib = lib
    return lib


def call(name, *args):
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed (status {rc}): {lib.adaface_last_error().decode()}")


def cross_capture_bwd_chunks(B, H, Lq):
    return int(load().adaface_attn_cross_capture_bwd_chunks(B, H, Lq))


def set_pdl(mask):
    """Programmatic dependent launch (adaface_set_pdl): bit 0 = projection GEMMs, bit 1 = attention kernels; True = both.
    Returns the previous mask."""
    return int(load().adaface_set_pdl(3 if mask is True else int(mask)))


def launch_count():
    return int(load().adaface_launch_count())
