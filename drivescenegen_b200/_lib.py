"""ctypes binding of libdsg_b200.so (the C ABI declared in include/dsg_b200.h).

There is NO fallback: if the shared library is missing or a call fails, a ``DsgError`` is raised.  The loader only
looks at the in-tree build product (``drivescenegen_b200/libdsg_b200.so``); build it with
``python -m drivescenegen_b200.build`` or ``__graft_entry__.build()``.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdsg_b200.so")


class DsgError(RuntimeError):
    pass


class ConvArgs(C.Structure):
    """mirror of ``dsg_conv_args`` (include/dsg_b200.h)."""
    _fields_ = [
        ("mode", C.c_int32),
        ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("cin", C.c_int32), ("cout", C.c_int32),
        ("x", C.c_void_p),
        ("sc1", C.c_void_p), ("csc1", C.c_int32),
        ("sc2", C.c_void_p), ("csc2", C.c_int32),
        ("wpacked", C.c_void_p),
        ("bias", C.c_void_p),
        ("temb", C.c_void_p),
        ("temb_stride", C.c_int32), ("temb_off", C.c_int32),
        ("residual", C.c_void_p),
        ("out", C.c_void_p),
        ("block_n", C.c_int32),
        ("impl", C.c_int32),
        ("out_nchw_f32", C.c_void_p),
        ("cout_real", C.c_int32),
        ("out_stats", C.c_void_p),
        ("x2", C.c_void_p),
        ("cin1", C.c_int32),
        ("gn_coef", C.c_void_p),
    ]


class WgradArgs(C.Structure):
    """mirror of ``dsg_wgrad_args`` (include/dsg_b200.h)."""
    _fields_ = [
        ("mode", C.c_int32),
        ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("cin", C.c_int32), ("cout", C.c_int32),
        ("x", C.c_void_p),
        ("dy", C.c_void_p),
        ("grad", C.c_void_p),
        ("ci_total", C.c_int32), ("ci_off", C.c_int32),
        ("accumulate", C.c_int32),
        ("inv_scale", C.c_void_p),
        ("workspace", C.c_void_p),
        ("workspace_bytes", C.c_int64),
        ("impl", C.c_int32),
    ]


class PackJob(C.Structure):
    """mirror of ``dsg_pack_job`` (include/dsg_b200.h)."""
    _fields_ = [
        ("mode", C.c_int32), ("cout", C.c_int32), ("cin", C.c_int32), ("csc", C.c_int32),
        ("w", C.c_void_p),
        ("w_sc", C.c_void_p),
        ("out", C.c_void_p),
        ("k_total", C.c_int64), ("rows", C.c_int64), ("chunk_begin", C.c_int64),
    ]


class ReduceJob(C.Structure):
    """mirror of ``dsg_reduce_job`` (include/dsg_b200.h)."""
    _fields_ = [
        ("src", C.c_void_p),
        ("n", C.c_int32), ("parts", C.c_int32), ("c", C.c_int32), ("comps", C.c_int32),
        ("sample_stride", C.c_int64), ("part_stride", C.c_int64),
        ("per_n", C.c_void_p),
        ("per_n_stride", C.c_int32), ("per_n_off", C.c_int32),
        ("inv_scale", C.c_void_p),
        ("out0", C.c_void_p), ("out0b", C.c_void_p), ("out1", C.c_void_p),
        ("block_begin", C.c_int32), ("pad_", C.c_int32),
    ]


_i32, _i64, _p, _f = C.c_int32, C.c_int64, C.c_void_p, C.c_float

# name -> (restype, argtypes); must list every symbol include/dsg_b200.h declares (tests check this)
SIGNATURES = {
    "dsg_version": (C.c_int, []),
    "dsg_last_error": (C.c_char_p, []),
    "dsg_launch_count": (_i64, []),
    "dsg_count_graph_launches": (None, [_i64]),
    "dsg_device_ok": (C.c_int, []),
    "dsg_ddpm_step": (C.c_int, [_p, _p, _p, _p, _i64, _p, _p, _i32, _p]),
    "dsg_ddim_step": (C.c_int, [_p, _p, _p, _p, _i64, _p, _p, _i32, _p]),
    "dsg_add_noise": (C.c_int, [_p, _p, _p, _p, _p, _i32, _p, _i32, _i64, _p]),
    "dsg_step_advance": (C.c_int, [_p, _p, _p, _i32, _p, _p]),
    "dsg_latent_to_image": (C.c_int, [_p, _p, _p, _i32, _i32, _i32, _i32, _p]),
    "dsg_image_to_sample": (C.c_int, [_p, _p, _i32, _i32, _i32, _i32, _i32, _p]),
    "dsg_resize_to_sample": (C.c_int, [_p, _i32, _p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p]),
    "dsg_gray_mask": (C.c_int, [_p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, C.c_double, _p]),
    "dsg_agent_threshold": (C.c_int, [_p, _i64, _p, _i32, _i64, _i32, _p]),
    "dsg_time_embed": (C.c_int, [_p, _p, _i32, _i32, _p, _p, _p, _p, _i32, _p, _p, _i32, _p, _p, _i32, _p]),
    "dsg_conv_in": (C.c_int, [_p, _p, _p, _p, _i32, _i32, _i32, _i32, _i32, _p]),
    "dsg_conv_in_scaled": (C.c_int, [_p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _i32, _p]),
    "dsg_conv_in_stats": (C.c_int, [_p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _i32, _p]),
    "dsg_conv_out_fused": (C.c_int, [_p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _i32, _p]),
    "dsg_conv_out": (C.c_int, [_p, _p, _p, _p, _i32, _i32, _i32, _i32, _i32, _p]),
    "dsg_gn_stats": (C.c_int, [_p, _i32, _p, _i32, _i64, _p]),
    "dsg_gn_apply": (C.c_int, [_p, _i32, _p, _p, _i32, _p, _p, _p, _f, _i32, _p, _i32, _i64, _i32, _p]),
    "dsg_conv": (C.c_int, [C.POINTER(ConvArgs), _p]),
    "dsg_conv_gn_fusable": (C.c_int, [_i32, _i32, _i32, _i32, _i32, _i32]),
    "dsg_gn_coef": (C.c_int, [_i32, _p, _i32, _p, _p, _p, _f, _p, _i32, _i64, _i32, _p]),
    "dsg_packed_k": (_i64, [_i32, _i32, _i32]),
    "dsg_packed_rows": (_i64, [_i32, _i32]),
    "dsg_pack_conv_weight": (C.c_int, [_i32, _p, _i32, _i32, _p, _i32, _p, _p]),
    "dsg_attention": (C.c_int, [_p, _p, _i32, _i32, _i32, _i32, _p]),
    "dsg_attention_ex": (C.c_int, [_p, _p, _i32, _i32, _i32, _i32, _i32, _p, _p]),
    # ---- training path
    "dsg_pack_job_blocks": (_i64, [_i32, _i32, _i32, _i32]),
    "dsg_pack_conv_weights_batched": (C.c_int, [_p, _i32, _i64, _p]),
    "dsg_packed_k_dgrad": (_i64, [_i32, _i32]),
    "dsg_packed_rows_dgrad": (_i64, [_i32, _i32]),
    "dsg_grad_scale": (C.c_int, [_p, _i64, _p, _i32, _p, _p]),
    "dsg_time_embed_ex": (C.c_int, [_p, _p, _i32, _i32, _p, _p, _p, _p, _i32, _p, _p, _i32, _p, _p, _i32, _p, _p]),
    "dsg_lin_dgrad_small": (C.c_int, [_p, _i32, _i32, _p, _i32, _i32, _p, _p, _i32, _p]),
    "dsg_lin_wgrad_small": (C.c_int, [_p, _i32, _i32, _p, _i32, _i32, _i32, _i32, _p, _p, _p, _p]),
    "dsg_gn_bwd": (C.c_int, [_p, _p, _i32, _p, _p, _i32, _p, _p, _p, _f, _i32, _p, _i32, _p, _p, _i32, _p, _i32, _p,
                             _p, _p, _i32, _i32, _i64, _i32, _p]),
    "dsg_gn_bwd_params": (C.c_int, [_p, _i32, _i32, _i32, _p, _p, _p, _p]),
    "dsg_colsum_h16": (C.c_int, [_p, _i64, _i32, _p, _i32, _p]),
    "dsg_colsum_finalize": (C.c_int, [_p, _i32, _i32, _i32, _p, _i32, _i32, _p, _p, _p, _p]),
    "dsg_reduce_rows_batched": (C.c_int, [_p, _i32, _i32, _p]),
    "dsg_conv_wgrad": (C.c_int, [C.POINTER(WgradArgs), _p]),
    "dsg_wgrad_workspace_bytes": (_i64, [_i32, _i32, _i32, _i32, _i32, _i32]),
    "dsg_conv_out_dgrad_weight": (C.c_int, [_p, _i32, _i32, _p, _p, _p]),
    "dsg_small_wgrad": (C.c_int, [_p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _p, _i32, _p, _p, _p, _p]),
    "dsg_attention_train_tc_ok": (C.c_int, [_i32, _i32]),
    "dsg_attention_train": (C.c_int, [_p, _p, _p, _i32, _i32, _i32, _i32, _p]),
    "dsg_attention_bwd": (C.c_int, [_p, _p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _p]),
    "dsg_grad_norm": (C.c_int, [_p, _i64, _p, _i32, _f, _f, _p, _p]),
    "dsg_adamw_step": (C.c_int, [_p, _p, _p, _p, _i64, _f, _f, _f, _f, _f, _i32, _p, _p]),
}

_lib = None


def load() -> C.CDLL:
    """Load the library (once) and bind every signature; raises DsgError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DsgError(f"{LIB_PATH} not found: the CUDA extension is required (no CPU fallback). "
                       "Build it with `python -m drivescenegen_b200.build`.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().dsg_last_error()
        raise DsgError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")


def launch_count() -> int:
    return int(load().dsg_launch_count())
