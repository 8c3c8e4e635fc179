"""ctypes binding of libdwb.so (the C ABI declared in include/dwb.h).

There is no CPU / eager fallback: if the shared object is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdwb.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "dwb.h")

_p = C.c_void_p
_i = C.c_int
_l = C.c_int64
_f = C.c_float

# name -> (restype, argtypes); must mirror include/dwb.h (tests/test_abi_symbols.py checks both directions)
SIGNATURES = {
    "dwb_last_error": (C.c_char_p, []),
    "dwb_abi_version": (_i, []),
    "dwb_check_device": (_i, []),
    "dwb_launch_count": (_l, [_i]),
    "dwb_set_row_walk": (_i, [_i]),
    "dwb_scale_bf16_dev": (_i, [_p, _l, _p, _p]),
    "dwb_gemm_bf16": (_i, [_p, _l, _i, _p, _l, _i, _p, _l, _i, _i, _i, _i, _p, _i, _f, _i, _i, _p]),
    "dwb_attention_fwd": (_i, [_p, _l, _p, _l, _p, _l, _p, _l, _p, _i, _i, _i, _i, _i, _i, _f, _p]),
    "dwb_attention_fwd_tc": (_i, [_p, _l, _p, _l, _p, _l, _p, _l, _p, _i, _i, _i, _i, _i, _i, _f, _p]),
    "dwb_attention_bwd": (_i, [_p, _l, _p, _l, _p, _l, _p, _l, _p, _l, _p, _p, _p, _p, _l, _p, _l,
                               _i, _i, _i, _i, _i, _i, _f, _p]),
    "dwb_attention_bwd_tc": (_i, [_p, _l, _p, _l, _p, _l, _p, _l, _p, _l, _p, _p, _p, _p, _l, _p, _l,
                                  _i, _i, _i, _i, _i, _i, _f, _p]),
    "dwb_add_layernorm": (_i, [_p, _i, _p, _p, _p, _p, _p, _p, _p, _i, _i, _f, _p]),
    "dwb_layernorm_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _p]),
    "dwb_cast_f32_to_bf16": (_i, [_p, _l, _p, _l, _i, _i, _f, _p]),
    "dwb_cast_bf16_to_f32": (_i, [_p, _l, _p, _l, _i, _i, _p]),
    "dwb_conv_weight_to_kc_bf16": (_i, [_p, _p, _i, _i, _p]),
    "dwb_conv_wgrad_kc_to_ck": (_i, [_p, _p, _i, _i, _i, _p]),
    "dwb_im2col_conv1": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "dwb_im2col_conv2": (_i, [_p, _p, _i, _i, _i, _p]),
    "dwb_col2im_conv2_gelu_bwd": (_i, [_p, _p, _p, _i, _i, _i, _p]),
    "dwb_embed_fwd": (_i, [_p, _p, _p, _i, _p, _i, _i, _i, _i, _p]),
    "dwb_embed_bwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "dwb_gemm_skinny_bf16": (_i, [_p, _l, _p, _l, _p, _l, _i, _i, _i, _i, _p, _i, _p]),
    "dwb_embed_decode": (_i, [_p, _i, _p, _p, _p, _i, _p, _i, _i, _i, _p]),
    "dwb_attention_decode": (_i, [_p, _l, _p, _p, _l, _p, _p, _l, _i, _p, _l, _i, _i, _i, _i, _p, _f, _p]),
    "dwb_greedy_pick": (_i, [_p, _l, _i, _p, _p, _i, _p, _i, _i, _p, _l, _l, _p, _i, _p]),
    "dwb_greedy_pick_timestamps": (_i, [_p, _l, _i, _p, _p, _i, _p, _i, _i, _p, _l, _l, _p, _i, _i, _i, _p]),
    "dwb_decode_advance": (_i, [_p, _p, _i, _p, _p]),
    "dwb_collate_labels": (_i, [_p, _p, _i, _i, _l, _p, _p, _p]),
    "dwb_colsum_bf16": (_i, [_p, _l, _p, _i, _i, _i, _p]),
    "dwb_gelu_bwd": (_i, [_p, _p, _p, _l, _p]),
    "dwb_gelu_fwd": (_i, [_p, _p, _l, _p]),
    "dwb_kd_loss_workspace_bytes": (_l, [_i]),
    "dwb_kd_loss": (_i, [_p, _p, _l, _p, _i, _i, _f, _f, _f, _p, _p, _l, _p, _p]),
    "dwb_set_tail_grid": (_i, [_i]),
    "dwb_allreduce_symm": (_i, [_p, _p, _i, _i, _l, _i, _p]),
    "dwb_grad_sumsq": (_i, [_p, _l, _p, _p]),
    "dwb_adamw_step": (_i, [_p, _p, _p, _p, _p, _l, _f, _f, _f, _f, _f, _i, _p, _f, _f, _i, _p]),
    "dwb_logmel_plan_create": (_i, [_p, _i, _i, C.POINTER(_p)]),
    "dwb_logmel_plan_destroy": (_i, [_p]),
    "dwb_logmel": (_i, [_p, _p, _i, _i, _p, _p]),
    "dwb_logmel_tc_plan_create": (_i, [_p, _i, _i, C.POINTER(_p)]),
    "dwb_logmel_tc_plan_destroy": (_i, [_p]),
    "dwb_logmel_tc_workspace_bytes": (_l, [_i, _i]),
    "dwb_logmel_tc": (_i, [_p, _p, _i, _i, _p, _p, _p]),
}

_NO_STATUS = {"dwb_last_error", "dwb_abi_version", "dwb_kd_loss_workspace_bytes", "dwb_launch_count", "dwb_logmel_tc_workspace_bytes"}


class DwbError(RuntimeError):
    pass


def header_symbols(path: str = HEADER_PATH):
    """Function names declared in include/dwb.h."""
    with open(path) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(dwb_[a-z0-9_]+)\s*\(", src)))


_lib = None

def launch_count(reset: bool = False) -> int:
    """Kernel launches issued by libdwb.so since the last reset (counted inside the library at every <<<>>>)."""
    return int(load().dwb_launch_count(1 if reset else 0))


def load() -> C.CDLL:
    """Load libdwb.so (built by __graft_entry__.build() / make -C distil_whisper_b200/csrc)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DwbError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  distil_whisper_b200 has no CPU or eager fallback.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError here == ABI drift; let it surface
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def call(name: str, *args):
    """Invoke an entry point; non-zero status -> DwbError carrying dwb_last_error()."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if name in _NO_STATUS:
        return rc
    if rc != 0:
        raise DwbError(f"{name} failed ({rc}): {lib.dwb_last_error().decode(errors='replace')}")
    return rc
