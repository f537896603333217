"""ctypes binding of libmvd_b200.so (the C ABI declared in include/mvd_b200.h).

The library is the product: there is no Python/PyTorch fallback for any op. If the shared object is
missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os
import re
import subprocess
from ctypes import c_char_p, c_float, c_int, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmvd_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "mvd_b200.h")

_lib = None


class MVDError(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into mvd_b200/libmvd_b200.so (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j", str(min(8, os.cpu_count() or 1))]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout)
    if r.returncode != 0:
        raise MVDError("building libmvd_b200.so failed")
    return LIB_PATH


def declared_symbols() -> list[str]:
    """Every function the public header declares (used by the CPU-side ABI test)."""
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mvd_[a-z0-9_]+)\s*\(", text)))


_P = c_void_p
_I = c_int
_L = c_int64
_F = c_float

_SIGNATURES = {
    "mvd_last_error": (c_char_p, []),
    "mvd_abi_version": (_I, []),
    "mvd_kernel_launch_count": (_L, []),
    "mvd_set_launch_overlap": (_I, [_I]),
    "mvd_linear_bf16": (_I, [_P, _L, _I, _P, _L, _I, _P, _L, _P, _P, _I, _I, _P, _L, _P, _L, _I, _I, _I, _I, _P]),
    "mvd_conv3x3_bf16": (_I, [_P, _I, _P, _I, _P, _P, _P, _I, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "mvd_linear_ex_bf16": (_I, [_P, _L, _I, _P, _L, _I, _P, _L, _P, _P, _I, _I, _P, _L, _P, _L, _I, _I, _I, _I, _P, _P]),
    "mvd_conv3x3_ex_bf16": (_I, [_P, _I, _P, _I, _P, _P, _P, _I, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P]),
    "mvd_gemm_plan": (_I, [_I] * 9 + [_P] * 3),
    "mvd_gemm_plan_streamk": (_I, [_I] * 8 + [_L] + [_P] * 6),
    "mvd_attention_bf16": (_I, [_P, _L, _L, _P, _L, _L, _P, _L, _L, _P, _L, _L, _I, _I, _I, _I, _F, _P]),
    "mvd_attention_workspace_bytes": (_L, []),
    "mvd_attention_bf16_ws": (_I, [_P, _L, _L, _P, _L, _L, _P, _L, _L, _P, _L, _L, _I, _I, _I, _I, _F, _P, _L, _I, _P]),
    "mvd_groupnorm_workspace_floats": (_L, [_I, _I, _I]),
    "mvd_groupnorm_bf16": (_I, [_P, _I, _P, _I, _P, _P, _P, _I, _I, _I, _F, _I, _P, _L, _P]),
    "mvd_layernorm_bf16": (_I, [_P, _L, _P, _P, _P, _L, _I, _I, _F, _P]),
    "mvd_layernorm_f32": (_I, [_P, _P, _P, _P, _I, _I, _F, _I, _P]),
    "mvd_refnorm_workspace_floats": (_L, [_I]),
    "mvd_refnorm_bf16": (_I, [_P, _P, _I, _I, _I, _I, _P, _L, _P]),
    "mvd_refnorm_replicated_bf16": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _L, _P]),
    "mvd_film_bf16": (_I, [_P, _P, _P, _I, _I, _I, _I, _F, _P]),
    "mvd_small_linear_f32": (_I, [_P, _L, _P, _P, _P, _L, _I, _I, _I, _I, _I, _P]),
    "mvd_timestep_embedding_f32": (_I, [_P, _I, _P, _I, _I, _P]),
    "mvd_camera_front_f32": (_I, [_P, _P, _P, _P, _P, _I, _I, _F, _P]),
    "mvd_conv_in_f32_bf16": (_I, [_P, _I, _P, _I, _F, _P, _P, _P, _I, _I, _I, _I, _P]),
    "mvd_conv_out_bf16_f32": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "mvd_upsample_nearest2x_bf16": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "mvd_head4_to_nchw_f32": (_I, [_P, _L, _P, _I, _L, _P]),
    "mvd_add_bf16": (_I, [_P, _P, _P, _L, _P]),
    "mvd_cast_f32_bf16": (_I, [_P, _P, _L, _P]),
    "mvd_transpose_batched": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "mvd_cfg_ddpm_step_f32": (_I, [_P, _P, _P, _L, _I, _F, _F, _F, _F, _F, _F, _P]),
    "mvd_cfg_ddpm_step_table_f32": (_I, [_P, _P, _P, _L, _I, _F, _P, _P, _P]),
    "mvd_advance_step": (_I, [_P, _P, _P, _I, _P]),
    "mvd_advance_step_rows": (_I, [_P, _P, _P, _I, _P, _P, _I, _P]),
}


def register_signature(name, restype, argtypes):
    _SIGNATURES[name] = (restype, argtypes)


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MVDError(
                f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no fallback path)"
            )
        _lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(_lib, name)
            fn.restype = res
            fn.argtypes = args
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().mvd_last_error()
        raise MVDError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")
