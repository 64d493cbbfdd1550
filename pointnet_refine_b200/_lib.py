"""ctypes binding of the C ABI in include/lrn_b200.h (in-tree liblrn_b200.so).

There is no fallback: importing this module raises if the library has not been built
(``python -c "import __graft_entry__ as g; g.build()"``), and every compute call raises
``RuntimeError`` on a non-zero status (e.g. LRN_ERR_UNSUPPORTED_ARCH off sm_100).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblrn_b200.so")
CSRC = os.path.join(_HERE, "csrc")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "lrn_b200.h")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

# enums of include/lrn_b200.h
LRN_OK = 0
PREC_BF16, PREC_TF32 = 0, 1
OUT_POOL, OUT_ARGMAX, OUT_FUSED, OUT_MEMORY = 1, 2, 4, 8
PRECISIONS = {"bf16": PREC_BF16, "tf32": PREC_TF32}

EXPORTS = [
    "lrn_abi_version", "lrn_status_string", "lrn_last_error", "lrn_device_check",
    "lrn_encoder_packed_bytes", "lrn_encoder_fold", "lrn_encoder_workspace_bytes", "lrn_encoder_forward",
    "lrn_head_forward", "lrn_gemm_bias_act",
]


class EncoderParams(C.Structure):
    """struct lrn_encoder_params"""
    _fields_ = (
        [(n, C.c_void_p * 5) for n in ("conv_w", "conv_b", "bn_w", "bn_b", "bn_mean", "bn_var")]
        + [(n, C.c_void_p) for n in ("fusion_w", "fusion_b", "fusion_bn_w", "fusion_bn_b", "fusion_bn_mean",
                                     "fusion_bn_var", "gate0_w", "gate0_b", "gate2_w", "gate2_b", "proj_w", "proj_b")]
        + [("bn_eps", C.c_float)]
    )


def build(verbose: bool = False) -> str:
    """Compile csrc/lrn_abi.cu for sm_100a into the in-tree shared library (nvcc cross-compiles
    without a GPU).  Rebuilds only when a source is newer than the library."""
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))] + [HEADER]
    if os.path.exists(LIB_PATH) and all(os.path.getmtime(s) <= os.path.getmtime(LIB_PATH) for s in srcs):
        return LIB_PATH
    cmd = ["nvcc", *NVCC_FLAGS, "-o", LIB_PATH, os.path.join(CSRC, "lrn_abi.cu")]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB_PATH


def _load():
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA library must be built first "
            "(python -c 'import __graft_entry__ as g; g.build()').  There is no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i64, sz, ci = C.c_void_p, C.c_int64, C.c_size_t, C.c_int
    lib.lrn_abi_version.restype = ci
    lib.lrn_status_string.restype = C.c_char_p
    lib.lrn_status_string.argtypes = [ci]
    lib.lrn_last_error.restype = C.c_char_p
    lib.lrn_device_check.restype = ci
    lib.lrn_encoder_packed_bytes.restype = sz
    lib.lrn_encoder_packed_bytes.argtypes = [ci]
    lib.lrn_encoder_fold.restype = ci
    lib.lrn_encoder_fold.argtypes = [C.POINTER(EncoderParams), ci, vp, sz, vp]
    lib.lrn_encoder_workspace_bytes.restype = sz
    lib.lrn_encoder_workspace_bytes.argtypes = [i64, i64, ci, ci, i64]
    lib.lrn_encoder_forward.restype = ci
    lib.lrn_encoder_forward.argtypes = [vp, ci, vp, i64, i64, ci, vp, vp, vp, vp, i64, vp, sz, vp]
    lib.lrn_head_forward.restype = ci
    lib.lrn_head_forward.argtypes = [vp, vp, vp, vp, vp, i64, vp, vp, vp, vp]
    lib.lrn_gemm_bias_act.restype = ci
    lib.lrn_gemm_bias_act.argtypes = [ci, vp, i64, vp, i64, vp, vp, i64, ci, ci, i64, i64, i64, vp]
    if lib.lrn_abi_version() != 1:
        raise RuntimeError("liblrn_b200.so ABI version mismatch; rebuild it")
    return lib


lib = _load()

# Number of kernels of THIS library enqueued so far by this process (bench.py reports the delta).
launch_counter = 0


def check(status: int, what: str) -> None:
    if status != LRN_OK:
        raise RuntimeError(f"{what} failed: {lib.lrn_status_string(status).decode()} "
                           f"[{status}] {lib.lrn_last_error().decode()}")
