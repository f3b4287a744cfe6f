"""ctypes binding of libdiffbir_b200.so (the C ABI declared in include/diffbir_b200.h).

The product path fails loudly when the CUDA library is missing: there is no CPU or
PyTorch-eager fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import torch

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libdiffbir_b200.so"
_lib = None
_launches = 0  # kernels launched through this binding (bench.py reports it)


class DbirError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("b", C.c_void_p), ("out", C.c_void_p),
        ("bias", C.c_void_p), ("rowvec", C.c_void_p), ("residual", C.c_void_p),
        ("lda", C.c_int64), ("ldb", C.c_int64), ("ldo", C.c_int64), ("ldr", C.c_int64),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("a_mode", C.c_int32),
        ("img_n", C.c_int32), ("img_h", C.c_int32), ("img_w", C.c_int32), ("img_c", C.c_int32),
        ("ksize", C.c_int32),
        ("rows_per_vec", C.c_int32), ("out_kind", C.c_int32), ("act", C.c_int32),
        ("geglu", C.c_int32), ("force_bn", C.c_int32),
        ("alpha", C.c_float), ("act_param", C.c_float),
    ]


def load():
    """Loads the shared library (building is the job of __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise DbirError(
            f"{LIB_PATH} not found: build it with `python -m diffbir_b200.build` "
            "(there is no CPU fallback)")
    lib = C.CDLL(str(LIB_PATH))
    lib.dbir_last_error.restype = C.c_char_p
    lib.dbir_version.restype = C.c_char_p
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().dbir_last_error().decode(errors="replace")
        raise DbirError(f"{what} failed (rc={rc}): {msg}")


def count_launch(n: int = 1):
    global _launches
    _launches += n


def launches() -> int:
    return _launches


def operand_dtype() -> torch.dtype:
    return torch.float16 if load().dbir_operand_kind() == 1 else torch.bfloat16


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


ACT = {None: 0, "none": 0, "gelu": 1, "lrelu": 2, "silu": 3}


def gemm(a, b, out, *, M, N, K, bias=None, rowvec=None, rows_per_vec=0, residual=None,
         lda=0, ldb=0, ldo=None, ldr=None, conv=None, act=None, act_param=0.0, alpha=1.0,
         geglu=False, force_bn=0):
    """out = residual + alpha * act(A @ B^T + bias + rowvec). conv = (n, h, w, c, ksize)."""
    lib = load()
    g = GemmArgs()
    g.a, g.b, g.out = a.data_ptr(), b.data_ptr(), out.data_ptr()
    g.bias, g.rowvec, g.residual = _ptr(bias), _ptr(rowvec), _ptr(residual)
    n_out = N // 2 if geglu else N
    g.lda, g.ldb = lda, ldb
    g.ldo = n_out if ldo is None else ldo
    g.ldr = n_out if ldr is None else ldr
    g.M, g.N, g.K = M, N, K
    if conv is not None:
        g.a_mode = 1
        g.img_n, g.img_h, g.img_w, g.img_c, g.ksize = conv
    else:
        g.a_mode = 0
    g.rows_per_vec = rows_per_vec
    g.out_kind = 0 if out.dtype == torch.float32 else 1
    g.act = ACT[act]
    g.act_param = act_param
    g.alpha = alpha
    g.geglu = 1 if geglu else 0
    g.force_bn = force_bn
    check(lib.dbir_gemm(C.byref(g), C.c_void_p(stream_ptr())), "dbir_gemm")
    count_launch()
