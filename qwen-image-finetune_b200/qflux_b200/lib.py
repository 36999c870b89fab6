"""ctypes binding of libqfx_b200.so (include/qfx.h).  Fails loudly when the library or a CUDA device is missing."""
from __future__ import annotations

import ctypes as C
import os

import torch  # noqa: F401  (loads libcudart.so.12 into the process before our library)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libqfx_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing — build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(there is no CPU / eager fallback for the B200 hot path)")

_lib = C.CDLL(LIB_PATH)
_lib.qfx_last_error.restype = C.c_char_p
_lib.qfx_version.restype = C.c_int


class QfxError(RuntimeError):
    pass


def check(rc: int, what: str):
    if rc != 0:
        raise QfxError(f"{what} failed (rc={rc}): {_lib.qfx_last_error().decode()}")


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise QfxError("qflux_b200 ops need CUDA tensors (no CPU fallback)")


def ptr(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def cur_stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


# ---------------------------------------------------------------------------------------------------- GEMM
EPI_BIAS, EPI_GELU, EPI_RESID_GATE, EPI_DGELU = 0, 1, 2, 3


class GemmProblem(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("lda", C.c_int64),
        ("B", C.c_void_p), ("ldb", C.c_int64),
        ("M", C.c_int),
        ("A2", C.c_void_p), ("lda2", C.c_int64),
        ("B2", C.c_void_p), ("ldb2", C.c_int64),
        ("kb2", C.c_int), ("a2_col0", C.c_int),
        ("bias", C.c_void_p),
        ("out", C.c_void_p), ("ldo", C.c_int64),
        ("out2", C.c_void_p), ("ldo2", C.c_int64),
        ("resid", C.c_void_p), ("ldr", C.c_int64),
        ("gate", C.c_void_p), ("ldg", C.c_int64), ("rows_per_batch", C.c_int),
        ("aux", C.c_void_p), ("ldaux", C.c_int64),
    ]


_lib.qfx_gemm_bf16.argtypes = [C.POINTER(GemmProblem), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
                               C.c_int, C.c_void_p]
_lib.qfx_gemm_bf16.restype = C.c_int


def _dp(t):
    return 0 if t is None else t.data_ptr()


def _ld(t):
    return 0 if t is None else t.stride(0)


def gemm_problem(A, B, out, *, A2=None, B2=None, kb2=0, a2_col0=0, bias=None, out2=None, resid=None, gate=None,
                 rows_per_batch=0, aux=None) -> GemmProblem:
    require_cuda(A, B, out, A2, B2, bias, out2, resid, gate, aux)
    for t in (A, B, out, A2, B2, out2, resid, gate, aux):
        assert t is None or (t.dtype == torch.bfloat16 and t.dim() == 2 and t.stride(1) == 1), "bf16 row-major 2-D expected"
    return GemmProblem(_dp(A), _ld(A), _dp(B), _ld(B), A.shape[0], _dp(A2), _ld(A2), _dp(B2), _ld(B2), kb2, a2_col0,
                       _dp(bias), _dp(out), _ld(out), _dp(out2), _ld(out2), _dp(resid), _ld(resid), _dp(gate), _ld(gate),
                       rows_per_batch, _dp(aux), _ld(aux))


def gemm(problems, N, K, *, trans_b=False, epilogue=EPI_BIAS, alpha=1.0, lora_group_n=0, block_n=0):
    arr = (GemmProblem * len(problems))(*problems)
    check(_lib.qfx_gemm_bf16(arr, len(problems), N, K, int(trans_b), epilogue, float(alpha), lora_group_n, block_n,
                             cur_stream()), "qfx_gemm_bf16")
