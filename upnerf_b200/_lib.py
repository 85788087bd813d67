"""ctypes binding of libupnerf_b200.so (the C ABI declared in include/upnerf_b200.h).

There is deliberately no fallback: if the shared library has not been built, or a call
returns a non-zero status, this module raises.  torch is used only for device memory and
streams -- tensors are handed to the library as raw device pointers.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import torch

_LIB_PATH = Path(__file__).resolve().parent / "lib" / "libupnerf_b200.so"
_lib = None


class UpnerfError(RuntimeError):
    pass


class Epilogue(C.Structure):
    """Mirror of `upnerf_epilogue` (include/upnerf_b200.h)."""

    _fields_ = [
        ("bias", C.c_void_p),
        ("ray_bias", C.c_void_p),
        ("rows_per_ray", C.c_int),
        ("rank1_row", C.c_void_p),
        ("rank1_col", C.c_void_p),
        ("aux", C.c_void_p),
        ("ldaux", C.c_int64),
        ("aux_mode", C.c_int),
        ("act", C.c_int),
        ("n_heads", C.c_int),
        ("head_w", C.c_void_p),
        ("head_b", C.c_void_p),
        ("head_act", C.c_int),
        ("head_out", C.c_void_p),
    ]


def lib() -> C.CDLL:
    """Load the shared library once; fail loudly when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise UpnerfError(
            f"{_LIB_PATH} is missing: build it with `python -m upnerf_b200.build` "
            "(there is no CPU or PyTorch fallback for the CUDA path)"
        )
    _lib = C.CDLL(str(_LIB_PATH))
    _lib.upnerf_last_error.restype = C.c_char_p
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().upnerf_last_error().decode("utf-8", "replace")
        raise UpnerfError(f"{what} failed (status {status}): {msg}")


def ptr(t: torch.Tensor | None) -> C.c_void_p:
    if t is None:
        return C.c_void_p(0)
    if not t.is_cuda:
        raise UpnerfError("upnerf_b200 kernels take CUDA tensors only (no CPU fallback)")
    return C.c_void_p(t.data_ptr())


def stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def make_epilogue(bias=None, ray_bias=None, rows_per_ray=0, rank1_row=None, rank1_col=None,
                  aux=None, ldaux=0, aux_mode=0, act=0, head_w=None, head_b=None, head_act=0,
                  head_out=None) -> Epilogue:
    ep = Epilogue()
    ep.bias = ptr(bias)
    ep.ray_bias = ptr(ray_bias)
    ep.rows_per_ray = int(rows_per_ray)
    ep.rank1_row = ptr(rank1_row)
    ep.rank1_col = ptr(rank1_col)
    ep.aux = ptr(aux)
    ep.ldaux = int(ldaux)
    ep.aux_mode = int(aux_mode)
    ep.act = int(act)
    ep.n_heads = 0 if head_w is None else int(head_w.shape[0])
    ep.head_w = ptr(head_w)
    ep.head_b = ptr(head_b)
    ep.head_act = int(head_act)
    ep.head_out = ptr(head_out)
    return ep


def _i64(x) -> C.c_int64:
    return C.c_int64(int(x))


def gemm_bf16(A, B, C_out, M, N, K, lda=None, ldb=None, ldc=None, ep: Epilogue | None = None):
    """C[M,N] = epi(A[M,K] @ B[N,K]^T) on tcgen05 (see upnerf_gemm_bf16)."""
    lda = A.stride(0) if lda is None else lda
    ldb = B.stride(0) if ldb is None else ldb
    ldc = C_out.stride(0) if ldc is None else ldc
    st = lib().upnerf_gemm_bf16(ptr(A), _i64(lda), ptr(B), _i64(ldb), ptr(C_out), _i64(ldc),
                                _i64(M), C.c_int(N), C.c_int(K),
                                C.byref(ep) if ep is not None else None, stream_ptr())
    check(st, "upnerf_gemm_bf16")


def wgrad_bf16(dY, X, dW, db, M, N, K, segs, lddy=None, ldx=None, lddw=None):
    """dW[n, map(k)] += dY^T X, db[n] += colsum(dY) on tcgen05 (see upnerf_wgrad_bf16)."""
    lddy = dY.stride(0) if lddy is None else lddy
    ldx = X.stride(0) if ldx is None else ldx
    lddw = dW.stride(0) if lddw is None else lddw
    n = len(segs)
    arr = C.c_int * n
    src = arr(*[s[0] for s in segs])
    ln = arr(*[s[1] for s in segs])
    dst = arr(*[s[2] for s in segs])
    st = lib().upnerf_wgrad_bf16(ptr(dY), _i64(lddy), ptr(X), _i64(ldx), ptr(dW), _i64(lddw),
                                 ptr(db), _i64(M), C.c_int(N), C.c_int(K), C.c_int(n), src, ln, dst,
                                 stream_ptr())
    check(st, "upnerf_wgrad_bf16")


def gemm_f32(A, sa, B, sb, C_out, sc, M, N, K, ep: Epilogue | None = None, accumulate=False,
             split_k=1):
    """Strided fp32 SIMT GEMM (see upnerf_gemm_f32). sa=(sam,sak), sb=(sbn,sbk), sc=(scm,scn)."""
    st = lib().upnerf_gemm_f32(ptr(A), _i64(sa[0]), _i64(sa[1]), ptr(B), _i64(sb[0]), _i64(sb[1]),
                               ptr(C_out), _i64(sc[0]), _i64(sc[1]), _i64(M), _i64(N), _i64(K),
                               C.byref(ep) if ep is not None else None, C.c_int(int(accumulate)),
                               C.c_int(split_k), stream_ptr())
    check(st, "upnerf_gemm_f32")
