"""ctypes binding of liboat.so (the C ABI declared in include/oat.h).

No fallback: if the shared library is missing or a launch fails, the caller gets a RuntimeError carrying
oat_last_error(). PyTorch only provides device memory (tensor.data_ptr()) and the current CUDA stream.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboat.so")
_lib = None

c_i32, c_i64, c_f32, c_vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p


class OatError(RuntimeError):
    pass


class GemmArgs(ctypes.Structure):
    _fields_ = [
        ("A", c_vp), ("lda", c_i64), ("a_major", c_i32),
        ("B", c_vp), ("ldb", c_i64), ("b_major", c_i32),
        ("M", c_i32), ("N", c_i32), ("K", c_i32),
        ("alpha", c_f32),
        ("bias", c_vp),
        ("scale_cols", c_i32), ("scale", c_f32),
        ("act", c_i32),
        ("aux_bf16", c_vp), ("ld_aux", c_i64),
        ("residual", c_vp), ("ldr", c_i64),
        ("out_f32", c_vp), ("ld_f32", c_i64),
        ("out_bf16", c_vp), ("ld_bf16", c_i64),
        ("out2_bf16", c_vp), ("ld2", c_i64),
        ("accumulate", c_i32),
        ("split_k", c_i32),
        ("rowdot", c_vp), ("ld_rowdot", c_i64),
    ]


def lib():
    """Load liboat.so once. Fails loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OatError(
                "liboat.so is missing (%s). Build it with `python -m oa_transformer_b200.build` "
                "(or __graft_entry__.build()); there is no CPU fallback." % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.oat_last_error.restype = ctypes.c_char_p
        _lib.oat_version.restype = c_i32
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().oat_last_error()
        raise OatError("%s failed (%d): %s" % (what or "liboat call", rc, msg.decode() if msg else "?"))


def stream_ptr():
    import torch
    if not torch.cuda.is_available():
        raise OatError("no CUDA device: the liboat kernels are the only implementation of this path (no CPU fallback)")
    return c_vp(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return c_vp(0) if t is None else c_vp(t.data_ptr())
