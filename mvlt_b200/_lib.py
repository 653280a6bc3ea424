"""ctypes binding of the C-ABI library (``include/mvlt_b200.h``).

There is NO CPU fallback: every op in this package enqueues hand-written sm_100a kernels from
``libmvlt_b200.so`` on the current CUDA stream. Loading fails loudly when the library is missing.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import torch

# MVLT_LIB: A/B runs of two in-tree builds of the SAME C-ABI (tuning aid; there is still no non-CUDA path)
_LIBPATH = Path(os.environ.get("MVLT_LIB") or (Path(__file__).resolve().parent / "lib" / "libmvlt_b200.so"))
_lib = None


class MvltError(RuntimeError):
    pass


class GemmDesc(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("B", C.c_void_p), ("D", C.c_void_p), ("D2", C.c_void_p),
        ("bias", C.c_void_p), ("aux", C.c_void_p), ("residual", C.c_void_p), ("rowscale", C.c_void_p), ("rowsum", C.c_void_p),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("a_mn", C.c_int32), ("b_mn", C.c_int32),
        ("lda", C.c_int64), ("ldb", C.c_int64), ("ldd", C.c_int64),
        ("batch1", C.c_int32), ("batch2", C.c_int32),
        ("sA1", C.c_int64), ("sA2", C.c_int64), ("sB1", C.c_int64), ("sB2", C.c_int64),
        ("sD1", C.c_int64), ("sD2", C.c_int64),
        ("alpha", C.c_float),
        ("act", C.c_int32), ("out_f32", C.c_int32), ("atomic_add", C.c_int32),
        ("rows_per_scale", C.c_int32), ("split_k", C.c_int32), ("block_n", C.c_int32),
        ("conv_mode", C.c_int32), ("conv_B", C.c_int32), ("conv_H", C.c_int32), ("conv_W", C.c_int32),
        ("conv_C", C.c_int32), ("conv_pix_stride", C.c_int64), ("conv_batch_stride", C.c_int64),
        ("ln_gamma", C.c_void_p), ("ln_beta", C.c_void_p), ("ln_mean", C.c_void_p), ("ln_rstd", C.c_void_p),
        ("ln_eps", C.c_float), ("conv_R", C.c_int32), ("patch_store", C.c_int32),
    ]


ACT_NONE, ACT_GELU, ACT_DGELU, ACT_GELU_SAVE_GRAD, ACT_MUL_AUX, ACT_SOFTMAX, ACT_SOFTMAX_BWD = 0, 1, 2, 3, 4, 5, 6
CONV_NONE, CONV_A, CONV_BT, CONV_PATCH_A = 0, 1, 2, 3


def lib_path() -> Path:
    return _LIBPATH


def load():
    """dlopen the C-ABI library (once). Raises MvltError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIBPATH.exists():
        raise MvltError(
            f"{_LIBPATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(mvlt_b200 has no CPU / eager fallback)")
    _lib = C.CDLL(str(_LIBPATH), mode=os.RTLD_GLOBAL if hasattr(os, "RTLD_GLOBAL") else C.DEFAULT_MODE)
    _lib.mvlt_last_error.restype = C.c_char_p
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().mvlt_last_error().decode(errors="replace")
        raise MvltError(f"{what} failed (rc={rc}): {msg}")


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_ptr() -> C.c_void_p:
    """cudaStream_t of torch's current stream on the current device (the fast private accessor when torch has it)."""
    if _raw_stream is not None:
        return C.c_void_p(_raw_stream(torch.cuda.current_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t) -> C.c_void_p:
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise MvltError("mvlt_b200 ops need CUDA tensors (sm_100a); there is no CPU fallback")


PARAM_EPOCH = 0       # bumped by every raw-pointer parameter write (own optimizer step, broadcasts): derived tables are stale
WEIGHT_EPOCH = 0      # bumped when such a write could not refresh the engine's bf16 weight copies itself: recast them


def params_written():
    """Tell every engine that parameters were rewritten behind autograd's back (``p.data`` writes such as collectives)."""
    global PARAM_EPOCH, WEIGHT_EPOCH
    PARAM_EPOCH += 1
    WEIGHT_EPOCH += 1


LAUNCHES = 0          # number of C-ABI kernel entry points invoked (bench.py's gpu_launches)
GEMM_FLOPS = 0.0      # executed GEMM flops accumulated while PROFILE is active
GEMM_BYTES = 0.0      # algorithmic (compulsory) operand + result bytes of the same launches
GEMM_LOG = None       # when a list: one (flops, bytes) tuple per GEMM launch, in launch order


def account_gemm(flops: float, nbytes: float, what: str = ""):
    global GEMM_FLOPS, GEMM_BYTES
    GEMM_FLOPS += flops
    GEMM_BYTES += nbytes
    if GEMM_LOG is not None:
        GEMM_LOG.append((flops, nbytes, what))
PROFILE = None        # when a dict: name -> list of (start_event, end_event) recorded around every call
BYTES = None          # when a dict: name -> algorithmic (compulsory) bytes of the memory-bound kernels launched under that name


def account_bytes(tag: str, nbytes: float):
    """Algorithmic bytes of one launch of a memory-bound kernel (bench.py's per-kernel HBM fractions); no-op unless
    ``BYTES`` is a dict."""
    if BYTES is not None:
        BYTES[tag] = BYTES.get(tag, 0.0) + float(nbytes)


_FN = {}


def call(name: str, *args, tag: str = None):
    """Call ``int mvlt_<name>(..., void* stream)`` with the current torch stream appended."""
    global LAUNCHES
    fn = _FN.get(name)
    if fn is None:
        fn = _FN[name] = getattr(load(), "mvlt_" + name)
    LAUNCHES += 1
    if PROFILE is None:
        rc = fn(*args, stream_ptr())
        if rc != 0:
            check(rc, "mvlt_" + name)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(fn(*args, stream_ptr()), "mvlt_" + name)
    e1.record()
    PROFILE.setdefault(tag or name, []).append((e0, e1))
