"""Thin Python wrappers (no autograd) over the C-ABI kernels. Tensors are torch CUDA tensors used purely as
device-memory handles; all arithmetic happens in ``libmvlt_b200.so``.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import ACT_DGELU, ACT_GELU, ACT_NONE, GemmDesc, call, ptr, require_cuda

BF16 = torch.bfloat16
F32 = torch.float32


def _operand(t: torch.Tensor, name: str):
    """Logical [..., R, K] bf16 view -> (mn_major, ld, batch sizes, batch strides)."""
    if t.dtype != BF16:
        raise _lib.MvltError(f"gemm operand {name} must be bf16, got {t.dtype}")
    if t.dim() < 2 or t.dim() > 4:
        raise _lib.MvltError(f"gemm operand {name} must be 2-4 D")
    R, K = t.shape[-2], t.shape[-1]
    sr, sk = t.stride(-2), t.stride(-1)
    if sk == 1 or K == 1:
        mn, ld = 0, sr if R > 1 else max(sr, K)
    elif sr == 1 or R == 1:
        mn, ld = 1, sk if K > 1 else max(sk, R)
    else:
        raise _lib.MvltError(f"gemm operand {name} has no unit stride: shape {tuple(t.shape)} strides {t.stride()}")
    bs = list(t.shape[:-2])
    st = list(t.stride()[:-2])
    while len(bs) < 2:
        bs.insert(0, 1)
        st.insert(0, 0)
    return mn, ld, bs, st


def gemm(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, *, alpha: float = 1.0,
         bias: Optional[torch.Tensor] = None, act: int = ACT_NONE, preact_out: Optional[torch.Tensor] = None,
         aux: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
         rowscale: Optional[torch.Tensor] = None, rows_per_scale: int = 0, atomic_add: bool = False,
         split_k: int = 0, block_n: int = 0, impl: str = "tcgen05") -> torch.Tensor:
    """out[..., M, N] = epilogue(alpha * a[..., M, K] @ b[..., N, K]^T)   (see csrc/gemm_desc.h).

    ``a``/``b`` may be arbitrary 2-4 D views with one unit stride among the last two dims (K-major or
    MN-major); leading dims are batch dims (broadcast with stride 0 / size 1 allowed).
    """
    require_cuda(a, b, out)
    a_mn, lda, abs_, ast = _operand(a, "a")
    b_mn, ldb, bbs, bst = _operand(b, "b")
    M, K = a.shape[-2], a.shape[-1]
    N = b.shape[-2]
    if b.shape[-1] != K:
        raise _lib.MvltError(f"gemm K mismatch: a {tuple(a.shape)} b {tuple(b.shape)}")
    if out.shape[-2] != M or out.shape[-1] != N or (out.stride(-1) != 1 and N > 1):
        raise _lib.MvltError(f"gemm out shape/stride mismatch: {tuple(out.shape)} {out.stride()} vs M={M} N={N}")
    obs = list(out.shape[:-2])
    ost = list(out.stride()[:-2])
    while len(obs) < 2:
        obs.insert(0, 1)
        ost.insert(0, 0)
    b1, b2 = obs
    for (bs, st, nm) in ((abs_, ast, "a"), (bbs, bst, "b")):
        for i in range(2):
            if bs[i] != obs[i]:
                if bs[i] != 1:
                    raise _lib.MvltError(f"gemm batch mismatch on {nm}: {bs} vs out {obs}")
                st[i] = 0
    d = GemmDesc()
    d.A, d.B, d.D = a.data_ptr(), b.data_ptr(), out.data_ptr()
    d.D2 = preact_out.data_ptr() if preact_out is not None else None
    d.bias = bias.data_ptr() if bias is not None else None
    d.aux = aux.data_ptr() if aux is not None else None
    d.residual = residual.data_ptr() if residual is not None else None
    d.rowscale = rowscale.data_ptr() if rowscale is not None else None
    d.M, d.N, d.K = M, N, K
    d.a_mn, d.b_mn = a_mn, b_mn
    d.lda, d.ldb, d.ldd = lda, ldb, out.stride(-2) if M > 1 else max(out.stride(-2), N)
    d.batch1, d.batch2 = b1, b2
    d.sA1, d.sA2 = ast
    d.sB1, d.sB2 = bst
    d.sD1, d.sD2 = ost
    d.alpha = alpha
    d.act = act
    if out.dtype == F32:
        d.out_f32 = 1
    elif out.dtype == BF16:
        d.out_f32 = 0
    else:
        raise _lib.MvltError(f"gemm out dtype {out.dtype} unsupported")
    d.atomic_add = 1 if atomic_add else 0
    d.rows_per_scale = rows_per_scale
    d.split_k = split_k
    d.block_n = block_n
    if bias is not None and (bias.dtype != F32 or bias.numel() != N):
        raise _lib.MvltError("gemm bias must be fp32 [N]")
    if residual is not None and (residual.dtype != F32 or residual.stride() != out.stride()):
        raise _lib.MvltError("gemm residual must be fp32 with out's strides")
    for t in (aux, preact_out):
        if t is not None and (t.dtype != BF16 or t.stride() != out.stride()):
            raise _lib.MvltError("gemm aux/preact must be bf16 with out's strides")
    call("gemm" if impl == "tcgen05" else "gemm_ref", C.byref(d))
    return out
