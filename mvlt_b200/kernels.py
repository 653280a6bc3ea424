"""Thin Python wrappers (no autograd) over the C-ABI kernels. Tensors are torch CUDA tensors used purely as
device-memory handles; all arithmetic happens in ``libmvlt_b200.so``.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import CONV_PATCH_A  # noqa: F401
from ._lib import (ACT_DGELU, ACT_GELU, ACT_GELU_SAVE_GRAD, ACT_MUL_AUX, ACT_NONE, ACT_SOFTMAX, ACT_SOFTMAX_BWD, CONV_A,
                   CONV_BT, GemmDesc, call, ptr, require_cuda)

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
    if K == 1 and R > 1 and sr == 1 and sk >= R:
        mn, ld = 1, sk      # a transposed single row (dy^T of a one-sample batch): keep the MN-major reading of its strides
    elif sk == 1 or K == 1:
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


def _build_gemm_desc(a, b, out, alpha, bias, act, preact_out, aux, residual, rowscale, rows_per_scale, atomic_add, split_k,
                     block_n, rowsum, ln_out=None):
    """Validate one GEMM signature and build its descriptor (everything except the device pointers)."""
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
    if rowsum is not None and (rowsum.dtype != F32 or rowsum.numel() != M or not rowsum.is_contiguous()):
        raise _lib.MvltError("gemm rowsum must be contiguous fp32 [M]")
    if out.dtype not in (F32, BF16):
        raise _lib.MvltError(f"gemm out dtype {out.dtype} unsupported")
    if bias is not None and (bias.dtype != F32 or bias.numel() != N):
        raise _lib.MvltError("gemm bias must be fp32 [N]")
    if residual is not None and (residual.dtype != F32 or residual.stride() != out.stride()):
        raise _lib.MvltError("gemm residual must be fp32 with out's strides")
    for t in (aux, preact_out):
        if t is not None and (t.dtype != BF16 or t.stride() != out.stride()):
            raise _lib.MvltError("gemm aux/preact must be bf16 with out's strides")
    d = GemmDesc()
    d.M, d.N, d.K = M, N, K
    d.a_mn, d.b_mn = a_mn, b_mn
    d.lda, d.ldb, d.ldd = lda, ldb, out.stride(-2) if M > 1 else max(out.stride(-2), N)
    d.batch1, d.batch2 = b1, b2
    d.sA1, d.sA2 = ast
    d.sB1, d.sB2 = bst
    d.sD1, d.sD2 = ost
    d.alpha = alpha
    d.act = act
    d.out_f32 = 1 if out.dtype == F32 else 0
    d.atomic_add = 1 if atomic_add else 0
    d.rows_per_scale = rows_per_scale
    d.split_k = split_k
    d.block_n = block_n
    nb = b1 * b2
    extra = sum(M * N * nb * t.element_size() for t in (aux, preact_out, residual, ln_out) if t is not None)
    acct = (2.0 * M * N * K * nb, (M * K + N * K) * 2.0 * nb + M * N * nb * out.element_size() + extra,
            f"gemm M={M} N={N} K={K} b={nb} amn={a_mn} bmn={b_mn} out={'f32' if out.dtype == F32 else 'bf16'} act={act} "
            f"res={int(residual is not None)} atomic={int(atomic_add)} split={split_k} rowsum={int(rowsum is not None)}"
            + (" ln=1" if ln_out is not None else ""))
    return d, C.byref(d), acct


_GEMM_CACHE = {}


def gemm(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, *, alpha: float = 1.0,
         bias: Optional[torch.Tensor] = None, act: int = ACT_NONE, preact_out: Optional[torch.Tensor] = None,
         aux: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
         rowscale: Optional[torch.Tensor] = None, rows_per_scale: int = 0, atomic_add: bool = False,
         split_k: int = 0, block_n: int = 0, rowsum: Optional[torch.Tensor] = None,
         ln=None, impl: str = "tcgen05") -> torch.Tensor:
    """out[..., M, N] = epilogue(alpha * a[..., M, K] @ b[..., N, K]^T)   (see csrc/gemm_desc.h).

    ``a``/``b`` may be arbitrary 2-4 D views with one unit stride among the last two dims (K-major or
    MN-major); leading dims are batch dims (broadcast with stride 0 / size 1 allowed).

    ``ln = (gamma, beta, out_bf16, mean, rstd, eps)``: LayerNorm of the finished fp32 rows fused into the residual epilogue
    (N = 64 | 128; ``mean`` / ``rstd`` fp32 [M] or None): ``out_bf16`` = the normalised rows, the next GEMM's operand.

    A training step repeats the same ~250 GEMM signatures every iteration: the validated descriptor of each signature
    (geometry, strides, dtypes, epilogue options) is cached and only its device pointers are refreshed per call, which
    keeps the host-side cost of a launch to a few microseconds.
    """
    if not (a.is_cuda and b.is_cuda and out.is_cuda):
        raise _lib.MvltError("mvlt_b200 ops need CUDA tensors (sm_100a); there is no CPU fallback")
    key = (a.shape, a.stride(), b.shape, b.stride(), out.shape, out.stride(), a.dtype, b.dtype, out.dtype, alpha, act,
           rows_per_scale, atomic_add, split_k, block_n,
           None if bias is None else (bias.dtype, bias.shape), None if preact_out is None else (preact_out.dtype, preact_out.stride()),
           None if aux is None else (aux.dtype, aux.stride()), None if residual is None else (residual.dtype, residual.stride()),
           rowscale is None, None if rowsum is None else (rowsum.dtype, rowsum.shape, rowsum.stride()),
           None if ln is None else (ln[2].stride(), ln[3] is None, ln[4] is None, float(ln[5])))
    ent = _GEMM_CACHE.get(key)
    if ent is None:
        if ln is not None:
            g_, b_, o_, m_, r_, _eps = ln
            if (residual is None or out.dtype != F32 or o_.dtype != BF16 or o_.stride() != out.stride() or g_.dtype != F32
                    or b_.dtype != F32 or g_.numel() != out.shape[-1] or b_.numel() != out.shape[-1] or preact_out is not None
                    or any(t is not None and (t.dtype != F32 or t.numel() != out.shape[-2] or not t.is_contiguous()) for t in (m_, r_))):
                raise _lib.MvltError("gemm ln=(gamma, beta, out_bf16, mean, rstd, eps): needs residual, fp32 out, bf16 out_bf16 with "
                                     "out's strides, fp32 gamma / beta [N], fp32 mean / rstd [M]")
        for t in (bias, preact_out, aux, residual, rowscale, rowsum) + (tuple(ln[:5]) if ln is not None else ()):
            if t is not None and not t.is_cuda:
                raise _lib.MvltError("mvlt_b200 ops need CUDA tensors (sm_100a); there is no CPU fallback")
        ent = _build_gemm_desc(a, b, out, alpha, bias, act, preact_out, aux, residual, rowscale, rows_per_scale, atomic_add,
                               split_k, block_n, rowsum, None if ln is None else ln[2])
        if len(_GEMM_CACHE) > 4096:
            _GEMM_CACHE.clear()
        _GEMM_CACHE[key] = ent
    d, ref, acct = ent
    d.A = a.data_ptr()
    d.B = b.data_ptr()
    d.D = out.data_ptr()
    d.D2 = None if preact_out is None else preact_out.data_ptr()
    d.bias = None if bias is None else bias.data_ptr()
    d.aux = None if aux is None else aux.data_ptr()
    d.residual = None if residual is None else residual.data_ptr()
    d.rowscale = None if rowscale is None else rowscale.data_ptr()
    d.rowsum = None if rowsum is None else rowsum.data_ptr()
    if ln is not None:
        d.ln_gamma, d.ln_beta, d.D2 = ln[0].data_ptr(), ln[1].data_ptr(), ln[2].data_ptr()
        d.ln_mean = None if ln[3] is None else ln[3].data_ptr()
        d.ln_rstd = None if ln[4] is None else ln[4].data_ptr()
        d.ln_eps = float(ln[5])
    else:
        d.ln_gamma = None
    if _lib.PROFILE is not None or _lib.GEMM_LOG is not None:
        _lib.account_gemm(*acct)
    call("gemm" if impl == "tcgen05" else "gemm_ref", ref)
    return out


def _conv_fields(d, x, B, H, W, Cdim, pix_stride, batch_stride, mode):
    if x.dtype != BF16:
        raise _lib.MvltError("implicit convolution needs a bf16 NHWC operand")
    d.conv_mode, d.conv_B, d.conv_H, d.conv_W, d.conv_C = mode, B, H, W, Cdim
    d.conv_pix_stride, d.conv_batch_stride = pix_stride, batch_stride


def conv3x3_gemm(x, B, H, W, Cdim, pix_stride, batch_stride, w, out, *, residual=None, block_n: int = 0):
    """out[(b,y,x), n] = sum_{tap,c} X[b, y+tap//3-1, x+tap%3-1, c] * w[n, tap*C + c]  (3x3, stride 1, zero pad 1).

    ``x``: bf16 NHWC storage addressed by (batch_stride, W*pix_stride, pix_stride, 1); the im2col matrix is never
    materialised (csrc/gemm_desc.h, MVLT_CONV_A). ``w``: bf16 [N, 9*C] K-major. ``out``: [B*H*W, N] bf16 / fp32.
    """
    require_cuda(x, w, out)
    M, N, K = B * H * W, w.shape[0], 9 * Cdim
    if w.shape[1] != K or w.stride(1) != 1 or out.shape[0] != M or out.shape[1] != N or out.stride(1) != 1:
        raise _lib.MvltError(f"conv3x3_gemm shape mismatch: w {tuple(w.shape)} out {tuple(out.shape)} M={M} K={K}")
    d = GemmDesc()
    d.A, d.B, d.D = x.data_ptr(), w.data_ptr(), out.data_ptr()
    d.residual = residual.data_ptr() if residual is not None else None
    d.M, d.N, d.K = M, N, K
    d.a_mn, d.b_mn = 0, 0
    d.lda, d.ldb, d.ldd = K, w.stride(0), out.stride(0)
    d.batch1 = d.batch2 = 1
    d.alpha = 1.0
    d.out_f32 = 1 if out.dtype == F32 else 0
    d.block_n = block_n
    _conv_fields(d, x, B, H, W, Cdim, pix_stride, batch_stride, CONV_A)
    if residual is not None and (residual.dtype != F32 or residual.stride() != out.stride()):
        raise _lib.MvltError("conv3x3_gemm residual must be fp32 with out's strides")
    if _lib.PROFILE is not None or _lib.GEMM_LOG is not None:   # algorithmic bytes: X once (not the 9x im2col matrix) + weights + output (+ residual)
        _lib.account_gemm(2.0 * M * N * K, (M * Cdim + N * K) * 2.0 + M * N * out.element_size() * (2 if residual is not None else 1),
                          f"conv3x3 M={M} N={N} K={K} HxW={H}x{W} out={'f32' if out.dtype == F32 else 'bf16'} res={int(residual is not None)}")
    call("gemm", C.byref(d))
    return out


def conv_patch_supported(H, W, Cdim, R):
    """Geometry the in-place patch view handles (csrc/gemm_desc.h MVLT_CONV_PATCH_A): 64 output pixels per image."""
    return R > 1 and H % R == 0 and W % R == 0 and (H // R) * (W // R) == 64 and (R * Cdim) % 64 == 0


def conv_patch_gemm(x, B, H, W, Cdim, R, batch_stride, w, out, *, bias=None):
    """Convolution with kernel = stride = R over NHWC bf16 ``x`` (pixel stride Cdim, ``batch_stride`` elements between images) as
    ONE GEMM whose A operand is the patch matrix read in place through a 5-D TMA view -- no patchify pass, no patch buffer:
    out[(b, oy, ox), n] = sum_k patches[(b, oy, ox), k] * w[n, k] (+ bias), k = (ky*R + kx)*C + c. ``w``: bf16 [N, R*R*C]
    (the engine's permuted conv weight), ``out``: bf16 / fp32 [B*64, N]."""
    require_cuda(x, w, out, bias)
    M, N, K = B * 64, w.shape[0], R * R * Cdim
    if (x.dtype != BF16 or w.dtype != BF16 or not conv_patch_supported(H, W, Cdim, R) or w.shape[1] != K or w.stride(1) != 1
            or tuple(out.shape) != (M, N) or out.stride(1) != 1 or (bias is not None and (bias.dtype != F32 or bias.numel() != N))):
        raise _lib.MvltError(f"conv_patch_gemm: unsupported operands (B={B} H={H} W={W} C={Cdim} R={R}, w {tuple(w.shape)}, out {tuple(out.shape)})")
    d = GemmDesc()
    d.A, d.B, d.D = x.data_ptr(), w.data_ptr(), out.data_ptr()
    d.bias = None if bias is None else bias.data_ptr()
    d.M, d.N, d.K = M, N, K
    d.a_mn, d.b_mn = 0, 0
    d.lda, d.ldb, d.ldd = K, w.stride(0), out.stride(0)
    d.batch1 = d.batch2 = 1
    d.alpha = 1.0
    d.out_f32 = 1 if out.dtype == F32 else 0
    d.conv_mode, d.conv_B, d.conv_H, d.conv_W, d.conv_C, d.conv_R = CONV_PATCH_A, B, H, W, Cdim, R
    d.conv_pix_stride, d.conv_batch_stride = Cdim, batch_stride
    if _lib.PROFILE is not None or _lib.GEMM_LOG is not None:
        _lib.account_gemm(2.0 * M * N * K, (M * K + N * K) * 2.0 + M * N * out.element_size(),
                          f"conv_patch M={M} N={N} K={K} HxW={H}x{W} R={R} out={'f32' if out.dtype == F32 else 'bf16'}")
    call("gemm", C.byref(d))
    return out


def conv_patch_dgrad(dy, w, dx, B, H, W, Cdim, R, batch_stride):
    """Input gradient of the kernel = stride = R convolution: dX[b, oy*R + ky, ox*R + kx, c] = sum_n dy[(b, oy, ox), n] *
    w[n, (ky*R + kx)*C + c], written straight into the fp32 NHWC rows of ``dx`` (pixel stride Cdim, ``batch_stride`` elements
    between images) through the 5-D TMA view -- the "unpatchify" pass and the bf16 patch-gradient buffer disappear.
    ``dy``: bf16 [B*64, N] rows, ``w``: bf16 [N, R*R*C]. Every image pixel is written exactly once; other rows of ``dx`` untouched."""
    require_cuda(dy, w, dx)
    M, Nn, K = B * 64, R * R * Cdim, w.shape[0]
    if (dy.dtype != BF16 or w.dtype != BF16 or dx.dtype != F32 or not conv_patch_supported(H, W, Cdim, R) or H // R != 8
            or tuple(dy.shape) != (M, K) or dy.stride(1) != 1 or w.shape[1] != Nn or w.stride(1) != 1):
        raise _lib.MvltError(f"conv_patch_dgrad: unsupported operands (B={B} H={H} W={W} C={Cdim} R={R}, dy {tuple(dy.shape)}, w {tuple(w.shape)})")
    d = GemmDesc()
    d.A, d.B, d.D = dy.data_ptr(), w.data_ptr(), dx.data_ptr()
    d.M, d.N, d.K = M, Nn, K
    d.a_mn, d.b_mn = 0, 1                      # B = w^T: [N = R*R*C, K = Co] read MN-major from w [Co, R*R*C]
    d.lda, d.ldb, d.ldd = dy.stride(0), w.stride(0), Nn
    d.batch1 = d.batch2 = 1
    d.alpha = 1.0
    d.out_f32 = 1
    d.patch_store = 1
    d.conv_B, d.conv_H, d.conv_W, d.conv_C, d.conv_R = B, H, W, Cdim, R
    d.conv_pix_stride, d.conv_batch_stride = Cdim, batch_stride
    if _lib.PROFILE is not None or _lib.GEMM_LOG is not None:
        _lib.account_gemm(2.0 * M * Nn * K, (M * K + Nn * K) * 2.0 + M * Nn * 4.0, f"conv_patch_dgrad M={M} N={Nn} K={K} HxW={H}x{W} R={R}")
    call("gemm", C.byref(d))
    return dx


def conv3x3_wgrad(dy, x, B, H, W, Cdim, pix_stride, batch_stride, out, *, split_k: int = 0):
    """out[co, tap*C + c] += sum_{b,y,x} dy[(b,y,x), co] * X[b, y+tap//3-1, x+tap%3-1, c]  (fp32 atomic accumulation).

    ``dy``: bf16 [B*H*W, Co] contiguous rows; ``x`` as in conv3x3_gemm (MVLT_CONV_BT: col^T is the MN-major B operand).
    """
    require_cuda(dy, x, out)
    Kp, Co, N = B * H * W, dy.shape[1], 9 * Cdim
    if dy.shape[0] != Kp or dy.stride(1) != 1 or out.shape[0] != Co or out.shape[1] != N or out.dtype != F32:
        raise _lib.MvltError(f"conv3x3_wgrad shape mismatch: dy {tuple(dy.shape)} out {tuple(out.shape)}")
    d = GemmDesc()
    d.A, d.B, d.D = dy.data_ptr(), x.data_ptr(), out.data_ptr()
    d.M, d.N, d.K = Co, N, Kp
    d.a_mn, d.b_mn = 1, 1
    d.lda, d.ldb, d.ldd = dy.stride(0), N, out.stride(0)
    d.batch1 = d.batch2 = 1
    d.alpha = 1.0
    d.out_f32, d.atomic_add = 1, 1
    d.split_k = split_k
    _conv_fields(d, x, B, H, W, Cdim, pix_stride, batch_stride, CONV_BT)
    if _lib.PROFILE is not None or _lib.GEMM_LOG is not None:
        _lib.account_gemm(2.0 * Co * N * Kp, (Kp * Co + Kp * Cdim) * 2.0 + Co * N * 4.0,
                          f"conv3x3_wgrad M={Co} N={N} K={Kp} HxW={H}x{W} split={split_k}")
    call("gemm", C.byref(d))
    return out


# ------------------------------------------------------------------------------------------------------
# Non-GEMM kernels
# ------------------------------------------------------------------------------------------------------
_I3 = C.c_int * 3


def _map(m):
    """(group, stride, offset) row map -> C int[3]; None = identity."""
    if m is None:
        return _I3(0, 0, 0)
    return _I3(int(m[0]), int(m[1]), int(m[2]))


def _f32(t):
    if t.dtype == F32:
        return 1
    if t.dtype == BF16:
        return 0
    raise _lib.MvltError(f"expected fp32/bf16 tensor, got {t.dtype}")


def layernorm_fwd(x, gamma, beta, y, eps, rows, Cdim, xmap=None, ymap=None, post_add=None, mean=None, rstd=None, chain=None):
    """``chain = (gamma2, beta2, y2_bf16, mean2, rstd2, eps2)``: a second LayerNorm of the row just written, from the same launch
    (y2 uses y's row map; mean2 / rstd2 are indexed by the mapped row)."""
    require_cuda(x, y)
    g2 = b2 = y2 = m2 = r2 = None
    eps2 = 0.0
    if chain is not None:
        g2, b2, y2, m2, r2, eps2 = chain
        require_cuda(g2, b2, y2, m2, r2)
        if y2.dtype != BF16 or g2.dtype != F32 or b2.dtype != F32 or g2.numel() != Cdim or b2.numel() != Cdim:
            raise _lib.MvltError("layernorm_fwd chain=(gamma2, beta2, y2_bf16, mean2, rstd2, eps2): fp32 gamma2 / beta2 [C], bf16 y2")
    if _lib.BYTES is not None:
        _lib.account_bytes("layernorm_fwd", rows * Cdim * (x.element_size() + y.element_size() + (2 if chain is not None else 0)))
    call("layernorm_fwd", ptr(x), _f32(x), _map(xmap), ptr(gamma), ptr(beta), ptr(y), _f32(y), _map(ymap),
         ptr(post_add), ptr(mean), ptr(rstd), C.c_int(rows), C.c_int(Cdim), C.c_float(eps), ptr(g2), ptr(b2), ptr(y2), ptr(m2),
         ptr(r2), C.c_float(float(eps2)))


def layernorm_bwd(dy, x, mean, rstd, gamma, dx, rows, Cdim, dymap=None, xmap=None, dxmap=None, dx_add=None,
                  dgamma=None, dbeta=None, dx_bf16=None, rowscale=None, rows_per_scale=0):
    """``dx_bf16`` (optional, dx's row layout): bf16 copy of dx times rowscale[row // rows_per_scale] (drop-path), i.e.
    the A operand of the backward GEMMs that consume dx next, written by the same pass."""
    require_cuda(dy, x, dx)
    if _lib.BYTES is not None:
        _lib.account_bytes("layernorm_bwd", rows * Cdim * (dy.element_size() + x.element_size() + dx.element_size()
                                                           + (4 if dx_add is not None else 0) + (2 if dx_bf16 is not None else 0)))
    call("layernorm_bwd", ptr(dy), _f32(dy), _map(dymap), ptr(x), _f32(x), _map(xmap), ptr(mean), ptr(rstd),
         ptr(gamma), ptr(dx), _f32(dx), _map(dxmap), ptr(dx_add), ptr(dgamma), ptr(dbeta), C.c_int(rows),
         C.c_int(Cdim), ptr(dx_bf16), ptr(rowscale), C.c_int(rows_per_scale))


SR_ATTENTION_MAX_NK = 192


def sr_attention_fwd(q, kv, o, p_out, B, N, Nk, heads, scale):
    """o[B*N, C] = merge_heads(softmax(scale * q_h k_h^T) v_h) in one tcgen05 kernel (csrc/attn_tcgen05.cu); kv = [B*Nk, 2C]
    (K | V column halves). ``p_out`` ([B, heads, N, Nk] bf16) receives the probabilities when given (training)."""
    require_cuda(q, kv, o, p_out)
    C_ = heads * 64
    if q.dtype != BF16 or kv.dtype != BF16 or o.dtype != BF16 or (p_out is not None and p_out.dtype != BF16):
        raise _lib.MvltError("sr_attention_fwd: bf16 operands required")
    if (not q.is_contiguous() or not kv.is_contiguous() or not o.is_contiguous() or q.numel() != B * N * C_
            or kv.numel() != B * Nk * 2 * C_ or o.numel() != B * N * C_
            or (p_out is not None and (not p_out.is_contiguous() or p_out.numel() != B * heads * N * Nk))):
        raise _lib.MvltError("sr_attention_fwd: contiguous q [B*N, C], kv [B*Nk, 2C], o [B*N, C], p [B, h, N, Nk] required")
    if _lib.BYTES is not None:
        _lib.account_bytes("sr_attention_fwd", 2 * (q.numel() + o.numel() + kv.numel() + (p_out.numel() if p_out is not None else 0)))
    call("sr_attention_fwd", ptr(q), ptr(kv), ptr(o), ptr(p_out), C.c_int(B), C.c_int(N), C.c_int(Nk), C.c_int(heads),
         C.c_float(scale))


def sr_attention_bwd(q, kv, do, p, dq, dkv, B, N, Nk, heads, scale):
    """dq [B*N, C] and dkv [B*Nk, 2C] (dK | dV) of the fused attention from the saved probabilities, one tcgen05 kernel
    (csrc/attn_bwd_tcgen05.cu)."""
    require_cuda(q, kv, do, p, dq, dkv)
    C_ = heads * 64
    for t in (q, kv, do, p, dq, dkv):
        if t.dtype != BF16 or not t.is_contiguous():
            raise _lib.MvltError("sr_attention_bwd: contiguous bf16 operands required")
    if (q.numel() != B * N * C_ or do.numel() != B * N * C_ or dq.numel() != B * N * C_ or kv.numel() != B * Nk * 2 * C_
            or dkv.numel() != B * Nk * 2 * C_ or p.numel() != B * heads * N * Nk):
        raise _lib.MvltError("sr_attention_bwd: q/do/dq [B*N, C], kv/dkv [B*Nk, 2C], p [B, h, N, Nk] required")
    call("sr_attention_bwd", ptr(q), ptr(kv), ptr(do), ptr(p), ptr(dq), ptr(dkv), C.c_int(B), C.c_int(N), C.c_int(Nk),
         C.c_int(heads), C.c_float(scale))


MLP_FUSED_DIMS = (64, 128)      # embedding widths the fused MLP forward supports (PVLT stages 1 and 2): used at inference
# ... and the widths whose recompute backward is used in TRAINING. Both backward kernels are validated (tests/test_mlp_gpu.py,
# model parity with MVLT_FUSED_MLP_TRAIN=64,128), but at C = 128 the fused pair (172 + 347 us per block) is no faster than the
# two-GEMM path it replaces (15.32 vs 15.25 ms per step, profiles/r2t_mlp_train_dims_ab.txt), so the default is stage 1 only.
MLP_FUSED_BWD_DIMS = tuple(int(v) for v in __import__("os").environ.get("MVLT_FUSED_MLP_TRAIN", "64").split(",") if v)        # ... and the widths whose recompute backward exists (training uses the fused path only there)


def mlp_fwd(x, w1, b1, w2, b2, residual, out, rowscale=None, rows_per_scale=0, ln=None):
    """out = residual + rowscale[row // rows_per_scale] * (gelu(x @ w1^T + b1) @ w2^T + b2), the hidden activation staying
    on-chip (csrc/mlp_tcgen05.cu). x bf16 [M, C], w1 bf16 [HD, C], w2 bf16 [C, HD], residual / out fp32 [M, C]; C in {64, 128}.
    ``ln = (gamma, beta, out_bf16, mean, rstd, eps)``: LayerNorm of the output rows (the next block's norm1) from the same launch."""
    require_cuda(x, w1, w2, residual, out, b1, b2, rowscale)
    M, C_ = x.shape
    HD = w1.shape[0]
    if (x.dtype != BF16 or w1.dtype != BF16 or w2.dtype != BF16 or residual.dtype != F32 or out.dtype != F32
            or b1.dtype != F32 or b2.dtype != F32 or tuple(w1.shape) != (HD, C_) or tuple(w2.shape) != (C_, HD)
            or tuple(residual.shape) != (M, C_) or tuple(out.shape) != (M, C_) or b1.numel() != HD or b2.numel() != C_):
        raise _lib.MvltError("mlp_fwd: x bf16 [M, C], w1 bf16 [HD, C], w2 bf16 [C, HD], b1 / b2 fp32, residual / out fp32 [M, C] required")
    for t in (x, w1, w2, residual, out, b1, b2):
        if not t.is_contiguous():
            raise _lib.MvltError("mlp_fwd: contiguous operands required")
    g_ = b_ = o_ = m_ = r_ = None
    eps = 0.0
    if ln is not None:
        g_, b_, o_, m_, r_, eps = ln
        require_cuda(g_, b_, o_, m_, r_)
        if (g_.dtype != F32 or b_.dtype != F32 or g_.numel() != C_ or b_.numel() != C_ or o_.dtype != BF16 or tuple(o_.shape) != (M, C_)
                or not o_.is_contiguous() or any(t is not None and (t.dtype != F32 or t.numel() != M or not t.is_contiguous()) for t in (m_, r_))):
            raise _lib.MvltError("mlp_fwd ln=(gamma, beta, out_bf16, mean, rstd, eps): fp32 gamma / beta [C], contiguous bf16 out [M, C], fp32 mean / rstd [M]")
    if _lib.BYTES is not None:
        _lib.account_bytes("mlp_fwd", M * C_ * (2 + 4 + 4 + (2 if ln is not None else 0)) + 4 * HD * C_)
    call("mlp_fwd", ptr(x), ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(residual), ptr(out), ptr(rowscale), C.c_int(rows_per_scale),
         C.c_int(M), C.c_int(C_), C.c_int(HD), ptr(g_), ptr(b_), ptr(o_), ptr(m_), ptr(r_), C.c_float(float(eps)))


def mlp_bwd(x, dy, w1, b1, w2, dh, dw1, dw2, db1):
    """Backward of the fused MLP branch, C in {64, 128} (csrc/mlp_tcgen05.cu): recomputes fc1 from ``x``; writes dh [M, HD] bf16 (=
    (dy w2) * gelu'(x w1^T + b1)) and ACCUMULATES dw1 [HD, C], dw2 [C, HD], db1 [HD] (fp32). dX = dh @ w1 is a GEMM of its own."""
    require_cuda(x, dy, w1, w2, dh, dw1, dw2, db1, b1)
    M, C_ = x.shape
    HD = w1.shape[0]
    if (x.dtype != BF16 or dy.dtype != BF16 or w1.dtype != BF16 or w2.dtype != BF16 or dh.dtype != BF16 or dw1.dtype != F32
            or dw2.dtype != F32 or db1.dtype != F32 or b1.dtype != F32 or tuple(dy.shape) != (M, C_) or tuple(w1.shape) != (HD, C_)
            or tuple(w2.shape) != (C_, HD) or tuple(dh.shape) != (M, HD) or tuple(dw1.shape) != (HD, C_) or tuple(dw2.shape) != (C_, HD)
            or db1.numel() != HD):
        raise _lib.MvltError("mlp_bwd: x / dy bf16 [M, C], w1 bf16 [HD, C], w2 bf16 [C, HD], dh bf16 [M, HD], dw1 / dw2 / db1 fp32 required")
    for t in (x, dy, w1, w2, dh, dw1, dw2, db1, b1):
        if not t.is_contiguous():
            raise _lib.MvltError("mlp_bwd: contiguous operands required")
    if _lib.BYTES is not None:
        _lib.account_bytes("mlp_bwd", M * C_ * 4 + M * HD * 2 + 12 * HD * C_)
    call("mlp_bwd", ptr(x), ptr(dy), ptr(w1), ptr(b1), ptr(w2), ptr(dh), ptr(dw1), ptr(dw2), ptr(db1), C.c_int(M), C.c_int(C_), C.c_int(HD))


def softmax_fwd(s, rows, nk):
    call("softmax_fwd", ptr(s), C.c_longlong(rows), C.c_int(nk))


def softmax_bwd(p, dp, rows, nk, scale):
    call("softmax_bwd", ptr(p), ptr(dp), C.c_longlong(rows), C.c_int(nk), C.c_float(scale))


def cast_scale_bf16(src, dst, rows, Cdim, rowscale=None, rows_per_scale=0, alpha=1.0):
    call("cast_scale_bf16", ptr(src), ptr(dst), C.c_longlong(rows), C.c_int(Cdim), ptr(rowscale),
         C.c_int(rows_per_scale), C.c_float(alpha))


def colsum(x, rows, Cdim, ld, out):
    call("colsum", ptr(x), _f32(x), C.c_longlong(rows), C.c_int(Cdim), C.c_longlong(ld), ptr(out))


def patchify(src, src_batch_stride, dst, B, H, W, Cdim, R):
    call("patchify", ptr(src), _f32(src), C.c_longlong(src_batch_stride), ptr(dst), C.c_int(B), C.c_int(H),
         C.c_int(W), C.c_int(Cdim), C.c_int(R))


def unpatchify(src, dst, dst_batch_stride, B, H, W, Cdim, R, accumulate=False):
    call("unpatchify", ptr(src), ptr(dst), C.c_longlong(dst_batch_stride), C.c_int(B), C.c_int(H), C.c_int(W),
         C.c_int(Cdim), C.c_int(R), C.c_int(1 if accumulate else 0))


def patchify_nchw(img, dst, B, Cin, H, W, P, Kpad):
    call("patchify_nchw", ptr(img), ptr(dst), C.c_int(B), C.c_int(Cin), C.c_int(H), C.c_int(W), C.c_int(P),
         C.c_int(Kpad))


def copy_rows(src, dst, rows, Cdim, smap=None, dmap=None, lds=None, ldd=None, accumulate=False):
    call("copy_rows", ptr(src), _f32(src), _map(smap), C.c_longlong(lds or Cdim), ptr(dst), _f32(dst), _map(dmap),
         C.c_longlong(ldd or Cdim), C.c_longlong(rows), C.c_int(Cdim), C.c_int(1 if accumulate else 0))


def batch_reduce(x, batch_stride, B, n, out, accumulate=False):
    call("batch_reduce", ptr(x), C.c_longlong(batch_stride), C.c_int(B), C.c_longlong(n), ptr(out),
         C.c_int(1 if accumulate else 0))


def pos_resize_fwd(table, out, h, w, H, W, Cdim):
    call("pos_resize_fwd", ptr(table), ptr(out), C.c_int(h), C.c_int(w), C.c_int(H), C.c_int(W), C.c_int(Cdim))


def pos_resize_bwd(dout, dtable, h, w, H, W, Cdim):
    call("pos_resize_bwd", ptr(dout), ptr(dtable), C.c_int(h), C.c_int(w), C.c_int(H), C.c_int(W), C.c_int(Cdim))


def cast_weight(src, dst):
    call("cast_weight", ptr(src), ptr(dst), C.c_longlong(src.numel()))


def cast_conv_weight(src, dst, Co, Ci, KK, dst_ld):
    call("cast_conv_weight", ptr(src), ptr(dst), C.c_int(Co), C.c_int(Ci), C.c_int(KK), C.c_int(dst_ld))


def cast_conv_weight_t(src, dst, Co, Ci, KK):
    call("cast_conv_weight_t", ptr(src), ptr(dst), C.c_int(Co), C.c_int(Ci), C.c_int(KK))


def uncast_conv_wgrad(dwp, dw, Co, Ci, KK, src_ld):
    call("uncast_conv_wgrad", ptr(dwp), ptr(dw), C.c_int(Co), C.c_int(Ci), C.c_int(KK), C.c_int(src_ld))


def uncast_conv_wgrad_multi(items):
    """items: [(dwp fp32 [Co, KK*Ci], dw fp32 [Co, Ci, kh, kw]), ...]: dw[co,ci,kk] += dwp[co,kk,ci] for every pair, one
    launch per 24 tensors."""
    import struct
    for i in range(0, len(items), 24):
        chunk = items[i:i + 24]
        buf = bytearray()
        for dwp, dw in chunk:
            Co, Ci = dw.shape[0], dw.shape[1]
            KK = dw.shape[2] * dw.shape[3]
            if dwp.dtype != F32 or dw.dtype != F32 or dwp.shape != (Co, KK * Ci) or not dw.is_contiguous() or dwp.stride(1) != 1:
                raise _lib.MvltError("uncast_conv_wgrad_multi: bad gradient buffers")
            buf += struct.pack("<QQiiii", dwp.data_ptr(), dw.data_ptr(), Co, Ci, KK, dwp.stride(0))
        call("uncast_conv_wgrad_multi", (C.c_char * len(buf)).from_buffer(buf), C.c_int(len(chunk)))


def bert_embed_fwd(ids, word, pos, typ, gamma, beta, out, mean, rstd, rows, T, eps, p_drop, seed, seed_dev=None):
    """``seed_dev`` (all three hash-driven entry points): 1-element uint64-sized device tensor the kernel reads the seed from at
    execution time instead of ``seed`` -- what lets a captured CUDA graph draw fresh masks on every replay."""
    call("bert_embed_fwd", ptr(ids), ptr(word), ptr(pos), ptr(typ), ptr(gamma), ptr(beta), ptr(out), ptr(mean),
         ptr(rstd), C.c_int(rows), C.c_int(T), C.c_float(eps), C.c_float(p_drop), C.c_ulonglong(seed), ptr(seed_dev))


def bert_embed_bwd(dy, ids, word, pos, typ, gamma, mean, rstd, dword, dpos, dtype_, dgamma, dbeta, rows, T, p_drop,
                   seed, pad_id=0, seed_dev=None):
    call("bert_embed_bwd", ptr(dy), ptr(ids), ptr(word), ptr(pos), ptr(typ), ptr(gamma), ptr(mean), ptr(rstd),
         ptr(dword), ptr(dpos), ptr(dtype_), ptr(dgamma), ptr(dbeta), C.c_int(rows), C.c_int(T), C.c_float(p_drop),
         C.c_ulonglong(seed), C.c_int(pad_id), ptr(seed_dev))


def keep_scale(out, rows, cols, rate_per_row=None, rate=0.0, seed=0, seed_dev=None):
    """DropPath keep factors [rows, cols] (per-row rates on the device) or, with ``rate_per_row=None``, the element-wise
    dropout factors ``bert_embed_fwd`` applies for (seed, rate) (csrc/embed.cu)."""
    call("keep_scale", ptr(out), C.c_int(rows), C.c_int(cols), ptr(rate_per_row), C.c_float(rate), C.c_ulonglong(seed),
         ptr(seed_dev))


def compact_labels(labels, n, ignore, idx_out, labels_out, count_out, cap=0, count_f32=None, inv_count=None, overflow=None):
    """``cap`` > 0: fixed-capacity mode (static shapes for CUDA-graph capture): exactly ``cap`` entries of idx_out / labels_out
    are written, the tail beyond the count padded with (row 0, ignore label) -- rows that contribute no loss and no gradient;
    ``count_f32`` / ``inv_count`` (device floats) receive the count and 1 / max(count, 1), ``overflow`` the count when it
    exceeds ``cap`` (untouched otherwise: zero it once and check it when convenient)."""
    call("compact_labels", ptr(labels), C.c_int(n), C.c_longlong(ignore), ptr(idx_out), ptr(labels_out),
         ptr(count_out), C.c_int(cap), ptr(count_f32), ptr(inv_count), ptr(overflow))


def gather_rows(src, idx, n_idx, dst, Cdim, smap=None, lds=None):
    call("gather_rows", ptr(src), _map(smap), C.c_longlong(lds or Cdim), ptr(idx), C.c_int(n_idx), ptr(dst),
         _f32(dst), C.c_int(Cdim))


def scatter_rows(src, idx, n_idx, dst, Cdim, dmap=None, ldd=None, accumulate=False):
    call("scatter_rows", ptr(src), _f32(src), ptr(idx), C.c_int(n_idx), ptr(dst), _map(dmap),
         C.c_longlong(ldd or Cdim), C.c_int(Cdim), C.c_int(1 if accumulate else 0))


def ce_fwd(logits, ld, labels, rows, n_cls, ignore, lse, loss_sum, scale, total_sum=None, argmax_out=None,
           correct=None, scale_dev=None):
    call("ce_fwd", ptr(logits), _f32(logits), C.c_longlong(ld), ptr(labels), C.c_int(rows), C.c_int(n_cls),
         C.c_longlong(ignore), ptr(lse), ptr(loss_sum), ptr(total_sum), C.c_float(scale), ptr(argmax_out),
         ptr(correct), ptr(scale_dev))


def ce_bwd(logits, ld, labels, rows, n_cls, ignore, lse, dlogits, ldd, scale, gscale=None, scale_dev=None):
    call("ce_bwd", ptr(logits), _f32(logits), C.c_longlong(ld), ptr(labels), C.c_int(rows), C.c_int(n_cls),
         C.c_longlong(ignore), ptr(lse), ptr(dlogits), C.c_longlong(ldd), C.c_float(scale), ptr(gscale), ptr(scale_dev))


def small_linear_fwd(h, W, b1, b2, out, M, n, K):
    call("small_linear_fwd", ptr(h), ptr(W), ptr(b1), ptr(b2), ptr(out), C.c_int(M), C.c_int(n), C.c_int(K))


def small_linear_bwd(dlogits, h, W, dh, dW, db1, db2, M, n, K):
    call("small_linear_bwd", ptr(dlogits), ptr(h), ptr(W), ptr(dh), ptr(dW), ptr(db1), ptr(db2), C.c_int(M),
         C.c_int(n), C.c_int(K))


def itm_rank(logits, n_query, n_cand, rank_out, prob_out=None):
    call("itm_rank", ptr(logits), C.c_int(n_query), C.c_int(n_cand), ptr(rank_out), ptr(prob_out))


def grid_mask(seeds_u32, grid_out, B, size_w, size_h, patch, ratio):
    call("grid_mask", ptr(seeds_u32), ptr(grid_out), C.c_int(B), C.c_int(size_w), C.c_int(size_h), C.c_int(patch),
         C.c_double(ratio))


def token_mask(seeds_u32, ori_ids, input_ids, labels, B, T, sep_id, mask_id, vocab_size, rate):
    call("token_mask", ptr(seeds_u32), ptr(ori_ids), ptr(input_ids), ptr(labels), C.c_int(B), C.c_int(T), C.c_int(sep_id),
         C.c_int(mask_id), C.c_int(vocab_size), C.c_double(rate))


def masked_fill(img, grid, out, mask_out, B, Cc, H, W, patch, fill=1e-6):
    call("masked_fill", ptr(img), ptr(grid), ptr(out), ptr(mask_out), C.c_int(B), C.c_int(Cc), C.c_int(H),
         C.c_int(W), C.c_int(patch), C.c_float(fill))


def gelu_bwd(dy, pre, out, n):
    call("gelu_bwd", ptr(dy), ptr(pre), ptr(out), C.c_longlong(n))


def gelu_ew(x, out, dy=None):
    """out = gelu_erf(x), or dy * gelu_erf'(x) when ``dy`` is given; contiguous fp32 or bf16 arrays of one dtype."""
    require_cuda(x, out, dy)
    for t in (x, out, dy):
        if t is not None and (t.dtype != x.dtype or not t.is_contiguous() or t.numel() != x.numel()):
            raise _lib.MvltError("gelu_ew: contiguous arrays of one dtype and size required")
    call("gelu_ew", ptr(x), ptr(dy), ptr(out), C.c_longlong(x.numel()), C.c_int(_f32(x)))


def cast2d(src, lds, dst, ldd, rows, Cdim, alpha=1.0):
    call("cast2d", ptr(src), _f32(src), C.c_longlong(lds), ptr(dst), _f32(dst), C.c_longlong(ldd),
         C.c_longlong(rows), C.c_int(Cdim), C.c_float(alpha))


def set_values(dst, payload: bytes):
    """Writes ``payload`` (4..64 bytes, a multiple of 4) to device memory at ``dst`` with a one-thread launch that carries the
    bytes in its kernel arguments (csrc/elementwise.cu: no pinned staging; safe to call again immediately)."""
    buf = (C.c_char * len(payload)).from_buffer_copy(payload)
    call("set_values", ptr(dst), buf, C.c_int(len(payload)))


def spin(ms: float):
    """Occupy the current stream for ``ms`` milliseconds (measurement aid, see csrc/elementwise.cu)."""
    call("spin", C.c_longlong(int(ms * 1e6)))


def memset_zero(t):
    call("memset_zero", ptr(t), C.c_longlong(t.numel() * t.element_size()))


def zeros(shape, dtype, dev):
    t = torch.empty(shape, dtype=dtype, device=dev)
    call("memset_zero", ptr(t), C.c_longlong(t.numel() * t.element_size()))
    return t


# ------------------------------------------------------------------------------------------------------
# t2i (MVM) head kernels
# ------------------------------------------------------------------------------------------------------
def im2col3x3(src, batch_stride, pix_stride, col, B, H, W, Cdim):
    call("im2col3x3", ptr(src), _f32(src), C.c_longlong(batch_stride), C.c_int(pix_stride), ptr(col), C.c_int(B),
         C.c_int(H), C.c_int(W), C.c_int(Cdim))


def col2im3x3(dcol, dst, batch_stride, pix_stride, B, H, W, Cdim, accumulate=False):
    call("col2im3x3", ptr(dcol), ptr(dst), _f32(dst), C.c_longlong(batch_stride), C.c_int(pix_stride), C.c_int(B),
         C.c_int(H), C.c_int(W), C.c_int(Cdim), C.c_int(1 if accumulate else 0))


def bn_stats(x, rows, Cdim, s, ss):
    call("bn_stats", ptr(x), C.c_longlong(rows), C.c_int(Cdim), ptr(s), ptr(ss))


def bn_finalize(s, ss, rows, gamma, beta, rmean, rvar, momentum, eps, training, scale, shift, mean, invstd, Cdim):
    call("bn_finalize", ptr(s), ptr(ss), C.c_longlong(rows), ptr(gamma), ptr(beta), ptr(rmean), ptr(rvar),
         C.c_float(momentum), C.c_float(eps), C.c_int(1 if training else 0), ptr(scale), ptr(shift), ptr(mean),
         ptr(invstd), C.c_int(Cdim))


def bn_apply(x, scale, shift, out, out_ld, out_coff, rows, Cdim, m1=None, m1_ld=0, m2=None, m2_ld=0):
    call("bn_apply", ptr(x), ptr(scale), ptr(shift), ptr(m1), C.c_int(m1_ld), ptr(m2), C.c_int(m2_ld), ptr(out),
         C.c_int(out_ld), C.c_int(out_coff), C.c_longlong(rows), C.c_int(Cdim))


def bn_bwd(dy, x, scale, mean, invstd, sum_dy, sum_dy_xhat, dx, rows, Cdim, training):
    call("bn_bwd", ptr(dy), ptr(x), ptr(scale), ptr(mean), ptr(invstd), ptr(sum_dy), ptr(sum_dy_xhat), ptr(dx),
         C.c_longlong(rows), C.c_int(Cdim), C.c_int(1 if training else 0))


def ew_mul(a, a_ld, a_coff, dst, d_ld, d_coff, rows, Cdim, b=None, b_ld=0, c2=None, c2_ld=0, accumulate=False):
    call("ew_mul", ptr(a), _f32(a), C.c_int(a_ld), C.c_int(a_coff), ptr(b), C.c_int(b_ld), ptr(c2), C.c_int(c2_ld),
         ptr(dst), _f32(dst), C.c_int(d_ld), C.c_int(d_coff), C.c_longlong(rows), C.c_int(Cdim),
         C.c_int(1 if accumulate else 0))


def upsample2x_fwd(src, batch_stride, pix_stride, dst, B, h, w, Cdim):
    call("upsample2x_fwd", ptr(src), _f32(src), C.c_longlong(batch_stride), C.c_int(pix_stride), ptr(dst), C.c_int(B),
         C.c_int(h), C.c_int(w), C.c_int(Cdim))


def upsample2x_bwd(dy, dx, batch_stride, pix_stride, B, h, w, Cdim, accumulate=False):
    call("upsample2x_bwd", ptr(dy), ptr(dx), _f32(dx), C.c_longlong(batch_stride), C.c_int(pix_stride), C.c_int(B),
         C.c_int(h), C.c_int(w), C.c_int(Cdim), C.c_int(1 if accumulate else 0))


def score_fwd(x, W, bias, out, rows, Cin):
    call("score_fwd", ptr(x), ptr(W), ptr(bias), ptr(out), C.c_longlong(rows), C.c_int(Cin))


def score_bwd(dscore, x, W, dx, dW, db, rows, Cin, gscale=None):
    call("score_bwd", ptr(dscore), ptr(x), ptr(W), ptr(dx), ptr(dW), ptr(db), C.c_longlong(rows), C.c_int(Cin), ptr(gscale))


def upsample8_fwd(score, out, B, h, w, S):
    call("upsample8_fwd", ptr(score), ptr(out), C.c_int(B), C.c_int(h), C.c_int(w), C.c_int(S))


def t2i_up_loss(score, target, dpred, dscore, loss_sum, total_sum, loss_scale, grad_scale, gscale, B, h, w, S, mode,
                want_grad):
    call("t2i_up_loss", ptr(score), ptr(target), ptr(dpred), ptr(dscore), ptr(loss_sum), ptr(total_sum),
         C.c_float(loss_scale), C.c_float(grad_scale), ptr(gscale), C.c_int(B), C.c_int(h), C.c_int(w), C.c_int(S),
         C.c_int(mode), C.c_int(1 if want_grad else 0))


def sq_diff_sum(a, b, n, out):
    call("sq_diff_sum", ptr(a), ptr(b), C.c_longlong(n), ptr(out))
