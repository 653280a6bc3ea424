"""Drop-in for the reference's ``libs/vl_scores.py`` (/root/reference/libs/vl_scores.py:5-63): same function
names, arguments and return types; the reductions run in the C-ABI kernels (csrc/loss.cu).
"""
from __future__ import annotations

import math

import torch

from .. import kernels as k
from .._lib import MvltError

F32, BF16 = torch.float32, torch.bfloat16


def compute_mlm_score(logits, target, index=-1):
    """argmax accuracy over positions whose target != index (vl_scores.py:5-34). Returns a Python float
    (NaN when no position is labelled, like the reference's 0/0)."""
    if not logits.is_cuda:
        raise MvltError("compute_mlm_score needs CUDA tensors (no CPU fallback)")
    logits, target = logits.detach(), target.detach().to(logits.device).contiguous().view(-1)
    V = logits.shape[-1]
    rows = target.numel()
    lg = logits.reshape(rows, V) if logits.dim() != 2 else logits
    if lg.stride(-1) != 1:
        lg = lg.contiguous()
    if lg.dtype not in (F32, BF16):
        lg = lg.to(F32)
    ld = lg.stride(0)
    dev = lg.device
    lse = torch.empty((rows,), dtype=F32, device=dev)
    dummy = k.zeros((1,), F32, dev)
    correct = k.zeros((1,), F32, dev)
    k.ce_fwd(lg, ld, target, rows, V, index, lse, dummy, 0.0, correct=correct)
    total = int((target != index).sum())
    return float(correct.item()) / total if total > 0 else float("nan")


def compute_score_with_logits(logits, labels):
    """argmax == label vector for multi-logit heads (vl_scores.py:37-51)."""
    if logits.shape[1] > 1:
        lg = logits.detach().to(F32).contiguous()
        rows, n = lg.shape[0], lg.shape[1]
        dev = lg.device
        am = torch.empty((rows,), dtype=torch.int32, device=dev)
        lse = torch.empty((rows,), dtype=F32, device=dev)
        dummy = k.zeros((1,), F32, dev)
        lab = labels.detach().to(dev).contiguous().view(-1)
        k.ce_fwd(lg, n, lab, rows, n, -100, lse, dummy, 0.0, argmax_out=am)
        return am.to(lab.dtype) == lab
    scores = torch.zeros_like(labels)
    p = torch.sigmoid(logits.view(-1))
    return ((p >= 0.5) == (labels.view(-1) == 1)).to(scores.dtype).view_as(scores)


def compute_psnr(logits, labels):
    """20*log10(255/sqrt(mse)), 100 when mse == 0 (vl_scores.py:54-63)."""
    logits, labels = logits.detach().to(F32).contiguous(), labels.detach().to(F32).contiguous()
    acc = k.zeros((1,), F32, logits.device)
    k.sq_diff_sum(logits, labels, logits.numel(), acc)
    mse = float(acc.item()) / logits.numel()
    if mse == 0:
        return 100
    return 20 * math.log10(255.0 / math.sqrt(mse))
