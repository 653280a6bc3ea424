"""Fused SR-attention forward (csrc/attn_tcgen05.cu) through the C-ABI: against an fp32 torch restatement of
/root/reference/libs/pvlt.py:113-117 on the same bf16 inputs, and against the two-GEMM tcgen05 path it replaces."""
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(180)]
BF16 = torch.bfloat16


def _inputs(B, N, heads, Nk, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    C = heads * 64
    q = torch.randn((B * N, C), generator=g, device="cuda").to(BF16)
    kv = torch.randn((B * Nk, 2 * C), generator=g, device="cuda").to(BF16)
    return q, kv


def _reference(q, kv, B, N, heads, Nk):
    C = heads * 64
    q4 = q.float().view(B, N, heads, 64).permute(0, 2, 1, 3)
    kv5 = kv.float().view(B, Nk, 2, heads, 64)
    k4, v4 = kv5[:, :, 0].permute(0, 2, 1, 3), kv5[:, :, 1].permute(0, 2, 1, 3)
    P = torch.softmax((q4 @ k4.transpose(-1, -2)) * 64 ** -0.5, dim=-1)          # pvlt.py:113-114
    O = (P @ v4).transpose(1, 2).reshape(B * N, C)                               # pvlt.py:117
    return P, O


SHAPES = [(2, 4224, 1, 192), (3, 192, 8, 192), (2, 384, 5, 160), (1, 1152, 2, 64), (2, 200, 2, 96)]


@pytest.mark.parametrize("B,N,heads,Nk", SHAPES)
@pytest.mark.parametrize("store_p", [True, False])
def test_fused_attention_matches_fp32_reference(B, N, heads, Nk, store_p):
    from mvlt_b200 import kernels as k
    q, kv = _inputs(B, N, heads, Nk, seed=N + Nk)
    C = heads * 64
    o = torch.full((B * N, C), float("nan"), device="cuda", dtype=BF16)
    P = torch.full((B, heads, N, Nk), float("nan"), device="cuda", dtype=BF16) if store_p else None
    k.sr_attention_fwd(q, kv, o, P, B, N, Nk, heads, 64 ** -0.5)
    torch.cuda.synchronize()
    Pr, Or = _reference(q, kv, B, N, heads, Nk)
    assert torch.isfinite(o.float()).all()
    # bf16 probabilities (relative 2^-8) feed the PV product: |dO| <= 2^-8 * sum_k P |v| plus the bf16 rounding of O
    assert (o.float() - Or).abs().max().item() <= 2e-2, (o.float() - Or).abs().max().item()
    if store_p:
        assert torch.isfinite(P.float()).all()
        err = (P.float() - Pr).abs()
        assert bool((err <= 1e-5 + 8e-3 * Pr).all()), err.max().item()
        assert (P.float().sum(-1) - 1).abs().max().item() <= 1e-2


@pytest.mark.parametrize("B,N,heads,Nk", SHAPES[:3])
def test_fused_attention_equals_two_gemm_path(B, N, heads, Nk):
    from mvlt_b200 import kernels as k
    q, kv = _inputs(B, N, heads, Nk, seed=7)
    C = heads * 64
    o1, o2 = torch.empty((B * N, C), device="cuda", dtype=BF16), torch.empty((B * N, C), device="cuda", dtype=BF16)
    P1, P2 = torch.empty((B, heads, N, Nk), device="cuda", dtype=BF16), torch.empty((B, heads, N, Nk), device="cuda", dtype=BF16)
    k.sr_attention_fwd(q, kv, o1, P1, B, N, Nk, heads, 64 ** -0.5)
    q4 = q.view(B, N, heads, 64).permute(0, 2, 1, 3)
    kv5 = kv.view(B, Nk, 2, heads, 64)
    k4, v4 = kv5[:, :, 0].permute(0, 2, 1, 3), kv5[:, :, 1].permute(0, 2, 1, 3)
    k.gemm(q4, k4, P2, alpha=64 ** -0.5, act=k.ACT_SOFTMAX)
    k.gemm(P2, v4.transpose(-1, -2), o2.view(B, N, heads, 64).permute(0, 2, 1, 3))
    torch.cuda.synchronize()
    # same formulas; only the order of the row-sum reduction differs (one warp per row vs two column halves): <= 1 bf16 ulp
    assert bool(((P1.float() - P2.float()).abs() <= 2 ** -7 * P2.float().abs() + 1e-7).all())
    assert (o1.float() - o2.float()).abs().max().item() <= 1.6e-2


def test_fused_attention_rejects_unsupported_lengths():
    from mvlt_b200 import kernels as k
    from mvlt_b200._lib import MvltError
    q, kv = _inputs(1, 128, 1, 224, seed=0)
    o = torch.empty_like(q)
    with pytest.raises(MvltError):
        k.sr_attention_fwd(q, kv, o, None, 1, 128, 224, 1, 0.125)


@pytest.mark.parametrize("B,N,heads,Nk", SHAPES)
def test_fused_attention_backward_matches_fp32_autograd(B, N, heads, Nk):
    from mvlt_b200 import kernels as k
    q, kv = _inputs(B, N, heads, Nk, seed=3 * N + Nk)
    C = heads * 64
    g = torch.Generator(device="cuda").manual_seed(11)
    do = torch.randn((B * N, C), generator=g, device="cuda").to(BF16)
    o = torch.empty((B * N, C), device="cuda", dtype=BF16)
    P = torch.empty((B, heads, N, Nk), device="cuda", dtype=BF16)
    k.sr_attention_fwd(q, kv, o, P, B, N, Nk, heads, 64 ** -0.5)
    dq = torch.full((B * N, C), float("nan"), device="cuda", dtype=BF16)
    dkv = torch.full((B * Nk, 2 * C), float("nan"), device="cuda", dtype=BF16)
    k.sr_attention_bwd(q, kv, do, P, dq, dkv, B, N, Nk, heads, 64 ** -0.5)
    torch.cuda.synchronize()
    qf = q.float().requires_grad_(True)
    kvf = kv.float().requires_grad_(True)
    _, Or = _reference(qf, kvf, B, N, heads, Nk)
    Or.backward(do.float())
    rel = lambda a, b: float((a.float() - b).norm() / (b.norm() + 1e-12))
    assert torch.isfinite(dq.float()).all() and torch.isfinite(dkv.float()).all()
    assert rel(dq, qf.grad) <= 2e-2, rel(dq, qf.grad)
    assert rel(dkv, kvf.grad) <= 2e-2, rel(dkv, kvf.grad)
