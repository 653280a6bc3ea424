"""Fused MLP kernels (csrc/mlp_tcgen05.cu) vs an fp32 torch restatement of /root/reference/libs/pvlt.py:65-71,142 on the
same bf16-rounded operands. Tolerance: the hidden activation is rounded to bf16 on-chip (as the two-GEMM path stores it):
relative L2 error of the MLP branch <= 1e-2."""
import pytest
import torch

pytestmark = pytest.mark.gpu
BF16, F32 = torch.bfloat16, torch.float32

SHAPES = [(128 * 5, 64, 512), (128 * 3 + 17, 64, 512), (4224 * 3, 64, 512), (128 * 4, 128, 1024), (1152 * 2 + 40, 128, 1024),
          (64, 64, 128)]


def _inputs(M, C, HD, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn((M, C), generator=g, device="cuda").to(BF16)
    w1 = (torch.randn((HD, C), generator=g, device="cuda") * C ** -0.5).to(BF16)
    w2 = (torch.randn((C, HD), generator=g, device="cuda") * HD ** -0.5).to(BF16)
    b1 = torch.randn((HD,), generator=g, device="cuda") * 0.3
    b2 = torch.randn((C,), generator=g, device="cuda") * 0.3
    res = torch.randn((M, C), generator=g, device="cuda")
    return x, w1, b1, w2, b2, res


@pytest.mark.parametrize("M,C,HD", [(128 * 3 + 17, 64, 512), (4224 * 3, 64, 512), (1152 * 2 + 40, 128, 1024)])
def test_fused_mlp_forward_with_layernorm_of_the_output(M, C, HD):
    """The next block's norm1 from the same launch (/root/reference/libs/pvlt.py:140-143): output unchanged, the normalised
    bf16 rows and (mean, rstd) equal LayerNorm of the kernel's own fp32 output."""
    from mvlt_b200 import kernels as k
    x, w1, b1, w2, b2, res = _inputs(M, C, HD, seed=M + C + 1)
    res = res * 2.0 + 0.5
    g = torch.Generator(device="cuda").manual_seed(9)
    gamma, beta = torch.randn(C, generator=g, device="cuda"), torch.randn(C, generator=g, device="cuda")
    rs = (torch.arange((M + 95) // 96, device="cuda") % 3 != 1).float() * 1.5
    out0 = torch.empty((M, C), device="cuda")
    k.mlp_fwd(x, w1, b1, w2, b2, res, out0, rowscale=rs, rows_per_scale=96)
    out = torch.full((M, C), float("nan"), device="cuda")
    xn = torch.full((M, C), float("nan"), device="cuda", dtype=BF16)
    mean, rstd = torch.full((M,), float("nan"), device="cuda"), torch.full((M,), float("nan"), device="cuda")
    k.mlp_fwd(x, w1, b1, w2, b2, res, out, rowscale=rs, rows_per_scale=96, ln=(gamma, beta, xn, mean, rstd, 1e-6))
    torch.cuda.synchronize()
    assert torch.equal(out, out0)
    mu, var = out.mean(1), out.var(1, unbiased=False)
    assert torch.allclose(mean, mu, rtol=1e-4, atol=1e-5) and torch.allclose(rstd, (var + 1e-6).rsqrt(), rtol=2e-4, atol=1e-5)
    ref = torch.nn.functional.layer_norm(out, (C,), gamma, beta, 1e-6)
    assert float((xn.float() - ref).abs().max()) <= 1e-2 * float(ref.abs().max()) + 1e-3
    with pytest.raises(Exception):      # in-place output is refused in this mode (the second pass re-reads the residual)
        k.mlp_fwd(x, w1, b1, w2, b2, res, res, ln=(gamma, beta, xn, mean, rstd, 1e-6))


@pytest.mark.parametrize("M,C,HD", SHAPES)
@pytest.mark.parametrize("droppath", [False, True])
def test_fused_mlp_forward_matches_fp32_reference(M, C, HD, droppath):
    from mvlt_b200 import kernels as k
    x, w1, b1, w2, b2, res = _inputs(M, C, HD, seed=M + C)
    rows_per_scale = 96
    rs = None
    if droppath:
        nb = (M + rows_per_scale - 1) // rows_per_scale
        rs = (torch.arange(nb, device="cuda") % 3 != 1).float() / (2.0 / 3.0)
    out = torch.full((M, C), float("nan"), device="cuda")
    k.mlp_fwd(x, w1, b1, w2, b2, res, out, rowscale=rs, rows_per_scale=rows_per_scale if droppath else 0)
    torch.cuda.synchronize()
    h = torch.nn.functional.gelu(x.float() @ w1.float().t() + b1)
    branch = h @ w2.float().t() + b2
    if droppath:
        branch = branch * rs[torch.arange(M, device="cuda") // rows_per_scale].view(-1, 1)
    ref = res + branch
    assert torch.isfinite(out).all()
    err = float((out - ref).norm() / branch.norm())
    assert err <= 1e-2, err
    # rows of dropped samples must carry the residual exactly
    if droppath:
        dropped = rs[torch.arange(M, device="cuda") // rows_per_scale] == 0
        assert torch.equal(out[dropped], res[dropped])
    # in-place form (out aliases the residual), as the engine uses it at inference
    res2 = res.clone()
    k.mlp_fwd(x, w1, b1, w2, b2, res2, res2, rowscale=rs, rows_per_scale=rows_per_scale if droppath else 0)
    torch.cuda.synchronize()
    assert torch.equal(res2, out)


@pytest.mark.parametrize("M,C,HD", [(128 * 5, 64, 512), (128 * 3 + 17, 64, 512), (4224 * 4, 64, 512), (64, 64, 128), (128 * 40 + 5, 64, 1024),
                                    (128 * 5, 128, 1024), (128 * 3 + 17, 128, 1024), (1152 * 8, 128, 1024), (64, 128, 128), (128 * 30 + 5, 128, 512)])
def test_fused_mlp_backward_matches_fp32_autograd(M, C, HD):
    """dh', dW1, dW2, db1 of the recompute backward vs fp32 autograd of the same branch (bf16-rounded operands); the
    accumulators must ADD to what the gradient buffers already hold."""
    from mvlt_b200 import kernels as k
    x, w1, b1, w2, b2, _ = _inputs(M, C, HD, seed=7 + M)
    g = torch.Generator(device="cuda").manual_seed(99)
    dy = torch.randn((M, C), generator=g, device="cuda").to(BF16)
    dh = torch.full((M, HD), float("nan"), device="cuda", dtype=BF16)
    dw1 = torch.full((HD, C), 0.5, device="cuda")
    dw2 = torch.full((C, HD), -0.25, device="cuda")
    db1 = torch.full((HD,), 2.0, device="cuda")
    k.mlp_bwd(x, dy, w1, b1, w2, dh, dw1, dw2, db1)
    torch.cuda.synchronize()
    xf = x.float()
    w1f, w2f, b1f = w1.float().requires_grad_(True), w2.float().requires_grad_(True), b1.clone().requires_grad_(True)
    h = xf @ w1f.t() + b1f
    h.retain_grad()
    y = torch.nn.functional.gelu(h) @ w2f.t()
    y.backward(dy.float())
    rel = lambda a, b: float((a.float() - b).norm() / (b.norm() + 1e-12))
    assert torch.isfinite(dh.float()).all()
    assert rel(dh, h.grad) <= 1e-2, rel(dh, h.grad)
    assert rel(dw1 - 0.5, w1f.grad) <= 1e-2, rel(dw1 - 0.5, w1f.grad)
    assert rel(dw2 + 0.25, w2f.grad) <= 1e-2, rel(dw2 + 0.25, w2f.grad)
    assert rel(db1 - 2.0, b1f.grad) <= 1e-2, rel(db1 - 2.0, b1f.grad)
