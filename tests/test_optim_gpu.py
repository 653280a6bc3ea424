"""Multi-tensor AdamW kernel (csrc/optim.cu) vs torch.optim.AdamW on identical parameters / gradients."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_adamw_matches_torch_over_steps():
    from mvlt_b200.optim import AdamW
    g = torch.Generator(device="cuda").manual_seed(3)
    shapes = [(30522, 768), (512,), (320, 64, 2, 2), (3,), (1, 50, 64), (7, 13), (64,), (1030,)]
    ours = [torch.nn.Parameter(torch.randn(s, generator=g, device="cuda")) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]

    def groups(ps):
        return [{"params": [p for p in ps if p.ndim > 1], "weight_decay": 0.05}, {"params": [p for p in ps if p.ndim <= 1], "weight_decay": 0.0}]
    o1 = AdamW(groups(ours), lr=3e-3, betas=(0.9, 0.999), eps=1e-8)
    o2 = torch.optim.AdamW(groups(ref), lr=3e-3, betas=(0.9, 0.999), eps=1e-8)
    for step in range(4):
        for a, b in zip(ours, ref):
            gr = torch.randn(a.shape, generator=g, device="cuda") * (0.1 + step)
            a.grad = gr.clone()
            b.grad = gr.clone()
        if step == 2:   # a scheduler rewrites the group's lr between steps
            for grp in o1.param_groups + o2.param_groups:
                grp["lr"] = 1e-3
        o1.step()
        o2.step()
    for a, b in zip(ours, ref):
        err = (a - b).abs().max().item()
        assert err <= 2e-6 * max(1.0, b.abs().max().item()), (tuple(a.shape), err)
    sd = o1.state_dict()
    assert sd["state"][0]["step"] == 4 and sd["state"][0]["exp_avg"].shape == ours[0].shape


def test_adamw_grad_scale_and_no_cpu_fallback():
    from mvlt_b200.optim import AdamW
    from mvlt_b200._lib import MvltError
    p = torch.nn.Parameter(torch.ones(1000, device="cuda"))
    q = torch.nn.Parameter(torch.ones(1000, device="cuda"))
    p.grad = torch.full((1000,), 8.0, device="cuda")
    q.grad = torch.full((1000,), 2.0, device="cuda")
    AdamW([p], lr=1e-2, weight_decay=0.0).step(grad_scale=torch.tensor([0.25], device="cuda"))
    AdamW([q], lr=1e-2, weight_decay=0.0).step()
    assert torch.allclose(p, q, atol=1e-7)
    c = torch.nn.Parameter(torch.ones(4))
    c.grad = torch.ones(4)
    with pytest.raises(MvltError):
        AdamW([c]).step()


def test_adamw_refreshes_registered_bf16_shadows():
    """The engine's GEMMs read bf16 copies of the fp32 masters; the own optimizer writes the masters through raw pointers
    (no autograd version bump), so the SAME kernel must rewrite the copies: plain, conv-permuted [Co, tap, Ci] and the
    flipped + transposed [Ci, tap', Co] layouts (csrc/optim.cu shadow modes 1 / 2)."""
    from mvlt_b200.optim import AdamW
    g = torch.Generator(device="cuda").manual_seed(5)
    lin = torch.nn.Parameter(torch.randn((1031, 67), generator=g, device="cuda"))
    conv = torch.nn.Parameter(torch.randn((24, 40, 2, 2), generator=g, device="cuda"))
    c3 = torch.nn.Parameter(torch.randn((64, 128, 3, 3), generator=g, device="cuda"))
    plain = torch.nn.Parameter(torch.randn((33, 5), generator=g, device="cuda"))
    w_lin = torch.zeros((1031, 67), device="cuda", dtype=torch.bfloat16)
    w_conv = torch.zeros((24, 4 * 40), device="cuda", dtype=torch.bfloat16)
    w_c3 = torch.zeros((64, 9 * 128), device="cuda", dtype=torch.bfloat16)
    w_c3t = torch.zeros((128, 9 * 64), device="cuda", dtype=torch.bfloat16)
    lin._mvlt_shadow = (1, w_lin, None, 0, 0, 0, 0)
    conv._mvlt_shadow = (2, w_conv, None, 24, 40, 4, 160)
    c3._mvlt_shadow = (2, w_c3, w_c3t, 64, 128, 9, 9 * 128)
    ps = [lin, conv, c3, plain]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    o1, o2 = AdamW(ps, lr=1e-2, weight_decay=0.05), torch.optim.AdamW(ref, lr=1e-2, weight_decay=0.05)
    from mvlt_b200 import _lib
    w_epoch = _lib.WEIGHT_EPOCH
    for step in range(3):
        for a, b in zip(ps, ref):
            a.grad = torch.randn(a.shape, generator=g, device="cuda")
            b.grad = a.grad.clone()
        o1.step()
        o2.step()
    assert _lib.WEIGHT_EPOCH == w_epoch + 3          # `plain` has no registered copy: engines must recast
    for a, b in zip(ps, ref):
        assert (a - b).abs().max().item() <= 2e-6 * max(1.0, b.abs().max().item())
    assert torch.equal(w_lin, lin.detach().to(torch.bfloat16))
    assert torch.equal(w_conv.view(24, 4, 40), conv.detach().view(24, 40, 4).permute(0, 2, 1).to(torch.bfloat16))
    assert torch.equal(w_c3.view(64, 9, 128), c3.detach().view(64, 128, 9).permute(0, 2, 1).to(torch.bfloat16))
    # input-gradient layout: w16t[ci, t, co] = w[co, ci, 8 - t]
    assert torch.equal(w_c3t.view(128, 9, 64), c3.detach().view(64, 128, 9).flip(-1).permute(1, 2, 0).to(torch.bfloat16))


@pytest.mark.parametrize("max_norm", [0.37, 1e6])
def test_adamw_fused_gradient_clipping_matches_torch_clip_grad_norm(max_norm):
    """AdamW.step(max_norm=...) (per-chunk squared sums + one coefficient launch, multiplied into the gradients as the AdamW
    kernel reads them) vs torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW. eps is large on purpose: Adam's update is
    invariant to a uniform gradient scale when eps -> 0, so a wrong coefficient would otherwise go unnoticed."""
    from mvlt_b200.optim import AdamW
    g = torch.Generator(device="cuda").manual_seed(9)
    shapes = [(3000, 768), (512,), (320, 64, 2, 2), (3,), (1, 50, 64), (7, 13), (16385,), (1030,)]
    ours = [torch.nn.Parameter(torch.randn(s, generator=g, device="cuda")) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]

    def groups(ps):
        return [{"params": [p for p in ps if p.ndim > 1], "weight_decay": 0.05}, {"params": [p for p in ps if p.ndim <= 1], "weight_decay": 0.0}]
    o1 = AdamW(groups(ours), lr=3e-3, eps=1e-2)
    o2 = torch.optim.AdamW(groups(ref), lr=3e-3, eps=1e-2)
    inv_scale = torch.tensor([0.5], device="cuda")
    for step in range(3):
        for a, b in zip(ours, ref):
            gr = torch.randn(a.shape, generator=g, device="cuda") * (1e-4 * 10 ** step)
            a.grad = gr.clone()
            b.grad = gr.clone()
        before = [a.grad.clone() for a in ours]
        if step == 2:       # with an incoming gradient scale (1 / loss scale): the norm is that of the unscaled gradients
            for b in ref:
                b.grad.mul_(0.5)
        want = torch.nn.utils.clip_grad_norm_(ref, max_norm)
        o1.step(max_norm=max_norm, grad_scale=inv_scale if step == 2 else None)
        o2.step()
        got = float(o1.last_grad_norm)
        assert abs(got - float(want)) <= 1e-5 * float(want), (step, got, float(want))
        coef = min(1.0, max_norm / (float(want) + 1e-6)) * (0.5 if step == 2 else 1.0)
        assert abs(float(o1._clip_out[0]) - coef) <= 1e-5 * coef, (step, float(o1._clip_out[0]), coef)
        for a, g0 in zip(ours, before):      # the gradients themselves are left as they were (the coefficient is applied on read)
            assert torch.equal(a.grad, g0)
    for a, b in zip(ours, ref):
        err = (a - b).abs().max().item()
        assert err <= 4e-6 * max(1.0, b.abs().max().item()), (tuple(a.shape), err)


def test_adamw_fused_clipping_is_deterministic():
    """No atomics in the norm: the same gradients give bit-identical coefficients (data-parallel ranks must stay bit-identical)."""
    from mvlt_b200.optim import AdamW
    g = torch.Generator(device="cuda").manual_seed(1)
    p = torch.nn.Parameter(torch.randn((4_000_037,), generator=g, device="cuda"))
    q = torch.nn.Parameter(torch.randn((129, 513), generator=g, device="cuda"))
    gp, gq = torch.randn(p.shape, generator=g, device="cuda"), torch.randn(q.shape, generator=g, device="cuda")
    outs = []
    for _ in range(3):
        o = AdamW([p, q], lr=0.0, weight_decay=0.0)
        p.grad, q.grad = gp, gq
        o.step(max_norm=1.0)
        outs.append(o._clip_out.clone())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    want = float(torch.sqrt(gp.double().pow(2).sum() + gq.double().pow(2).sum()))
    assert abs(float(outs[0][1]) - want) <= 1e-5 * want
