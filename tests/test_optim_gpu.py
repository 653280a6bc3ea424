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
