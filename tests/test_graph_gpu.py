"""CUDA-graph replay of the training iteration (mvlt_b200/graph.py) against the per-launch path: same kernels, same draws,
same losses and parameters; plus the device-side scalars the captured step depends on (fixed-capacity label compaction,
device-resident CE scale, AdamW hyper-parameters, seeds). Reference loop: /root/reference/engine_grid_masking.py:69-127."""
import struct

import pytest
import torch

pytestmark = pytest.mark.gpu

PRE = {"itm": 1, "mlm": 1, "t2i": 1, "cls": 0}
CLS = {"itm": 0, "mlm": 0, "t2i": 0, "cls": 1}


def _model(loss_type, seed=0, drop_path=0.1):
    import mvlt_b200
    torch.manual_seed(seed)
    return mvlt_b200.create_model("pvlt_tiny", pretrained=True, num_classes=1000, drop_rate=0.0, drop_path_rate=drop_path,
                                  drop_block_rate=None, token_hidden_size=768, num_text_tokens=128,
                                  loss_type=dict(loss_type), pretrained_pth="").cuda().train()


def _opt(m):
    from mvlt_b200.optim import AdamW, param_groups_no_decay
    return AdamW(param_groups_no_decay(m, 0.01), lr=3e-4)


def test_set_values_and_fixed_capacity_compaction():
    from mvlt_b200 import kernels as k
    dst = torch.zeros(16, dtype=torch.int32, device="cuda")
    k.set_values(dst[2:], struct.pack("<iiQ", 7, -3, 0x0123456789ABCDEF))
    torch.cuda.synchronize()
    got = dst.cpu()
    assert got[:2].tolist() == [0, 0] and got[2:4].tolist() == [7, -3] and got[6:].abs().sum() == 0
    assert (int(got[4]) & 0xFFFFFFFF) | ((int(got[5]) & 0xFFFFFFFF) << 32) == 0x0123456789ABCDEF

    g = torch.Generator().manual_seed(3)
    n = 5000
    labels = torch.where(torch.rand(n, generator=g) < 0.05, torch.randint(0, 30522, (n,), generator=g), torch.tensor(-1)).cuda()
    want = torch.nonzero(labels != -1).flatten()
    cnt_true = int(want.numel())
    for cap in (cnt_true + 37, cnt_true, max(cnt_true - 20, 1)):
        idx = torch.full((n,), -7, dtype=torch.int32, device="cuda")
        lab = torch.full((n,), -7, dtype=torch.int64, device="cuda")
        cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
        f = torch.zeros(3, dtype=torch.float32, device="cuda")
        k.compact_labels(labels, n, -1, idx, lab, cnt, cap=cap, count_f32=f[0:1], inv_count=f[1:2], overflow=f[2:3])
        torch.cuda.synchronize()
        m = min(cap, cnt_true)
        assert int(cnt) == cnt_true and float(f[0]) == cnt_true and abs(float(f[1]) - 1.0 / cnt_true) < 1e-9
        assert torch.equal(idx[:m].long(), want[:m]) and torch.equal(lab[:m], labels[want[:m]])
        assert (idx[m:cap] == 0).all() and (lab[m:cap] == -1).all()          # padded tail: row 0, ignore label
        assert (idx[cap:] == -7).all()                                         # nothing beyond the capacity is written
        assert (float(f[2]) > 0) == (cnt_true > cap)


def test_adamw_device_hyper_matches_host_hyper():
    from mvlt_b200.optim import AdamW
    torch.manual_seed(0)
    ps_a = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in ((300, 64), (64,), (17, 5, 3, 3))]
    ps_b = [torch.nn.Parameter(p.detach().clone()) for p in ps_a]
    oa = AdamW([{"params": ps_a[:1], "weight_decay": 0.05}, {"params": ps_a[1:], "weight_decay": 0.0}], lr=1e-2)
    ob = AdamW([{"params": ps_b[:1], "weight_decay": 0.05}, {"params": ps_b[1:], "weight_decay": 0.0}], lr=1e-2)
    ob.enable_device_hyper(True)
    for t in range(5):
        lr = 1e-2 * (1 + t)
        for o in (oa, ob):
            for g in o.param_groups:
                g["lr"] = lr
        for pa, pb in zip(ps_a, ps_b):
            gr = torch.randn_like(pa)
            pa.grad, pb.grad = gr.clone(), gr.clone()
        ob.advance()
        oa.step()
        ob.step()
    for pa, pb in zip(ps_a, ps_b):
        assert torch.equal(pa, pb)
    assert ob.state_dict()["state"][0]["step"] == 5


@pytest.mark.parametrize("loss_type,tag,B", [(PRE, "pre", 8), (CLS, "cls", 8), (PRE, "pre-b128", 128)])
def test_graph_replay_matches_per_launch_path(loss_type, tag, B):
    """Two identically initialised models take the same steps: one through the per-launch path (autograd node, one stream), one
    through GraphedStep (call 1 eager warm-up, call 2 capture + replay, then replays; weight-gradient launches, the t2i head and
    the key/value chains run as parallel branches of the graph). Same seeds -> same dropout / drop-path draws.
    Phase 1 (learning rate 0: parameters stay equal, nothing amplifies rounding noise): losses AND every gradient of every
    step must agree -- a missing dependency between branches, a stale seed or a stale input buffer shows up here.
    Phase 2 (learning rate raised through the param group, as an LR scheduler does): the parameter updates must agree."""
    from mvlt_b200.graph import GraphedStep
    from mvlt_b200.synthetic import make_batch
    batches = [{k: v.cuda() for k, v in make_batch(B, seed=i).items()} for i in range(3)]
    keys = ("sup_cls_labels", "sub_cls_labels") if loss_type["cls"] else ("mlm_labels", "itm_labels")
    n1, n2 = 4, 3

    def labels_of(b):
        d = {k: b[k] for k in keys}
        if loss_type["t2i"]:
            d["target_images"] = b["images"]
        return d

    def set_lr(opt, lr):
        for g in opt.param_groups:
            g["lr"] = lr

    torch.manual_seed(11)
    ma = _model(loss_type, seed=5)
    oa = _opt(ma)
    la, ga = [], []
    for i in range(n1 + n2):
        set_lr(oa, 0.0 if i < n1 else 3e-4)
        b = batches[i % 3]
        total, stats = ma(b["images"], b["input_ids"], **labels_of(b))
        total.backward()
        if i < n1:
            ga.append({n: p.grad.detach().clone() for n, p in ma.named_parameters() if p.grad is not None})
        oa.step()
        oa.zero_grad(set_to_none=True)
        la.append(stats.clone())

    torch.manual_seed(11)
    mb = _model(loss_type, seed=5)
    ob = _opt(mb)
    cnt = max(int((b["mlm_labels"] != -1).sum()) for b in batches)
    gs = GraphedStep(mb, ob, mlm_capacity=cnt + 9 if loss_type["mlm"] else None, warmup=1)
    assert gs.eng.wgrad_stream is not None and gs.eng.branch_streams is not None
    static = {k: torch.empty_like(v) for k, v in batches[0].items()}
    lb, gb, seeds = [], [], []
    layout = None
    for i in range(n1 + n2):
        set_lr(ob, 0.0 if i < n1 else 3e-4)
        for k, v in batches[i % 3].items():
            static[k].copy_(v)
        total, stats = gs(static["images"], static["input_ids"], **labels_of(static))
        lb.append(stats.clone())
        seeds.append(gs.state.host_seeds)
        if i < n1:      # the step's gradients are still in the persistent flat buffer (zeroed by the NEXT step)
            layout = layout or gs.eng._grad_layout[0]
            flat = gs.eng._static_flat
            gb.append({n: flat[off:off + gs.eng.P[n].numel()].view(gs.eng.P[n].shape).clone() for n, off in layout})
    assert gs.captured() and gs.launches_per_replay() > 100
    assert len(set(seeds)) == n1 + n2                 # fresh dropout / drop-path seeds on every replay
    assert not gs.check_overflow()
    # phase 1: equal parameters -> the two paths differ only in summation order (padded MLM rows, concurrent fp32 atomics)
    for i in range(n1):
        a, b = la[i], lb[i]
        assert torch.allclose(a[:6], b[:6], rtol=5e-4, atol=5e-4), (tag, i, a.tolist(), b.tolist())
        assert float(a[7]) == float(b[7])                                  # labelled-row count, produced on the device
        num = den = 0.0
        worst, worst_name = 0.0, ""
        for n, g in ga[i].items():
            d = float((g - gb[i][n]).double().pow(2).sum())
            r = float(g.double().pow(2).sum())
            num, den = num + d, den + r
            if r > 0 and (d / r) ** 0.5 > worst:
                worst, worst_name = (d / r) ** 0.5, n
        assert (num / den) ** 0.5 < 3e-3 and worst < 5e-2, (tag, i, (num / den) ** 0.5, worst, worst_name)
    # phase 2: the updates agree (a stale learning rate / bias correction / gradient in the replay gives a ratio of ~1)
    for i in range(n1, n1 + n2):
        assert torch.allclose(la[i][:6], lb[i][:6], rtol=3e-2, atol=3e-2), (tag, i, la[i].tolist(), lb[i].tolist())
    init = _model(loss_type, seed=5)
    num = den = 0.0
    for (n, pa), (_, pb), (_, p0) in zip(ma.named_parameters(), mb.named_parameters(), init.named_parameters()):
        num += float((pa.detach() - pb.detach()).double().pow(2).sum())
        den += float((pa.detach() - p0.detach()).double().pow(2).sum())
    assert den > 0 and (num / den) ** 0.5 < 5e-2, (tag, num, den)
    # the optimizer state carries the step count of the replays
    assert ob.state_dict()["state"][0]["step"] == n1 + n2
    if loss_type["t2i"]:
        assert int(mb.state_dict()["t2i_head.conv4.1.num_batches_tracked"]) == n1 + n2
        for (n, ba), (_, bb) in zip(ma.named_buffers(), mb.named_buffers()):      # BatchNorm running statistics followed the replays
            if n.endswith("running_var"):
                assert torch.allclose(ba, bb, rtol=2e-2, atol=1e-4), n
    # other buffers under the same key are refused instead of silently training on stale data
    with pytest.raises(Exception):
        gs(batches[0]["images"], static["input_ids"], **labels_of(static))
    gs.detach()
    b = batches[0]
    total, _ = mb(b["images"], b["input_ids"], **labels_of(b))     # back on the plain path
    total.backward()


def test_graph_capacity_overflow_is_flagged():
    from mvlt_b200.graph import GraphedStep
    from mvlt_b200.synthetic import make_batch
    b = {k: v.cuda() for k, v in make_batch(8, seed=0).items()}
    m = _model(PRE, seed=1)
    gs = GraphedStep(m, _opt(m), mlm_capacity=8, warmup=0)
    for _ in range(2):
        gs(b["images"], b["input_ids"], mlm_labels=b["mlm_labels"], itm_labels=b["itm_labels"], target_images=b["images"])
    assert gs.check_overflow()


def test_graph_step_with_fused_gradient_clipping():
    """--clip-grad (main_vl.py:62, engine_grid_masking.py:126) inside the captured iteration: GraphedStep(max_norm=...) computes
    the global gradient norm and the clip coefficient on the device. With the learning rate at 0 both models keep equal
    parameters, so the norm of every replayed step must equal torch.nn.utils.clip_grad_norm_'s on the per-launch model."""
    from mvlt_b200.graph import GraphedStep
    from mvlt_b200.synthetic import make_batch
    B, max_norm = 8, 0.05
    batches = [{k: v.cuda() for k, v in make_batch(B, seed=i).items()} for i in range(2)]
    keys = ("sup_cls_labels", "sub_cls_labels")

    def set_lr(opt, lr):
        for g in opt.param_groups:
            g["lr"] = lr

    torch.manual_seed(11)
    ma = _model(CLS, seed=5)
    oa = _opt(ma)
    set_lr(oa, 0.0)
    want = []
    for i in range(4):
        b = batches[i % 2]
        total, _ = ma(b["images"], b["input_ids"], **{k: b[k] for k in keys})
        total.backward()
        want.append(float(torch.nn.utils.clip_grad_norm_(ma.parameters(), max_norm)))
        oa.step()
        oa.zero_grad(set_to_none=True)

    torch.manual_seed(11)
    mb = _model(CLS, seed=5)
    ob = _opt(mb)
    set_lr(ob, 0.0)
    gs = GraphedStep(mb, ob, warmup=1, max_norm=max_norm)
    static = {k: torch.empty_like(v) for k, v in batches[0].items()}
    for i in range(4):
        for k, v in batches[i % 2].items():
            static[k].copy_(v)
        gs(static["images"], static["input_ids"], **{k: static[k] for k in keys})
        got, coef = float(ob.last_grad_norm), float(ob._clip_out[0])
        assert abs(got - want[i]) <= 1e-2 * want[i], (i, got, want[i])
        assert abs(coef - min(1.0, max_norm / (got + 1e-6))) <= 1e-5, (i, coef, got)
    assert gs.captured()
    assert min(want) > max_norm                         # the clipping was active
    p0 = {n: p.detach().clone() for n, p in mb.named_parameters()}
    set_lr(ob, 3e-4)
    for k, v in batches[0].items():
        static[k].copy_(v)
    gs(static["images"], static["input_ids"], **{k: static[k] for k in keys})
    moved = sum(float((p.detach() - p0[n]).abs().sum()) for n, p in mb.named_parameters())
    assert moved > 0 and all(torch.isfinite(p).all() for p in mb.parameters())
