"""The reference-named train / eval loops (engine_grid_masking.py) on synthetic loaders, plus size-independent
properties at the BASELINE batch size (determinism, candidate-permutation equivariance, shard equivalence)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

PRE = {"itm": 1, "mlm": 1, "t2i": 1, "cls": 0}
CLS = {"itm": 0, "mlm": 0, "t2i": 0, "cls": 1}


def _model(loss_type, drop_path=0.1, seed=0):
    import mvlt_b200
    torch.manual_seed(seed)
    return mvlt_b200.create_model("pvlt_tiny", pretrained=True, num_classes=1000, drop_rate=0.0, drop_path_rate=drop_path,
                                  drop_block_rate=None, token_hidden_size=768, num_text_tokens=128,
                                  loss_type=dict(loss_type), pretrained_pth="").cuda()


def _loader(n, B, keys_eval=False):
    from mvlt_b200.synthetic import make_batch
    out = []
    for i in range(n):
        b = make_batch(B, seed=i)
        b["image"] = b.pop("images")
        if keys_eval:
            b["images"] = b["image"]
        out.append(b)
    return out


class _Args:
    loss_type = PRE
    eval_retrieval_tir = True
    eval_retrieval_itr = False


@pytest.mark.parametrize("own_optimizer", [False, True, "graph"])
def test_train_and_eval_loops_run_and_learn(own_optimizer):
    """``"graph"``: the same loop with ``args.cuda_graph`` -- every iteration after the first two is a CUDA-graph replay."""
    import engine_grid_masking as E
    m = _model(PRE, drop_path=0.0)
    m.text_embeddings.dropout.p = 0.0
    args = _Args()
    args.cuda_graph = own_optimizer == "graph"
    if own_optimizer:     # the multi-tensor sm_100a AdamW with timm's no-decay grouping (main_vl.py:308)
        from mvlt_b200.optim import AdamW, param_groups_no_decay
        opt = AdamW(param_groups_no_decay(m, 0.01), lr=2e-4)
    else:
        opt = torch.optim.AdamW(m.parameters(), lr=2e-4)
    data = _loader(2, 8) * 6                        # 12 steps over two fixed batches: the loss must go down
    s0 = E.train_one_epoch_vl(m, None, data[:2], opt, torch.device("cuda"), 0, None, args=args)
    for ep in range(1, 5):
        s1 = E.train_one_epoch_vl(m, None, data[:2], opt, torch.device("cuda"), ep, None, args=args)
    if args.cuda_graph:
        gs = m.__dict__["_graphed_step"]
        assert len(gs._graphs) == 2 and gs.launches_per_replay((False, (8, 3, 256, 256))) > 100     # unmasked / masked iteration
    assert all(math.isfinite(v) for v in s1.values())
    assert s1["total_loss"] < s0["total_loss"] - 0.2, (s0, s1)
    ev = E.evaluate_vl(_loader(2, 8), m, torch.device("cuda"), _Args())
    assert set(ev) >= {"mlm_acc", "itm_acc", "t2i_psnr", "total_loss"} and 0 <= ev["mlm_acc"] <= 1 and ev["t2i_psnr"] > 0
    sd = m.state_dict()
    assert int(sd["t2i_head.conv4.1.num_batches_tracked"]) == 10   # one BatchNorm update per training forward


def test_eval_fused_metrics_equal_dict_path_metrics():
    """evaluate_vl takes accuracies from the fused kernels; they must equal vl_scores on materialised logits."""
    from mvlt_b200.libs.vl_scores import compute_mlm_score, compute_score_with_logits
    m = _model(PRE).eval()
    b = _loader(1, 16)[0]
    with torch.no_grad():
        out = m(b["image"].cuda(), b["input_ids"].cuda())
        _, st = m.forward_losses(b["image"].cuda(), b["input_ids"].cuda(), mlm_labels=b["mlm_labels"],
                                 itm_labels=b["itm_labels"], only=("mlm", "itm"))
    s = st.tolist()
    acc = compute_mlm_score(out["mlm_logits"], b["mlm_labels"].cuda())
    assert abs(acc - s[6] / s[7]) < 1e-6
    itm = compute_score_with_logits(out["itm_logits"].view(-1, 2), b["itm_labels"].view(-1).cuda()).sum().item()
    assert itm == s[8]
    ce = torch.nn.functional.cross_entropy(out["mlm_logits"].float().view(-1, 30522), b["mlm_labels"].view(-1).cuda(),
                                           ignore_index=-1).item()
    assert abs(ce - s[1]) < 2e-2 * ce          # fused path scores bf16 logits, dict path fp32 logits


def test_retrieval_loop_ranks_identical_to_oracle_ranking_of_same_logits():
    import engine_grid_masking as E
    from mvlt_b200 import retrieval
    from mvlt_b200.synthetic import make_batch
    from oracle import pvlt_oracle as O
    m = _model({"itm": 1, "mlm": 0, "t2i": 0, "cls": 0}).eval()
    pool = make_batch(101, seed=5)
    loader = []
    for q in range(3):
        ids = pool["ori_input_ids"][q:q + 1].repeat(101, 1)
        loader.append({"images_101": pool["images"].unsqueeze(0), "ori_input_ids_101": ids.unsqueeze(0), "info_list": []})
    res = E.evaluate_retrieval(loader, m, torch.device("cuda"), _Args())
    assert set(res) == {"acc@1", "acc@5", "acc@10"}
    imgs = pool["images"].cuda()
    for q in range(3):
        ids = loader[q]["ori_input_ids_101"][0].cuda()
        ranks, logits = retrieval.rank_queries(m, imgs, ids, 101)
        assert int(ranks[0]) == O.retrieval_rank(logits[0].cpu())          # identical ranking (same logits)
        # shard equivalence: scoring the two halves separately gives bit-identical logits => identical ranks
        lo, hi, _ = retrieval.shard_bounds(101, 0, 2)
        a = retrieval.score_pairs(m, imgs[lo:hi], ids[lo:hi])
        b = retrieval.score_pairs(m, imgs[hi:], ids[hi:])
        assert torch.equal(torch.cat([a, b]), logits[0])
        # permutation equivariance of the candidate axis
        perm = torch.randperm(101, generator=torch.Generator().manual_seed(q))
        lp = retrieval.score_pairs(m, imgs[perm.cuda()], ids[perm.cuda()])
        assert torch.equal(lp, logits[0][perm.cuda()])


def test_recognition_loop():
    import engine_grid_masking as E
    m = _model(CLS).eval()
    res = E.evaluate_recognition(_loader(2, 8, keys_eval=True), m, torch.device("cuda"), _Args())
    assert 0 <= res["sup"][0] <= 1 and 0 <= res["sub"][0] <= 1


def test_planted_recognition_loop_predictions_identical_to_fp32_oracle():
    """the reference-named recognition loop (engine_grid_masking.py:396-462) on the planted recognition protocol
    (mvlt_b200/synthetic.py:planted_cls_set; last linear layer of the two category heads fitted by the fp32 oracle on
    brightness-planted classes). The 48- / 122-way argmax decisions of the bf16 sm_100a kernels must equal the fp32 oracle's
    sample by sample, so accuracy and macro / micro / weighted F1 of the loop are the oracle's, and the protocol must be
    well conditioned (decision margin above twice the largest logit error of the sample)."""
    import json, os
    import engine_grid_masking as E
    import mvlt_b200
    from mvlt_b200.synthetic import planted_cls_set
    from oracle import pvlt_oracle as O
    sd = O.fit_cls_probes(O.make_state_dict("pvlt_tiny", CLS, seed=5), *planted_cls_set(96, seed=0))
    m = mvlt_b200.create_model("pvlt_tiny", pretrained=True, num_classes=1000, drop_rate=0.0, drop_path_rate=0.1,
                               drop_block_rate=None, token_hidden_size=768, num_text_tokens=128, loss_type=dict(CLS),
                               pretrained_pth="")
    m.load_state_dict(sd)
    m = m.cuda().eval()
    loader, rows, preds = [], [], {"sup": [], "sub": []}
    labels = {"sup": [], "sub": []}
    for i in range(3):
        images, ids, sup, sub = planted_cls_set(16, seed=1 + i)
        loader.append({"images": images, "ori_input_ids": ids, "sup_cls_labels": sup, "sub_cls_labels": sub})
        with torch.no_grad():
            got = m(images.cuda(), ids.cuda())
            ref = O.forward(sd, images, ids, CLS, training=False)
        for nm, lab in (("sup", sup), ("sub", sub)):
            lg, lr = got[f"{nm}_cls_logits"].view(16, -1).float().cpu(), ref[f"{nm}_cls_logits"].view(16, -1)
            top = lr.topk(2, dim=-1).values
            for j in range(16):
                rows.append(dict(head=nm, sample=16 * i + j, label=int(lab[j]), pred_gpu=int(lg[j].argmax()),
                                 pred_oracle=int(lr[j].argmax()), margin=float(top[j, 0] - top[j, 1]),
                                 max_err=float((lg[j] - lr[j]).abs().max())))
            preds[nm] += lr.argmax(-1).tolist()
            labels[nm] += lab.view(-1).tolist()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open("gpurun_out/planted_recognition.json", "w"), indent=1)
    for r in rows:
        assert r["pred_oracle"] == r["label"], r                      # the fp32 oracle solves the planted task
        assert r["margin"] > 2 * r["max_err"], r                      # ... with room: bf16 noise cannot flip a decision
        assert r["pred_gpu"] == r["pred_oracle"], r
    res = E.evaluate_recognition(loader, m, torch.device("cuda"), _Args())
    for nm in ("sup", "sub"):
        want = E.calculate_cls_metrics(labels[nm], preds[nm])
        assert tuple(res[nm]) == tuple(want), (nm, res[nm], want)


def test_full_batch_properties_b128():
    """BASELINE configs[1] size (B=128): finite step, MLM loss ~ ln(30522) at random init, forward repeatable."""
    from mvlt_b200.synthetic import make_batch
    m = _model(PRE, drop_path=0.0)
    m.text_embeddings.dropout.p = 0.0
    m.train()
    b = make_batch(128, seed=1)
    img, ids = b["images"].cuda(), b["input_ids"].cuda()
    kw = dict(mlm_labels=b["mlm_labels"], itm_labels=b["itm_labels"], target_images=img)
    t1, s1 = m(img, ids, **kw)
    t1.backward()
    g1 = m.block1[0].mlp.fc1.weight.grad.clone()
    m.zero_grad(set_to_none=True)
    t2, s2 = m(img, ids, **kw)
    assert torch.isfinite(s1).all() and torch.isfinite(g1).all()
    assert abs(s1[1].item() - math.log(30522)) < 0.5
    # repeatable up to the summation order of the fp32 atomics (loss / BatchNorm-statistic accumulators)
    assert abs(s1[1].item() - s2[1].item()) < 1e-4 and abs(s1[2].item() - s2[2].item()) < 1e-4
    assert float(g1.abs().sum()) > 0


def test_own_adamw_keeps_bf16_weight_copies_fresh_and_tracks_torch_adamw():
    """The own optimizer writes the fp32 masters through raw pointers (no autograd version bump): the bf16 copies the GEMMs
    read must still follow them every step (csrc/optim.cu writes them in the same launch), for Linear weights, the k=s
    convolutions (permuted) and the 3x3 t2i convolutions (permuted + flipped/transposed). Cross-check: the loss trajectory
    equals the one of ``torch.optim.AdamW`` (which bumps versions and therefore takes the recast path) on a twin model."""
    from mvlt_b200.optim import AdamW, param_groups_no_decay
    from mvlt_b200.synthetic import make_batch
    ma, mb = _model(PRE, drop_path=0.0, seed=3), _model(PRE, drop_path=0.0, seed=3)
    mb.load_state_dict(ma.state_dict())
    for m in (ma, mb):
        m.text_embeddings.dropout.p = 0.0
        m.train()
    oa = AdamW(param_groups_no_decay(ma, 0.05), lr=3e-4)
    ob = torch.optim.AdamW(param_groups_no_decay(mb, 0.05), lr=3e-4)
    batches = [make_batch(8, seed=s) for s in (0, 1)]
    la, lb = [], []
    from mvlt_b200 import _lib
    epoch0 = None
    for step in range(6):
        if step == 1:
            epoch0 = _lib.WEIGHT_EPOCH      # every parameter is registered after the first forward
        b = batches[step % 2]
        img, ids = b["images"].cuda(), b["input_ids"].cuda()
        for m, o, out in ((ma, oa, la), (mb, ob, lb)):
            total, stats = m(img, ids, mlm_labels=b["mlm_labels"], itm_labels=b["itm_labels"], target_images=img)
            o.zero_grad()
            total.backward()
            o.step()
            out.append(stats[:6].tolist())
    assert _lib.WEIGHT_EPOCH == epoch0, "the own optimizer forced a full recast although every copy is registered"
    eng = ma._engine()
    P = dict(ma.named_parameters())
    convs = eng._conv_names()
    checked = 0
    for name, w in eng.W.items():
        p = P[name].detach()
        if name in convs:
            co, ci = p.shape[0], p.shape[1]
            want = p.view(co, ci, -1).permute(0, 2, 1).reshape(co, -1).to(torch.bfloat16)
        else:
            want = p.view(w.shape).to(torch.bfloat16)
        assert torch.equal(w, want), name
        checked += 1
    for name, w in eng.t2i.W.items():
        p = P[name.replace("^T", "")].detach()
        co, ci = p.shape[0], p.shape[1]
        if name.endswith("^T"):
            want = p.view(co, ci, 9).flip(-1).permute(1, 2, 0).reshape(ci, -1).to(torch.bfloat16)
        else:
            want = p.view(co, ci, 9).permute(0, 2, 1).reshape(co, -1).to(torch.bfloat16)
        assert torch.equal(w, want), name
        checked += 1
    assert checked > 60
    # the weights really moved, and both optimizers walked the same path (bf16 GEMMs + fp32 atomics: small run-to-run noise)
    assert la[-2][0] < la[0][0] - 0.1 and la[-1][0] < la[1][0] - 0.1, la
    for sa, sb in zip(la, lb):
        for x, y in zip(sa, sb):
            assert abs(x - y) <= 2e-2 * max(1.0, abs(y)), (la, lb)
    pa, pb = dict(ma.named_parameters()), dict(mb.named_parameters())
    for name in ("block1.0.mlp.fc1.weight", "block3.1.attn.kv.weight", "patch_embed2.proj.weight", "t2i_head.conv4.0.weight"):
        d = (pa[name] - pb[name]).abs().max().item()
        assert d <= 6 * 3e-4, (name, d)     # at most a few lr-sized steps apart (sign flips of ~zero gradients)


@pytest.mark.parametrize("loss_type,tag", [(PRE, "pre"), ({"itm": 1, "mlm": 1, "t2i": 0, "cls": 1}, "enc")])
def test_evaluate_vl_metrics_match_oracle(loss_type, tag):
    """the reference-named eval loop (three forwards per batch on the fused metric path) against the fp32
    oracle's restatement of engine_grid_masking.py:186-300 on the same weights, batches and grid-masked images."""
    import numpy as np
    import engine_grid_masking as E
    from mvlt_b200 import masking
    from oracle import grid_mask as ogm
    from oracle import pvlt_oracle as O
    import mvlt_b200
    m = mvlt_b200.create_model("pvlt_tiny", pretrained=True, num_classes=1000, drop_rate=0.0, drop_path_rate=0.1,
                               drop_block_rate=None, token_hidden_size=768, num_text_tokens=128, loss_type=dict(loss_type),
                               pretrained_pth="")
    sd = O.make_state_dict("pvlt_tiny", loss_type, seed=2)
    m.load_state_dict(sd)
    m = m.cuda()
    loader, want, n = [], {}, 0
    for i in range(2):
        b = O.make_inputs(6, seed=20 + i)
        seeds = [masking.sample_seed(9, 6 * i + j) for j in range(6)]
        mk = np.stack([ogm.masked_fill(b["images"][j].numpy(), ogm.expand(ogm.grid_py(seeds[j]))) for j in range(6)])
        s = dict(b, image=b["images"], masked_images=torch.from_numpy(mk))
        loader.append(s)
        r = O.evaluate_vl_batch(sd, s, loss_type)
        for k_, v in r.items():
            want[k_] = want.get(k_, 0.0) + v * 6
        n += 6

    class A:
        pass
    a = A()
    a.loss_type = loss_type
    got = E.evaluate_vl(loader, m, torch.device("cuda"), a)
    want = {k_: v / n for k_, v in want.items()}
    assert abs(got["total_loss"] - want["total_loss"]) <= 2e-2 * max(1.0, want["total_loss"]), (got, want)
    for key in ("mlm_acc", "itm_acc", "sup_cls_acc", "sub_cls_acc"):      # argmax decisions: at most one near-tie apart
        assert abs(got[key] - want[key]) <= 1.0 / n + 1e-9, (key, got, want)
    if loss_type["t2i"]:
        assert abs(got["t2i_psnr"] - want["t2i_psnr"]) <= 0.05, (got["t2i_psnr"], want["t2i_psnr"])


def test_planted_positive_retrieval_ranks_identical_to_fp32_oracle():
    """north_star: "retrieval rankings must be identical for the synthetic protocol" (SURVEY 7 H3 option iii, README.md).
    The planted-positive TIR protocol of mvlt_b200/synthetic.py: random-init PVLT-tiny encoder, last ITM linear layer fitted
    by the fp32 oracle on brightness-planted pairs; 24 queries x 101 candidates. The rank of the positive -- the quantity
    engine_grid_masking.py:360-384 computes -- from the bf16 sm_100a kernels must equal the fp32 oracle's for every query
    (and the planted value), and so must acc@1/5/10 from the reference-named loop."""
    import engine_grid_masking as E
    import mvlt_b200
    from mvlt_b200 import retrieval
    from mvlt_b200.synthetic import planted_fit_set, planted_tir_query
    from oracle import pvlt_oracle as O
    lt = {"itm": 1, "mlm": 0, "t2i": 0, "cls": 0}
    sd = O.make_state_dict("pvlt_tiny", lt, seed=4)
    sd = O.fit_itm_probe(sd, *planted_fit_set(128))
    m = mvlt_b200.create_model("pvlt_tiny", pretrained=True, num_classes=1000, drop_rate=0.0, drop_path_rate=0.1,
                               drop_block_rate=None, token_hidden_size=768, num_text_tokens=128, loss_type=dict(lt),
                               pretrained_pth="")
    m.load_state_dict(sd)
    m = m.cuda().eval()
    loader, rows = [], []
    for q in range(24):
        img, ids, planted = planted_tir_query(q)
        ranks, logits = retrieval.rank_queries(m, img.cuda(), ids.cuda(), 101)
        with torch.no_grad():
            ref = torch.cat([O.forward(sd, img[i:i + 51], ids[i:i + 51], lt, training=False)["itm_logits"].view(-1, 2)
                             for i in range(0, 101, 51)])
        r_ref = O.retrieval_rank(ref)
        lg = logits[0].cpu()
        d_ref, d_gpu = ref[:, 1] - ref[:, 0], lg[:, 1] - lg[:, 0]
        gap = (d_ref[1:] - d_ref[0]).abs()
        near = 1 + int(gap.argmin())                       # the candidate whose score is closest to the positive's
        rows.append(dict(q=q, planted=planted, rank_gpu=int(ranks[0]), rank_oracle=r_ref,
                         max_err=float((d_ref - d_gpu).abs().max()),
                         near_err=float(max((d_ref[0] - d_gpu[0]).abs(), (d_ref[near] - d_gpu[near]).abs())),
                         margin=float(gap.min())))
        loader.append({"images_101": img.unsqueeze(0), "ori_input_ids_101": ids.unsqueeze(0), "info_list": []})
    import json, os
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open("gpurun_out/planted_retrieval.json", "w"), indent=1)
    for r in rows:
        assert r["rank_gpu"] == r["rank_oracle"] == r["planted"], r
        # the protocol really is well conditioned: the positive's gap to its nearest competitor dwarfs the bf16 error of the
        # two scores that decide the rank (and exceeds twice the largest error anywhere in the candidate list)
        assert r["margin"] > 4 * r["near_err"] and r["margin"] > 2 * r["max_err"], r
    res = E.evaluate_retrieval(loader, m, torch.device("cuda"), _Args())
    want = {f"acc@{k}": sum(int(r["rank_oracle"] < k) for r in rows) / 24 for k in (1, 5, 10)}
    assert res == want, (res, want)
