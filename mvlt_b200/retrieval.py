"""Candidate-sharded ITM retrieval sweep (zero-shot ITR / TIR).

Reference: /root/reference/engine_grid_masking.py:336-393 scores the 101 (query, candidate) pairs of one query per
forward on EVERY rank redundantly and also runs the unused MLM + t2i heads. Here the pairs of a chunk of queries
are block-partitioned across the ranks of one NVSwitch box, each rank runs encoder + ITM head only, the [.,2]
logits are all-gathered (a few hundred bytes per query over NCCL/NVLink) and the rank of candidate 0 is computed
by the ``itm_rank`` kernel (softmax p(match), descending, engine_grid_masking.py:360-384).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib
from . import kernels as k

F32 = torch.float32


def shard_bounds(n_pairs: int, rank: int, world: int):
    """Balanced contiguous block partition of the flattened pair list (101 is prime: per-query splits would not be)."""
    per = (n_pairs + world - 1) // world
    lo = min(rank * per, n_pairs)
    return lo, min(lo + per, n_pairs), per


def gather_shards(local, n, world, group=None):
    """local: [per, 2] logits of this rank's block (zero padded); returns the [n, 2] logits in global pair order."""
    if world == 1:
        return local[:n]
    full = torch.empty((world * local.shape[0], local.shape[1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(full, local.contiguous(), group=group)
    return full[:n].contiguous()


@torch.no_grad()
def score_pairs(model, images, input_ids):
    """ITM logits fp32 [n, 2] for n aligned (image, text) pairs: encoder + ITM head only."""
    return model.itm_logits(images, input_ids)


@torch.no_grad()
def rank_queries(model, images, input_ids, n_cand, rank=0, world=1, group=None):
    """images [Q*n_cand, 3, H, W], input_ids [Q*n_cand, T] (flattened query-major; candidate 0 is the positive).
    Every rank passes the SAME full tensors' shard-local view via ``shard_bounds``; returns int32 ranks [Q]."""
    n = images.shape[0]
    Q = n // n_cand
    dev = images.device
    lo, hi, per = shard_bounds(n, rank, world)
    local = torch.zeros((per, 2), dtype=F32, device=dev)
    if hi > lo:
        local[: hi - lo] = score_pairs(model, images[lo:hi], input_ids[lo:hi])
    logits = gather_shards(local, n, world, group)
    ranks = torch.empty((Q,), dtype=torch.int32, device=dev)
    k.itm_rank(logits, Q, n_cand, ranks)
    return ranks, logits.view(Q, n_cand, 2)


def accuracy_at(ranks, ks=(1, 5, 10)):
    """acc@k as engine_grid_masking.py:380-393 counts it (index < k)."""
    r = ranks.cpu()
    return {kk: float((r < kk).sum()) / max(r.numel(), 1) for kk in ks}


def bench_sweep(dev, rank, world, n_query=1000, n_cand=101, queries_per_step=8, warmup=1, pool=512):
    """Synthetic TIR protocol: per query one id row repeated n_cand times against n_cand images drawn
    from a device-resident pool. Returns the JSON sub-object bench.py prints."""
    import mvlt_b200
    from .synthetic import make_batch
    torch.manual_seed(4321)
    model = mvlt_b200.create_model("pvlt_tiny", pretrained=True, num_classes=1000, drop_rate=0.0, drop_path_rate=0.1,
                                   drop_block_rate=None, token_hidden_size=768, num_text_tokens=128,
                                   loss_type={"itm": 1, "mlm": 0, "t2i": 0, "cls": 0}, pretrained_pth="").to(dev).eval()
    if world > 1:   # identical weights on every rank
        for p in model.parameters():
            dist.broadcast(p.data, 0)
        _lib.params_written()   # .data writes bypass the version counters the engine's bf16 weight copies are keyed on
    b = make_batch(pool, seed=99)
    img_pool = b["images"].to(dev)
    ids_pool = b["ori_input_ids"].to(dev)
    qps = queries_per_step * world
    n_steps = (n_query + qps - 1) // qps
    cand = torch.arange(n_cand, device=dev)

    def one(step):
        q0 = step * qps
        qs = torch.arange(q0, q0 + qps, device=dev)
        img_idx = ((qs.unsqueeze(1) * 37 + cand.unsqueeze(0) * 11) % pool).reshape(-1)
        ids_idx = (qs % pool).repeat_interleave(n_cand)
        lo, hi, _ = shard_bounds(qps * n_cand, rank, world)
        # only this rank's shard is materialised; rank_queries slices [lo:hi] of a virtual full tensor
        images = img_pool[img_idx[lo:hi]]
        ids = ids_pool[ids_idx[lo:hi]]
        return _rank_shard(model, images, ids, qps, n_cand, lo, hi, rank, world)

    for s in range(warmup):
        one(s)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    acc1 = 0
    for s in range(n_steps):
        ranks = one(s)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    pairs = n_steps * qps * n_cand
    v = pairs / (ms / 1e3)
    return {"metric": "itm_retrieval_pairs_per_s", "value": round(v, 1), "unit": "pairs/s", "n_query": n_steps * qps,
            "n_cand": n_cand, "pairs": pairs, "ms_total": round(ms, 2), "scaling": "strong (candidate pairs sharded)",
            "model_tflops": round(v * 8.33 / 1e3, 2), "config": "BASELINE configs[2], ITM-only forward, bf16 operands"}


@torch.no_grad()
def _rank_shard(model, images, ids, Q, n_cand, lo, hi, rank, world):
    dev = images.device
    n = Q * n_cand
    per = (n + world - 1) // world
    local = torch.zeros((per, 2), dtype=F32, device=dev)
    if hi > lo:
        local[: hi - lo] = score_pairs(model, images, ids)
    logits = gather_shards(local, n, world)
    ranks = torch.empty((Q,), dtype=torch.int32, device=dev)
    k.itm_rank(logits, Q, n_cand, ranks)
    return ranks


def fit_planted_itm(model, steps: int = 300, batch: int = 64, lr: float = 5e-4, weight_decay: float = 0.01, n_classes=None,
                    log=None):
    """Fits ``model`` (ITM head enabled) on the planted matched / mismatched pairs of ``synthetic.planted_pairs`` with the
    own AdamW: the preparation step of the planted-positive retrieval protocol (see synthetic.py). Returns the list of
    logged (step, loss, accuracy)."""
    from .optim import AdamW, param_groups_no_decay
    from .synthetic import PLANTED_CLASSES, planted_pairs
    dev = next(model.parameters()).device
    opt = AdamW(param_groups_no_decay(model, weight_decay), lr=lr)
    model.train()
    hist = []
    for step in range(steps):
        b = planted_pairs(batch, seed=step, n_classes=n_classes or PLANTED_CLASSES, device=dev)
        for g in opt.param_groups:          # linear warm-up over the first 20 steps
            g["lr"] = lr * min(1.0, (step + 1) / 20.0)
        total, stats = model(b["images"], b["input_ids"], itm_labels=b["itm_labels"], only=("itm",))
        opt.zero_grad()
        total.backward()
        opt.step()
        if step % 25 == 0 or step == steps - 1:
            s = stats.tolist()
            hist.append((step, s[2], s[8] / batch))
            if log is not None:
                log(f"planted ITM fit step {step}: loss {s[2]:.4f} acc {s[8] / batch:.3f}")
    model.eval()
    return hist
