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


def bench_sweep(dev, rank, world, n_query=1000, n_cand=101, queries_per_step=8, warmup=1, pool=512, e2e=True, profile_hook=None):
    """Synthetic TIR protocol: per query one id row repeated n_cand times against n_cand images.
    ``value``: inputs gathered from a device-resident pool. ``e2e``: the same sweep with every step's (image, id) pairs
    copied from pinned HOST buffers (as the reference's loader hands them over, engine_grid_masking.py:349-352) on a side
    stream, double-buffered, and the ranks read back to the host: all copies inside the timed region.
    Returns the JSON sub-object bench.py prints."""
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
    import os
    model._engine().enable_branches(int(os.environ.get("MVLT_RETR_BRANCHES", "0")))     # A/B: key/value chain on a branch stream
    b = make_batch(pool, seed=99)
    img_pool = b["images"].to(dev)
    ids_pool = b["ori_input_ids"].to(dev)
    qps = queries_per_step * world
    n_steps = (n_query + qps - 1) // qps
    cand = torch.arange(n_cand, device=dev)
    lo, hi, _ = shard_bounds(qps * n_cand, rank, world)

    def indices(step):
        q0 = step * qps
        qs = torch.arange(q0, q0 + qps, device=dev)
        img_idx = ((qs.unsqueeze(1) * 37 + cand.unsqueeze(0) * 11) % pool).reshape(-1)
        ids_idx = (qs % pool).repeat_interleave(n_cand)
        return img_idx[lo:hi], ids_idx[lo:hi]

    def one(step):
        # only this rank's shard is materialised; rank_queries slices [lo:hi] of a virtual full tensor
        ii, ti = indices(step)
        return _rank_shard(model, img_pool[ii], ids_pool[ti], qps, n_cand, lo, hi, rank, world)

    def timed(fn):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for s in range(warmup):
        one(s)
    ms = timed(lambda: [one(s) for s in range(n_steps)])
    pairs = n_steps * qps * n_cand
    v = pairs / (ms / 1e3)
    out = {"metric": "itm_retrieval_pairs_per_s", "value": round(v, 1), "unit": "pairs/s", "n_query": n_steps * qps,
           "n_cand": n_cand, "pairs": pairs, "ms_total": round(ms, 2), "scaling": "strong (candidate pairs sharded)",
           "model_tflops": round(v * 8.33 / 1e3, 2), "config": "BASELINE configs[2], ITM-only forward, bf16 operands"}

    if e2e:
        # two pinned host step-batches (this rank's shard of the pairs of `qps` queries), alternated; H2D on a side stream
        host = []
        for s in range(2):
            ii, ti = indices(s)
            host.append((b["images"][ii.cpu()].pin_memory(), b["ori_input_ids"][ti.cpu()].pin_memory()))
        stage = [(torch.empty_like(host[0][0], device=dev), torch.empty_like(host[0][1], device=dev)) for _ in range(2)]
        ranks_host = [torch.empty((qps,), dtype=torch.int32).pin_memory() for _ in range(2)]
        copy_stream = torch.cuda.Stream(device=dev)

        def run_e2e():
            main = torch.cuda.current_stream()
            copied = [torch.cuda.Event() for _ in range(2)]
            consumed = [None, None]
            ready = [None, None]

            def put(i):
                j = i % 2
                with torch.cuda.stream(copy_stream):
                    if consumed[j] is not None:
                        copy_stream.wait_event(consumed[j])
                    stage[j][0].copy_(host[j][0], non_blocking=True)
                    stage[j][1].copy_(host[j][1], non_blocking=True)
                    copied[j].record(copy_stream)
            put(0)
            for i in range(n_steps):
                j = i % 2
                if i + 1 < n_steps:
                    put(i + 1)
                main.wait_event(copied[j])
                r = _rank_shard(model, stage[j][0], stage[j][1], qps, n_cand, lo, hi, rank, world)
                consumed[j] = torch.cuda.Event()
                consumed[j].record(main)
                ranks_host[j].copy_(r, non_blocking=True)
                ready[j] = torch.cuda.Event()
                ready[j].record(main)
                if i > 0:
                    ready[1 - j].synchronize()
            ready[(n_steps - 1) % 2].synchronize()
        ms2 = timed(run_e2e)
        h2d = host[0][0].numel() * 4 + host[0][1].numel() * 8
        out["e2e"] = {"value": round(pairs / (ms2 / 1e3), 1), "unit": "pairs/s", "h2d_bytes_per_step": h2d,
                      "d2h_bytes_per_step": qps * 4, "ms_total": round(ms2, 2),
                      "what": "pairs of every step copied from pinned host buffers (side stream, double-buffered), ranks read back"}
    if profile_hook is not None and rank == 0 and world == 1:
        try:
            out.update(profile_hook(one))
        except Exception as ex:      # reporting only
            out["profile_error"] = repr(ex)[:200]
    return out


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
