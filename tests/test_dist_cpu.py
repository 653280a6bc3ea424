"""Host-side multi-process logic on CPU (gloo, world_size 2): candidate-pair sharding + score gather of the retrieval
sweep, and the metric reduction of the train/eval loops. The CUDA side of the same paths is covered by -m gpu tests."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mvlt_b200 import retrieval
        from mvlt_b200.utils import MetricLogger
        from oracle import pvlt_oracle as O
        n_cand, Q = 101, 3
        n = n_cand * Q
        g = torch.Generator().manual_seed(0)
        truth = torch.randn((n, 2), generator=g)                  # the logits every pair "should" get
        lo, hi, per = retrieval.shard_bounds(n, rank, world)
        local = torch.zeros((per, 2))
        local[: hi - lo] = truth[lo:hi]                            # this rank scores only its block
        full = retrieval.gather_shards(local, n, world)
        assert torch.equal(full, truth)
        ranks = [O.retrieval_rank(full.view(Q, n_cand, 2)[i]) for i in range(Q)]
        ml = MetricLogger()
        ml.update(loss=float(rank + 1), n=2)
        ml.synchronize_between_processes()
        q.put((rank, lo, hi, ranks, ml.meters["loss"].global_avg))
    finally:
        dist.destroy_process_group()


def test_retrieval_sharding_and_metric_reduce_gloo_world2():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, ranks0, avg0), (r1, lo1, hi1, ranks1, avg1) = res
    assert lo0 == 0 and hi0 == lo1 and hi1 == 303          # contiguous, disjoint, complete cover of 3 x 101 pairs
    assert abs((hi0 - lo0) - (hi1 - lo1)) <= 1              # balanced although 101 is prime
    assert ranks0 == ranks1                                 # every rank derives identical rankings
    assert avg0 == avg1 == pytest.approx(1.5)


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_shard_bounds_cover(world):
    from mvlt_b200 import retrieval
    n = 101 * 4
    seen = []
    for r in range(world):
        lo, hi, per = retrieval.shard_bounds(n, r, world)
        assert hi - lo <= per
        seen += list(range(lo, hi))
    assert seen == list(range(n))


def _sync_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mvlt_b200.libs.pvlt import allreduce_flat_
        flat = torch.arange(10, dtype=torch.float32) * (rank + 1)      # this rank's "flat gradient buffer"
        views = [flat[:4].view(2, 2), flat[4:]]                         # per-parameter views handed to autograd
        allreduce_flat_(flat, True)
        q.put((rank, flat.tolist(), views[0].tolist()))
    finally:
        dist.destroy_process_group()


def test_flat_gradient_allreduce_gloo_world2():
    """The data-parallel exchange of the training step: ONE all-reduce (average) over the flat gradient buffer; the
    per-parameter views alias it, so they see the reduced values."""
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sync_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [1.5 * i for i in range(10)]
    for _, flat, v0 in res:
        assert flat == pytest.approx(want)
        assert v0 == [[0.0, 1.5], [3.0, 4.5]]


def _segment_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mvlt_b200.libs.pvlt import SegmentReducer
        flat = torch.arange(12, dtype=torch.float32) * (rank + 1)
        red = SegmentReducer(flat, [(0, 4), (4, 4), (4, 9), (9, 9), (9, 12)], True)
        red.segment_done(0)
        after0 = flat.tolist()
        red.segment_done(2, None)    # (the engine's call form: an optional event of the weight-gradient stream, None on CPU)
        red.segment_done(0)          # a repeated notification must not reduce the segment twice
        red.finish()                 # reduces whatever was not announced (segment 4)
        q.put((rank, after0, flat.tolist()))
    finally:
        dist.destroy_process_group()


def test_segment_ordered_gradient_exchange_gloo_world2():
    """The overlapped exchange of the training step: the flat gradient buffer is averaged segment by segment, in the order
    the backward pass completes them; every element is reduced exactly once."""
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_segment_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, after0, final in res:
        assert after0[:4] == pytest.approx([1.5 * i for i in range(4)])
        assert after0[4:] == pytest.approx([float(i * (rank + 1)) for i in range(4, 12)])
        assert final == pytest.approx([1.5 * i for i in range(12)])


def test_gradient_layout_is_in_completion_order():
    from mvlt_b200.engine import PVLTEngine
    seg = PVLTEngine.grad_segment
    assert seg("mlm_head.bias") == 0 and seg("t2i_head.conv4.0.weight") == 0 and seg("itm_head_embed.0.weight") == 0
    assert seg("block4.1.mlp.fc1.weight") == 1 and seg("pos_embed4") == 1 and seg("patch_embed4.proj.weight") == 1
    assert seg("block3.0.attn.sr.weight") == 2 and seg("text_embed2.0.weight") == 3 and seg("text_pos_embed1") == 4
    assert seg("text_embeddings.word_embeddings.weight") == 4
