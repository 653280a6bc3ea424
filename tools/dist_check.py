"""Multi-GPU check of the segment-ordered, overlapped gradient exchange (mvlt_b200/libs/pvlt.py:SegmentReducer):
  torchrun --nproc-per-node 2 tools/dist_check.py
(1) the gradients a rank receives equal the average over ranks of the gradients the same step produces WITHOUT the exchange
    (same weights, same per-rank data), element for element up to the fp32 summation order of NCCL;
(2) after optimizer steps the parameters are bit-identical on every rank."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mvlt_b200  # noqa: E402
from mvlt_b200.optim import AdamW, param_groups_no_decay  # noqa: E402
from mvlt_b200.synthetic import make_batch  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lt = {"itm": 1, "mlm": 1, "t2i": 1, "cls": 0}


def build():
    torch.manual_seed(5)
    m = mvlt_b200.create_model("pvlt_tiny", pretrained=True, num_classes=1000, drop_rate=0.0, drop_path_rate=0.0, drop_block_rate=None,
                               token_hidden_size=768, num_text_tokens=128, loss_type=dict(lt), pretrained_pth="").to(dev).train()
    m.text_embeddings.dropout.p = 0.0
    return m


def grads(m, b):
    m.zero_grad(set_to_none=True)
    img = b["images"].to(dev)
    total, _ = m(img, b["input_ids"].to(dev), mlm_labels=b["mlm_labels"], itm_labels=b["itm_labels"], target_images=img)
    total.backward()
    return {n: p.grad.detach().clone() for n, p in m.named_parameters()}


B = 16
b = make_batch(B, seed=100 + rank)
plain, synced = build(), build()
synced.enable_grad_sync(True)
g_plain = grads(plain, b)
g_sync = grads(synced, b)
worst, worst_name = 0.0, ""
for n, g in g_plain.items():
    avg = g.clone()
    dist.all_reduce(avg, op=dist.ReduceOp.AVG)
    num = float((g_sync[n] - avg).norm())
    den = float(avg.norm()) + 1e-12
    if num / den > worst:
        worst, worst_name = num / den, n
# (bf16 GEMMs with fp32 atomics: two runs of the same step differ by summation order, ~1e-3 relative on the smallest tensors)
print(f"rank {rank}: worst relative difference synced vs averaged-unsynced gradient: {worst:.3e} ({worst_name})", flush=True)
assert worst < 2e-2, (worst, worst_name)

opt = AdamW(param_groups_no_decay(synced, 0.01), lr=1e-3)
for step in range(3):
    bb = make_batch(B, seed=1000 * step + rank)
    img = bb["images"].to(dev)
    total, _ = synced(img, bb["input_ids"].to(dev), mlm_labels=bb["mlm_labels"], itm_labels=bb["itm_labels"], target_images=img)
    opt.zero_grad()
    total.backward()
    opt.step()
bad = 0
for n, p in synced.named_parameters():
    lo, hi = p.detach().clone(), p.detach().clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if not torch.equal(lo, hi):
        bad += 1
print(f"rank {rank}: parameters differing across ranks after 3 steps: {bad}", flush=True)
assert bad == 0

# (3) the same exchange captured inside the CUDA graph of the whole iteration (mvlt_b200/graph.py): after replays the parameters
#     are still bit-identical on every rank, and they moved (the replays really stepped the optimizer)
from mvlt_b200.graph import GraphedStep  # noqa: E402
gs = GraphedStep(synced, opt, mlm_capacity=256, warmup=1)
static = {k: v.to(dev) for k, v in make_batch(B, seed=rank).items()}
before = {n: p.detach().clone() for n, p in synced.named_parameters()}
for step in range(5):
    bb = make_batch(B, seed=2000 * step + rank)
    for k, v in bb.items():
        static[k].copy_(v.to(dev))
    total, stats = gs(static["images"], static["input_ids"], mlm_labels=static["mlm_labels"], itm_labels=static["itm_labels"],
                      target_images=static["images"])
    assert torch.isfinite(total).item()
assert gs.captured() and not gs.check_overflow()
bad = moved = 0
for n, p in synced.named_parameters():
    lo, hi = p.detach().clone(), p.detach().clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    bad += int(not torch.equal(lo, hi))
    moved += int(not torch.equal(p.detach(), before[n]))
print(f"rank {rank}: graph replays: parameters differing across ranks: {bad}; parameters that moved: {moved}", flush=True)
assert bad == 0 and moved > 200
gs.detach()          # graphs that captured NCCL work must be destroyed before the communicator is
del gs
import gc  # noqa: E402
gc.collect()
torch.cuda.synchronize()
dist.destroy_process_group()
if rank == 0:
    print("dist_check OK")
