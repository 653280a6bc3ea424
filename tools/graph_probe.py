"""Probe: how much device time would CUDA-graph replay of the forward + backward save over stream launches (with programmatic
dependent launch)? Fixed inputs / seeds / label count (randomness is NOT advanced between replays: timing probe only)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mvlt_b200  # noqa: E402
from mvlt_b200.synthetic import make_batch  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
lt = {"itm": 1, "mlm": 1, "t2i": 1, "cls": 0}
m = mvlt_b200.create_model("pvlt_tiny", pretrained=True, num_classes=1000, drop_rate=0.0, drop_path_rate=0.1, drop_block_rate=None,
                           token_hidden_size=768, num_text_tokens=128, loss_type=lt, pretrained_pth="").to(dev).train()
b = make_batch(128, 0)
n = int((b["mlm_labels"] != -1).sum())
b = {k: v.to(dev) for k, v in b.items()}


def fwd_bwd():
    total, stats = m(b["images"], b["input_ids"], mlm_labels=b["mlm_labels"], itm_labels=b["itm_labels"], target_images=b["images"], mlm_count=n)
    total.backward()
    m.zero_grad(set_to_none=True)
    return total


def timed(fn, k):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(k):
        fn()
    e1.record()
    host = (time.perf_counter() - t0) / k * 1e3
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k, host


for _ in range(4):
    fwd_bwd()
dev_ms, host_ms = timed(fwd_bwd, 10)
print(f"eager  fwd+bwd: {dev_ms:.3f} ms device, {host_ms:.3f} ms host enqueue per step", flush=True)
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        fwd_bwd()
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    out = fwd_bwd()
torch.cuda.synchronize()
for _ in range(3):
    g.replay()
dev_ms, host_ms = timed(g.replay, 10)
print(f"graph  fwd+bwd: {dev_ms:.3f} ms device, {host_ms:.3f} ms host enqueue per step; loss {float(out):.4f}", flush=True)
