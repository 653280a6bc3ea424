"""One profiled pre-training step (B=128) between cudaProfilerStart/Stop, for ncu:

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -c 3 \
      -o gpurun_out/prof python tools/profile_step.py

With --dump-gemms it also prints one line per GEMM launch (shape, layouts, CUDA-event time) for the tuning tables.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mvlt_b200  # noqa: E402
from mvlt_b200 import _lib, kernels, masking  # noqa: E402
from mvlt_b200.synthetic import make_batch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--dump-gemms", action="store_true")
ap.add_argument("--retrieval", action="store_true")
ap.add_argument("--graph", action="store_true",
                help="profile ONE CUDA-graph replay of the whole iteration (forward, losses, backward, AdamW; mvlt_b200/graph.py): "
                     "ncu reports the kernel nodes of the graph individually")
args = ap.parse_args()

dev = torch.device("cuda", 0)
torch.manual_seed(0)
lt = {"itm": 1, "mlm": 1, "t2i": 1, "cls": 0}
m = mvlt_b200.create_model("pvlt_tiny", pretrained=True, num_classes=1000, drop_rate=0.0, drop_path_rate=0.1,
                           drop_block_rate=None, token_hidden_size=768, num_text_tokens=128, loss_type=lt,
                           pretrained_pth="").to(dev).train()
b = make_batch(args.batch, 0)
n = int((b["mlm_labels"] != -1).sum())
b = {k: v.to(dev) for k, v in b.items()}
seeds = torch.arange(args.batch, device=dev)


def step(i):
    img = b["images"]
    x = masking.apply_grid_mask(img, masking.grid_mask_batch(seeds + i)) if i % 2 else img
    if args.retrieval:
        with torch.no_grad():
            m.eval()
            return m.itm_logits(x, b["ori_input_ids"])
    total, _ = m(x, b["input_ids"], mlm_labels=b["mlm_labels"], itm_labels=b["itm_labels"], target_images=img, mlm_count=n)
    total.backward()
    m.zero_grad(set_to_none=True)


for i in range(2):
    step(i)
torch.cuda.synchronize()

if args.graph:
    from mvlt_b200.graph import GraphedStep
    from mvlt_b200.optim import AdamW, param_groups_no_decay
    opt = AdamW(param_groups_no_decay(m, 0.01), lr=1e-4)
    gs = GraphedStep(m, opt, mlm_capacity=-(-int(n * 1.1) // 128) * 128, warmup=1)
    xm = torch.empty_like(b["images"])
    masking.apply_grid_mask(b["images"], masking.grid_mask_batch(seeds + 1), out=xm)

    def gstep():
        return gs(xm, b["input_ids"], mlm_labels=b["mlm_labels"], itm_labels=b["itm_labels"], target_images=b["images"])
    gstep()                      # eager warm-up of the graph-mode code path
    _lib.GEMM_LOG = []           # descriptors of the captured iteration, in node order
    gstep()                      # capture + first replay
    log, _lib.GEMM_LOG = _lib.GEMM_LOG, None
    gstep()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    gstep()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(log, open("gpurun_out/gemm_desc_log.json", "w"))
    print(f"graph replay profiled: {gs.launches_per_replay()} kernel nodes, {len(log)} GEMM descriptors")
    sys.exit(0)

if args.dump_gemms:
    rows = []
    orig = kernels.gemm

    def traced(a, bb, out, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig(a, bb, out, **kw)
        e1.record()
        rows.append((a, bb, out, kw, e0, e1))
        return r
    kernels.gemm = traced
    import mvlt_b200.engine as E
    import mvlt_b200.t2i as T
    E.k.gemm = traced
    step(1)
    torch.cuda.synchronize()
    out = []
    for a, bb, o, kw, e0, e1 in rows:
        M, K = a.shape[-2], a.shape[-1]
        N = bb.shape[-2]
        batch = 1
        for d in o.shape[:-2]:
            batch *= d
        ms = e0.elapsed_time(e1)
        fl = 2.0 * M * N * K * batch
        byt = (M * K + N * K) * 2 * batch + M * N * batch * o.element_size()
        out.append(dict(M=M, N=N, K=K, batch=batch, a_mn=int(a.stride(-1) != 1), b_mn=int(bb.stride(-1) != 1),
                        out=str(o.dtype)[6:], atomic=bool(kw.get("atomic_add")), split=kw.get("split_k", 0),
                        act=kw.get("act", 0), res=kw.get("residual") is not None, ms=round(ms, 4),
                        tflops=round(fl / ms / 1e9, 1), gbs=round(byt / ms / 1e6, 1)))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/gemm_launches.json", "w"))
    tot = sum(r["ms"] for r in out)
    print(f"{len(out)} GEMM launches, {tot:.2f} ms")
    for r in sorted(out, key=lambda r: -r["ms"])[:40]:
        print(r)
else:
    _lib.GEMM_LOG = []          # one (flops, bytes, description) per GEMM launch of the profiled step, in launch order
    torch.cuda.profiler.start()
    step(1)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(_lib.GEMM_LOG, open("gpurun_out/gemm_desc_log.json", "w"))
