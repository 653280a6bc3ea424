"""dW = dy^T x split-K GEMMs of the PVLT-tiny step, with and without the fused bias row-sum."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvlt_b200 import kernels as k
BF16, F32 = torch.bfloat16, torch.float32
dev = "cuda"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def bench(fn, iters=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


for (tok, co, ci, split) in [(540672, 512, 64, 74), (540672, 64, 512, 148), (540672, 64, 64, 296), (147456, 128, 1024, 74),
                             (147456, 1024, 128, 37), (147456, 128, 128, 296), (49152, 320, 1280, 20), (49152, 1280, 320, 15),
                             (24576, 2048, 512, 10), (24576, 512, 2048, 10), (24576, 512, 512, 37)]:
    dy = torch.randn((tok, co), device=dev).to(BF16)
    x = torch.randn((tok, ci), device=dev).to(BF16)
    dw = torch.zeros((co, ci), device=dev, dtype=F32)
    db = torch.zeros((co,), device=dev, dtype=F32)
    byt = tok * (co + ci) * 2
    for bn in (0, 64, 128):
        if bn and ci % bn:
            continue
        t0 = bench(lambda: k.gemm(dy.t(), x.t(), dw, atomic_add=True, split_k=split, block_n=bn))
        t1 = bench(lambda: k.gemm(dy.t(), x.t(), dw, atomic_add=True, split_k=split, block_n=bn, rowsum=db))
        print(f"tok={tok:7d} co={co:5d} ci={ci:5d} split={split:4d} bn={bn:3d}: plain {t0*1e3:7.1f} us ({byt/t0/1e6:5.0f} GB/s)   "
              f"+rowsum {t1*1e3:7.1f} us ({byt/t1/1e6:5.0f} GB/s)", flush=True)
