"""LayerNorm backward at the PVLT-tiny stage shapes (B = 128) under the launcher's tuning knobs (environment, read once
per process): python tools/ln_sweep.py   -> one line per shape; run once per knob setting."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvlt_b200 import kernels as k
dev = "cuda"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def bench(fn, iters=7):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


tag = " ".join(f"{kk[8:]}={v}" for kk, v in sorted(os.environ.items()) if kk.startswith("MVLT_LN_"))
tot = 0.0
out = []
for (rows, C) in [(540672, 64), (147456, 128), (49152, 320), (24576, 512)]:
    for dy_dt in (torch.bfloat16, torch.float32):
        dy = torch.randn((rows, C), device=dev).to(dy_dt)
        x = torch.randn((rows, C), device=dev)
        mean, rstd = torch.zeros(rows, device=dev), torch.ones(rows, device=dev)
        gamma = torch.ones(C, device=dev)
        dx, dx_add = torch.empty((rows, C), device=dev), torch.randn((rows, C), device=dev)
        dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
        dx16 = torch.empty((rows, C), device=dev, dtype=torch.bfloat16)
        t = bench(lambda: k.layernorm_bwd(dy, x, mean, rstd, gamma, dx, rows, C, dx_add=dx_add, dgamma=dg, dbeta=db, dx_bf16=dx16))
        nbytes = rows * C * (dy.element_size() + 4 + 4 + 4 + 2)
        tot += t
        out.append(f"C={C:3d} dy={'bf16' if dy_dt == torch.bfloat16 else 'f32 '} {t * 1e3:6.1f}us {nbytes / t / 1e6:5.0f}GB/s")
print(f"[{tag or 'defaults'}] total {tot * 1e3:.0f}us | " + " | ".join(out), flush=True)
