"""Times the fused MLP kernels against the two-GEMM path at the PVLT-tiny stage-1/2 shapes (B = 128), CUDA events, cold L2
between launches is not forced: the operands (> 300 MB) exceed the 126 MB L2.   python tools/mlp_bench.py [--bwd]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvlt_b200 import kernels as k  # noqa: E402

BF16, F32 = torch.bfloat16, torch.float32
ap = argparse.ArgumentParser()
ap.add_argument("--bwd", action="store_true")
ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()


def timeit(fn, n):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for M, C, HD in ((540672, 64, 512), (147456, 128, 1024)):
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn((M, C), generator=g, device="cuda").to(BF16)
    w1 = (torch.randn((HD, C), generator=g, device="cuda") * C ** -0.5).to(BF16)
    w2 = (torch.randn((C, HD), generator=g, device="cuda") * HD ** -0.5).to(BF16)
    b1 = torch.randn((HD,), generator=g, device="cuda") * 0.3
    b2 = torch.randn((C,), generator=g, device="cuda") * 0.3
    res = torch.randn((M, C), generator=g, device="cuda")
    out = torch.empty_like(res)
    act = torch.empty((M, HD), dtype=BF16, device="cuda")
    hpre = torch.empty((M, HD), dtype=BF16, device="cuda")
    t_f = timeit(lambda: k.mlp_fwd(x, w1, b1, w2, b2, res, out), a.iters)

    def two_gemm(save):
        k.gemm(x, w1, act, bias=b1, act=k.ACT_GELU_SAVE_GRAD if save else k.ACT_GELU, preact_out=hpre if save else None)
        k.gemm(act, w2, out, bias=b2, residual=res)
    t_i = timeit(lambda: two_gemm(False), a.iters)
    t_t = timeit(lambda: two_gemm(True), a.iters)
    alg = M * C * 10 / 1e9
    print(f"M={M} C={C} HD={HD}: fused fwd {t_f:.1f} us ({alg / t_f * 1e6:.0f} GB/s algorithmic, {2 * 2 * M * C * HD / t_f / 1e6:.0f} TF/s) | "
          f"two GEMMs inference {t_i:.1f} us, training (act + gelu') {t_t:.1f} us", flush=True)
    if a.bwd:
        dy = torch.randn((M, C), generator=g, device="cuda").to(BF16)
        dh = torch.empty((M, HD), dtype=BF16, device="cuda")
        dw1 = torch.zeros((HD, C), device="cuda")
        dw2 = torch.zeros((C, HD), device="cuda")
        db1 = torch.zeros((HD,), device="cuda")
        t_b = timeit(lambda: k.mlp_bwd(x, dy, w1, b1, w2, dh, dw1, dw2, db1), a.iters)
        print(f"   fused bwd (recompute, dh', dW1, dW2, db1) {t_b:.1f} us", flush=True)
