"""Fixed cost of one launch: back-to-back tiny GEMMs / LayerNorms / alternating, CUDA-event timed (warm)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvlt_b200 import kernels as k
BF16, F32 = torch.bfloat16, torch.float32
dev = "cuda"


def timeit(fn, n=200):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for (M, N, K) in [(128, 64, 64), (148 * 128, 64, 64), (148 * 128, 256, 64), (24576, 320, 640), (8192, 64, 4096)]:
    a = torch.randn((M, K), device=dev).to(BF16)
    b = torch.randn((N, K), device=dev).to(BF16)
    o = torch.empty((M, N), device=dev, dtype=BF16)
    print(f"gemm {M}x{N}x{K}: {timeit(lambda: k.gemm(a, b, o)):7.2f} us per launch (back to back)")
x = torch.randn((4096, 64), device=dev)
g = torch.ones(64, device=dev)
y = torch.empty((4096, 64), device=dev, dtype=BF16)
print(f"layernorm 4096x64: {timeit(lambda: k.layernorm_fwd(x, g, g, y, 1e-6, 4096, 64)):7.2f} us per launch")
a = torch.randn((128, 64), device=dev).to(BF16)
b = torch.randn((64, 64), device=dev).to(BF16)
o = torch.empty((128, 64), device=dev, dtype=BF16)


def alt():
    k.gemm(a, b, o)
    k.layernorm_fwd(x, g, g, y, 1e-6, 4096, 64)


print(f"gemm + layernorm alternating: {timeit(alt):7.2f} us per pair")
