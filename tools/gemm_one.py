"""Launch one mvlt_gemm configuration a few times (for ncu): python tools/gemm_one.py M N K mode [block_n] [f32]"""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvlt_b200 import kernels as k

M, N, K = (int(x) for x in sys.argv[1:4])
mode = sys.argv[4] if len(sys.argv) > 4 else "plain"
bn = int(sys.argv[5]) if len(sys.argv) > 5 else 0
f32 = len(sys.argv) > 6 and sys.argv[6] == "f32"
dev = "cuda"
a = torch.randn((M, K), device=dev).to(torch.bfloat16)
b = torch.randn((N, K), device=dev).to(torch.bfloat16)
out = torch.empty((M, N), dtype=torch.float32 if f32 else torch.bfloat16, device=dev)
kw = {}
if mode == "res":
    kw["residual"] = torch.randn((M, N), device=dev)
elif mode == "mul":
    kw["aux"] = torch.randn((M, N), device=dev).to(torch.bfloat16)
    kw["act"] = k.ACT_MUL_AUX
elif mode == "gelu":
    kw["act"] = k.ACT_GELU_SAVE_GRAD
    kw["preact_out"] = torch.empty((M, N), dtype=torch.bfloat16, device=dev)
for _ in range(3):
    k.gemm(a, b, out, block_n=bn, **kw)
torch.cuda.synchronize()
