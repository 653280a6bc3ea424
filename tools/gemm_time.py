"""Times one mvlt_gemm configuration with CUDA events and checks it against torch fp32 on a row sample:
python tools/gemm_time.py M N K mode   (mode: plain | res | gelu | mul)"""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvlt_b200 import kernels as k

M, N, K = (int(x) for x in sys.argv[1:4])
mode = sys.argv[4] if len(sys.argv) > 4 else "plain"
g = torch.Generator(device="cuda").manual_seed(0)
a = torch.randn((M, K), generator=g, device="cuda").to(torch.bfloat16)
b = (torch.randn((N, K), generator=g, device="cuda") * K ** -0.5).to(torch.bfloat16)
bias = torch.randn((N,), generator=g, device="cuda")
f32 = mode == "res"
out = torch.empty((M, N), dtype=torch.float32 if f32 else torch.bfloat16, device="cuda")
kw = dict(bias=bias)
if mode == "res":
    kw["residual"] = torch.randn((M, N), generator=g, device="cuda")
elif mode == "mul":
    kw["aux"] = torch.randn((M, N), generator=g, device="cuda").to(torch.bfloat16)
    kw["act"] = k.ACT_MUL_AUX
elif mode == "gelu":
    kw["act"] = k.ACT_GELU_SAVE_GRAD
    kw["preact_out"] = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
for _ in range(3):
    k.gemm(a, b, out, **kw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 10
for _ in range(n):
    k.gemm(a, b, out, **kw)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / n * 1e3
rows = torch.cat([torch.arange(0, 300, device="cuda"), torch.arange(M - 300, M, device="cuda"), torch.randint(0, M, (400,), device="cuda")])
ref = a[rows].float() @ b.float().t() + bias
if mode == "res":
    ref = ref + kw["residual"][rows]
elif mode == "mul":
    ref = ref * kw["aux"][rows].float()
elif mode == "gelu":
    ref = torch.nn.functional.gelu(ref)
err = float((out[rows].float() - ref).norm() / ref.norm())
print(f"{us:.1f} us  {2.0 * M * N * K / us / 1e6:.0f} TF/s  rel err {err:.2e}")
