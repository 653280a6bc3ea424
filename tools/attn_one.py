"""Launch the fused SR-attention forward at one PVLT-tiny stage shape (for ncu): python tools/attn_one.py [stage 1-4] [B]
   launch 1 = with the probability store (training), launch 2 = without (inference / retrieval)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvlt_b200 import kernels as k

stage = int(sys.argv[1]) if len(sys.argv) > 1 else 1
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
H, N = [(1, 4224), (2, 1152), (5, 384), (8, 192)][stage - 1]
Nk, C = 192, H * 64
q = torch.randn((B * N, C), device="cuda").to(torch.bfloat16)
kv = torch.randn((B * Nk, 2 * C), device="cuda").to(torch.bfloat16)
o = torch.empty_like(q)
P = torch.empty((B, H, N, Nk), device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    k.sr_attention_fwd(q, kv, o, P, B, N, Nk, H, 0.125)
    k.sr_attention_fwd(q, kv, o, None, B, N, Nk, H, 0.125)
torch.cuda.synchronize()
