"""Attention-product GEMMs of one PVLT-tiny block per stage (fused softmax fwd / bwd epilogues and the plain ones)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvlt_b200 import kernels as k
BF16 = torch.bfloat16
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


B, D, Nk = 128, 64, 192
for (H, N) in [(1, 4224), (2, 1152), (5, 384), (8, 192)]:
    C = H * D
    q = torch.randn((B * N, C), device=dev).to(BF16)
    kv = torch.randn((B * Nk, 2 * C), device=dev).to(BF16)
    do = torch.randn((B * N, C), device=dev).to(BF16)
    q4 = q.view(B, N, H, D).permute(0, 2, 1, 3)
    do4 = do.view(B, N, H, D).permute(0, 2, 1, 3)
    k4 = kv.view(B, Nk, 2, H, D)[:, :, 0].permute(0, 2, 1, 3)
    v4 = kv.view(B, Nk, 2, H, D)[:, :, 1].permute(0, 2, 1, 3)
    P = torch.empty((B, H, N, Nk), device=dev, dtype=BF16)
    dS = torch.empty_like(P)
    o = torch.empty((B * N, C), device=dev, dtype=BF16)
    o4 = o.view(B, N, H, D).permute(0, 2, 1, 3)
    pb = P.numel() * 2
    qb = q.numel() * 2
    t = bench(lambda: k.gemm(q4, k4, P, alpha=0.125, act=k.ACT_SOFTMAX))
    print(f"H={H} N={N:5d}  QK^T+softmax   {t*1e3:7.1f} us  {(pb + qb)/t/1e6:6.0f} GB/s")
    t = bench(lambda: k.gemm(P, v4.transpose(-1, -2), o4))
    print(f"H={H} N={N:5d}  PV             {t*1e3:7.1f} us  {(pb + qb)/t/1e6:6.0f} GB/s")
    kvb = kv.numel() * 2
    if "--fused" in sys.argv:   # one-kernel forward (csrc/attn_tcgen05.cu): with / without the probability store
        t = bench(lambda: k.sr_attention_fwd(q, kv, o, P, B, N, Nk, H, 0.125))
        print(f"H={H} N={N:5d}  fused fwd (+P) {t*1e3:7.1f} us  {(pb + 2 * qb + kvb)/t/1e6:6.0f} GB/s")
        t = bench(lambda: k.sr_attention_fwd(q, kv, o, None, B, N, Nk, H, 0.125))
        print(f"H={H} N={N:5d}  fused fwd      {t*1e3:7.1f} us  {(2 * qb + kvb)/t/1e6:6.0f} GB/s")
    t = bench(lambda: k.gemm(do4, v4, dS, alpha=0.125, act=k.ACT_SOFTMAX_BWD, aux=P))
    print(f"H={H} N={N:5d}  dP+softmax_bwd {t*1e3:7.1f} us  {(2 * pb + qb)/t/1e6:6.0f} GB/s")
    t = bench(lambda: k.gemm(dS, k4.transpose(-1, -2), o4))
    print(f"H={H} N={N:5d}  dQ = dS K      {t*1e3:7.1f} us  {(pb + qb)/t/1e6:6.0f} GB/s", flush=True)
