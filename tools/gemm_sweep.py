"""Micro-benchmarks of mvlt_gemm over shape / epilogue variants (CUDA-event timing, L2 flushed between runs)."""
import sys, os, json
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


def run(M, N, K, out_dtype=BF16, mode="plain", block_n=0, b_mn=False):
    a = torch.randn((M, K), device=dev).to(BF16)
    b = (torch.randn((K, N), device=dev).to(BF16).t() if b_mn else torch.randn((N, K), device=dev).to(BF16))
    out = torch.empty((M, N), dtype=out_dtype, device=dev)
    kw = {}
    extra_bytes = 0
    if mode == "res":
        kw["residual"] = torch.randn((M, N), device=dev); extra_bytes = M * N * 4
    elif mode == "mul":
        kw["aux"] = torch.randn((M, N), device=dev).to(BF16); kw["act"] = k.ACT_MUL_AUX; extra_bytes = M * N * 2
    elif mode == "gelu":
        kw["act"] = k.ACT_GELU_SAVE_GRAD; kw["preact_out"] = torch.empty((M, N), dtype=BF16, device=dev); extra_bytes = M * N * 2
    elif mode == "softmax":
        kw["act"] = k.ACT_SOFTMAX; kw["alpha"] = 0.125
    ms = bench(lambda: k.gemm(a, b, out, block_n=block_n, **kw))
    byt = (M * K + N * K) * 2 + M * N * out.element_size() + extra_bytes
    tiles = ((M + 127) // 128) * ((N + (block_n or min(N, 256)) - 1) // (block_n or min(N, 256)))
    print(f"M={M:7d} N={N:5d} K={K:5d} {str(out_dtype)[6:]:8s} {mode:7s} bn={block_n:3d} bmn={int(b_mn)} : {ms*1e3:8.1f} us  "
          f"{byt/ms/1e6:7.0f} GB/s  {2.0*M*N*K/ms/1e9:7.1f} TF  us/tile/SM={ms*1e3/max(tiles/148,1):6.2f}", flush=True)


M = 540672
for N in (64, 128, 256, 512):
    run(M, N, 64)
for N in (64, 128, 256):
    run(M, N, 64, F32)
run(M, 64, 64, F32, "res")
run(M, 64, 512, F32, "res")
run(M, 512, 64, BF16, "mul", b_mn=True)
run(M, 512, 64, BF16, "mul", block_n=128, b_mn=True)
run(M, 512, 64, BF16, "gelu")
run(M, 512, 64, BF16, "gelu", block_n=128)
run(M, 256, 64, block_n=128)
run(M, 256, 64, block_n=64)
run(M, 512, 64, block_n=128)
run(147456, 1024, 128)
run(49152, 1280, 320)
run(24576, 2048, 512)
run(24576, 2048, 512, block_n=128)
run(8192, 8192, 8192)
run(8192, 8192, 8192, block_n=128)
run(16384, 30522, 768)
