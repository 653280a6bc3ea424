"""Launch heuristics shared by the hand-scheduled engine modules."""
from functools import lru_cache


def pick_block_n(N, b_mn=True):
    """Mirror of pick_block_n() in csrc/gemm_tcgen05.cu (column-tile width the kernel will choose)."""
    step = 64 if b_mn else 32
    if N <= 256:
        return (N + step - 1) // step * step
    for bn in range(256, 127, -step):
        if N % bn == 0:
            return bn
    if not b_mn:
        for bn in range(224, 95, -32):
            if N % bn == 0:
                return bn
    return 128 if b_mn else 256


@lru_cache(maxsize=4096)
def split_k(M, N, K, b_mn=True):
    """split-K factor for GEMMs with a small M x N output and a long K (dW = dy^T x, the vocabulary dH): the persistent
    grid runs ceil(tiles * s / 148) waves of ceil(kb / s) k-blocks each (+ ~8 k-blocks worth of per-tile prologue /
    epilogue); pick the s that minimises that, i.e. fill the SMs without leaving a nearly empty last wave."""
    bn = pick_block_n(N, b_mn)
    tiles = ((M + 127) // 128) * ((N + bn - 1) // bn)
    kb = (K + 63) // 64
    best, best_cost = 1, None
    for s in range(1, max(1, kb // 4) + 1):
        waves = (tiles * s + 147) // 148
        cost = waves * ((kb + s - 1) // s + 8)
        if best_cost is None or cost < best_cost:
            best, best_cost = s, cost
        if tiles * s > 8 * 148:
            break
    return best
