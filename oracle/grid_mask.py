"""ORACLE (test infrastructure): random grid mask + pixel fill, CPU.

Follows /root/reference/mcloader/fashion_gen.py:225-254 (mask) and :176-177 (fill). Two independent
restatements: ``grid_c`` (oracle/grid_mask.c, compiled with gcc) and ``grid_py`` (pure Python MT19937);
the CPU tests check them against each other and against the golden vectors recorded from the reference.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_build" / "liboracle_mask.so"


def build() -> Path:
    src = _HERE / "grid_mask.c"
    _SO.parent.mkdir(exist_ok=True)
    if not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", str(_SO), str(src)])
    return _SO


def sample_seed(seed: int, sample_idx: int) -> int:
    """The harness-defined per-sample MT19937 seed: np.random.seed(seed*1000003 + idx)."""
    return (seed * 1000003 + sample_idx) & 0xFFFFFFFF


def grid_c(seed: int, size=(256, 256), patch: int = 16, ratio: float = 0.5) -> np.ndarray:
    lib = ctypes.CDLL(str(build()))
    lib.oracle_grid_mask.restype = ctypes.c_uint32
    nw, nh = size[0] // patch, size[1] // patch
    grid = np.zeros((nh, nw), dtype=np.uint8)
    lib.oracle_grid_mask(ctypes.c_uint32(seed), size[0], size[1], patch, ctypes.c_double(ratio),
                         grid.ctypes.data_as(ctypes.c_void_p))
    return grid


class _MT:
    def __init__(self, seed):
        mt = [seed & 0xFFFFFFFF]
        for i in range(1, 624):
            mt.append((1812433253 * (mt[-1] ^ (mt[-1] >> 30)) + i) & 0xFFFFFFFF)
        self.mt, self.idx = mt, 624

    def next(self):
        if self.idx >= 624:
            mt = self.mt
            for k in range(624):
                y = (mt[k] & 0x80000000) | (mt[(k + 1) % 624] & 0x7FFFFFFF)
                mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
            self.idx = 0
        y = self.mt[self.idx]
        self.idx += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF

    def interval(self, mx):
        if mx == 0:
            return 0
        mask = mx
        for s in (1, 2, 4, 8, 16):
            mask |= mask >> s
        while True:
            v = self.next() & mask
            if v <= mx:
                return v

    def shuffle(self, lst):
        for i in range(len(lst) - 1, 0, -1):
            j = self.interval(i)
            lst[i], lst[j] = lst[j], lst[i]


def grid_py(seed: int, size=(256, 256), patch: int = 16, ratio: float = 0.5) -> np.ndarray:
    nw, nh = size[0] // patch, size[1] // patch
    n = nw * nh
    nm = int(ratio * n)
    vals = [0] * (n - nm) + [1] * nm
    rng = _MT(seed)
    rng.shuffle(vals)
    grid = np.zeros((nh, nw), dtype=np.uint8)
    for r in range(nh):
        row = vals[r:r + nw]
        rng.shuffle(row)
        grid[r, :len(row)] = row
    return grid


def expand(grid: np.ndarray, patch: int = 16) -> np.ndarray:
    """[nh,nw] -> float64 [1, nh*patch, nw*patch], the array generate_grid_mask returns."""
    return np.kron(grid, np.ones((patch, patch)))[None].astype(np.float64)


def masked_fill(image: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """fashion_gen.py:176: image.clone().masked_fill_(mask.byte().bool(), 1e-6); image [3,S,S], mask [1,S,S]."""
    out = image.copy()
    out[np.broadcast_to(mask.astype(bool), image.shape)] = np.float32(1e-6)
    return out
