"""In-tree nvcc build of the C-ABI library ``mvlt_b200/lib/libmvlt_b200.so`` (sm_100a only).

The .so is git-ignored but travels to the GPU box with the gpurun snapshot. Nothing is JIT-compiled at
import time: if the library is missing on a GPU box the package fails loudly (see ``_lib.py``).
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
LIBDIR = ROOT / "lib"
OBJDIR = ROOT / "lib" / "obj"
LIBPATH = LIBDIR / "libmvlt_b200.so"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
          "--expt-relaxed-constexpr"]


def _sources():
    return sorted(CSRC.glob("*.cu"))


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(ARCH_FLAGS + COMMON).encode())
    return h.hexdigest()


def _compile_one(src: Path, deps_digest: str, verbose: bool) -> Path:
    obj = OBJDIR / (src.stem + ".o")
    stamp = OBJDIR / (src.stem + ".stamp")
    want = hashlib.sha256((deps_digest + src.read_text()).encode()).hexdigest()
    if obj.exists() and stamp.exists() and stamp.read_text() == want:
        return obj
    cmd = [NVCC, *ARCH_FLAGS, *COMMON, "-I", str(CSRC), "-c", str(src), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"nvcc failed for {src.name}")
    if verbose:
        (OBJDIR / (src.stem + ".ptxas.log")).write_text(r.stderr)
    stamp.write_text(want)
    return obj


def build(verbose: bool = True, force: bool = False) -> Path:
    """Compile every ``csrc/*.cu`` for sm_100a and link ``libmvlt_b200.so``. Incremental per source file."""
    LIBDIR.mkdir(exist_ok=True)
    OBJDIR.mkdir(exist_ok=True)
    srcs = _sources()
    headers = sorted(list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")))
    hdig = _digest(headers)
    if force:
        for f in OBJDIR.glob("*.stamp"):
            f.unlink()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile_one(s, hdig, verbose), srcs))
    link_stamp = OBJDIR / "link.stamp"
    want = _digest(objs)
    if not (LIBPATH.exists() and link_stamp.exists() and link_stamp.read_text() == want):
        cmd = [NVCC, *ARCH_FLAGS, "-shared", "-o", str(LIBPATH), *map(str, objs), "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
        link_stamp.write_text(want)
    return LIBPATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
