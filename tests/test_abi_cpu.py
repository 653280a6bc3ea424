"""The C-ABI library loads without a GPU and exports every symbol include/mvlt_b200.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "mvlt_b200.h")).read()
    return sorted(set(re.findall(r"\b(mvlt_[a-z0-9_]+)\s*\(", hdr)))


def test_library_builds_and_exports_header_symbols():
    from mvlt_b200 import build
    lib = ctypes.CDLL(str(build.build(verbose=False)))
    names = _declared()
    assert len(names) >= 45
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mvlt_b200.h but not exported"
    lib.mvlt_last_error.restype = ctypes.c_char_p
    assert lib.mvlt_abi_version() == 1


def test_header_is_in_sync_with_sources():
    import subprocess
    import sys
    before = open(os.path.join(ROOT, "include", "mvlt_b200.h")).read()
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "gen_header.py")], stdout=subprocess.DEVNULL)
    assert open(os.path.join(ROOT, "include", "mvlt_b200.h")).read() == before


def test_no_cpu_fallback_error_paths():
    """Product ops fail loudly on CPU tensors instead of falling back."""
    import pytest
    import torch
    from mvlt_b200 import kernels
    from mvlt_b200._lib import MvltError
    a = torch.zeros((8, 8), dtype=torch.bfloat16)
    with pytest.raises(MvltError):
        kernels.gemm(a, a, torch.zeros((8, 8)))
    import mvlt_b200
    m = mvlt_b200.create_model("pvlt_tiny", pretrained=False, token_hidden_size=768, num_text_tokens=128,
                               loss_type={"itm": 1, "mlm": 0, "t2i": 0, "cls": 0}, pretrained_pth="")
    with pytest.raises(MvltError):
        m(torch.zeros(1, 3, 256, 256), torch.zeros(1, 128, dtype=torch.long))


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "mvlt_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), f"{f} references the oracle"
