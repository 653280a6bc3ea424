"""Stages the UNMODIFIED reference model sources into the git-ignored ``baseline/_ref/`` so that the reference arm of
bench.py (``--impl reference``, ``cpu_baseline``, ``gpu_eager_reference``) can run the LIVE reference on the GPU box,
which has no /root/reference (gpurun ships ignored files; history stays source-only).

    python tools/stage_reference.py          # build container only; also called by __graft_entry__.build()

Copied verbatim (never edited, never committed): libs/{__init__,pvlt,vl_heads,vl_scores}.py -- the files SURVEY 8(a)
cites for the model. Third-party gaps are papered over exactly as SURVEY 8(c) describes, by files this script WRITES
(own code, not reference code): a minimal ``timm`` stub package (DropPath, to_2tuple, trunc_normal_, register_model,
_cfg -- the five names libs/pvlt.py:6-8 imports) and, at load time (baseline/ref_loader.py),
``BertConfig.from_pretrained -> BertConfig()`` (defaults == bert-base-uncased; the hub is unreachable).
"""
import hashlib
import os
import shutil
import sys

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = ["libs/__init__.py", "libs/pvlt.py", "libs/vl_heads.py", "libs/vl_scores.py"]

TIMM_LAYERS = '''"""Stub of the five timm 0.3.2 names libs/pvlt.py imports (timm is not installed here; SURVEY 8c gap 1)."""
import torch


class DropPath(torch.nn.Module):
    """Per-sample stochastic depth: x / keep * Bernoulli(keep) in training, identity otherwise."""

    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if not self.drop_prob or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
        return x.div(keep) * mask


def to_2tuple(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


trunc_normal_ = torch.nn.init.trunc_normal_
'''


def stage(verbose=True):
    if not os.path.isdir(REF):
        if verbose:
            print("stage_reference: /root/reference not present (GPU box): using what is already staged")
        return os.path.isdir(os.path.join(DST, "libs"))
    os.makedirs(os.path.join(DST, "libs"), exist_ok=True)
    manifest = []
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        shutil.copyfile(src, dst)
        manifest.append(f"{hashlib.sha256(open(src, 'rb').read()).hexdigest()}  {rel}")
    for pkg in ("timm", "timm/models"):
        os.makedirs(os.path.join(DST, pkg), exist_ok=True)
        open(os.path.join(DST, pkg, "__init__.py"), "w").write("")
    open(os.path.join(DST, "timm/models/layers.py"), "w").write(TIMM_LAYERS)
    open(os.path.join(DST, "timm/models/registry.py"), "w").write("def register_model(fn):\n    return fn\n")
    open(os.path.join(DST, "timm/models/vision_transformer.py"), "w").write("def _cfg(**kwargs):\n    return dict(kwargs)\n")
    open(os.path.join(DST, "MANIFEST.sha256"), "w").write("\n".join(manifest) + "\n")
    if verbose:
        print("staged the reference model sources into", DST)
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
