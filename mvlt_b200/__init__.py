"""mvlt_b200: B200-native (sm_100a) implementation of the MVLT / PVLT data-parallel hot path.

Public surface mirrors the reference (GewelsJI/MVLT): ``mvlt_b200.libs.pvlt`` (model entry points),
``mvlt_b200.libs.vl_heads``, ``mvlt_b200.libs.vl_scores``, ``mvlt_b200.masking`` (grid masking) and the
train/eval loops in the top-level ``engine_grid_masking.py``; ``hubconf.py`` exports the ``pvlt_*`` entry points.
Compute lives in ``mvlt_b200/csrc`` (hand-written CUDA behind the C-ABI of ``include/mvlt_b200.h``).
"""
__version__ = "0.1.0"

_MODELS = ("pvlt_tiny", "pvlt_small", "pvlt_medium", "pvlt_large")


def create_model(name, pretrained=False, **kwargs):
    """Stand-in for ``timm.create_model`` as main_vl.py:259-270 calls it (timm is not in this image):
    drops ``None`` kwargs and injects ``in_chans=3`` exactly like timm 0.3.2's factory."""
    from .libs import pvlt
    if name not in _MODELS:
        raise ValueError(f"unknown model {name!r}; available: {_MODELS}")
    kwargs = {k: v for k, v in kwargs.items() if v is not None}
    kwargs.setdefault("in_chans", 3)
    return getattr(pvlt, name)(pretrained=pretrained, **kwargs)
