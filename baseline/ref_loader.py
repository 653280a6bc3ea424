"""Loads the LIVE, unmodified reference model (libs/pvlt.py) from the git-ignored ``baseline/_ref/`` staged by
``tools/stage_reference.py``. Used only by bench.py's reference legs and by tests that cross-check the oracle; never by the
product package (mvlt_b200 has no dependency on it)."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_mod = None


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "libs", "pvlt.py"))


def load():
    """Returns the reference's ``libs.pvlt`` module. transformers probes for a real ``timm`` at import time, so it is
    imported BEFORE the stub package becomes importable."""
    global _mod
    if _mod is not None:
        return _mod
    if not available():
        raise RuntimeError("baseline/_ref is not staged: run `python tools/stage_reference.py` in the build container")
    from transformers.models.bert.modeling_bert import BertConfig, BertEmbeddings  # noqa: F401
    from transformers import BartConfig, BartForConditionalGeneration, BartModel   # noqa: F401
    BertConfig.from_pretrained = classmethod(lambda cls, *a, **k: cls())   # hub unreachable; defaults == bert-base-uncased
    for name in [n for n in sys.modules if n == "libs" or n.startswith("libs.")]:
        del sys.modules[name]
    sys.path.insert(0, REF_DIR)
    try:
        from libs import pvlt as ref_pvlt
    finally:
        sys.path.remove(REF_DIR)
    _mod = ref_pvlt
    return ref_pvlt


def build_model(name="pvlt_tiny", loss_type=None, drop_path_rate=0.0, state_dict=None):
    """The reference model as main_vl.py:259-270 builds it (timm.create_model strips None kwargs and injects in_chans)."""
    loss_type = dict(loss_type or {"itm": 1, "mlm": 1, "t2i": 1, "cls": 0})
    m = getattr(load(), name)(pretrained=True, token_hidden_size=768, num_text_tokens=128, loss_type=loss_type,
                              pretrained_pth="", num_classes=1000, in_chans=3, drop_rate=0.0, drop_path_rate=drop_path_rate)
    if state_dict is not None:
        m.load_state_dict(state_dict)
    return m
