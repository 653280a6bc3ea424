"""Host-side utilities of the SURVEY 8f rows (no GPU): text pipeline (tokenizer wrapper, caption rows, ITM pair
sampling) and checkpoint compatibility with the reference's file formats."""
import os
import random

import pytest
import torch

from mvlt_b200 import text as T
from oracle import pvlt_oracle as O

PRE = {"itm": 1, "mlm": 1, "t2i": 1, "cls": 0}
VOCAB = ["[PAD]"] + [f"[unused{i}]" for i in range(99)] + ["[UNK]", "[CLS]", "[SEP]", "[MASK]"] + \
        ["a", "black", "dress", "with", "long", "sleeve", "##s", "and", "pocket", "##less", ",", "."]


@pytest.fixture()
def tokenizer(tmp_path):
    p = tmp_path / "vocab.txt"
    p.write_text("\n".join(VOCAB) + "\n", encoding="utf-8")
    return T.load_tokenizer(str(p))


def test_encode_captions_layout_truncation_padding(tokenizer):
    assert (tokenizer.pad_token_id, tokenizer.cls_token_id, tokenizer.sep_token_id, tokenizer.mask_token_id) == (0, 101, 102, 103)
    caps = ["A black dress with long sleeves.", "pocketless dress , " * 40, "zzz"]
    ori, att, seg = T.encode_captions(tokenizer, caps, max_token_length=16)
    assert ori.shape == att.shape == seg.shape == (3, 16) and ori.dtype == torch.long
    v = {t: i for i, t in enumerate(VOCAB)}
    want0 = [101, v["a"], v["black"], v["dress"], v["with"], v["long"], v["sleeve"], v["##s"], v["."], 102]
    assert ori[0, :10].tolist() == want0 and ori[0, 10:].eq(0).all()
    assert att[0].tolist() == [1] * 10 + [0] * 6
    # fashion_gen.py:327-329: at most T-2 word pieces, [SEP] always present
    assert ori[1, 0] == 101 and ori[1, 15] == 102 and att[1].all()
    assert ori[1, 1:5].tolist() == [v["pocket"], v["##less"], v["dress"], v[","]]
    assert ori[2, :3].tolist() == [101, 100, 102]          # unknown word -> [UNK]
    assert seg.eq(0).all()


def test_itm_pair_sampling_follows_the_reference_stream():
    """fashion_gen.py:121-146 restated inline on the module-level RNG, against itm_text_index on a seeded Random."""
    size = 1000
    for seed in (0, 1, 12345):
        random.seed(seed)
        want = []
        for index in range(0, size, 37):
            if random.random() <= 0.5:
                want.append((index, 1))
            else:
                j = index + random.randint(50, size // 2)
                if j > size - 1:
                    j -= size
                want.append((j, 0))
        idx, lab = T.itm_pairs(seed, list(range(0, size, 37)), size)
        assert idx == [w[0] for w in want] and lab.view(-1).tolist() == [w[1] for w in want]
        assert all(0 <= j < size for j in idx)
        neg = [(i, j) for i, (j, l) in zip(range(0, size, 37), want) if l == 0]
        assert all(50 <= (j - i) % size <= size // 2 for i, j in neg)


def _model():
    import mvlt_b200
    return mvlt_b200.create_model("pvlt_tiny", pretrained=True, num_classes=1000, drop_rate=0.0, drop_path_rate=0.0,
                                  drop_block_rate=None, token_hidden_size=768, num_text_tokens=128, loss_type=dict(PRE),
                                  pretrained_pth="")


def test_load_checkpoint_reference_formats(tmp_path):
    from mvlt_b200.utils import load_checkpoint
    sd = O.make_state_dict("pvlt_tiny", PRE, seed=3)      # the reference's key names / shapes (pinned by the golden script)
    m = _model()
    # (1) bare state_dict file, DDP-prefixed, with a foreign ImageNet classifier row that must be dropped
    f1 = tmp_path / "checkpoint_retrieval.pth"
    bare = {"module." + k: v for k, v in sd.items()}
    bare["module.head.weight"] = torch.zeros(1000, 512)
    torch.save(bare, f1)
    missing, unexpected, nxt = load_checkpoint(m, str(f1))
    assert missing == [] and unexpected == [] and nxt is None
    for k, v in m.state_dict().items():
        assert torch.equal(v, sd[k]), k
    assert m.mlm_head.mlm_decoder.weight.data_ptr() == m.text_embeddings.word_embeddings.weight.data_ptr()
    # (2) training checkpoint (main_vl.py:327-346): model + torch.optim.AdamW state + epoch
    from mvlt_b200.optim import AdamW, param_groups_no_decay
    ref_opt = torch.optim.AdamW(param_groups_no_decay(m, 0.05), lr=1e-3)
    for p in m.parameters():
        p.grad = torch.full_like(p, 1e-3)
    ref_opt.step()
    after = {k: v.clone() for k, v in m.state_dict().items()}
    f2 = tmp_path / "checkpoint.pth"
    torch.save({"model": m.state_dict(), "optimizer": ref_opt.state_dict(), "lr_scheduler": {}, "epoch": 7}, f2)
    m2 = _model()
    opt2 = AdamW(param_groups_no_decay(m2, 0.05), lr=5e-4)
    missing, unexpected, nxt = load_checkpoint(m2, str(f2), optimizer=opt2)
    assert missing == [] and unexpected == [] and nxt == 8
    for k, v in m2.state_dict().items():
        assert torch.equal(v, after[k]), k
    assert opt2.param_groups[0]["lr"] == 1e-3 and opt2.param_groups[1]["weight_decay"] == 0.0
    p0 = opt2.param_groups[0]["params"][0]
    st = opt2.state[p0]
    assert int(st["step"]) == 1 and st["exp_avg"].shape == p0.shape and float(st["exp_avg"].abs().max()) > 0


REF_VOCAB = "/root/reference/preweights/bert-base-uncased-vocab.txt"


@pytest.mark.skipif(not os.path.exists(REF_VOCAB), reason="needs the vocabulary the reference vendors (build container only)")
def test_encode_captions_matches_reference_text_process_golden():
    """tests/golden/text_golden.npz was recorded from the reference's own text_process (make_golden.py:golden_text)."""
    import numpy as np
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "text_golden.npz"))
    tok = T.load_tokenizer(REF_VOCAB)
    assert len(tok) == 30522
    caps = [str(c) for c in g["captions"]]
    row = 0
    for L in (int(v) for v in g["lengths"]):
        ori, att, seg = T.encode_captions(tok, caps, max_token_length=L)
        n = len(caps)
        assert (ori.numpy() == g["ori"][row:row + n, :L]).all()
        assert (att.numpy() == g["att"][row:row + n, :L]).all()
        assert (seg.numpy() == g["seg"][row:row + n, :L]).all()
        row += n


def test_bench_hbm_kernel_table_and_byte_accounting():
    """bench.py's per-kernel HBM fractions: the wrappers count algorithmic bytes only while _lib.BYTES is a dict."""
    import importlib.util
    from mvlt_b200 import _lib
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(__file__)), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert _lib.BYTES is None
    _lib.account_bytes("layernorm_fwd", 10.0)          # inactive: must not create state
    assert _lib.BYTES is None
    _lib.BYTES = {}
    try:
        _lib.account_bytes("layernorm_fwd", 3.0e9)
        _lib.account_bytes("layernorm_fwd", 3.0e9)
        _lib.account_bytes("adamw_multi", 1.0e9)
        tab = bench.hbm_kernel_table(_lib.BYTES, {"layernorm_fwd": 2.0, "adamw_multi": 0.0, "gemm": 5.0},
                                     {"layernorm_fwd": 4}, 2, 6000.0)
    finally:
        _lib.BYTES = None
    assert set(tab) == {"layernorm_fwd"}                # kernels without a measured time are dropped
    row = tab["layernorm_fwd"]
    assert row["achieved_gbs"] == 3000.0 and row["frac_of_hbm_peak"] == 0.5
    assert row["ms_per_step"] == 1.0 and row["launches_per_step"] == 2.0 and row["algorithmic_gb_per_step"] == 3.0


def test_adamw_step_bytes_counts_the_bf16_copies():
    """bench.py's AdamW roofline: 28 B per parameter + 2 B per bf16 compute copy the launch rewrites."""
    import torch
    from mvlt_b200.optim import step_bytes
    a, b, c, d = (torch.nn.Parameter(torch.zeros(s)) for s in ((10, 4), (8, 2, 3, 3), (7,), (5, 5)))
    a._mvlt_shadow = (1, torch.zeros(40, dtype=torch.bfloat16), None, 0, 0, 0, 0)
    b._mvlt_shadow = (2, torch.zeros(144, dtype=torch.bfloat16), torch.zeros(144, dtype=torch.bfloat16), 8, 2, 9, 18)
    c._mvlt_shadow = (0, None, None, 0, 0, 0, 0)
    assert step_bytes([a, b, c, d]) == 40 * 30 + 144 * 32 + 7 * 28 + 25 * 28


def test_entrypoints_register_with_timm_and_hubconf(monkeypatch):
    """main_vl.py:16,259: ``from libs import pvlt`` registers pvlt_* with timm's registry; hubconf exports them."""
    import importlib
    import sys
    import types
    registered = {}
    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    registry = types.ModuleType("timm.models.registry")

    def register_model(fn):
        registered[fn.__name__] = fn
        return fn
    registry.register_model = register_model
    for name, mod in (("timm", timm), ("timm.models", models), ("timm.models.registry", registry)):
        monkeypatch.setitem(sys.modules, name, mod)
    import mvlt_b200.libs.pvlt as P
    importlib.reload(P)
    try:
        assert set(registered) == {"pvlt_tiny", "pvlt_small", "pvlt_medium", "pvlt_large"}
    finally:
        for name in ("timm", "timm.models", "timm.models.registry"):
            monkeypatch.delitem(sys.modules, name)
        importlib.reload(P)
    import hubconf
    for name in ("pvlt_tiny", "pvlt_small", "pvlt_medium", "pvlt_large"):
        assert callable(getattr(hubconf, name))
    import mvlt_b200
    m = mvlt_b200.create_model("pvlt_small", pretrained=False, drop_block_rate=None, token_hidden_size=768, num_text_tokens=128,
                               loss_type={"itm": 1, "mlm": 0, "t2i": 0, "cls": 1}, pretrained_pth="", drop_path_rate=0.1)
    assert len(m.block3) == 6      # depths [3, 4, 6, 3]
    with pytest.raises(TypeError):      # the reference constructor has no drop_block_rate either: timm strips None values
        mvlt_b200.create_model("pvlt_tiny", pretrained=False, drop_block_rate=0.1, token_hidden_size=768, num_text_tokens=128,
                               loss_type=dict(PRE), pretrained_pth="")


def test_validated_defaults_are_the_ones_shipped():
    """The switches the GPU runs validated: fused attention forward and backward on, programmatic dependent launch on."""
    import mvlt_b200.engine as E
    if "MVLT_FUSED_ATTN" not in os.environ:
        assert E.FUSED_ATTENTION is True
    if "MVLT_FUSED_ATTN_BWD" not in os.environ:
        assert E.FUSED_ATTENTION_BWD is True
    if "MVLT_FUSED_MLP" not in os.environ:
        assert E.FUSED_MLP is True
    if "MVLT_FUSED_MLP_TRAIN" not in os.environ:
        from mvlt_b200 import kernels
        assert kernels.MLP_FUSED_DIMS == (64, 128) and kernels.MLP_FUSED_BWD_DIMS == (64,)
    src = open(os.path.join(os.path.dirname(os.path.dirname(__file__)), "mvlt_b200", "csrc", "common.cuh")).read()
    assert "#define MVLT_PDL_DEFAULT 1" in src


def test_split_k_heuristic_invariants():
    """engine_util.split_k (dW = dy^T x and the vocabulary dH GEMMs): 1 <= s <= k-blocks / 4, more splits never leave the persistent
    grid with fewer than one wave of work per split, long-K / tiny-output shapes are split, full-grid shapes are not."""
    from mvlt_b200.engine_util import pick_block_n, split_k
    shapes = [(64, 64, 540672), (128, 128, 147456), (320, 320, 49152), (512, 2048, 24576), (30522, 768, 896), (896, 768, 30522),
              (192, 1728, 131072), (64, 48, 524288), (2, 768, 128), (147456, 128, 1024)]
    for M, N, K in shapes:
        s = split_k(M, N, K)
        kb = (K + 63) // 64
        assert 1 <= s <= max(1, kb // 4), (M, N, K, s)
        assert s == split_k(M, N, K)
    assert split_k(64, 64, 540672) > 32 and split_k(896, 768, 30522) > 1        # tiny outputs, long K: fill the 148 SMs
    assert split_k(147456, 128, 1024) == 1 and split_k(30522, 768, 896) == 1     # already >= several waves of tiles
    for n, want in ((64, 64), (192, 192), (320, 128), (512, 256), (768, 256), (1728, 192), (30522, 128)):
        assert pick_block_n(n) == want, (n, pick_block_n(n), want)
