"""Generates the committed golden fixtures by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py

Writes
  tests/golden/grid_mask_golden.npz   reference generate_grid_mask under np.random.seed(s) for many seeds
  tests/golden/token_mask_golden.npz  reference random_masking_features (BERT word-piece masking) under random.seed(s)
  tests/golden/pvlt_tiny_golden.npz   reference PVLT-tiny (libs/pvlt.py) outputs, losses and gradient summaries
                                      for oracle.make_state_dict(seed) weights and oracle.make_inputs batches
  tests/golden/text_golden.npz        reference text_process (tokenize / truncate / pad) on a few captions, vendored vocabulary
  tests/golden/pvlt_small_golden.npz  the same for pvlt_small (depths [3,4,6,3]; BASELINE configs[4] stand-in), batch 1

Third-party gaps papered over exactly as SURVEY 8c describes: a ~30-line timm stub and
BertConfig.from_pretrained -> BertConfig() (defaults == bert-base-uncased).
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def install_timm_stub():
    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    layers = types.ModuleType("timm.models.layers")
    registry = types.ModuleType("timm.models.registry")
    vit = types.ModuleType("timm.models.vision_transformer")

    class DropPath(torch.nn.Module):
        def __init__(self, p=0.0):
            super().__init__()
            self.p = p

        def forward(self, x):
            if self.p == 0.0 or not self.training:
                return x
            keep = 1 - self.p
            mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
            return x.div(keep) * mask

    layers.DropPath = DropPath
    layers.to_2tuple = lambda v: v if isinstance(v, tuple) else (v, v)
    layers.trunc_normal_ = torch.nn.init.trunc_normal_
    registry.register_model = lambda f: f
    vit._cfg = lambda **kw: dict(kw)
    for name, mod in [("timm", timm), ("timm.models", models), ("timm.models.layers", layers),
                      ("timm.models.registry", registry), ("timm.models.vision_transformer", vit)]:
        sys.modules[name] = mod


def import_reference():
    # transformers probes for a real `timm` at import time, so import it BEFORE the stub is installed
    from transformers.models.bert.modeling_bert import BertConfig, BertEmbeddings  # noqa: F401
    from transformers import BartConfig, BartForConditionalGeneration, BartModel   # noqa: F401
    install_timm_stub()
    BertConfig.from_pretrained = classmethod(lambda cls, *a, **k: cls())
    sys.path.insert(0, REF)
    from libs import pvlt as ref_pvlt
    return ref_pvlt


def golden_grid_masks():
    """Calls the reference's own method (fashion_gen.py:225-254) with the legacy global numpy RNG."""
    src = open(os.path.join(REF, "mcloader", "fashion_gen.py")).read()
    start = src.index("    def generate_grid_mask")
    end = src.index("    def generate_square_mask")
    ns = {"np": np}
    exec("class _Holder:\n" + src[start:end], ns)          # executes the reference text in place, nothing copied
    fn = ns["_Holder"].generate_grid_mask
    seeds = list(range(64)) + [12345, 2 ** 31 - 1, 2 ** 32 - 1, 1000003 * 7 + 5]
    grids = []
    for s in seeds:
        np.random.seed(s)
        m = fn(None, input_size=(256, 256), mask_ratio=0.5, patch_size=16)
        assert m.shape == (1, 256, 256) and m.dtype == np.float64
        g = m[0, ::16, ::16]
        assert (np.kron(g, np.ones((16, 16))) == m[0]).all()
        grids.append(g.astype(np.uint8))
    # a second geometry / ratio (the function's defaults) to pin the general case
    extra = []
    for s in range(8):
        np.random.seed(s)
        m = fn(None, input_size=(352, 352), mask_ratio=0.75, patch_size=16)
        extra.append(m[0, ::16, ::16].astype(np.uint8))
    np.savez_compressed(os.path.join(HERE, "grid_mask_golden.npz"), seeds=np.array(seeds, dtype=np.uint64),
                        grids=np.stack(grids), extra_352_075=np.stack(extra))
    print("grid masks:", len(seeds), "seeds; masked fraction mean", float(np.mean(grids)))


TINY_CASES = (("pre", {"itm": 1, "mlm": 1, "t2i": 1, "cls": 0}, 2), ("cls", {"itm": 0, "mlm": 0, "t2i": 0, "cls": 1}, 2))
SMALL_CASES = (("pre", {"itm": 1, "mlm": 1, "t2i": 1, "cls": 0}, 1),)    # BASELINE configs[4] stand-in


def golden_pvlt(model="pvlt_tiny", cases=TINY_CASES):
    ref_pvlt = import_reference()
    from oracle import pvlt_oracle as O
    out = {}
    for tag, loss_type, B in cases:
        torch.manual_seed(0)
        m = getattr(ref_pvlt, model)(pretrained=True, token_hidden_size=768, num_text_tokens=128, loss_type=loss_type,
                               pretrained_pth="", num_classes=1000, in_chans=3, drop_rate=0.0, drop_path_rate=0.0)
        sd = O.make_state_dict(model, loss_type, seed=0)
        ref_keys = set(m.state_dict().keys())
        assert ref_keys == set(sd.keys()), (sorted(ref_keys - set(sd)), sorted(set(sd) - ref_keys))
        for k, v in m.state_dict().items():
            assert tuple(v.shape) == tuple(sd[k].shape), (k, v.shape, sd[k].shape)
        m.load_state_dict(sd)
        assert m.mlm_head.mlm_decoder.weight.data_ptr() == m.text_embeddings.word_embeddings.weight.data_ptr() \
            if loss_type["mlm"] else True
        m.text_embeddings.dropout.p = 0.0
        m.train()
        batch = O.make_inputs(B, seed=0)
        o = m(batch["images"], batch["input_ids"])
        ls = O.losses(o, batch, batch["images"])     # same formulas as engine_grid_masking.py:81-102
        ls["total"].backward()
        for k in ("mlm", "itm", "t2i", "sup_cls", "sub_cls", "total"):
            if k in ls:
                out[f"{tag}_loss_{k}"] = np.float64(ls[k].item())
        if o["mlm_logits"] is not None:
            out[f"{tag}_mlm_logits_sub"] = o["mlm_logits"][:, :8, ::257].detach().numpy()
            out[f"{tag}_mlm_logits_lse"] = torch.logsumexp(o["mlm_logits"], -1).detach().numpy()
        if o["itm_logits"] is not None:
            out[f"{tag}_itm_logits"] = o["itm_logits"].detach().numpy()
        if o["sup_cls_logits"] is not None:
            out[f"{tag}_sup_cls_logits"] = o["sup_cls_logits"].detach().numpy()
            out[f"{tag}_sub_cls_logits"] = o["sub_cls_logits"].detach().numpy()
        if o["t2i_logits"] is not None:
            out[f"{tag}_t2i_logits_sub"] = o["t2i_logits"][:, :, ::16, ::16].detach().numpy()
        names, gsum, gabs = [], [], []
        for n, p in m.named_parameters():
            names.append(n)
            g = p.grad if p.grad is not None else torch.zeros_like(p)
            gsum.append(g.double().sum().item())
            gabs.append(g.double().abs().sum().item())
        out[f"{tag}_grad_names"] = np.array(names)
        out[f"{tag}_grad_sum"] = np.array(gsum)
        out[f"{tag}_grad_abs"] = np.array(gabs)
        # eval-mode forward (BN running stats) for the retrieval path
        m.eval()
        with torch.no_grad():
            oe = m(batch["images"], batch["input_ids"])
        if oe["itm_logits"] is not None:
            out[f"{tag}_eval_itm_logits"] = oe["itm_logits"].numpy()
        if oe["t2i_logits"] is not None:
            out[f"{tag}_eval_t2i_logits_sub"] = oe["t2i_logits"][:, :, ::16, ::16].numpy()
        print(tag, {k: float(v) for k, v in ls.items()})
    np.savez_compressed(os.path.join(HERE, f"{model}_golden.npz"), **out)


def golden_token_masks():
    """Runs the reference's own ``random_masking_features`` (fashion_gen.py:383-409) on word-piece STRINGS with the
    vendored BERT vocabulary under ``random.seed(s)``; records ids so that the fixture does not depend on the tokenizer."""
    import collections
    import random
    src = open(os.path.join(REF, "mcloader", "fashion_gen.py")).read()
    start = src.index("    def random_masking_features")
    end = src.index("    def rgb_loader")
    ns = {"random": random}
    exec("class _Holder:\n" + src[start:end], ns)          # executes the reference text in place, nothing copied
    fn = ns["_Holder"].random_masking_features
    vocab = collections.OrderedDict()
    with open(os.path.join(REF, "preweights", "bert-base-uncased-vocab.txt"), encoding="utf-8") as f:
        for i, line in enumerate(f):
            vocab[line.rstrip("\n")] = i
    inv = list(vocab.keys())
    assert len(vocab) == 30522 and vocab["[MASK]"] == 103 and vocab["[SEP]"] == 102 and vocab["[CLS]"] == 101

    class _Tok:
        pass
    holder = types.SimpleNamespace(word_mask_rate=0.15, tokenizer=_Tok())
    holder.tokenizer.vocab = vocab
    T = 128
    rs = np.random.RandomState(7)
    seeds, ori, ids, labels = [], [], [], []
    for s in list(range(48)) + [12345, 2 ** 31 - 1, 2 ** 32 - 1, 1000003 * 9 + 1]:
        L = int(rs.randint(3, T - 1))                          # number of word pieces (max T - 2, text_process :326-328)
        pieces = rs.randint(1000, 30522, size=L)
        toks = [inv[i] for i in pieces]
        random.seed(s)
        out_toks, lab = fn(holder, list(toks))
        row_ori = [101] + [int(i) for i in pieces] + [102] + [0] * (T - L - 2)
        row_ids = [101] + [vocab[t] for t in out_toks] + [102] + [0] * (T - L - 2)
        row_lab = [-1] + [int(v) for v in lab] + [-1] + [-1] * (T - L - 2)
        seeds.append(s); ori.append(row_ori); ids.append(row_ids); labels.append(row_lab)
    np.savez_compressed(os.path.join(HERE, "token_mask_golden.npz"), seeds=np.array(seeds, dtype=np.uint64),
                        ori=np.array(ori, dtype=np.int64), ids=np.array(ids, dtype=np.int64),
                        labels=np.array(labels, dtype=np.int64))
    print("token_mask_golden.npz:", len(seeds), "samples,", int((np.array(labels) != -1).sum()), "labelled pieces")


CAPTIONS = [
    "Long sleeve cotton poplin shirt in white. Spread collar. Button closure at front.",
    "Slim-fit jeans in 'washed' indigo blue; fading, whiskering & distressing throughout!",
    "Leather ankle boots in black. Almond toe. Zip closure at inner side. Tonal stitching. Approx. 1.5\" heel.",
    "T-shirt",
    "",
    "Relaxed-fit hoodie " * 60,
    "Crêpe de Chine blouse — naïve floral print, 100% silk (Made in Italy)",
    "BOMBER JACKET IN NAVY / RIB KNIT COLLAR, CUFFS, AND HEM / TWO-WAY ZIP",
]


def golden_text():
    """Runs the reference's own ``text_process`` (fashion_gen.py:321-381) with the vendored vocabulary; records the
    RNG-independent outputs (ori_input_ids, attention_mask, segment_ids) for mvlt_b200.text.encode_captions."""
    import random
    from transformers import BertTokenizer
    import inspect
    vf = os.path.join(REF, "preweights", "bert-base-uncased-vocab.txt")
    if "vocab" in inspect.signature(BertTokenizer.__init__).parameters:
        with open(vf, encoding="utf-8") as f:
            tok = BertTokenizer(vocab={line.rstrip("\n"): i for i, line in enumerate(f)}, do_lower_case=True)
    else:
        tok = BertTokenizer(vocab_file=vf, do_lower_case=True)
    src = open(os.path.join(REF, "mcloader", "fashion_gen.py")).read()
    start = src.index("    def text_process")
    end = src.index("    def rgb_loader")
    ns = {"random": random, "torch": torch,
          "shift_tokens_right": lambda ids, pad_token_id, decoder_start_token_id: ids}   # BART leftover, unused output
    exec("class _Holder:\n" + src[start:end], ns)          # executes the reference text in place, nothing copied
    holder = types.SimpleNamespace(word_mask_rate=0.15, tokenizer=tok)
    holder.random_masking_features = lambda toks: ns["_Holder"].random_masking_features(holder, toks)
    ori, att, seg = [], [], []
    for T in (128, 24):
        for cap in CAPTIONS:
            random.seed(0)
            out = ns["_Holder"].text_process(holder, cap, T)
            input_ids, attention_mask, mlm_labels, segment_ids, ori_input_ids = out[:5]
            row = lambda t: np.pad(t.numpy(), (0, 128 - T))
            ori.append(row(ori_input_ids)); att.append(row(attention_mask)); seg.append(row(segment_ids))
    np.savez_compressed(os.path.join(HERE, "text_golden.npz"), captions=np.array(CAPTIONS), lengths=np.array([128, 24]),
                        ori=np.stack(ori), att=np.stack(att), seg=np.stack(seg))
    print("text_golden.npz:", len(ori), "rows")


if __name__ == "__main__":
    if "--text-only" in sys.argv:
        golden_text()
        sys.exit(0)
    if "--small-only" not in sys.argv:
        golden_text()
        golden_grid_masks()
        golden_token_masks()
        golden_pvlt()
    golden_pvlt("pvlt_small", SMALL_CASES)
