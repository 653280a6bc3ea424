"""Synthetic Fashion-Gen-shaped inputs for benchmarks and smoke runs: 256x256 ``rand`` images,
128-token BERT id rows ([CLS] tokens [SEP] pad...), 15 % MLM corruption (80/10/10, /root/reference/mcloader/
fashion_gen.py:383-409), Bernoulli ITM labels (:123) and uniform category labels."""
from __future__ import annotations

import torch

VOCAB = 30522


def make_batch(batch: int, seed: int = 0, T: int = 128, img: int = 256, pin: bool = False):
    g = torch.Generator().manual_seed(7919 * (seed + 1))
    images = torch.rand((batch, 3, img, img), generator=g)
    L = torch.randint(16, 65, (batch,), generator=g)
    pos = torch.arange(T).unsqueeze(0)
    toks = torch.randint(1000, VOCAB, (batch, T), generator=g)
    ori = torch.where(pos == 0, torch.full_like(toks, 101), toks)
    ori = torch.where(pos == (L - 1).unsqueeze(1), torch.full_like(toks, 102), ori)
    ori = torch.where(pos >= L.unsqueeze(1), torch.zeros_like(toks), ori)
    inner = (pos >= 1) & (pos < (L - 1).unsqueeze(1))
    r = torch.rand((batch, T), generator=g)
    r2 = torch.rand((batch, T), generator=g)
    rnd = torch.randint(1000, VOCAB, (batch, T), generator=g)
    chosen = inner & (r < 0.15)
    chosen[:, 1] |= ~chosen.any(dim=1)          # at least one labelled token per row keeps the CE well defined
    mlm = torch.where(chosen, ori, torch.full_like(ori, -1))
    ids = torch.where(chosen & (r2 < 0.8), torch.full_like(ori, 103), ori)
    ids = torch.where(chosen & (r2 >= 0.8) & (r2 < 0.9), rnd, ids)
    out = dict(images=images, input_ids=ids, ori_input_ids=ori, mlm_labels=mlm,
               itm_labels=torch.randint(0, 2, (batch, 1), generator=g),
               sup_cls_labels=torch.randint(0, 48, (batch, 1), generator=g),
               sub_cls_labels=torch.randint(0, 122, (batch, 1), generator=g))
    if pin:
        out = {k: v.pin_memory() for k, v in out.items()}
    return out
