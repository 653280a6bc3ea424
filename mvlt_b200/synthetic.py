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


# ------------------------------------------------------------------------------------------------------------
# Planted-positive retrieval protocol (SURVEY 7 H3 option iii; stated in README.md). At random init the ITM score spread
# over 101 candidates (6e-3) is below bf16 noise, so "identical rankings" cannot be tested there. The protocol plants the
# match signal in a quantity the random encoder already transmits to text token 0 -- image brightness -- and fits ONLY the
# last ITM linear layer (768 -> 2) on reference features with a few optimiser steps (tests/test_engine_gpu.py does the
# fit with the fp32 reference restatement). The fitted score is monotone in brightness with gaps of several logits per
# 0.1 of brightness, so the rank of the positive (engine_grid_masking.py:360-384) is a well-conditioned quantity that the
# bf16 kernels and the fp32 reference must agree on exactly: the positive sits at brightness 0.55, ``n_brighter`` decoys
# are brighter (>= 0.76, they outrank it), the remaining candidates are darker (<= 0.34).
# ------------------------------------------------------------------------------------------------------------
def brightness_images(levels: torch.Tensor, g: torch.Generator, img: int = 256, noise: float = 0.3):
    n = levels.shape[0]
    return (levels.view(n, 1, 1, 1) + (torch.rand((n, 3, img, img), generator=g) - 0.5) * noise).clamp_(0.0, 1.0)


def planted_fit_set(n: int = 128, seed: int = 0):
    """Images of uniform random brightness with varied synthetic captions; label 1 = brighter than 0.5."""
    g = torch.Generator().manual_seed(5150 + seed)
    levels = torch.rand(n, generator=g) * 0.8 + 0.1
    return brightness_images(levels, g), make_batch(n, seed=3 + seed)["ori_input_ids"], (levels > 0.5).long()


def planted_tir_query(q: int, n_cand: int = 101):
    """TIR query q: one caption against n_cand images, positive at index 0 (fashion_gen.py:436-508 layout). Returns
    (images, input_ids, expected_rank) with expected_rank = the number of brighter decoys = q % 11."""
    g = torch.Generator().manual_seed(8086 + 977 * q)
    n_brighter = q % 11
    n_dark = n_cand - 1 - n_brighter
    levels = torch.cat([torch.tensor([0.55]), 0.76 + 0.16 * torch.rand(n_brighter, generator=g),
                        0.10 + 0.24 * torch.rand(n_dark, generator=g)])
    perm = torch.cat([torch.zeros(1, dtype=torch.long), 1 + torch.randperm(n_cand - 1, generator=g)])
    images = brightness_images(levels[perm], g)
    ids = make_batch(1, seed=1000 + q)["ori_input_ids"].repeat(n_cand, 1)
    return images, ids, n_brighter


# ------------------------------------------------------------------------------------------------------------
# Planted recognition protocol (the recognition loop, engine_grid_masking.py:396-462, against the fp32 reference). Same idea as
# the retrieval protocol: the class is planted in image brightness and ONLY the last linear layer of the two category heads
# (768 -> 48 / 122) is fitted on reference features. The random encoder's response to brightness saturates for bright
# images, so the four levels sit where it is steep (equal steps of the response, not of the brightness): the argmax of a
# sample is then decided by a margin of a few logits, and the bf16 kernels and the fp32 reference must predict the same class.
# ------------------------------------------------------------------------------------------------------------
PLANTED_CLS_LEVELS = (0.05, 0.12, 0.22, 0.55)
PLANTED_SUP_CLASSES = (5, 17, 29, 41)        # of the 48 M-CR categories
PLANTED_SUB_CLASSES = (3, 40, 77, 110)       # of the 122 S-CR categories


def planted_cls_set(n: int, seed: int = 0):
    """(images, ori_input_ids, sup_cls_labels [n, 1], sub_cls_labels [n, 1]): a sample's category pair is one of four,
    carried by its brightness level (+- 0.01); captions are varied and carry no signal."""
    g = torch.Generator().manual_seed(4004 + seed)
    c = torch.randint(0, 4, (n,), generator=g)
    levels = torch.tensor(PLANTED_CLS_LEVELS)[c] + (torch.rand(n, generator=g) - 0.5) * 0.02
    images = brightness_images(levels, g, noise=0.1)
    ids = make_batch(n, seed=50 + seed)["ori_input_ids"]
    return images, ids, torch.tensor(PLANTED_SUP_CLASSES)[c].view(n, 1), torch.tensor(PLANTED_SUB_CLASSES)[c].view(n, 1)
