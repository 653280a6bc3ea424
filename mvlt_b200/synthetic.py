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
# Planted-positive retrieval protocol (SURVEY 7 H3 option iii). At random init the ITM score spread over 101 candidates
# (6e-3) is below bf16 noise, so "identical rankings" cannot be tested there. Here images and captions carry a planted
# class: an image is a coarse 4x4 grid of class colours plus pixel noise, a caption is [CLS] + 6 class tokens + [SEP].
# A model fitted for a few hundred steps on matched / mismatched pairs (mvlt_b200.retrieval.fit_planted_itm) separates
# positives from negatives by several logits, and the rank of the positive (engine_grid_masking.py:360-384) becomes a
# well-conditioned quantity that the bf16 kernels and the fp32 reference must agree on exactly.
# ------------------------------------------------------------------------------------------------------------
PLANTED_CLASSES = 8


def planted_palette(n_classes: int = PLANTED_CLASSES, seed: int = 0):
    g = torch.Generator().manual_seed(424242 + seed)
    colours = torch.rand((n_classes, 3, 4, 4), generator=g) * 0.9 + 0.05
    tokens = torch.randint(1000, VOCAB, (n_classes, 6), generator=g)
    return colours, tokens


def _planted_images(classes, g, colours, noise, img, device="cpu"):
    base = torch.nn.functional.interpolate(colours.to(device)[classes], size=(img, img), mode="nearest")
    return (base + (torch.rand(base.shape, generator=g, device=device) - 0.5) * 2 * noise).clamp_(0.0, 1.0)


def _planted_ids(classes, tokens, T, device="cpu"):
    ids = torch.zeros((classes.shape[0], T), dtype=torch.long, device=device)
    ids[:, 0] = 101
    ids[:, 1:7] = tokens.to(device)[classes]
    ids[:, 7] = 102
    return ids


def planted_pairs(batch: int, seed: int, n_classes: int = PLANTED_CLASSES, noise: float = 0.1, T: int = 128, img: int = 256,
                  device="cpu"):
    """Training pairs: ITM label 1 = the caption names the image's class, 0 = another class (fashion_gen.py:121-146).
    ``device``: where the batch is generated (a CUDA generator makes fitting fast; the draws differ from the CPU ones)."""
    colours, tokens = planted_palette(n_classes)
    g = torch.Generator(device=device).manual_seed(99991 * (seed + 1))
    cls_img = torch.randint(0, n_classes, (batch,), generator=g, device=device)
    match = torch.randint(0, 2, (batch,), generator=g, device=device)
    shift = torch.randint(1, n_classes, (batch,), generator=g, device=device)
    cls_txt = torch.where(match == 1, cls_img, (cls_img + shift) % n_classes)
    return dict(images=_planted_images(cls_img, g, colours, noise, img, device), input_ids=_planted_ids(cls_txt, tokens, T, device),
                itm_labels=match.view(batch, 1))


def planted_query(q: int, n_cand: int = 101, n_classes: int = PLANTED_CLASSES, noise: float = 0.1, T: int = 128, img: int = 256,
                  mode: str = "tir"):
    """One retrieval query on the CPU (the fp32 reference and the kernels see the same tensors), positive at index 0.
    ``tir``: one caption against n_cand images; ``itr``: one image against n_cand captions (fashion_gen.py:436-508).
    Negatives are drawn from the other classes."""
    colours, tokens = planted_palette(n_classes)
    g = torch.Generator().manual_seed(777 + 31337 * q)
    c = int(torch.randint(0, n_classes, (1,), generator=g))
    neg = (c + torch.randint(1, n_classes, (n_cand - 1,), generator=g)) % n_classes
    classes = torch.cat([torch.tensor([c]), neg])
    if mode == "tir":
        images = _planted_images(classes, g, colours, noise, img)
        ids = _planted_ids(torch.full((n_cand,), c), tokens, T)
    else:
        images = _planted_images(torch.full((1,), c), g, colours, noise, img).expand(n_cand, -1, -1, -1).contiguous()
        ids = _planted_ids(classes, tokens, T)
    return images, ids
