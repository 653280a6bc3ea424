"""Random grid masking and BERT token masking on the GPU, bit-exact against the reference's host implementations for a
given per-sample seed.

Reference: /root/reference/mcloader/fashion_gen.py:225-254 (``generate_grid_mask``; note the sliding-window quirk at
:246) and :176-177 (``masked_fill_`` with 1e-6). The reference draws from numpy's process-global legacy RNG inside
DataLoader workers; here every sample owns an MT19937 stream seeded explicitly (``sample_seed``), which reproduces
``np.random.seed(s); generate_grid_mask(...)`` bit for bit (csrc/mask.cu).
"""
from __future__ import annotations

import numpy as np
import torch

from . import kernels as k
from ._lib import MvltError


def sample_seed(seed: int, sample_idx: int) -> int:
    """Per-sample MT19937 seed used by the harness: np.random.seed(seed * 1000003 + idx) (mod 2^32)."""
    return (int(seed) * 1000003 + int(sample_idx)) & 0xFFFFFFFF


def grid_mask_batch(seeds, input_size=(256, 256), mask_ratio=0.5, patch_size=16, device="cuda") -> torch.Tensor:
    """uint8 [B, H/patch, W/patch] grid (1 = masked patch) for one MT19937 seed per sample."""
    if input_size[0] % patch_size or input_size[1] % patch_size:
        raise AssertionError("input size must be divisible by patch_size")   # fashion_gen.py:227-228
    if torch.is_tensor(seeds):
        s = seeds.to(device=device, dtype=torch.int64)
    else:
        s = torch.tensor([int(v) & 0xFFFFFFFF for v in seeds], dtype=torch.int64, device=device)
    B = s.numel()
    s32 = (s & 0xFFFFFFFF).to(torch.uint32) if hasattr(torch, "uint32") else s.to(torch.int32)
    nw, nh = input_size[0] // patch_size, input_size[1] // patch_size
    grid = torch.empty((B, nh, nw), dtype=torch.uint8, device=device)
    k.grid_mask(s32.contiguous(), grid, B, input_size[0], input_size[1], patch_size, float(mask_ratio))
    return grid


def apply_grid_mask(images: torch.Tensor, grid: torch.Tensor, patch_size=16, fill=1e-6, return_mask=False, out=None):
    """masked_images = images.masked_fill(mask, 1e-6) with mask[b, 0, y, x] = grid[b, y//P, x//P]
    (fashion_gen.py:176). Optionally also returns the float mask [B,1,H,W] (the dataset's ``t2i_labels``)."""
    if not images.is_cuda or images.dtype != torch.float32:
        raise MvltError("apply_grid_mask expects a CUDA float32 [B,C,H,W] tensor")
    images = images.contiguous()
    B, Cc, H, W = images.shape
    if out is None:
        out = torch.empty_like(images)
    elif out.shape != images.shape or out.dtype != images.dtype or not out.is_cuda or not out.is_contiguous():
        raise MvltError("apply_grid_mask: `out` must be a contiguous CUDA tensor shaped like the images")
    mask = torch.empty((B, 1, H, W), dtype=torch.float32, device=images.device) if return_mask else None
    k.masked_fill(images, grid.contiguous(), out, mask, B, Cc, H, W, patch_size, fill)
    return (out, mask) if return_mask else out


def generate_grid_mask(input_size=(352, 352), mask_ratio=0.75, patch_size=16, seed=None) -> np.ndarray:
    """Signature and return value of the reference method (float64 ndarray [1, H, W] of {0,1}); ``seed`` selects the
    MT19937 stream (``None`` draws one from numpy's global RNG, as the reference implicitly does)."""
    if seed is None:
        seed = int(np.random.randint(0, 2 ** 32, dtype=np.uint64))
    grid = grid_mask_batch([seed], input_size, mask_ratio, patch_size)[0].cpu().numpy()
    return np.kron(grid, np.ones((patch_size, patch_size)))[None].astype(np.float64)


def _seeds_u32(seeds, device):
    if torch.is_tensor(seeds):
        s = seeds.to(device=device, dtype=torch.int64)
    else:
        s = torch.tensor([int(v) & 0xFFFFFFFF for v in seeds], dtype=torch.int64, device=device)
    return ((s & 0xFFFFFFFF).to(torch.uint32) if hasattr(torch, "uint32") else s.to(torch.int32)).contiguous()


def mask_tokens_batch(seeds, ori_input_ids: torch.Tensor, mask_rate=0.15, sep_id=102, mask_id=103, vocab_size=30522):
    """BERT word-piece masking of the reference's text pipeline on the device (mcloader/fashion_gen.py:383-409, called
    from text_process :334): every word piece between [CLS] and [SEP] is selected with probability ``mask_rate``; a
    selected piece becomes [MASK] (80 %), a uniformly random vocabulary id (10 %) or stays (10 %) and its original id is
    the MLM label, all other positions get label -1. Sample b reproduces ``random.seed(seeds[b])`` followed by the
    reference loop bit for bit (CPython MT19937 / random() / choice()).

    ori_input_ids: int64 [B, T] rows ``[CLS] pieces... [SEP] [PAD]...``. Returns (input_ids, mlm_labels), int64 [B, T]."""
    if not ori_input_ids.is_cuda or ori_input_ids.dtype != torch.int64 or ori_input_ids.dim() != 2:
        raise MvltError("mask_tokens_batch expects a CUDA int64 [B, T] tensor")
    ori = ori_input_ids.contiguous()
    B, T = ori.shape
    ids = torch.empty_like(ori)
    labels = torch.empty_like(ori)
    k.token_mask(_seeds_u32(seeds, ori.device), ori, ids, labels, B, T, sep_id, mask_id, vocab_size, float(mask_rate))
    return ids, labels
