"""Drop-in for the reference's ``libs/pvlt.py`` (/root/reference/libs/pvlt.py).

Same entry points (``pvlt_tiny|small|medium|large``, pvlt.py:415-483), same constructor arguments, same
``forward(input_images, input_ids) -> dict`` contract (pvlt.py:358-401) and the same parameter / buffer names
(SURVEY Appendix A), so ``main_vl.py`` (create_model / DDP / optimizer / state_dict / resume) keeps working.
The ``nn.Module`` tree below only OWNS parameters; no torch op computes anything: ``forward`` is one
``torch.autograd.Function`` that drives the hand-written sm_100a kernels through ``mvlt_b200.engine``.

Added (not in the reference): ``forward_losses`` -- the fused pre-training step used by our
``engine_grid_masking.train_one_epoch_vl`` (labels in, scalar losses out; the MLM head runs only on labelled
rows, which changes no loss or gradient value).
"""
from __future__ import annotations

from functools import partial
from typing import Dict, Optional

import torch
import torch.nn as nn

from .. import _lib
from .. import kernels as k
from .._lib import MvltError
from ..engine import DEPTHS, EMBED_DIMS, HIDDEN, MLP_RATIOS, NUM_HEADS, PATCH, SR_RATIOS, VOCAB, VOCAB_PAD, PVLTEngine
from .vl_heads import CLSHead, ITGHead, ITMHead, MLMHead

__all__ = ["pvlt_tiny", "pvlt_small", "pvlt_medium", "pvlt_large", "PyramidVisionLanguageTransformer"]

F32, BF16 = torch.float32, torch.bfloat16
MLM_LOSS_WEIGHT, ITM_LOSS_WEIGHT, T2I_LOSS_WEIGHT = 1.0, 1.0, 10.0   # engine_grid_masking.py:23


# ---- parameter containers (names == reference) -----------------------------------------------------------
class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.fc2 = nn.Linear(hidden_features, in_features)


class Attention(nn.Module):
    def __init__(self, dim, num_heads, qkv_bias, sr_ratio):
        super().__init__()
        assert dim % num_heads == 0, f"dim {dim} should be divided by num_heads {num_heads}."
        self.q = nn.Linear(dim, dim, bias=qkv_bias)
        self.kv = nn.Linear(dim, dim * 2, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        self.sr_ratio = sr_ratio
        if sr_ratio > 1:
            self.sr = nn.Conv2d(dim, dim, kernel_size=sr_ratio, stride=sr_ratio)
            self.norm = nn.LayerNorm(dim)


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio, qkv_bias, norm_layer, sr_ratio):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads, qkv_bias, sr_ratio)
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))


class PatchEmbed(nn.Module):
    def __init__(self, img_size, patch_size, in_chans, embed_dim):
        super().__init__()
        assert img_size % patch_size == 0, f"img_size {img_size} should be divided by patch_size {patch_size}."
        self.img_size, self.patch_size = (img_size, img_size), (patch_size, patch_size)
        self.H = self.W = img_size // patch_size
        self.num_patches = self.H * self.W
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.norm = nn.LayerNorm(embed_dim)


class TextEmbeddings(nn.Module):
    """Parameter layout of transformers' BertEmbeddings (bert-base-uncased config), pvlt.py:232-233."""

    def __init__(self, vocab=VOCAB, hidden=HIDDEN, max_pos=512, type_vocab=2, p_drop=0.1):
        super().__init__()
        self.word_embeddings = nn.Embedding(vocab, hidden, padding_idx=0)
        self.position_embeddings = nn.Embedding(max_pos, hidden)
        self.token_type_embeddings = nn.Embedding(type_vocab, hidden)
        self.LayerNorm = nn.LayerNorm(hidden, eps=1e-12)
        self.dropout = nn.Dropout(p_drop)


# ---- the single autograd node ------------------------------------------------------------------------------
class _PVLTFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, mode, batch, images, input_ids, *params):
        eng: PVLTEngine = model._engine()
        training = model.training
        mode, grad_mode, heads = mode
        need_grad = grad_mode and any(p.requires_grad for p in params)   # (grad mode is always off inside forward)
        eng.embed_dropout = model.text_embeddings.dropout.p
        eng.prepare_weights()
        images = images.contiguous().to(F32)
        if mode == "logits":
            outs, saved = eng_forward_logits(eng, images, input_ids, training, need_grad, heads)
        else:
            outs, saved = eng_forward_losses(eng, images, input_ids, batch, training, need_grad)
        ctx.model, ctx.mode, ctx.saved = model, mode, saved
        ctx.set_materialize_grads(False)
        if mode == "losses":
            ctx.mark_non_differentiable(outs[1])
        return outs

    @staticmethod
    def backward(ctx, *gouts):
        model, saved = ctx.model, ctx.saved
        if saved is None:
            raise MvltError("backward called on a forward that ran without gradient tracking")
        G = run_backward(model, ctx.mode, saved, gouts)
        ctx.saved = None
        names = model._param_names
        return (None, None, None, None, None) + tuple(G[n] for n in names)


def run_backward(model, mode, saved, gouts):
    """The hand-scheduled backward of one forward (``saved``) into a flat gradient buffer, including the data-parallel
    exchange when ``model.enable_grad_sync`` is on. Returns the gradient dict (``G[name]`` = view of the flat buffer).
    Called by the autograd node above and, without autograd, by ``mvlt_b200.graph.GraphedStep``."""
    eng: PVLTEngine = model._engine()
    G = eng.new_grads()
    sync = model.__dict__.get("_grad_sync")
    # data parallelism without DistributedDataParallel's bucket copies: every gradient of the step lives in ONE flat fp32
    # buffer laid out in completion order, and each segment (heads, stage 4, 3, 2, stage 1 + embeddings) is all-reduced
    # (averaged) over NVLink on a side stream as soon as the hand-scheduled backward has finished it -- the exchange of
    # everything but the last segment overlaps the remaining backward kernels (DDP's overlap at main_vl.py:297-299)
    reducer = SegmentReducer(G["__flat__"], G["__segments__"], sync) if sync is not None else None
    on_seg = reducer.segment_done if (reducer is not None and GRAD_SYNC_OVERLAP) else None
    if mode == "logits":
        eng_backward_logits(eng, saved, gouts, G, on_seg)
    else:
        eng_backward_losses(eng, saved, gouts[0], G, on_seg)
    eng.wgrad_join()
    if reducer is not None:
        reducer.finish()
    return G


def allreduce_flat_(flat, group=None):
    """In-place average of a flat gradient buffer over the data-parallel group (the role DDP's reducer plays at
    /root/reference/main_vl.py:297-299)."""
    import torch.distributed as dist
    grp = None if group is True else group
    world = dist.get_world_size(grp)
    if world == 1:
        return flat
    if dist.get_backend(grp) == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=grp)
    else:   # gloo (CPU tests) has no AVG
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=grp)
        flat.div_(world)
    return flat


# MVLT_GRAD_OVERLAP=0: exchange all segments after the last backward kernel (A/B switch; same result)
GRAD_SYNC_OVERLAP = __import__("os").environ.get("MVLT_GRAD_OVERLAP", "1") != "0"


class SegmentReducer:
    """Averages the segments of a flat gradient buffer over the data-parallel group, each as soon as it is complete.
    CUDA: the all-reduce of segment k is enqueued on a side stream behind an event recorded on the compute stream, so it
    runs under the backward kernels of the following segments; ``finish`` makes the compute stream wait for all of them.
    CPU tensors (gloo tests): the same calls, synchronously."""

    def __init__(self, flat, segments, group=True):
        import torch.distributed as dist
        self.flat, self.segments = flat, segments
        self.group = None if group is True else group
        self.world = dist.get_world_size(self.group)
        self.cuda = flat.is_cuda
        self.done = set()
        self.events = []
        if self.cuda and self.world > 1:
            key = flat.device.index
            side = _SIDE_STREAMS.get(key)
            if side is None:
                side = _SIDE_STREAMS[key] = torch.cuda.Stream(device=flat.device)
            self.side = side

    def segment_done(self, k, also_wait=None):
        """``also_wait``: optional extra CUDA event (work on another stream that also writes the segment) the exchange waits for."""
        if self.world == 1 or k in self.done:
            return
        self.done.add(k)
        b, e = self.segments[k]
        if e <= b:
            return
        part = self.flat[b:e]
        if not self.cuda:
            allreduce_flat_(part, True if self.group is None else self.group)
            return
        main = torch.cuda.current_stream()
        ready = torch.cuda.Event()
        ready.record(main)
        with torch.cuda.stream(self.side):
            self.side.wait_event(ready)
            if also_wait is not None:
                self.side.wait_event(also_wait)
            allreduce_flat_(part, True if self.group is None else self.group)
            ev = torch.cuda.Event()
            ev.record(self.side)
        self.events.append(ev)

    def finish(self):
        for k in range(len(self.segments)):
            self.segment_done(k)
        if self.cuda and self.world > 1:
            main = torch.cuda.current_stream()
            for ev in self.events:
                main.wait_event(ev)


_SIDE_STREAMS = {}


def _heads_common_fwd(eng, enc, B):
    st4 = enc["stages"][3]
    X4 = st4["out"]
    HW4 = st4["H"] * st4["W"]
    return X4, HW4


def eng_forward_logits(eng: PVLTEngine, images, ids, training, save, heads=None):
    """pvlt.py:358-401: returns the five logits tensors (None for disabled heads). ``heads``: optional subset of
    {"mlm", "itm", "cls", "t2i"} to evaluate (the others come back as None, e.g. no [B, 128, 30522] logits for a caller
    that only wants the reconstruction)."""
    B = images.shape[0]
    T = eng.T
    enc = eng.encoder_fwd(images, ids, training, save)
    X4, HW4 = _heads_common_fwd(eng, enc, B)
    lt = eng.loss_type
    if heads is not None:
        lt = {key: (v if key in heads else 0) for key, v in lt.items()}
    hc = {}
    mlm = itm = sup = sub = t2i = None
    if lt.get("mlm"):
        lg, hc["mlm"] = eng.mlm_fwd(X4, B, HW4, None, B * T, out_f32=True)
        mlm = lg.as_strided((B, T, VOCAB), (T * VOCAB_PAD, VOCAB_PAD, 1))   # [B,T,V] view, row stride VOCAB_PAD
    if lt.get("itm"):
        lg, hc["itm"] = eng.small_head_fwd(X4, B, HW4, "itm")
        itm = lg.view(B, 1, 2)
    if lt.get("cls"):
        lg, hc["sup_cls"] = eng.small_head_fwd(X4, B, HW4, "sup_cls")
        sup = lg.view(B, 1, -1)
        lg, hc["sub_cls"] = eng.small_head_fwd(X4, B, HW4, "sub_cls")
        sub = lg.view(B, 1, -1)
    if lt.get("t2i"):
        feats = [(enc["stages"][i]["out"], enc["stages"][i]["H"], enc["stages"][i]["W"], EMBED_DIMS[i]) for i in (1, 2, 3)]
        score, hc["t2i"] = eng.t2i.forward(feats, B, training)
        h, w = feats[0][1], feats[0][2]
        t2i = torch.empty((B, 3, h * 8, w * 8), dtype=F32, device=images.device)
        k.upsample8_fwd(score, t2i, B, h, w, 8)
        hc["t2i"]["score"] = score
    outs = tuple(o if o is not None else torch.empty(0, device=images.device) for o in (mlm, itm, sup, sub, t2i))
    saved = dict(enc=enc, hc=hc, B=B, HW4=HW4) if save else None
    if not save:
        enc = None
    return outs, saved


def eng_backward_logits(eng: PVLTEngine, saved, gouts, G, on_segment=None):
    enc, hc, B, HW4 = saved["enc"], saved["hc"], saved["B"], saved["HW4"]
    T = eng.T
    dev = enc["ids"].device
    N4 = HW4 + T
    dX4 = k.zeros((B, N4, EMBED_DIMS[-1]), F32, dev)
    g_mlm, g_itm, g_sup, g_sub, g_t2i = gouts
    dXs = [None, None, None, dX4]
    if "mlm" in hc and g_mlm is not None:
        g = g_mlm.contiguous().view(B * T, VOCAB)
        dl = torch.empty((B * T, VOCAB_PAD), dtype=BF16, device=dev)
        k.cast2d(g, VOCAB, dl, VOCAB_PAD, B * T, VOCAB)
        eng.mlm_bwd(dl[:, :VOCAB], hc["mlm"], B, HW4, dX4, G)
    for name, g in (("itm", g_itm), ("sup_cls", g_sup), ("sub_cls", g_sub)):
        if name in hc and g is not None:
            eng.small_head_bwd(g.contiguous().view(B, -1).to(F32), hc[name], name, B, HW4, dX4, G)
    if "t2i" in hc and g_t2i is not None:
        c = hc["t2i"]
        h, w = c["dims"][1], c["dims"][2]
        dscore = k.zeros((B * h * w, 3), F32, dev)
        k.t2i_up_loss(c["score"], None, g_t2i.contiguous().to(F32), dscore, None, None, 0.0, 1.0, None, B, h, w, 8, 0, True)
        df2, df3 = eng.t2i.backward(dscore, c, G, dX4)
        dXs[1], dXs[2] = df2, df3
    if on_segment is not None:
        on_segment(0, eng.wgrad_event())       # every head gradient is enqueued (weight gradients: on the side stream)
    eng.encoder_bwd(enc, dXs, G, on_segment)


def eng_forward_losses(eng: PVLTEngine, images, ids, batch, training, save):
    """engine_grid_masking.py:81-102 fused onto the heads. ``batch`` carries labels / target images / weights.
    Returns (total_loss [], stats fp32 [12] = [total, mlm, itm, sup_cls, sub_cls, t2i, mlm_correct, mlm_count,
    itm_correct, sup_cls_correct, sub_cls_correct, 0])."""
    B = images.shape[0]
    T = eng.T
    dev = images.device
    enc = eng.encoder_fwd(images, ids, training, save)
    X4, HW4 = _heads_common_fwd(eng, enc, B)
    lt = eng.loss_type
    w = batch.get("weights", {})
    stats = k.zeros((12,), F32, dev)
    hc = {}
    only = batch.get("only")
    want = lambda name: only is None or name in only
    if lt.get("t2i") and want("t2i") and batch.get("target_images") is not None:
        # (first, and as a parallel branch of a captured graph: its ~60 small launches run next to the MLM / ITM heads)
        with eng.branch(1):
            feats = [(enc["stages"][i]["out"], enc["stages"][i]["H"], enc["stages"][i]["W"], EMBED_DIMS[i]) for i in (1, 2, 3)]
            score, c = eng.t2i.forward(feats, B, training)
            h, wd = feats[0][1], feats[0][2]
            target = batch["target_images"].contiguous().to(F32)
            wt = w.get("t2i", T2I_LOSS_WEIGHT)
            numel = B * 3 * h * 8 * wd * 8
            # training: the same launch also produces d loss / d score (for an upstream gradient of 1; the real one is a device
            # scalar folded into the first backward kernel), so the target image is read once per step, not twice
            dscore = k.zeros((B * h * wd, 3), F32, dev) if save else None
            k.t2i_up_loss(score, target, None, dscore, stats[5:6], stats[0:1], wt / numel, wt / numel if save else 0.0, None, B, h,
                          wd, 8, 1, bool(save))
            c.update(score=score, dscore=dscore)
            hc["t2i"] = c
    if lt.get("mlm") and want("mlm") and batch.get("mlm_labels") is not None:
        labels = batch["mlm_labels"]
        lab_dev = labels.to(dev, non_blocking=True).contiguous().view(-1)
        idx = torch.empty((B * T,), dtype=torch.int32, device=dev)
        lab_c = torch.empty((B * T,), dtype=torch.int64, device=dev)
        cnt = torch.empty((1,), dtype=torch.int32, device=dev)
        gs = eng.graph_state
        wm = w.get("mlm", MLM_LOSS_WEIGHT)
        if gs is not None and gs.mlm_cap and training and save:
            # static shapes (CUDA-graph capture): the head always runs on `mlm_cap` rows; rows beyond the labelled count are
            # padded with (token 0, ignore label) and contribute neither loss nor gradient; the count, 1 / count (the CE
            # scale) and an overflow flag stay on the device
            n = min(int(gs.mlm_cap), B * T)
            k.compact_labels(lab_dev, B * T, -1, idx, lab_c, cnt, cap=n, count_f32=stats[7:8], inv_count=gs.mlm_inv,
                             overflow=gs.mlm_overflow)
            scale, scale_dev = wm, gs.mlm_inv
        else:
            if batch.get("mlm_count") is not None:
                # supplied by the data pipeline: no device->host sync. The count is the caller's claim, so the compaction runs
                # in its fixed-capacity form: exactly n index rows are written (a claim above the real count is padded with
                # ignored rows, one below it drops the surplus), and the head can never gather through an unwritten index
                n = min(max(int(batch["mlm_count"]), 0), B * T)
                k.compact_labels(lab_dev, B * T, -1, idx, lab_c, cnt, cap=n)
            else:
                k.compact_labels(lab_dev, B * T, -1, idx, lab_c, cnt)
                n = int((labels != -1).sum()) if not labels.is_cuda else int(cnt.item())
            scale, scale_dev = wm / max(n, 1), None
        if n > 0:
            lg, c = eng.mlm_fwd(X4, B, HW4, idx, n)
            lse = torch.empty((n,), dtype=F32, device=dev)
            k.ce_fwd(lg, VOCAB_PAD, lab_c, n, VOCAB, -1, lse, stats[1:2], scale, total_sum=stats[0:1],
                     correct=stats[6:7], scale_dev=scale_dev)
            if scale_dev is None:
                stats[7:8].fill_(float(n))
            c.update(logits=lg, lse=lse, labels=lab_c, scale=scale, scale_dev=scale_dev, n=n)
            hc["mlm"] = c
    for name, key, wkey in (("itm", "itm_labels", "itm"), ("sup_cls", "sup_cls_labels", "cls"),
                            ("sub_cls", "sub_cls_labels", "cls")):
        on = lt.get("itm") if name == "itm" else lt.get("cls")
        if not on or not want(wkey) or batch.get(key) is None:
            continue
        lab = batch[key].to(dev, non_blocking=True).contiguous().view(-1)
        lg, c = eng.small_head_fwd(X4, B, HW4, name)
        n_cls = c["n"]
        lse = torch.empty((B,), dtype=F32, device=dev)
        wt = w.get(wkey, ITM_LOSS_WEIGHT if name == "itm" else 1.0)
        slot = {"itm": 2, "sup_cls": 3, "sub_cls": 4}[name]
        k.ce_fwd(lg, n_cls, lab, B, n_cls, -100, lse, stats[slot:slot + 1], wt / B, total_sum=stats[0:1],
                 correct=stats[slot + 6:slot + 7])
        c.update(logits=lg, lse=lse, labels=lab, scale=wt / B)
        hc[name] = c
    eng.join_branches()
    total = stats[0]
    saved = dict(enc=enc, hc=hc, B=B, HW4=HW4) if save else None
    return (total, stats), saved


def eng_backward_losses(eng: PVLTEngine, saved, gtotal, G, on_segment=None):
    enc, hc, B, HW4 = saved["enc"], saved["hc"], saved["B"], saved["HW4"]
    T = eng.T
    dev = enc["ids"].device
    N4 = HW4 + T
    dX4 = k.zeros((B, N4, EMBED_DIMS[-1]), F32, dev)
    dXs = [None, None, None, dX4]
    gs = gtotal.to(F32).contiguous() if gtotal is not None else None
    if "t2i" in hc:      # parallel branch of a captured graph (disjoint rows of dX4: image rows here, text rows below)
        with eng.branch(1):
            c = hc["t2i"]
            df2, df3 = eng.t2i.backward(c["dscore"], c, G, dX4, gscale=gs)
            dXs[1], dXs[2] = df2, df3
    if "mlm" in hc:
        c = hc["mlm"]
        lg = c["logits"]
        k.ce_bwd(lg, VOCAB_PAD, c["labels"], c["n"], VOCAB, -1, c["lse"], lg, VOCAB_PAD, c["scale"], gs,
                 scale_dev=c.get("scale_dev"))   # in place
        eng.mlm_bwd(lg, c, B, HW4, dX4, G)
    for name in ("itm", "sup_cls", "sub_cls"):
        if name in hc:
            c = hc[name]
            n_cls = c["n"]
            dl = torch.empty((B, n_cls), dtype=F32, device=dev)
            k.ce_bwd(c["logits"], n_cls, c["labels"], B, n_cls, -100, c["lse"], dl, n_cls, c["scale"], gs)
            eng.small_head_bwd(dl, c, name, B, HW4, dX4, G)
    eng.join_branches()
    if on_segment is not None:
        on_segment(0, eng.wgrad_event())       # every head gradient is enqueued (weight gradients: on the side stream)
    eng.encoder_bwd(enc, dXs, G, on_segment)


# ---- the model ---------------------------------------------------------------------------------------------
class PyramidVisionLanguageTransformer(nn.Module):
    """Pyramid Vision Language Transformer (PVLT); constructor signature of pvlt.py:179-185."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dims=[64, 128, 256, 512],
                 num_heads=[1, 2, 4, 8], mlp_ratios=[4, 4, 4, 4], qkv_bias=False, qk_scale=None, drop_rate=0.,
                 attn_drop_rate=0., drop_path_rate=0., norm_layer=nn.LayerNorm, depths=[3, 4, 6, 3],
                 sr_ratios=[8, 4, 2, 1], num_stages=4, F4=False,
                 token_hidden_size=768, num_text_tokens=128, loss_type={'itm': 1, 'mlm': 1, 'itg': 1, 'rtd': 1}):
        super().__init__()
        if (list(embed_dims) != EMBED_DIMS or list(num_heads) != NUM_HEADS or list(mlp_ratios) != MLP_RATIOS
                or list(sr_ratios) != SR_RATIOS or patch_size != 4 or num_stages != 4 or token_hidden_size != HIDDEN
                or not qkv_bias or qk_scale is not None or drop_rate != 0. or attn_drop_rate != 0. or in_chans != 3):
            raise MvltError("the sm_100a kernels are specialised to the PVLT family the reference registers "
                            "(pvlt.py:415-483: dims [64,128,320,512], heads [1,2,5,8], patch 4, qkv_bias, no dropout)")
        for key in ("mlm", "itm", "cls", "t2i"):
            loss_type[key]  # KeyError on a missing key, like pvlt.py:242-275
        self.num_classes, self.depths, self.F4, self.num_stages = num_classes, list(depths), F4, num_stages
        self.T_num = num_text_tokens
        self.loss_type = dict(loss_type)
        self.drop_path_rate = drop_path_rate
        for i in range(num_stages):
            patch_embed = PatchEmbed(img_size=img_size if i == 0 else img_size // (2 ** (i + 1)),
                                     patch_size=patch_size if i == 0 else 2,
                                     in_chans=in_chans if i == 0 else embed_dims[i - 1], embed_dim=embed_dims[i])
            text_embed = nn.Sequential(nn.Linear(token_hidden_size if i == 0 else embed_dims[i - 1], embed_dims[i]),
                                       nn.LayerNorm(embed_dims[i]))
            num_patches = patch_embed.num_patches if i != num_stages - 1 else patch_embed.num_patches + 1
            pos_embed = nn.Parameter(torch.zeros(1, num_patches, embed_dims[i]))
            text_pos_embed = nn.Parameter(torch.zeros(1, num_text_tokens, embed_dims[i]))
            block = nn.ModuleList([Block(embed_dims[i], num_heads[i], mlp_ratios[i], qkv_bias, norm_layer, sr_ratios[i])
                                   for _ in range(depths[i])])
            setattr(self, f"patch_embed{i + 1}", patch_embed)
            setattr(self, f"text_embed{i + 1}", text_embed)
            setattr(self, f"pos_embed{i + 1}", pos_embed)
            setattr(self, f"text_pos_embed{i + 1}", text_pos_embed)
            setattr(self, f"pos_drop{i + 1}", nn.Dropout(p=drop_rate))
            setattr(self, f"block{i + 1}", block)
            nn.init.trunc_normal_(pos_embed, std=.02)
            nn.init.trunc_normal_(text_pos_embed, std=.02)
        self.text_embeddings = TextEmbeddings()
        _config = {'vocab_size': VOCAB, 'hidden_size': token_hidden_size, 'num_layers': 2, 'num_heads': 12,
                   'mlp_ratio': 4, 'max_text_len': num_text_tokens, 'drop_rate': 0.1, 'hidden_act': 'gelu'}
        head_embed = lambda: nn.Sequential(nn.Linear(embed_dims[-1], token_hidden_size), nn.LayerNorm(token_hidden_size))
        if self.loss_type['mlm'] == 1:
            self.mlm_head_embed = head_embed()
            self.mlm_head = MLMHead(_config, self.text_embeddings.word_embeddings.weight)
        if self.loss_type['itm'] == 1:
            self.itm_head_embed = head_embed()
            self.itm_head = ITMHead(_config)
        if self.loss_type['cls'] == 1:
            self.sup_cls_head_embed = head_embed()
            self.sup_cls_head = CLSHead(_config, 48)
            self.sub_cls_head_embed = head_embed()
            self.sub_cls_head = CLSHead(_config, 122)
        if self.loss_type['t2i'] == 1:
            self.t2i_head = ITGHead(embed_dims=embed_dims, channel=64)
        self.apply(self._init_weights)
        self.__dict__["_eng"] = None
        self.__dict__["_grad_sync"] = None

    def _init_weights(self, m):   # pvlt.py:282-289
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    # -- engine plumbing
    def _params_list(self):
        """Parameters in named_parameters() order. Walking the module tree costs ~1 ms of host time per call, so the list
        is cached and revalidated against the modules' own parameter dicts only when a parameter object was replaced
        (``_apply`` / ``load_state_dict(assign=True)`` / re-registration bump this through ``_eng = None`` or identity)."""
        cached = self.__dict__.get("_plist")
        if cached is not None:
            owners, names, plist = cached
            ok = True
            for (mod, key), p in zip(owners, plist):
                if mod._parameters.get(key) is not p:
                    ok = False
                    break
            if ok:
                return names, plist
        owners, names, plist, seen = [], [], [], set()
        for mname, mod in self.named_modules():
            for key, p in mod._parameters.items():
                if p is None or id(p) in seen:
                    continue
                seen.add(id(p))
                owners.append((mod, key))
                names.append((mname + "." if mname else "") + key)
                plist.append(p)
        assert names == [n for n, _ in self.named_parameters()]
        self.__dict__["_plist"] = (owners, names, plist)
        return names, plist

    def _engine(self) -> PVLTEngine:
        eng = self.__dict__.get("_eng")
        names, plist = self._params_list()
        dev = plist[0].device
        if eng is None or eng._device != dev or eng._plist_id is not plist:
            params = dict(zip(names, plist))
            if dev.type != "cuda":
                raise MvltError("mvlt_b200 runs on CUDA (sm_100a) only: move the model with .to('cuda'); "
                                "there is no CPU fallback")
            eng = PVLTEngine(params, dict(self.named_buffers()), self.depths, self.loss_type, self.T_num,
                             self.drop_path_rate, self.text_embeddings.dropout.p,
                             self.__dict__.setdefault("_step_counter", [0]))
            eng._device = dev
            eng._plist_id = plist
            self.__dict__["_eng"] = eng
            self.__dict__["_param_names"] = list(params.keys())
        return eng

    def enable_grad_sync(self, group=True, broadcast_from: Optional[int] = 0):
        """Data-parallel training without a DistributedDataParallel wrapper: the backward pass all-reduces (averages) its
        single flat gradient buffer over ``group`` (``True`` = the default process group) before handing the per-parameter
        views to autograd. ``broadcast_from``: rank whose parameters / buffers are copied to every rank first (DDP's
        construction-time broadcast); ``None`` skips it. ``enable_grad_sync(None)`` turns the exchange off."""
        import torch.distributed as dist
        self.__dict__["_grad_sync"] = group
        if group is not None and broadcast_from is not None:
            grp = None if group is True else group
            for t in list(self.parameters()) + list(self.buffers()):
                dist.broadcast(t.data, src=broadcast_from, group=grp)
            _lib.params_written()   # .data writes bypass the version counters the engine's bf16 weight copies are keyed on
        return self

    def state_dict(self, *args, **kwargs):
        eng = self.__dict__.get("_eng")
        if eng is not None and eng.t2i is not None:
            eng.t2i.sync_buffers()
        return super().state_dict(*args, **kwargs)

    def _apply(self, fn, *a, **kw):
        self.__dict__["_eng"] = None
        self.__dict__["_plist"] = None
        return super()._apply(fn, *a, **kw)

    # -- public API
    def forward(self, input_images, input_ids, **fused):
        """Returns the reference's logits dict (pvlt.py:358-401); disabled heads map to None. ``heads=("t2i",)`` (an
        extension) evaluates only the named heads: the others map to None and cost nothing.

        With label keyword arguments (``mlm_labels=..., itm_labels=..., target_images=...``) it takes the fused
        loss path instead (see ``forward_losses``) -- routed through ``forward`` so that DistributedDataParallel's
        forward hook arms its gradient reducer for it as well."""
        heads = fused.pop("heads", None)
        if fused:
            return self.forward_losses(input_images, input_ids, **fused)
        self._engine()
        params = self._params_list()[1]
        mlm, itm, sup, sub, t2i = _PVLTFunction.apply(self, ("logits", torch.is_grad_enabled(), None if heads is None else tuple(heads)),
                                                      None, input_images, input_ids, *params)
        lt = self.loss_type
        if heads is not None:
            lt = {key: (v if key in heads else 0) for key, v in lt.items()}
        return dict(mlm_logits=mlm if lt['mlm'] else None, itm_logits=itm if lt['itm'] else None,
                    sup_cls_logits=sup if lt['cls'] else None, sub_cls_logits=sub if lt['cls'] else None,
                    t2i_logits=t2i if lt['t2i'] else None)

    @torch.no_grad()
    def itm_logits(self, input_images, input_ids):
        """Retrieval fast path: encoder + ITM head only (the reference also runs its unused MLM / t2i heads,
        engine_grid_masking.py:356-358). Returns fp32 [B, 2]."""
        if not self.loss_type['itm']:
            raise MvltError("this model was built without the ITM head")
        eng = self._engine()
        eng.prepare_weights()
        images = input_images.contiguous().to(F32)
        enc = eng.encoder_fwd(images, input_ids, False, False)
        X4, HW4 = _heads_common_fwd(eng, enc, images.shape[0])
        lg, _ = eng.small_head_fwd(X4, images.shape[0], HW4, "itm")
        return lg

    def forward_losses(self, input_images, input_ids, *, mlm_labels=None, itm_labels=None, sup_cls_labels=None,
                       sub_cls_labels=None, target_images=None, weights: Optional[Dict[str, float]] = None,
                       mlm_count: Optional[int] = None, only=None):
        """Fused step: heads + losses of engine_grid_masking.py:81-102. Returns (total_loss, stats[8])."""
        self._engine()
        params = self._params_list()[1]
        batch = dict(mlm_labels=mlm_labels, itm_labels=itm_labels, sup_cls_labels=sup_cls_labels,
                     sub_cls_labels=sub_cls_labels, target_images=target_images, weights=weights or {},
                     mlm_count=mlm_count, only=only)
        return _PVLTFunction.apply(self, ("losses", torch.is_grad_enabled(), None), batch, input_images, input_ids, *params)


def _cfg(**kwargs):
    return dict(url='', num_classes=1000, input_size=(3, 224, 224), pool_size=None, crop_pct=.9,
                interpolation='bicubic', first_conv='patch_embed.proj', classifier='head', **kwargs)


def _make(name, pretrained, token_hidden_size, num_text_tokens, loss_type, pretrained_pth, **kwargs):
    kwargs = {kk: v for kk, v in kwargs.items() if v is not None}   # timm's create_model strips None kwargs
    model = PyramidVisionLanguageTransformer(
        patch_size=4, embed_dims=[64, 128, 320, 512], num_heads=[1, 2, 5, 8], mlp_ratios=[8, 8, 4, 4], qkv_bias=True,
        norm_layer=partial(nn.LayerNorm, eps=1e-6), depths=DEPTHS[name], sr_ratios=[8, 4, 2, 1],
        token_hidden_size=token_hidden_size, num_text_tokens=num_text_tokens, loss_type=loss_type, **kwargs)
    model.default_cfg = _cfg()
    if pretrained_pth:
        from ..utils import load_file
        model.load_state_dict(load_file(pretrained_pth), strict=False)   # tensors only (weights_only=True)
        print('>>> load pretrained weights (backbone part) from:', pretrained_pth)
    return model


def pvlt_tiny(pretrained, token_hidden_size, num_text_tokens, loss_type, pretrained_pth, **kwargs):
    return _make("pvlt_tiny", pretrained, token_hidden_size, num_text_tokens, loss_type, pretrained_pth, **kwargs)


def pvlt_small(pretrained, token_hidden_size, num_text_tokens, loss_type, pretrained_pth, **kwargs):
    return _make("pvlt_small", pretrained, token_hidden_size, num_text_tokens, loss_type, pretrained_pth, **kwargs)


def pvlt_medium(pretrained, token_hidden_size, num_text_tokens, loss_type, pretrained_pth, **kwargs):
    return _make("pvlt_medium", pretrained, token_hidden_size, num_text_tokens, loss_type, pretrained_pth, **kwargs)


def pvlt_large(pretrained, token_hidden_size, num_text_tokens, loss_type, pretrained_pth, **kwargs):
    return _make("pvlt_large", pretrained, token_hidden_size, num_text_tokens, loss_type, pretrained_pth, **kwargs)


try:  # register with timm when it is installed, so that timm.create_model('pvlt_tiny', ...) resolves (main_vl.py:259)
    from timm.models.registry import register_model as _register
    for _f in (pvlt_tiny, pvlt_small, pvlt_medium, pvlt_large):
        _register(_f)
except Exception:  # timm is absent in this image; mvlt_b200.create_model is the stand-in
    pass
