"""ORACLE (test infrastructure, never shipped on the product path).

CPU fp32 restatement, in plain functional PyTorch, of the reference's PVLT hot path. It is what the CUDA
kernels are checked against in ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` -- nothing under ``mvlt_b200/`` may import it.

Pinned against the real reference: ``tests/golden/make_golden.py`` imports /root/reference/libs/pvlt.py in
the build container, loads it with ``make_state_dict`` weights and stores its outputs / losses / gradient
summaries in ``tests/golden/pvlt_tiny_golden.npz``; ``tests/test_oracle_cpu.py`` replays them through this
file.

Each function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

# libs/pvlt.py:415-483 -- the four registered variants differ only in depths
ARCH = {
    "pvlt_tiny": dict(depths=[2, 2, 2, 2]),
    "pvlt_small": dict(depths=[3, 4, 6, 3]),
    "pvlt_medium": dict(depths=[3, 4, 18, 3]),
    "pvlt_large": dict(depths=[3, 8, 27, 3]),
}
EMBED_DIMS = [64, 128, 320, 512]
NUM_HEADS = [1, 2, 5, 8]
MLP_RATIOS = [8, 8, 4, 4]
SR_RATIOS = [8, 4, 2, 1]
PATCH = [4, 2, 2, 2]
IMG_SIZE_DEFAULT = 224        # pvlt.py:179 -- never overridden by main_vl.py, so pos_embeds are 56/28/14/7 (+1)
VOCAB, HIDDEN, MAX_POS = 30522, 768, 512
T2I_CH = 64


# --------------------------------------------------------------------------------------------------------
# Procedural weights: same names/shapes as the reference state_dict, values drawn from
# the reference's init DISTRIBUTIONS (pvlt.py:228-229,282-289; torch defaults for conv/BN) with a private
# per-tensor seed, so that the very same tensors can be rebuilt on the GPU box without the reference.
# --------------------------------------------------------------------------------------------------------
def _gen(seed: int, name: str) -> torch.Generator:
    h = 1469598103934665603
    for ch in f"{seed}:{name}".encode():
        h = ((h ^ ch) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return torch.Generator().manual_seed(h & 0x7FFFFFFFFFFFFFFF)


def _tn(shape, seed, name, std=0.02):
    return torch.randn(shape, generator=_gen(seed, name)).clamp_(-2 / std, 2 / std) * std


def _uni(shape, seed, name, bound):
    return (torch.rand(shape, generator=_gen(seed, name)) * 2 - 1) * bound


def make_state_dict(model: str = "pvlt_tiny", loss_type: Optional[Dict[str, int]] = None, seed: int = 0,
                    perturb: float = 0.02) -> Dict[str, torch.Tensor]:
    """state_dict with the reference's key names. ``perturb`` adds noise to LN/BN affine params and biases
    (which the reference initialises to exactly 1/0) so that parity tests exercise them."""
    loss_type = loss_type or {"itm": 1, "mlm": 1, "t2i": 1, "cls": 0}
    depths = ARCH[model]["depths"]
    sd: Dict[str, torch.Tensor] = {}

    def linear(prefix, out_f, in_f):
        sd[prefix + ".weight"] = _tn((out_f, in_f), seed, prefix + ".weight")
        sd[prefix + ".bias"] = _tn((out_f,), seed, prefix + ".bias", perturb) if perturb else torch.zeros(out_f)

    def lnorm(prefix, c):
        sd[prefix + ".weight"] = 1.0 + (_tn((c,), seed, prefix + ".weight", perturb) if perturb else 0)
        sd[prefix + ".bias"] = _tn((c,), seed, prefix + ".bias", perturb) if perturb else torch.zeros(c)

    def conv(prefix, co, ci, k, bias=True):
        bound = 1.0 / math.sqrt(ci * k * k)   # torch Conv2d default: kaiming_uniform(a=sqrt(5))
        sd[prefix + ".weight"] = _uni((co, ci, k, k), seed, prefix + ".weight", bound)
        if bias:
            sd[prefix + ".bias"] = _uni((co,), seed, prefix + ".bias", bound)

    for i in range(4):
        s = i + 1
        C = EMBED_DIMS[i]
        cin = 3 if i == 0 else EMBED_DIMS[i - 1]
        side = IMG_SIZE_DEFAULT // 4 if i == 0 else (IMG_SIZE_DEFAULT // (2 ** (i + 1))) // 2
        npatch = side * side + (1 if i == 3 else 0)
        sd[f"pos_embed{s}"] = _tn((1, npatch, C), seed, f"pos_embed{s}")
        sd[f"text_pos_embed{s}"] = _tn((1, 128, C), seed, f"text_pos_embed{s}")
        conv(f"patch_embed{s}.proj", C, cin, PATCH[i])
        lnorm(f"patch_embed{s}.norm", C)
        linear(f"text_embed{s}.0", C, HIDDEN if i == 0 else cin)
        lnorm(f"text_embed{s}.1", C)
        for j in range(depths[i]):
            p = f"block{s}.{j}"
            lnorm(p + ".norm1", C)
            lnorm(p + ".norm2", C)
            linear(p + ".attn.q", C, C)
            linear(p + ".attn.kv", 2 * C, C)
            linear(p + ".attn.proj", C, C)
            if SR_RATIOS[i] > 1:
                conv(p + ".attn.sr", C, C, SR_RATIOS[i])
                lnorm(p + ".attn.norm", C)
            linear(p + ".mlp.fc1", C * MLP_RATIOS[i], C)
            linear(p + ".mlp.fc2", C, C * MLP_RATIOS[i])
    # BertEmbeddings (transformers); word embedding is tied to the MLM decoder and therefore re-initialised
    # trunc-normal(0.02) by pvlt.py:280-284
    we = _tn((VOCAB, HIDDEN), seed, "word_embeddings")
    sd["text_embeddings.word_embeddings.weight"] = we
    sd["text_embeddings.position_embeddings.weight"] = _tn((MAX_POS, HIDDEN), seed, "position_embeddings", 1.0) * 0.05
    sd["text_embeddings.token_type_embeddings.weight"] = _tn((2, HIDDEN), seed, "token_type_embeddings", 1.0) * 0.05
    lnorm("text_embeddings.LayerNorm", HIDDEN)
    # (position_ids / token_type_ids are non-persistent buffers in transformers >= 4.31: not in the state_dict)

    def head_embed(prefix):
        linear(prefix + ".0", HIDDEN, EMBED_DIMS[-1])
        lnorm(prefix + ".1", HIDDEN)

    if loss_type.get("mlm"):
        head_embed("mlm_head_embed")
        sd["mlm_head.bias"] = _tn((VOCAB,), seed, "mlm_head.bias", perturb) if perturb else torch.zeros(VOCAB)
        linear("mlm_head.transform.dense", HIDDEN, HIDDEN)
        lnorm("mlm_head.transform.LayerNorm", HIDDEN)
        sd["mlm_head.mlm_decoder.weight"] = we  # same storage (vl_heads.py:62)
    if loss_type.get("itm"):
        head_embed("itm_head_embed")
        linear("itm_head.linear", 2, HIDDEN)
        sd["itm_head.linear_bias"] = _tn((2,), seed, "itm_head.linear_bias", perturb) if perturb else torch.zeros(2)
    if loss_type.get("cls"):
        for nm, n in (("sup", 48), ("sub", 122)):
            head_embed(f"{nm}_cls_head_embed")
            linear(f"{nm}_cls_head.linear", n, HIDDEN)
            sd[f"{nm}_cls_head.linear_bias"] = (_tn((n,), seed, f"{nm}_cls_head.linear_bias", perturb)
                                                if perturb else torch.zeros(n))
    if loss_type.get("t2i"):
        ch = T2I_CH
        convbn = [("reduction1", EMBED_DIMS[1], ch), ("reduction2", EMBED_DIMS[2], ch),
                  ("reduction3", EMBED_DIMS[3], ch), ("conv_upsample1", ch, ch), ("conv_upsample2", ch, ch),
                  ("conv_upsample3", ch, ch), ("conv_upsample4", ch, ch), ("conv_upsample5", 2 * ch, 2 * ch),
                  ("conv_concat2", 2 * ch, 2 * ch), ("conv_concat3", 3 * ch, 3 * ch), ("conv4", 3 * ch, 3 * ch)]
        for nm, ci, co in convbn:
            p = f"t2i_head.{nm}"
            conv(p + ".0", co, ci, 3, bias=False)
            sd[p + ".1.weight"] = 1.0 + (_tn((co,), seed, p + ".1.weight", perturb) if perturb else 0)
            sd[p + ".1.bias"] = _tn((co,), seed, p + ".1.bias", perturb) if perturb else torch.zeros(co)
            sd[p + ".1.running_mean"] = torch.zeros(co)
            sd[p + ".1.running_var"] = torch.ones(co)
            sd[p + ".1.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
        conv("t2i_head.score.0", 3, 3 * ch, 1)
    return sd


# --------------------------------------------------------------------------------------------------------
# Synthetic inputs
# --------------------------------------------------------------------------------------------------------
def make_inputs(batch: int, seed: int = 0, T: int = 128, img: int = 256):
    g = torch.Generator().manual_seed(1000 + seed)
    images = torch.rand((batch, 3, img, img), generator=g)
    ids = torch.zeros((batch, T), dtype=torch.long)
    ori = torch.zeros((batch, T), dtype=torch.long)
    mlm = torch.full((batch, T), -1, dtype=torch.long)
    for b in range(batch):
        L = int(torch.randint(16, 65, (1,), generator=g))
        toks = torch.randint(1000, VOCAB, (L - 2,), generator=g)
        row = torch.cat([torch.tensor([101]), toks, torch.tensor([102])])
        ori[b, :L] = row
        r = torch.rand((L,), generator=g)
        r2 = torch.rand((L,), generator=g)
        rnd = torch.randint(1000, VOCAB, (L,), generator=g)
        new = row.clone()
        for t in range(1, L - 1):          # fashion_gen.py:383-409: 15 % -> 80 % [MASK] / 10 % random / 10 % keep
            if r[t] < 0.15:
                mlm[b, t] = row[t]
                if r2[t] < 0.8:
                    new[t] = 103
                elif r2[t] < 0.9:
                    new[t] = rnd[t]
        if (mlm[b] != -1).sum() == 0:      # keep CE well defined (reference would give NaN)
            mlm[b, 1] = row[1]
            new[1] = 103
        ids[b, :L] = new
    itm = torch.randint(0, 2, (batch, 1), generator=g)
    sup = torch.randint(0, 48, (batch, 1), generator=g)
    sub = torch.randint(0, 122, (batch, 1), generator=g)
    return dict(images=images, input_ids=ids, ori_input_ids=ori, mlm_labels=mlm, itm_labels=itm,
                sup_cls_labels=sup, sub_cls_labels=sub)


# --------------------------------------------------------------------------------------------------------
# Forward restatement
# --------------------------------------------------------------------------------------------------------
def _ln(x, sd, prefix, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], eps)


def _lin(x, sd, prefix):
    return F.linear(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"))


def bert_embeddings(sd, ids, p_drop=0.0, training=False, keep_scale=None):
    """transformers BertEmbeddings.forward (absolute positions, token type 0); call site pvlt.py:326.
    ``keep_scale`` [B, T, 768]: an explicit dropout mask (0 or 1/(1-p)) instead of torch's RNG draw."""
    T = ids.shape[1]
    # padding_idx=0 (BertConfig.pad_token_id): pad rows get no gradient from the gather
    x = (F.embedding(ids, sd["text_embeddings.word_embeddings.weight"], padding_idx=0)
         + sd["text_embeddings.token_type_embeddings.weight"][0]
         + sd["text_embeddings.position_embeddings.weight"][:T])
    x = _ln(x, sd, "text_embeddings.LayerNorm", 1e-12)
    if keep_scale is not None:
        return x * keep_scale
    return F.dropout(x, p_drop, training)


def resized_pos_embed(sd, stage: int, H: int, W: int):
    """pvlt.py:291-297 + :341-344. The comparison is against the STAGE-1 patch count (quirk, Appendix D)."""
    pe = sd[f"pos_embed{stage}"]
    if stage == 4:
        pe = pe[:, 1:]
    side = int(round(math.sqrt(pe.shape[1])))
    n1 = sd["pos_embed1"].shape[1]
    if H * W == n1:
        return pe
    C = pe.shape[-1]
    return F.interpolate(pe.reshape(1, side, side, C).permute(0, 3, 1, 2), size=(H, W),
                         mode="bilinear").reshape(1, C, H * W).permute(0, 2, 1)


def attention(sd, p, x, H, W, T, heads, sr):
    """pvlt.py:95-121."""
    B, N, C = x.shape
    hd = C // heads
    q = _lin(x, sd, p + ".q").reshape(B, N, heads, hd).permute(0, 2, 1, 3)
    if sr > 1:
        xi, xt = x[:, :H * W], x[:, H * W:]
        xi = xi.permute(0, 2, 1).reshape(B, C, H, W)
        xi = F.conv2d(xi, sd[p + ".sr.weight"], sd[p + ".sr.bias"], stride=sr).reshape(B, C, -1).permute(0, 2, 1)
        xi = _ln(xi, sd, p + ".norm", 1e-5)
        kv_in = torch.cat((xi, xt), 1)
    else:
        kv_in = x
    kv = _lin(kv_in, sd, p + ".kv").reshape(B, -1, 2, heads, hd).permute(2, 0, 3, 1, 4)
    k, v = kv[0], kv[1]
    a = (q @ k.transpose(-2, -1)) * (hd ** -0.5)
    a = a.softmax(-1)
    o = (a @ v).transpose(1, 2).reshape(B, N, C)
    return _lin(o, sd, p + ".proj")


def block(sd, p, x, H, W, T, heads, sr, dp_scale=None):
    """pvlt.py:140-144; dp_scale = optional per-sample drop-path factors (mask/keep) [2][B]."""
    a = attention(sd, p + ".attn", _ln(x, sd, p + ".norm1", 1e-6), H, W, T, heads, sr)
    if dp_scale is not None:
        a = a * dp_scale[0].view(-1, 1, 1)
    x = x + a
    h = _lin(_ln(x, sd, p + ".norm2", 1e-6), sd, p + ".mlp.fc1")
    h = _lin(F.gelu(h), sd, p + ".mlp.fc2")
    if dp_scale is not None:
        h = h * dp_scale[1].view(-1, 1, 1)
    return x + h


def pyramid_features(sd, images, ids, model="pvlt_tiny", T=128, dp_scales=None, embed_keep_scale=None):
    """pvlt.py:322-356. ``dp_scales`` [2 * n_blocks, B]: timm DropPath factors (mask / keep_prob) of the attention and MLP
    branch of every block in execution order (pvlt.py:141-142); ``embed_keep_scale``: BertEmbeddings dropout mask."""
    depths = ARCH[model]["depths"]
    B = images.shape[0]
    y = bert_embeddings(sd, ids, keep_scale=embed_keep_scale)
    blk = 0
    x = images
    img_feats, text_feats = [], []
    for i in range(4):
        s = i + 1
        x = F.conv2d(x, sd[f"patch_embed{s}.proj.weight"], sd[f"patch_embed{s}.proj.bias"], stride=PATCH[i])
        H, W = x.shape[2], x.shape[3]
        x = _ln(x.flatten(2).transpose(1, 2), sd, f"patch_embed{s}.norm", 1e-5)
        y = _ln(_lin(y, sd, f"text_embed{s}.0"), sd, f"text_embed{s}.1", 1e-5)
        pe = resized_pos_embed(sd, s, H, W)
        x = torch.cat((x + pe, y + sd[f"text_pos_embed{s}"]), 1)
        for j in range(depths[i]):
            dp = None if dp_scales is None else (dp_scales[2 * blk], dp_scales[2 * blk + 1])
            x = block(sd, f"block{s}.{j}", x, H, W, T, NUM_HEADS[i], SR_RATIOS[i], dp)
            blk += 1
        x, y = x[:, :H * W], x[:, H * W:]
        x = x.reshape(B, H, W, -1).permute(0, 3, 1, 2).contiguous()
        img_feats.append(x)
        text_feats.append(y)
    return img_feats, text_feats


def _head_embed(sd, prefix, x):
    return _ln(_lin(x, sd, prefix + ".0"), sd, prefix + ".1", 1e-5)


def mlm_head(sd, text_feat):
    """pvlt.py:368-369, vl_heads.py:30-35,65-70."""
    h = _head_embed(sd, "mlm_head_embed", text_feat)
    h = _lin(h, sd, "mlm_head.transform.dense")
    h = h * 0.5 * (1.0 + torch.erf(h / math.sqrt(2.0)))
    h = _ln(h, sd, "mlm_head.transform.LayerNorm", 1e-5)
    return F.linear(h, sd["mlm_head.mlm_decoder.weight"]) + sd["mlm_head.bias"]


def small_head(sd, name, text_feat):
    """ITM / CLS heads on text token 0: pvlt.py:375-388, vl_heads.py:84-87,101-104 (two biases)."""
    h = _head_embed(sd, f"{name}_head_embed", text_feat[:, 0:1, :])
    return _lin(h, sd, f"{name}_head.linear") + sd[f"{name}_head.linear_bias"]


def _convbn(sd, p, x, training, stats=None):
    x = F.conv2d(x, sd[p + ".0.weight"], None, padding=1)
    rm, rv = sd[p + ".1.running_mean"], sd[p + ".1.running_var"]
    if training:
        rm, rv = rm.clone(), rv.clone()   # do not mutate the caller's buffers
    y = F.batch_norm(x, rm, rv, sd[p + ".1.weight"], sd[p + ".1.bias"], training, 0.1, 1e-5)
    if stats is not None and training:
        stats[p] = (rm, rv)
    return y


def t2i_head(sd, low, mid, high, training=True, stats=None):
    """ITGHead: vl_heads.py:136-165 (no activations; BatchNorm in train mode uses batch statistics)."""
    up2 = lambda t: F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=True)
    cb = lambda n, t: _convbn(sd, "t2i_head." + n, t, training, stats)
    low, mid, high = cb("reduction1", low), cb("reduction2", mid), cb("reduction3", high)
    x1_1 = high
    x2_1 = cb("conv_upsample1", up2(x1_1)) * mid
    x3_1 = cb("conv_upsample2", up2(mid)) * cb("conv_upsample3", up2(x2_1)) * low
    x2_2 = cb("conv_concat2", torch.cat((x2_1, cb("conv_upsample4", up2(x1_1))), 1))
    x3_2 = cb("conv_concat3", torch.cat((x3_1, cb("conv_upsample5", up2(x2_2))), 1))
    r = cb("conv4", x3_2)
    r = F.conv2d(r, sd["t2i_head.score.0.weight"], sd["t2i_head.score.0.bias"])
    return F.interpolate(r, scale_factor=8, mode="bilinear", align_corners=True)


def forward(sd, images, ids, loss_type, model="pvlt_tiny", training=True, bn_stats=None, dp_scales=None,
            embed_keep_scale=None):
    """pvlt.py:358-401: the logits dict with exactly the reference's keys."""
    img_feats, text_feats = pyramid_features(sd, images, ids, model, dp_scales=dp_scales, embed_keep_scale=embed_keep_scale)
    out = dict(mlm_logits=None, itm_logits=None, sup_cls_logits=None, sub_cls_logits=None, t2i_logits=None)
    tf = text_feats[-1]
    if loss_type.get("mlm"):
        out["mlm_logits"] = mlm_head(sd, tf)
    if loss_type.get("itm"):
        out["itm_logits"] = small_head(sd, "itm", tf)
    if loss_type.get("cls"):
        out["sup_cls_logits"] = small_head(sd, "sup_cls", tf)
        out["sub_cls_logits"] = small_head(sd, "sub_cls", tf)
    if loss_type.get("t2i"):
        out["t2i_logits"] = t2i_head(sd, img_feats[1], img_feats[2], img_feats[3], training, bn_stats)
    out["_img_feats"], out["_text_feats"] = img_feats, text_feats
    return out


def losses(out, batch, images_target):
    """engine_grid_masking.py:81-102: MLM x1 (ignore_index -1), ITM x1, cls x1 + x1, t2i x10 SmoothL1."""
    res = {}
    total = 0
    if out["mlm_logits"] is not None:
        res["mlm"] = F.cross_entropy(out["mlm_logits"].view(-1, VOCAB), batch["mlm_labels"].view(-1), ignore_index=-1)
        total = total + res["mlm"]
    if out["itm_logits"] is not None:
        res["itm"] = F.cross_entropy(out["itm_logits"].view(-1, 2), batch["itm_labels"].view(-1))
        total = total + res["itm"]
    if out["sup_cls_logits"] is not None:
        res["sup_cls"] = F.cross_entropy(out["sup_cls_logits"].view(-1, 48), batch["sup_cls_labels"].view(-1))
        res["sub_cls"] = F.cross_entropy(out["sub_cls_logits"].view(-1, 122), batch["sub_cls_labels"].view(-1))
        total = total + res["sup_cls"] + res["sub_cls"]
    if out["t2i_logits"] is not None:
        res["t2i"] = 10 * F.smooth_l1_loss(out["t2i_logits"], images_target)
        total = total + res["t2i"]
    res["total"] = total
    return res


def train_step_grads(sd, batch, loss_type, model="pvlt_tiny", masked_images=None, dp_scales=None, embed_keep_scale=None):
    """One fwd+bwd of engine_grid_masking.py:69-127 (no optimizer). Returns (losses, grads by name)."""
    names = [k for k, v in sd.items() if v.is_floating_point() and "running_" not in k
             and k != "mlm_head.mlm_decoder.weight"]
    leaf = {k: sd[k].detach().clone().requires_grad_(True) for k in names}
    sdl = dict(sd)
    sdl.update(leaf)
    if "mlm_head.mlm_decoder.weight" in sd:
        sdl["mlm_head.mlm_decoder.weight"] = leaf["text_embeddings.word_embeddings.weight"]
    x = masked_images if masked_images is not None else batch["images"]
    out = forward(sdl, x, batch["input_ids"], loss_type, model, training=True, dp_scales=dp_scales,
                  embed_keep_scale=embed_keep_scale)
    ls = losses(out, batch, batch["images"])
    ls["total"].backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaf.items()}
    return {k: float(v.detach()) for k, v in ls.items()}, grads, out


# --------------------------------------------------------------------------------------------------------
# Scores / retrieval (libs/vl_scores.py, engine_grid_masking.py:356-384)
# --------------------------------------------------------------------------------------------------------
def compute_mlm_score(logits, target, index=-1):
    """vl_scores.py:5-34."""
    preds = logits.argmax(-1)
    sel = target != index
    return float((preds[sel] == target[sel]).sum()) / max(int(sel.sum()), 1)


def compute_score_with_logits(logits, labels):
    """vl_scores.py:37-51 (multi-logit branch)."""
    return logits.max(1)[1] == labels


def compute_psnr(logits, labels):
    """vl_scores.py:54-63."""
    mse = float(torch.mean((logits - labels) ** 2))
    return 100 if mse == 0 else 20 * math.log10(255.0 / math.sqrt(mse))


def retrieval_rank(itm_logits):
    """engine_grid_masking.py:360-384: rank of candidate 0 among the softmax p(match), descending sort."""
    p = F.softmax(itm_logits.view(-1, 2).float(), -1)[:, 1]
    order = torch.sort(p, descending=True)[1]
    return int((order == 0).nonzero()[0, 0])


def evaluate_vl_batch(sd, samples, loss_type, model="pvlt_tiny"):
    """engine_grid_masking.py:186-300 for ONE batch (eval mode, fp32): three forwards (masked ids, original ids, masked
    image) -> metrics dict with the reference's meter names and the summed weighted loss (weights :23)."""
    B = samples["image"].shape[0]
    res = dict(mlm_acc=0.0, itm_acc=0.0, sup_cls_acc=0.0, sub_cls_acc=0.0, t2i_psnr=0.0)
    total = 0.0
    with torch.no_grad():
        o = forward(sd, samples["image"], samples["input_ids"], loss_type, model, training=False)
        if o["mlm_logits"] is not None:      # :201-208
            total += float(F.cross_entropy(o["mlm_logits"].view(-1, VOCAB), samples["mlm_labels"].view(-1), ignore_index=-1))
            res["mlm_acc"] = compute_mlm_score(o["mlm_logits"], samples["mlm_labels"])
        o = forward(sd, samples["image"], samples["ori_input_ids"], loss_type, model, training=False)
        if o["itm_logits"] is not None:      # :223-229
            total += float(F.cross_entropy(o["itm_logits"].view(-1, 2), samples["itm_labels"].view(-1)))
            res["itm_acc"] = float(compute_score_with_logits(o["itm_logits"].view(-1, 2), samples["itm_labels"].view(-1)).sum()) / B
        if o["sup_cls_logits"] is not None:  # :236-250
            total += float(F.cross_entropy(o["sup_cls_logits"].view(-1, 48), samples["sup_cls_labels"].view(-1)))
            total += float(F.cross_entropy(o["sub_cls_logits"].view(-1, 122), samples["sub_cls_labels"].view(-1)))
            res["sup_cls_acc"] = float(compute_score_with_logits(o["sup_cls_logits"].view(-1, 48), samples["sup_cls_labels"].view(-1)).sum()) / B
            res["sub_cls_acc"] = float(compute_score_with_logits(o["sub_cls_logits"].view(-1, 122), samples["sub_cls_labels"].view(-1)).sum()) / B
        if loss_type.get("t2i"):             # :305-316
            o = forward(sd, samples["masked_images"], samples["ori_input_ids"], loss_type, model, training=False)
            total += float(10 * F.smooth_l1_loss(o["t2i_logits"], samples["image"]))
            res["t2i_psnr"] = compute_psnr(o["t2i_logits"], samples["image"])
    res["total_loss"] = total
    return res


def fit_itm_probe(sd, images, ids, labels, model="pvlt_tiny", steps=200, l2=1e-4):
    """Planted-positive retrieval protocol (mvlt_b200/synthetic.py): K optimiser steps (L-BFGS, deterministic, zero init) on
    the LAST ITM linear layer only (vl_heads.py:84-87), on fp32 reference features of the fit set. Returns a copy of ``sd``
    with ``itm_head.linear.{weight,bias}`` replaced and ``itm_head.linear_bias`` zeroed."""
    with torch.no_grad():
        feats = []
        for i in range(0, images.shape[0], 32):
            _, tf = pyramid_features(sd, images[i:i + 32], ids[i:i + 32], model)
            feats.append(_head_embed(sd, "itm_head_embed", tf[-1][:, 0:1, :]).reshape(-1, HIDDEN))
        feats = torch.cat(feats)
    W = torch.zeros((2, HIDDEN), requires_grad=True)
    b = torch.zeros((2,), requires_grad=True)
    opt = torch.optim.LBFGS([W, b], lr=1.0, max_iter=steps)

    def closure():
        opt.zero_grad()
        loss = F.cross_entropy(feats @ W.t() + b, labels) + l2 * (W ** 2).sum()
        loss.backward()
        return loss
    opt.step(closure)
    out = dict(sd)
    out["itm_head.linear.weight"] = W.detach().clone()
    out["itm_head.linear.bias"] = b.detach().clone()
    out["itm_head.linear_bias"] = torch.zeros(2)
    return out


def fit_cls_probes(sd, images, ids, sup_labels, sub_labels, model="pvlt_tiny", steps=500, l2=1e-4):
    """Planted recognition protocol (mvlt_b200/synthetic.py:planted_cls_set): the LAST linear layer of the two category
    heads (vl_heads.py:101-104; 768 -> 48 / 122) fitted on fp32 reference features of the fit set (L-BFGS with a strong-Wolfe
    line search, zero init, deterministic). Returns a copy of ``sd`` with ``{sup,sub}_cls_head.linear.{weight,bias}`` replaced
    and ``linear_bias`` zeroed."""
    with torch.no_grad():
        tfs = []
        for i in range(0, images.shape[0], 32):
            _, tf = pyramid_features(sd, images[i:i + 32], ids[i:i + 32], model)
            tfs.append(tf[-1][:, 0:1, :])
        tf0 = torch.cat(tfs)
    out = dict(sd)
    for name, labels in (("sup_cls", sup_labels), ("sub_cls", sub_labels)):
        with torch.no_grad():
            feats = _head_embed(sd, f"{name}_head_embed", tf0).reshape(-1, HIDDEN)
        n_cls = sd[f"{name}_head.linear.weight"].shape[0]
        W = torch.zeros((n_cls, HIDDEN), requires_grad=True)
        b = torch.zeros((n_cls,), requires_grad=True)
        opt = torch.optim.LBFGS([W, b], lr=1.0, max_iter=steps, line_search_fn="strong_wolfe")
        y = labels.view(-1)

        def closure():
            opt.zero_grad()
            loss = F.cross_entropy(feats @ W.t() + b, y) + l2 * (W ** 2).sum()
            loss.backward()
            return loss
        opt.step(closure)
        out[f"{name}_head.linear.weight"] = W.detach().clone()
        out[f"{name}_head.linear.bias"] = b.detach().clone()
        out[f"{name}_head.linear_bias"] = torch.zeros(n_cls)
    return out
