"""Train / eval loops with the reference's names and signatures (/root/reference/engine_grid_masking.py), driving the
B200-native PVLT model. ``main_vl.py`` imports ``train_one_epoch_vl, evaluate_vl, evaluate_retrieval,
evaluate_recognition, visual_vl`` from this module (main_vl.py:198).

Differences that are deliberate (and documented in DESIGN.md):
  * the training step uses the model's fused loss path (heads + losses inside the model's single autograd node; the MLM
    head runs on labelled rows only) instead of materialising [B,128,30522] logits -- loss values and gradients match;
  * grid masks may be generated on the device (``mvlt_b200.masking``) when the batch carries no ``masked_images``;
  * with ``t2i`` disabled the published engine never calls the model on odd steps and crashes;
    here every step runs a forward;
  * retrieval shards the candidate pairs across ranks and runs encoder + ITM head only;
  * one device->host read per step for logging instead of seven ``.item()`` calls plus ``cuda.synchronize()``.
"""
from __future__ import annotations

import math
from typing import Iterable, Optional

import torch

from mvlt_b200 import masking, retrieval
from mvlt_b200.libs.vl_scores import compute_psnr
from mvlt_b200.utils import MetricLogger, SmoothedValue, is_dist

MLM_LOSS_WEIGHT, ITM_LOSS_WEIGHT, T2I_LOSS_WEIGHT = 1, 1, 10
USE_ORI_INPUT_IDS = False


def _net(model):
    return model.module if hasattr(model, "module") else model


def _dev(x, device):
    return x.to(device, non_blocking=True) if torch.is_tensor(x) else x


def _masked(samples, images, device, step, seed=0, out=None):
    """``out``: optional static device buffer the masked batch is written into (CUDA-graph mode)."""
    if "masked_images" in samples:
        if out is not None:
            return out.copy_(samples["masked_images"], non_blocking=True)
        return _dev(samples["masked_images"], device)
    B = images.shape[0]
    seeds = [masking.sample_seed(seed, step * B + i) for i in range(B)]
    return masking.apply_grid_mask(images, masking.grid_mask_batch(seeds, (images.shape[3], images.shape[2]), 0.5, 16, device),
                                   out=out)


def _own_adamw(optimizer):
    from mvlt_b200.optim import AdamW
    return isinstance(optimizer, AdamW)


def _graphed_step(model, optimizer, loss_scaler, max_norm, model_ema, args, batch_rows):
    """CUDA-graph replay of the iteration (mvlt_b200/graph.py) when asked for (``args.cuda_graph`` or MVLT_CUDA_GRAPH=1) and
    possible: own AdamW, no loss scaler / EMA hooks between backward and step, no DistributedDataParallel wrapper
    (``model.enable_grad_sync()`` is the data-parallel mode that can be captured); gradient clipping (``max_norm``) is captured
    with the step (AdamW.step(max_norm=...): norm and coefficient stay on the device). One GraphedStep per model, kept across epochs."""
    import os
    want = bool(getattr(args, "cuda_graph", False)) or os.environ.get("MVLT_CUDA_GRAPH", "0") == "1"
    if not want or loss_scaler is not None or model_ema is not None or hasattr(model, "module"):
        return None
    from mvlt_b200.graph import GraphedStep
    from mvlt_b200.optim import AdamW
    if not isinstance(optimizer, AdamW):
        return None
    gs = model.__dict__.get("_graphed_step")
    if gs is None or gs.opt is not optimizer or gs.model._engine() is not gs.eng or gs.max_norm != (float(max_norm) if max_norm else None):
        cap = getattr(args, "mlm_capacity", None)
        if not cap and model.loss_type.get("mlm"):      # 15 % of the word pieces are masked: 20 % of all rows bounds the count
            cap = -(-int(0.2 * batch_rows) // 128) * 128
        gs = GraphedStep(model, optimizer, mlm_capacity=cap, warmup=1, max_norm=max_norm or None)
        model.__dict__["_graphed_step"] = gs
        model.__dict__["_graphed_static"] = {}
    return gs


def _static_copy(model, name, src, device):
    """The captured graph reads its inputs from fixed device buffers: copy the batch into them (H2D for host batches)."""
    bufs = model.__dict__["_graphed_static"]
    key = (name, tuple(src.shape), src.dtype)       # one buffer per shape: a short last batch must not move the full-size one
    t = bufs.get(key)
    if t is None:
        t = bufs[key] = torch.empty(src.shape, dtype=src.dtype, device=device)
    t.copy_(src, non_blocking=True)
    return t


def train_one_epoch_vl(model: torch.nn.Module, criterion, data_loader: Iterable, optimizer: torch.optim.Optimizer,
                       device: torch.device, epoch: int, loss_scaler=None, max_norm: float = 0, model_ema=None,
                       mixup_fn=None, set_training_mode=True, fp32=False, args=None):
    """One epoch of engine_grid_masking.py:27-150. ``criterion`` / ``mixup_fn`` are accepted and unused (as in the
    reference, where DistillationLoss is constructed but never called)."""
    model.train(set_training_mode)
    logger = MetricLogger(delimiter="  ")
    logger.add_meter("lr", SmoothedValue(window_size=1, fmt="{value:.6f}"))
    loss_type = _net(model).loss_type
    weights = {"mlm": MLM_LOSS_WEIGHT, "itm": ITM_LOSS_WEIGHT, "t2i": T2I_LOSS_WEIGHT, "cls": 1}
    gstep = None
    for idx, samples in enumerate(logger.log_every(data_loader, 10, f"Epoch: [{epoch}]")):
        if idx == 0 and set_training_mode:
            B0, T0 = samples["input_ids"].shape[:2]
            gstep = _graphed_step(model, optimizer, loss_scaler, max_norm, model_ema, args, B0 * T0)
        masked_step = idx % 2 == 1 and bool(loss_type.get("t2i"))      # :72-78 odd steps feed the grid-masked image
        if gstep is not None:
            # whole iteration (forward, losses, backward, gradient exchange, AdamW, zero_grad) as one captured graph per
            # (masked?, batch shape): the batch is copied into static device buffers, per-step scalars live on the device
            net = _net(model)
            images = _static_copy(net, "image", samples["image"], device)
            ids = _static_copy(net, "ids", samples["ori_input_ids" if USE_ORI_INPUT_IDS else "input_ids"], device)
            labels = {k: _static_copy(net, k, samples[k], device)
                      for k in ("mlm_labels", "itm_labels", "sup_cls_labels", "sub_cls_labels") if samples.get(k) is not None}
            x = images
            if masked_step:
                bufs = net.__dict__["_graphed_static"]
                mkey = ("masked", tuple(images.shape), images.dtype)
                if mkey not in bufs:
                    bufs[mkey] = torch.empty_like(images)
                x = _masked(samples, images, device, idx, seed=epoch, out=bufs[mkey])
            total, stats = gstep(x, ids, key=(masked_step, tuple(images.shape)), target_images=images, weights=weights, **labels)
        else:
            images = _dev(samples["image"], device)
            ids = _dev(samples["ori_input_ids" if USE_ORI_INPUT_IDS else "input_ids"], device)
            x = images
            if masked_step:
                x = _masked(samples, images, device, idx, seed=epoch)
            mlm_labels = samples.get("mlm_labels")
            total, stats = model(x, ids, mlm_labels=mlm_labels, itm_labels=samples.get("itm_labels"),
                                 sup_cls_labels=samples.get("sup_cls_labels"), sub_cls_labels=samples.get("sub_cls_labels"),
                                 target_images=images, weights=weights)
            optimizer.zero_grad()
            if loss_scaler is not None:   # timm NativeScaler call convention (engine_grid_masking.py:126-127)
                loss_scaler(total, optimizer, clip_grad=max_norm if max_norm else None, parameters=model.parameters(),
                            create_graph=False)
            else:
                total.backward()
                if max_norm and _own_adamw(optimizer):
                    optimizer.step(max_norm=max_norm)       # norm + clip coefficient on the device, folded into the AdamW read
                else:
                    if max_norm:
                        torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm)
                    optimizer.step()
        if model_ema is not None:
            model_ema.update(model)
        s = stats.tolist()                                   # the step's single device->host read
        if not math.isfinite(s[0]):
            print(f" [ Warning!!! ] Total Loss is {s[0]} (mlm={s[1]} | itm={s[2]} | sup_cls={s[3]} | sub_cls={s[4]} | "
                  f"t2i={s[5]}), raise NaN value")
        logger.update(total_loss=s[0], loss_mlm=s[1], loss_itm=s[2], loss_sup_cls=s[3], loss_sub_cls=s[4], loss_t2i=s[5])
        logger.update(lr=optimizer.param_groups[0]["lr"])
    if gstep is not None and gstep.check_overflow():
        from mvlt_b200._lib import MvltError
        raise MvltError(f"a batch held more MLM-labelled tokens than the captured step's row capacity ({gstep.state.mlm_cap}): "
                        "set args.mlm_capacity higher")
    logger.synchronize_between_processes()
    print("Averaged stats:", logger)
    return {k: m.global_avg for k, m in logger.meters.items()}


@torch.no_grad()
def evaluate_vl(data_loader, model, device, args):
    """engine_grid_masking.py:153-333: MLM accuracy on masked ids, ITM / category accuracy on original ids, PSNR of
    the reconstruction from the masked image, summed loss."""
    logger = MetricLogger(delimiter="  ")
    model.eval()
    net = _net(model)
    lt = net.loss_type
    for step, samples in enumerate(logger.log_every(data_loader, 10, "Test:")):
        images = _dev(samples["image"], device)
        B = images.shape[0]
        total = 0.0
        mlm_acc = itm_acc = sup_acc = sub_acc = psnr = 0.0
        if lt.get("mlm"):
            _, st = net.forward_losses(images, _dev(samples["input_ids"], device), mlm_labels=samples["mlm_labels"],
                                       only=("mlm",))
            s = st.tolist()
            total += s[1]
            mlm_acc = s[6] / max(s[7], 1.0)
        if lt.get("itm") or lt.get("cls"):
            _, st = net.forward_losses(images, _dev(samples["ori_input_ids"], device), itm_labels=samples.get("itm_labels"),
                                       sup_cls_labels=samples.get("sup_cls_labels"),
                                       sub_cls_labels=samples.get("sub_cls_labels"), only=("itm", "cls"))
            s = st.tolist()
            total += s[2] + s[3] + s[4]
            itm_acc, sup_acc, sub_acc = s[8] / B, s[9] / B, s[10] / B
        if lt.get("t2i"):
            x = _masked(samples, images, device, step)
            out = net(x, _dev(samples["ori_input_ids"], device), heads=("t2i",))   # no [B, 128, 30522] MLM logits for a PSNR
            if out["t2i_logits"] is None:
                raise Exception("t2i_logits is none, please check the settings!")
            psnr = compute_psnr(out["t2i_logits"], images)
            mse = 10 ** (-(psnr - 20 * math.log10(255.0)) / 10) if psnr != 100 else 0.0
            del out
            _, st = net.forward_losses(x, _dev(samples["ori_input_ids"], device), target_images=images, only=("t2i",))
            total += st.tolist()[5]
        logger.meters["mlm_acc"].update(mlm_acc, n=B)
        logger.meters["itm_acc"].update(itm_acc, n=B)
        logger.meters["sup_cls_acc"].update(sup_acc, n=B)
        logger.meters["sub_cls_acc"].update(sub_acc, n=B)
        logger.meters["t2i_psnr"].update(psnr, n=B)
        logger.update(total_loss=total, n=B)
    logger.synchronize_between_processes()
    m = logger.meters
    print("** mlm@acc {:.5f} itm@acc {:.5f} sup_cls@acc {:.5f} sub_cls@acc {:.5f} t2i@psnr {:.5f} total_loss {:.5f}".format(
        m["mlm_acc"].global_avg, m["itm_acc"].global_avg, m["sup_cls_acc"].global_avg, m["sub_cls_acc"].global_avg,
        m["t2i_psnr"].global_avg, m["total_loss"].global_avg))
    return {k: meter.global_avg for k, meter in logger.meters.items()}


@torch.no_grad()
def evaluate_retrieval(data_loader, model, device, args):
    """engine_grid_masking.py:336-393: rank of candidate 0 among the 101 pairs of each query; acc@1/5/10.
    The pairs of every query are block-partitioned across the ranks (mvlt_b200.retrieval)."""
    model.eval()
    net = _net(model)
    rank = torch.distributed.get_rank() if is_dist() else 0
    world = torch.distributed.get_world_size() if is_dist() else 1
    logger = MetricLogger(delimiter="  ")
    counts = {1: 0, 5: 0, 10: 0}
    n_query = 0
    for samples in logger.log_every(data_loader, 10, "Test:"):
        images = _dev(samples["images_101"], device).squeeze(0)
        ids = _dev(samples["ori_input_ids_101"], device).squeeze(0)
        n_cand = images.shape[0]
        ranks, _ = retrieval.rank_queries(net, images, ids, n_cand, rank, world)
        r = int(ranks[0].item())
        n_query += 1
        for kk in counts:
            counts[kk] += int(r < kk)
    flag = "TIR" if getattr(args, "eval_retrieval_tir", False) else "ITR"
    denom = max(n_query, 1)   # the reference hard-codes 1000 (engine_grid_masking.py:393)
    print("\n", "#" * 30, "retrieval evaluation", "#" * 30)
    print(">>> retrieval {}: acc@1: {}, acc@5: {}, acc@10: {}".format(flag, counts[1] / denom, counts[5] / denom, counts[10] / denom))
    return {f"acc@{kk}": v / denom for kk, v in counts.items()}


@torch.no_grad()
def evaluate_recognition(data_loader, model, device, args):
    """engine_grid_masking.py:396-462: argmax of the 48 / 122-way heads; accuracy and macro / micro / weighted F1."""
    model.eval()
    net = _net(model)
    sup_l, sup_p, sub_l, sub_p = [], [], [], []
    for samples in data_loader:
        out = net(_dev(samples["images"], device), _dev(samples["ori_input_ids"], device))
        sup_p += out["sup_cls_logits"].view(-1, 48).argmax(-1).cpu().tolist()
        sub_p += out["sub_cls_logits"].view(-1, 122).argmax(-1).cpu().tolist()
        sup_l += samples["sup_cls_labels"].view(-1).tolist()
        sub_l += samples["sub_cls_labels"].view(-1).tolist()
    sup_m, sub_m = calculate_cls_metrics(sup_l, sup_p), calculate_cls_metrics(sub_l, sub_p)
    print("\n", "#" * 30, "recognition evaluation", "#" * 30)
    print("> logging-sup: accuracy ({}) macro_f1 ({}) micro_f1 ({}) weighted_f1 ({})\n"
          "> logging-sub: accuracy ({}) macro_f1 ({}) micro_f1 ({}) weighted_f1 ({})".format(*sup_m, *sub_m))
    return {"sup": sup_m, "sub": sub_m}


def calculate_cls_metrics(cls_labels, preds):
    """engine_grid_masking.py:465-474 (sklearn metrics on host integer lists)."""
    from sklearn.metrics import accuracy_score, f1_score
    return (accuracy_score(cls_labels, preds), f1_score(cls_labels, preds, average="macro"),
            f1_score(cls_labels, preds, average="micro"), f1_score(cls_labels, preds, average="weighted"))


def visual_vl(data_loader, model, device, args):
    """The reference's debug image dump (engine_grid_masking.py:502-685) reads keys its datasets no longer emit
    (SURVEY 2.1 #4); out of scope for the hot path."""
    raise NotImplementedError("visual_vl is a debug visualisation outside the PVLT hot path (see DESIGN.md, out of scope)")
