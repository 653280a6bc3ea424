#!/usr/bin/env python
"""The reference's main_vl.py flow on synthetic Fashion-Gen-shaped data, end to end on the B200-native hot path.

  python examples/train_synthetic.py --epochs 2 --steps 6 --batch-size 32 --cuda-graph --clip-grad 1.0
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 examples/train_synthetic.py --cuda-graph

What a maintainer's main_vl.py does, in the same order and through the same names (file:line = /root/reference):
  main_vl.py:205-208   seeds (args.seed + rank)
  main_vl.py:259-270   create_model('pvlt_tiny', ..., loss_type=...)          -> mvlt_b200.create_model
  main_vl.py:298-302   DistributedDataParallel                                -> model.enable_grad_sync() (or the DDP wrapper)
  main_vl.py:308-310   create_optimizer / NativeScaler / create_scheduler     -> mvlt_b200.optim.AdamW + a cosine lambda schedule
  main_vl.py:420-447   train_one_epoch_vl, lr_scheduler.step, evaluate_vl     -> engine_grid_masking (this repository's)
  main_vl.py:352-372   evaluate_retrieval / evaluate_recognition              -> same
Datasets, tokenizer and checkpoint files are the reference's own business (out of scope); the loaders here are lists of
synthetic batches with the key set of mcloader/fashion_gen.py's sample dicts. Needs an sm_100 GPU: there is no CPU path.
"""
from __future__ import annotations

import argparse
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import engine_grid_masking as E  # noqa: E402
import mvlt_b200  # noqa: E402
from mvlt_b200.optim import AdamW, param_groups_no_decay  # noqa: E402
from mvlt_b200.synthetic import make_batch, planted_tir_query  # noqa: E402


def train_loader(n, B, seed0):
    out = []
    for i in range(n):
        b = make_batch(B, seed=seed0 + i)
        b["image"] = b.pop("images")          # fashion_gen.py:180-196 key names
        out.append(b)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="pvlt_tiny", choices=["pvlt_tiny", "pvlt_small", "pvlt_medium", "pvlt_large"])
    ap.add_argument("--task", default="pretrain", choices=["pretrain", "recognition"])
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--steps", type=int, default=6, help="batches per epoch")
    ap.add_argument("--batch-size", type=int, default=32)
    ap.add_argument("--lr", type=float, default=2.5e-4)
    ap.add_argument("--weight-decay", type=float, default=0.05)
    ap.add_argument("--clip-grad", type=float, default=None)
    ap.add_argument("--cuda-graph", action="store_true")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--retrieval-queries", type=int, default=4)
    ap.add_argument("--checkpoint", default=None, help="write a reference-format checkpoint here and resume from it (main_vl.py:327-346)")
    args = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=device)
    torch.manual_seed(args.seed + rank)

    args.loss_type = ({"itm": 1, "mlm": 1, "t2i": 1, "cls": 0} if args.task == "pretrain"
                      else {"itm": 0, "mlm": 0, "t2i": 0, "cls": 1})
    args.eval_retrieval_tir, args.eval_retrieval_itr = True, False
    model = mvlt_b200.create_model(args.model, pretrained=True, num_classes=1000, drop_rate=0.0, drop_path_rate=0.1,
                                   drop_block_rate=None, token_hidden_size=768, num_text_tokens=128,
                                   loss_type=dict(args.loss_type), pretrained_pth="").to(device)
    if world > 1:
        model.enable_grad_sync()              # parameters broadcast from rank 0; gradients averaged segment by segment
    optimizer = AdamW(param_groups_no_decay(model, args.weight_decay), lr=args.lr)
    total_steps = args.epochs * args.steps
    sched = torch.optim.lr_scheduler.LambdaLR(optimizer, lambda e: 0.5 * (1 + math.cos(math.pi * e / max(args.epochs, 1))))

    first = last = None
    for epoch in range(args.epochs):
        data = train_loader(args.steps, args.batch_size, seed0=1000 * rank + epoch * args.steps)
        stats = E.train_one_epoch_vl(model, None, data, optimizer, device, epoch, loss_scaler=None,
                                     max_norm=args.clip_grad or 0, args=args)
        sched.step()
        first = first or stats
        last = stats
        if rank == 0:
            norm = optimizer.last_grad_norm
            print(f"epoch {epoch}: total_loss {stats['total_loss']:.4f}  lr {optimizer.param_groups[0]['lr']:.2e}"
                  + (f"  grad_norm {float(norm):.3f}" if norm is not None else ""))

    if args.task == "pretrain":
        ev = E.evaluate_vl(train_loader(2, args.batch_size, seed0=777), model, device, args)
        loader = []
        for q in range(args.retrieval_queries):
            img, ids, _ = planted_tir_query(q)
            loader.append({"images_101": img.unsqueeze(0), "ori_input_ids_101": ids.unsqueeze(0), "info_list": []})
        rt = E.evaluate_retrieval(loader, model, device, args)
        if rank == 0:
            print("evaluate_vl:", {k: round(v, 4) for k, v in ev.items()})
            print("evaluate_retrieval:", rt)
    else:
        loader = []
        for i in range(2):
            b = make_batch(args.batch_size, seed=900 + i)
            loader.append(b)
        rc = E.evaluate_recognition(loader, model, device, args)
        if rank == 0:
            print("evaluate_recognition:", rc)
    if args.checkpoint and rank == 0:
        from mvlt_b200.utils import load_checkpoint
        torch.save({"model": model.state_dict(), "optimizer": optimizer.state_dict(), "lr_scheduler": sched.state_dict(),
                    "epoch": args.epochs - 1}, args.checkpoint)
        missing, unexpected, next_epoch = load_checkpoint(model, args.checkpoint, optimizer, sched)
        print(f"checkpoint round trip: missing {len(missing)}, unexpected {len(unexpected)}, next epoch {next_epoch}")
    if rank == 0:
        print(f"done: {total_steps} steps/rank, total_loss {first['total_loss']:.4f} -> {last['total_loss']:.4f}")
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
