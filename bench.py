#!/usr/bin/env python
"""Benchmark of the PVLT hot path on B200 (contract: see the task statement / DESIGN.md §Measurement).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

ours:       PVLT-tiny pre-training step (BASELINE.json configs[1]): bf16 operands, batch 128 per GPU, synthetic
            Fashion-Gen-shaped data, grid masking on odd steps, MLM + ITM + t2i(MVM) losses, backward, AdamW step.
            `value` = samples/s with inputs resident in HBM; `e2e` = same metric from pinned HOST buffers (H2D of the
            step's inputs + D2H of the loss inside the timed region). Also reports the candidate-sharded ITM retrieval
            sweep (configs[2]) as `retrieval`, the GEMM-kernel roofline and the CPU baseline (live reference, rank 0).
            Sub-lines of the same run: `sub_benches.recognition` (configs[3], cls-only fine-tune step) and
            `sub_benches.pvlt_small` (configs[4] stand-in), `gpu_eager_reference` (the LIVE reference model under bf16
            autocast + torch.optim.AdamW on the same GPU: the library-kernel bar the hand-written kernels must beat).
reference:  the reference's own CPU implementation of the same step on the host cores: the LIVE unmodified reference
            model staged under baseline/_ref (tools/stage_reference.py; `kind: "reference"`), or -- when that directory is
            absent -- the oracle port (oracle/pvlt_oracle.py, pinned to the reference by golden vectors; `kind: "port"`).
            All host threads, a bounded sample (batch 4) per step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PRE = {"itm": 1, "mlm": 1, "t2i": 1, "cls": 0}
CLS = {"itm": 0, "mlm": 0, "t2i": 0, "cls": 1}   # recognition fine-tune (BASELINE configs[3], dws_mvlt_ft_exp48.py:11)
GF_PER_SAMPLE_TRAIN = 50.35   # dense model FLOPs fwd+bwd, PVLT-tiny, MLM+ITM+t2i
# SURVEY §8d [probe]: (model, heads) -> dense fwd+bwd GFLOP per sample; the default line is ("pvlt_tiny", "pretrain")
GF_TABLE = {("pvlt_tiny", "pretrain"): 50.35, ("pvlt_tiny", "recognition"): 24.97, ("pvlt_small", "pretrain"): 73.98}
WORKLOADS = {
    "pretrain": "pretraining step (grid masking on odd steps + MLM/ITM/t2i losses + backward + AdamW), 256x256 images, "
                "128 BERT tokens",
    "recognition": "recognition fine-tune step (M-CR 48-way + S-CR 122-way classification heads, cls losses + backward + "
                   "AdamW), 256x256 images, 128 BERT tokens",
}
GF_PER_PAIR_RETR = 8.33       # encoder + ITM head forward


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [int(float(r[2])) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for r in self.rows:
            for j, nm in enumerate(names):
                if len(r) > 5 + j and r[5 + j].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def hbm_kernel_table(byte_counts, times_ms, launches, nprof, hbm_gbs):
    """Per-kernel-family view of the memory-bound kernels of the instrumented pass: algorithmic bytes (counted by the
    wrappers in mvlt_b200/kernels.py) / CUDA-event time vs the measured HBM peak."""
    out = {}
    for name, nbytes in sorted(byte_counts.items()):
        ms = times_ms.get(name, 0.0)
        if ms <= 0.0 or nbytes <= 0.0:
            continue
        gbs = nbytes / (ms / 1e3) / 1e9
        out[name] = {"ms_per_step": round(ms / nprof, 4), "launches_per_step": launches.get(name, 0) / nprof,
                     "algorithmic_gb_per_step": round(nbytes / nprof / 1e9, 4), "achieved_gbs": round(gbs, 1),
                     "frac_of_hbm_peak": round(gbs / hbm_gbs, 4)}
    return out


def synth_batch(B, seed, pin=True):
    """SURVEY 8d synthetic Fashion-Gen-shaped batch, on the (pinned) host."""
    from mvlt_b200.synthetic import make_batch
    return make_batch(B, seed=seed, pin=pin and torch.cuda.is_available())


def make_optimizer(model, lr, wd=0.01):
    """timm create_optimizer semantics (main_vl.py:308): AdamW, no weight decay on 1-D / bias parameters; stepped by the
    multi-tensor sm_100a kernel (mvlt_b200/optim.py, csrc/optim.cu)."""
    from mvlt_b200.optim import AdamW, param_groups_no_decay
    return AdamW(param_groups_no_decay(model, wd), lr=lr)


def losses_reference(out, batch, target):
    """The loss block of /root/reference/engine_grid_masking.py:81-102 (weights :23) on a logits dict."""
    F = torch.nn.functional
    total = 0
    if out.get("mlm_logits") is not None:
        total = total + F.cross_entropy(out["mlm_logits"].reshape(-1, 30522).float(), batch["mlm_labels"].view(-1), ignore_index=-1)
    if out.get("itm_logits") is not None:
        total = total + F.cross_entropy(out["itm_logits"].reshape(-1, 2).float(), batch["itm_labels"].view(-1))
    if out.get("sup_cls_logits") is not None:
        total = total + F.cross_entropy(out["sup_cls_logits"].reshape(-1, 48).float(), batch["sup_cls_labels"].view(-1))
        total = total + F.cross_entropy(out["sub_cls_logits"].reshape(-1, 122).float(), batch["sub_cls_labels"].view(-1))
    if out.get("t2i_logits") is not None:
        total = total + 10 * F.smooth_l1_loss(out["t2i_logits"].float(), target)
    return total


def no_decay_groups(model, wd):
    """timm add_weight_decay (main_vl.py:308): no weight decay on 1-D parameters and biases."""
    decay, no_decay = [], []
    for n, p in model.named_parameters():
        (no_decay if p.ndim <= 1 or n.endswith(".bias") else decay).append(p)
    return [{"params": decay, "weight_decay": wd}, {"params": no_decay, "weight_decay": 0.0}]


def train_line(args, model_name, workload, K, Wm, rank, world, local, dev, detail):
    """Times the training step of one (model, workload): `value` (device-resident inputs), `e2e` (pinned host inputs),
    and -- with ``detail`` -- host enqueue cost, clocks and the instrumented per-kernel pass. Returns a dict."""
    import torch.distributed as dist
    import mvlt_b200
    from mvlt_b200 import _lib, masking

    B = args.batch
    hbm, tf_burst, tf_sus, peak_src = _peaks()
    torch.manual_seed(1234)
    heads = dict(PRE if workload == "pretrain" else CLS)
    gf_per_sample = GF_TABLE.get((model_name, workload))
    model = mvlt_b200.create_model(model_name, pretrained=True, num_classes=1000, drop_rate=0.0, drop_path_rate=0.1,
                                   drop_block_rate=None, token_hidden_size=768, num_text_tokens=128, loss_type=heads,
                                   pretrained_pth="").to(dev)
    model.train()
    net = model
    if world > 1:
        if args.ddp:     # the reference's wrapper (main_vl.py:297): works unchanged, pays DDP's bucket copies
            net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], broadcast_buffers=False,
                                                            gradient_as_bucket_view=True)
        else:            # the flat-gradient exchange (mvlt_b200/libs/pvlt.py:enable_grad_sync)
            model.enable_grad_sync(True)
    opt = make_optimizer(model, lr=2.5e-4 * B * world / 512.0)

    host = [synth_batch(B, seed=100 * rank + i) for i in range(2)]
    for h in host:
        h["mlm_count"] = int((h["mlm_labels"] != -1).sum())   # known to the data pipeline; sizes the compacted MLM GEMMs
    devb = [{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in h.items()} for h in host]
    seeds = torch.tensor([masking.sample_seed(rank, i) for i in range(B)], dtype=torch.int64, device=dev)

    # CUDA-graph mode (default; --no-graph keeps the per-launch path): the whole iteration -- forward, losses, backward, the
    # gradient exchange and AdamW -- is captured once per static input buffer set and replayed (mvlt_b200/graph.py). The MLM
    # head then runs on a FIXED row capacity (the labelled count + 10 % headroom, rounded up to the 128-row GEMM tile; padded
    # rows carry the ignore label), and the grid masking of odd steps stays outside the graph (two launches).
    gstep = None
    use_graph = args.graph and not args.ddp
    if use_graph:
        from mvlt_b200.graph import GraphedStep
        cap = None
        if heads["mlm"]:
            cap = -(-int(max(h["mlm_count"] for h in host) * 1.1) // 128) * 128
        gstep = GraphedStep(model, opt, mlm_capacity=cap, warmup=1, enabled=False)
    xm_bufs = {}

    def step(i, b, fwd=None, tag="dev"):
        fwd = fwd or net
        img = b["images"]
        odd = i % 2 == 1 and bool(heads["t2i"])
        if odd:   # engine_grid_masking.py:72-78: odd steps feed the grid-masked image
            grid = masking.grid_mask_batch(seeds + i * B, (img.shape[3], img.shape[2]), 0.5, 16, device=dev)
            if gstep is not None:       # static masked-image buffer per input buffer set
                xm = xm_bufs.get(img.data_ptr())
                if xm is None:
                    xm = xm_bufs[img.data_ptr()] = torch.empty_like(img)
                x = masking.apply_grid_mask(img, grid, 16, out=xm)
            else:
                x = masking.apply_grid_mask(img, grid, 16)
        else:
            x = img
        if heads["cls"]:
            labels = dict(sup_cls_labels=b["sup_cls_labels"], sub_cls_labels=b["sub_cls_labels"])
        else:
            labels = dict(mlm_labels=b["mlm_labels"], itm_labels=b["itm_labels"], target_images=img)
        if gstep is not None:
            total, stats = gstep(x, b["input_ids"], key=(img.data_ptr(), odd), **labels)
            return total
        if heads["mlm"]:
            labels["mlm_count"] = b["mlm_count"]
        total, stats = fwd(x, b["input_ids"], **labels)
        total.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return total

    # ---- end-to-end loop: every step copies its inputs from pinned HOST buffers (H2D) and reads its loss back (D2H).
    # The copies are double-buffered on a side stream so that the H2D of step i+1 overlaps the compute of step i, and
    # the loss of step i is read (pinned buffer + event) while step i+1 is already enqueued: all K input copies and all
    # K loss reads happen inside the timed region.
    E2E_KEYS = ("images", "input_ids", "sup_cls_labels", "sub_cls_labels") if heads["cls"] else \
        ("images", "input_ids", "itm_labels", "mlm_labels")
    copy_stream = torch.cuda.Stream(device=dev)
    stage_bufs = [{k: torch.empty_like(host[0][k], device=dev) for k in E2E_KEYS} for _ in range(2)]
    loss_host = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]

    def run_e2e(n):
        copied = [torch.cuda.Event() for _ in range(2)]
        consumed = [None, None]
        loss_ready = [None, None]
        losses = []
        main = torch.cuda.current_stream()

        def stage(i):
            j = i % 2
            with torch.cuda.stream(copy_stream):
                if consumed[j] is not None:
                    copy_stream.wait_event(consumed[j])       # the step that last read this buffer set has finished
                for k in E2E_KEYS:
                    stage_bufs[j][k].copy_(host[i % 2][k], non_blocking=True)
                copied[j].record(copy_stream)

        stage(0)
        for i in range(n):
            j = i % 2
            if i + 1 < n:
                stage(i + 1)
            main.wait_event(copied[j])
            b = dict(stage_bufs[j])
            b["mlm_count"] = host[i % 2]["mlm_count"]
            loss = step(i, b)
            consumed[j] = torch.cuda.Event()
            consumed[j].record(main)
            loss_host[j].copy_(loss.detach().reshape(1), non_blocking=True)      # D2H read of the step's result
            loss_ready[j] = torch.cuda.Event()
            loss_ready[j].record(main)
            if i > 0:
                loss_ready[1 - j].synchronize()
                losses.append(float(loss_host[1 - j][0]))
        loss_ready[(n - 1) % 2].synchronize()
        losses.append(float(loss_host[(n - 1) % 2][0]))
        return losses

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n, whole_loop=False):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if whole_loop:
            fn(n)
        else:
            for i in range(n):
                fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for i in range(max(Wm, 3)):
        step(i, devb[i % 2])
    if gstep is not None:        # capture (one graph per (buffer set, masked?) pair), then two replays of each
        gstep.enabled = True
        for i in range(6):
            step(i, devb[i % 2])
    sampler = ClockSampler(local)
    if rank == 0 and detail:
        sampler.start()
    l0 = _lib.LAUNCHES
    ms = timed(lambda i: step(i, devb[i % 2]), K)
    launches = _lib.LAUNCHES - l0
    clocks = sampler.stop() if (rank == 0 and detail) else None
    value = B * world * K / (ms / 1e3)

    run_e2e(6 if gstep is not None else 2)      # graph mode: eager warm-up, capture and one replay per staging buffer set
    ms_e2e = timed(run_e2e, K, whole_loop=True)
    e2e_value = B * world * K / (ms_e2e / 1e3)
    h2d = sum(host[0][k].numel() * host[0][k].element_size() for k in E2E_KEYS)
    cfg_idx = 1 if (model_name, workload) == ("pvlt_tiny", "pretrain") else 3 if workload == "recognition" else 4
    res = {
        "value": round(value, 2), "ms_per_step": round(ms / K, 3), "steps": K, "launches": launches,
        "workload": f"{model_name} {WORKLOADS[workload]}, BASELINE configs[{cfg_idx}]",
        "e2e": {"value": round(e2e_value, 2), "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": round(ms_e2e / K, 3)},
        "clocks": clocks,
        "model_tflops": round(value * gf_per_sample / 1e3, 2) if gf_per_sample else None,
        "model_flops_frac_of_bf16_peak": round(value * gf_per_sample / 1e3 / tf_sus, 4) if gf_per_sample else None,
        "cuda_graph": None if gstep is None else {
            "graphs": len(gstep._graphs), "kernels_per_replay": max(g["launches"] for g in gstep._graphs.values()) if gstep._graphs else 0,
            "parallel_branches": "weight-gradient stream + t2i head + key/value chain" if gstep.eng.wgrad_stream is not None else "none",
            "mlm_row_capacity": gstep.state.mlm_cap, "mlm_rows_labelled": [h["mlm_count"] for h in host] if heads["mlm"] else None,
            "mlm_capacity_overflow": gstep.check_overflow()},
    }
    if not detail:
        model.enable_grad_sync(None)
        if gstep is not None:
            gstep.detach()
        return res

    # host-side cost of enqueueing one step (Python + ctypes + allocator), measured from an idle GPU without synchronising:
    # as long as it stays below ms_per_step the device, not the launch path, bounds the step
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step(0, devb[0])
    res["host_enqueue_ms_per_step"] = round((time.perf_counter() - t0) * 1e3, 3)
    torch.cuda.synchronize()

    # ---- per-kernel breakdown (instrumented pass, NOT the reported value) -> roofline of the dominant kernel
    model.enable_grad_sync(None)      # the instrumented pass below runs on rank 0 alone: no collectives from here on
    if gstep is not None:
        gstep.enabled = False         # the instrumented pass times every launch: same kernels, launched one by one ...
        gstep._graphs = {}
        gstep.eng.wgrad_stream = None     # ... and on ONE stream (no parallel branches), so that every kernel is timed alone
        gstep.eng.branch_streams = None
    if rank == 0:
        res.update(instrumented_pass(lambda i: step(i, devb[i % 2], fwd=model), 2, hbm, tf_sus, peak_src))
    return res


def instrumented_pass(run, nprof, hbm, tf_sus, peak_src, spin_ms=25.0):
    """Runs ``run(i)`` nprof times with a CUDA-event pair around every C-ABI launch: per-kernel times, the roofline of the
    dominant kernel (gemm_tcgen05_kernel) and the HBM fractions of the memory-bound kernels."""
    from mvlt_b200 import _lib
    _lib.PROFILE, _lib.GEMM_FLOPS, _lib.GEMM_BYTES, _lib.GEMM_LOG = {}, 0.0, 0.0, []
    _lib.BYTES = {}
    from mvlt_b200 import kernels as _k
    for i in range(nprof):
        # head start for the host: the GPU spins while the step's launches and event records are enqueued, then executes them
        # back to back -- an event interval is then the kernel's duration, not the Python launch path's latency
        torch.cuda.synchronize()
        _k.spin(spin_ms)
        run(i)
    torch.cuda.synchronize()
    prof, flops, nbytes, glog = _lib.PROFILE, _lib.GEMM_FLOPS, _lib.GEMM_BYTES, _lib.GEMM_LOG
    prof.pop("spin", None)
    _lib.PROFILE, _lib.GEMM_LOG = None, None
    byte_counts, _lib.BYTES = (_lib.BYTES or {}), None
    tot = {n: sum(a.elapsed_time(b) for a, b in ev) for n, ev in prof.items()}
    cnt = {n: len(ev) for n, ev in prof.items()}
    try:
        hbm_kernels = hbm_kernel_table(byte_counts, tot, cnt, nprof, hbm)
    except Exception as ex:      # reporting only: never lose the bench line over it
        hbm_kernels = {"error": repr(ex)[:200]}
    allms = sum(tot.values())
    breakdown = {n: {"ms_per_step": round(tot[n] / nprof, 4), "launches_per_step": cnt[n] / nprof,
                     "share": round(tot[n] / allms, 4)} for n in sorted(tot, key=lambda n: -tot[n])[:40]}
    gemm_ms = tot.get("gemm", 0.0)
    gemm_n = cnt.get("gemm", 0)
    # every launch against ITS OWN bound: max(algorithmic bytes / HBM peak, flops / tensor peak)
    gev = prof.get("gemm", [])
    ideal_ms = sum(max(b / (hbm * 1e9), f / (tf_sus * 1e12)) for f, b, _ in glog) * 1e3
    hbm_bound_ms = sum(a.elapsed_time(e) for (a, e), (f, b, _) in zip(gev, glog) if b / (hbm * 1e9) >= f / (tf_sus * 1e12))
    tensor_rows = [(a.elapsed_time(e), f) for (a, e), (f, b, _) in zip(gev, glog) if b / (hbm * 1e9) < f / (tf_sus * 1e12)]
    t_ms, t_fl = sum(r[0] for r in tensor_rows), sum(r[1] for r in tensor_rows)
    hbm_rows = [(a.elapsed_time(e), b) for (a, e), (f, b, _) in zip(gev, glog) if b / (hbm * 1e9) >= f / (tf_sus * 1e12)]
    h_ms, h_by = sum(r[0] for r in hbm_rows), sum(r[1] for r in hbm_rows)
    ach_gbs = nbytes / (gemm_ms / 1e3) / 1e9 if gemm_ms > 0 else 0.0
    ach_tf = flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "gemm_traffic.json")      # written by tools/launch_table.py from the ncu pass
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roof = {"kernel": "gemm_tcgen05_kernel (all GEMM launches of one step; thin-K PVLT-tiny shapes: "
                      f"{100 * hbm_bound_ms / gemm_ms if gemm_ms else 0:.0f}% of its time is in HBM-bound launches)",
            "bound": "hbm", "achieved": round(ach_gbs, 1), "peak": hbm, "unit": "GB/s", "frac": round(ach_gbs / hbm, 4),
            "peak_source": f"{peak_src} hbm_gbs (MEASURED_PEAKS.json copy bandwidth)",
            "algorithmic_bytes_per_launch": round(nbytes / max(gemm_n, 1)), "launches_per_step": gemm_n / nprof,
            "avg_launch_us": round(gemm_ms / max(gemm_n, 1) * 1e3, 2), "traffic": traffic,
            "tensor_view": {"achieved": round(ach_tf, 2), "peak": tf_sus, "unit": "TFLOP/s", "frac": round(ach_tf / tf_sus, 4),
                            "peak_source": f"{peak_src} bf16_tflops_sustained"},
            "tensor_bound_launches": {"ms_per_step": round(t_ms / nprof, 3), "launches_per_step": len(tensor_rows) / nprof,
                                      "achieved_tflops": round(t_fl / (t_ms / 1e3) / 1e12, 1) if t_ms > 0 else None,
                                      "frac_of_sustained_peak": round(t_fl / (t_ms / 1e3) / 1e12 / tf_sus, 4) if t_ms > 0 else None},
            "hbm_bound_launches": {"ms_per_step": round(h_ms / nprof, 3), "launches_per_step": len(hbm_rows) / nprof,
                                   "achieved_gbs": round(h_by / (h_ms / 1e3) / 1e9, 1) if h_ms > 0 else None,
                                   "frac_of_hbm_peak": round(h_by / (h_ms / 1e3) / 1e9 / hbm, 4) if h_ms > 0 else None},
            "frac_of_own_roofline": round(ideal_ms / gemm_ms, 4) if gemm_ms else None,
            "flops_per_step": flops / nprof, "bytes_per_step": nbytes / nprof, "gemm_ms_per_step": round(gemm_ms / nprof, 3),
            "gemm_share_of_step": round(gemm_ms / allms, 4) if allms else None}
    return {"roofline": roof, "kernel_breakdown": breakdown, "hbm_bound_kernels": hbm_kernels}


def gpu_eager_reference(B, dev, steps=5, warmup=2):
    """INFORMATIONAL comparator: the LIVE reference model (baseline/_ref, unmodified libs/pvlt.py) under
    torch.autocast('cuda', bf16) with the loss block of engine_grid_masking.py:81-102 and torch.optim.AdamW, same B, same
    synthetic data, on the same GPU -- i.e. cuBLAS / cuDNN / ATen sm_100 kernels. None of our kernels run here."""
    from baseline import ref_loader
    if not ref_loader.available():
        return {"unavailable": "baseline/_ref not staged (tools/stage_reference.py)"}
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        torch.manual_seed(1234)
        m = ref_loader.build_model("pvlt_tiny", PRE, drop_path_rate=0.1).to(dev).train()
    opt = torch.optim.AdamW(no_decay_groups(m, 0.01), lr=2.5e-4 * B / 512.0)
    hb = [synth_batch(B, seed=i, pin=False) for i in range(2)]
    db = [{k: v.to(dev) for k, v in h.items()} for h in hb]
    masked = [b["images"].masked_fill(torch.rand((B, 1, 16, 16), device=dev).repeat_interleave(16, 2).repeat_interleave(16, 3) < 0.5, 1e-6)
              for b in db]

    def step(i):
        b = db[i % 2]
        x = masked[i % 2] if i % 2 == 1 else b["images"]
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = m(x, b["input_ids"])
            loss = losses_reference(out, b, b["images"])
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()

    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    peak_gb = torch.cuda.max_memory_allocated(dev) / 2 ** 30
    del m, opt, db, masked
    torch.cuda.empty_cache()
    return {"value": round(B / (ms / 1e3), 2), "unit": "samples/s", "ms_per_step": round(ms, 3), "steps": steps, "batch": B,
            "what": "live reference libs/pvlt.py (baseline/_ref) on the same GPU: torch.autocast(bf16) + engine_grid_masking.py:81-102 "
                    "losses + torch.optim.AdamW, device-resident inputs; informational, library sm_100 kernels",
            "peak_mem_gib": round(peak_gb, 1)}


def run_ours(args):
    import torch.distributed as dist
    from mvlt_b200 import _lib

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rc = _lib.load().mvlt_check_device()
    if rc != 0:
        raise SystemExit("mvlt_b200 needs an sm_100 device: " + _lib.load().mvlt_last_error().decode())
    B, K, Wm = args.batch, args.steps, args.warmup
    hbm, tf_burst, tf_sus, peak_src = _peaks()

    main_res = train_line(args, args.model, args.workload, K, Wm, rank, world, local, dev, detail=True)
    torch.cuda.empty_cache()

    # ---- the other training configurations of BASELINE.json as sub-lines of the same run (driver-visible at every N)
    subs = {}
    if args.sub_benches and (args.model, args.workload) == ("pvlt_tiny", "pretrain"):
        for tag, mname, wl in (("recognition", "pvlt_tiny", "recognition"), ("pvlt_small", "pvlt_small", "pretrain")):
            try:
                subs[tag] = train_line(args, mname, wl, max(4, K // 2), 3, rank, world, local, dev, detail=False)
            except Exception as ex:
                subs[tag] = {"error": repr(ex)[:200]}
            torch.cuda.empty_cache()

    # ---- retrieval sweep (configs[2]): 1000 queries x 101 candidates, candidates sharded across ranks
    retr = None
    if args.retrieval_queries > 0:
        try:
            from mvlt_b200 import retrieval
            hook = None
            if rank == 0 and world == 1:      # per-kernel view of one 808-pair ITM-only forward (instrumented, not the timed value)
                hook = lambda run: instrumented_pass(run, 2, hbm, tf_sus, peak_src)   # noqa: E731
            retr = retrieval.bench_sweep(dev, rank, world, n_query=args.retrieval_queries, n_cand=101, warmup=1, profile_hook=hook)
            if rank == 0 and world == 1 and not args.no_cpu:
                retr["cpu_baseline"] = cpu_retrieval_baseline()
        except Exception as ex:   # the training number must still be reported
            retr = {"error": repr(ex)[:200]}
        torch.cuda.empty_cache()

    eager = None
    if rank == 0 and world == 1 and not args.no_eager:
        try:
            eager = gpu_eager_reference(B, dev)
            if "value" in eager:
                eager["ours_over_eager"] = round(main_res["value"] / eager["value"], 3)
        except Exception as ex:
            eager = {"error": repr(ex)[:200]}
    cpu = cpu_baseline(args) if (rank == 0 and world == 1 and not args.no_cpu) else None

    if rank == 0:
        line = {
            "metric": "train_samples_per_s", "value": main_res["value"], "unit": "samples/s", "n_gpus": world, "steps": K,
            "warmup": max(Wm, 3), "ms_per_step": main_res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": main_res["workload"],
                       "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}",
                       "l2": "inputs and activations (>2 GB/step) exceed the 126 MB L2",
                       "optimizer": "mvlt_b200.optim.AdamW (own multi-tensor kernel, one launch per parameter group; refreshes the "
                                    "bf16 weight copies in the same launch)",
                       "mlm_rows": "MLM head evaluated on labelled rows only (identical loss/gradients)",
                       "fused_attention": bool(__import__("mvlt_b200.engine", fromlist=["x"]).FUSED_ATTENTION),
                       "fused_attention_bwd": bool(__import__("mvlt_b200.engine", fromlist=["x"]).FUSED_ATTENTION_BWD),
                       "cuda_graph": main_res.get("cuda_graph")},
            "e2e": main_res["e2e"],
            "gpu_launches": main_res["launches"], "host_enqueue_ms_per_step": main_res.get("host_enqueue_ms_per_step"),
            "clocks": main_res["clocks"],
            "model_tflops": main_res["model_tflops"],
            "model_flops_frac_of_bf16_peak": main_res["model_flops_frac_of_bf16_peak"],
            "roofline": main_res.get("roofline"), "kernel_breakdown": main_res.get("kernel_breakdown"),
            "hbm_bound_kernels": main_res.get("hbm_bound_kernels"), "retrieval": retr, "sub_benches": subs or None,
            "gpu_eager_reference": eager, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _quiet(fn, *a, **kw):
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **kw)


def _cpu_stepper(B, threads):
    """The same workload on the host cores. With baseline/_ref staged: the LIVE, unmodified reference model (libs/pvlt.py,
    train mode, its own dropout / DropPath) + the loss block of engine_grid_masking.py:81-102 + backward + torch.optim.AdamW
    with timm's no-decay grouping (main_vl.py:308) -- ``kind`` "reference". Otherwise the oracle port (oracle/pvlt_oracle.py,
    pinned to the reference by golden vectors) -- ``kind`` "port". Returns (step, kind)."""
    torch.set_num_threads(threads)
    from baseline import ref_loader
    from mvlt_b200.synthetic import make_batch
    if ref_loader.available():
        torch.manual_seed(1234)
        m = _quiet(ref_loader.build_model, "pvlt_tiny", PRE, drop_path_rate=0.1).train()
        opt = torch.optim.AdamW(no_decay_groups(m, 0.01), lr=2.5e-4 * B / 512.0)
        batch = make_batch(B, seed=0)

        def step():
            out = m(batch["images"], batch["input_ids"])
            loss = losses_reference(out, batch, batch["images"])
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
        return step, "reference"
    from oracle import pvlt_oracle as O
    sd = O.make_state_dict("pvlt_tiny", PRE, seed=0)
    batch = O.make_inputs(B, seed=0)
    state = {}

    def step():
        _, grads, _ = O.train_step_grads(sd, batch, PRE)
        if not state:
            state["names"] = list(grads)
            state["params"] = [torch.nn.Parameter(sd[k]) for k in state["names"]]    # shares storage with sd
            state["opt"] = torch.optim.AdamW(state["params"], lr=1e-5, weight_decay=0.05)
        for p, k in zip(state["params"], state["names"]):
            p.grad = grads[k]
        state["opt"].step()
    return step, "port"


CPU_WHAT = {"reference": "LIVE reference libs/pvlt.py (baseline/_ref) fp32 on the host cores",
            "port": "oracle/pvlt_oracle.py fp32 port of the reference on the host cores"}


def cpu_baseline(args):
    threads = os.cpu_count() or 1
    step, kind = _cpu_stepper(4, threads)
    ts = []
    for i in range(4):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    ts = sorted(ts[1:])
    v = 4 / ts[len(ts) // 2]
    return {"value": round(v, 3), "unit": "samples/s", "cores": threads, "kind": kind,
            "sample": f"{CPU_WHAT[kind]}: train step (fwd + MLM/ITM/t2i losses + backward + torch AdamW), batch 4 "
                      "(BASELINE configs[0]; NOT the GPU arm's batch 128), median of 3 steps after 1 warm-up"}


def cpu_retrieval_baseline():
    """one 101-pair query in fp32 on the host cores, as the reference runs it (all pre-training heads computed,
    engine_grid_masking.py:356-358) and ITM-only (what the sharded sweep computes)."""
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    from baseline import ref_loader
    from mvlt_b200.synthetic import make_batch
    b = make_batch(101, seed=99)
    out = {"cores": threads, "unit": "pairs/s", "sample": "one query = 101 (image, text) pairs, batch 101 forward, fp32, eval()"}
    ITM = {"itm": 1, "mlm": 0, "t2i": 0, "cls": 0}
    if ref_loader.available():
        out["kind"] = "reference"
        for tag, lt in (("as_reference_all_heads", PRE), ("itm_only", ITM)):
            m = _quiet(ref_loader.build_model, "pvlt_tiny", lt).eval()
            with torch.no_grad():
                m(b["images"][:8], b["ori_input_ids"][:8])
                t0 = time.perf_counter()
                m(b["images"], b["ori_input_ids"])
                out[tag] = round(101 / (time.perf_counter() - t0), 2)
            del m
    else:
        from oracle import pvlt_oracle as O
        out["kind"] = "port"
        for tag, lt in (("as_reference_all_heads", PRE), ("itm_only", ITM)):
            sd = O.make_state_dict("pvlt_tiny", lt, seed=0)
            with torch.no_grad():
                O.forward(sd, b["images"][:8], b["ori_input_ids"][:8], lt, training=False)
                t0 = time.perf_counter()
                O.forward(sd, b["images"], b["ori_input_ids"], lt, training=False)
                out[tag] = round(101 / (time.perf_counter() - t0), 2)
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    K, Wm = args.steps, max(args.warmup, 1)
    B = 4
    step, kind = _cpu_stepper(B, threads)
    K = min(K, 8)
    for _ in range(min(Wm, 2)):
        step()
    t0 = time.perf_counter()
    for _ in range(K):
        step()
    dt = time.perf_counter() - t0
    v = B * K / dt
    line = {"impl": "reference", "metric": "train_samples_per_s", "value": round(v, 3), "unit": "samples/s",
            "n_gpus": int(os.environ.get("WORLD_SIZE", 1)), "steps": K, "warmup": min(Wm, 2), "ms_per_step": round(dt / K * 1e3, 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"pvlt_tiny {WORKLOADS['pretrain']}, BASELINE configs[1]",
                       "sample": f"{CPU_WHAT[kind]}, bounded sample: batch {B} per step (the GPU arm runs batch 128 per GPU)",
                       "batch_per_step": B},
            "cpu_baseline": {"value": round(v, 3), "unit": "samples/s", "cores": threads, "kind": kind,
                             "sample": f"{K} steps of batch {B}, fwd + MLM/ITM/t2i losses + backward + torch AdamW"},
            "e2e": {"value": round(v, 3), "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--retrieval-queries", type=int, default=1000)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-eager", action="store_true", help="skip the informational gpu_eager_reference leg")
    ap.add_argument("--no-sub", dest="sub_benches", action="store_false", help="skip the recognition / pvlt_small sub-lines")
    ap.add_argument("--model", default="pvlt_tiny", choices=["pvlt_tiny", "pvlt_small", "pvlt_medium", "pvlt_large"],
                    help="default pvlt_tiny = BASELINE configs[1]; pvlt_small = the configs[4] stand-in (SURVEY H9)")
    ap.add_argument("--workload", default="pretrain", choices=["pretrain", "recognition"],
                    help="pretrain = MLM+ITM+t2i heads (configs[1]); recognition = cls-only fine-tune step (configs[3])")
    ap.add_argument("--no-graph", dest="graph", action="store_false",
                    help="launch every kernel of the step individually instead of replaying the captured CUDA graph")
    ap.add_argument("--ddp", action="store_true", help="wrap the model in torch DistributedDataParallel instead of the flat-buffer all-reduce")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
