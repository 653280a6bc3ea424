"""CUDA-graph capture of the pre-training / fine-tuning step (fused forward + losses + backward + own AdamW).

Reference loop restated: /root/reference/engine_grid_masking.py:69-127 (one iteration: forward, the loss block :81-102,
``loss.backward()``, ``optimizer.step()``, ``optimizer.zero_grad()``). The eager path enqueues ~470 kernels through ctypes per
step, which costs ~10 ms of host time against a ~15 ms device step; here the whole iteration is captured ONCE per set of static
input buffers and replayed with one ``cudaGraphLaunch``. Everything that changes from step to step is read from device memory:

* dropout / drop-path seeds        -> ``GraphState.seeds`` (uint64 x 2), written ahead of every replay by ``mvlt_set_values``
* AdamW lr and bias corrections    -> ``AdamW.enable_device_hyper`` / ``advance`` (same mechanism)
* the number of MLM-labelled rows  -> fixed row capacity (``mlm_capacity``); the tail rows are padded with ignore labels, the
                                      count and 1 / count (the CE scale) are produced on the device by ``compact_labels``
* the gradients                    -> ONE persistent flat buffer (``PVLTEngine.static_grads``), zeroed inside the graph

The data-parallel gradient exchange (``model.enable_grad_sync``: NCCL all-reduces on a side stream, forked / joined with events)
is captured with the step. There is no CPU path here either: capture needs the sm_100a library like everything else.
"""
from __future__ import annotations

import os
import struct
from typing import Dict, Optional

import torch

from . import _lib
from . import kernels as k
from ._lib import MvltError


class GraphState:
    """Device-resident per-step scalars of a captured step (owned by the engine while a GraphedStep is attached)."""

    def __init__(self, dev, mlm_cap: int):
        self.seeds = torch.zeros((2,), dtype=torch.int64, device=dev)        # [embedding-dropout seed, drop-path seed]
        self.host_seeds = (0, 0)
        self.mlm_cap = int(mlm_cap)
        self.mlm_inv = torch.ones((1,), dtype=torch.float32, device=dev)     # 1 / max(#labelled rows, 1)
        self.mlm_overflow = torch.zeros((1,), dtype=torch.float32, device=dev)   # > 0: a batch had more labelled rows than mlm_cap


class GraphedStep:
    """``step = GraphedStep(model, optimizer, mlm_capacity=896)`` then ``total, stats = step(images, input_ids, **labels)``
    every iteration, with the SAME tensor objects (static buffers the data loader copies into) for a given ``key``.

    The first ``warmup`` calls of a key run the identical code eagerly (real training steps); the next call captures the graph
    (capture does not execute) and replays it; later calls only replay. ``total`` / ``stats`` are static output tensors that the
    next replay overwrites. ``model`` may have ``enable_grad_sync`` on (the exchange is captured); a DistributedDataParallel
    wrapper is not supported here. ``mlm_capacity`` bounds the labelled rows per batch (a multiple of 128 just above the
    largest expected count is best); ``check_overflow()`` tells, with a device->host read, whether any batch exceeded it.
    ``max_norm``: clip the global gradient norm ahead of the update (``--clip-grad``), on the device, inside the graph.
    """

    def __init__(self, model, optimizer, mlm_capacity: Optional[int] = None, warmup: int = 1, enabled: bool = True,
                 parallel_wgrad: Optional[bool] = None, max_norm: Optional[float] = None):
        from .optim import AdamW
        if not isinstance(optimizer, AdamW):
            raise MvltError("GraphedStep needs mvlt_b200.optim.AdamW (its hyper-parameters must be readable from device memory)")
        self.model, self.opt = model, optimizer
        self.eng = model._engine()
        dev = self.eng._device
        if model.loss_type.get("mlm") and not mlm_capacity:
            raise MvltError("GraphedStep needs mlm_capacity (static MLM row count) for a model with the MLM head")
        self.state = GraphState(dev, mlm_capacity or 0)
        self.warmup = max(int(warmup), 0)
        self.enabled = enabled
        # gradient-norm clipping (main_vl.py --clip-grad) folded into the captured optimizer step: norm and coefficient stay on
        # the device (AdamW.step(max_norm=...)); ``optimizer.last_grad_norm`` holds the norm of the last step
        self.max_norm = float(max_norm) if max_norm else None
        # weight-gradient launches as a parallel branch of the graph (engine.side_launch); MVLT_GRAPH_WGRAD_BRANCH=0 for the A/B
        if parallel_wgrad is None:
            parallel_wgrad = os.environ.get("MVLT_GRAPH_WGRAD_BRANCH", "1") != "0"
        self.parallel_wgrad = bool(parallel_wgrad)
        self._graphs: Dict[object, dict] = {}
        self._calls: Dict[object, int] = {}
        self._pool = None
        self._attach()

    # ---- engine / optimizer wiring ----------------------------------------------------------------------------------------
    def _attach(self):
        eng, opt = self.eng, self.opt
        eng.graph_state = self.state
        eng.static_grads = True
        eng.wgrad_stream = torch.cuda.Stream(device=eng._device) if self.parallel_wgrad else None
        nb = int(os.environ.get("MVLT_GRAPH_BRANCHES", "2")) if self.parallel_wgrad else 0      # A/B: 0 / 1 / 2 extra branches
        eng.branch_streams = [torch.cuda.Stream(device=eng._device) if i < nb else None for i in range(2)] if nb else None
        eng.prepare_weights()
        eng.prepare_static(eng._device)
        opt.enable_device_hyper(True)
        # moment buffers + pointer tables for the persistent gradient views, built outside any capture
        G = eng.new_grads()
        for name, p in eng.P.items():
            if p.requires_grad:
                p.grad = G[name]
        opt.prepare()
        for p in eng.P.values():
            p.grad = None

    def detach(self):
        """Back to the plain eager path (host-side seeds / hyper-parameters, freshly allocated gradients)."""
        self.eng.graph_state = None
        self.eng.static_grads = False
        self.eng.wgrad_stream = None
        self.eng.branch_streams = None
        self.opt.enable_device_hyper(False)
        self._graphs = {}

    def _check_engine(self):
        if self.model._engine() is not self.eng:
            raise MvltError("the model was moved / rebuilt after GraphedStep was created: build a new GraphedStep")

    # ---- one iteration ------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def _iteration(self, images, input_ids, labels):
        """Forward + losses, backward, gradient exchange, AdamW: the engine's hand-scheduled passes called directly (what the
        model's single autograd node does, minus autograd: no AccumulateGrad nodes whose streams a capture would have to match)."""
        from .libs import pvlt as _pvlt
        model, eng = self.model, self.eng
        eng.embed_dropout = model.text_embeddings.dropout.p
        eng.prepare_weights()
        batch = dict(mlm_labels=labels.get("mlm_labels"), itm_labels=labels.get("itm_labels"),
                     sup_cls_labels=labels.get("sup_cls_labels"), sub_cls_labels=labels.get("sub_cls_labels"),
                     target_images=labels.get("target_images"), weights=labels.get("weights") or {}, mlm_count=None,
                     only=labels.get("only"))
        (total, stats), saved = _pvlt.eng_forward_losses(eng, images.contiguous().to(torch.float32), input_ids, batch, True, True)
        G = _pvlt.run_backward(model, "losses", saved, (None,))
        del saved
        for name, p in eng.P.items():
            if p.requires_grad:
                p.grad = G[name]
        self.opt.step(max_norm=self.max_norm)
        self.opt.zero_grad(set_to_none=True)
        return total, stats

    def _write_step_scalars(self):
        seed, dp_seed = self.eng._next_seeds()
        self.state.host_seeds = (seed, dp_seed)
        k.set_values(self.state.seeds, struct.pack("<QQ", seed, dp_seed))
        self.opt.advance()

    def __call__(self, images, input_ids, key=0, **labels):
        self._check_engine()
        if not self.model.training:
            raise MvltError("GraphedStep runs training steps: call model.train() first")
        labels.pop("mlm_count", None)        # static shapes: the count stays on the device
        g = self._graphs.get(key)
        if g is not None:
            ptrs = tuple(t.data_ptr() for t in (images, input_ids, *labels.values()) if torch.is_tensor(t))
            if ptrs != g["ptrs"]:
                raise MvltError(f"GraphedStep key {key!r}: the graph was captured on other buffers (copy the batch into the "
                                "static tensors, or use one key per buffer set)")
        self._write_step_scalars()
        if g is None:
            n = self._calls.get(key, 0)
            self._calls[key] = n + 1
            if not self.enabled or n < self.warmup:
                return self._iteration(images, input_ids, labels)
            g = self._capture(key, images, input_ids, labels)
        g["graph"].replay()
        _lib.LAUNCHES += g["launches"]
        _lib.PARAM_EPOCH += 1
        if self.eng.t2i is not None:
            for pfx, cnt in g["nbt"].items():
                self.eng.t2i.nbt_pending[pfx] = self.eng.t2i.nbt_pending.get(pfx, 0) + cnt
        return g["total"], g["stats"]

    def _capture(self, key, images, input_ids, labels):
        eng = self.eng
        eng._pos_cache = {}                # derived tables must be recomputed INSIDE the graph (the weights change between replays)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        nbt0 = dict(eng.t2i.nbt_pending) if eng.t2i is not None else {}
        l0 = _lib.LAUNCHES
        kw = {} if self._pool is None else {"pool": self._pool}
        with torch.cuda.graph(graph, **kw):
            total, stats = self._iteration(images, input_ids, labels)
        if self._pool is None:
            self._pool = graph.pool()
        launches = _lib.LAUNCHES - l0
        _lib.LAUNCHES = l0                 # nothing ran during capture: replays add the launch count
        nbt = {}
        if eng.t2i is not None:            # the capture pass counted one (not executed) BatchNorm update per unit: undo, replays add it
            for pfx, cnt in eng.t2i.nbt_pending.items():
                d = cnt - nbt0.get(pfx, 0)
                if d:
                    nbt[pfx] = d
            eng.t2i.nbt_pending = nbt0
        eng._pos_cache = {}                # tables computed under capture live in graph memory: never reuse them eagerly
        g = dict(graph=graph, total=total, stats=stats, launches=launches, nbt=nbt,
                 ptrs=tuple(t.data_ptr() for t in (images, input_ids, *labels.values()) if torch.is_tensor(t)))
        self._graphs[key] = g
        return g

    def check_overflow(self) -> bool:
        """True when some batch held more MLM-labelled rows than ``mlm_capacity`` (device->host read: call it rarely)."""
        return bool(float(self.state.mlm_overflow.item()) > 0.0)

    def captured(self, key=0) -> bool:
        return key in self._graphs

    def launches_per_replay(self, key=0) -> int:
        return self._graphs[key]["launches"] if key in self._graphs else 0
