"""Multi-tensor AdamW on the hand-written sm_100a kernel (csrc/optim.cu): one launch per parameter group.

Same update rule and constructor surface as ``torch.optim.AdamW`` (the optimizer timm's ``create_optimizer(opt='adamw')``
builds for the reference, /root/reference/main_vl.py:308): ``param_groups`` with ``lr`` / ``weight_decay`` / ``betas`` /
``eps`` that LR schedulers may rewrite between steps, ``state_dict`` / ``load_state_dict`` in torch's format, and
``zero_grad(set_to_none=True)``. There is no CPU fallback: parameters must live on an sm_100 device.
"""
from __future__ import annotations

import ctypes as C
import struct

import torch

from . import _lib
from ._lib import MvltError, call, ptr

CHUNK = 16384   # fp32 elements per CTA (64 KB of each of p, g, m, v)


def param_groups_no_decay(model, weight_decay):
    """timm ``add_weight_decay`` semantics: no weight decay on 1-D parameters and biases."""
    decay, no_decay = [], []
    for n, p in model.named_parameters():
        if not p.requires_grad:
            continue
        (no_decay if p.ndim <= 1 or n.endswith(".bias") else decay).append(p)
    return [{"params": decay, "weight_decay": weight_decay}, {"params": no_decay, "weight_decay": 0.0}]


def step_bytes(params) -> int:
    """Algorithmic bytes of one AdamW launch over ``params``: p, m, v read + written and g read (28 B per parameter), plus
    2 B per bf16 compute copy the same launch rewrites (``p._mvlt_shadow``: one copy for Linear / embedding / k = s
    convolution weights, two for the 3x3 convolutions, none for parameters the kernels read in fp32)."""
    total = 0
    for p in params:
        sh = getattr(p, "_mvlt_shadow", None)
        copies = 0 if sh is None or not sh[0] else (1 if sh[2] is None else 2)
        total += p.numel() * (28 + 2 * copies)
    return total


class AdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("invalid AdamW hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._tables = {}
        self._hyper = None      # device fp32 [n_groups, 4] = {lr, bias_correction1, bias_correction2, 0}: see enable_device_hyper
        self._hyper_t = None
        self._clip_part = self._clip_out = None     # gradient-clipping scratch: per-chunk squared sums, {coefficient, total norm}
        self.last_grad_norm = None                  # 1-element device tensor: total gradient norm of the last clipped step

    # ---- device-resident hyper-parameters (CUDA-graph replays) -------------------------------------------------------------
    def enable_device_hyper(self, on=True):
        """Make ``step`` read {lr, bias corrections} from device memory instead of passing them by value, so that a captured
        step picks up the current values on every replay. The caller then runs ``advance()`` once ahead of every step (eager
        or replayed): it increments the step count and writes the values of THAT step (one small launch, no host sync)."""
        if not on:
            self._sync_steps()
            self._hyper = self._hyper_t = None
            return
        ps = [p for g in self.param_groups for p in g["params"]]
        dev = ps[0].device
        steps = {int(self.state[p]["step"]) for p in ps if len(self.state.get(p, {}))}
        if len(steps) > 1:
            raise MvltError("parameters must share their step count")
        self._hyper_t = steps.pop() if steps else 0
        self._hyper = torch.zeros((len(self.param_groups), 4), dtype=torch.float32, device=dev)

    def prepare(self):
        """Allocate the moment buffers and build the device pointer tables for the CURRENT ``p.grad`` tensors without taking a
        step (the host -> device table upload cannot happen inside a stream capture)."""
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            for p in ps:
                self._state(p)
            if ps:
                self._group_tables(gi, ps)
        self._clip_buffers()

    def advance(self):
        if self._hyper is None:
            raise MvltError("advance() needs enable_device_hyper()")
        self._hyper_t += 1
        t = self._hyper_t
        payload = b""
        for group in self.param_groups:
            b1, b2 = group["betas"]
            payload += struct.pack("<ffff", group["lr"], 1.0 - b1 ** t, 1.0 - b2 ** t, 0.0)
        from . import kernels as k
        for off in range(0, len(payload), 64):      # 4 groups per launch
            k.set_values(self._hyper.view(-1)[off // 4:], payload[off:off + 64])

    def _sync_steps(self):
        if self._hyper_t is not None:
            for st in self.state.values():
                if "step" in st:
                    st["step"] = self._hyper_t

    def state_dict(self):
        self._sync_steps()
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._tables = {}
        if self._hyper is not None:
            steps = {int(st["step"]) for st in self.state.values() if "step" in st}
            self._hyper_t = steps.pop() if len(steps) == 1 else 0

    def _state(self, p):
        st = self.state[p]
        if len(st) == 0:
            st["step"] = 0
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    def _group_tables(self, gi, ps):
        """Device tables of (p, g, m, v, n) per tensor and of the 64 KB chunks; rebuilt only when a pointer moved."""
        shadows = [getattr(p, "_mvlt_shadow", None) for p in ps]
        key = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]["exp_avg"].data_ptr(),
                     None if sh is None else (sh[0], 0 if sh[1] is None else sh[1].data_ptr(), 0 if sh[2] is None else sh[2].data_ptr()))
                    for p, sh in zip(ps, shadows))
        cached = self._tables.get(gi)
        if cached is not None and cached[0] == key:
            return cached[1], cached[2], cached[3]
        tb, cb, nchunks = bytearray(), bytearray(), 0
        for ti, (p, sh) in enumerate(zip(ps, shadows)):
            st = self.state[p]
            n = p.numel()
            # bf16 compute copies registered by the engine (engine.PVLTEngine._register_shadow): refreshed by the same launch
            mode, w16, w16t, co, ci, kk, ld = sh if sh is not None else (0, None, None, 0, 0, 0, 0)
            if mode and (w16.device != p.device or w16.dtype != torch.bfloat16 or w16.numel() < n):
                raise MvltError("stale bf16 shadow registered on a parameter")
            tb += struct.pack("<QQQQqfiQQiiii", p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(),
                              st["exp_avg_sq"].data_ptr(), n, 1.0, mode, 0 if w16 is None else w16.data_ptr(),
                              0 if w16t is None else w16t.data_ptr(), co, ci, kk, ld)
            for off in range(0, n, CHUNK):
                cb += struct.pack("<qii", off, ti, 0)
                nchunks += 1
        dev = ps[0].device
        tt = torch.frombuffer(tb, dtype=torch.uint8).to(dev)
        ct = torch.frombuffer(cb, dtype=torch.uint8).to(dev)
        self._tables[gi] = (key, tt, ct, nchunks)
        return tt, ct, nchunks

    def _checked(self, group):
        ps = [p for p in group["params"] if p.grad is not None]
        for p in ps:
            if not p.is_cuda or p.dtype != torch.float32 or p.grad.dtype != torch.float32:
                raise MvltError("mvlt_b200.optim.AdamW needs fp32 CUDA parameters and gradients (no CPU fallback)")
            if not p.is_contiguous() or not p.grad.is_contiguous():
                raise MvltError("mvlt_b200.optim.AdamW needs contiguous parameters and gradients")
            self._state(p)
        return ps

    def _clip_buffers(self):
        """Scratch of the fused gradient clipping, sized for the chunk tables built so far (allocated outside stream captures
        when ``prepare`` is used)."""
        total = sum(t[3] for t in self._tables.values())
        if total and (self._clip_part is None or self._clip_part.numel() < total):
            dev = next(iter(self._tables.values()))[1].device
            self._clip_part = torch.empty((total,), dtype=torch.float32, device=dev)
            self._clip_out = torch.empty((2,), dtype=torch.float32, device=dev)
        return total

    def _clip_scale(self, max_norm, grad_scale):
        """torch.nn.utils.clip_grad_norm_ folded into the step (csrc/optim.cu): per-chunk squared sums of every gradient (one
        launch per group, the AdamW chunk tables), then ONE small launch that sums them in a fixed order and writes
        {grad_scale * min(1, max_norm / (norm + 1e-6)), norm} to device memory. The gradients themselves are not rewritten: the
        AdamW kernel multiplies the coefficient in as it reads them. No host synchronisation (``last_grad_norm`` stays on the
        device), deterministic, capturable in a CUDA graph."""
        plan = []
        for gi, group in enumerate(self.param_groups):
            ps = self._checked(group)
            if ps:
                plan.append(self._group_tables(gi, ps))
        if not plan:
            return grad_scale
        total = self._clip_buffers()
        off = 0
        for tt, ct, nchunks in plan:
            call("grad_sumsq_multi", ptr(tt), ptr(ct), C.c_int(nchunks), C.c_int(CHUNK), ptr(self._clip_part[off:]))
            off += nchunks
        call("grad_clip_scale", ptr(self._clip_part), C.c_int(off), C.c_float(float(max_norm)), ptr(grad_scale), ptr(self._clip_out))
        self.last_grad_norm = self._clip_out[1:2]
        return self._clip_out[0:1]

    @torch.no_grad()
    def step(self, closure=None, grad_scale=None, max_norm=None):
        """``grad_scale``: optional 1-element fp32 device tensor multiplied into every gradient (e.g. 1 / loss scale).
        ``max_norm`` > 0: clip the global gradient norm like ``torch.nn.utils.clip_grad_norm_(params, max_norm)`` ahead of the
        update (timm NativeScaler's ``clip_grad``, engine_grid_masking.py:126-127), computed and applied on the device; the
        norm (of the ``grad_scale``-d gradients) is left in ``last_grad_norm``."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if max_norm is not None and max_norm > 0:
            grad_scale = self._clip_scale(max_norm, grad_scale)
        for gi, group in enumerate(self.param_groups):
            ps = self._checked(group)
            if not ps:
                continue
            # torch.optim.AdamW checkpoints (main_vl.py:340) carry ``step`` as a 0-dim tensor: accept both forms
            if self._hyper is not None:
                t = max(self._hyper_t, 1)      # the kernel reads lr / bias corrections of this step from self._hyper[gi]
            else:
                steps = {int(self.state[p]["step"]) for p in ps}
                if len(steps) != 1:
                    raise MvltError("parameters of one group must share their step count")
                t = steps.pop() + 1
            b1, b2 = group["betas"]
            tt, ct, nchunks = self._group_tables(gi, ps)
            if _lib.BYTES is not None:
                _lib.account_bytes("adamw_multi", step_bytes(ps))
            call("adamw_multi", ptr(tt), ptr(ct), C.c_int(nchunks), C.c_int(CHUNK), C.c_float(group["lr"]), C.c_float(b1),
                 C.c_float(b2), C.c_float(group["eps"]), C.c_float(group["weight_decay"]), C.c_float(1.0 - b1 ** t),
                 C.c_float(1.0 - b2 ** t), ptr(grad_scale), C.c_int(0),
                 ptr(self._hyper[gi]) if self._hyper is not None else C.c_void_p(0))
            if self._hyper is None:
                for p in ps:
                    self.state[p]["step"] = t
        # the kernel writes through raw pointers, which autograd's version counters do not see: parameters whose bf16 compute
        # copy this launch could not refresh (none registered yet) and cached derived tables (resized position embeddings) are
        # invalidated through the package-wide epoch instead
        _lib.PARAM_EPOCH += 1
        if any(p.dim() >= 2 and getattr(p, "_mvlt_shadow", None) is None for g in self.param_groups for p in g["params"]):
            _lib.WEIGHT_EPOCH += 1
        return loss
