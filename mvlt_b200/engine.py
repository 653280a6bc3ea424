"""Hand-scheduled forward/backward of the PVLT hot path on top of the C-ABI kernels.

No autograd graph is built inside: ``forward`` records exactly the buffers its ``backward`` needs, and both
enqueue only ``libmvlt_b200.so`` kernels on the current stream (tensors are device-memory handles).
The whole thing is exposed to PyTorch as ONE autograd node (``mvlt_b200/libs/pvlt.py``), so ``nn.Parameter``
ownership, ``state_dict`` names, optimizers and ``DistributedDataParallel`` work unchanged.

Reference path restated (file:line relative to /root/reference):
  libs/pvlt.py:322-356  forward_pyramid_features_vl      -> PVLTEngine._stage_fwd / _block_fwd
  libs/pvlt.py:95-121   Attention.forward (SR attention)  -> _block_fwd (q / sr / kv / softmax / proj)
  libs/pvlt.py:65-71    Mlp.forward                       -> _block_fwd (fc1+GELU epilogue, fc2+residual epilogue)
  libs/pvlt.py:358-401  forward (heads)                   -> heads_fwd
  libs/vl_heads.py      MLMHead / ITMHead / CLSHead / ITGHead
  engine_grid_masking.py:81-102 losses                    -> fused loss path (CE / SmoothL1 kernels)

dtype policy (== the reference under torch.cuda.amp.autocast): fp32 master weights, fp32
residual stream, LayerNorm / softmax / losses in fp32, bf16 GEMM operands with fp32 (TMEM) accumulation.
"""
from __future__ import annotations

import contextlib
import math
import os
from typing import Dict, List, Optional

import torch

from . import _lib
from . import kernels as k
from ._lib import MvltError
from .engine_util import split_k as _split_k

BF16, F32 = torch.bfloat16, torch.float32

EMBED_DIMS = [64, 128, 320, 512]
NUM_HEADS = [1, 2, 5, 8]
MLP_RATIOS = [8, 8, 4, 4]
SR_RATIOS = [8, 4, 2, 1]
PATCH = [4, 2, 2, 2]
DEPTHS = {"pvlt_tiny": [2, 2, 2, 2], "pvlt_small": [3, 4, 6, 3], "pvlt_medium": [3, 4, 18, 3],
          "pvlt_large": [3, 8, 27, 3]}
VOCAB, HIDDEN = 30522, 768
VOCAB_PAD = 30528  # leading dimension of logits buffers (TMA needs 16-byte row strides)
HEAD_DIM = 64
# fused SR-attention forward (csrc/attn_tcgen05.cu); MVLT_FUSED_ATTN=0 keeps the two-GEMM path (QK^T + softmax epilogue, PV)
# for A/B measurements -- both are sm_100a tcgen05 kernels
FUSED_ATTENTION = os.environ.get("MVLT_FUSED_ATTN", "1") != "0"
# fused attention backward (csrc/attn_bwd_tcgen05.cu: dQ, dK, dV in one kernel); MVLT_FUSED_ATTN_BWD=0 keeps the four-GEMM
# path (dV, dP + softmax-backward epilogue, dQ, dK) for A/B measurements
FUSED_ATTENTION_BWD = os.environ.get("MVLT_FUSED_ATTN_BWD", "1") != "0"
# fused MLP (csrc/mlp_tcgen05.cu) for the thin stages (C = 64 / 128); MVLT_FUSED_MLP=0 keeps the two-GEMM path for A/B runs
FUSED_MLP = os.environ.get("MVLT_FUSED_MLP", "1") != "0"
# LayerNorm (norm2) folded into the epilogue of the attention-projection GEMM for the thin stages (C = 64 / 128: the row fits
# one tile): the fp32 rows are normalised on their way out instead of being re-read by a LayerNorm pass; MVLT_FUSED_LN=0 for A/B
FUSED_LN = os.environ.get("MVLT_FUSED_LN", "1") != "0"
FUSED_LN_DIMS = (64, 128)
# spatial-reduction convolution (kernel = stride = R) with its patch matrix read in place through a 5-D TMA view instead of a
# patchify pass (csrc/gemm_desc.h MVLT_CONV_PATCH_A); MVLT_PATCH_VIEW=0 for the A/B
PATCH_VIEW = os.environ.get("MVLT_PATCH_VIEW", "1") != "0"
PATCH_STORE = PATCH_VIEW and os.environ.get("MVLT_PATCH_STORE", "1") != "0"     # the input gradient stored through the same view


def _empty(shape, dtype, dev):
    return torch.empty(shape, dtype=dtype, device=dev)


class PVLTEngine:
    def __init__(self, params: Dict[str, torch.Tensor], buffers: Dict[str, torch.Tensor], depths: List[int],
                 loss_type: Dict[str, int], num_text_tokens: int = 128, drop_path_rate: float = 0.0,
                 embed_dropout: float = 0.1, step_counter: Optional[list] = None):
        self.P = params          # name -> fp32 parameter tensor (live nn.Parameter data)
        self.Bf = buffers        # name -> buffer (BatchNorm running stats)
        self.depths = depths
        self.loss_type = loss_type
        self.T = num_text_tokens
        self.embed_dropout = embed_dropout
        nblk = sum(depths)
        self.dpr = [drop_path_rate * i / max(nblk - 1, 1) for i in range(nblk)]  # torch.linspace(0, r, n)
        self.W: Dict[str, torch.Tensor] = {}   # bf16 compute copies (conv weights permuted to [Co, kh, kw, Ci])
        self._w_version = None
        self._pos_cache = {}
        self.step_counter = step_counter if step_counter is not None else [0]
        self._seed_base = None
        self._dp_rates = None
        self.last_rng = None
        self._grad_layout = None
        # CUDA-graph support (mvlt_b200/graph.py): ``graph_state`` carries the device-resident per-step scalars a captured step
        # reads instead of host values (dropout / drop-path seeds, fixed MLM row capacity and 1 / #labelled rows);
        # ``static_grads`` keeps ONE persistent flat gradient buffer (same addresses every step) instead of a fresh allocation
        self.graph_state = None
        self.static_grads = False
        self._static_flat = self._static_arena = None
        # weight-gradient side stream (set by GraphedStep): the dW / bias-gradient launches feed nothing but the optimizer, so
        # they run as a parallel branch of the captured graph and fill the SMs that the tails and the small launches of the
        # dX chain leave idle. Operands are kept alive (and never reused in place) until the branch is joined.
        self.wgrad_stream = None
        self._wgrad_keep = []
        self._wgrad_pending = False
        # further parallel branches of the captured graph (set by GraphedStep): [0] the spatial-reduction key/value chain of a
        # block next to its query projection, [1] the t2i head next to the MLM / ITM heads (forward and backward)
        self.branch_streams = None
        self._branch_done = []
        from . import t2i as _t2i
        self.t2i = _t2i.T2IHead(self) if loss_type.get("t2i") else None

    # ------------------------------------------------------------------------------------------------
    # weights
    # ------------------------------------------------------------------------------------------------
    def _conv_names(self):
        names = {}
        for i in range(4):
            s = i + 1
            if i > 0:
                names[f"patch_embed{s}.proj.weight"] = PATCH[i] * PATCH[i]
            if SR_RATIOS[i] > 1:
                for j in range(self.depths[i]):
                    names[f"block{s}.{j}.attn.sr.weight"] = SR_RATIOS[i] * SR_RATIOS[i]
        return names

    def prepare_weights(self):
        """fp32 master -> bf16 compute copies, refreshed only when a parameter changed (optimizer step / load)."""
        # invalidation: autograd version counters (torch optimizers, load_state_dict, in-place edits) + the package epoch for
        # writers that go through raw pointers. The own AdamW kernel refreshes the registered copies itself (optim.cu)
        ver = (tuple(p._version for p in self.P.values()), _lib.WEIGHT_EPOCH)
        dev = next(iter(self.P.values())).device
        if self._w_version == ver and self.W:
            return
        convs = self._conv_names()
        no_copy = (0, None, None, 0, 0, 0, 0)   # read in fp32 by the kernels: nothing for the optimizer kernel to refresh
        for name, p in self.P.items():
            if p.dim() < 2 or name.startswith("pos_embed") or name.startswith("text_pos_embed"):
                p._mvlt_shadow = no_copy
                continue
            if name.startswith("t2i_head."):
                if not name.endswith(".0.weight") or "score" in name:
                    p._mvlt_shadow = no_copy
                continue  # handled by the t2i module (3x3 layout)
            if name in ("text_embeddings.position_embeddings.weight", "text_embeddings.token_type_embeddings.weight"):
                p._mvlt_shadow = no_copy
                continue
            if name.endswith("linear.weight") and ("itm_head" in name or "cls_head" in name):
                p._mvlt_shadow = no_copy
                continue  # small heads read fp32 weights directly
            if name not in self.W:
                shape = (p.shape[0], p[0].numel())
                self.W[name] = _empty(shape, BF16, dev)
            if name in convs:
                k.cast_conv_weight(p, self.W[name], p.shape[0], p.shape[1], convs[name], self.W[name].shape[1])
                p._mvlt_shadow = (2, self.W[name], None, p.shape[0], p.shape[1], convs[name], self.W[name].shape[1])
            else:
                k.cast_weight(p, self.W[name])
                p._mvlt_shadow = (1, self.W[name], None, 0, 0, 0, 0)
        if self.t2i is not None:
            self.t2i.prepare_weights()
        self._w_version = ver

    def prepare_static(self, dev):
        """Device tables that depend on the configuration only (built once: no per-step host->device copy, and none inside a
        stream capture)."""
        if self._dp_rates is None or self._dp_rates.device != dev:
            self._dp_rates = torch.tensor([r for r in self.dpr for _ in (0, 1)], device=dev, dtype=F32)

    def _next_seeds(self):
        """Seeds of this step's dropout / drop-path draws (counter-based hash, csrc/common.cuh). The stream is derived from
        torch's seed (``torch.manual_seed``; main_vl.py:205-208 seeds ``args.seed + rank``) and the data-parallel rank, so
        runs are reproducible and ranks draw different masks; the step counter lives on the model (survives engine rebuilds)."""
        if self._seed_base is None:
            rank = 0
            try:
                import torch.distributed as dist
                if dist.is_available() and dist.is_initialized():
                    rank = dist.get_rank()
            except Exception:
                rank = 0
            self._seed_base = (torch.initial_seed() * 0x9E3779B97F4A7C15 + (rank + 1) * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF
        self.step_counter[0] += 1
        z = (self._seed_base + self.step_counter[0] * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z ^= z >> 31
        return z & 0xFFFFFFFFFFFF, (z * 0x94D049BB133111EB + 1) & 0xFFFFFFFFFFFF

    def invalidate(self):
        """Call after writing parameters behind autograd's back (``p.data`` writes, broadcasts, raw pointers)."""
        self._w_version = None
        self._pos_cache = {}

    # ------------------------------------------------------------------------------------------------
    # helpers
    # ------------------------------------------------------------------------------------------------
    def side_launch(self, fn, *keep):
        """Run ``fn()`` (launches that only produce parameter gradients) on the weight-gradient side stream, ordered after
        everything enqueued on the current stream so far; ``keep``: the tensors it reads (held until ``wgrad_join``). Without
        a side stream it simply runs in line."""
        side = self.wgrad_stream
        if side is None:
            fn()
            return
        main = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(main)
        self._wgrad_keep.append(keep)
        with torch.cuda.stream(side):
            side.wait_event(ev)
            fn()
        self._wgrad_pending = True

    def wgrad_join(self):
        """The current stream waits for every side-stream launch so far (before gradients are folded, exchanged or applied)."""
        if self.wgrad_stream is not None and self._wgrad_pending:
            ev = torch.cuda.Event()
            ev.record(self.wgrad_stream)
            torch.cuda.current_stream().wait_event(ev)
            self._wgrad_pending = False
        self._wgrad_keep.clear()

    @contextlib.contextmanager
    def branch(self, idx):
        """``with eng.branch(i):`` -- the launches inside run on branch stream i, ordered after everything enqueued on the
        current stream so far; ``join_branches()`` makes the current stream wait for them. Without branch streams (the
        per-launch path) the body simply runs in line."""
        st = self.branch_streams[idx] if self.branch_streams else None
        if st is None:
            yield
            return
        main = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(main)
        with torch.cuda.stream(st):
            st.wait_event(ev)
            yield
            done = torch.cuda.Event()
            done.record(st)
        self._branch_done.append(done)

    def enable_branches(self, n: int):
        """``n`` > 0: give this engine ``n`` (<= 2) branch streams outside a GraphedStep as well, so that e.g. the key/value
        chain of every block overlaps the query projection in a forward-only sweep; 0 turns them off."""
        dev = next(iter(self.P.values())).device
        self.branch_streams = [torch.cuda.Stream(device=dev) if i < n else None for i in range(2)] if n > 0 else None

    def join_branches(self):
        if self._branch_done:
            main = torch.cuda.current_stream()
            for ev in self._branch_done:
                main.wait_event(ev)
            self._branch_done.clear()

    def wgrad_event(self):
        """An event marking everything enqueued on the weight-gradient stream so far (None without pending side launches): the
        gradient exchange of a finished segment waits for it ON ITS OWN stream -- the compute stream does not stall."""
        if self.wgrad_stream is None or not self._wgrad_pending:
            return None
        ev = torch.cuda.Event()
        ev.record(self.wgrad_stream)
        return ev

    def _lin_param_grads(self, G, wname, bname, dy, x, wgrad=None):
        """dW += dy^T x (split-K, fp32 atomics straight into the gradient buffer); db += column sums of dy (same launch)."""
        rows, co = dy.shape
        ci = x.shape[1]
        tgt = wgrad if wgrad is not None else G[wname].view(co, -1)
        # the bias gradient db = dy^T 1 rides on the same GEMM (one extra N=16 MMA per k-step against a tile of ones)
        self.side_launch(lambda: k.gemm(dy.t(), x.t(), tgt, atomic_add=True, split_k=_split_k(co, ci, rows),
                                        rowsum=G[bname] if bname is not None else None), dy, x)

    def _pos(self, stage, H, W, dev):
        """pvlt.py:291-297,341-344: bilinear resize of the position table (cached per weight version)."""
        key = (stage, H, W, self.P[f"pos_embed{stage}"]._version, _lib.PARAM_EPOCH)
        if key in self._pos_cache:
            return self._pos_cache[key]
        pe = self.P[f"pos_embed{stage}"]
        C = pe.shape[-1]
        tab = pe[0, 1:] if stage == 4 else pe[0]
        side = int(round(math.sqrt(tab.shape[0])))
        n1 = self.P["pos_embed1"].shape[1]
        if H * W == n1:
            out = tab
        else:
            out = _empty((H * W, C), F32, dev)
            k.pos_resize_fwd(tab, out, side, side, H, W, C)
        self._pos_cache = {kk: v for kk, v in self._pos_cache.items() if kk[0] != stage}
        self._pos_cache[key] = out
        return out

    # ------------------------------------------------------------------------------------------------
    # transformer block
    # ------------------------------------------------------------------------------------------------
    def _block_fwd(self, X, pfx, i, B, H, W, dp, save, pre_norm=None, next_pfx=None):
        """``pre_norm``: (xn, mean, rstd) of this block's norm1 when the previous block's MLP already produced it;
        ``next_pfx``: name of the next block of the stage -- its norm1 is then folded into this block's last kernel when that
        kernel can do it (C = 64 / 128), and returned as the third value (None otherwise)."""
        P, Wb, T = self.P, self.W, self.T
        C, R, heads = EMBED_DIMS[i], SR_RATIOS[i], NUM_HEADS[i]
        HW, N = H * W, H * W + T
        M = B * N
        dev = X.device
        hidden = C * MLP_RATIOS[i]
        c = {}
        # ---- attention branch
        if pre_norm is not None:
            xn, mean1, rstd1 = pre_norm
        else:
            xn = _empty((M, C), BF16, dev)
            mean1, rstd1 = _empty((M,), F32, dev), _empty((M,), F32, dev)
            k.layernorm_fwd(X, P[pfx + ".norm1.weight"], P[pfx + ".norm1.bias"], xn, 1e-6, M, C, mean=mean1, rstd=rstd1)
        q = _empty((M, C), BF16, dev)
        Nk = (H // R) * (W // R) + T if R > 1 else N
        if Nk % 32 != 0 or Nk > 256:
            raise MvltError(f"K/V length {Nk} unsupported by the fused softmax epilogue (need a multiple of 32, <= 256)")
        kv = _empty((B * Nk, 2 * C), BF16, dev)
        # the key/value chain (spatial reduction: patchify, conv-as-GEMM, LayerNorm, text rows, then the kv projection) only
        # meets the query projection at the attention kernel: its small launches run as a parallel branch of a captured graph
        with self.branch(0):
            if R > 1:
                oh, ow = H // R, W // R
                sr = _empty((B * oh * ow, C), BF16, dev)
                if PATCH_VIEW and k.conv_patch_supported(H, W, C, R):
                    # the GEMM reads the patches of xn in place (5-D TMA view): no patchify pass on the forward path; the
                    # backward materialises them on the weight-gradient stream, where only the dW GEMM wants them
                    patches = None
                    k.conv_patch_gemm(xn, B, H, W, C, R, N * C, Wb[pfx + ".attn.sr.weight"], sr, bias=P[pfx + ".attn.sr.bias"])
                else:
                    patches = _empty((B * oh * ow, R * R * C), BF16, dev)
                    k.patchify(xn, N * C, patches, B, H, W, C, R)
                    k.gemm(patches, Wb[pfx + ".attn.sr.weight"], sr, bias=P[pfx + ".attn.sr.bias"])
                kvin = _empty((B * Nk, C), BF16, dev)
                srm, srr = _empty((B * oh * ow,), F32, dev), _empty((B * oh * ow,), F32, dev)
                k.layernorm_fwd(sr, P[pfx + ".attn.norm.weight"], P[pfx + ".attn.norm.bias"], kvin, 1e-5, B * oh * ow, C,
                                ymap=(oh * ow, Nk, 0), mean=srm, rstd=srr)
                k.copy_rows(xn, kvin, B * T, C, smap=(T, N, HW), dmap=(T, Nk, oh * ow))
                c.update(patches=patches, sr=sr, srm=srm, srr=srr)
            else:
                kvin = xn
            k.gemm(kvin, Wb[pfx + ".attn.kv.weight"], kv, bias=P[pfx + ".attn.kv.bias"])
        k.gemm(xn, Wb[pfx + ".attn.q.weight"], q, bias=P[pfx + ".attn.q.bias"])
        self.join_branches()
        q4 = q.view(B, N, heads, HEAD_DIM).permute(0, 2, 1, 3)
        kv5 = kv.view(B, Nk, 2, heads, HEAD_DIM)
        k4, v4 = kv5[:, :, 0].permute(0, 2, 1, 3), kv5[:, :, 1].permute(0, 2, 1, 3)
        o = _empty((M, C), BF16, dev)
        if FUSED_ATTENTION and Nk <= k.SR_ATTENTION_MAX_NK:
            # one kernel: S = QK^T in TMEM, softmax in registers, P staged in shared memory as the A operand of PV;
            # the probabilities reach HBM only when the backward needs them
            Pm = _empty((B, heads, N, Nk), BF16, dev) if save else None
            k.sr_attention_fwd(q, kv, o, Pm, B, N, Nk, heads, HEAD_DIM ** -0.5)
        else:
            Pm = _empty((B, heads, N, Nk), BF16, dev)
            k.gemm(q4, k4, Pm, alpha=HEAD_DIM ** -0.5, act=k.ACT_SOFTMAX)   # softmax fused into the QK^T epilogue
            k.gemm(Pm, v4.transpose(-1, -2), o.view(B, N, heads, HEAD_DIM).permute(0, 2, 1, 3))
        X1 = _empty((B, N, C), F32, dev)
        xn2 = _empty((M, C), BF16, dev)
        mean2, rstd2 = _empty((M,), F32, dev), _empty((M,), F32, dev)
        if FUSED_LN and C in FUSED_LN_DIMS:
            # x = x + drop_path(proj(attn)) and norm2(x) in ONE kernel (pvlt.py:141-142): the epilogue normalises the row it just
            # finished (it spans one tile) and also writes the bf16 operand of the MLP
            k.gemm(o, Wb[pfx + ".attn.proj.weight"], X1.view(M, C), bias=P[pfx + ".attn.proj.bias"],
                   residual=X.view(M, C), rowscale=dp[0] if dp else None, rows_per_scale=N,
                   ln=(P[pfx + ".norm2.weight"], P[pfx + ".norm2.bias"], xn2, mean2, rstd2, 1e-6))
        else:
            k.gemm(o, Wb[pfx + ".attn.proj.weight"], X1.view(M, C), bias=P[pfx + ".attn.proj.bias"],
                   residual=X.view(M, C), rowscale=dp[0] if dp else None, rows_per_scale=N)
            # ---- MLP branch
            k.layernorm_fwd(X1, P[pfx + ".norm2.weight"], P[pfx + ".norm2.bias"], xn2, 1e-6, M, C, mean=mean2, rstd=rstd2)
        X2 = _empty((B, N, C), F32, dev)
        fused_mlp = FUSED_MLP and C in k.MLP_FUSED_DIMS and (not save or C in k.MLP_FUSED_BWD_DIMS)
        act = hpre = None
        # the next block's norm1 from this block's last kernel (its rows are complete there): (gamma, beta, xn, mean, rstd, eps)
        nxt = ln_next = None
        if next_pfx is not None and FUSED_LN and C in FUSED_LN_DIMS:
            nxt = (_empty((M, C), BF16, dev), _empty((M,), F32, dev), _empty((M,), F32, dev))
            ln_next = (P[next_pfx + ".norm1.weight"], P[next_pfx + ".norm1.bias"], nxt[0], nxt[1], nxt[2], 1e-6)
        if fused_mlp:
            # one kernel: fc1 -> GELU -> fc2 -> + bias, x drop-path, + residual; the [M, hidden] activation stays in TMEM /
            # shared memory (csrc/mlp_tcgen05.cu). Training saves nothing: the backward recomputes fc1 from xn2.
            k.mlp_fwd(xn2, Wb[pfx + ".mlp.fc1.weight"], P[pfx + ".mlp.fc1.bias"], Wb[pfx + ".mlp.fc2.weight"],
                      P[pfx + ".mlp.fc2.bias"], X1.view(M, C), X2.view(M, C), rowscale=dp[1] if dp else None, rows_per_scale=N,
                      ln=ln_next)
        else:
            act = _empty((M, hidden), BF16, dev)
            # training: the epilogue also stores gelu'(pre-activation) so that the backward GEMM epilogue is a multiply
            hpre = _empty((M, hidden), BF16, dev) if save else None
            k.gemm(xn2, Wb[pfx + ".mlp.fc1.weight"], act, bias=P[pfx + ".mlp.fc1.bias"],
                   act=k.ACT_GELU_SAVE_GRAD if save else k.ACT_GELU, preact_out=hpre)
            k.gemm(act, Wb[pfx + ".mlp.fc2.weight"], X2.view(M, C), bias=P[pfx + ".mlp.fc2.bias"],
                   residual=X1.view(M, C), rowscale=dp[1] if dp else None, rows_per_scale=N, ln=ln_next)
        if save:
            c.update(X=X, xn=xn, mean1=mean1, rstd1=rstd1, q=q, kvin=kvin, kv=kv, Pm=Pm, o=o, X1=X1, xn2=xn2,
                     mean2=mean2, rstd2=rstd2, act=act, hpre=hpre, dp=dp, Nk=Nk)
        return X2, c, nxt

    def _block_bwd(self, dX2, c, pfx, i, B, H, W, G, dy2=None, next_dp=None):
        """``dy2``: bf16 copy of dX2 already scaled by this block's MLP drop-path mask (written by the LayerNorm backward
        that produced dX2), or None. ``next_dp``: MLP drop-path mask of the block that runs next in backward order; when
        given, its scaled bf16 copy of the returned gradient is produced here and returned as the second value."""
        P, Wb, T = self.P, self.W, self.T
        C, R, heads = EMBED_DIMS[i], SR_RATIOS[i], NUM_HEADS[i]
        HW, N = H * W, H * W + T
        M = B * N
        Nk = c["Nk"]
        dev = dX2.device
        hidden = C * MLP_RATIOS[i]
        dp = c["dp"]
        # ---- MLP branch
        if dy2 is None:
            dy2 = _empty((M, C), BF16, dev)
            k.cast_scale_bf16(dX2, dy2, M, C, rowscale=dp[1] if dp else None, rows_per_scale=N)
        dh = _empty((M, hidden), BF16, dev)
        if c["act"] is None:
            # fused forward saved nothing: one kernel recomputes fc1 from xn2 and produces dh = (dy W2) * gelu' together with the
            # fc1 / fc2 weight gradients and the fc1 bias gradient (accumulated in TMEM over the rows); db2 = column sums of dy
            k.mlp_bwd(c["xn2"], dy2, Wb[pfx + ".mlp.fc1.weight"], P[pfx + ".mlp.fc1.bias"], Wb[pfx + ".mlp.fc2.weight"], dh,
                      G[pfx + ".mlp.fc1.weight"], G[pfx + ".mlp.fc2.weight"], G[pfx + ".mlp.fc1.bias"])
            self.side_launch(lambda: k.colsum(dy2, M, C, C, G[pfx + ".mlp.fc2.bias"]), dy2)
        else:
            self._lin_param_grads(G, pfx + ".mlp.fc2.weight", pfx + ".mlp.fc2.bias", dy2, c["act"])
            k.gemm(dy2, Wb[pfx + ".mlp.fc2.weight"].t(), dh, act=k.ACT_MUL_AUX, aux=c["hpre"])
            self._lin_param_grads(G, pfx + ".mlp.fc1.weight", pfx + ".mlp.fc1.bias", dh, c["xn2"])
        # (in place over dy2 -- unless weight-gradient launches that still read dy2 may be pending on the side stream)
        dxn2 = dy2 if self.wgrad_stream is None else _empty((M, C), BF16, dev)
        k.gemm(dh, Wb[pfx + ".mlp.fc1.weight"].t(), dxn2)
        del dh
        dX1 = _empty((B, N, C), F32, dev)
        # ---- attention branch: its bf16, drop-path-scaled input gradient is written by the same LayerNorm pass
        dyp = _empty((M, C), BF16, dev)
        k.layernorm_bwd(dxn2, c["X1"], c["mean2"], c["rstd2"], P[pfx + ".norm2.weight"], dX1, M, C, dx_add=dX2,
                        dgamma=G[pfx + ".norm2.weight"], dbeta=G[pfx + ".norm2.bias"], dx_bf16=dyp,
                        rowscale=dp[0] if dp else None, rows_per_scale=N)
        self._lin_param_grads(G, pfx + ".attn.proj.weight", pfx + ".attn.proj.bias", dyp, c["o"])
        do = _empty((M, C), BF16, dev)
        k.gemm(dyp, Wb[pfx + ".attn.proj.weight"].t(), do)
        do4 = do.view(B, N, heads, HEAD_DIM).permute(0, 2, 1, 3)
        q4 = c["q"].view(B, N, heads, HEAD_DIM).permute(0, 2, 1, 3)
        kv5 = c["kv"].view(B, Nk, 2, heads, HEAD_DIM)
        k4, v4 = kv5[:, :, 0].permute(0, 2, 1, 3), kv5[:, :, 1].permute(0, 2, 1, 3)
        dkv = _empty((B * Nk, 2 * C), BF16, dev)
        dkv5 = dkv.view(B, Nk, 2, heads, HEAD_DIM)
        dk4, dv4 = dkv5[:, :, 0].permute(0, 2, 1, 3), dkv5[:, :, 1].permute(0, 2, 1, 3)
        Pm = c["Pm"]
        dq = dyp if self.wgrad_stream is None else _empty((M, C), BF16, dev)   # dyp is still read by the side-stream dW GEMM
        if FUSED_ATTENTION_BWD and Nk <= k.SR_ATTENTION_MAX_NK:
            k.sr_attention_bwd(c["q"], c["kv"], do, Pm, dq, dkv, B, N, Nk, heads, HEAD_DIM ** -0.5)
        else:
            k.gemm(Pm.transpose(-1, -2), do4.transpose(-1, -2), dv4)          # dV = P^T dO
            dS = _empty((B, heads, N, Nk), BF16, dev)
            # dP = dO V^T with the softmax backward (and the qk scale) fused into the epilogue: writes dS directly
            k.gemm(do4, v4, dS, alpha=HEAD_DIM ** -0.5, act=k.ACT_SOFTMAX_BWD, aux=Pm)
            k.gemm(dS, k4.transpose(-1, -2), dq.view(B, N, heads, HEAD_DIM).permute(0, 2, 1, 3))   # dQ = dS K
            k.gemm(dS.transpose(-1, -2), q4.transpose(-1, -2), dk4)           # dK = dS^T Q
            del dS
        self._lin_param_grads(G, pfx + ".attn.kv.weight", pfx + ".attn.kv.bias", dkv, c["kvin"])
        dxn = _empty((M, C), F32, dev)
        if R > 1:
            oh, ow = H // R, W // R
            dkvin = _empty((B * Nk, C), BF16, dev)
            k.gemm(dkv, Wb[pfx + ".attn.kv.weight"].t(), dkvin)
            k.copy_rows(dkvin, dxn, B * T, C, smap=(T, Nk, oh * ow), dmap=(T, N, HW))
            dsr = _empty((B * oh * ow, C), BF16, dev)
            k.layernorm_bwd(dkvin, c["sr"], c["srm"], c["srr"], P[pfx + ".attn.norm.weight"], dsr, B * oh * ow, C,
                            dymap=(oh * ow, Nk, 0), dgamma=G[pfx + ".attn.norm.weight"],
                            dbeta=G[pfx + ".attn.norm.bias"])
            if c["patches"] is None:      # forward read them in place: materialise for the weight-gradient GEMM, off the dX chain
                xn_s = c["xn"]

                def sr_wgrad(dsr=dsr, xn_s=xn_s):
                    patches = _empty((B * oh * ow, R * R * C), BF16, dev)
                    k.patchify(xn_s, N * C, patches, B, H, W, C, R)
                    k.gemm(dsr.t(), patches.t(), self._conv_wgrad(G, pfx + ".attn.sr.weight"), atomic_add=True,
                           split_k=_split_k(C, R * R * C, B * oh * ow), rowsum=G[pfx + ".attn.sr.bias"])
                self.side_launch(sr_wgrad, dsr, xn_s)
            else:
                self._lin_param_grads(G, None, pfx + ".attn.sr.bias", dsr, c["patches"],
                                      wgrad=self._conv_wgrad(G, pfx + ".attn.sr.weight"))
            if PATCH_STORE and k.conv_patch_supported(H, W, C, R) and oh == 8 and ow == 8:
                # the input-gradient GEMM stores through the patch view: fp32 straight into the image rows of dxn
                k.conv_patch_dgrad(dsr, Wb[pfx + ".attn.sr.weight"], dxn, B, H, W, C, R, N * C)
            else:
                dpatch = _empty((B * oh * ow, R * R * C), BF16, dev)
                k.gemm(dsr, Wb[pfx + ".attn.sr.weight"].t(), dpatch)
                k.unpatchify(dpatch, dxn, N * C, B, H, W, C, R)
            k.gemm(dq, Wb[pfx + ".attn.q.weight"].t(), dxn, residual=dxn)
        else:
            k.gemm(dkv, Wb[pfx + ".attn.kv.weight"].t(), dxn)
            k.gemm(dq, Wb[pfx + ".attn.q.weight"].t(), dxn, residual=dxn)
        self._lin_param_grads(G, pfx + ".attn.q.weight", pfx + ".attn.q.bias", dq, c["xn"])
        dX = _empty((B, N, C), F32, dev)
        dy_next = _empty((M, C), BF16, dev) if next_dp is not None else None      # next_dp False = "no drop-path mask"
        k.layernorm_bwd(dxn, c["X"], c["mean1"], c["rstd1"], P[pfx + ".norm1.weight"], dX, M, C, dx_add=dX1,
                        dgamma=G[pfx + ".norm1.weight"], dbeta=G[pfx + ".norm1.bias"], dx_bf16=dy_next,
                        rowscale=next_dp if torch.is_tensor(next_dp) else None, rows_per_scale=N)
        return dX, dy_next

    def _conv_wgrad(self, G, name):
        """fp32 gradient buffer in the permuted [Co, kh*kw*Ci] layout; folded back into the master layout at the end."""
        return G["__perm__" + name]

    # ------------------------------------------------------------------------------------------------
    # encoder
    # ------------------------------------------------------------------------------------------------
    def encoder_fwd(self, images, ids, training, save):
        P, Wb, T = self.P, self.W, self.T
        dev = images.device
        B, Cin, IH, IW = images.shape
        if ids.shape != (B, T):
            raise MvltError(f"input_ids must be [B, {T}], got {tuple(ids.shape)}")
        ctx = {"B": B, "stages": [], "IH": IH, "IW": IW}
        # --- BERT embeddings (pvlt.py:326)
        p_drop = self.embed_dropout if training else 0.0
        gs = self.graph_state if training else None
        if gs is not None:      # the seeds live in device memory, refreshed by GraphedStep ahead of every run / replay
            seed, dp_seed, seed_dev, dp_seed_dev = 0, 0, gs.seeds[0:1], gs.seeds[1:2]
        else:
            seed, dp_seed = self._next_seeds() if training else (0, 0)
            seed_dev = dp_seed_dev = None
        y768 = _empty((B * T, HIDDEN), BF16, dev)
        em, er = _empty((B * T,), F32, dev), _empty((B * T,), F32, dev)
        ids = ids.contiguous()
        k.bert_embed_fwd(ids, P["text_embeddings.word_embeddings.weight"], P["text_embeddings.position_embeddings.weight"],
                         P["text_embeddings.token_type_embeddings.weight"], P["text_embeddings.LayerNorm.weight"],
                         P["text_embeddings.LayerNorm.bias"], y768, em, er, B * T, T, 1e-12, p_drop, seed, seed_dev=seed_dev)
        ctx.update(ids=ids, em=em, er=er, p_drop=p_drop, seed=seed, seed_dev=seed_dev)
        # --- drop path factors: one [2*nblocks, B] draw per step
        nblk = sum(self.depths)
        dps = None
        if training and any(r > 0 for r in self.dpr):
            self.prepare_static(dev)
            dps = _empty((2 * nblk, B), F32, dev)
            k.keep_scale(dps, 2 * nblk, B, rate_per_row=self._dp_rates, seed=dp_seed, seed_dev=dp_seed_dev)
        self.last_rng = dict(seed=gs.host_seeds[0] if gs is not None else seed, p_drop=p_drop,
                             dp_seed=gs.host_seeds[1] if gs is not None else dp_seed, dps=dps)   # what this step drew (read back by the parity tests)
        Xprev, Hp, Wp = None, IH, IW
        te_in = y768
        blk = 0
        for i in range(4):
            s = i + 1
            C, p = EMBED_DIMS[i], PATCH[i]
            H, W = Hp // p, Wp // p
            HW, N = H * W, H * W + T
            sc = {"H": H, "W": W}
            # patch embed (pvlt.py:165-172): conv k=s=p as patchify + GEMM, LN(1e-5) + pos add written in place
            if i == 0:
                Kp = Cin * p * p
                patches = _empty((B * HW, Kp), BF16, dev)
                k.patchify_nchw(images, patches, B, Cin, IH, IW, p, Kp)
            else:
                Cp = EMBED_DIMS[i - 1]
                Np = Hp * Wp + T
                patches = _empty((B * HW, p * p * Cp), BF16, dev)
                k.patchify(Xprev, Np * Cp, patches, B, Hp, Wp, Cp, p)
            pe = _empty((B * HW, C), BF16, dev)
            k.gemm(patches, Wb[f"patch_embed{s}.proj.weight"], pe, bias=P[f"patch_embed{s}.proj.bias"])
            X = _empty((B, N, C), F32, dev)
            pem, per = _empty((B * HW,), F32, dev), _empty((B * HW,), F32, dev)
            # the first block's norm1 is chained onto the two stage-embedding LayerNorms (they write its input rows): the fp32
            # token buffer is not re-read by a LayerNorm pass of its own
            pre0 = chain0 = None
            if FUSED_LN and self.depths[i] > 0:
                pre0 = (_empty((B * N, C), BF16, dev), _empty((B * N,), F32, dev), _empty((B * N,), F32, dev))
                chain0 = (P[f"block{s}.0.norm1.weight"], P[f"block{s}.0.norm1.bias"], pre0[0], pre0[1], pre0[2], 1e-6)
            k.layernorm_fwd(pe, P[f"patch_embed{s}.norm.weight"], P[f"patch_embed{s}.norm.bias"], X, 1e-5, B * HW, C,
                            ymap=(HW, N, 0), post_add=self._pos(s, H, W, dev), mean=pem, rstd=per, chain=chain0)
            # text embed (pvlt.py:205-208,339)
            if i > 0:
                Cp = EMBED_DIMS[i - 1]
                Np = Hp * Wp + T
                te_in = _empty((B * T, Cp), BF16, dev)
                k.copy_rows(Xprev, te_in, B * T, Cp, smap=(T, Np, Hp * Wp))
            te = _empty((B * T, C), BF16, dev)
            k.gemm(te_in, Wb[f"text_embed{s}.0.weight"], te, bias=P[f"text_embed{s}.0.bias"])
            tem, ter = _empty((B * T,), F32, dev), _empty((B * T,), F32, dev)
            k.layernorm_fwd(te, P[f"text_embed{s}.1.weight"], P[f"text_embed{s}.1.bias"], X, 1e-5, B * T, C,
                            ymap=(T, N, HW), post_add=P[f"text_pos_embed{s}"], mean=tem, rstd=ter, chain=chain0)
            if save:
                sc.update(patches=patches, pe=pe, pem=pem, per=per, te_in=te_in, te=te, tem=tem, ter=ter)
            sc["blocks"] = []
            pre_norm = pre0
            for j in range(self.depths[i]):
                dp = (dps[2 * blk], dps[2 * blk + 1]) if dps is not None else None
                X, bc, pre_norm = self._block_fwd(X, f"block{s}.{j}", i, B, H, W, dp, save, pre_norm=pre_norm,
                                                  next_pfx=f"block{s}.{j + 1}" if j + 1 < self.depths[i] else None)
                sc["blocks"].append(bc)
                blk += 1
            sc["out"] = X
            ctx["stages"].append(sc)
            Xprev, Hp, Wp = X, H, W
        return ctx

    def encoder_bwd(self, ctx, dXs, G, on_segment=None):
        """dXs[i]: fp32 gradient wrt the output token buffer of stage i (or None). Accumulates into G.
        ``on_segment(k)`` is called as soon as every gradient of flat-buffer segment k (1 = stage 4 ... 4 = stage 1 + BERT
        embeddings, see ``grad_segment``) has been enqueued: the data-parallel exchange of that segment can start there."""
        P, Wb, T = self.P, self.W, self.T
        B = ctx["B"]
        dev = ctx["ids"].device
        dX = dXs[3]
        for i in (3, 2, 1, 0):
            s = i + 1
            sc = ctx["stages"][i]
            C, p = EMBED_DIMS[i], PATCH[i]
            H, W = sc["H"], sc["W"]
            HW, N = H * W, H * W + T
            if dX is None:
                dX = torch.zeros((B, N, C), dtype=F32, device=dev)
            dy2 = None
            for j in reversed(range(self.depths[i])):
                # the block that runs next (j-1) needs bf16(dX * its MLP drop-path mask): produced by this block's last pass
                nxt = sc["blocks"][j - 1]["dp"] if j > 0 else None
                next_dp = None if j == 0 else (nxt[1] if nxt else False)
                dX, dy2 = self._block_bwd(dX, sc["blocks"][j], f"block{s}.{j}", i, B, H, W, G, dy2=dy2, next_dp=next_dp)
            # position embeddings (pvlt.py:341-346): batch-sum of the token grads, image part through the
            # transposed bilinear resize
            pe_tab = P[f"pos_embed{s}"]
            side = int(round(math.sqrt(pe_tab.shape[1] - (1 if s == 4 else 0))))
            gtab = G[f"pos_embed{s}"][0, 1:] if s == 4 else G[f"pos_embed{s}"][0]
            def pos_grads(dX=dX, gtab=gtab, side=side, H=H, W=W, C=C, HW=HW, N=N, s=s):
                if HW == P["pos_embed1"].shape[1]:
                    k.batch_reduce(dX, N * C, B, HW * C, gtab, accumulate=True)
                else:
                    dpos = _empty((HW, C), F32, dev)
                    k.batch_reduce(dX, N * C, B, HW * C, dpos)
                    k.pos_resize_bwd(dpos, gtab, side, side, H, W, C)
                k.batch_reduce(dX.view(-1)[HW * C:], N * C, B, T * C, G[f"text_pos_embed{s}"], accumulate=True)
            self.side_launch(pos_grads, dX)      # parameter gradients only: off the dX chain
            # patch embed backward
            dpe = _empty((B * HW, C), BF16, dev)
            k.layernorm_bwd(dX, sc["pe"], sc["pem"], sc["per"], P[f"patch_embed{s}.norm.weight"], dpe, B * HW, C,
                            dymap=(HW, N, 0), dgamma=G[f"patch_embed{s}.norm.weight"],
                            dbeta=G[f"patch_embed{s}.norm.bias"])
            if i == 0:
                self._lin_param_grads(G, f"patch_embed{s}.proj.weight", f"patch_embed{s}.proj.bias", dpe, sc["patches"])
            else:
                self._lin_param_grads(G, None, f"patch_embed{s}.proj.bias", dpe, sc["patches"],
                                      wgrad=self._conv_wgrad(G, f"patch_embed{s}.proj.weight"))
            # text embed backward
            dte = _empty((B * T, C), BF16, dev)
            k.layernorm_bwd(dX, sc["te"], sc["tem"], sc["ter"], P[f"text_embed{s}.1.weight"], dte, B * T, C,
                            dymap=(T, N, HW), dgamma=G[f"text_embed{s}.1.weight"], dbeta=G[f"text_embed{s}.1.bias"])
            self._lin_param_grads(G, f"text_embed{s}.0.weight", f"text_embed{s}.0.bias", dte, sc["te_in"])
            if i > 0:
                Cp = EMBED_DIMS[i - 1]
                Hp, Wp = ctx["stages"][i - 1]["H"], ctx["stages"][i - 1]["W"]
                Np = Hp * Wp + T
                dXp = _empty((B, Np, Cp), F32, dev)
                dpatch = _empty((B * HW, p * p * Cp), BF16, dev)
                k.gemm(dpe, Wb[f"patch_embed{s}.proj.weight"].t(), dpatch)
                k.unpatchify(dpatch, dXp, Np * Cp, B, Hp, Wp, Cp, p)
                k.gemm(dte.view(B, T, C), Wb[f"text_embed{s}.0.weight"].t(), dXp[:, Hp * Wp:, :])
                if dXs[i - 1] is not None:   # t2i head gradients: fp32 [B, HWp, Cp], image rows of the previous stage
                    k.copy_rows(dXs[i - 1], dXp, B * Hp * Wp, Cp, dmap=(Hp * Wp, Np, 0), accumulate=True)
                dX = dXp
                # fold this stage's permuted conv-weight gradients back into the master [Co, Ci, kh, kw] layout (one launch):
                # the stage's segment of the flat gradient buffer is complete after it
                items = [(G[key], G[key[len("__perm__"):]]) for key in G
                         if key.startswith(f"__perm__block{s}.") or key.startswith(f"__perm__patch_embed{s}.")]
                if items:      # (on the weight-gradient stream when there is one: it follows the GEMMs that fill the arena)
                    self.side_launch(lambda: k.uncast_conv_wgrad_multi(items))
                if on_segment is not None:
                    on_segment(4 - i, self.wgrad_event())      # (the exchange also waits for the stage's side-stream weight gradients)
            else:
                dy768 = _empty((B * T, HIDDEN), BF16, dev)
                k.gemm(dte, Wb["text_embed1.0.weight"].t(), dy768)
                k.bert_embed_bwd(dy768, ctx["ids"], P["text_embeddings.word_embeddings.weight"],
                                 P["text_embeddings.position_embeddings.weight"],
                                 P["text_embeddings.token_type_embeddings.weight"], P["text_embeddings.LayerNorm.weight"],
                                 ctx["em"], ctx["er"], G["text_embeddings.word_embeddings.weight"],
                                 G["text_embeddings.position_embeddings.weight"],
                                 G["text_embeddings.token_type_embeddings.weight"], G["text_embeddings.LayerNorm.weight"],
                                 G["text_embeddings.LayerNorm.bias"], B * T, T, ctx["p_drop"], ctx["seed"],
                                 seed_dev=ctx.get("seed_dev"))
        items = [(G[key], G[key[len("__perm__"):]]) for key in G if key.startswith("__perm__block1.")]
        if items:
            self.side_launch(lambda: k.uncast_conv_wgrad_multi(items))
        self.wgrad_join()
        if on_segment is not None:
            on_segment(4, None)

    # ------------------------------------------------------------------------------------------------
    # heads (pvlt.py:365-397)
    # ------------------------------------------------------------------------------------------------
    def _head_embed_fwd(self, feat, name, rows):
        """Linear(512->768) + LayerNorm(1e-5): pvlt.py:244-248,253-257,262-273."""
        P, Wb = self.P, self.W
        dev = feat.device
        he = _empty((rows, HIDDEN), BF16, dev)
        k.gemm(feat, Wb[name + ".0.weight"], he, bias=P[name + ".0.bias"])
        hn = _empty((rows, HIDDEN), BF16, dev)
        m, r = _empty((rows,), F32, dev), _empty((rows,), F32, dev)
        k.layernorm_fwd(he, P[name + ".1.weight"], P[name + ".1.bias"], hn, 1e-5, rows, HIDDEN, mean=m, rstd=r)
        return hn, dict(feat=feat, he=he, m=m, r=r)

    def _head_embed_bwd(self, dhn, c, name, rows, G):
        P, Wb = self.P, self.W
        dhe = _empty((rows, HIDDEN), BF16, dhn.device)
        k.layernorm_bwd(dhn, c["he"], c["m"], c["r"], P[name + ".1.weight"], dhe, rows, HIDDEN,
                        dgamma=G[name + ".1.weight"], dbeta=G[name + ".1.bias"])
        self._lin_param_grads(G, name + ".0.weight", name + ".0.bias", dhe, c["feat"])
        dfeat = _empty((rows, EMBED_DIMS[-1]), BF16, dhn.device)
        k.gemm(dhe, Wb[name + ".0.weight"].t(), dfeat)
        return dfeat

    def small_head_fwd(self, X4, B, HW4, name):
        """ITM / CLS heads on text token 0 (vl_heads.py:84-87,101-104: linear.bias + linear_bias)."""
        P = self.P
        N4 = HW4 + self.T
        dev = X4.device
        feat = _empty((B, EMBED_DIMS[-1]), BF16, dev)
        k.copy_rows(X4, feat, B, EMBED_DIMS[-1], smap=(1, N4, HW4))
        hn, c = self._head_embed_fwd(feat, name + "_head_embed", B)
        n = P[name + "_head.linear.weight"].shape[0]
        logits = _empty((B, n), F32, dev)
        k.small_linear_fwd(hn, P[name + "_head.linear.weight"], P[name + "_head.linear.bias"],
                           P[name + "_head.linear_bias"], logits, B, n, HIDDEN)
        c.update(hn=hn, n=n)
        return logits, c

    def small_head_bwd(self, dlogits, c, name, B, HW4, dX4, G):
        P = self.P
        N4 = HW4 + self.T
        n = c["n"]
        dhn = _empty((B, HIDDEN), BF16, dlogits.device)
        k.small_linear_bwd(dlogits, c["hn"], P[name + "_head.linear.weight"], dhn, G[name + "_head.linear.weight"],
                           G[name + "_head.linear.bias"], G[name + "_head.linear_bias"], B, n, HIDDEN)
        dfeat = self._head_embed_bwd(dhn, c, name + "_head_embed", B, G)
        k.copy_rows(dfeat, dX4, B, EMBED_DIMS[-1], dmap=(1, N4, HW4), accumulate=True)

    def mlm_fwd(self, X4, B, HW4, idx, n_rows, out_f32=False):
        """MLM head (pvlt.py:368-369, vl_heads.py:30-35,65-70) on the rows listed in ``idx`` (None = all B*T)."""
        P, Wb, T = self.P, self.W, self.T
        N4 = HW4 + T
        dev = X4.device
        feat = _empty((n_rows, EMBED_DIMS[-1]), BF16, dev)
        if idx is None:
            k.copy_rows(X4, feat, n_rows, EMBED_DIMS[-1], smap=(T, N4, HW4))
        else:
            k.gather_rows(X4, idx, n_rows, feat, EMBED_DIMS[-1], smap=(T, N4, HW4))
        hn, c = self._head_embed_fwd(feat, "mlm_head_embed", n_rows)
        ha = _empty((n_rows, HIDDEN), BF16, dev)
        hpre = _empty((n_rows, HIDDEN), BF16, dev)
        k.gemm(hn, Wb["mlm_head.transform.dense.weight"], ha, bias=P["mlm_head.transform.dense.bias"],
               act=k.ACT_GELU_SAVE_GRAD, preact_out=hpre)
        hl = _empty((n_rows, HIDDEN), BF16, dev)
        m2, r2 = _empty((n_rows,), F32, dev), _empty((n_rows,), F32, dev)
        k.layernorm_fwd(ha, P["mlm_head.transform.LayerNorm.weight"], P["mlm_head.transform.LayerNorm.bias"], hl,
                        1e-5, n_rows, HIDDEN, mean=m2, rstd=r2)
        logits = _empty((n_rows, VOCAB_PAD), F32 if out_f32 else BF16, dev)[:, :VOCAB]
        k.gemm(hl, Wb["text_embeddings.word_embeddings.weight"], logits, bias=P["mlm_head.bias"])
        c.update(hn=hn, ha=ha, hpre=hpre, hl=hl, m2=m2, r2=r2, idx=idx, n_rows=n_rows)
        return logits, c

    def mlm_bwd(self, dlogits, c, B, HW4, dX4, G):
        """dlogits: bf16 [n_rows, VOCAB] view with leading dimension VOCAB_PAD."""
        P, Wb, T = self.P, self.W, self.T
        N4 = HW4 + T
        n_rows = c["n_rows"]
        dev = dlogits.device
        # tied decoder: dE += dlogits^T hl  (lands in word_embeddings.grad next to the gather's scatter-add)
        def decoder_grads():
            k.gemm(dlogits.t(), c["hl"].t(), G["text_embeddings.word_embeddings.weight"], atomic_add=True,
                   split_k=_split_k(VOCAB, HIDDEN, n_rows))
            k.colsum(dlogits, n_rows, VOCAB, VOCAB_PAD, G["mlm_head.bias"])
        self.side_launch(decoder_grads, dlogits, c["hl"])
        # dH = dlogits E: only ~700 labelled rows but K = 30522 -> split-K into a zeroed fp32 buffer (18 tiles otherwise)
        dhl = k.zeros((n_rows, HIDDEN), F32, dev)
        k.gemm(dlogits, Wb["text_embeddings.word_embeddings.weight"].t(), dhl, atomic_add=True,
               split_k=_split_k(n_rows, HIDDEN, VOCAB))
        dha = _empty((n_rows, HIDDEN), BF16, dev)
        k.layernorm_bwd(dhl, c["ha"], c["m2"], c["r2"], P["mlm_head.transform.LayerNorm.weight"], dha, n_rows, HIDDEN,
                        dgamma=G["mlm_head.transform.LayerNorm.weight"], dbeta=G["mlm_head.transform.LayerNorm.bias"])
        # GELU backward: dpre = dha * gelu'(pre), with gelu' saved by the forward epilogue
        dpre = _empty((n_rows, HIDDEN), BF16, dev)
        k.ew_mul(dha, HIDDEN, 0, dpre, HIDDEN, 0, n_rows, HIDDEN, b=c["hpre"], b_ld=HIDDEN)
        self._lin_param_grads(G, "mlm_head.transform.dense.weight", "mlm_head.transform.dense.bias", dpre, c["hn"])
        dhn = dha
        k.gemm(dpre, Wb["mlm_head.transform.dense.weight"].t(), dhn)
        dfeat = self._head_embed_bwd(dhn, c, "mlm_head_embed", n_rows, G)
        if c["idx"] is None:
            k.copy_rows(dfeat, dX4, n_rows, EMBED_DIMS[-1], dmap=(T, N4, HW4), accumulate=True)
        else:
            k.scatter_rows(dfeat, c["idx"], n_rows, dX4, EMBED_DIMS[-1], dmap=(T, N4, HW4), accumulate=True)

    # ------------------------------------------------------------------------------------------------
    # gradient buffers
    # ------------------------------------------------------------------------------------------------
    @staticmethod
    def grad_segment(name: str) -> int:
        """Order in which the backward pass COMPLETES the gradient of a parameter: 0 = heads, 1..4 = stages 4..1 (stage 1
        together with the BERT embeddings, whose scatter-add is the very last kernel). The flat gradient buffer is laid out
        in this order so that each segment can be exchanged between data-parallel ranks while the rest is still computed."""
        for s in (4, 3, 2, 1):
            if (name.startswith(f"block{s}.") or name.startswith(f"patch_embed{s}.") or name.startswith(f"text_embed{s}.")
                    or name == f"pos_embed{s}" or name == f"text_pos_embed{s}"):
                return 5 - s
        if name.startswith("text_embeddings."):
            return 4
        return 0

    def new_grads(self):
        """One flat zeroed fp32 buffer, viewed per parameter (the tied decoder weight has no separate entry), laid out in
        gradient-completion order (``grad_segment``); ``G["__segments__"]`` = [(begin, end)] element ranges of the segments."""
        dev = next(iter(self.P.values())).device
        total = sum((p.numel() + 3) // 4 * 4 for p in self.P.values())
        if self.static_grads:
            # persistent buffer: the addresses the optimizer's pointer tables and a captured CUDA graph hold stay valid. The
            # caller owns the protocol (one backward per optimizer step, gradients dropped before the next backward)
            if self._static_flat is None or self._static_flat.device != dev or self._static_flat.numel() != total:
                self._static_flat = torch.empty(total, dtype=F32, device=dev)
            flat = self._static_flat
            k.memset_zero(flat)
        else:
            flat = torch.zeros(total, dtype=F32, device=dev)
        if self._grad_layout is None:
            order = sorted(self.P.keys(), key=lambda n: (self.grad_segment(n), n.startswith("text_embeddings.word")))
            layout, bounds, off, cur = [], [], 0, 0
            for name in order:
                seg = self.grad_segment(name)
                while cur < seg:
                    bounds.append(off)
                    cur += 1
                layout.append((name, off))
                off += (self.P[name].numel() + 3) // 4 * 4
            while cur < 5:
                bounds.append(off)
                cur += 1
            self._grad_layout = (layout, [(b, e) for b, e in zip([0] + bounds[:-1], bounds)])
        layout, segments = self._grad_layout
        G = {}
        for name, off in layout:
            p = self.P[name]
            G[name] = flat[off:off + p.numel()].view(p.shape)
        G["__flat__"] = flat
        G["__segments__"] = segments
        # convolution weights accumulate their gradient in the GEMM-friendly permuted layout [Co, kh*kw*Ci]: one zeroed
        # arena for all of them, folded back into the master [Co, Ci, kh, kw] views by one launch per head / encoder
        convs = [(name, p) for name, p in self.P.items() if p.dim() == 4]
        ptotal = sum((p.numel() + 3) // 4 * 4 for _, p in convs)
        if ptotal:
            if self.static_grads:
                if self._static_arena is None or self._static_arena.device != dev or self._static_arena.numel() != ptotal:
                    self._static_arena = torch.empty(ptotal, dtype=F32, device=dev)
                arena = self._static_arena
                k.memset_zero(arena)
            else:
                arena = torch.zeros(ptotal, dtype=F32, device=dev)
            off = 0
            for name, p in convs:
                co, ci, kh, kw = p.shape
                G["__perm__" + name] = arena[off:off + p.numel()].view(co, kh * kw * ci)
                off += (p.numel() + 3) // 4 * 4
        return G
