"""MVM / text-to-image head (the reference's ITGHead, /root/reference/libs/vl_heads.py:107-165) scheduled
by hand on the C-ABI kernels: implicit-GEMM 3x3 convolutions (TMA-shifted NHWC boxes feeding the tcgen05 GEMM; no
im2col / col2im buffers) for the eleven conv+BN units, forward, input-gradient and weight-gradient, train-mode BatchNorm,
x2 bilinear upsampling, elementwise products written into channel slices of the concat buffers, the 1x1 score
conv, and the x8 upsample fused with the SmoothL1 loss (engine_grid_masking.py:101).

Activations are NHWC bf16; the three inputs are read in place from the fp32 token buffers of stages 2-4
(image rows are already NHWC, so the reference's permute+contiguous at pvlt.py:352 disappears).
"""
from __future__ import annotations

import torch

from . import kernels as k
from .engine_util import split_k as _split_k

BF16, F32 = torch.bfloat16, torch.float32
CH = 64
UNITS = ["reduction1", "reduction2", "reduction3", "conv_upsample1", "conv_upsample2", "conv_upsample3",
         "conv_upsample4", "conv_upsample5", "conv_concat2", "conv_concat3", "conv4"]


class T2IHead:
    def __init__(self, engine):
        self.e = engine
        self.W = {}
        self.nbt_pending = {}
        self._arena, self._arena_used = None, 0

    def _scratch(self, dev):
        """[2, 192] fp32 zeros for one unit's BatchNorm sums, carved from an arena that is zeroed once per pass."""
        if self._arena is None or self._arena_used >= self._arena.shape[0]:
            self._arena = k.zeros((len(UNITS), 2, 3 * CH), F32, dev)
            self._arena_used = 0
        t = self._arena[self._arena_used]
        self._arena_used += 1
        return t

    def sync_buffers(self):
        """BatchNorm's num_batches_tracked is bookkeeping only (momentum is fixed): counted on the host and written
        to the registered buffers when a state_dict is taken."""
        for pfx, n in self.nbt_pending.items():
            self.e.Bf[pfx + ".1.num_batches_tracked"] += n
        self.nbt_pending = {}

    def prepare_weights(self):
        P = self.e.P
        for u in UNITS:
            w = P[f"t2i_head.{u}.0.weight"]
            name = f"t2i_head.{u}.0.weight"
            if name not in self.W:
                self.W[name] = torch.empty((w.shape[0], 9 * w.shape[1]), dtype=BF16, device=w.device)
            k.cast_conv_weight(w, self.W[name], w.shape[0], w.shape[1], 9, 9 * w.shape[1])
            tname = name + "^T"             # flipped + transposed copy for the input-gradient convolution
            if tname not in self.W:
                self.W[tname] = torch.empty((w.shape[1], 9 * w.shape[0]), dtype=BF16, device=w.device)
            k.cast_conv_weight_t(w, self.W[tname], w.shape[0], w.shape[1], 9)
            w._mvlt_shadow = (2, self.W[name], self.W[tname], w.shape[0], w.shape[1], 9, 9 * w.shape[1])

    # ---- conv3x3 (no bias) + BatchNorm ----------------------------------------------------------------
    def _convbn_fwd(self, u, src, batch_stride, pix_stride, B, H, W, Ci, training, out=None, out_ld=None, out_coff=0):
        P, Bf = self.e.P, self.e.Bf
        pfx = f"t2i_head.{u}"
        Wp = self.W[pfx + ".0.weight"]
        Co = Wp.shape[0]
        rows = B * H * W
        dev = Wp.device
        if src.dtype != BF16:   # encoder token buffers are fp32: one bf16 NHWC copy of the image rows (kept for the weight gradient)
            xb = torch.empty((rows, Ci), dtype=BF16, device=dev)
            k.copy_rows(src, xb, rows, Ci, smap=(H * W, batch_stride // pix_stride, 0), lds=pix_stride)
            src, batch_stride, pix_stride = xb, H * W * Ci, Ci
        y = torch.empty((rows, Co), dtype=BF16, device=dev)
        k.conv3x3_gemm(src, B, H, W, Ci, pix_stride, batch_stride, Wp, y)
        st = self._scratch(dev)
        if training:
            k.bn_stats(y, rows, Co, st[0], st[1])
        aff = torch.empty((4, Co), dtype=F32, device=dev)  # scale, shift, mean, invstd
        k.bn_finalize(st[0], st[1], rows, P[pfx + ".1.weight"], P[pfx + ".1.bias"], Bf[pfx + ".1.running_mean"],
                      Bf[pfx + ".1.running_var"], 0.1, 1e-5, training, aff[0], aff[1], aff[2], aff[3], Co)
        if training:
            self.nbt_pending[pfx] = self.nbt_pending.get(pfx, 0) + 1   # folded into the buffer by sync_buffers()
        if out is None:
            out = torch.empty((rows, Co), dtype=BF16, device=dev)
            out_ld = Co
        k.bn_apply(y, aff[0], aff[1], out, out_ld, out_coff, rows, Co)
        return out, dict(x=src, xs=(batch_stride, pix_stride), y=y, aff=aff, B=B, H=H, W=W, Ci=Ci, Co=Co, training=training)

    def _convbn_bwd(self, u, dout, c, G, dst, dst_batch_stride, dst_pix_stride, accumulate=False):
        """dout: contiguous bf16 [rows, Co]. Writes/accumulates the input gradient into ``dst`` (NHWC, fp32 or bf16)."""
        pfx = f"t2i_head.{u}"
        Wp = self.W[pfx + ".0.weight"]
        B, H, W, Ci, Co = c["B"], c["H"], c["W"], c["Ci"], c["Co"]
        rows = B * H * W
        dev = dout.device
        aff = c["aff"]
        dy = torch.empty((rows, Co), dtype=BF16, device=dev)
        # aff[0] = gamma * invstd is exactly the leading factor of the BatchNorm backward formula. The two column sums the
        # backward needs (sum dy, sum dy*xhat) ARE dbeta and dgamma: they are reduced straight into the (zero-initialised,
        # written once per step) gradient views, no scratch buffers and no copies.
        k.bn_bwd(dout, c["y"], aff[0], aff[2], aff[3], G[pfx + ".1.bias"], G[pfx + ".1.weight"], dy, rows, Co, c["training"])
        key = "__perm__" + pfx + ".0.weight"
        x_in, xs = c["x"], c["xs"]
        self.e.side_launch(lambda: k.conv3x3_wgrad(dy, x_in, B, H, W, Ci, xs[1], xs[0], G[key], split_k=_split_k(Co, 9 * Ci, rows)),
                           dy, x_in)      # weight gradient: parallel branch of the captured graph (engine.side_launch)
        if dst is not None:
            Wt = self.W[pfx + ".0.weight^T"]
            linear = dst_pix_stride == Ci and dst_batch_stride == H * W * Ci
            if linear and dst.dtype == F32:
                k.conv3x3_gemm(dy, B, H, W, Co, Co, H * W * Co, Wt, dst.view(rows, Ci), residual=dst.view(rows, Ci) if accumulate else None)
            elif linear and not accumulate:
                k.conv3x3_gemm(dy, B, H, W, Co, Co, H * W * Co, Wt, dst.view(rows, Ci))
            else:   # bf16 accumulation, or rows scattered in a token buffer: one temporary + a strided (accumulating) copy
                tmp = torch.empty((rows, Ci), dtype=BF16 if dst.dtype == BF16 else F32, device=dev)
                k.conv3x3_gemm(dy, B, H, W, Co, Co, H * W * Co, Wt, tmp)
                k.copy_rows(tmp, dst, rows, Ci, dmap=(H * W, dst_batch_stride // dst_pix_stride, 0), ldd=dst_pix_stride,
                            accumulate=accumulate)

    def _slice(self, t, ld, coff, rows, Cdim):
        out = torch.empty((rows, Cdim), dtype=t.dtype, device=t.device)
        k.cast2d(t.view(-1)[coff:], ld, out, Cdim, rows, Cdim)
        return out

    # ---- forward ------------------------------------------------------------------------------------
    def forward(self, feats, B, training):
        """feats: [(X fp32 [B,N,C], H, W, C)] for stages 2, 3, 4. Returns (score fp32 [B*h*w, 3], ctx)."""
        (Xl, Hl, Wl, Cl), (Xm, Hm, Wm, Cm), (Xh, Hh, Wh, Chh) = feats
        T = self.e.T
        dev = Xl.device
        c = {}
        # BatchNorm sums of this pass: zeroed HERE, ahead of any branch (a unit on a branch stream must not be the one that zeroes)
        self._arena = k.zeros((len(UNITS), 2, 3 * CH), F32, dev)
        self._arena_used = 0
        # reduction1 (the large 32x32 map) is not needed before cat3: as a parallel branch of a captured graph its full-size
        # launches run next to the small 16x16 / 8x8 chains below (engine.branch; in line on the per-launch path)
        with self.e.branch(0):
            low, c["r1"] = self._convbn_fwd("reduction1", Xl, (Hl * Wl + T) * Cl, Cl, B, Hl, Wl, Cl, training)
        mid, c["r2"] = self._convbn_fwd("reduction2", Xm, (Hm * Wm + T) * Cm, Cm, B, Hm, Wm, Cm, training)
        high, c["r3"] = self._convbn_fwd("reduction3", Xh, (Hh * Wh + T) * Chh, Chh, B, Hh, Wh, Chh, training)
        rows_m, rows_l = B * Hm * Wm, B * Hl * Wl
        up_high = torch.empty((rows_m, CH), dtype=BF16, device=dev)
        k.upsample2x_fwd(high, Hh * Wh * CH, CH, up_high, B, Hh, Wh, CH)
        cat2 = torch.empty((rows_m, 2 * CH), dtype=BF16, device=dev)
        A1, c["u1"] = self._convbn_fwd("conv_upsample1", up_high, Hm * Wm * CH, CH, B, Hm, Wm, CH, training)
        k.ew_mul(A1, CH, 0, cat2, 2 * CH, 0, rows_m, CH, b=mid, b_ld=CH)                     # x2_1
        _, c["u4"] = self._convbn_fwd("conv_upsample4", up_high, Hm * Wm * CH, CH, B, Hm, Wm, CH, training,
                                      out=cat2, out_ld=2 * CH, out_coff=CH)
        x2_2, c["c2"] = self._convbn_fwd("conv_concat2", cat2, Hm * Wm * 2 * CH, 2 * CH, B, Hm, Wm, 2 * CH, training)
        up_mid = torch.empty((rows_l, CH), dtype=BF16, device=dev)
        k.upsample2x_fwd(mid, Hm * Wm * CH, CH, up_mid, B, Hm, Wm, CH)
        A2, c["u2"] = self._convbn_fwd("conv_upsample2", up_mid, Hl * Wl * CH, CH, B, Hl, Wl, CH, training)
        up_x21 = torch.empty((rows_l, CH), dtype=BF16, device=dev)
        k.upsample2x_fwd(cat2, Hm * Wm * 2 * CH, 2 * CH, up_x21, B, Hm, Wm, CH)
        A3, c["u3"] = self._convbn_fwd("conv_upsample3", up_x21, Hl * Wl * CH, CH, B, Hl, Wl, CH, training)
        cat3 = torch.empty((rows_l, 3 * CH), dtype=BF16, device=dev)
        self.e.join_branches()
        k.ew_mul(A2, CH, 0, cat3, 3 * CH, 0, rows_l, CH, b=A3, b_ld=CH, c2=low, c2_ld=CH)    # x3_1
        up_x22 = torch.empty((rows_l, 2 * CH), dtype=BF16, device=dev)
        k.upsample2x_fwd(x2_2, Hm * Wm * 2 * CH, 2 * CH, up_x22, B, Hm, Wm, 2 * CH)
        _, c["u5"] = self._convbn_fwd("conv_upsample5", up_x22, Hl * Wl * 2 * CH, 2 * CH, B, Hl, Wl, 2 * CH, training,
                                      out=cat3, out_ld=3 * CH, out_coff=CH)
        x3_2, c["c3"] = self._convbn_fwd("conv_concat3", cat3, Hl * Wl * 3 * CH, 3 * CH, B, Hl, Wl, 3 * CH, training)
        r, c["c4"] = self._convbn_fwd("conv4", x3_2, Hl * Wl * 3 * CH, 3 * CH, B, Hl, Wl, 3 * CH, training)
        P = self.e.P
        score = torch.empty((rows_l, 3), dtype=F32, device=dev)
        k.score_fwd(r, P["t2i_head.score.0.weight"], P["t2i_head.score.0.bias"], score, rows_l, 3 * CH)
        c.update(low=low, mid=mid, A1=A1, A2=A2, A3=A3, cat2=cat2, r=r, dims=(B, Hl, Wl, Hm, Wm, Hh, Wh),
                 chans=(Cl, Cm, Chh))
        return score, c

    # ---- backward -----------------------------------------------------------------------------------
    def backward(self, dscore, c, G, dX4, gscale=None):
        """dscore fp32 [B*Hl*Wl, 3] (times the device scalar ``gscale`` when given). Returns (dfeat2 fp32 [B,Hl*Wl,Cl], dfeat3 fp32 [B,Hm*Wm,Cm]) and accumulates the
        stage-4 image-row gradient into dX4 (fp32 [B, N4, C4])."""
        P = self.e.P
        T = self.e.T
        B, Hl, Wl, Hm, Wm, Hh, Wh = c["dims"]
        Cl, Cm, Chh = c["chans"]
        rows_l, rows_m, rows_h = B * Hl * Wl, B * Hm * Wm, B * Hh * Wh
        dev = dscore.device
        new = lambda rows, ch: torch.empty((rows, ch), dtype=BF16, device=dev)
        dr = new(rows_l, 3 * CH)
        k.score_bwd(dscore, c["r"], P["t2i_head.score.0.weight"], dr, G["t2i_head.score.0.weight"],
                    G["t2i_head.score.0.bias"], rows_l, 3 * CH, gscale=gscale)
        d_x32 = new(rows_l, 3 * CH)
        self._convbn_bwd("conv4", dr, c["c4"], G, d_x32, Hl * Wl * 3 * CH, 3 * CH)
        d_cat3 = new(rows_l, 3 * CH)
        self._convbn_bwd("conv_concat3", d_x32, c["c3"], G, d_cat3, Hl * Wl * 3 * CH, 3 * CH)
        # cat3 = [x3_1 | bn(conv_upsample5(up(x2_2)))]
        d_u5 = self._slice(d_cat3, 3 * CH, CH, rows_l, 2 * CH)
        d_up_x22 = new(rows_l, 2 * CH)
        self._convbn_bwd("conv_upsample5", d_u5, c["u5"], G, d_up_x22, Hl * Wl * 2 * CH, 2 * CH)
        d_x22 = new(rows_m, 2 * CH)
        k.upsample2x_bwd(d_up_x22, d_x22, Hm * Wm * 2 * CH, 2 * CH, B, Hm, Wm, 2 * CH)
        # x3_1 = A2 * A3 * low
        dA2, dA3, g_low = new(rows_l, CH), new(rows_l, CH), new(rows_l, CH)
        k.ew_mul(d_cat3, 3 * CH, 0, dA2, CH, 0, rows_l, CH, b=c["A3"], b_ld=CH, c2=c["low"], c2_ld=CH)
        k.ew_mul(d_cat3, 3 * CH, 0, dA3, CH, 0, rows_l, CH, b=c["A2"], b_ld=CH, c2=c["low"], c2_ld=CH)
        k.ew_mul(d_cat3, 3 * CH, 0, g_low, CH, 0, rows_l, CH, b=c["A2"], b_ld=CH, c2=c["A3"], c2_ld=CH)
        # reduction1's backward (large map; feeds only the stage-2 token gradient) as a parallel branch next to the chains below
        dfeat2 = torch.empty((B, Hl * Wl, Cl), dtype=F32, device=dev)
        with self.e.branch(0):
            self._convbn_bwd("reduction1", g_low, c["r1"], G, dfeat2, Hl * Wl * Cl, Cl)
        d_up_x21 = new(rows_l, CH)
        self._convbn_bwd("conv_upsample3", dA3, c["u3"], G, d_up_x21, Hl * Wl * CH, CH)
        g_x21 = new(rows_m, CH)
        k.upsample2x_bwd(d_up_x21, g_x21, Hm * Wm * CH, CH, B, Hm, Wm, CH)
        d_up_mid = new(rows_l, CH)
        self._convbn_bwd("conv_upsample2", dA2, c["u2"], G, d_up_mid, Hl * Wl * CH, CH)
        g_mid = new(rows_m, CH)
        k.upsample2x_bwd(d_up_mid, g_mid, Hm * Wm * CH, CH, B, Hm, Wm, CH)
        # cat2 = [x2_1 | bn(conv_upsample4(up(high)))]
        d_cat2 = new(rows_m, 2 * CH)
        self._convbn_bwd("conv_concat2", d_x22, c["c2"], G, d_cat2, Hm * Wm * 2 * CH, 2 * CH)
        d_u4 = self._slice(d_cat2, 2 * CH, CH, rows_m, CH)
        d_up_high = new(rows_m, CH)
        self._convbn_bwd("conv_upsample4", d_u4, c["u4"], G, d_up_high, Hm * Wm * CH, CH)
        # x2_1 = A1 * mid ; total gradient of x2_1 = g_x21 + d_cat2[:, :CH]
        k.ew_mul(d_cat2, 2 * CH, 0, g_x21, CH, 0, rows_m, CH, accumulate=True)
        dA1 = new(rows_m, CH)
        k.ew_mul(g_x21, CH, 0, dA1, CH, 0, rows_m, CH, b=c["mid"], b_ld=CH)
        k.ew_mul(g_x21, CH, 0, g_mid, CH, 0, rows_m, CH, b=c["A1"], b_ld=CH, accumulate=True)
        self._convbn_bwd("conv_upsample1", dA1, c["u1"], G, d_up_high, Hm * Wm * CH, CH, accumulate=True)
        g_high = new(rows_h, CH)
        k.upsample2x_bwd(d_up_high, g_high, Hh * Wh * CH, CH, B, Hh, Wh, CH)
        # reductions back into the encoder's token-gradient buffers
        N4 = Hh * Wh + T
        self._convbn_bwd("reduction3", g_high, c["r3"], G, dX4, N4 * Chh, Chh, accumulate=True)
        dfeat3 = torch.empty((B, Hm * Wm, Cm), dtype=F32, device=dev)
        self._convbn_bwd("reduction2", g_mid, c["r2"], G, dfeat3, Hm * Wm * Cm, Cm)
        self.e.join_branches()
        # fold the permuted 3x3 weight gradients back to [Co, Ci, 3, 3] (one launch for the eleven units)
        self.e.side_launch(lambda: k.uncast_conv_wgrad_multi([(G["__perm__t2i_head.%s.0.weight" % u], G["t2i_head.%s.0.weight" % u])
                                                              for u in UNITS]))
        return dfeat2, dfeat3
