"""Drop-in for the reference's ``libs/vl_heads.py`` (/root/reference/libs/vl_heads.py:17-165): same class names,
constructor arguments and parameter names. Inside PVLT the heads are computed by the fused engine
(``mvlt_b200/engine.py``, ``mvlt_b200/t2i.py``); the ``forward`` methods here run the same sm_100a kernels for
stand-alone (inference) use of a head, as the reference's classes allow.
"""
from __future__ import annotations

import torch
from torch import nn

from .. import kernels as k
from .._lib import MvltError

BF16, F32 = torch.bfloat16, torch.float32


def _rows(x):
    return x.reshape(-1, x.shape[-1])


def _lin(x_bf16, lin: nn.Linear, out_dtype=BF16, **epi):
    out = torch.empty((x_bf16.shape[0], lin.out_features), dtype=out_dtype, device=x_bf16.device)
    w = torch.empty(lin.weight.shape, dtype=BF16, device=x_bf16.device)
    k.cast_weight(lin.weight.detach(), w)
    k.gemm(x_bf16, w, out, bias=lin.bias.detach() if lin.bias is not None else None, **epi)
    return out


def _to_bf16(x):
    x = _rows(x).contiguous()
    if x.dtype == BF16:
        return x
    out = torch.empty(x.shape, dtype=BF16, device=x.device)
    k.cast2d(x.to(F32), x.shape[1], out, x.shape[1], x.shape[0], x.shape[1])
    return out


class _GeluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        xc = x.contiguous()
        out = torch.empty_like(xc)
        k.gelu_ew(xc, out)
        ctx.save_for_backward(xc)
        return out

    @staticmethod
    def backward(ctx, g):
        (xc,) = ctx.saved_tensors
        dx = torch.empty_like(xc)
        k.gelu_ew(xc, dx, dy=g.contiguous().to(xc.dtype))
        return dx


class GELU(nn.Module):
    """exact-erf GELU (vl_heads.py:7-14). Inside the heads it is fused into the producing GEMM's epilogue; called on its
    own (as the reference's class can be) it runs the stand-alone sm_100a kernel (fp32 / bf16 CUDA tensors)."""

    def forward(self, x):
        if not x.is_cuda or x.dtype not in (torch.float32, torch.bfloat16):
            raise MvltError("mvlt_b200 GELU needs an fp32 / bf16 CUDA tensor (sm_100a); there is no CPU fallback")
        return _GeluFn.apply(x)


class BertHeadTransform(nn.Module):
    """dense -> GELU -> LayerNorm (vl_heads.py:17-35)."""

    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config['hidden_size'], config['hidden_size'])
        self.transform_act_fn = GELU()
        self.LayerNorm = nn.LayerNorm(config['hidden_size'])

    @torch.no_grad()
    def forward(self, hidden_states):
        x = _to_bf16(hidden_states)
        h = _lin(x, self.dense, act=k.ACT_GELU)
        out = torch.empty_like(h)
        k.layernorm_fwd(h, self.LayerNorm.weight, self.LayerNorm.bias, out, self.LayerNorm.eps, h.shape[0], h.shape[1])
        return out.view(*hidden_states.shape[:-1], -1)


class MLMHead(nn.Module):
    """Masked language modelling head; decoder weight tied to the BERT word embeddings (vl_heads.py:38-70)."""

    def __init__(self, config, bert_model_embedding_weights):
        super().__init__()
        self.transform = BertHeadTransform(config)
        self.hidden_size = config['hidden_size']
        self.vocab_size = config['vocab_size']
        rows, cols = bert_model_embedding_weights.shape
        if (rows, cols) != (self.vocab_size, self.hidden_size):   # same two checks as vl_heads.py:50-55
            raise MvltError(f"tied word-embedding table is {rows}x{cols}, the head was configured for "
                            f"{self.vocab_size}x{self.hidden_size}")
        self.mlm_decoder = nn.Linear(cols, rows, bias=False)
        self.mlm_decoder.weight = bert_model_embedding_weights
        self.bias = nn.Parameter(torch.zeros(self.vocab_size))

    @torch.no_grad()
    def forward(self, input_):
        h = _rows(self.transform(input_))
        V = self.vocab_size
        ld = (V + 7) // 8 * 8
        logits = torch.empty((h.shape[0], ld), dtype=F32, device=h.device)[:, :V]
        w = torch.empty(self.mlm_decoder.weight.shape, dtype=BF16, device=h.device)
        k.cast_weight(self.mlm_decoder.weight.detach(), w)
        k.gemm(h, w, logits, bias=self.bias.detach())
        return logits.reshape(*input_.shape[:-1], V)


class _SmallHead(nn.Module):
    def __init__(self, config, n_out):
        super().__init__()
        self.dim = config['hidden_size']
        self.linear = nn.Linear(self.dim, n_out)
        self.linear_bias = nn.Parameter(torch.zeros(n_out))

    @torch.no_grad()
    def forward(self, input_):
        x = _to_bf16(input_)
        n = self.linear.out_features
        out = torch.empty((x.shape[0], n), dtype=F32, device=x.device)
        k.small_linear_fwd(x, self.linear.weight, self.linear.bias, self.linear_bias, out, x.shape[0], n, self.dim)
        return out.view(*input_.shape[:-1], n)


class ITMHead(_SmallHead):
    """Image-text matching: Linear(768, 2) + extra bias (vl_heads.py:73-87)."""

    def __init__(self, config):
        super().__init__(config, 2)


class CLSHead(_SmallHead):
    """Super-/sub-category classifier (vl_heads.py:90-104)."""

    def __init__(self, config, cls_num):
        super().__init__(config, cls_num)


class ITGHead(nn.Module):
    """Image reconstruction ("t2i" / MVM) head (vl_heads.py:107-165): parameter container with the reference names;
    computed by ``mvlt_b200.t2i.T2IHead``."""

    def __init__(self, embed_dims, channel=64):
        super().__init__()
        self.reduction1 = self.ConvBN(embed_dims[1], channel, 3, padding=1)
        self.reduction2 = self.ConvBN(embed_dims[2], channel, 3, padding=1)
        self.reduction3 = self.ConvBN(embed_dims[3], channel, 3, padding=1)
        self.conv_upsample1 = self.ConvBN(channel, channel, 3, padding=1)
        self.conv_upsample2 = self.ConvBN(channel, channel, 3, padding=1)
        self.conv_upsample3 = self.ConvBN(channel, channel, 3, padding=1)
        self.conv_upsample4 = self.ConvBN(channel, channel, 3, padding=1)
        self.conv_upsample5 = self.ConvBN(2 * channel, 2 * channel, 3, padding=1)
        self.conv_concat2 = self.ConvBN(2 * channel, 2 * channel, 3, padding=1)
        self.conv_concat3 = self.ConvBN(3 * channel, 3 * channel, 3, padding=1)
        self.conv4 = self.ConvBN(3 * channel, 3 * channel, 3, padding=1)
        self.score = nn.Sequential(nn.Conv2d(3 * channel, 3, 1))

    @staticmethod
    def ConvBN(in_planes, out_planes, kernel_size, stride=1, padding=0, dilation=1):
        return nn.Sequential(nn.Conv2d(in_planes, out_planes, kernel_size=kernel_size, stride=stride, padding=padding,
                                       dilation=dilation, bias=False),
                             nn.BatchNorm2d(out_planes))

    @torch.no_grad()
    def forward(self, low_feat, mid_feat, high_feat):
        """Stand-alone inference on NCHW fp32 feature maps (the PVLT model feeds NHWC token buffers instead)."""
        from ..t2i import T2IHead

        class _Ctx:
            pass
        e = _Ctx()
        e.P = {"t2i_head." + n: p for n, p in self.named_parameters()}
        e.Bf = {"t2i_head." + n: b for n, b in self.named_buffers()}
        e.T = 0
        head = T2IHead(e)
        head.prepare_weights()
        feats = []
        for f in (low_feat, mid_feat, high_feat):
            B, C, H, W = f.shape
            feats.append((f.permute(0, 2, 3, 1).contiguous().to(F32).view(B, H * W, C), H, W, C))
        score, _ = head.forward(feats, low_feat.shape[0], self.training)
        B, _, H, W = low_feat.shape
        out = torch.empty((B, 3, H * 8, W * 8), dtype=F32, device=low_feat.device)
        k.upsample8_fwd(score, out, B, H, W, 8)
        return out
