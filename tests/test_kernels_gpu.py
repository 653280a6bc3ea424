"""Kernel-level parity through the C-ABI. Integer / index kernels are checked bit-exactly against the oracle and the
reference's golden vectors; floating-point kernels against fp32 torch references with the tolerance in each test."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
BF16, F32 = torch.bfloat16, torch.float32


def _g(seed):
    return torch.Generator(device="cuda").manual_seed(seed)


def _close(a, b, rtol, atol, tag=""):
    a, b = a.float(), b.float()
    err = (a - b).abs()
    lim = atol + rtol * b.abs()
    assert bool((err <= lim).all()), f"{tag}: max err {err.max().item():.4g} (ref max {b.abs().max().item():.4g})"


# ------------------------------------------------------------------------------------------------------------
# integer kernels: bit exact
# ------------------------------------------------------------------------------------------------------------
def test_grid_mask_bit_exact_vs_reference_golden_and_oracle():
    from mvlt_b200 import masking
    from oracle import grid_mask as ogm
    g = np.load(os.path.join(GOLD, "grid_mask_golden.npz"))
    seeds = [int(s) for s in g["seeds"]]
    got = masking.grid_mask_batch(seeds).cpu().numpy()
    assert (got == g["grids"]).all()                               # reference's own generate_grid_mask outputs
    got2 = masking.grid_mask_batch(list(range(8)), (352, 352), 0.75, 16).cpu().numpy()
    assert (got2 == g["extra_352_075"]).all()
    big = [masking.sample_seed(3, i) for i in range(512)]          # full-batch sizes: vs the C oracle
    gb = masking.grid_mask_batch(big).cpu().numpy()
    for s, grid in zip(big, gb):
        assert (grid == ogm.grid_c(s)).all()
    m = masking.generate_grid_mask((256, 256), 0.5, 16, seed=7)    # reference signature / return type
    assert m.shape == (1, 256, 256) and m.dtype == np.float64 and (m == ogm.expand(ogm.grid_c(7))).all()


def test_masked_fill_exact():
    from mvlt_b200 import masking
    from oracle import grid_mask as ogm
    img = torch.rand((5, 3, 256, 256), generator=_g(0), device="cuda")
    seeds = [11, 12, 13, 14, 15]
    grid = masking.grid_mask_batch(seeds)
    out, mask = masking.apply_grid_mask(img, grid, return_mask=True)
    for i, s in enumerate(seeds):
        m = ogm.expand(ogm.grid_py(s))
        ref = ogm.masked_fill(img[i].cpu().numpy(), m)
        assert (out[i].cpu().numpy() == ref).all()
        assert (mask[i].cpu().numpy() == m.astype(np.float32)).all()
    # idempotence: masking an already masked image with the same grid changes nothing
    assert torch.equal(masking.apply_grid_mask(out, grid), out)


def test_compact_gather_scatter_exact():
    from mvlt_b200 import kernels as k
    n, C = 128 * 128, 512
    labels = torch.full((n,), -1, dtype=torch.int64, device="cuda")
    pos = torch.randperm(n, generator=_g(1), device="cuda")[:777]
    labels[pos] = torch.randint(0, 30522, (777,), generator=_g(2), device="cuda")
    idx = torch.empty((n,), dtype=torch.int32, device="cuda")
    lab = torch.empty((n,), dtype=torch.int64, device="cuda")
    cnt = torch.zeros((1,), dtype=torch.int32, device="cuda")
    k.compact_labels(labels, n, -1, idx, lab, cnt)
    ref_idx = (labels != -1).nonzero().view(-1)
    assert int(cnt.item()) == 777
    assert torch.equal(idx[:777].long(), ref_idx) and torch.equal(lab[:777], labels[ref_idx])
    # empty and full edge cases
    k.compact_labels(torch.full((300,), -1, dtype=torch.int64, device="cuda"), 300, -1, idx, lab, cnt)
    assert int(cnt.item()) == 0
    k.compact_labels(torch.zeros((300,), dtype=torch.int64, device="cuda"), 300, -1, idx, lab, cnt)
    assert int(cnt.item()) == 300 and torch.equal(idx[:300].long(), torch.arange(300, device="cuda"))
    # gather (fp32 -> fp32 is a bit-exact copy) and scatter round trip with a token row map
    T, HW, N = 128, 64, 192
    X = torch.randn((128, N, C), generator=_g(3), device="cuda")
    k.compact_labels(labels, n, -1, idx, lab, cnt)
    out = torch.empty((777, C), dtype=F32, device="cuda")
    k.gather_rows(X, idx, 777, out, C, smap=(T, N, HW))
    ref = X[:, HW:, :].reshape(-1, C)[ref_idx]
    assert torch.equal(out, ref)
    back = torch.zeros_like(X)
    k.scatter_rows(out, idx, 777, back, C, dmap=(T, N, HW))
    exp = torch.zeros_like(X)
    expv = torch.zeros((128 * T, C), device="cuda")
    expv[ref_idx] = ref
    exp[:, HW:, :] = expv.view(128, T, C)
    assert torch.equal(back, exp)


def test_itm_rank_identical_to_oracle():
    from mvlt_b200 import kernels as k
    from oracle import pvlt_oracle as O
    Q, n_cand = 1000, 101
    logits = torch.randn((Q, n_cand, 2), generator=_g(4), device="cuda") * 0.05
    ranks = torch.empty((Q,), dtype=torch.int32, device="cuda")
    probs = torch.empty((Q, n_cand), dtype=F32, device="cuda")
    k.itm_rank(logits.contiguous(), Q, n_cand, ranks, probs)
    lc = logits.cpu()
    ref = torch.tensor([O.retrieval_rank(lc[i]) for i in range(Q)], dtype=torch.int32)
    assert torch.equal(ranks.cpu(), ref)
    _close(probs.cpu(), F.softmax(lc, -1)[..., 1], 1e-6, 1e-6, "p(match)")
    # permutation property: moving the positive elsewhere and back keeps rank; a candidate beating everyone -> rank 0
    logits[:, 0, 1] += 10.0
    k.itm_rank(logits.contiguous(), Q, n_cand, ranks)
    assert int(ranks.max().item()) == 0


# ------------------------------------------------------------------------------------------------------------
# floating point kernels
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("C", [64, 128, 320, 512, 768])
@pytest.mark.parametrize("in_dtype,out_dtype", [(F32, BF16), (BF16, F32), (F32, F32)])
def test_layernorm_fwd_bwd(C, in_dtype, out_dtype):
    from mvlt_b200 import kernels as k
    rows = 1000 + C // 64
    x = torch.randn((rows, C), generator=_g(C), device="cuda").to(in_dtype)
    gamma = 1 + 0.1 * torch.randn(C, generator=_g(1), device="cuda")
    beta = 0.1 * torch.randn(C, generator=_g(2), device="cuda")
    y = torch.empty((rows, C), dtype=out_dtype, device="cuda")
    mean, rstd = torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
    k.layernorm_fwd(x, gamma, beta, y, 1e-6, rows, C, mean=mean, rstd=rstd)
    xr = x.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ref = F.layer_norm(xr, (C,), gr, br, 1e-6)
    tol = 1e-2 if out_dtype == BF16 else 2e-5
    _close(y, ref, tol, tol, "ln fwd")
    dy = torch.randn((rows, C), generator=_g(3), device="cuda")
    ref.backward(dy)
    dx = torch.empty((rows, C), dtype=F32, device="cuda")
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    add = torch.randn((rows, C), generator=_g(5), device="cuda")
    dx16 = torch.empty((rows, C), dtype=BF16, device="cuda")
    rs = torch.rand((rows + 49) // 50, generator=_g(6), device="cuda")
    k.layernorm_bwd(dy, x, mean, rstd, gamma, dx, rows, C, dx_add=add, dgamma=dg, dbeta=db, dx_bf16=dx16, rowscale=rs,
                    rows_per_scale=50)
    _close(dx, xr.grad + add, 1e-4, 1e-4, "ln dx")
    _close(dx16, (xr.grad + add) * rs.repeat_interleave(50)[:rows, None], 1e-2, 1e-2, "ln dx (bf16, drop-path scaled copy)")
    _close(dg, gr.grad, 1e-3, 1e-3, "ln dgamma")
    _close(db, br.grad, 1e-3, 1e-3, "ln dbeta")


def test_layernorm_row_maps_and_post_add():
    from mvlt_b200 import kernels as k
    B, HW, T, C = 3, 64, 128, 128
    N = HW + T
    pe = torch.randn((B * HW, C), generator=_g(1), device="cuda").to(BF16)
    pos = torch.randn((HW, C), generator=_g(2), device="cuda")
    gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    X = torch.zeros((B, N, C), device="cuda")
    k.layernorm_fwd(pe, gamma, beta, X, 1e-5, B * HW, C, ymap=(HW, N, 0), post_add=pos)
    ref = F.layer_norm(pe.float(), (C,)).view(B, HW, C) + pos
    _close(X[:, :HW], ref, 1e-5, 1e-5, "rowmap+pos")
    assert float(X[:, HW:].abs().max()) == 0.0
    # chained second LayerNorm of the rows just written (the first block's norm1), for every width the encoder uses
    for C2 in (64, 128, 320, 512):
        pe2 = torch.randn((B * HW, C2), generator=_g(C2), device="cuda").to(BF16)
        pos2 = torch.randn((HW, C2), generator=_g(C2 + 1), device="cuda") * 0.5 + 0.3
        g1, b1 = torch.randn(C2, generator=_g(3), device="cuda"), torch.randn(C2, generator=_g(4), device="cuda")
        g2, b2 = torch.randn(C2, generator=_g(5), device="cuda"), torch.randn(C2, generator=_g(6), device="cuda")
        X2 = torch.zeros((B, N, C2), device="cuda")
        xn = torch.zeros((B, N, C2), device="cuda", dtype=BF16)
        m2, r2 = torch.zeros(B * N, device="cuda"), torch.zeros(B * N, device="cuda")
        k.layernorm_fwd(pe2, g1, b1, X2, 1e-5, B * HW, C2, ymap=(HW, N, 0), post_add=pos2, chain=(g2, b2, xn, m2, r2, 1e-6))
        y1 = F.layer_norm(pe2.float(), (C2,), g1, b1, 1e-5).view(B, HW, C2) + pos2
        _close(X2[:, :HW], y1, 1e-5, 1e-4, "chain: first norm")
        y2 = F.layer_norm(X2[:, :HW], (C2,), g2, b2, 1e-6)
        _close(xn[:, :HW], y2, 1e-2, 1e-2, "chain: second norm")
        _close(m2.view(B, N)[:, :HW], X2[:, :HW].mean(-1), 1e-4, 1e-5, "chain: mean")
        _close(r2.view(B, N)[:, :HW], (X2[:, :HW].var(-1, unbiased=False) + 1e-6).rsqrt(), 1e-3, 1e-5, "chain: rstd")
        assert float(xn[:, HW:].float().abs().max()) == 0.0 and float(m2.view(B, N)[:, HW:].abs().max()) == 0.0


def test_softmax_fwd_bwd():
    from mvlt_b200 import kernels as k
    rows, nk = 4096, 192
    s = (torch.randn((rows, nk), generator=_g(1), device="cuda") * 2).to(BF16)
    p = s.clone()
    k.softmax_fwd(p, rows, nk)
    ref = torch.softmax(s.float(), -1)
    _close(p, ref, 1e-2, 1e-4, "softmax")
    dp = torch.randn((rows, nk), generator=_g(2), device="cuda").to(BF16)
    ds = dp.clone()
    k.softmax_bwd(p, ds, rows, nk, 0.125)
    pf = p.float()
    refd = 0.125 * pf * (dp.float() - (pf * dp.float()).sum(-1, keepdim=True))
    _close(ds, refd, 2e-2, 1e-3, "softmax bwd")


def test_colsum_patchify_copy_rows():
    from mvlt_b200 import kernels as k
    x = torch.randn((5000, 320), generator=_g(1), device="cuda").to(BF16)
    out = torch.ones(320, device="cuda")
    k.colsum(x, 5000, 320, 320, out)
    _close(out, 1 + x.float().sum(0), 1e-3, 1e-2, "colsum")
    xv = torch.randn((100, 30528), generator=_g(2), device="cuda").to(BF16)[:, :30522]
    out = torch.zeros(30522, device="cuda")
    k.colsum(xv, 100, 30522, 30528, out)
    _close(out, xv.float().sum(0), 1e-3, 1e-2, "colsum vocab tail")
    B, H, W, C, R, T = 2, 16, 16, 64, 4, 128
    N = H * W + T
    tok = torch.randn((B, N, C), generator=_g(3), device="cuda")
    pat = torch.empty((B * (H // R) * (W // R), R * R * C), dtype=BF16, device="cuda")
    k.patchify(tok, N * C, pat, B, H, W, C, R)
    img = tok[:, :H * W].view(B, H, W, C).permute(0, 3, 1, 2)
    ref = F.unfold(img, R, stride=R).view(B, C, R * R, -1).permute(0, 3, 2, 1).reshape(pat.shape)   # (ky,kx,c) order
    assert torch.equal(pat, ref.to(BF16))
    back = torch.zeros((B, N, C), device="cuda")
    k.unpatchify(pat, back, N * C, B, H, W, C, R)
    assert torch.equal(back[:, :H * W], tok[:, :H * W].to(BF16).float()) and float(back[:, H * W:].abs().max()) == 0
    base = torch.randn((B, N, C), generator=_g(4), device="cuda")
    acc = base.clone()
    k.unpatchify(pat, acc, N * C, B, H, W, C, R, accumulate=True)        # += mode (the branch-parallel backward)
    assert torch.equal(acc[:, :H * W], base[:, :H * W] + tok[:, :H * W].to(BF16).float()) and torch.equal(acc[:, H * W:], base[:, H * W:])
    dst = torch.zeros((B * T, C), dtype=BF16, device="cuda")
    k.copy_rows(tok, dst, B * T, C, smap=(T, N, H * W))
    assert torch.equal(dst, tok[:, H * W:].reshape(B * T, C).to(BF16))


def test_patchify_nchw_matches_conv_layout():
    from mvlt_b200 import kernels as k
    B, P = 2, 4
    img = torch.rand((B, 3, 32, 32), generator=_g(1), device="cuda")
    pat = torch.empty((B * 64, 48), dtype=BF16, device="cuda")
    k.patchify_nchw(img, pat, B, 3, 32, 32, P, 48)
    ref = F.unfold(img, P, stride=P).permute(0, 2, 1).reshape(B * 64, 48)     # (ci,ky,kx) = conv weight flattening
    assert torch.equal(pat, ref.to(BF16))


@pytest.mark.parametrize("h,H", [(56, 64), (28, 32), (14, 16), (7, 8)])
def test_pos_resize_fwd_bwd(h, H):
    from mvlt_b200 import kernels as k
    C = 64
    tab = torch.randn((h * h, C), generator=_g(h), device="cuda")
    out = torch.empty((H * H, C), device="cuda")
    k.pos_resize_fwd(tab, out, h, h, H, H, C)
    tr = tab.clone().requires_grad_(True)
    ref = F.interpolate(tr.view(1, h, h, C).permute(0, 3, 1, 2), size=(H, H), mode="bilinear").reshape(1, C, H * H).permute(0, 2, 1)[0]
    _close(out, ref, 1e-5, 1e-5, "pos resize")
    g = torch.randn((H * H, C), generator=_g(9), device="cuda")
    ref.backward(g)
    dt = torch.zeros_like(tab)
    k.pos_resize_bwd(g, dt, h, h, H, H, C)
    _close(dt, tr.grad, 1e-4, 1e-4, "pos resize bwd")


def test_upsample2x_and_im2col():
    from mvlt_b200 import kernels as k
    B, h, w, C = 2, 8, 8, 64
    x = torch.randn((B, h, w, C), generator=_g(1), device="cuda").to(BF16)
    up = torch.empty((B, 2 * h, 2 * w, C), dtype=BF16, device="cuda")
    k.upsample2x_fwd(x, h * w * C, C, up, B, h, w, C)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    ref = F.interpolate(xr, scale_factor=2, mode="bilinear", align_corners=True)
    _close(up.float().permute(0, 3, 1, 2), ref, 1e-2, 1e-2, "up2")
    g = torch.randn((B, 2 * h, 2 * w, C), generator=_g(2), device="cuda").to(BF16)
    ref.backward(g.float().permute(0, 3, 1, 2))
    dx = torch.empty((B, h, w, C), dtype=F32, device="cuda")
    k.upsample2x_bwd(g, dx, h * w * C, C, B, h, w, C)
    _close(dx.permute(0, 3, 1, 2), xr.grad, 1e-4, 1e-4, "up2 bwd")
    col = torch.empty((B * h * w, 9 * C), dtype=BF16, device="cuda")
    k.im2col3x3(x, h * w * C, C, col, B, h, w, C)
    refc = F.unfold(x.float().permute(0, 3, 1, 2), 3, padding=1).view(B, C, 9, h * w).permute(0, 3, 2, 1).reshape(col.shape)
    assert torch.equal(col, refc.to(BF16))
    dxc = torch.empty((B, h, w, C), dtype=F32, device="cuda")
    k.col2im3x3(col, dxc, h * w * C, C, B, h, w, C)
    reff = F.fold(col.float().view(B, h * w, 9, C).permute(0, 3, 2, 1).reshape(B, C * 9, h * w), (h, w), 3, padding=1)
    _close(dxc.permute(0, 3, 1, 2), reff, 1e-5, 1e-4, "col2im")


def test_batchnorm_fwd_bwd_train_and_running_stats():
    from mvlt_b200 import kernels as k
    rows, C = 4096, 128
    x = (torch.randn((rows, C), generator=_g(1), device="cuda") * 2 + 0.5).to(BF16)
    gamma = 1 + 0.1 * torch.randn(C, generator=_g(2), device="cuda")
    beta = 0.1 * torch.randn(C, generator=_g(3), device="cuda")
    rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    st = torch.zeros((2, C), device="cuda")
    aff = torch.empty((4, C), device="cuda")
    k.bn_stats(x, rows, C, st[0], st[1])
    k.bn_finalize(st[0], st[1], rows, gamma, beta, rm, rv, 0.1, 1e-5, True, aff[0], aff[1], aff[2], aff[3], C)
    y = torch.empty((rows, C), dtype=BF16, device="cuda")
    k.bn_apply(x, aff[0], aff[1], y, C, 0, rows, C)
    xr = x.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm2, rv2 = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    ref = F.batch_norm(xr, rm2, rv2, gr, br, True, 0.1, 1e-5)
    _close(y, ref, 1e-2, 1e-2, "bn fwd")
    _close(rm, rm2, 1e-4, 1e-5, "running mean")
    _close(rv, rv2, 1e-3, 1e-4, "running var")
    dy = torch.randn((rows, C), generator=_g(4), device="cuda").to(BF16)
    ref.backward(dy.float())
    red = torch.zeros((2, C), device="cuda")
    dx = torch.empty((rows, C), dtype=BF16, device="cuda")
    k.bn_bwd(dy, x, aff[0], aff[2], aff[3], red[0], red[1], dx, rows, C, True)
    _close(dx, xr.grad, 2e-2, 2e-3, "bn dx")
    _close(red[0], br.grad, 1e-3, 1e-2, "bn dbeta")
    _close(red[1], gr.grad, 1e-3, 5e-2, "bn dgamma")


@pytest.mark.parametrize("n_cls,dtype", [(2, F32), (48, F32), (122, F32), (30522, BF16)])
def test_cross_entropy_fwd_bwd(n_cls, dtype):
    from mvlt_b200 import kernels as k
    rows = 257
    ld = (n_cls + 7) // 8 * 8
    logits = (torch.randn((rows, ld), generator=_g(n_cls), device="cuda") * 2).to(dtype)[:, :n_cls]
    labels = torch.randint(0, n_cls, (rows,), generator=_g(1), device="cuda")
    labels[::5] = -1
    n = int((labels != -1).sum())
    lse = torch.empty(rows, device="cuda")
    loss, total, corr = torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda")
    am = torch.empty(rows, dtype=torch.int32, device="cuda")
    k.ce_fwd(logits, ld, labels, rows, n_cls, -1, lse, loss, 1.0 / n, total_sum=total, argmax_out=am, correct=corr)
    lr = logits.float().clone().requires_grad_(True)
    ref = F.cross_entropy(lr, labels, ignore_index=-1)
    assert abs(loss.item() - ref.item()) < 1e-4 * max(1, abs(ref.item()))
    # loss and total are accumulated by independent fp32 atomics (order differs run to run): a few ulp apart
    assert abs(total.item() - loss.item()) < 1e-5 * max(1, abs(ref.item()))
    assert torch.equal(am.long(), logits.float().argmax(-1))
    assert int(corr.item()) == int(((logits.float().argmax(-1) == labels) & (labels != -1)).sum())
    ref.backward()
    dl = torch.empty((rows, ld), dtype=dtype, device="cuda")[:, :n_cls]
    gs = torch.tensor([2.0], device="cuda")
    k.ce_bwd(logits, ld, labels, rows, n_cls, -1, lse, dl, ld, 1.0 / n, gs)
    tol = 1e-2 if dtype == BF16 else 1e-5
    _close(dl, 2.0 * lr.grad, tol, 1e-7 if dtype == F32 else 1e-6, "ce bwd")


def test_t2i_upsample8_smoothl1_fwd_bwd():
    from mvlt_b200 import kernels as k
    B, h, w = 3, 32, 32
    score = torch.randn((B, h, w, 3), generator=_g(1), device="cuda") * 1.5
    target = torch.rand((B, 3, 256, 256), generator=_g(2), device="cuda")
    pred = torch.empty((B, 3, 256, 256), device="cuda")
    k.upsample8_fwd(score, pred, B, h, w, 8)
    sr = score.clone().requires_grad_(True)
    refp = F.interpolate(sr.permute(0, 3, 1, 2), scale_factor=8, mode="bilinear", align_corners=True)
    _close(pred, refp, 1e-5, 1e-5, "up8")
    ref = 10 * F.smooth_l1_loss(refp, target)
    ref.backward()
    numel = pred.numel()
    loss, total = torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda")
    k.t2i_up_loss(score, target, None, None, loss, total, 10.0 / numel, 0.0, None, B, h, w, 8, 1, False)
    assert abs(loss.item() - ref.item()) < 1e-4 * ref.item()
    ds = torch.zeros((B, h, w, 3), device="cuda")
    k.t2i_up_loss(score, target, None, ds, None, None, 0.0, 10.0 / numel, None, B, h, w, 8, 1, True)
    _close(ds, sr.grad, 1e-3, 1e-9, "t2i loss grad")
    g = torch.randn((B, 3, 256, 256), generator=_g(3), device="cuda")
    sr.grad = None
    refp2 = F.interpolate(sr.permute(0, 3, 1, 2), scale_factor=8, mode="bilinear", align_corners=True)
    refp2.backward(g)
    ds2 = torch.zeros((B, h, w, 3), device="cuda")
    k.t2i_up_loss(score, None, g, ds2, None, None, 0.0, 1.0, None, B, h, w, 8, 0, True)
    _close(ds2, sr.grad, 1e-3, 1e-4, "up8 bwd")


def test_bert_embed_fwd_bwd():
    from mvlt_b200 import kernels as k
    B, T, H, V = 4, 128, 768, 30522
    word = torch.randn((V, H), generator=_g(1), device="cuda") * 0.02
    pos = torch.randn((512, H), generator=_g(2), device="cuda") * 0.02
    typ = torch.randn((2, H), generator=_g(3), device="cuda") * 0.02
    gamma = 1 + 0.1 * torch.randn(H, generator=_g(4), device="cuda")
    beta = 0.1 * torch.randn(H, generator=_g(5), device="cuda")
    ids = torch.randint(1000, V, (B, T), generator=_g(6), device="cuda")
    ids[:, 40:] = 0
    out = torch.empty((B * T, H), dtype=BF16, device="cuda")
    mean, rstd = torch.empty(B * T, device="cuda"), torch.empty(B * T, device="cuda")
    k.bert_embed_fwd(ids, word, pos, typ, gamma, beta, out, mean, rstd, B * T, T, 1e-12, 0.0, 0)
    wr, pr, tr = word.clone().requires_grad_(True), pos.clone().requires_grad_(True), typ.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    x = F.embedding(ids, wr, padding_idx=0) + pr[:T] + tr[0]
    ref = F.layer_norm(x, (H,), gr, br, 1e-12)
    _close(out.view(B, T, H), ref, 1e-2, 1e-2, "bert embed")
    dy = (torch.randn((B * T, H), generator=_g(7), device="cuda") * 0.1).to(BF16)
    ref.backward(dy.float().view(B, T, H))
    dw, dp, dt = torch.zeros_like(word), torch.zeros_like(pos), torch.zeros_like(typ)
    dg, db = torch.zeros(H, device="cuda"), torch.zeros(H, device="cuda")
    k.bert_embed_bwd(dy, ids, word, pos, typ, gamma, mean, rstd, dw, dp, dt, dg, db, B * T, T, 0.0, 0)
    _close(dw, wr.grad, 1e-3, 1e-4, "dword")
    assert float(dw[0].abs().max()) == 0.0                       # padding row: no gradient from the gather (H6)
    _close(dp, pr.grad, 1e-3, 1e-4, "dpos")
    _close(dt, tr.grad, 1e-3, 1e-3, "dtype")
    _close(dg, gr.grad, 1e-3, 1e-3, "dgamma")
    _close(db, br.grad, 1e-3, 1e-3, "dbeta")
    # dropout: deterministic per seed, right keep rate, and mean-preserving scale
    o1, o2 = torch.empty_like(out), torch.empty_like(out)
    k.bert_embed_fwd(ids, word, pos, typ, gamma, beta, o1, mean, rstd, B * T, T, 1e-12, 0.1, 42)
    k.bert_embed_fwd(ids, word, pos, typ, gamma, beta, o2, mean, rstd, B * T, T, 1e-12, 0.1, 42)
    assert torch.equal(o1, o2)
    keep = float((o1 != 0).float().mean())
    assert abs(keep - 0.9) < 0.01


@pytest.mark.parametrize("B,n,stride_pad,acc", [(128, 64 * 512, 128 * 512, False), (128, 4096 * 64, 128 * 64, True),
                                                (5, 1024, 0, False), (33, 8192, 64, True)])
def test_batch_reduce_chunked(B, n, stride_pad, acc):
    """Position-embedding gradient reduction: batch split across CTAs with vector atomics when the row is short."""
    from mvlt_b200 import kernels as k
    stride = n + stride_pad
    x = torch.randn((B, stride), generator=_g(B + n), device="cuda")
    out = torch.randn(n, generator=_g(3), device="cuda")
    ref = x[:, :n].double().sum(0).float() + (out if acc else 0)
    k.batch_reduce(x, stride, B, n, out, accumulate=acc)
    _close(out, ref, 1e-5, 1e-4, "batch_reduce")


@pytest.mark.parametrize("Kp", [64, 48])      # 48 = the engine's geometry (float4 fast path), 64 = the padded generic kernel
def test_patchify_nchw_full_image(Kp):
    from mvlt_b200 import kernels as k
    B, P = 3, 4
    img = torch.rand((B, 3, 256, 256), generator=_g(11), device="cuda")
    pat = torch.full((B * 4096, Kp), 7.0, dtype=BF16, device="cuda")
    k.patchify_nchw(img, pat, B, 3, 256, 256, P, Kp)
    ref = F.unfold(img, P, stride=P).permute(0, 2, 1).reshape(B * 4096, 48)
    assert torch.equal(pat[:, :48], ref.to(BF16))
    if Kp > 48:
        assert float(pat[:, 48:].abs().max()) == 0.0


def test_token_mask_bit_exact_vs_oracle_and_reference_golden():
    """Device BERT word-piece masking == CPython random-driven reference loop, per sample seed (integer path: bit-exact)."""
    import os
    import numpy as np
    from mvlt_b200 import masking
    from oracle import token_mask as tm
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "token_mask_golden.npz"))
    ori = torch.from_numpy(g["ori"]).cuda()
    ids, labels = masking.mask_tokens_batch([int(s) for s in g["seeds"]], ori)
    assert torch.equal(ids.cpu(), torch.from_numpy(g["ids"])) and torch.equal(labels.cpu(), torch.from_numpy(g["labels"]))
    # a full-size batch against the oracle: 512 samples, ragged lengths incl. the extremes (1 piece, 126 pieces)
    rs = np.random.RandomState(3)
    B, T = 512, 128
    rows = np.zeros((B, T), dtype=np.int64)
    for b in range(B):
        L = [1, T - 2][b] if b < 2 else int(rs.randint(1, T - 1))
        rows[b, 0], rows[b, 1:1 + L], rows[b, 1 + L] = 101, rs.randint(1000, 30522, size=L), 102
    seeds = [masking.sample_seed(5, b) for b in range(B)]
    ids, labels = masking.mask_tokens_batch(seeds, torch.from_numpy(rows).cuda())
    ids, labels = ids.cpu().numpy(), labels.cpu().numpy()
    for b in range(B):
        i2, l2 = tm.mask_tokens(seeds[b], rows[b])
        assert (ids[b] == i2).all() and (labels[b] == l2).all(), b
    frac = (labels != -1).sum() / (rows[:, 1:] > 102).sum()
    assert 0.13 < frac < 0.17


def test_vl_scores_equal_live_reference_functions():
    """mvlt_b200.libs.vl_scores (reductions in csrc/loss.cu) next to the staged, unmodified reference's
    libs/vl_scores.py on the same CUDA tensors: MLM accuracy over labelled positions at the real vocabulary width, argmax scoring
    of 2 / 48 / 122-way heads, the single-logit sigmoid branch, PSNR."""
    import importlib.util
    import os
    from baseline import ref_loader
    from mvlt_b200.libs import vl_scores as S
    path = os.path.join(ref_loader.REF_DIR, "libs", "vl_scores.py")
    if not os.path.isfile(path):
        pytest.skip("baseline/_ref not staged")
    spec = importlib.util.spec_from_file_location("ref_vl_scores_gpu", path)
    R = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(R)
    g = torch.Generator(device="cuda").manual_seed(0)
    for trial in range(3):
        logits = torch.randn((2, 64, 30522), generator=g, device="cuda")
        target = torch.randint(0, 30522, (2, 64), generator=g, device="cuda")
        target[torch.rand((2, 64), generator=g, device="cuda") < 0.6] = -1
        am = logits.argmax(-1)
        keep = torch.rand((2, 64), generator=g, device="cuda") < 0.5
        target = torch.where((target != -1) & keep, am, target)        # about half of the labelled positions are "correct"
        target[0, 0] = am[0, 0]
        got, want = S.compute_mlm_score(logits, target), R.compute_mlm_score(logits, target)
        assert 0.2 < want < 0.9 and abs(got - want) < 1e-6, (got, want)
        for n_cls in (2, 48, 122):
            lg = torch.randn((37, n_cls), generator=g, device="cuda")
            lab = torch.where(torch.rand((37,), generator=g, device="cuda") < 0.5, lg.argmax(-1),
                              torch.randint(0, n_cls, (37,), generator=g, device="cuda"))
            assert torch.equal(S.compute_score_with_logits(lg, lab), R.compute_score_with_logits(lg, lab))
        lg1 = torch.randn((11, 1), generator=g, device="cuda")
        lab1 = torch.randint(0, 2, (11,), generator=g, device="cuda")
        assert torch.equal(S.compute_score_with_logits(lg1, lab1), R.compute_score_with_logits(lg1, lab1))
        a = torch.rand((2, 3, 256, 256), generator=g, device="cuda")
        b = torch.rand((2, 3, 256, 256), generator=g, device="cuda")
        assert abs(S.compute_psnr(a, b) - R.compute_psnr(a, b)) < 1e-3
    assert S.compute_psnr(a, a) == 100 == R.compute_psnr(a, a)
