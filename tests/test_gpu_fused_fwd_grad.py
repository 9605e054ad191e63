"""GPU parity of the fused softmax at full size and of its "flash" form (rs_ce_fwd_grad / rs_ce_bwd_from_grad):
the forward pass that also accumulates the row side of the backward.

Reference here = plain fp32 torch on the GPU, computed from the SAME 16-bit-rounded operands in row chunks (the
[M, N] logits never exist in one piece), so the comparison isolates the kernels' own arithmetic: fp32 accumulation
order, the 16-bit rounding of P / dS before the second contraction, the polynomial 2^x on a share of the elements
(7.5e-5 relative).  Bounds: lse 2e-3 absolute; gradients max-norm 1.5 % of the largest entry AND relative Frobenius
norm 1e-2 (a max-norm bound alone would hide a systematic error in small-magnitude rows).
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"
NEG_INF = float("-inf")


def _ref_chunked(a16, b16, scale, bias, ka_row, ka_col, kb_row, kb_col, diag_offset, no_diag, w, w_diag=None,
                 chunk=4096):
    """fp32 reference of lse / diag and of dA, dB for loss = sum_i w_i * lse_i + w_diag_i * diag_i."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        A, B = a16.float(), b16.float()
        M, N = A.shape[0], B.shape[0]
        lse = torch.empty(M, device=A.device)
        diag = torch.zeros(M, device=A.device)
        dA = torch.empty_like(A)
        dB = torch.zeros_like(B)
        cols = torch.arange(N, device=A.device)
        for r0 in range(0, M, chunk):
            r1 = min(M, r0 + chunk)
            S = (A[r0:r1] @ B.T) * scale
            if bias is not None:
                S = S - bias.view(1, -1)
            rows = torch.arange(r0, r1, device=A.device)
            lab = rows + diag_offset
            is_lab = (cols.view(1, -1) == lab.view(-1, 1)) if not no_diag else torch.zeros_like(S, dtype=torch.bool)
            masked = torch.zeros_like(S, dtype=torch.bool)
            if ka_row is not None:
                masked |= ka_row[r0:r1].view(-1, 1) == ka_col.view(1, -1)
            if kb_row is not None:
                masked |= kb_row[r0:r1].view(-1, 1) == kb_col.view(1, -1)
            masked &= ~is_lab
            S = S.masked_fill(masked, NEG_INF)
            l = torch.logsumexp(S, dim=1)
            lse[r0:r1] = l
            P = torch.exp(S - l.view(-1, 1)) * w[r0:r1].view(-1, 1)
            if not no_diag:
                ok = (lab >= 0) & (lab < N)
                diag[r0:r1][ok] = S[ok, lab[ok]]
                if w_diag is not None:
                    P[ok, lab[ok]] += w_diag[r0:r1][ok]
            dA[r0:r1] = (P @ B) * scale
            dB += (P.T @ A[r0:r1]) * scale
        return lse, diag, dA, dB
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def _rel_fro(x, y):
    return ((x - y).norm() / y.norm().clamp_min(1e-30)).item()


def _check_grads(got, want, what, max_tol=1.5e-2, fro_tol=1e-2):
    assert torch.isfinite(got).all(), what
    mx = ((got - want).abs().max() / want.abs().max().clamp_min(1e-30)).item()
    fro = _rel_fro(got, want)
    assert mx <= max_tol and fro <= fro_tol, f"{what}: max-norm {mx:.3e} (<= {max_tol}), rel-Frobenius {fro:.3e} (<= {fro_tol})"
    return mx, fro


def _problem(M, N, seed, n_items=None, users=None, sort_cols=False):
    g = torch.Generator().manual_seed(seed)
    n_items = n_items or max(4, N // 2)
    b = F.normalize(torch.randn(N, 128, generator=g), dim=1)
    ka_col = torch.randint(1, n_items, (N,), generator=g)
    if sort_cols:
        ka_col = torch.sort(ka_col).values
    ka_row = ka_col[torch.randint(0, N, (M,), generator=g)]
    a = F.normalize(torch.randn(M, 128, generator=g) + 1.5 * b[torch.randint(0, N, (M,), generator=g)], dim=1)
    bias = torch.log(torch.rand(N, generator=g) * 0.01 + 1e-6)
    kb_row = torch.sort(torch.randint(0, users or max(2, M // 12), (M,), generator=g)).values
    w = torch.rand(M, generator=g) / M
    d = lambda t: t.to(DEV)
    return dict(a=d(a).bfloat16(), b=d(b).bfloat16(), bias=d(bias), ka_row=d(ka_row), ka_col=d(ka_col), kb_row=d(kb_row),
                w=d(w))


def _run_stats(rs, pr, fuse, scale=10.0, keys=True, bias=True, flags=0, diag_offset=0, w_diag=None, kb=False,
               logit_bound=True):
    Ls = rs.losses
    old = Ls.FUSE_ROW_GRAD
    Ls.FUSE_ROW_GRAD = fuse
    try:
        a = pr["a"].clone().requires_grad_(True)
        b = pr["b"].clone().requires_grad_(True)
        kw = {}
        if keys:
            kw.update(key_a_row=pr["ka_row"], key_a_col=pr["ka_col"])
        if kb:
            kw.update(key_b_row=pr["kb_row"], key_b_col=pr["kb_col"])
        lse, diag, _, _ = Ls.fused_softmax_stats(a, b, scale, col_bias=pr["bias"] if bias else None, diag_offset=diag_offset,
                                                 mask_value=NEG_INF, flags=flags, dtype=torch.bfloat16,
                                                 unit_norm=logit_bound, **kw)
        loss = (lse * pr["w"]).sum()
        if w_diag is not None:
            loss = loss + (diag * w_diag).sum()
        loss.backward()
        return lse.detach(), diag.detach(), a.grad.float(), b.grad.float()
    finally:
        Ls.FUSE_ROW_GRAD = old


@pytest.mark.parametrize("M,N", [(1, 5), (127, 129), (300, 300), (1000, 4100), (5000, 2500), (20000, 700)])
def test_flash_form_plain_vs_fp32(rs, M, N):
    """MODE_PLAIN (no bias, no keys), square and rectangular, with the label term (w_diag) of an InfoNCE."""
    pr = _problem(M, N, seed=M + N)
    g = torch.Generator().manual_seed(1)
    wd = (-torch.rand(M, generator=g) / M).to(DEV)
    got = _run_stats(rs, pr, True, keys=False, bias=False, w_diag=wd)
    lse0, diag0, dA0, dB0 = _ref_chunked(pr["a"], pr["b"], 10.0, None, None, None, None, None, 0, False, pr["w"], wd)
    torch.testing.assert_close(got[0], lse0, rtol=0, atol=2e-3)
    nd = min(M, N)
    torch.testing.assert_close(got[1][:nd], diag0[:nd], rtol=0, atol=2e-3)
    _check_grads(got[2], dA0, "dA")
    _check_grads(got[3], dB0, "dB")
    # and against the two-pass backward of the same library (rs_ce_bwd): same rounding points, tighter agreement
    old = _run_stats(rs, pr, False, keys=False, bias=False, w_diag=wd)
    torch.testing.assert_close(got[0], old[0], rtol=0, atol=1e-4)
    _check_grads(got[2], old[2], "dA vs two-pass", 1e-2, 5e-3)


@pytest.mark.parametrize("M,N,off", [(129, 64, 0), (1000, 3000, 1000), (4096, 4096, 0), (5000, 20000, 7000)])
def test_flash_form_rows_with_masks_vs_fp32(rs, M, N, off):
    """MODE_GENERAL with a label: column bias, same-item and same-user masks (-inf), diagonal tiles, diag_offset."""
    pr = _problem(M, N, seed=3 * M + N, users=max(2, M // 12))
    g = torch.Generator().manual_seed(2)
    # the label column of row i carries row i's keys (as in the in-batch loss: column i + off IS row i's positive)
    pr["kb_col"] = torch.randint(0, max(2, M // 12), (N,), generator=g).to(DEV)      # same-user hits off the label too
    lab = torch.arange(M, device=DEV) + off
    ok = lab < N
    pr["ka_row"][ok] = pr["ka_col"][lab[ok]]
    pr["kb_col"][lab[ok]] = pr["kb_row"][ok]
    wd = (-torch.rand(M, generator=g) / M).to(DEV)
    got = _run_stats(rs, pr, True, kb=True, diag_offset=off, w_diag=wd)
    lse0, diag0, dA0, dB0 = _ref_chunked(pr["a"], pr["b"], 10.0, pr["bias"], pr["ka_row"], pr["ka_col"], pr["kb_row"],
                                         pr["kb_col"], off, False, pr["w"], wd)
    torch.testing.assert_close(got[0], lse0, rtol=0, atol=2e-3)
    torch.testing.assert_close(got[1][ok], diag0[ok], rtol=0, atol=2e-3)
    _check_grads(got[2], dA0, "dA")
    _check_grads(got[3], dB0, "dB")


@pytest.mark.parametrize("M,N", [(1000, 300), (5000, 2500), (30000, 6000)])
def test_flash_form_distinct_columns_vs_fp32(rs, M, N):
    """RS_CE_NO_DIAG: sorted distinct-item columns with the row's own item masked by key (the main loss's launch)."""
    pr = _problem(M, N, seed=M - N, n_items=10 ** 6, sort_cols=True)
    pr["ka_col"] = torch.unique(pr["ka_col"])          # distinct, sorted
    N = pr["ka_col"].numel()
    pr["b"], pr["bias"] = pr["b"][:N].contiguous(), pr["bias"][:N].contiguous()
    pr["ka_row"] = pr["ka_col"][torch.randint(0, N, (M,), device=DEV)]
    got = _run_stats(rs, pr, True, flags=rs._lib.RS_CE_NO_DIAG)
    lse0, _, dA0, dB0 = _ref_chunked(pr["a"], pr["b"], 10.0, pr["bias"], pr["ka_row"], pr["ka_col"], None, None, 0, True,
                                     pr["w"])
    torch.testing.assert_close(got[0], lse0, rtol=0, atol=2e-3)
    _check_grads(got[2], dA0, "dA")
    _check_grads(got[3], dB0, "dB")


def test_flash_form_falls_back_on_device_when_the_exponent_range_is_large(rs):
    """a bias range beyond what the fixed-offset softmax can hold: g_info[1] == 0 on the device, the backward runs its
    own row-side pass -- same results as the two-pass path, no host decision involved."""
    pr = _problem(700, 900, seed=5)
    pr["bias"] = pr["bias"] * 40.0                      # |bias| up to ~550 nats
    a = _run_stats(rs, pr, True)
    b = _run_stats(rs, pr, False)
    torch.testing.assert_close(a[0], b[0], rtol=0, atol=1e-5)
    _check_grads(a[2], b[2], "dA vs two-pass", 2e-3, 1e-3)        # (lse differs in the last bit: other split plan)
    _check_grads(a[3], b[3], "dB vs two-pass", 2e-3, 1e-3)
    lse0, _, dA0, dB0 = _ref_chunked(pr["a"], pr["b"], 10.0, pr["bias"], pr["ka_row"], pr["ka_col"], None, None, 0, False,
                                     pr["w"])
    torch.testing.assert_close(a[0], lse0, rtol=0, atol=5e-3)
    _check_grads(a[2], dA0, "dA")


def test_fully_masked_rows_and_zero_weights(rs):
    """rows whose weight is 0 get an exactly-zero gradient; +inf bias columns (count 0) contribute nothing."""
    pr = _problem(500, 400, seed=9)
    pr["w"][::3] = 0.0
    pr["bias"][::5] = float("inf")
    got = _run_stats(rs, pr, True, flags=rs._lib.RS_CE_NO_DIAG)
    lse0, _, dA0, dB0 = _ref_chunked(pr["a"], pr["b"], 10.0, pr["bias"], pr["ka_row"], pr["ka_col"], None, None, 0, True,
                                     pr["w"])
    torch.testing.assert_close(got[0], lse0, rtol=0, atol=2e-3)
    assert (got[2][::3] == 0).all()
    assert (got[3][::5] == 0).all()
    _check_grads(got[2], dA0, "dA")
    _check_grads(got[3], dB0, "dB")


# ---------------------------------------------------------------------------------------------- BASELINE's own size
@pytest.mark.parametrize("M,N", [(103_976, 21_435), (103_976, 70_832)])
def test_full_size_main_loss_vs_fp32(rs, M, N):
    """The main loss's launch at BASELINE config 2's size (M = valid time steps of a B=8192 batch, N = distinct targets:
    21,435 on one GPU, 70,832 box-wide at 8 GPUs): every row's lse and the complete dU / dV against the chunked fp32
    reference.  Exercises the wave-aware column split, the range-test fast path and the fixed-offset softmax at the
    size they were tuned for."""
    g = torch.Generator().manual_seed(11)
    ids = torch.sort(torch.randperm(105_542, generator=g)[:N] + 1).values
    # Zipf-like multiplicities: bias = logq - log m
    m = torch.floor(1 + 200 * torch.rand(N, generator=g) ** 8)
    logq = torch.log(torch.rand(N, generator=g) * 1e-3 + 1e-6)
    b = F.normalize(torch.randn(N, 128, generator=g), dim=1)
    tgt_col = torch.randint(0, N, (M,), generator=g)
    a = F.normalize(torch.randn(M, 128, generator=g) + 1.5 * b[tgt_col], dim=1)
    w = torch.full((M,), 1.0 / M)
    pr = dict(a=a.to(DEV).bfloat16(), b=b.to(DEV).bfloat16(), bias=(logq - torch.log(m)).to(DEV), ka_row=ids[tgt_col].to(DEV),
              ka_col=ids.to(DEV), w=w.to(DEV))
    lse0, _, dA0, dB0 = _ref_chunked(pr["a"], pr["b"], 10.0, pr["bias"], pr["ka_row"], pr["ka_col"], None, None, 0, True,
                                     pr["w"], chunk=8192)
    for fuse in (True, False):
        got = _run_stats(rs, pr, fuse, flags=rs._lib.RS_CE_NO_DIAG)
        torch.testing.assert_close(got[0], lse0, rtol=0, atol=2e-3)
        ma, fa = _check_grads(got[2], dA0, f"dU (fuse={fuse})")
        mb, fb = _check_grads(got[3], dB0, f"dV (fuse={fuse})")
        print(f"[full size {M}x{N} fuse={fuse}] lse max err {(got[0] - lse0).abs().max().item():.2e}; "
              f"dU max-norm {ma:.2e} fro {fa:.2e}; dV max-norm {mb:.2e} fro {fb:.2e}")
    # 512 sampled rows once more, straight from the definition (no chunk bookkeeping shared with the helper above)
    rows = torch.randperm(M, generator=g)[:512].to(DEV)
    S = (pr["a"][rows].float() @ pr["b"].float().T) * 10.0 - pr["bias"].view(1, -1)
    S = S.masked_fill(pr["ka_row"][rows].view(-1, 1) == pr["ka_col"].view(1, -1), NEG_INF)
    torch.testing.assert_close(got[0][rows], torch.logsumexp(S, 1), rtol=0, atol=2e-3)
    assert math.isfinite(got[0].sum().item())
