"""GPU parity: fused tcgen05 softmax losses (C1, C2, C3, C5), retrieval (R1) and FM (F1).

Tolerances.  The similarity contraction runs with bf16 (or fp16) operands and fp32 accumulation, the
reference fixtures are fp32: with unit-norm rows |u.v| error <= ~2^-9 (bf16) / 2^-12 (fp16) per operand,
times 1/temperature <= 12.5 on the logits.  Loss: atol 2e-2 (bf16) / 3e-3 (fp16).  Gradients: compared
to the fp32 gradient with atol = 3% (bf16) / 0.5% (fp16) of the largest gradient entry.
"""
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden
from oracle import fm as ofm, losses as olosses, retrieval as oretr

pytestmark = pytest.mark.gpu
DEV = "cuda"
LOSS_TOL = {torch.bfloat16: 2e-2, torch.float16: 3e-3}
GRAD_TOL = {torch.bfloat16: 3e-2, torch.float16: 5e-3}
# relative Frobenius norm of the gradient error (a max-norm bound alone would hide a systematic error in the many
# small-magnitude rows).  The fixtures are fp32 end to end; rounding the unit-norm operands to 16 bits perturbs every
# logit by ~ 2^-9 * 10 (bf16) and hence every softmax weight by ~ 2 % (bf16) / 0.25 % (fp16), independently per entry.
FRO_TOL = {torch.bfloat16: 2e-2, torch.float16: 4e-3}


@pytest.fixture(scope="module")
def lg():
    return load_golden("losses.pt")


def _run(rs, fn, wrt, dtype, *a, **k):
    old = rs.losses.COMPUTE_DTYPE
    rs.losses.COMPUTE_DTYPE = dtype
    try:
        leaves = [x.to(DEV).requires_grad_(True) for x in wrt]
        loss = fn(*leaves, *a, **k)
        loss.backward()
        return loss.detach().cpu(), [x.grad.cpu() for x in leaves]
    finally:
        rs.losses.COMPUTE_DTYPE = old


def _cmp(got, want, dtype):
    loss, grads = got
    assert torch.isfinite(loss)
    assert abs(loss.item() - want["loss"].item()) < LOSS_TOL[dtype], (loss.item(), want["loss"].item())
    for g, w in zip(grads, want["grads"]):
        assert torch.isfinite(g).all()
        assert (g - w).abs().max() <= GRAD_TOL[dtype] * w.abs().max() + 1e-7, ((g - w).abs().max(), w.abs().max())
        assert (g - w).norm() <= FRO_TOL[dtype] * w.norm() + 1e-7, ((g - w).norm(), w.norm())


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_c1_simcse_vs_reference(rs, lg, dtype):
    c = lg["c1"]
    _cmp(_run(rs, rs.simcse_loss, [c["E1"], c["E2"]], dtype, c["temperature"]), c, dtype)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_c2_vs_reference(rs, lg, dtype):
    d = lambda k: lg[k].to(DEV)
    _cmp(_run(rs, rs.inbatch_corrected_logq_loss, [lg["U"], lg["table"]], dtype, d("tgt"), d("uid"), d("logq"),
              temperature=0.1, lambda_logq=1.0), lg["c2"], dtype)
    _cmp(_run(rs, rs.inbatch_corrected_logq_loss, [lg["U"], lg["table"]], dtype, d("tgt"), d("uid"), d("logq"),
              temperature=0.07, lambda_logq=0.0), lg["c2_nologq"], dtype)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_c3_vs_reference(rs, lg, dtype):
    d = lambda k: lg[k].to(DEV)
    _cmp(_run(rs, rs.duorec_loss_refined, [lg["U"], lg["U2"]], dtype, d("tgt"), temperature=0.1, lambda_sup=0.1),
         lg["c3"], dtype)
    _cmp(_run(rs, rs.duorec_loss_refined, [lg["U"], lg["U2"]], dtype, d("tgt"), temperature=0.1, lambda_sup=0.0),
         lg["c3_nosup"], dtype)
    _cmp(_run(rs, rs.duorec_loss_refined, [lg["U"], lg["U2"]], dtype, d("tgt_distinct"), temperature=0.1,
              lambda_sup=0.1), lg["c3_distinct"], dtype)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_c5_vs_reference(rs, lg, dtype):
    d = lambda k: lg[k].to(DEV)
    v = lg["table"][lg["tgt"]]
    _cmp(_run(rs, rs.logq_correction_loss, [lg["U"], v], dtype, d("tgt"), d("probs"), temperature=0.07,
              lambda_logq=0.5), lg["c5_logq"], dtype)
    _cmp(_run(rs, rs.efficient_corrected_logq_loss, [lg["U"], v], dtype, d("tgt"), d("logq"), temperature=0.1,
              lambda_logq=0.1), lg["c5_eff"], dtype)


@pytest.mark.parametrize("N,V", [(1, 5), (127, 40), (128, 1000), (129, 64), (1000, 300), (4096, 3000), (5000, 100000)])
def test_c2_sizes_vs_oracle(rs, N, V):
    """ragged sizes around the 128-tile edge, heavy collisions (V << N) and none (V >> N)."""
    g = torch.Generator().manual_seed(N)
    table = F.normalize(torch.randn(V, 128, generator=g), dim=1)
    tgt = torch.randint(0, V, (N,), generator=g)
    U = F.normalize(torch.randn(N, 128, generator=g) + 2.0 * table[tgt], dim=1)
    uid = torch.randint(0, max(2, N // 8), (N,), generator=g)
    logq = torch.log(torch.rand(V, generator=g) + 1e-6)
    u, t = U.clone().requires_grad_(True), table.clone().requires_grad_(True)
    want = olosses.inbatch_corrected_logq_loss(u, t, tgt, uid, logq, 0.1, 1.0)
    want.backward()
    got = _run(rs, rs.inbatch_corrected_logq_loss, [U, table], torch.bfloat16, tgt.to(DEV), uid.to(DEV), logq.to(DEV),
               temperature=0.1, lambda_logq=1.0)
    _cmp(got, dict(loss=want.detach(), grads=[u.grad, t.grad]), torch.bfloat16)


def _own_cols(pos_col, uid):
    """[N, K] columns of the same user's targets (-1 padded), built the slow obvious way."""
    n = uid.numel()
    groups = {}
    for i, u in enumerate(uid.tolist()):
        groups.setdefault(u, []).append(i)
    k = max(len(v) for v in groups.values())
    own = torch.full((n, k), -1, dtype=torch.long)
    for i, u in enumerate(uid.tolist()):
        js = groups[u]
        own[i, :len(js)] = pos_col[js]
    return own


def _columns_loss(rs, U, table, tgt, uid, logq, mode, temperature, lam, dtype=torch.bfloat16, unit_norm=False):
    """C2 through the distinct-item column form (losses.logq_infonce_columns), columns built on the host."""
    ids, counts, pos_col = rs.losses.item_columns(tgt, None if mode == "unique" else table.shape[0])
    own = _own_cols(pos_col, uid)

    def fn(u, t):
        return rs.logq_infonce_columns(u, rs.ops.gather_rows(t, ids.to(DEV)), ids.to(DEV), counts.to(DEV), tgt.to(DEV),
                                       pos_col.to(DEV), own.to(DEV), logq.to(DEV), temperature, lam, unit_norm=unit_norm)
    return _run(rs, fn, [U, table], dtype)


@pytest.mark.parametrize("mode", ["unique", "catalog"])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_c2_columns_vs_reference(rs, lg, mode, dtype):
    """the [N, U] distinct-item form equals the reference's [N, N] loss and gradients (golden fixture)."""
    _cmp(_columns_loss(rs, lg["U"], lg["table"], lg["tgt"], lg["uid"], lg["logq"], mode, 0.1, 1.0, dtype), lg["c2"], dtype)
    _cmp(_columns_loss(rs, lg["U"], lg["table"], lg["tgt"], lg["uid"], lg["logq"], mode, 0.07, 0.0, dtype),
         lg["c2_nologq"], dtype)


@pytest.mark.parametrize("N,V,mode", [(1, 5, "unique"), (127, 40, "catalog"), (129, 64, "unique"), (1000, 300, "unique"),
                                      (4096, 3000, "catalog"), (5000, 100000, "unique")])
def test_c2_columns_sizes_vs_oracle(rs, N, V, mode):
    g = torch.Generator().manual_seed(N + 1)
    table = F.normalize(torch.randn(V, 128, generator=g), dim=1)
    tgt = torch.randint(0, V, (N,), generator=g)
    U = F.normalize(torch.randn(N, 128, generator=g) + 2.0 * table[tgt], dim=1)
    uid = torch.randint(0, max(2, N // 8), (N,), generator=g)
    logq = torch.log(torch.rand(V, generator=g) + 1e-6)
    u, t = U.clone().requires_grad_(True), table.clone().requires_grad_(True)
    want = olosses.inbatch_corrected_logq_loss(u, t, tgt, uid, logq, 0.1, 1.0)
    want.backward()
    got = _columns_loss(rs, U, table, tgt, uid, logq, mode, 0.1, 1.0)
    _cmp(got, dict(loss=want.detach(), grads=[u.grad, t.grad]), torch.bfloat16)


def test_rectangular_with_diag_offset(rs):
    """[B, G*B] block of the cross-GPU negatives layout: label of row i is column i + off."""
    g = torch.Generator().manual_seed(4)
    B, G, r = 200, 3, 1
    U = F.normalize(torch.randn(B, 128, generator=g), dim=1)
    V = F.normalize(torch.randn(G * B, 128, generator=g), dim=1)
    tgt = torch.randint(1, 50, (G * B,), generator=g)
    uid = torch.arange(G * B)
    logq = torch.log(torch.rand(60, generator=g) + 1e-4)
    u, v = U.clone().requires_grad_(True), V.clone().requires_grad_(True)
    s = u @ v.T / 0.1 - logq[tgt].view(1, -1)
    lab = torch.arange(B) + r * B
    same = (tgt[lab].view(-1, 1) == tgt.view(1, -1)) | (uid[lab].view(-1, 1) == uid.view(1, -1))
    same[torch.arange(B), lab] = False
    want = F.cross_entropy(s.masked_fill(same, float("-inf")), lab)
    want.backward()
    ug, vg = U.to(DEV).requires_grad_(True), V.to(DEV).requires_grad_(True)
    got = rs.logq_infonce_rows(ug, vg[r * B:(r + 1) * B], tgt[lab].to(DEV), uid[lab].to(DEV), logq.to(DEV), 0.1, 1.0,
                               col_rows=vg, col_target_ids=tgt.to(DEV), col_user_ids=uid.to(DEV), diag_offset=r * B)
    got.backward()
    _cmp((got.detach().cpu(), [ug.grad.cpu(), vg.grad.cpu()]), dict(loss=want.detach(), grads=[u.grad, v.grad]),
         torch.bfloat16)


def test_full_size_infonce_properties(rs):
    """B = 8192 (BASELINE config 2): lse >= diag, lse bounded by log N + max logit, gradient rows of the
    softmax part sum to ~0 along the label direction, loss equals log N for identical columns."""
    g = torch.Generator().manual_seed(0)
    N = 8192
    U = F.normalize(torch.randn(N, 128, generator=g), dim=1).to(DEV)
    lse, diag, _, _ = rs.losses.fused_softmax_stats(U, U, 10.0)
    assert (lse >= diag - 1e-3).all() and (lse <= 10.0 + torch.log(torch.tensor(float(N))) + 1e-3).all()
    torch.testing.assert_close(diag, torch.full_like(diag, 10.0), rtol=0, atol=0.1)        # <u,u> = 1 (bf16 rounding)
    ones = F.normalize(torch.ones(N, 128), dim=1).to(DEV)
    loss = rs.info_nce(ones, ones, 0.1)
    torch.testing.assert_close(loss.cpu(), torch.log(torch.tensor(float(N))), rtol=1e-3, atol=1e-3)
    # gradient w.r.t. a: sum_j softmax_ij * b_j - b_i ; for identical rows it vanishes
    a = ones.clone().requires_grad_(True)
    rs.info_nce(a, ones, 0.1).backward()
    assert a.grad.abs().max() < 1e-4


# ------------------------------------------------------------------------------------------ R1
def test_retrieval_vs_reference(rs):
    r = load_golden("retrieval.pt")
    U, I = r["U"].to(DEV), r["I"].to(DEV)
    for k in (12, 20, 100, 500):
        sc, ids = rs.retrieve_topk(U, I, k)
        assert torch.equal(ids.cpu(), r[f"k{k}"]["ids"]), k                 # bit-exact ids (no ties in this fixture)
        torch.testing.assert_close(sc.cpu(), r[f"k{k}"]["scores"], rtol=0, atol=2e-6)
    sc, ids = rs.retrieve_topk(U, I, 20, mask_index0=True)
    assert torch.equal(ids.cpu(), r["gnn_k20"]["ids"]) and not (ids == 0).any()
    t = r["ties_k12"]
    sc, ids = rs.retrieve_topk(U, t["I"].to(DEV), 12)
    assert torch.equal(ids.cpu(), oretr.canonical_ids(t["scores"], t["ids"]))    # ties: lower id first


@pytest.mark.parametrize("nu,ni,k", [(1, 200, 5), (777, 105542, 12), (5000, 20000, 100), (33, 5000, 1000), (300, 130, 130)])
def test_retrieval_sizes_vs_oracle(rs, nu, ni, k):
    g = torch.Generator().manual_seed(nu + ni)
    U = F.normalize(torch.randn(nu, 128, generator=g), dim=1)
    I = F.normalize(torch.randn(ni, 128, generator=g), dim=1)
    sc_w, ids_w = oretr.retrieve_topk(U, I, k)
    sc, ids = rs.retrieve_topk(U.to(DEV), I.to(DEV), k)
    torch.testing.assert_close(sc.cpu(), sc_w, rtol=0, atol=3e-6)
    # the ids we return really have those scores (fp32 CPU re-scoring of OUR ids) ...
    rescored = torch.gather(U @ I.T, 1, ids.cpu())
    torch.testing.assert_close(rescored, sc_w, rtol=0, atol=3e-6)
    # ... and they are bit-exact wherever both neighbours in the ranking are further away than the fp32
    # summation noise (1e-5); inside such near-ties the order between MKL and the kernel is arbitrary
    d = (sc_w[:, :-1] - sc_w[:, 1:]) > 1e-5
    t = torch.ones(nu, 1, dtype=torch.bool)
    safe = torch.cat([t, d], 1) & torch.cat([d[:, : k - 1], t], 1)
    safe[:, -1] = False                       # the k-th entry competes with the unseen (k+1)-th
    assert safe.float().mean() > 0.5
    assert torch.equal(ids.cpu()[safe], ids_w[safe])
    srt = torch.sort(ids, dim=1).values
    assert (srt[:, 1:] != srt[:, :-1]).all()                                                      # no duplicates


def test_retrieval_all_rows_bit_exact_on_gap_separated_inputs(rs):
    """inputs are REGENERATED until every user's top-13 scores are separated by more than the fp32
    summation noise (1e-5); then the ids of ALL rows are compared, exactly, against the fp32 CPU oracle (105,542 items,
    k = 12)."""
    g = torch.Generator().manual_seed(77)
    nu, ni, k = 1024, 105542, 12
    I = F.normalize(torch.randn(ni, 128, generator=g), dim=1)
    U = F.normalize(torch.randn(nu, 128, generator=g), dim=1)
    for _ in range(20):
        sc = torch.topk(U @ I.T, k + 1, dim=1).values
        bad = ((sc[:, :-1] - sc[:, 1:]).min(dim=1).values <= 1e-5).nonzero().flatten()
        if bad.numel() == 0:
            break
        U[bad] = F.normalize(torch.randn(bad.numel(), 128, generator=g), dim=1)
    assert bad.numel() == 0
    sc_w, ids_w = oretr.retrieve_topk(U, I, k)
    sc, ids = rs.retrieve_topk(U.to(DEV), I.to(DEV), k)
    assert torch.equal(ids.cpu(), ids_w)                                   # every row, every rank
    torch.testing.assert_close(sc.cpu(), sc_w, rtol=0, atol=3e-6)


@pytest.mark.parametrize("nu,ni,k,mask0", [(1, 1024, 1, False), (130, 5000, 12, True), (3000, 105542, 12, False),
                                           (200, 20000, 32, False), (70000, 2048, 17, False)])
def test_retrieval_tensor_core_path_equals_fp32_path(rs, nu, ni, k, mask0):
    """bf16 tcgen05 candidate pass + exact re-scoring vs the fp32 kernel: same ids everywhere both are unambiguous, same
    scores; non-unit norms (the rounding bound scales with |u| |i|max), few users (column splits), k up to 32."""
    g = torch.Generator().manual_seed(nu + ni + k)
    U = (torch.randn(nu, 128, generator=g) * (0.2 + 3 * torch.rand(nu, 1, generator=g))).to(DEV)
    I = (torch.randn(ni, 128, generator=g) * (0.5 + torch.rand(ni, 1, generator=g))).to(DEV)
    sc0, id0 = rs.ops.retrieve_topk(U, I, k, mask0, tensor_cores=False)
    sc1, id1 = rs.ops.retrieve_topk(U, I, k, mask0, tensor_cores=True)
    torch.testing.assert_close(sc1, sc0, rtol=1e-5, atol=1e-5)
    gap = sc0.abs().max(dim=1, keepdim=True).values * 1e-5
    d = (sc0[:, :-1] - sc0[:, 1:]) > gap if k > 1 else torch.ones(nu, 0, dtype=torch.bool, device=DEV)
    t = torch.ones(nu, 1, dtype=torch.bool, device=DEV)
    safe = torch.cat([t, d], 1) & torch.cat([d, t], 1)
    safe[:, -1] &= (k == ni)                      # the k-th entry competes with the unseen (k+1)-th
    if k > 1:
        assert safe.float().mean() > 0.5
    assert torch.equal(id1[safe], id0[safe])
    if mask0:
        assert not (id1 == 0).any()
    # the ids it returns really carry those scores
    torch.testing.assert_close(torch.gather(U @ I.T, 1, id1), sc1, rtol=1e-4, atol=1e-4)


def test_retrieval_tensor_core_path_falls_back_on_overflow(rs):
    """every item equal -> every score ties -> every element passes the candidate threshold -> the lists overflow: the
    device flag routes the call through the exact kernel, ids = ties by ascending id."""
    g = torch.Generator().manual_seed(5)
    U = F.normalize(torch.randn(300, 128, generator=g), dim=1).to(DEV)
    I = F.normalize(torch.randn(1, 128, generator=g), dim=1).repeat(4096, 1).to(DEV)
    sc0, id0 = rs.ops.retrieve_topk(U, I, 12, False, tensor_cores=False)
    sc1, id1 = rs.ops.retrieve_topk(U, I, 12, False, tensor_cores=True)
    assert torch.equal(id1, id0) and torch.equal(id1[0].cpu(), torch.arange(12))
    torch.testing.assert_close(sc1, sc0, rtol=0, atol=1e-6)


# ------------------------------------------------------------------------------------------ F1
@pytest.mark.parametrize("k", [4, 16, 128])
def test_fm_vs_oracle(rs, k):
    syn = rs.synthetic
    vocab = [50, 7, 1000, 3, 64, 64, 20000][: (7 if k > 4 else 5)] + [64] * 32
    F_ = len(vocab)
    B = 1000
    ids = syn.make_fm_batch(B, vocab, seed=1)
    m = rs.FM(vocab, k=k, init_std=0.1)
    emb, lin, offs = m.embedding.detach().clone(), m.linear.detach().clone(), m.offsets.clone()
    e0, l0 = emb.clone().requires_grad_(True), lin.clone().requires_grad_(True)
    x = ofm.field_rows(ids, e0, offs)
    y_want = ofm.fm_second_order(x) + l0[ids + offs].squeeze(-1).sum(1)
    cot_y, cot_c = torch.randn(B), torch.randn(B, F_ * k)
    ((y_want * cot_y).sum() + (x.flatten(1) * cot_c).sum()).backward()
    m = m.to(DEV)
    y, concat = m(ids.to(DEV))
    torch.testing.assert_close(y.detach().cpu(), y_want.detach(), rtol=1e-4, atol=1e-5)
    assert torch.equal(concat.detach().cpu(), x.detach().flatten(1))                      # gathered rows bit-exact
    ((y * cot_y.to(DEV)).sum() + (concat * cot_c.to(DEV)).sum()).backward()
    torch.testing.assert_close(m.embedding.grad.cpu(), e0.grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(m.linear.grad.cpu(), l0.grad, rtol=1e-4, atol=1e-4)


def test_deepfm_full_size(rs):
    """BASELINE config 3: B=65536, F=39; FM identity against the explicit pairwise sum on a slice."""
    syn = rs.synthetic
    vocab = syn.criteo_vocab_sizes()
    ids = syn.make_fm_batch(65536, vocab)
    m = rs.DeepFM(vocab, k=16, init_std=0.05).to(DEV)
    p = m(ids.to(DEV))
    assert p.shape == (65536,) and torch.isfinite(p).all() and (p > 0).all() and (p < 1).all()
    y, concat = m.fm(ids.to(DEV))
    x = concat[:512].view(512, 39, 16).double().cpu()
    lin = m.fm.linear.detach().cpu()[ids[:512] + m.fm.offsets.cpu()].squeeze(-1).sum(1).double()
    torch.testing.assert_close(y[:512].double().cpu(), ofm.fm_pairwise(x) + lin, rtol=1e-4, atol=1e-5)
    F.binary_cross_entropy(p, torch.full_like(p, 0.25)).backward()
    assert torch.isfinite(m.fm.embedding.grad).all() and m.fm.embedding.grad.abs().sum() > 0


# ------------------------------------------------------------------------------------------ C4 / C5 hard negatives
def test_mining_vs_oracle(rs):
    g = torch.Generator().manual_seed(12)
    for N, V, k in ((96, 50, 4), (1000, 400, 9), (3000, 100000, 29), (700, 30, 700 // 10)):
        table = F.normalize(torch.randn(V, 128, generator=g), dim=1)
        table[V // 2] = F.normalize(table[V // 2 - 1] + 0.05 * torch.randn(128, generator=g), dim=0)   # a too-similar pair
        tgt = torch.randint(0, V, (N,), generator=g)
        U = F.normalize(torch.randn(N, 128, generator=g) + table[tgt], dim=1)
        v = table[tgt]
        cos = U @ v.T
        same = tgt.view(-1, 1) == tgt.view(1, -1)
        sim = ((v @ v.T) > 0.9) & ~torch.eye(N, dtype=torch.bool)
        ign = same | sim
        want_s, want_i = torch.topk(cos.masked_fill(ign, float("-inf")), k, dim=1)
        sc, ids, avail = rs.ops.mine_hard_negatives(U.to(DEV), v.to(DEV), tgt.to(DEV), k, 0.9)
        assert torch.equal(avail.cpu().long(), (~ign).sum(1))
        fin = torch.isfinite(want_s)
        torch.testing.assert_close(sc.cpu()[fin], want_s[fin], rtol=0, atol=3e-6)
        assert (ids.cpu()[~fin] == -1).all()
        # every returned id is a legal (non-ignored) column with the score we report
        ii = ids.cpu().clamp(min=0)
        assert not torch.gather(ign, 1, ii)[fin].any()
        torch.testing.assert_close(torch.gather(cos, 1, ii)[fin], want_s[fin], rtol=0, atol=3e-6)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_c4_c5_hnm_vs_reference(rs, lg, dtype):
    d = lambda k: lg[k].to(DEV)

    def run(fn, *a, **k):
        old = rs.losses.COMPUTE_DTYPE
        rs.losses.COMPUTE_DTYPE = dtype
        try:
            leaves = [lg["U"].to(DEV).requires_grad_(True), lg["table"].to(DEV).requires_grad_(True)]
            loss, stats = fn(*leaves, *a, **k)
            loss.backward()
            return (loss.detach().cpu(), [x.grad.cpu() for x in leaves]), stats
        finally:
            rs.losses.COMPUTE_DTYPE = old

    got, stats = run(rs.full_batch_hard_emphasis_loss, d("tgt"), d("logq"), top_k_percent=0.05, hard_margin=0.01,
                     hnm_threshold=0.90, temperature=0.15, lambda_logq=1.0)
    _cmp(got, lg["c4"], dtype)
    assert stats["num_hard"] == lg["c4"]["stats"]["num_hard"]
    assert stats["avg_hn_similarity"] == pytest.approx(lg["c4"]["stats"]["avg_hn_similarity"], abs=1e-5)
    got, stats = run(rs.inbatch_hnm_corrected_loss_with_stats, d("tgt"), d("logq"), top_k_percent=0.05,
                     hnm_threshold=0.90, temperature=0.1, lambda_logq=0.7)
    _cmp(got, lg["c5_hnm"], dtype)
    assert stats["num_active_hard_negs"] == lg["c5_hnm"]["stats"]["num_active_hard_negs"]
    got, stats = run(rs.inbatch_mixed_hnm_loss_with_stats, d("tgt"), d("logq"), top_k_percent=0.05,
                     random_indices=d("mixed_random_indices"))
    _cmp(got, lg["c5_mixed"], dtype)


def test_c4_larger_vs_oracle(rs):
    g = torch.Generator().manual_seed(21)
    N, V = 3000, 2000
    table = F.normalize(torch.randn(V, 128, generator=g), dim=1)
    tgt = torch.randint(0, V, (N,), generator=g)
    U = F.normalize(torch.randn(N, 128, generator=g) + 1.5 * table[tgt], dim=1)
    logq = torch.log(torch.rand(V, generator=g) + 1e-6)
    u, t = U.clone().requires_grad_(True), table.clone().requires_grad_(True)
    want, wstats = olosses.full_batch_hard_emphasis_loss(u, t, tgt, logq, 0.01, 0.2, 0.9, 0.1, 1.0)
    want.backward()
    ug, tg = U.to(DEV).requires_grad_(True), table.to(DEV).requires_grad_(True)
    got, stats = rs.full_batch_hard_emphasis_loss(ug, tg, tgt.to(DEV), logq.to(DEV), 0.01, 0.2, 0.9, 0.1, 1.0)
    got.backward()
    _cmp((got.detach().cpu(), [ug.grad.cpu(), tg.grad.cpu()]), dict(loss=want.detach(), grads=[u.grad, t.grad]),
         torch.bfloat16)
    assert stats["num_hard"] == wstats["num_hard"]
    assert stats["avg_hn_similarity"] == pytest.approx(wstats["avg_hn_similarity"], abs=1e-4)


def test_train_step_column_modes_agree(rs):
    """One full train step (eval-mode towers: no dropout noise) gives the same main loss and the same
    item-matrix / user-tower gradients whichever way the in-batch softmax enumerates its columns."""
    syn = rs.synthetic
    n_items, B, SL = 3000, 96, 50
    torch.manual_seed(0)
    model = rs.SASRecUserTower(syn.tower_args(num_items=n_items, max_len=SL)).to(DEV).eval()
    item = rs.SASRecItemTower(n_items, 128, syn.log_q(n_items)).to(DEV)
    lookup = syn.pretrained_table(n_items).to(DEV)
    item.init_from_pretrained(lookup)
    batch = rs.train.prepare_batch(rs.train.add_host_index(syn.make_batch(B, SL, n_items, seed=5)), DEV)
    res = {}
    for mode in ("batch", "unique", "catalog"):
        model.zero_grad(set_to_none=True)
        item.zero_grad(set_to_none=True)
        opt = torch.optim.SGD(list(model.parameters()) + list(item.parameters()), lr=0.0)
        total, main, cl = rs.train.two_tower_step(model, item, batch, lookup, opt, columns=mode)
        res[mode] = (main.item(), cl.item(), item.item_matrix.weight.grad.clone(), model.item_id_emb.weight.grad.clone(),
                     model.output_proj[0].weight.grad.clone())
    for mode in ("unique", "catalog"):
        assert abs(res[mode][0] - res["batch"][0]) < 5e-3, (mode, res[mode][0], res["batch"][0])
        assert abs(res[mode][1] - res["batch"][1]) < 1e-5
        for g, w in zip(res[mode][2:], res["batch"][2:]):
            assert (g - w).abs().max() <= 2e-2 * w.abs().max() + 1e-8, (mode, (g - w).abs().max(), w.abs().max())


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_user_block_logits_vs_torch(rs, dtype):
    """per-user dense blocks: label logit, log-sum-exp over the user's OTHER items, and both gradients (fp32: the SIMT
    kernels; 16-bit operands: the mma.sync tile kernels with head + remainder coefficients -- same tolerance)."""
    g = torch.Generator().manual_seed(11)
    lens = [1, 5, 50, 2, 17, 33, 1, 8]
    n, n_cols = sum(lens), 40
    cu = torch.zeros(len(lens) + 1, dtype=torch.int32)
    cu[1:] = torch.cumsum(torch.tensor(lens), 0)
    U = F.normalize(torch.randn(n, 128, generator=g), dim=1).to(dtype).float()
    C = F.normalize(torch.randn(n_cols, 128, generator=g), dim=1).to(dtype).float()
    pos = torch.randint(0, n_cols, (n,), generator=g)
    bias = torch.randn(n_cols, generator=g)
    wp, wo = torch.randn(n, generator=g), torch.randn(n, generator=g)
    u0, c0 = U.clone().requires_grad_(True), C.clone().requires_grad_(True)
    s = u0 @ c0.T / 0.1 - bias.view(1, -1)
    s_pos0 = s[torch.arange(n), pos]
    own0 = []
    for b in range(len(lens)):
        r = torch.arange(cu[b], cu[b + 1])
        blk = s[r][:, pos[r]]                                          # [len, len]
        blk = blk.masked_fill(pos[r].view(-1, 1) == pos[r].view(1, -1), float("-inf"))
        own0.append(torch.logsumexp(blk, dim=1))
    own0 = torch.cat(own0)
    fin = torch.isfinite(own0)
    (s_pos0 * wp).sum().backward(retain_graph=True)
    (torch.where(fin, own0, torch.zeros_like(own0)) * wo).sum().backward()
    u1, c1 = U.to(DEV).requires_grad_(True), C.to(DEV).requires_grad_(True)
    s_pos, own = rs.ops.user_block_logits(u1, c1, pos.to(DEV), cu.to(DEV), 50, 10.0, bias.to(DEV), compute_dtype=dtype)
    tol = dict(rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(s_pos.detach().cpu(), s_pos0.detach(), **tol)
    assert torch.equal(torch.isfinite(own.detach().cpu()), fin)
    torch.testing.assert_close(own.detach().cpu()[fin], own0.detach()[fin], **tol)
    ((s_pos * wp.to(DEV)).sum() + (torch.where(fin.to(DEV), own, torch.zeros_like(own)) * wo.to(DEV)).sum()).backward()
    torch.testing.assert_close(u1.grad.cpu(), u0.grad, rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(c1.grad.cpu(), c0.grad, rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("N,V", [(129, 64), (1000, 300), (4096, 3000), (5000, 100000)])
def test_fixed_offset_and_folded_paths_vs_oracle(rs, N, V):
    """unit-norm operands -> fixed softmax offset in the forward, folded exponent in the backward (logit_bound > 0):
    same tolerance as the general path, for the [N, N] form, the distinct-item form and a finite-mask (C5) case."""
    g = torch.Generator().manual_seed(N + 7)
    table = F.normalize(torch.randn(V, 128, generator=g), dim=1)
    tgt = torch.randint(0, V, (N,), generator=g)
    U = F.normalize(torch.randn(N, 128, generator=g) + 2.0 * table[tgt], dim=1)
    uid = torch.randint(0, max(2, N // 8), (N,), generator=g)
    logq = torch.log(torch.rand(V, generator=g) + 1e-6)
    u, t = U.clone().requires_grad_(True), table.clone().requires_grad_(True)
    want = olosses.inbatch_corrected_logq_loss(u, t, tgt, uid, logq, 0.1, 1.0)
    want.backward()
    ref = dict(loss=want.detach(), grads=[u.grad, t.grad])

    def rows_fn(uu, tt):
        return rs.logq_infonce_rows(uu, rs.ops.gather_rows(tt, tgt.to(DEV)), tgt.to(DEV), uid.to(DEV), logq.to(DEV), 0.1,
                                    1.0, unit_norm=True)
    _cmp(_run(rs, rows_fn, [U, table], torch.bfloat16), ref, torch.bfloat16)
    _cmp(_columns_loss(rs, U, table, tgt, uid, logq, "unique", 0.1, 1.0, unit_norm=True), ref, torch.bfloat16)
    # the two kernel paths agree with each other much more tightly than with the fp32 oracle
    a = _run(rs, rows_fn, [U, table], torch.bfloat16)
    b = _run(rs, lambda uu, tt: rs.logq_infonce_rows(uu, rs.ops.gather_rows(tt, tgt.to(DEV)), tgt.to(DEV), uid.to(DEV),
                                                      logq.to(DEV), 0.1, 1.0), [U, table], torch.bfloat16)
    assert abs(a[0].item() - b[0].item()) < 1e-4
    for x, y in zip(a[1], b[1]):
        assert (x - y).abs().max() <= 1e-2 * y.abs().max() + 1e-8


def test_row_combine_kernel_matches_torch_expression(rs):
    """rs_ce_row_combine (per-row tail of logq_infonce_columns: value + three gradient vectors) against the torch
    expression it replaces, incl. zero-weight rows with non-finite inputs and rows without a same-user term."""
    g = torch.Generator().manual_seed(3)
    n = 5000
    lse0 = (torch.randn(n, generator=g) * 3).to(DEV).requires_grad_(True)
    pos = (torch.randn(n, generator=g) * 3).to(DEV).requires_grad_(True)
    own_v = pos.detach() - torch.rand(n, generator=g).to(DEV) * 4 - 0.05
    own_v[::7] = float("-inf")                                  # no other target of the same user
    own = own_v.clone().requires_grad_(True)
    w = torch.rand(n, generator=g).to(DEV) / n
    w[-100:] = 0
    mx = torch.maximum(lse0, pos).detach()
    z = torch.exp(lse0 - mx) + torch.exp(pos - mx) - torch.exp(own - mx)
    want = (((mx + torch.log(z.clamp_min(1e-30))) - pos) * w).sum()
    want.backward()
    ref = [lse0.grad.clone(), pos.grad.clone(), own.grad.clone()]
    lse0.grad = pos.grad = own.grad = None
    bad = lse0.detach().clone()
    bad[-50:] = float("nan")                                    # padding rows: weight 0 -> ignored whatever they hold
    bad.requires_grad_(True)
    got = rs.losses.row_combine(bad, pos, own, w)
    (got * 2.0).backward()
    torch.testing.assert_close(got, want.detach(), rtol=1e-5, atol=1e-6)
    for a, b in zip([bad.grad, pos.grad, own.grad], ref):
        torch.testing.assert_close(a, 2.0 * b, rtol=1e-4, atol=1e-9)
    # mean form, no same-user term
    got2 = rs.losses.row_combine(lse0.detach(), pos.detach(), None, None)
    mx = torch.maximum(lse0, pos).detach()
    want2 = (mx + torch.log(torch.exp(lse0 - mx) + torch.exp(pos - mx)) - pos).mean()
    torch.testing.assert_close(got2, want2.detach(), rtol=1e-5, atol=1e-6)


def test_columns_loss_fused_tail_matches_torch_tail(rs):
    """logq_infonce_columns with the fused per-row tail == with the torch tail (same kernels before it)."""
    g = torch.Generator().manual_seed(4)
    B, Lr, U, D = 64, 6, 300, 128
    n = B * Lr
    u = torch.nn.functional.normalize(torch.randn(n, D, generator=g), dim=1).to(DEV)
    v = torch.nn.functional.normalize(torch.randn(U, D, generator=g), dim=1).to(DEV)
    tgt_cols = torch.randint(0, U, (n,), generator=g)
    cid = (torch.randperm(5000, generator=g)[:U].sort().values).to(DEV)
    cnt = torch.bincount(tgt_cols, minlength=U).float().to(DEV)
    tgt = cid[tgt_cols.to(DEV)]
    row_cu = (torch.arange(B + 1, dtype=torch.int32) * Lr).to(DEV)
    logq = torch.randn(5000, generator=g).to(DEV) * 0.3 - 8
    w = torch.full((n,), 1.0 / n, device=DEV)
    outs = []
    for fused in (True, False):
        rs.losses.FUSE_ROW_COMBINE = fused
        try:
            uu, vv = u.clone().requires_grad_(True), v.clone().requires_grad_(True)
            loss = rs.losses.logq_infonce_columns(uu, vv, cid, cnt, tgt, tgt_cols.to(DEV), None, logq, 0.1, 1.0,
                                                  row_cu=row_cu, max_rows_per_user=Lr, unit_norm=True, row_weight=w)
            loss.backward()
            outs.append((loss.detach(), uu.grad, vv.grad))
        finally:
            rs.losses.FUSE_ROW_COMBINE = True
    torch.testing.assert_close(outs[0][0], outs[1][0], rtol=1e-5, atol=1e-5)
    for a, b in zip(outs[0][1:], outs[1][1:]):
        assert (a - b).norm() / b.norm() < 1e-3
