"""GPU parity (through torch.ops.rs.* -> C ABI): gathers, fused fronts and their sparse backward against
the oracle and the reference-generated golden fixtures.  Integer / row-copy work is bit-exact; sums in a
different order get an fp32 tolerance written next to the assertion."""
from types import SimpleNamespace

import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden
from oracle import embed, towers as otowers

pytestmark = pytest.mark.gpu
DEV = "cuda"


def cu(x):
    if isinstance(x, dict):
        return {k: cu(v) for k, v in x.items()}
    return x.to(DEV) if torch.is_tensor(x) else x


# ------------------------------------------------------------------------------------------ gather_rows
@pytest.mark.parametrize("rows,dim,n", [(1000, 128, 4097), (37, 64, 513), (50, 768, 300), (11, 16, 1000),
                                        (4, 4, 77), (1371, 64, 1), (100, 128, 0), (300, 100, 257)])
def test_gather_rows_bit_exact(rs, rows, dim, n):
    g = torch.Generator().manual_seed(rows + dim + n)
    table = torch.randn(rows, dim, generator=g)
    ids = torch.randint(0, rows, (n,), generator=g)
    out = rs.gather_rows(table.to(DEV), ids.to(DEV))
    assert torch.equal(out.cpu(), embed.gather_rows(table, ids))


def test_gather_rows_shapes_dtypes_clamp(rs):
    g = torch.Generator().manual_seed(1)
    table = torch.randn(1001, 128, generator=g)
    deltas = torch.randint(0, 5000, (7, 9), generator=g)
    out = rs.gather_rows(table.to(DEV), deltas.to(DEV), clamp_max=1000)
    assert torch.equal(out.cpu(), embed.hybrid_time_rows(table, deltas))                 # H1 time_emb clamp
    bf = rs.gather_rows(table.to(DEV), deltas.clamp(max=1000).to(DEV), out_dtype=torch.bfloat16)
    assert torch.equal(bf.cpu(), table[deltas.clamp(max=1000)].to(torch.bfloat16))      # one rounding, no more
    t16 = table.to(torch.bfloat16)
    assert torch.equal(rs.gather_rows(t16.to(DEV), deltas.clamp(max=1000).to(DEV)).cpu(), t16[deltas.clamp(max=1000)])


def test_out_of_range_id_raises_index_error(rs):
    table = torch.randn(10, 128, device=DEV)
    rs.check_ids()
    rs.gather_rows(table, torch.tensor([3, 10], device=DEV))
    with pytest.raises(IndexError):
        rs.check_ids()
    rs.check_ids()       # flag was cleared


@pytest.mark.parametrize("deterministic", [True, False])
@pytest.mark.parametrize("rows,dim,n,pad", [(500, 128, 5000, 0), (12, 128, 3000, 0), (30522 // 10, 768, 2048, -1),
                                             (1371, 64, 900, -1), (7, 128, 40, 0), (100000, 128, 20000, 0)])
def test_embedding_backward_matches_aten(rs, rows, dim, n, pad, deterministic):
    g = torch.Generator().manual_seed(n)
    table = torch.randn(rows, dim, generator=g)
    # skewed ids: heavy hitters exercise the cross-tile segments
    ids = (torch.randint(0, rows, (n,), generator=g) * (torch.rand(n, generator=g) < 0.7)).long()
    ids[: n // 3] = min(3, rows - 1)
    cot = torch.randn(n, dim, generator=g)
    t = table.clone().requires_grad_(True)
    (F.embedding(ids, t, padding_idx=pad if pad >= 0 else None) * cot).sum().backward()
    rs.ops.DETERMINISTIC = deterministic
    try:
        tg = table.to(DEV).requires_grad_(True)
        (rs.gather_rows(tg, ids.to(DEV), padding_idx=pad) * cot.to(DEV)).sum().backward()
    finally:
        rs.ops.DETERMINISTIC = True
    # fp32 sums of up to n/3 terms in a different order: rtol 1e-5 on the row, atol scaled to the sum size
    torch.testing.assert_close(tg.grad.cpu(), t.grad, rtol=1e-5, atol=1e-5 * (n ** 0.5))
    if pad >= 0:
        assert tg.grad[pad].abs().sum() == 0


def test_sorted_backward_is_deterministic(rs):
    g = torch.Generator().manual_seed(5)
    ids = torch.randint(0, 50, (20000,), generator=g).to(DEV)
    cot = torch.randn(20000, 128, generator=g).to(DEV)
    a = torch.ops.rs.embedding_dense_bwd(cot, ids, 50, 0, -1, True)
    b = torch.ops.rs.embedding_dense_bwd(cot, ids.clone(), 50, 0, -1, True)
    assert torch.equal(a, b)


def test_sort_ids_is_a_stable_sort(rs):
    g = torch.Generator().manual_seed(6)
    for rows, n in ((105543, 409600), (12, 5000), (70000, 1), (1 << 20, 33333)):
        ids = torch.randint(0, rows, (n,), generator=g)
        sk, sp = rs.ops.sorted_ids(ids.to(DEV), rows)
        want_k, want_p = torch.sort(ids, stable=True)
        assert torch.equal(sk.cpu().long(), want_k) and torch.equal(sp.cpu().long(), want_p)


# ------------------------------------------------------------------------------------------ U1 / U2 / U3
@pytest.fixture(scope="module")
def ut():
    return load_golden("user_tower.pt")


def _product_tower(rs, ut):
    m = rs.SASRecUserTower(SimpleNamespace(**ut["args"]))
    m.load_state_dict(ut["state"], strict=True)
    return m.to(DEV).eval()


def test_seq_front_bit_exact_vs_reference(rs, ut):
    m = _product_tower(rs, ut)
    i = cu(ut["inputs"])
    with torch.no_grad():
        x = m.embed_front(i["pretrained_vecs"], i["item_ids"], i["time_bucket_ids"], i["type_ids"], i["color_ids"],
                          i["graphic_ids"], i["section_ids"])
    # item_proj runs in cuBLAS on the GPU and MKL in the fixture: feed the fixture's own base to compare the kernel
    base = F.linear(ut["inputs"]["pretrained_vecs"], ut["state"]["item_proj.weight"], ut["state"]["item_proj.bias"])
    gates = torch.sigmoid(ut["state"]["seq_gate"]) * torch.tensor(otowers.SEQ_GATE_MASK)
    ids = [i[k] for k in otowers.SEQ_INPUTS]
    tables = [ut["state"][n + ".weight"].to(DEV) for n in otowers.SEQ_TABLES]
    out = rs.seq_front(base.to(DEV), ids, tables, gates.to(DEV), ut["state"]["pos_emb.weight"].to(DEV))
    assert torch.equal(out.cpu(), ut["seq_front"])                           # same op order -> bit-exact
    torch.testing.assert_close(x.cpu(), ut["seq_front"], rtol=1e-5, atol=1e-5)   # only item_proj differs (GEMM order)
    # all six tables live (no masked gates): still bit-exact against the oracle
    gates6 = torch.linspace(0.2, 0.9, 6)
    want = embed.seq_front(base, [ut["inputs"][k] for k in otowers.SEQ_INPUTS],
                           [ut["state"][n + ".weight"] for n in otowers.SEQ_TABLES], gates6, ut["state"]["pos_emb.weight"])
    got = rs.seq_front(base.to(DEV), ids, tables, gates6.to(DEV), ut["state"]["pos_emb.weight"].to(DEV))
    assert torch.equal(got.cpu(), want)


def test_seq_front_sequence_shorter_or_longer_than_the_position_table(rs):
    """pos_emb has max_len rows; a batch with seq_len < max_len trains fine in the reference (rows arange(seq_len) get
    gradient, the others none), seq_len > max_len raises IndexError."""
    g = torch.Generator().manual_seed(3)
    B, L, D, max_len = 8, 20, 128, 50
    table = (torch.randn(300, D, generator=g) * 0.02)
    pos = (torch.randn(max_len, D, generator=g) * 0.02)
    ids = torch.randint(0, 300, (B, L), generator=g)
    base = torch.randn(B, L, D, generator=g)
    gates = torch.tensor([0.6])
    cot = torch.randn(B, L, D, generator=g)
    t0, p0 = table.clone().requires_grad_(True), pos.clone().requires_grad_(True)
    want = embed.seq_front(base, [ids], [t0], gates, p0)
    (want * cot).sum().backward()
    t1, p1 = table.to(DEV).requires_grad_(True), pos.to(DEV).requires_grad_(True)
    got = rs.seq_front(base.to(DEV), [ids.to(DEV)], [t1], gates.to(DEV), p1, out_dtype=torch.float32)
    assert torch.equal(got.detach().cpu(), want.detach())
    (got * cot.to(DEV)).sum().backward()
    assert p1.grad.shape == pos.shape
    torch.testing.assert_close(p1.grad.cpu(), p0.grad, rtol=1e-4, atol=1e-5)
    assert (p1.grad[L:] == 0).all()
    torch.testing.assert_close(t1.grad.cpu(), t0.grad, rtol=1e-4, atol=1e-5)
    with pytest.raises(IndexError):
        rs.seq_front(base.to(DEV), [ids.to(DEV)], [t1], gates.to(DEV), p1[:L - 1])


@pytest.mark.parametrize("deterministic", [True, False])
def test_seq_front_backward_vs_oracle(rs, deterministic):
    g = torch.Generator().manual_seed(9)
    B, L, D = 64, 50, 128
    rows = [5000, 12, 101, 101, 101, 101]
    tables = [torch.randn(r, D, generator=g) * 0.02 for r in rows]
    pos = torch.randn(L, D, generator=g) * 0.02
    ids = [torch.randint(0, r, (B, L), generator=g) for r in rows]
    ids[0][:, :20] = 0
    ids[0][:, 20:30] = 7                                           # heavy hitter
    base = torch.randn(B, L, D, generator=g)
    gate_raw = torch.linspace(-0.3, 0.6, 6)
    cot = torch.randn(B, L, D, generator=g)
    mask = torch.tensor(otowers.SEQ_GATE_MASK)
    # oracle
    ot = [t.clone().requires_grad_(True) for t in tables]
    op, ob, og = pos.clone().requires_grad_(True), base.clone().requires_grad_(True), gate_raw.clone().requires_grad_(True)
    (embed.seq_front(ob, ids, ot, torch.sigmoid(og) * mask, op) * cot).sum().backward()
    # product
    rs.ops.DETERMINISTIC = deterministic
    try:
        pt = [t.to(DEV).requires_grad_(True) for t in tables]
        pp, pb = pos.to(DEV).requires_grad_(True), base.to(DEV).requires_grad_(True)
        pg = gate_raw.to(DEV).requires_grad_(True)
        out = rs.seq_front(pb, [i.to(DEV) for i in ids], pt, torch.sigmoid(pg) * mask.to(DEV), pp, n_live=2)
        (out * cot.to(DEV)).sum().backward()
    finally:
        rs.ops.DETERMINISTIC = True
    tol = dict(rtol=1e-4, atol=2e-4)          # fp32 sums of up to B*L terms, different order
    for a, b, name in zip(pt, ot, otowers.SEQ_TABLES):
        torch.testing.assert_close(a.grad.cpu(), b.grad, msg=name, **tol)
    assert pt[0].grad[0].abs().sum() == 0                          # padding row never accumulates (invariant 1)
    assert all(pt[k].grad.abs().sum() == 0 for k in range(2, 6))   # masked gates: exact zeros (invariant 2)
    torch.testing.assert_close(pp.grad.cpu(), op.grad, **tol)
    torch.testing.assert_close(pb.grad.cpu(), ob.grad, rtol=0, atol=0)
    torch.testing.assert_close(pg.grad.cpu(), og.grad, rtol=1e-4, atol=1e-3)
    assert (pg.grad[2:] == 0).all()


def test_static_front_vs_reference(rs, ut):
    m = _product_tower(rs, ut)
    i = cu(ut["inputs"])
    names = otowers.STATIC_INPUTS
    with torch.no_grad():
        out = m.static_front(*[i[k] for k in names], i["cont_feats"])
    # through the module the gates come from torch.sigmoid on the GPU (1 ulp from the CPU's): tolerance
    torch.testing.assert_close(out.cpu(), ut["static_front"], rtol=1e-6, atol=1e-7)
    # with the fixture's own gate values the gathered columns are bit-exact; the 16 Linear(4->16)
    # columns differ by fma order only
    st = ut["state"]
    gates = torch.sigmoid(st["static_gate"]).to(DEV)
    out = rs.static_front([i[k] for k in names], [st[n + ".weight"].to(DEV) for n, _, _ in otowers.STATIC_TABLES],
                          i["cont_feats"], st["cont_proj.weight"].to(DEV), st["cont_proj.bias"].to(DEV), gates)
    assert torch.equal(out[:, :84].cpu(), ut["static_front"][:, :84])
    torch.testing.assert_close(out[:, 84:].cpu(), ut["static_front"][:, 84:], rtol=1e-6, atol=1e-6)


def test_static_front_backward_vs_oracle(rs):
    g = torch.Generator().manual_seed(3)
    B = 300
    spec = otowers.STATIC_TABLES
    tables = [torch.randn(r, d, generator=g) for _, r, d in spec]
    ids = [torch.randint(0, r, (B,), generator=g) for _, r, _ in spec]
    cont = torch.randn(B, 4, generator=g)
    W, bias, graw = torch.randn(16, 4, generator=g), torch.randn(16, generator=g), torch.randn(10, generator=g)
    cot = torch.randn(B, 100, generator=g)
    ot = [t.clone().requires_grad_(True) for t in tables]
    oW, ob, og = W.clone().requires_grad_(True), bias.clone().requires_grad_(True), graw.clone().requires_grad_(True)
    (embed.static_front(ids, ot, cont, oW, ob, torch.sigmoid(og)) * cot).sum().backward()
    pt = [t.to(DEV).requires_grad_(True) for t in tables]
    pW, pb, pg = (x.to(DEV).requires_grad_(True) for x in (W, bias, graw))
    out = rs.static_front([x.to(DEV) for x in ids], pt, cont.to(DEV), pW, pb, torch.sigmoid(pg))
    (out * cot.to(DEV)).sum().backward()
    tol = dict(rtol=1e-4, atol=1e-4)
    for a, b in zip(pt, ot):
        torch.testing.assert_close(a.grad.cpu(), b.grad, **tol)
        assert a.grad[0].abs().sum() == 0
    torch.testing.assert_close(pW.grad.cpu(), oW.grad, **tol)
    torch.testing.assert_close(pb.grad.cpu(), ob.grad, **tol)
    torch.testing.assert_close(pg.grad.cpu(), og.grad, **tol)


def test_user_tower_forward_backward_vs_reference(rs, ut):
    m = _product_tower(rs, ut)
    i = cu(ut["inputs"])
    out = m(**i, training_mode=True)
    # fp32 end to end; stock transformer/MLP layers run in cuBLAS vs MKL -> 1e-4
    torch.testing.assert_close(out.detach().cpu(), ut["out_train"], rtol=1e-3, atol=1e-4)
    ev = m(**i, training_mode=False)
    torch.testing.assert_close(ev.detach().cpu(), ut["out_eval"], rtol=1e-3, atol=1e-4)
    (out * ut["cotangent"].to(DEV)).sum().backward()
    grads = {k: p.grad.cpu() for k, p in m.named_parameters() if p.grad is not None}
    assert set(grads) == set(ut["grads"])
    for k, gref in ut["grads"].items():
        torch.testing.assert_close(grads[k], gref, rtol=2e-3, atol=2e-4, msg=k)
    assert grads["item_id_emb.weight"][0].abs().sum() == 0
    assert grads["type_emb.weight"].abs().sum() == 0 and (grads["seq_gate"][2:] == 0).all()


def test_user_tower_autocast_bf16(rs, ut):
    m = _product_tower(rs, ut)
    i = cu(ut["inputs"])
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = m(**i, training_mode=True)
    assert out.dtype == torch.float32                                   # F.normalize autocasts to fp32
    valid = ~ut["inputs"]["padding_mask"]
    # bf16 activations through 2 transformer layers: cosine agreement on the valid steps
    cos = F.cosine_similarity(out.detach().cpu()[valid], ut["out_train"][valid], dim=-1)
    assert cos.min() > 0.99


# ------------------------------------------------------------------------------------------ U4
def test_normalized_rows_vs_reference(rs):
    im = load_golden("item_matrix.pt")
    it = rs.SASRecItemTower(300, 128)
    it.load_state_dict(im["state"], strict=True)
    it = it.to(DEV)
    rows = it.normalized_rows(im["target_ids"].to(DEV))
    # norm summed in a different order than ATen's reduction: 1 ulp on the quotient
    torch.testing.assert_close(rows.detach().cpu(), im["rows"], rtol=3e-7, atol=1e-7)
    (rows * im["cotangent"].to(DEV)).sum().backward()
    torch.testing.assert_close(it.item_matrix.weight.grad.cpu(), im["grad_weight"], rtol=1e-4, atol=1e-6)
    assert torch.equal(it.get_log_q().cpu(), im["log_q"])


# ------------------------------------------------------------------------------------------ I1 / I2 / I3
def test_item_fronts_vs_reference(rs):
    g = load_golden("item_front.pt")
    s, i = cu(g["state"]), cu(g["inputs"])
    std = torch.ops.rs.std_front(s["std_embedding.weight"], i["std_input"], s["std_field_emb"], s["std_ln.weight"],
                                 s["std_ln.bias"], 1e-5, 0)
    torch.testing.assert_close(std.cpu(), g["std_out"], rtol=1e-5, atol=1e-5)
    e = "bert_model.embeddings."
    T = i["re_input_ids"].shape[-1]
    we = torch.ops.rs.bert_embed(s[e + "word_embeddings.weight"], s[e + "position_embeddings.weight"],
                                 s[e + "token_type_embeddings.weight"], s[e + "LayerNorm.weight"],
                                 s[e + "LayerNorm.bias"], g["bert_ln_eps"], i["re_input_ids"].reshape(-1, T), 0.0, 0, 0)
    torch.testing.assert_close(we.cpu(), g["word_embs"], rtol=1e-5, atol=1e-5)
    # masked mean over the reference's re_proj output
    h = F.gelu(F.layer_norm(F.linear(g["word_embs"], g["state"]["re_proj.0.weight"], g["state"]["re_proj.0.bias"]),
                            (128,), g["state"]["re_proj.1.weight"], g["state"]["re_proj.1.bias"]))
    mask = g["inputs"]["re_attn_mask"].reshape(-1, T)
    hp = h.to(DEV).requires_grad_(True)
    pooled = rs.masked_mean(hp, mask.to(DEV))
    torch.testing.assert_close(pooled.detach().cpu(), embed.masked_mean_pool(h, mask), rtol=1e-5, atol=1e-6)
    ho = h.clone().requires_grad_(True)
    cot = torch.randn(pooled.shape)
    (embed.masked_mean_pool(ho, mask) * cot).sum().backward()
    (pooled * cot.to(DEV)).sum().backward()
    torch.testing.assert_close(hp.grad.cpu(), ho.grad, rtol=1e-5, atol=1e-7)
    # dropout path: keep-rate and scaling (train-mode BertEmbeddings dropout, invariant 6)
    wd = torch.ops.rs.bert_embed(s[e + "word_embeddings.weight"], s[e + "position_embeddings.weight"],
                                 s[e + "token_type_embeddings.weight"], s[e + "LayerNorm.weight"],
                                 s[e + "LayerNorm.bias"], g["bert_ln_eps"], i["re_input_ids"].reshape(-1, T), 0.1, 123, 0)
    kept = wd != 0
    assert abs(kept.float().mean().item() - 0.9) < 0.01
    torch.testing.assert_close(wd[kept].cpu(), (g["word_embs"].to(DEV)[kept] / 0.9).cpu(), rtol=1e-4, atol=1e-5)


def test_word_embedding_scatter_768(rs):
    """I3: [B*32] rows x 768 scattered into the BERT word table."""
    g = torch.Generator().manual_seed(2)
    V, D, n = 3000, 768, 4096
    ids = torch.randint(0, V, (n,), generator=g)
    cot = torch.randn(n, D, generator=g)
    want = torch.zeros(V, D).index_add_(0, ids, cot)
    got = torch.ops.rs.embedding_dense_bwd(cot.to(DEV), ids.to(DEV), V, -1, -1, True)
    torch.testing.assert_close(got.cpu(), want, rtol=1e-5, atol=1e-4)


# ------------------------------------------------------------------------------------------ H1
def test_hybrid_user_gathers_vs_reference(rs):
    h = load_golden("hybrid_user.pt")
    t, i = h["tables"], h["inputs"]
    m = rs.HybridUserEmbeddings(t["gnn_user_emb"], t["gnn_item_emb"], t["item_content_emb"])
    with torch.no_grad():
        m.time_emb.weight.copy_(t["time_emb"]); m.channel_emb.weight.copy_(t["channel_emb"])
    m = m.to(DEV)
    out = m(i["u_idx"].to(DEV), i["seq_ids"].to(DEV), i["seq_deltas"].to(DEV), i["u_cat"].to(DEV))
    for k, v in h["gathered"].items():
        assert torch.equal(out[k].detach().cpu(), v), k
    # row 0 has no padding_idx here: it DOES receive gradient
    (out["item_content_emb"].sum()).backward()
    assert m.item_content_emb.weight.grad[0].abs().sum() > 0


# ------------------------------------------------------------------------------------------ full-size properties
def test_full_size_seq_front_properties(rs):
    """BASELINE config 2 sizes (B=8192, L=50): size-independent checks."""
    syn = rs.synthetic
    B, L, D, NI = 8192, 50, 128, syn.N_ITEMS
    b = syn.make_batch(B, L, NI)
    g = torch.Generator().manual_seed(0)
    item_tab = (torch.randn(NI + 1, D, generator=g) * 0.02).to(DEV)
    time_tab = (torch.randn(12, D, generator=g) * 0.02).to(DEV)
    pos = (torch.randn(L, D, generator=g) * 0.02).to(DEV)
    ids = [b["item_ids"].to(DEV), b["time_bucket_ids"].to(DEV)]
    gates = torch.tensor([0.7, 0.4], device=DEV)
    out = rs.seq_front(None, ids, [item_tab, time_tab], gates, pos)
    # linearity in the gates: f(2g) - f(g) == f(g) - f(0)
    out2 = rs.seq_front(None, ids, [item_tab, time_tab], 2 * gates, pos)
    out0 = rs.seq_front(None, ids, [item_tab, time_tab], 0 * gates, pos)
    torch.testing.assert_close(out2 - out, out - out0, rtol=1e-4, atol=1e-6)
    assert torch.equal(out0, pos.unsqueeze(0).expand(B, L, D))
    # spot rows against the definition
    for (bi, li) in ((0, 49), (8191, 0), (4096, 25)):
        want = (item_tab[ids[0][bi, li]] * gates[0] + time_tab[ids[1][bi, li]] * gates[1]) + pos[li]
        assert torch.equal(out[bi, li], want)
    # backward: column sums of the dense gradient equal the gate-weighted sum of cotangent rows (checksum)
    cot = torch.randn(B, L, D, generator=g).to(DEV)
    res = torch.ops.rs.seq_front_bwd(cot, ids, [item_tab, time_tab], gates, L, 0, True)
    live = (ids[0] != 0).unsqueeze(-1)
    torch.testing.assert_close(res[0].sum(0), 0.7 * (cot * live).sum((0, 1)), rtol=1e-3, atol=1e-2)
    torch.testing.assert_close(res[3], cot.sum(0), rtol=1e-4, atol=1e-3)
    assert res[0][0].abs().sum() == 0
    res2 = torch.ops.rs.seq_front_bwd(cot, ids, [item_tab, time_tab], gates, L, 0, False)     # atomics
    torch.testing.assert_close(res2[0], res[0], rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(res2[2], res[2], rtol=1e-3, atol=1e-1)
