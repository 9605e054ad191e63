"""N1: the device-built, bucketed batch index (rs_batch_index_build) against the host index
(train.add_host_index -- torch nonzero / unique / cumsum, i.e. what the reference's boolean indexing computes,
tower_code/v1_usertower_train.py:794-804, :830) -- bit-exact on every real entry -- and the train step on it."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _both(rs, B, SL, n_items, seed, tok_q=256, col_q=128, ragged=False):
    syn = rs.synthetic
    hb = syn.make_batch(B, SL, n_items, seed=seed)
    if ragged:      # valid positions need not be contiguous for the builder (the reference always left-pads)
        g = torch.Generator().manual_seed(seed)
        hole = (torch.rand(B, SL, generator=g) < 0.15) & ~hb["padding_mask"]
        hole[:, -1] = False                                                  # keep every sequence non-empty
        hb["padding_mask"] = hb["padding_mask"] | hole
    host = rs.train.add_host_index(hb)
    dev = {k: v.to(DEV) for k, v in hb.items()}
    T, U = host["valid_index"].numel(), host["col_item_ids"].numel()
    tok_cap, col_cap = rs.train.bucket_of(T, U, tok_q, col_q)
    d = rs.train.device_index(dev, n_items + 1, tok_cap, col_cap)
    return hb, host, d, tok_cap, col_cap


@pytest.mark.parametrize("B,SL,n_items,ragged", [(1, 50, 100, False), (7, 50, 300, False), (64, 50, 3000, False),
                                                 (64, 50, 3000, True), (333, 64, 50, False), (2048, 50, 105542, False)])
def test_device_index_equals_host_index(rs, B, SL, n_items, ragged):
    hb, host, d, tok_cap, col_cap = _both(rs, B, SL, n_items, seed=B + SL, ragged=ragged)
    c = rs.train.check_index(d)
    T, U = host["valid_index"].numel(), host["col_item_ids"].numel()
    E = (host["cu_seqlens_2v"].numel() - 1 - 2 * B) // 2
    assert (c["tokens"], c["extras"], c["columns"]) == (T, E, U)
    G = d["fold_inv1"].numel()
    cpu = {k: v.cpu() for k, v in d.items() if isinstance(v, torch.Tensor)}
    # U1 grid: valid tokens, extras, zero padding
    for k in ("pk_item_ids", "pk_time_ids", "pk_pos_ids"):
        got, want = cpu[k].reshape(-1), host[k].reshape(-1)
        assert torch.equal(got[:T + E], want[:T + E]), k
        assert (got[T + E:] == 0).all(), k
    # encoder tokens: [v1 valid | v2 valid | v1 extras | v2 extras | padding]
    assert torch.equal(cpu["pk_index_2v"][:2 * T + 2 * E], host["pk_index_2v"])
    assert torch.equal(cpu["cu_seqlens_2v"][:2 * B + 1], host["cu_seqlens_2v"][:2 * B + 1])
    assert cpu["cu_seqlens_2v"][2 * B + 1].item() == 2 * G
    assert torch.equal(cpu["row_cu"], host["cu_seqlens"][:B + 1])
    # every U1 row is read by exactly its two encoder slots, and those cover all slots
    inv = torch.cat([cpu["fold_inv1"], cpu["fold_inv2"]])
    assert torch.equal(torch.sort(inv).values, torch.arange(2 * G))
    assert torch.equal(cpu["pk_index_2v"][cpu["fold_inv1"]], torch.arange(G))
    assert torch.equal(cpu["pk_index_2v"][cpu["fold_inv2"]], torch.arange(G))
    # rows the head selects
    sel, usr = cpu["select_2v"], cpu["users_2v"]
    assert torch.equal(sel[:T], host["select_2v_all"][:T]) and torch.equal(usr[:T], host["users_2v_all"][:T])
    assert torch.equal(sel[tok_cap:], host["select_2v_all"][T:]) and torch.equal(usr[tok_cap:], host["users_2v_all"][T:])
    # targets, columns, weights
    tgt = hb["target_ids"].reshape(-1)
    assert torch.equal(cpu["main_tgt"][:T], tgt[host["valid_index"]]) and (cpu["main_tgt"][T:] == 0).all()
    assert torch.equal(cpu["last_tgt"], tgt[host["last_index"]])
    assert torch.equal(cpu["col_item_ids"][:U], host["col_item_ids"]) and (cpu["col_item_ids"][U:] == 0).all()
    assert torch.equal(cpu["col_counts"][:U], host["col_counts"].float()) and (cpu["col_counts"][U:] == 0).all()
    assert torch.equal(cpu["pos_col"][:T], host["pos_col"])
    w = cpu["row_weight"]
    assert (w[:T] == 1.0 / T).all() and (w[T:] == 0).all()
    # the counts-only pre-pass agrees
    m = rs.ops.batch_index_counts(d["padding_mask"], d["target_ids"], n_items + 1).cpu()
    assert m[:3].tolist() == [T, E, U]


def test_overflow_and_bad_batches_are_flagged(rs):
    hb, host, d, tok_cap, col_cap = _both(rs, 64, 50, 3000, seed=3)
    dev = {k: v.to(DEV) for k, v in hb.items()}
    small = rs.train.device_index(dev, 3001, 256, col_cap)             # too few rows
    with pytest.raises(RuntimeError, match="overflow"):
        rs.train.check_index(small)
    small = rs.train.device_index(dev, 3001, tok_cap, 128)              # too few columns
    with pytest.raises(RuntimeError, match="overflow"):
        rs.train.check_index(small)
    bad = dict(dev)
    bad["padding_mask"] = dev["padding_mask"].clone()
    bad["padding_mask"][5] = True                                        # an empty sequence
    with pytest.raises(ValueError, match="empty"):
        rs.train.check_index(rs.train.device_index(bad, 3001, tok_cap, col_cap))
    with pytest.raises(IndexError):
        rs.train.check_index(rs.train.device_index(dev, 100, tok_cap, col_cap))      # targets beyond the table


def _models(rs, n_items, SL):
    syn = rs.synthetic
    torch.manual_seed(0)
    model = rs.SASRecUserTower(syn.tower_args(num_items=n_items, max_len=SL)).to(DEV).eval()
    item = rs.SASRecItemTower(n_items, 128, syn.log_q(n_items)).to(DEV)
    lookup = syn.pretrained_table(n_items).to(DEV)
    item.init_from_pretrained(lookup)
    return model, item, lookup


@pytest.mark.parametrize("tok_q,col_q", [(256, 128), (4096, 2048)])
def test_step_on_device_index_matches_step_on_host_index(rs, tok_q, col_q):
    """full train step (eval-mode towers: no dropout), bf16 autocast: exact host index vs bucketed device index (padding
    rows / columns / tokens must be inert): losses and every gradient."""
    n_items, B, SL = 3000, 64, 50
    model, item, lookup = _models(rs, n_items, SL)
    hb, host, d, _, _ = _both(rs, B, SL, n_items, seed=9, tok_q=tok_q, col_q=col_q)
    res = {}
    for name, batch in (("host", rs.train.prepare_batch(host, DEV)), ("device", d)):
        model.zero_grad(set_to_none=True)
        item.zero_grad(set_to_none=True)
        opt = torch.optim.SGD(list(model.parameters()) + list(item.parameters()), lr=0.0)
        t, mn, c = rs.train.two_tower_step(model, item, batch, lookup, opt)
        grads = {k: p.grad.clone() for k, p in list(model.named_parameters()) + list(item.named_parameters())
                 if p.grad is not None}
        res[name] = (mn.item(), c.item(), grads)
    assert abs(res["device"][0] - res["host"][0]) < 2e-3 and abs(res["device"][1] - res["host"][1]) < 2e-3, \
        (res["device"][:2], res["host"][:2])
    assert res["device"][2].keys() == res["host"][2].keys()
    for k, gref in res["host"][2].items():
        g = res["device"][2][k]
        assert torch.isfinite(g).all(), k
        # (the padded problem runs other GEMM tile shapes: bf16 activations differ in their last bit here and there)
        assert (g - gref).abs().max() <= 3e-2 * gref.abs().max() + 1e-7, (k, (g - gref).abs().max(), gref.abs().max())
        assert (g - gref).norm() <= 3e-2 * gref.norm() + 1e-7, (k, (g - gref).norm(), gref.norm())


def test_bucketed_graphs_serve_fresh_batches(rs):
    """one captured graph per shape bucket, replayed on batches it has never seen: same losses as the eager step on the
    host-indexed batch (lr = 0, eval-mode towers -> deterministic)."""
    syn, tr = rs.synthetic, rs.train
    n_items, B, SL = 3000, 64, 50
    model, item, lookup = _models(rs, n_items, SL)
    params = list(model.parameters()) + list(item.parameters())
    opt = torch.optim.AdamW(params, lr=0.0, weight_decay=0.0, fused=True, capturable=True)
    step = lambda b: tr.two_tower_step(model, item, b, lookup, opt)
    bs = tr.BucketedStep(step, B, SL, n_items + 1, DEV, use_graph=True, tok_q=256, col_q=128)
    seen = set()
    for seed in range(20, 32):
        hb = syn.make_batch(B, SL, n_items, seed=seed)
        fb = tr.FlatBatch(B, SL, pin=True).fill(hb)
        bs.raw.copy_(fb)
        t_, e_, u_ = bs.counts(bs.raw)[:3].tolist()
        key = bs.bucket(t_, u_)
        new = bs.ensure(key)
        seen.add(key)
        if new:
            continue                                # this batch was used to capture: replay on the NEXT ones only
        tot, mn, cl = bs.run(key)
        tr.check_index({**bs.index[key]})
        want = tr.two_tower_step(model, item, tr.prepare_batch(tr.add_host_index(hb), DEV), lookup, None)
        assert abs(mn.item() - want[1].item()) < 5e-3 and abs(cl.item() - want[2].item()) < 5e-3, (seed, key)
    assert len(bs.graphs) == len(seen) and len(bs.graphs) < 12


def _owner_compact_ref(cnt, world, R, cap):
    """plain-Python statement of rs_owner_compact: owner by owner, ascending local row, `cap` slots per owner"""
    n_ids = cnt.numel()
    rows = torch.full((world * cap,), -1, dtype=torch.int64)
    ids = torch.zeros(world * cap, dtype=torch.int64)
    counts = torch.zeros(world * cap)
    slot_of = torch.full((n_ids,), -1, dtype=torch.int32)
    sizes = []
    for r in range(world):
        present = [i for i in range(r, n_ids, world) if cnt[i] > 0]
        sizes.append(len(present))
        for s, i in enumerate(present[:cap]):
            o = r * cap + s
            rows[o], ids[o], counts[o], slot_of[i] = i // world, i, float(cnt[i]), o
    return rows, ids, counts, slot_of, sizes


@pytest.mark.parametrize("n_ids,world,cap,density", [(1, 1, 4, 1.0), (37, 2, 16, 0.5), (5000, 3, 900, 0.4), (105543, 8, 4096, 0.2),
                                                     (105543, 1, 30000, 0.2), (3001, 4, 64, 0.3), (1371981, 8, 2048, 0.005)])
def test_owner_compact_vs_reference(rs, n_ids, world, cap, density):
    g = torch.Generator().manual_seed(n_ids + world)
    cnt = (torch.rand(n_ids, generator=g) < density).to(torch.int32) * torch.randint(1, 50, (n_ids,), generator=g, dtype=torch.int32)
    R = (n_ids + world - 1) // world
    rows, ids, counts, slot_of, meta = rs.ops.owner_compact(cnt.to(DEV), world, R, cap, want_ids=True)
    w_rows, w_ids, w_counts, w_slot, sizes = _owner_compact_ref(cnt, world, R, cap)
    m = meta.cpu().tolist()
    assert m[0] == max(sizes) and m[1] == int(max(sizes) > cap) and m[2] == sum(sizes)
    assert torch.equal(rows.cpu(), w_rows) and torch.equal(ids.cpu(), w_ids) and torch.equal(counts.cpu(), w_counts)
    assert torch.equal(slot_of.cpu(), w_slot)
    # histogram + slot lookup
    toks = torch.randint(0, n_ids, (4000,), generator=g)
    h = rs.ops.id_histogram(toks.to(DEV), n_ids, force_bin0=True).cpu()
    want_h = torch.bincount(toks, minlength=n_ids).to(torch.int32)
    want_h[0] += 1
    assert torch.equal(h, want_h)
    nv = torch.tensor([1234], dtype=torch.int32, device=DEV)
    h2 = rs.ops.id_histogram(toks.to(DEV), n_ids, n_valid=nv).cpu()
    assert torch.equal(h2, torch.bincount(toks[:1234], minlength=n_ids).to(torch.int32))
    got = rs.ops.lookup_i32(slot_of, toks.to(DEV), fill=0).cpu()
    assert torch.equal(got, torch.where(w_slot[toks] >= 0, w_slot[toks].long(), torch.zeros(4000, dtype=torch.int64)))
