"""CPU suite: the N>1 path (row-sharded lookup all-to-all, negatives all-gather with reduce-scatter
backward) on 2 gloo ranks.  The product's local gather/scatter are CUDA kernels, so the oracle's CPU
gather is injected for the owner-side step; what is under test is the routing and the collectives."""
import importlib
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rs = importlib.import_module(PKG)
        from oracle import embed, losses as olosses
        sh = rs.sharded
        torch.manual_seed(0)
        full = torch.randn(37, 8)
        shard = sh.shard_rows(full, rank, world).clone().requires_grad_(True)
        g = torch.Generator().manual_seed(100 + rank)
        ids = torch.randint(0, 37, (5, 3), generator=g)
        gather = lambda t, i: embed.gather_rows(t, i)
        scatter = lambda gr, i, rows: torch.zeros(rows, gr.shape[1]).index_add_(0, i, gr)
        out = sh.sharded_lookup(shard, ids, None, gather, scatter)
        ok = torch.equal(out.detach(), full[ids])
        w = torch.randn(5, 3, 8, generator=g)
        (out * w).sum().backward()
        # expected gradient of the FULL table = sum over ranks of the local index_add
        exp = torch.zeros(37, 8).index_add_(0, ids.reshape(-1), w.reshape(-1, 8))
        dist.all_reduce(exp)
        ok = ok and torch.allclose(shard.grad, sh.shard_rows(exp, rank, world), atol=1e-6)

        # negatives all-gather: local rows [B] against gathered [G*B]; compare with the oracle on the full problem
        B = 6
        gen = torch.Generator().manual_seed(7)
        U = torch.nn.functional.normalize(torch.randn(world * B, 16, generator=gen), dim=1)
        V = torch.nn.functional.normalize(torch.randn(world * B, 16, generator=gen), dim=1)
        tgt = torch.randint(1, 9, (world * B,), generator=gen)
        logq = torch.log(torch.rand(10, generator=gen) + 1e-3)
        u = U[rank * B:(rank + 1) * B].clone().requires_grad_(True)
        v = V[rank * B:(rank + 1) * B].clone().requires_grad_(True)

        def cpu_loss(ue, rows, t, uid, lq, temp, lam, col_rows, col_target_ids, col_user_ids, diag_offset):
            s = ue @ col_rows.T / temp - lq[col_target_ids].view(1, -1) * lam
            same = (t.view(-1, 1) == col_target_ids.view(1, -1)) | (uid.view(-1, 1) == col_user_ids.view(1, -1))
            lab = torch.arange(ue.shape[0]) + diag_offset
            same[torch.arange(ue.shape[0]), lab] = False
            return torch.nn.functional.cross_entropy(s.masked_fill(same, float("-inf")), lab)

        uid = torch.arange(B)
        loss = sh.cross_rank_logq_infonce(u, v, tgt[rank * B:(rank + 1) * B], uid, logq, 0.1, 1.0, None, cpu_loss)
        loss.backward()
        Uf, Vf = U.clone().requires_grad_(True), V.clone().requires_grad_(True)
        gid = torch.arange(world * B) // B * (1 << 24) + torch.arange(world * B) % B
        ref = olosses.inbatch_corrected_logq_loss(Uf, Vf, torch.arange(world * B), gid, torch.zeros(1), 0.1, 0.0) * 0
        # oracle C2 takes a table + ids: use V as the table and remap targets to keep the same-item mask
        s = Uf @ Vf.T / 0.1 - logq[tgt].view(1, -1)
        same = (tgt.view(-1, 1) == tgt.view(1, -1)) | (gid.view(-1, 1) == gid.view(1, -1))
        same.fill_diagonal_(False)
        per_row = torch.nn.functional.cross_entropy(s.masked_fill(same, float("-inf")), torch.arange(world * B),
                                                    reduction="none")
        full_loss = per_row.view(world, B).mean(dim=1)          # local means
        full_loss.sum().backward()
        ok = ok and torch.allclose(loss.detach(), full_loss[rank].detach(), atol=1e-5)
        ok = ok and torch.allclose(u.grad, Uf.grad[rank * B:(rank + 1) * B], atol=1e-5)
        ok = ok and torch.allclose(v.grad, Vf.grad[rank * B:(rank + 1) * B], atol=1e-5)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def _worker_catalog(rank, world, port, q):
    """planned lookup (loader-stage plan, step-stage row exchange) + catalogue-wide negatives with global
    multiplicities: the per-rank losses must add up to the reference loss of the concatenated global batch and
    the sharded gradients must be the shards of its gradient."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rs = importlib.import_module(PKG)
        from oracle import embed, losses as olosses
        sh = rs.sharded
        F = torch.nn.functional
        gen = torch.Generator().manual_seed(3)
        n_rows, D, n = 41, 16, 9                              # 41 rows: not a multiple of the world size
        full = torch.randn(n_rows, D, generator=gen)
        shard = sh.shard_padded(full, rank, world).clone().requires_grad_(True)
        R = sh.padded_rows(n_rows, world)
        ok = shard.shape[0] == R

        # ---- planned lookup with a leading unused row and a padding row that takes no gradient
        g = torch.Generator().manual_seed(50 + rank)
        ids = torch.randint(0, n_rows, (4, 5), generator=g)
        ids[0, 0] = 0
        plan = sh.plan_lookup(ids)
        gather = lambda t, i: embed.gather_rows(t, i)

        def scatter(gr, i, rows, pad):
            out = torch.zeros(rows, gr.shape[1]).index_add_(0, i, gr)
            if pad >= 0:
                out[pad] = 0
            return out
        out = sh.planned_lookup(shard, plan, None, gather, scatter, lead_rows=1, pad_local_row=0 if rank == 0 else -1)
        ok = ok and torch.equal(out[1:].detach(), full[ids.reshape(-1)]) and bool((out[0] == 0).all())
        w = torch.randn(ids.numel() + 1, D, generator=g)
        (out * w).sum().backward()
        exp = torch.zeros(n_rows, D).index_add_(0, ids.reshape(-1), w[1:])
        dist.all_reduce(exp)
        exp[0] = 0                                            # padding row never receives gradient
        ok = ok and torch.allclose(shard.grad, sh.shard_padded(exp, rank, world), atol=1e-6)

        # ---- catalogue-wide negatives
        U = F.normalize(torch.randn(world * n, D, generator=gen), dim=1)
        tgt = torch.randint(1, 12, (world * n,), generator=gen)          # heavy collisions
        uid_loc = torch.randint(0, 3, (world * n,), generator=gen)
        uid_glob = uid_loc + (torch.arange(world * n) // n) * 1000
        logq = torch.log(torch.rand(n_rows, generator=gen) + 1e-3)
        sl = slice(rank * n, (rank + 1) * n)
        u = U[sl].clone().requires_grad_(True)
        shard2 = sh.shard_padded(full, rank, world).clone().requires_grad_(True)
        cols = sh.CatalogColumns(n_rows, world, "cpu")
        ok = ok and torch.equal(torch.sort(cols.col_item_ids).values, torch.arange(world * R))
        ok = ok and torch.equal(cols.col_item_ids[cols.col_of(tgt)], tgt)
        v_cols = sh.all_gather_rows(F.normalize(shard2, dim=1))
        cnt = cols.counts(tgt[sl])
        ok = ok and torch.equal(cnt[cols.col_of(torch.arange(n_rows))], torch.bincount(tgt, minlength=n_rows).float())
        pos = cols.col_of(tgt[sl])
        own = torch.full((n, n), -1, dtype=torch.long)
        for i in range(n):
            js = torch.nonzero(uid_loc[sl] == uid_loc[sl][i]).squeeze(1)
            own[i, :js.numel()] = pos[js]
        lq = torch.zeros(cols.n_cols)
        lq[:n_rows] = logq
        loss = olosses.inbatch_corrected_logq_loss_columns(u, v_cols, cols.col_item_ids, cnt, tgt[sl], pos, own, lq,
                                                           0.1, 1.0) * (n / (world * n))
        loss.backward()
        Uf, Tf = U.clone().requires_grad_(True), full.clone().requires_grad_(True)
        ref = olosses.inbatch_corrected_logq_loss(Uf, F.normalize(Tf, dim=1), tgt, uid_glob, logq, 0.1, 1.0)
        ref.backward()
        tot = loss.detach().clone()
        dist.all_reduce(tot)
        ok = ok and torch.allclose(tot, ref.detach(), atol=1e-5)
        ok = ok and torch.allclose(u.grad, Uf.grad[sl], atol=1e-6)
        ok = ok and torch.allclose(shard2.grad, sh.shard_padded(Tf.grad, rank, world), atol=1e-6)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def _run_world(worker, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in range(world)]
    for p in procs:
        p.join(30)
    return sorted(res)


@pytest.mark.timeout(180)
def test_two_rank_catalog_negatives_and_planned_lookup():
    assert _run_world(_worker_catalog) == [(0, True), (1, True)]


@pytest.mark.timeout(180)
def test_two_rank_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in range(world)]
    for p in procs:
        p.join(30)
    assert sorted(res) == [(0, True), (1, True)]


def owner_compact_ref(cnt, world, R, cap):
    """torch restatement of rs_owner_compact (csrc/shard_route.cu) for the CPU tests: owner by owner, ascending local
    row, `cap` slots per owner.  Returns (rows [world*cap], ids, counts, slot_of [n_ids])."""
    n_ids = cnt.numel()
    rows = torch.full((world * cap,), -1, dtype=torch.int64)
    ids = torch.zeros(world * cap, dtype=torch.int64)
    counts = torch.zeros(world * cap)
    slot_of = torch.full((n_ids,), -1, dtype=torch.int32)
    for r in range(world):
        present = [i for i in range(r, n_ids, world) if cnt[i] > 0][:cap]
        for s, i in enumerate(present):
            o = r * cap + s
            rows[o], ids[o], counts[o], slot_of[i] = i // world, i, float(cnt[i]), o
    return rows, ids, counts, slot_of


def _worker_dedup(rank, world, port, q):
    """sharded.dedup_lookup: de-duplicated equal-split exchange (forward rows, backward gradient rows) on gloo."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rs = importlib.import_module(PKG)
        sh = rs.sharded
        torch.manual_seed(0)
        n_ids, D, cap = 41, 8, 24
        R = sh.padded_rows(n_ids, world)
        full = torch.randn(n_ids, D)
        shard = sh.shard_padded(full, rank, world).clone().requires_grad_(True)
        g = torch.Generator().manual_seed(200 + rank)
        tok_ids = torch.randint(0, n_ids, (60,), generator=g)            # many repeats: 60 tokens over <= 41 ids
        cnt = torch.bincount(tok_ids, minlength=n_ids)
        cnt[0] += 1                                                      # the padding id always owns slot 0
        req, _, _, slot_of = owner_compact_ref(cnt, world, R, cap)
        slots = slot_of[tok_ids].long()

        def gather(t, i):                                                # -1 = empty request slot -> zeros
            out = t[i.clamp(min=0)].clone()
            out[i < 0] = 0
            return out

        def scatter(gr, i, rows, pad):
            keep = (i >= 0) & (i != pad)
            return torch.zeros(rows, gr.shape[1]).index_add_(0, i[keep], gr[keep])
        buf = sh.dedup_lookup(shard, req, None, gather, scatter, pad_local_row=0 if rank == 0 else -1)
        ok = slot_of[0] == 0 and torch.equal(buf.detach()[slots], full[tok_ids])
        w = torch.randn(60, D, generator=g)
        (buf[slots] * w).sum().backward()
        # reference: dense gradient of the full table from ALL ranks' tokens, row 0 (padding) excluded
        ws, ts = [torch.zeros(60, D) for _ in range(world)], [torch.zeros(60, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(ws, w)
        dist.all_gather(ts, tok_ids)
        want = torch.zeros(n_ids, D)
        for w_r, t_r in zip(ws, ts):
            want.index_add_(0, t_r, w_r)
        want[0] = 0
        ok = ok and torch.allclose(shard.grad, sh.shard_padded(want, rank, world), atol=1e-6)
        # gather-to-full of the shards (state_dict hook)
        parts = [torch.empty_like(shard.data) for _ in range(world)]
        dist.all_gather(parts, shard.data)
        ok = ok and torch.equal(sh.unshard_rows(parts)[:n_ids], full)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_dedup_exchange():
    assert _run_world(_worker_dedup) == [(0, True), (1, True)]
