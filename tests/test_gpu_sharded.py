"""GPU parity of the N>1 path (needs >= 2 GPUs; skipped otherwise): the row-sharded / catalogue-negatives step on
2 NCCL ranks against the single-process step on the concatenated global batch (eval-mode towers: no dropout)."""
import importlib
import os
import socket
import sys

import pytest
import torch

from conftest import PKG, ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        rs = importlib.import_module(PKG)
        syn, sh = rs.synthetic, rs.sharded
        n_items, B, SL = 3001, 48, 50                     # 3002 rows: padded shards are exercised for world = 4

        def build():
            torch.manual_seed(0)
            m = rs.SASRecUserTower(syn.tower_args(num_items=n_items, max_len=SL)).to(dev).eval()
            it = rs.SASRecItemTower(n_items, 128, syn.log_q(n_items)).to(dev)
            lk = syn.pretrained_table(n_items).to(dev)
            it.init_from_pretrained(lk)
            return m, it, lk

        parts = [syn.make_batch(B, SL, n_items, seed=11 + r) for r in range(world)]
        # ---- single process, global batch
        model, item, lookup = build()
        glob = {k: torch.cat([p[k] for p in parts]) for k in parts[0]}
        gb = rs.train.prepare_batch(rs.train.add_host_index(glob), dev)
        opt = torch.optim.SGD(list(model.parameters()) + list(item.parameters()), lr=0.0)
        t0, m0, c0 = rs.train.two_tower_step(model, item, gb, lookup, opt, columns="catalog")
        ref = dict(im=item.item_matrix.weight.grad.clone(), ie=model.item_id_emb.weight.grad.clone(),
                   op=model.output_proj[0].weight.grad.clone(), te=model.time_emb.weight.grad.clone(),
                   sg=model.seq_gate.grad.clone())
        # ---- sharded
        model, item, lookup = build()
        tr = rs.train.ShardedTwoTower(model, item)
        opt = torch.optim.SGD(list(model.parameters()) + list(item.parameters()), lr=0.0)
        lb = tr.plan(rs.train.prepare_batch(rs.train.add_host_index(parts[rank]), dev), "unique")
        t1, m1, c1 = tr.step(lb, lookup, opt)
        ok = True
        msgs = []

        def close(name, a, b, tol):
            nonlocal ok
            err, ref_mag = (a - b).abs().max().item(), b.abs().max().item()
            good = err <= tol * ref_mag + 1e-8
            ok = ok and good
            msgs.append(f"{name}: err {err:.3e} of {ref_mag:.3e} {'ok' if good else 'FAIL'}")

        close("main", m1, m0, 2e-3)
        close("cl", c1, c0, 2e-3)
        close("item_matrix", item.item_matrix.weight.grad, sh.shard_padded(ref["im"], rank, world), 3e-2)
        close("item_id_emb", model.item_id_emb.weight.grad, sh.shard_padded(ref["ie"], rank, world), 3e-2)
        close("output_proj", model.output_proj[0].weight.grad, ref["op"], 3e-2)
        close("time_emb", model.time_emb.weight.grad, ref["te"], 3e-2)
        close("seq_gate", model.seq_gate.grad, ref["sg"], 3e-2)
        # same step with every item as a column (static shapes, all-gathered item matrix)
        opt.zero_grad(set_to_none=True)
        lb2 = tr.plan(rs.train.prepare_batch(rs.train.add_host_index(parts[rank]), dev), "catalog")
        t2, m2, c2 = tr.step(lb2, lookup, opt)
        close("main(catalog)", m2, m0, 2e-3)
        close("item_matrix(catalog)", item.item_matrix.weight.grad, sh.shard_padded(ref["im"], rank, world), 3e-2)
        q.put((rank, ok, msgs))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_sharded_step_matches_single_process():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=500) for _ in range(world)]
    for p in procs:
        p.join(60)
    for rank, ok, msgs in sorted(res):
        print(rank, msgs)
    assert all(ok for _, ok, _ in res), res


def _worker_device(rank, world, port, q):
    """the device-routed sharded step (train.ShardedDeviceStep) through bench.py's own parity check"""
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        rs = importlib.import_module(PKG)
        import bench
        q.put((rank, bench.sharded_parity_check(rs, dev, rank, world)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_device_routed_sharded_step_matches_single_process():
    """de-duplicated equal-split exchange + box-wide columns from the all-reduced target histogram vs the single-process
    step on the concatenated batch (losses, all gradients, gather-to-full state_dict)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_device, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=500) for _ in range(world)]
    for p in procs:
        p.join(60)
    assert all(v.startswith("ok") for _, v in res), res
