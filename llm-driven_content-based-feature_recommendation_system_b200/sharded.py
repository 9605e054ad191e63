"""Multi-GPU plumbing for the path.  The reference is single-process; what shards is

  * the wide tables (`item_id_emb`, `item_matrix`, `gnn_user_emb`): row-sharded `owner = row % G`,
    looked up with ids all-to-all -> owner-side row gather (the CUDA kernel) -> rows all-to-all back;
    the backward sends gradient rows to their owners, which scatter-add them into their shard;
  * the in-batch negatives: item rows, target ids and user ids are all-gathered so that every rank
    scores its B users against G*B columns (`diag_offset = rank*B`); the gradient of the gathered
    columns is reduce-scattered back to the rank that owns them.

One process per GPU, `torch.distributed` (NCCL over NVLink on the box; gloo in the CPU tests, where
the local gather/scatter callables are injected because the product ops are CUDA-only).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist

from . import _lib as _L


def _default_gather(table, ids):
    from . import ops  # noqa: F401  (registers the ops)
    return _L.direct.gather_rows(table, ids, -1, _L.dt(table))


def _default_scatter(grad_rows, ids, rows):
    return _L.direct.embedding_dense_bwd(grad_rows, ids, rows, -1, -1, True)


def shard_rows(full: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Rows owned by `rank` under owner = row % world, in local order local = row // world."""
    return full[rank::world].contiguous()


def unshard_rows(shards) -> torch.Tensor:
    """Inverse of shard_rows over all ranks (gather-to-full for state_dict)."""
    world = len(shards)
    n = sum(s.shape[0] for s in shards)
    out = shards[0].new_empty(n, *shards[0].shape[1:])
    for r, s in enumerate(shards):
        out[r::world] = s
    return out


def route(ids: torch.Tensor, world: int):
    """Bucket a flat id vector by owner.  Returns (order, send_counts[world], local_rows sorted by owner)."""
    owner = ids % world
    order = torch.argsort(owner, stable=True)
    counts = torch.bincount(owner, minlength=world)
    return order, counts, (ids // world)[order]


class _ShardedLookup(torch.autograd.Function):
    @staticmethod
    def forward(ctx, shard, ids, group, gather_fn, scatter_fn):
        world = dist.get_world_size(group)
        flat = ids.reshape(-1)
        order, send_counts, local_rows = route(flat, world)
        recv_counts = torch.empty_like(send_counts)
        dist.all_to_all_single(recv_counts, send_counts, group=group)
        sc, rc = send_counts.tolist(), recv_counts.tolist()          # split sizes must live on the host
        req = local_rows.new_empty(sum(rc))
        dist.all_to_all_single(req, local_rows, rc, sc, group=group)                 # ids -> owners
        rows = gather_fn(shard, req)                                                 # owner-side gather
        back = rows.new_empty(flat.numel(), shard.shape[1])
        dist.all_to_all_single(back, rows.contiguous(), sc, rc, group=group)         # rows -> requesters
        out = torch.empty_like(back)
        out[order] = back
        ctx.save_for_backward(order, req)
        ctx.meta = (sc, rc, shard.shape[0], group, scatter_fn, shard.dtype)
        return out.view(*ids.shape, shard.shape[1])

    @staticmethod
    def backward(ctx, g):
        order, req = ctx.saved_tensors
        sc, rc, nrows, group, scatter_fn, sdt = ctx.meta
        g = g.reshape(-1, g.shape[-1])[order].contiguous()
        recv = g.new_empty(sum(rc), g.shape[1])
        dist.all_to_all_single(recv, g, rc, sc, group=group)                         # grad rows -> owners
        return scatter_fn(recv, req, nrows).to(sdt), None, None, None, None


def sharded_lookup(shard: torch.Tensor, ids: torch.Tensor, group=None, gather_fn: Optional[Callable] = None,
                   scatter_fn: Optional[Callable] = None) -> torch.Tensor:
    """rows `full[ids]` of a table row-sharded as `shard = full[rank::world]` on every rank."""
    return _ShardedLookup.apply(shard, ids, group, gather_fn or _default_gather, scatter_fn or _default_scatter)


class _AllGatherRows(torch.autograd.Function):
    """all_gather along dim 0 whose backward reduce-scatters the gradient of the gathered copy."""

    @staticmethod
    def forward(ctx, x, group):
        world = dist.get_world_size(group)
        out = x.new_empty(world * x.shape[0], *x.shape[1:])
        dist.all_gather_into_tensor(out, x.contiguous(), group=group)
        ctx.group = group
        return out

    @staticmethod
    def backward(ctx, g):
        world = dist.get_world_size(ctx.group)
        out = g.new_empty(g.shape[0] // world, *g.shape[1:])
        if dist.get_backend(ctx.group) == "gloo":            # gloo has no reduce_scatter_tensor
            g = g.contiguous()
            dist.all_reduce(g, group=ctx.group)
            r = dist.get_rank(ctx.group)
            out.copy_(g[r * out.shape[0]:(r + 1) * out.shape[0]])
        else:
            dist.reduce_scatter_tensor(out, g.contiguous(), group=ctx.group)
        return out, None


def all_gather_rows(x: torch.Tensor, group=None) -> torch.Tensor:
    return _AllGatherRows.apply(x, group)


def all_gather_ids(x: torch.Tensor, group=None) -> torch.Tensor:
    world = dist.get_world_size(group)
    out = x.new_empty(world * x.shape[0])
    dist.all_gather_into_tensor(out, x.contiguous(), group=group)
    return out


def cross_rank_logq_infonce(user_emb, item_rows, target_ids, user_ids, log_q_tensor, temperature=0.1,
                            lambda_logq=1.0, group=None, loss_fn=None):
    """C2 with the negatives of every rank: this rank's [B] users against the all-gathered [G*B] item rows.
    Requires the same B on every rank.  Returns the LOCAL mean; average over ranks for the global loss
    (DDP's gradient averaging does exactly that)."""
    from . import losses
    rank = dist.get_rank(group)
    B = user_emb.shape[0]
    cols = all_gather_rows(item_rows, group)
    ct = all_gather_ids(target_ids, group)
    # user ids are only unique within a rank (batch-row indices): make them globally unique
    cu = all_gather_ids(user_ids + rank * (1 << 24), group)
    fn = loss_fn or losses.logq_infonce_rows
    return fn(user_emb, item_rows, target_ids, user_ids + rank * (1 << 24), log_q_tensor, temperature, lambda_logq,
              col_rows=cols, col_target_ids=ct, col_user_ids=cu, diag_offset=rank * B)


# ----------------------------------------------------------------------------------------------------
# Planned exchange: everything that depends only on the batch's ids (bucketing by owner, split sizes, the id
# all-to-all itself) is done once per batch, where the loader prepares it -- the training step then contains
# only the two row exchanges and no host synchronisation.
# ----------------------------------------------------------------------------------------------------
class LookupPlan:
    """Routing of one id tensor against a table row-sharded as owner = id % world, local row = id // world.
    order[n]        positions sorted by owner (stable)
    send / recv     per-rank split sizes (host lists): rows this rank asks rank r for / is asked for by rank r
    req[sum(recv)]  LOCAL rows the other ranks ask this rank for, grouped by requesting rank"""

    def __init__(self, order, send, recv, req, n):
        self.order, self.send, self.recv, self.req, self.n = order, send, recv, req, n

    def to(self, device, non_blocking=True):
        return LookupPlan(self.order.to(device, non_blocking=non_blocking), self.send, self.recv,
                          self.req.to(device, non_blocking=non_blocking), self.n)


def plan_lookup(ids: torch.Tensor, group=None) -> LookupPlan:
    """Build the plan for `ids` (any shape; host or device tensor).  Exchanges the split sizes and the requested
    rows with the other ranks (collective, synchronises: call it from the loader / prefetch stage)."""
    world = dist.get_world_size(group)
    flat = ids.reshape(-1)
    order, counts, local_rows = route(flat, world)
    recv_counts = torch.empty_like(counts)
    dist.all_to_all_single(recv_counts, counts, group=group)
    sc, rc = counts.tolist(), recv_counts.tolist()
    req = local_rows.new_empty(sum(rc))
    dist.all_to_all_single(req, local_rows, rc, sc, group=group)
    return LookupPlan(order, sc, rc, req, flat.numel())


class _PlannedLookup(torch.autograd.Function):
    @staticmethod
    def forward(ctx, shard, plan, group, gather_fn, scatter_fn, lead_rows, pad_local_row):
        rows = gather_fn(shard, plan.req)                                            # owner-side gather (CUDA kernel)
        back = rows.new_empty(plan.n, shard.shape[1])
        dist.all_to_all_single(back, rows.contiguous(), plan.send, plan.recv, group=group)   # rows -> requesters
        out = rows.new_zeros(lead_rows + plan.n, shard.shape[1]) if lead_rows else rows.new_empty(plan.n, shard.shape[1])
        out[lead_rows:].index_copy_(0, plan.order, back)
        ctx.plan, ctx.meta = plan, (shard.shape[0], group, scatter_fn, shard.dtype, lead_rows, pad_local_row)
        return out

    @staticmethod
    def backward(ctx, g):
        plan = ctx.plan
        nrows, group, scatter_fn, sdt, lead_rows, pad_local_row = ctx.meta
        g = g[lead_rows:].index_select(0, plan.order)
        recv = g.new_empty(sum(plan.recv), g.shape[1])
        dist.all_to_all_single(recv, g, plan.recv, plan.send, group=group)           # grad rows -> owners
        return scatter_fn(recv, plan.req, nrows, pad_local_row).to(sdt), None, None, None, None, None, None


def _scatter_pad(grad_rows, ids, rows, pad_local_row):
    return _L.direct.embedding_dense_bwd(grad_rows, ids, rows, pad_local_row, -1, True)


def planned_lookup(shard: torch.Tensor, plan: LookupPlan, group=None, gather_fn: Optional[Callable] = None,
                   scatter_fn: Optional[Callable] = None, lead_rows: int = 0, pad_local_row: int = -1) -> torch.Tensor:
    """rows `full[ids]` ([lead_rows + n, D], the first `lead_rows` rows zero and without gradient) of a table
    row-sharded as `shard = full[rank::world]`, with the routing precomputed by `plan_lookup`.  Backward: the
    gradient rows travel back to their owners, which reduce them into their shard with the deterministic
    sort + segment-reduce kernel (`pad_local_row`: a local row that never receives gradient -- the padding row
    0 lives on rank 0 as local row 0; -1 elsewhere)."""
    return _PlannedLookup.apply(shard, plan, group, gather_fn or _default_gather, scatter_fn or _scatter_pad,
                                lead_rows, pad_local_row)


# ----------------------------------------------------------------------------------------------------
# De-duplicated, equal-split exchange: a rank asks every owner for the DISTINCT rows it needs, through a request list of
# static capacity per owner built on the device (ops.owner_compact): no host-side split sizes, no synchronisation, the
# id exchange itself is part of the step.  -1 = empty request slot (the owner answers zeros, the gradient is dropped).
# ----------------------------------------------------------------------------------------------------
def _all_to_all_equal(x: torch.Tensor, group) -> torch.Tensor:
    out = torch.empty_like(x)
    dist.all_to_all_single(out, x.contiguous(), group=group)
    return out


class _DedupLookup(torch.autograd.Function):
    @staticmethod
    def forward(ctx, shard, req_rows, group, gather_fn, scatter_fn, pad_local_row):
        got = _all_to_all_equal(req_rows, group)                       # ids -> owners   ([world * cap] each way)
        rows = gather_fn(shard, got)                                   # owner-side gather (CUDA kernel; -1 -> zeros)
        buf = _all_to_all_equal(rows, group)                           # rows -> requesters: slot o = owner * cap + s
        ctx.save_for_backward(got)
        ctx.meta = (shard.shape[0], group, scatter_fn, shard.dtype, pad_local_row)
        return buf

    @staticmethod
    def backward(ctx, g):
        (got,) = ctx.saved_tensors
        nrows, group, scatter_fn, sdt, pad_local_row = ctx.meta
        recv = _all_to_all_equal(g.contiguous(), group)                # gradient rows -> owners
        return scatter_fn(recv, got, nrows, pad_local_row).to(sdt), None, None, None, None, None


def dedup_lookup(shard: torch.Tensor, req_rows: torch.Tensor, group=None, gather_fn: Optional[Callable] = None,
                 scatter_fn: Optional[Callable] = None, pad_local_row: int = -1) -> torch.Tensor:
    """[world * cap, D] buffer whose slot o = owner * cap + s holds row `req_rows[o]` of owner's shard (zeros for -1).
    Backward: the buffer's gradient travels back slot by slot; every owner reduces what it receives into its shard
    with the deterministic sort + segment-reduce kernel (rows requested by several ranks add up; -1 slots and
    `pad_local_row` are skipped)."""
    return _DedupLookup.apply(shard, req_rows, group, gather_fn or _default_gather, scatter_fn or _scatter_pad,
                              pad_local_row)


def padded_rows(n_rows: int, world: int) -> int:
    """Rows per shard when a table of `n_rows` is padded to a multiple of `world`."""
    return (n_rows + world - 1) // world


def shard_padded(full: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """`shard_rows` of the table zero-padded to world * padded_rows rows (equal shards: all-gather friendly)."""
    R = padded_rows(full.shape[0], world)
    out = full.new_zeros(R, *full.shape[1:])
    mine = full[rank::world]
    out[:mine.shape[0]] = mine
    return out


class CatalogColumns:
    """Column layout of the all-gathered, row-sharded item matrix (rank-major: column r*R + j holds item
    j*world + r) for `losses.logq_infonce_columns` with catalogue-wide negatives: every rank scores its rows
    against ALL items, weighted by how often each item occurs as a target anywhere on the box (the reference's
    in-batch negatives of the global batch, grouped by item).  Static shapes, no id exchange, no dedup."""

    def __init__(self, n_rows: int, world: int, device):
        self.world, self.R = world, padded_rows(n_rows, world)
        self.n_cols = self.world * self.R
        c = torch.arange(self.n_cols, device=device)
        self.col_item_ids = (c % self.R) * world + c // self.R       # ids >= n_rows are padding columns (count 0)

    def col_of(self, item_ids: torch.Tensor) -> torch.Tensor:
        """column index of an item id (elementwise; negative ids stay negative)."""
        col = (item_ids % self.world) * self.R + torch.div(item_ids, self.world, rounding_mode="floor")
        return torch.where(item_ids < 0, item_ids, col)

    def counts(self, target_ids: torch.Tensor, group=None) -> torch.Tensor:
        """occurrences of every column's item among the targets of ALL ranks (one small all-reduce)."""
        from .losses import count_ids
        cnt = count_ids(self.col_of(target_ids), self.n_cols)
        if self.world > 1:
            dist.all_reduce(cnt, group=group)
        return cnt
