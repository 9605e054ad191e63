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


def _default_gather(table, ids):
    from . import ops
    return torch.ops.rs.gather_rows(table, ids, -1, ops.L.dt(table))


def _default_scatter(grad_rows, ids, rows):
    return torch.ops.rs.embedding_dense_bwd(grad_rows, ids, rows, -1, -1, True)


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
