"""One training step of the two-tower model, mirroring the body of `train_user_tower_all_time`
(tower_code/v1_usertower_train.py:729-859) without its per-step host synchronisations:

  * `pretrained_lookup[item_ids.cpu()].to(device)` (:760, a D2H sync + CPU gather + H2D of B*L*128
    floats) becomes a device-resident table gathered by the row-gather kernel;
  * `output_1[valid_mask]` (:797, boolean indexing = nonzero + sync) becomes a gather with the
    batch's precomputed valid-position index;
  * `F.normalize(item_matrix.weight)[targets]` (:810-811) normalises only the gathered rows;
  * the three `.item()` calls per step (:857-859) are dropped: losses are returned as device scalars.
"""
from __future__ import annotations

from typing import Dict, Optional

import contextlib

import torch
import torch.nn.functional as F
from torch.nn.attention import SDPBackend, sdpa_kernel

from . import encoder, losses, ops
from .synthetic import FORWARD_KEYS


def prepare_batch(batch: Dict[str, torch.Tensor], device, non_blocking=True) -> Dict[str, torch.Tensor]:
    """Host -> device copy of one collated batch (+ the flat index of its valid time steps, computed
    on the host where the padding mask already lives).  Pin `batch` for an asynchronous copy."""
    out = {k: v.to(device, non_blocking=non_blocking) for k, v in batch.items()}
    if "valid_index" not in batch:
        out["valid_index"] = torch.nonzero(~batch["padding_mask"].reshape(-1)).squeeze(1).to(device,
                                                                                             non_blocking=non_blocking)
    return out


def add_host_index(batch: Dict[str, torch.Tensor], columns: bool = True) -> Dict[str, torch.Tensor]:
    """Add `valid_index` / `last_index` (host tensors) so that the device step needs no nonzero(), and
    (columns=True) the distinct-item column set of the batch's targets for `losses.logq_infonce_columns`:
    `col_item_ids[U]`, `col_counts[U]`, `pos_col[N]`, `own_grid[B, L]` (column of the target at (b, l), -1 at
    padding).  All of it is index arithmetic on the collated batch, done where the ids live (the loader)."""
    valid = ~batch["padding_mask"]
    B, L = valid.shape
    batch = dict(batch)
    batch["valid_index"] = torch.nonzero(valid.reshape(-1)).squeeze(1)
    last = (valid.sum(dim=1) - 1).clamp(min=0)
    batch["last_index"] = torch.arange(B) * L + last
    batch["select_index"] = torch.cat([batch["valid_index"], batch["last_index"]])
    # packed-token view of the same batch (encoder.py): the valid time steps in batch-major order, their sequence
    # offsets, and the row DuoRec reads for every user.  The reference takes `last_indices = valid.sum(1) - 1`
    # (v1_usertower_train.py:830) on its LEFT-padded grid, i.e. column len-1 counted from the left: a padded
    # position whenever the sequence fills less than half of the window.  To stay a drop-in those positions are
    # carried as extra one-token pseudo-sequences behind the valid tokens (`packed_zero_tail` of them: every key of
    # such a query is masked in the reference, attention output 0).
    lens = valid.sum(dim=1)
    if bool((lens > 0).all()):
        T = int(lens.sum())
        pos2packed = torch.full((B * L,), -1, dtype=torch.int64)
        pos2packed[batch["valid_index"]] = torch.arange(T)
        lp = pos2packed[batch["last_index"]]
        extra = torch.nonzero(lp < 0).squeeze(1)                       # users whose "last" position is padding
        lp[extra] = T + torch.arange(extra.numel())
        cu = torch.zeros(B + 1 + extra.numel(), dtype=torch.int32)
        cu[1:B + 1] = torch.cumsum(lens, 0)
        cu[B + 1:] = T + 1 + torch.arange(extra.numel(), dtype=torch.int32)
        batch["packed_index"] = torch.cat([batch["valid_index"], batch["last_index"][extra]])
        batch["cu_seqlens"] = cu
        batch["last_packed"] = lp
        batch["select_packed"] = torch.cat([torch.arange(T), lp])
        # both dropout views in one pass: tokens [view-1 valid | view-2 valid | view-1 extras | view-2 extras]
        E = extra.numel()
        ext_pos = batch["last_index"][extra]
        batch["packed_index_2v"] = torch.cat([batch["valid_index"], batch["valid_index"], ext_pos, ext_pos])
        cu2 = torch.zeros(2 * B + 1 + 2 * E, dtype=torch.int32)
        cs = torch.cumsum(lens, 0).to(torch.int32)
        cu2[1:B + 1] = cs
        cu2[B + 1:2 * B + 1] = T + cs
        cu2[2 * B + 1:] = 2 * T + 1 + torch.arange(2 * E, dtype=torch.int32)
        batch["cu_seqlens_2v"] = cu2
        lp1 = torch.where(lp < T, lp, lp + T)                     # extras of view 1 start at 2T
        lp2 = torch.where(lp < T, lp + T, lp + T + E)             # view-2 valid tokens at T.., its extras at 2T+E..
        ar = torch.arange(B)
        batch["select_2v_all"] = torch.cat([torch.arange(T), lp1, lp2])       # main rows + DuoRec rows of both views
        batch["select_2v_last"] = torch.cat([lp1, lp1, lp2])                   # loss_scope == "last"
        batch["users_2v_all"] = torch.cat([batch["valid_index"] // L, ar, B + ar])
        batch["users_2v_last"] = torch.cat([ar, ar, B + ar])
        # U1 on packed tokens as well: [R, 64] grids holding the T valid tokens, then the E extras, zero-padded
        PK = 64
        R = (T + E + PK - 1) // PK
        flat = torch.cat([batch["valid_index"], ext_pos])

        def grid_of(x):
            g_ = torch.zeros(R * PK, dtype=torch.int64)
            g_[:T + E] = x
            return g_.view(R, PK)
        batch["pk_item_ids"] = grid_of(batch["item_ids"].reshape(-1)[flat])
        batch["pk_time_ids"] = grid_of(batch["time_bucket_ids"].reshape(-1)[flat])
        batch["pk_pos_ids"] = grid_of(flat % L + 1)
        tok = torch.arange(T)
        ext = T + torch.arange(E)
        batch["pk_index_2v"] = torch.cat([tok, tok, ext, ext])
    if columns:
        tgt = batch["target_ids"].reshape(-1)[batch["valid_index"]]
        ids, counts, pos_col = losses.item_columns(tgt)
        grid = torch.full((B * L,), -1, dtype=torch.int64)
        grid[batch["valid_index"]] = pos_col
        batch.update(col_item_ids=ids, col_counts=counts, pos_col=pos_col, own_grid=grid.view(B, L))
    return batch


# ------------------------------------------------------------------------------------------------------------
# Device-built, bucketed batch index: the collated [B, L] tensors go to the device as they are and
# one stream-ordered call (ops.batch_index_build -> rs_batch_index_build) derives every index the packed two-view step
# consumes, into arrays whose shapes depend only on the bucket (tok_cap, col_cap).  Padding rows carry loss weight 0,
# padding columns count 0, padding tokens sit in the attention kernel's zero tail -- so ONE captured graph serves
# every batch of a bucket, and nothing of the reference's per-step `nonzero` / `unique` work stays on the host.
# ------------------------------------------------------------------------------------------------------------
TOK_BUCKET = 2048          # main-loss rows (valid time steps) are padded to a multiple of this
COL_BUCKET = 512           # distinct-target columns are padded to a multiple of this


def bucket_of(n_tokens: int, n_cols: int, tok_q: int = TOK_BUCKET, col_q: int = COL_BUCKET):
    """(tok_cap, col_cap) of the shape bucket that holds a batch with `n_tokens` valid steps and `n_cols` distinct
    targets."""
    return ops.round_up(max(n_tokens, 1), tok_q), ops.round_up(max(n_cols, 1), col_q)


def device_index(batch: Dict[str, torch.Tensor], n_item_rows: int, tok_cap: int, col_cap: int,
                 out: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
    """`batch` (device tensors of one collated batch) + its device-built index for the bucket (tok_cap, col_cap).
    `out`: preallocated index arrays (ops.batch_index_alloc) to fill in place -- the static buffers of a graph."""
    B, L = batch["item_ids"].shape
    if out is None:
        out = ops.batch_index_alloc(B, L, tok_cap, col_cap, batch["item_ids"].device)
    ops.batch_index_build(batch["padding_mask"], batch["item_ids"], batch["time_bucket_ids"], batch["target_ids"],
                          n_item_rows, out)
    merged = dict(batch)
    merged.update(out)
    return merged


def check_index(batch: Dict[str, torch.Tensor]) -> Dict[str, int]:
    """Read the index's counters back (synchronises) and raise if a capacity was too small / the batch is malformed."""
    t, e, u, flags = (int(x) for x in batch["meta"][:4].tolist())
    if flags & 1:
        raise RuntimeError(f"batch index overflow: T={t}, E={e}, U={u} do not fit tok_cap={batch['main_tgt'].numel()}, "
                           f"col_cap={batch['col_item_ids'].numel()}")
    if flags & 2:
        raise ValueError("batch index: a user has an empty sequence")
    if flags & 4:
        raise IndexError("batch index: a target id lies outside the item table")
    return dict(tokens=t, extras=e, columns=u)


def _front_item_ids(batch, packed):
    """the item-id grid U1 runs on: the packed [R, 64] grid when the batch carries it, else the padded [B, L] one."""
    return batch["pk_item_ids"] if (packed and "pk_item_ids" in batch and ("cu_seqlens_2v" in batch)) else batch["item_ids"]


def _two_views(model, batch, pretrained_vecs, kw, loss_scope, packed, **extra):
    """The two dropout views (v1_usertower_train.py:788-792).  View 1 feeds the main loss (valid steps, or the last
    step) and DuoRec (last step); view 2 only DuoRec: the late-fusion head runs on exactly those rows (same values
    as slicing the full [B, L, 128] output).  Returns ([n_main + B, 128], [B, 128])."""
    if packed and "row_weight" in batch:          # device-built, bucketed index (device_index)
        out = model(pretrained_vecs=pretrained_vecs, **kw, training_mode=True, select_index=batch["select_2v"],
                    select_users=batch["users_2v"], cu_seqlens=batch["cu_seqlens_2v"], packed_zero_tail=1, views=2,
                    packed_index=batch["pk_index_2v"], packed_fold_inv=(batch["fold_inv1"], batch["fold_inv2"]),
                    select_prefix=batch["main_tgt"].numel(),
                    packed_inputs=dict(item_ids=batch["pk_item_ids"], time_bucket_ids=batch["pk_time_ids"],
                                       pos_ids=batch["pk_pos_ids"]), **extra)
        B = batch["item_ids"].shape[0]
        # (main rows, view 1's last steps, view 2's last steps): one split, one concatenation in the backward
        return ops.split_rows(out, out.shape[0] - 2 * B, B, B)
    if packed and "cu_seqlens_2v" in batch:
        B = batch["item_ids"].shape[0]
        k = "all" if loss_scope == "all" else "last"
        cu2 = batch["cu_seqlens_2v"]
        pk = {}
        if "pk_item_ids" in batch:
            T_ = batch["valid_index"].numel()
            pk = dict(packed_index=batch["pk_index_2v"], packed_fold=(T_, (batch["pk_index_2v"].numel() - 2 * T_) // 2),
                      packed_inputs=dict(item_ids=batch["pk_item_ids"], time_bucket_ids=batch["pk_time_ids"],
                                         pos_ids=batch["pk_pos_ids"]))
        else:
            pk = dict(packed_index=batch["packed_index_2v"])
        out = model(pretrained_vecs=pretrained_vecs, **kw, training_mode=True, select_index=batch["select_2v_" + k],
                    select_users=batch["users_2v_" + k], cu_seqlens=cu2,
                    packed_zero_tail=cu2.numel() - 1 - 2 * B, views=2, **pk, **extra)
        return out[:-B], out[-B:]
    li = batch["last_index"]
    sel1 = batch.get("select_index") if loss_scope == "all" else torch.cat([li, li])
    if sel1 is None:
        sel1 = torch.cat([batch["valid_index"], li])
    return (model(pretrained_vecs=pretrained_vecs, **kw, training_mode=True, select_index=sel1, **extra),
            model(pretrained_vecs=pretrained_vecs, **kw, training_mode=True, select_index=li, **extra))


def _losses_device_index(model, item_tower, batch, pretrained_vecs, kw, loss_scope, packed, columns, lambda_logq,
                         lambda_sup):
    """main (C2, all valid steps, distinct-item columns) + DuoRec on a batch indexed by `device_index`: static shapes --
    tok_cap main rows (weight 0 beyond the true count), col_cap columns (count 0 beyond it)."""
    L = batch["item_ids"].shape[1]
    tgt = batch["main_tgt"]
    n_main = tgt.numel()
    out_main, last1, last2 = _two_views(model, batch, pretrained_vecs, kw, loss_scope, packed)
    u = encoder.l2_normalize(out_main)                                               # :794-807
    cid = batch["col_item_ids"]
    v = item_tower.normalized_rows(cid)                                              # :810-811 + :833
    main = losses.logq_infonce_columns(u, v, cid, batch["col_counts"], tgt, batch["pos_col"], None,
                                       item_tower.get_log_q(), 0.1, lambda_logq, unit_norm=True,
                                       row_cu=batch["row_cu"], max_rows_per_user=L, row_weight=batch["row_weight"])
    cl = losses.duorec_loss_refined(last1, last2, batch["last_tgt"], lambda_sup=lambda_sup)   # :830-842
    return main, cl


def _losses_host_index(model, item_tower, batch, pretrained_vecs, kw, loss_scope, packed, columns, lambda_logq,
                       lambda_sup):
    """the same two losses on a batch indexed on the host (add_host_index): exact, data-dependent shapes."""
    B, L = batch["item_ids"].shape
    tgt_flat = batch["target_ids"].reshape(-1)
    idx = batch["valid_index"] if loss_scope == "all" else batch["last_index"]
    li = batch["last_index"]
    n_main = idx.numel()
    # view 1 feeds the main loss (valid steps) and DuoRec (last step); view 2 only DuoRec: the late-fusion
    # head runs on exactly those rows (same values as slicing the full [B,L,128] output)
    out1, out2 = _two_views(model, batch, pretrained_vecs, kw, loss_scope, packed)
    u = encoder.l2_normalize(out1[:n_main])                                      # :794-807
    tgt = tgt_flat[idx]
    uid = idx // L                                                               # batch row = user id (:801-804)
    if columns == "batch" or (columns == "unique" and ("col_item_ids" not in batch or loss_scope != "all")):
        v = item_tower.normalized_rows(tgt)                                      # :810-811 + :833
        main = losses.logq_infonce_rows(u, v, tgt, uid, item_tower.get_log_q(), 0.1, lambda_logq, unit_norm=True)
    else:
        if columns == "unique":
            cid, cnt, pos_col, grid = (batch[k] for k in ("col_item_ids", "col_counts", "pos_col", "own_grid"))
            v = item_tower.normalized_rows(cid)
        else:
            v = F.normalize(item_tower.get_all_embeddings(), p=2, dim=1)         # :810, all rows
            cid, cnt, pos_col = losses.item_columns(tgt, v.shape[0])
            grid = batch["target_ids"].masked_fill(batch["padding_mask"], -1)
        # same-user mask: the rows are grouped by user (batch-major valid steps) -> one dense block per user
        blk = {}
        if loss_scope == "all" and "cu_seqlens" in batch and L <= 64:
            blk, own = dict(row_cu=batch["cu_seqlens"][:B + 1], max_rows_per_user=L), None
        else:
            own = grid[uid] if loss_scope == "all" else None     # one row per user otherwise: nothing to mask
        main = losses.logq_infonce_columns(u, v, cid, cnt, tgt, pos_col, own, item_tower.get_log_q(), 0.1,
                                           lambda_logq, unit_norm=True, **blk)     # u and v are F.normalize'd
    cl = losses.duorec_loss_refined(out1[n_main:], out2, tgt_flat[li], lambda_sup=lambda_sup)   # :830-842
    return main, cl


def two_tower_step(model, item_tower, batch, pretrained_lookup, optimizer=None, lambda_logq=1.0, lambda_sup=0.1,
                   lambda_cl=0.2, loss_scope="all", amp_dtype: Optional[torch.dtype] = torch.bfloat16,
                   scaler=None, max_norm=5.0, grad_hook=None, sdpa_efficient=True, columns="unique", packed=True):
    """forward x2 (two dropout views) + C2 + C3 + backward + clip + optimizer step.
    Returns (total, main, cl) as device scalars.  `batch` comes from prepare_batch(add_host_index(...)).
    columns: how the in-batch softmax of the main loss enumerates its columns (same loss value, see
    losses.logq_infonce_columns) -- "unique": the distinct target items of the batch with multiplicities
    (default); "catalog": every item, static shape; "batch": one column per row, the reference's [N, N].
    packed: run the sequence encoder on the packed valid tokens (encoder.py) instead of the padded [B, L] grid
    (needs the batch's `cu_seqlens`, i.e. no empty sequence); same values at the positions the losses read."""
    item_ids = batch["item_ids"]
    B, L = item_ids.shape
    encoder.rng_advance()            # new dropout epoch (also what makes a captured graph draw new masks per replay)
    encoder.prepare_weights(model, amp_dtype)      # 16-bit copies of the GEMM weights: one multi-tensor launch
    if optimizer is not None:
        optimizer.zero_grad(set_to_none=True)
    with torch.no_grad():
        pretrained_vecs = ops.gather_rows(pretrained_lookup, _front_item_ids(batch, packed), out_dtype=amp_dtype)  # (item_proj's autocast cast, folded in)
    kw = {k: batch[k] for k in FORWARD_KEYS}
    # The encoder is stock nn.TransformerEncoder (as in the reference).  For its shape (L=50, 4 heads x 32, explicit
    # causal + padding mask) PyTorch's default pick on sm_100, the cuDNN flash kernel with 128-wide tiles, is 28 %
    # slower than the memory-efficient backend (tools/sdpa_probe.py: 28.4 vs 22.2 ms per view, fwd+bwd): select it.
    sdpa = sdpa_kernel([SDPBackend.EFFICIENT_ATTENTION, SDPBackend.MATH]) if sdpa_efficient else contextlib.nullcontext()
    bucketed = packed and "row_weight" in batch
    if bucketed and (loss_scope != "all" or columns != "unique"):
        raise ValueError("the device-built batch index serves loss_scope='all' with columns='unique'")
    with sdpa, torch.autocast("cuda", dtype=amp_dtype, enabled=amp_dtype is not None):
        fn = _losses_device_index if bucketed else _losses_host_index
        main, cl = fn(model, item_tower, batch, pretrained_vecs, kw, loss_scope, packed, columns, lambda_logq, lambda_sup)
        total = main + lambda_cl * cl
    if optimizer is not None:
        if scaler is not None:
            scaler.scale(total).backward()
            if grad_hook is not None:
                grad_hook()                      # e.g. the data-parallel gradient all-reduce
            scaler.unscale_(optimizer)
            torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=max_norm)
            scaler.step(optimizer)
            scaler.update()
        else:
            total.backward()
            if grad_hook is not None:
                grad_hook()
            torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=max_norm)
            optimizer.step()
    encoder.release_weights()
    return total.detach(), main.detach(), cl.detach()


# ------------------------------------------------------------------------------------------------------------
# N > 1: row-sharded item tables + catalogue-wide negatives.  One process per GPU.
# ------------------------------------------------------------------------------------------------------------
class ShardedTwoTower:
    """The same step on `world` ranks with
      * `item_id_emb` (user tower) and `item_matrix` (item tower) ROW-SHARDED, owner = row % world: each rank keeps
        [ceil(rows/world), 128] of each (parameters + AdamW state); lookups go through the planned all-to-all
        (sharded.planned_lookup: owner-side gather kernel -> rows all-to-all; backward: gradient rows all-to-all ->
        owner-side deterministic segment reduce);
      * the negatives of the main loss spanning the box: every rank scores its rows against the DISTINCT target
        items of all ranks (their rows fetched from the owners by all-to-all; or, columns="catalog", against the
        all-gathered item matrix), each item weighted by its number of occurrences among the targets of ALL ranks
        -- the reference's in-batch softmax over the global batch, grouped by item (losses.logq_infonce_columns);
        the gradient of the fetched rows travels back to the owners (all-to-all / reduce-scatter);
      * DuoRec with all-gathered second views / targets (rectangular [B, world*B] blocks, diag_offset = rank*B);
      * every other parameter replicated, gradients summed with one flat all-reduce.
    Each rank's loss is its share of the GLOBAL mean (local sum / global row count), so that summing the
    gradients over ranks -- which the collectives above do -- gives the gradient of the global-batch loss."""

    def __init__(self, model, item_tower, group=None):
        import torch.distributed as dist
        from . import sharded
        self.dist, self.sh = dist, sharded
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.model, self.item_tower = model, item_tower
        dev = item_tower.item_matrix.weight.device
        n_rows = item_tower.item_matrix.weight.shape[0]
        self.n_rows = n_rows
        self.cols = sharded.CatalogColumns(n_rows, self.world, dev)
        for emb in (model.item_id_emb, item_tower.item_matrix):                 # keep only this rank's rows
            full = emb.weight.data
            emb.weight = torch.nn.Parameter(sharded.shard_padded(full, self.rank, self.world))
        lq = item_tower.get_log_q().float()
        lq_pad = torch.zeros(self.cols.n_cols, device=dev)
        lq_pad[:n_rows] = lq
        self.log_q_by_id = lq_pad                                                # indexed by item id (padded)
        self.sharded_params = [model.item_id_emb.weight, item_tower.item_matrix.weight]
        sp = {id(p) for p in self.sharded_params}
        self.replicated_params = [p for p in list(model.parameters()) + list(item_tower.parameters())
                                  if id(p) not in sp and p.requires_grad]
        self.pad_local_row = 0 if self.rank == 0 else -1                         # global row 0 = rank 0, local row 0

    def plan(self, batch, columns="unique"):
        """Loader-stage work for one (device) batch -- everything that depends only on its ids, done with
        collectives and host reads BEFORE the step so that the step itself never synchronises:
          * routing of the item ids against the row-sharded `item_id_emb` (sharded.plan_lookup);
          * columns="unique": the distinct target items of ALL ranks (sorted), their box-wide counts, the routing of
            that list against the row-sharded `item_matrix`, and the column of every local target;
            columns="catalog": nothing (every item is a column, static; counts are all-reduced inside the step)."""
        dist, sh = self.dist, self.sh
        batch = dict(batch)
        batch["lookup_plan"] = sh.plan_lookup(_front_item_ids(batch, True), self.group)
        if columns == "unique":
            tgt = batch["target_ids"].reshape(-1)[batch["valid_index"]]
            mine = torch.unique(tgt)
            n = torch.tensor([mine.numel()], device=tgt.device)
            sizes = [torch.zeros_like(n) for _ in range(self.world)]
            dist.all_gather(sizes, n, group=self.group)
            cap = int(max(int(x) for x in sizes))
            padded = torch.full((cap,), -1, dtype=mine.dtype, device=tgt.device)
            padded[:mine.numel()] = mine
            allv = torch.empty(self.world * cap, dtype=mine.dtype, device=tgt.device)
            dist.all_gather_into_tensor(allv, padded, group=self.group)
            gq = torch.unique(allv[allv >= 0])                               # sorted, identical on every rank
            pos_col = torch.searchsorted(gq, tgt)
            cnt = losses.count_ids(pos_col, gq.numel())
            dist.all_reduce(cnt, group=self.group)
            batch.update(col_item_ids=gq, col_counts=cnt, pos_col=pos_col,
                         col_plan=sh.plan_lookup(gq, self.group))
        return batch

    def _sync_replicated(self):
        gs = [p.grad for p in self.replicated_params if p.grad is not None]
        flat = torch.cat([g.reshape(-1) for g in gs])
        self.dist.all_reduce(flat, group=self.group)
        torch._foreach_copy_(gs, [t.view_as(g) for t, g in zip(flat.split([g.numel() for g in gs]), gs)])

    def _clip(self, max_norm):
        """clip_grad_norm_ over the user tower's parameters with its sharded table counted once, globally."""
        mp = {id(p) for p in self.model.parameters()}
        rep = [p.grad for p in self.replicated_params if id(p) in mp and p.grad is not None]
        sq_rep = torch.stack(torch._foreach_norm(rep)).square().sum()
        sq_sh = self.model.item_id_emb.weight.grad.float().square().sum()
        self.dist.all_reduce(sq_sh, group=self.group)
        total = (sq_rep + sq_sh).sqrt()
        coef = (max_norm / (total + 1e-6)).clamp(max=1.0)
        torch._foreach_mul_(rep + [self.model.item_id_emb.weight.grad], coef)
        return total

    def step(self, batch, pretrained_lookup, optimizer=None, lambda_logq=1.0, lambda_sup=0.1, lambda_cl=0.2,
             amp_dtype: Optional[torch.dtype] = torch.bfloat16, max_norm=5.0, sdpa_efficient=True, packed=True):
        dist, sh, model, item_tower = self.dist, self.sh, self.model, self.item_tower
        item_ids = batch["item_ids"]
        B, L = item_ids.shape
        encoder.rng_advance()
        encoder.prepare_weights(model, amp_dtype)
        if optimizer is not None:
            optimizer.zero_grad(set_to_none=True)
        with torch.no_grad():
            pretrained_vecs = ops.gather_rows(pretrained_lookup, _front_item_ids(batch, packed), out_dtype=amp_dtype)  # (item_proj's autocast cast, folded in)
        kw = {k: batch[k] for k in FORWARD_KEYS}
        # one exchange serves both dropout views (their gradients add up in the buffer before travelling back)
        id_rows = sh.planned_lookup(model.item_id_emb.weight, batch["lookup_plan"], self.group, lead_rows=1,
                                    pad_local_row=self.pad_local_row)
        sdpa = sdpa_kernel([SDPBackend.EFFICIENT_ATTENTION, SDPBackend.MATH]) if sdpa_efficient else contextlib.nullcontext()
        with sdpa, torch.autocast("cuda", dtype=amp_dtype, enabled=amp_dtype is not None):
            tgt_flat = batch["target_ids"].reshape(-1)
            idx, li = batch["valid_index"], batch["last_index"]
            n_main = idx.numel()
            out1, out2 = _two_views(model, batch, pretrained_vecs, kw, "all", packed, item_id_rows=id_rows)
            u = encoder.l2_normalize(out1[:n_main])
            tgt = tgt_flat[idx]
            uid = idx // L
            # main loss: all items as columns, global multiplicities
            n_glob = torch.full((1,), float(n_main), device=u.device)          # (a fill kernel: graph-capturable)
            dist.all_reduce(n_glob, group=self.group)
            if "col_plan" in batch:
                # columns = the distinct targets of the whole box: fetch their rows from the owners (all-to-all), normalise
                rows = sh.planned_lookup(item_tower.item_matrix.weight, batch["col_plan"], self.group)
                v_cols = encoder.l2_normalize(rows)
                cid, cnt, pos_col = batch["col_item_ids"], batch["col_counts"], batch["pos_col"]
            else:
                # columns = every item: all-gather the normalised shards (reduce-scatter backward), static shapes
                v_cols = sh.all_gather_rows(F.normalize(item_tower.item_matrix.weight, p=2, dim=1), self.group)
                cid, cnt, pos_col = self.cols.col_item_ids, self.cols.counts(tgt, self.group), self.cols.col_of(tgt)
            if "cu_seqlens" in batch and L <= 64:
                blk, own = dict(row_cu=batch["cu_seqlens"][:B + 1], max_rows_per_user=L), None
            elif "col_plan" in batch:
                g_ = torch.full((B * L,), -1, dtype=torch.int64, device=u.device).index_copy_(0, idx, pos_col)
                blk, own = {}, g_.view(B, L)[uid]
            else:
                blk, own = {}, self.cols.col_of(batch["target_ids"].masked_fill(batch["padding_mask"], -1))[uid]
            main_local = losses.logq_infonce_columns(u, v_cols, cid, cnt, tgt, pos_col, own, self.log_q_by_id, 0.1,
                                                     lambda_logq, unit_norm=True, **blk)
            main = main_local * (n_main / n_glob.squeeze(0))
            # DuoRec across the box
            cl = self._duorec(out1[n_main:], out2, tgt_flat[li], lambda_sup) / self.world
            total = main + lambda_cl * cl
        if optimizer is not None:
            total.backward()
            self._sync_replicated()
            self._clip(max_norm)
            optimizer.step()
        out = torch.stack([total.detach(), main.detach(), cl.detach()])
        dist.all_reduce(out, group=self.group)                                   # global-batch losses (for logging)
        encoder.release_weights()
        return out[0], out[1], out[2]

    def _duorec(self, e1, e2, tgt, lambda_sup, temperature=0.1):
        """C3 (v1_refine_usertower.py:576-627) with the columns of every rank; returns this rank's mean over its
        B rows (the caller divides by world: equal B on all ranks)."""
        sh, rank = self.sh, self.rank
        B = e1.shape[0]
        z1, z2 = encoder.l2_normalize(e1), encoder.l2_normalize(e2)
        z2_all = sh.all_gather_rows(z2, self.group)
        loss = losses.info_nce(z1, z2_all, temperature, diag_offset=rank * B, unit_norm=True)
        if lambda_sup > 0:
            z1_all = sh.all_gather_rows(z1, self.group)
            tgt_all = sh.all_gather_ids(tgt, self.group)
            lse, _, pos_sum, pos_cnt = losses.fused_softmax_stats(
                z1, z1_all, 1.0 / temperature, key_a_row=tgt, key_a_col=tgt_all, diag_offset=rank * B,
                mask_value=losses.NEG_INF, flags=losses.L.RS_CE_DIAG_MASK | losses.L.RS_CE_SUPCON, unit_norm=True)
            valid = pos_cnt > 0
            per_row = torch.where(valid, lse - pos_sum / pos_cnt.clamp(min=1.0), torch.zeros_like(lse))
            n_valid = valid.sum().to(per_row.dtype)
            self.dist.all_reduce(n_valid, group=self.group)
            # mean over the valid rows of ALL ranks; x world because the caller averages the rank means
            loss = loss + lambda_sup * per_row.sum() / n_valid.clamp(min=1) * self.world
        return loss


class ShardedDeviceStep(ShardedTwoTower):
    """The N > 1 step on device-indexed batches (device_index): same sharding and the same loss as ShardedTwoTower, but
    every routing decision is made on the device inside the step, with static shapes:

      * U1's item rows: histogram of the batch's item ids -> owner-major request list of `front_cap` slots per owner
        (ops.owner_compact) -> DE-DUPLICATED equal-split exchange (sharded.dedup_lookup: ids all-to-all, owner-side
        gather kernel, rows all-to-all) -> the front kernel reads row `slot` of the arrival buffer for every token.
        ~20 k distinct rows travel instead of ~110 k token rows; the backward reduces token gradients into the
        buffer locally (the front's own sparse backward), then ONE gradient row per (rank, item) travels back;
      * negatives spanning the box: the ranks' target histograms are all-reduced (int32 [n_items], 0.4 MB), every rank
        derives the identical owner-major list of distinct targets (`col_cap` per owner) from it, normalises ITS OWN
        segment's rows from its shard and all-gathers the segments (reduce-scatter backward): no id lists exchanged,
        no unique / searchsorted, no host read;
      * capacities are fixed at start-up from a few batches (`calibrate`), overflow raises a device flag (`check`).
    Nothing in the step synchronises with the host, and all collective sizes are static."""

    def __init__(self, model, item_tower, group=None):
        super().__init__(model, item_tower, group)
        self.R = self.sh.padded_rows(self.n_rows, self.world)
        self.front_cap = self.col_cap = None
        # overflow flags of every step since the last check(), OR-ed into ONE persistent word (a captured graph keeps
        # writing to it on every replay; tensors created inside a capture live in the graphs' shared pool and are reused)
        self.flag_acc = torch.zeros(4, dtype=torch.int32, device=item_tower.item_matrix.weight.device)

    # -- routing (device, stream-ordered)
    def _front_route(self, batch, cap):
        cnt = ops.id_histogram(batch["pk_item_ids"], self.n_rows, force_bin0=True)
        req, _, _, slot_of, meta = ops.owner_compact(cnt, self.world, self.R, cap)
        return req, ops.lookup_i32(slot_of, batch["pk_item_ids"], 0), meta

    def _col_route(self, batch, cap):
        cnt = ops.id_histogram(batch["main_tgt"], self.n_rows, n_valid=batch["meta"][0:1])
        if self.world > 1:
            self.dist.all_reduce(cnt, group=self.group)                  # occurrences among the targets of ALL ranks
        rows, ids, counts, slot_of, meta = ops.owner_compact(cnt, self.world, self.R, cap, want_ids=True)
        return rows, ids, counts, ops.lookup_i32(slot_of, batch["main_tgt"], 0), meta

    def calibrate(self, batches, margin: float = 1.15, col_margin: float = 1.03, q: int = 128):
        """Fix the per-owner capacities from a few (device-indexed) batches of every rank: largest list seen anywhere,
        plus a margin (request slots only cost exchange volume: generous; column slots cost tensor-core work: the
        box-wide distinct-target count varies by well under 1 % between batches, 3 % is plenty).  Collective + host
        read: start-up only."""
        mf = mc = 0
        for b in batches:
            mf = max(mf, int(self._front_route(b, self.R)[2][0]))
            mc = max(mc, int(self._col_route(b, self.R)[4][0]))
        t = torch.tensor([mf, mc], device=batches[0]["main_tgt"].device)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        mf, mc = (int(x) for x in t.tolist())
        self.front_cap = min(self.R, ops.round_up(int(mf * margin) + 1, q))
        self.col_cap = min(self.R, ops.round_up(int(mc * col_margin) + 1, q))
        return self.front_cap, self.col_cap

    def check(self):
        """Raise if a request list overflowed its capacity in any step since the last check (synchronises)."""
        bad = int(self.flag_acc[1])
        self.flag_acc.zero_()
        if bad:
            raise RuntimeError("sharded step: a per-owner request list exceeded its capacity; re-run calibrate()")

    def step(self, batch, pretrained_lookup, optimizer=None, lambda_logq=1.0, lambda_sup=0.1, lambda_cl=0.2,
             amp_dtype: Optional[torch.dtype] = torch.bfloat16, max_norm=5.0, sdpa_efficient=True, packed=True):
        if "row_weight" not in batch:
            raise ValueError("ShardedDeviceStep.step needs a device-indexed batch (train.device_index)")
        if self.front_cap is None:
            raise RuntimeError("call calibrate() first")
        dist, sh, model, item_tower = self.dist, self.sh, self.model, self.item_tower
        B, L = batch["item_ids"].shape
        encoder.rng_advance()
        encoder.prepare_weights(model, amp_dtype)
        if optimizer is not None:
            optimizer.zero_grad(set_to_none=True)
        with torch.no_grad():
            pretrained_vecs = ops.gather_rows(pretrained_lookup, batch["pk_item_ids"], out_dtype=amp_dtype)
        kw = {k: batch[k] for k in FORWARD_KEYS}
        # one de-duplicated exchange serves every token of both dropout views
        req, slots, meta_f = self._front_route(batch, self.front_cap)
        buf = sh.dedup_lookup(model.item_id_emb.weight, req, self.group, pad_local_row=self.pad_local_row)
        sdpa = sdpa_kernel([SDPBackend.EFFICIENT_ATTENTION, SDPBackend.MATH]) if sdpa_efficient else contextlib.nullcontext()
        with sdpa, torch.autocast("cuda", dtype=amp_dtype, enabled=amp_dtype is not None):
            tgt = batch["main_tgt"]
            n_main = tgt.numel()
            out_main, last1, last2 = _two_views(model, batch, pretrained_vecs, kw, "all", packed, item_id_rows=(buf, slots))
            u = encoder.l2_normalize(out_main)
            # columns: the distinct targets of the whole box, owner-major; this rank contributes its own segment
            rows, cid, cnt, pos_col, meta_c = self._col_route(batch, self.col_cap)
            own = rows[self.rank * self.col_cap:(self.rank + 1) * self.col_cap]
            v_cols = sh.all_gather_rows(ops.normalized_rows(item_tower.item_matrix.weight, own), self.group)
            t_loc = batch["meta"][0:1].to(torch.float32)
            t_glob = t_loc.clone()
            dist.all_reduce(t_glob, group=self.group)
            rw = batch["row_weight"] * (t_loc / t_glob)                           # 1 / (valid rows of ALL ranks)
            main = losses.logq_infonce_columns(u, v_cols, cid, cnt, tgt, pos_col, None, self.log_q_by_id, 0.1,
                                               lambda_logq, unit_norm=True, row_cu=batch["row_cu"],
                                               max_rows_per_user=L, row_weight=rw)
            cl = self._duorec(last1, last2, batch["last_tgt"], lambda_sup) / self.world
            total = main + lambda_cl * cl
        torch.maximum(self.flag_acc, torch.maximum(meta_f, meta_c), out=self.flag_acc)
        if optimizer is not None:
            total.backward()
            self._sync_replicated()
            self._clip(max_norm)
            optimizer.step()
        out = torch.stack([total.detach(), main.detach(), cl.detach()])
        dist.all_reduce(out, group=self.group)                                   # global-batch losses (for logging)
        encoder.release_weights()
        return out[0], out[1], out[2]

    def full_state_dict(self):
        """`state_dict()` of both towers with the row-sharded tables gathered back to full size, under the reference's
        parameter names (checkpoints are saved / loaded by name,
        tower_code/v1_usertower_train.py:1020-1021,1079-1083).  Collective: call it on every rank."""
        return gathered_state_dicts(self)


def gathered_state_dicts(trainer) -> Dict[str, Dict[str, torch.Tensor]]:
    """{'user_tower': ..., 'item_tower': ...}: full-size state dicts of a (row-sharded) trainer, every rank gets them."""
    dist, sh = trainer.dist, trainer.sh
    out = {}
    for name, mod, key in (("user_tower", trainer.model, "item_id_emb.weight"), ("item_tower", trainer.item_tower, "item_matrix.weight")):
        sd = {k: v.detach().clone() for k, v in mod.state_dict().items()}
        shard = sd[key].contiguous()
        parts = [torch.empty_like(shard) for _ in range(trainer.world)]
        dist.all_gather(parts, shard, group=trainer.group)
        sd[key] = sh.unshard_rows(parts)[:trainer.n_rows].contiguous()        # drop the equal-shard padding rows
        out[name] = sd
    return out


# the tensors of one collated batch (SASRecDataset's default collate, tower_code/v1_refine_usertower.py:204-306)
_GRID_I64 = ("item_ids", "target_ids", "time_bucket_ids", "type_ids", "color_ids", "graphic_ids", "section_ids")
_USER_I64 = ("age_bucket", "price_bucket", "cnt_bucket", "recency_bucket", "channel_ids", "club_status_ids",
             "news_freq_ids", "fn_ids", "active_ids")


class FlatBatch:
    """One collated batch as views of ONE byte buffer (pinned host memory or device memory): a batch crosses PCIe, or
    moves between a staging buffer and a graph's static inputs, with a single copy."""

    def __init__(self, B: int, L: int, device=None, pin: bool = False):
        spec = [(k, (B, L), torch.int64) for k in _GRID_I64] + [(k, (B,), torch.int64) for k in _USER_I64]
        spec += [("cont_feats", (B, 4), torch.float32), ("padding_mask", (B, L), torch.bool)]
        offs, o = [], 0
        for _, shape, dt in spec:
            offs.append(o)
            n = torch.tensor([], dtype=dt).element_size()
            for d in shape:
                n *= d
            o += (n + 255) // 256 * 256
        self.B, self.L, self.nbytes = B, L, o
        self.buf = torch.empty(o, dtype=torch.uint8, device=device if device is not None else "cpu",
                               pin_memory=bool(pin and device is None))
        self.views: Dict[str, torch.Tensor] = {}
        for (k, shape, dt), off in zip(spec, offs):
            n = torch.tensor([], dtype=dt).element_size()
            for d in shape:
                n *= d
            self.views[k] = self.buf[off:off + n].view(dt).view(*shape)

    def fill(self, batch: Dict[str, torch.Tensor]) -> "FlatBatch":
        for k, v in self.views.items():
            v.copy_(batch[k])
        return self

    def copy_(self, other: "FlatBatch", non_blocking: bool = True) -> "FlatBatch":
        self.buf.copy_(other.buf, non_blocking=non_blocking)
        return self


class BucketedStep:
    """The train step on device-indexed batches, one captured CUDA graph per SHAPE BUCKET (tok_cap, col_cap).

    Every graph reads the same static raw-batch buffer (`self.raw`, a FlatBatch on the device), builds the batch index
    inside the graph (ops.batch_index_build: static output shapes) and runs forward, losses, backward, clip and the
    optimizer.  A loader only needs two numbers per batch to pick the graph -- valid time steps T and distinct targets
    U (`counts`, computed on the device while the previous step runs).  step_fn(indexed_batch) -> (total, main, cl)."""

    def __init__(self, step_fn, B: int, L: int, n_item_rows: int, device, use_graph: bool = True,
                 tok_q: int = TOK_BUCKET, col_q: int = COL_BUCKET):
        self.step_fn, self.B, self.L, self.n_item_rows, self.device = step_fn, B, L, n_item_rows, device
        self.use_graph, self.tok_q, self.col_q = use_graph, tok_q, col_q
        self.raw = FlatBatch(B, L, device=device)
        self.index: Dict[tuple, Dict[str, torch.Tensor]] = {}
        self.graphs: Dict[tuple, "GraphedStep"] = {}
        self.pool = None

    def counts(self, fb: FlatBatch, meta: Optional[torch.Tensor] = None) -> torch.Tensor:
        """int32[8] device tensor (T, E, U, ...) of the batch in `fb` -- stream-ordered, no synchronisation."""
        return ops.batch_index_counts(fb.views["padding_mask"], fb.views["target_ids"], self.n_item_rows, meta)

    def bucket(self, n_tokens: int, n_cols: int):
        """`n_cols` None: the step does not use the index's LOCAL column list (the sharded step derives box-wide
        columns itself): give it the capacity that can never overflow."""
        if n_cols is None:
            t = ops.round_up(max(n_tokens, 1), self.tok_q)
            return t, t
        return bucket_of(n_tokens, n_cols, self.tok_q, self.col_q)

    def _run_eager(self, key):
        idx = self.index.get(key)
        if idx is None:
            idx = self.index[key] = ops.batch_index_alloc(self.B, self.L, key[0], key[1], self.device)
        return self.step_fn(device_index(self.raw.views, self.n_item_rows, key[0], key[1], out=idx))

    def ensure(self, key) -> bool:
        """Capture the graph of bucket `key` if it does not exist yet; `self.raw` must hold a batch of that bucket.
        Returns True when a capture happened."""
        if not self.use_graph or key in self.graphs:
            return False
        g = GraphedStep(lambda _views: self._run_eager(key), self.raw.views, pool=self.pool)
        if self.pool is None:
            self.pool = g.graph.pool()         # later graphs share this one's memory pool (they never run concurrently)
        self.graphs[key] = g
        return True

    def run(self, key):
        """One step on the batch in `self.raw` (which must belong to bucket `key`)."""
        if not self.use_graph:
            return self._run_eager(key)
        self.ensure(key)
        return self.graphs[key].replay()

    def meta_of(self, key) -> torch.Tensor:
        """the index's counters / overflow flags of the last step run in bucket `key` (device int32[8])"""
        return self.index[key]["meta"]


class GraphedStep:
    """One train step captured in a CUDA graph (whole step: forward, losses, backward, clip, optimizer), replayed with
    a single launch.  The step is ~600 small kernels; eager Python enqueues them about as fast as the GPU retires
    them, so one graph launch per step takes the host off the critical path.

    Shapes inside the step depend on the batch (valid tokens, distinct items), so a graph belongs to ONE static batch
    (a dict of device tensors): refill those tensors in place (`load`) and `replay`.  A loader that buckets batches by
    shape keeps one GraphedStep per bucket.  Dropout: the kernels' seeds are frozen at capture, the device-side epoch
    advanced inside the step (rs_rng_advance) makes every replay draw new masks; torch's own dropout is graph-safe.
    The optimizer must be capturable (`torch.optim.AdamW(..., fused=True, capturable=True)`)."""

    def __init__(self, step_fn, static_batch, warmup: int = 3, pool=None):
        self.batch = static_batch
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):                  # allocator, workspaces, sort caches, autotuned GEMMs: all warm
                step_fn(static_batch)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        # the (id, position) sorts of the sparse backward are cached per ids tensor: drop the warm-up's entries so that
        # the sorts are captured INSIDE the graph (a replay after `load` must sort the new ids), and drop the captured
        # entries afterwards (they live in the graph's private pool)
        ops._sort_cache.clear()
        lib = ops._lib
        c0 = lib.rs_launch_count()
        self.graph = torch.cuda.CUDAGraph()
        # with torch.distributed initialised NCCL's watchdog thread polls events meanwhile, which a "global" capture
        # would reject (capturing the N > 1 step itself is not supported: see DESIGN.md 7)
        mode = "thread_local" if (torch.distributed.is_available() and torch.distributed.is_initialized()) else "global"
        with torch.cuda.graph(self.graph, pool=pool, capture_error_mode=mode):
            self.out = step_fn(static_batch)
        self.launches = int(lib.rs_launch_count() - c0)      # this library's kernel nodes in the graph
        ops._sort_cache.clear()

    def load(self, host_batch, non_blocking=True):
        """copy a host batch with the SAME shapes into the static inputs (H2D, stream-ordered before the replay)"""
        for k, v in host_batch.items():
            dst = self.batch.get(k)
            if isinstance(dst, torch.Tensor) and isinstance(v, torch.Tensor):
                if dst.shape != v.shape:
                    raise ValueError(f"GraphedStep.load: '{k}' has shape {tuple(v.shape)}, the graph was captured with "
                                     f"{tuple(dst.shape)}")
                dst.copy_(v, non_blocking=non_blocking)
            elif dst is not None and not isinstance(dst, torch.Tensor):
                raise TypeError(f"GraphedStep.load: '{k}' is not a tensor (routing plans of the sharded step hold host "
                                f"split sizes that a captured graph has frozen)")

    def replay(self):
        self.graph.replay()
        return self.out
