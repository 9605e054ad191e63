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

from . import losses, ops
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
    if columns:
        tgt = batch["target_ids"].reshape(-1)[batch["valid_index"]]
        ids, counts, pos_col = losses.item_columns(tgt)
        grid = torch.full((B * L,), -1, dtype=torch.int64)
        grid[batch["valid_index"]] = pos_col
        batch.update(col_item_ids=ids, col_counts=counts, pos_col=pos_col, own_grid=grid.view(B, L))
    return batch


def two_tower_step(model, item_tower, batch, pretrained_lookup, optimizer=None, lambda_logq=1.0, lambda_sup=0.1,
                   lambda_cl=0.2, loss_scope="all", amp_dtype: Optional[torch.dtype] = torch.bfloat16,
                   scaler=None, max_norm=5.0, grad_hook=None, sdpa_efficient=True, columns="unique"):
    """forward x2 (two dropout views) + C2 + C3 + backward + clip + optimizer step.
    Returns (total, main, cl) as device scalars.  `batch` comes from prepare_batch(add_host_index(...)).
    columns: how the in-batch softmax of the main loss enumerates its columns (same loss value, see
    losses.logq_infonce_columns) -- "unique": the distinct target items of the batch with multiplicities
    (default); "catalog": every item, static shape; "batch": one column per row, the reference's [N, N]."""
    item_ids = batch["item_ids"]
    B, L = item_ids.shape
    if optimizer is not None:
        optimizer.zero_grad(set_to_none=True)
    with torch.no_grad():
        pretrained_vecs = ops.gather_rows(pretrained_lookup, item_ids)
    kw = {k: batch[k] for k in FORWARD_KEYS}
    # The encoder is stock nn.TransformerEncoder (as in the reference).  For its shape (L=50, 4 heads x 32, explicit
    # causal + padding mask) PyTorch's default pick on sm_100, the cuDNN flash kernel with 128-wide tiles, is 28 %
    # slower than the memory-efficient backend (tools/sdpa_probe.py: 28.4 vs 22.2 ms per view, fwd+bwd): select it.
    sdpa = sdpa_kernel([SDPBackend.EFFICIENT_ATTENTION, SDPBackend.MATH]) if sdpa_efficient else contextlib.nullcontext()
    with sdpa, torch.autocast("cuda", dtype=amp_dtype, enabled=amp_dtype is not None):
        tgt_flat = batch["target_ids"].reshape(-1)
        idx = batch["valid_index"] if loss_scope == "all" else batch["last_index"]
        li = batch["last_index"]
        n_main = idx.numel()
        # view 1 feeds the main loss (valid steps) and DuoRec (last step); view 2 only DuoRec: the late-fusion
        # head runs on exactly those rows (same values as slicing the full [B,L,128] output)
        sel1 = batch.get("select_index")
        if sel1 is None:
            sel1 = torch.cat([idx, li])
        out1 = model(pretrained_vecs=pretrained_vecs, **kw, training_mode=True, select_index=sel1)
        out2 = model(pretrained_vecs=pretrained_vecs, **kw, training_mode=True, select_index=li)
        u = F.normalize(out1[:n_main], p=2, dim=1)                                   # :794-807
        tgt = tgt_flat[idx]
        uid = idx // L                                                               # batch row = user id (:801-804)
        if columns == "batch" or (columns == "unique" and ("col_item_ids" not in batch or loss_scope != "all")):
            v = item_tower.normalized_rows(tgt)                                      # :810-811 + :833
            main = losses.logq_infonce_rows(u, v, tgt, uid, item_tower.get_log_q(), 0.1, lambda_logq)
        else:
            if columns == "unique":
                cid, cnt, pos_col, grid = (batch[k] for k in ("col_item_ids", "col_counts", "pos_col", "own_grid"))
                v = item_tower.normalized_rows(cid)
            else:
                v = F.normalize(item_tower.get_all_embeddings(), p=2, dim=1)         # :810, all rows
                cid, cnt, pos_col = losses.item_columns(tgt, v.shape[0])
                grid = batch["target_ids"].masked_fill(batch["padding_mask"], -1)
            own = grid[uid] if loss_scope == "all" else None     # one row per user otherwise: nothing to mask
            main = losses.logq_infonce_columns(u, v, cid, cnt, tgt, pos_col, own, item_tower.get_log_q(), 0.1,
                                               lambda_logq)
        cl = losses.duorec_loss_refined(out1[n_main:], out2, tgt_flat[li], lambda_sup=lambda_sup)   # :830-842
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
    return total.detach(), main.detach(), cl.detach()
