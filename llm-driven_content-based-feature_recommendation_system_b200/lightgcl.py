"""N3: LightGCL's loss terms and retrieval on the path's kernels.

    calc_bpr_loss   gnn_model/v1_lightgcl.py:188-195
    calc_ssl_loss   gnn_model/v1_lightgcl.py:197-213   InfoNCE between the local and the SVD view (temp 0.2, clamp 100)
    get_l2_reg      gnn_model/v1_lightgcl.py:215-219
    retrieval       gnn_model/v1_evaluate_lightgcl.py:312-318   pure dot product, item 0 masked, top-k

The graph propagation (SpMM) is out of scope; these are the same contraction / gather ops as the two-tower path:
row gathers with dense scatter-add backward (rs_gather_rows), the fused tcgen05 softmax (rs_ce_*; LightGCL's
64-wide rows are zero-padded to the kernel's K = 128, which leaves every dot product unchanged) and the fused top-k."""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import Tensor

from . import losses, ops


def calc_bpr_loss(local_emb: Tensor, users: Tensor, pos_items: Tensor, neg_items: Tensor) -> Tensor:
    u = ops.gather_rows(local_emb, users)
    p = ops.gather_rows(local_emb, pos_items)
    n = ops.gather_rows(local_emb, neg_items)
    pos_scores = torch.sum(u * p, dim=1)
    neg_scores = torch.sum(u * n, dim=1)
    return -torch.mean(torch.log(torch.sigmoid(pos_scores - neg_scores) + 1e-10))


def _unit_rows_128(table: Tensor, idx: Tensor) -> Tensor:
    """F.normalize(table, dim=1)[idx] (only the rows asked for are normalised: same values) zero-padded to 128 columns."""
    rows = F.normalize(ops.gather_rows(table, idx).float(), dim=1)
    d = rows.shape[1]
    if d > 128 or d % 8:
        raise ValueError(f"embedding width {d}: the fused softmax contracts over K = 128 (widths up to 128, multiple of 8)")
    return F.pad(rows, (0, 128 - d)) if d < 128 else rows


def calc_ssl_loss(local_emb: Tensor, global_emb: Tensor, users: Tensor, items: Tensor, temp: float = 0.2) -> Tensor:
    """`torch.clamp(logits, max=100)` (:207) never binds for unit rows when 1/temp <= 100, i.e. temp >= 0.01 (LightGCL
    trains at 0.2); a smaller temperature would need the clamp inside the kernel and is rejected."""
    if 1.0 / temp * losses.UNIT_NORM_BOUND > 100.0:
        raise NotImplementedError(f"temp = {temp}: the reference's clamp(max=100) on the logits would bind; not supported")

    def nce(idx):
        v1, v2 = _unit_rows_128(local_emb, idx), _unit_rows_128(global_emb, idx)
        return losses.info_nce(v1, v2, temp, unit_norm=True)
    return nce(torch.unique(users)) + nce(torch.unique(items))


def get_l2_reg(user_weight: Tensor, item_weight: Tensor, users: Tensor, pos_items: Tensor, neg_items: Tensor) -> Tensor:
    return 0.5 * (ops.gather_rows(user_weight, users).norm(2).pow(2) + ops.gather_rows(item_weight, pos_items).norm(2).pow(2)
                  + ops.gather_rows(item_weight, neg_items).norm(2).pow(2))


def retrieve(user_emb: Tensor, all_items: Tensor, max_k: int):
    """scores = user_emb @ all_items.T; scores[:, 0] = -inf; topk -> (scores, ids)"""
    return ops.retrieve_topk(user_emb, all_items, max_k, mask_index0=True)
