"""Oracle: dot-product top-k retrieval (fp32, CPU).

TEST INFRASTRUCTURE -- see oracle/__init__.py for who may import this.

Row R1 of SURVEY.md section 8a.
"""
from __future__ import annotations

import torch

NEG_INF = float("-inf")


def retrieve_topk(user_emb: torch.Tensor, item_emb: torch.Tensor, k: int, mask_index0: bool = False,
                  chunk: int = 4096):
    """`topk(user_emb @ item_emb.T, k)` -> (scores [b,k], ids [b,k] int64).

    Follows tower_code/v1_usertower_train.py:672-675 (scores materialised per
    user batch, fp32, then `torch.topk`); with `mask_index0` column 0 is set to
    -inf first, as tower_code/mined_inference.py:1536-1542 does for the GNN
    variant.  Users are processed in chunks of 4096 (the reference's eval batch,
    mined_inference.py:799) so that the CPU baseline has the reference's shape.
    `torch.topk` leaves the order of equal scores unspecified (SURVEY.md 8c
    invariant 7): compare ids only after `canonical_ids`.
    """
    out_s, out_i = [], []
    for lo in range(0, user_emb.shape[0], chunk):
        s = user_emb[lo:lo + chunk] @ item_emb.T
        if mask_index0:
            s[:, 0] = NEG_INF
        sc, idx = torch.topk(s, k=k, dim=-1)
        out_s.append(sc)
        out_i.append(idx)
    return torch.cat(out_s), torch.cat(out_i)


def canonical_ids(scores: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    """Re-order each row so that equal scores appear by ascending id (the CUDA
    path's documented tie policy: higher score first, lower id first)."""
    # sort by id ascending first, then stable sort by score descending
    ids_sorted, p = torch.sort(ids, dim=1, stable=True)
    sc = torch.gather(scores, 1, p)
    _, q = torch.sort(sc, dim=1, descending=True, stable=True)
    return torch.gather(ids_sorted, 1, q)
