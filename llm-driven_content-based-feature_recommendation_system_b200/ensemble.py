"""N4: ensemble of two retrieval models over the union of their top-M candidates.

    evaluate_weighted_score_ensemble   tower_code/mined_inference.py:1103-1189   min-max normalised weighted sum
    evaluate_rrf_ensemble              tower_code/mined_inference.py:1330-1411   weighted reciprocal-rank fusion

Reference per 4096-user batch: two materialised [b, n_items] score matrices + topk, a gather of 2M item vectors per
user for each model, and -- per alpha and per USER, on the host -- np.unique to de-duplicate the ranking.  Here: the two
fused top-k retrievals (rs_retrieve_topk), the re-scoring of the union by the sparse-logits kernel, and one kernel that
normalises, blends, sorts, cuts and de-duplicates for every alpha (rs_ensemble_merge)."""
from __future__ import annotations

from typing import Sequence, Tuple

import torch
from torch import Tensor

from . import _lib as L
from . import ops

_lib = L.load()


def candidate_union(user_a: Tensor, items_a: Tensor, user_b: Tensor, items_b: Tensor, pool_k: int):
    """(:1103-1133) -> (combined_indices [b, 2*pool_k] int64, s_a [b, 2*pool_k], s_b): each model's global top-`pool_k`
    side by side and both models' scores of every candidate (fp32)."""
    _, ia = ops.retrieve_topk(user_a, items_a, pool_k)
    _, ib = ops.retrieve_topk(user_b, items_b, pool_k)
    comb = torch.cat([ia, ib], dim=1)
    sa = L.direct.sparse_logits(user_a.float().contiguous(), items_a.float().contiguous(), comb, 1.0, None, None, None)
    sb = L.direct.sparse_logits(user_b.float().contiguous(), items_b.float().contiguous(), comb, 1.0, None, None, None)
    return comb, sa, sb


def merge(cand_ids: Tensor, s_a: Tensor, s_b: Tensor, alphas: Sequence[float], k_sel: int, mode: str = "minmax",
          k_rrf: float = 60.0, return_norm: bool = False):
    """(:1139-1183 / :1358-1407) -> (ids [len(alphas), b, k_sel] int64, counts [len(alphas), b] int32): for every alpha
    the k_sel best candidates by alpha * n_a + (1 - alpha) * n_b with duplicates dropped in ranking order; rows are -1
    padded behind `counts` distinct ids (the reference's `pred_unique`)."""
    L.require_cuda(cand_ids, s_a, s_b)
    if mode not in ("minmax", "rrf"):
        raise ValueError("mode must be 'minmax' or 'rrf'")
    cand_ids, s_a, s_b = cand_ids.contiguous().to(torch.int64), s_a.float().contiguous(), s_b.float().contiguous()
    b, P = cand_ids.shape
    A = len(alphas)
    ids = torch.empty(A, b, k_sel, dtype=torch.int64, device=cand_ids.device)
    cnt = torch.empty(A, b, dtype=torch.int32, device=cand_ids.device)
    n1 = n2 = None
    if return_norm:
        n1, n2 = torch.empty_like(s_a), torch.empty_like(s_b)
    al = (L.C.c_double * A)(*[float(a) for a in alphas])
    L.check(_lib.rs_ensemble_merge(L.ptr(cand_ids), L.ptr(s_a), L.ptr(s_b), b, P, 0 if mode == "minmax" else 1, float(k_rrf),
                                   al, A, k_sel, L.ptr(ids), L.ptr(cnt), L.ptr(n1), L.ptr(n2), L.stream()), "rs_ensemble_merge")
    return (ids, cnt, n1, n2) if return_norm else (ids, cnt)


def weighted_score_ensemble(user_a, items_a, user_b, items_b, alphas: Sequence[float], pool_k: int = 1000,
                            max_k: int = 500) -> Tuple[Tensor, Tensor]:
    """evaluate_weighted_score_ensemble's retrieval arithmetic for one batch of users."""
    comb, sa, sb = candidate_union(user_a, items_a, user_b, items_b, pool_k)
    return merge(comb, sa, sb, alphas, min(max_k + 20, comb.shape[1]), "minmax")


def rrf_ensemble(user_a, items_a, user_b, items_b, alphas: Sequence[float], pool_k: int = 1000, max_k: int = 500,
                 k_rrf: float = 60.0) -> Tuple[Tensor, Tensor]:
    """evaluate_rrf_ensemble's retrieval arithmetic for one batch of users."""
    comb, sa, sb = candidate_union(user_a, items_a, user_b, items_b, pool_k)
    return merge(comb, sa, sb, alphas, min(max_k + 20, comb.shape[1]), "rrf", k_rrf)
