"""CPU restatement (test infrastructure only) of the callers / data formats either side of the hot path
(SURVEY.md 8f, rows N2-N4).  Plain torch / numpy on the host, every function citing the reference lines it follows;
pinned by tests/golden/{alignment,ensemble,lightgcl}.pt, which tests/golden/make_golden.py produced by running the
reference's own code."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------- N2
def align_pretrained(pretrained, pretrained_ids, item_ids, dim):
    """tower_code/v1_usertower_train.py:131-160: rows of the exported matrix placed at 1 + (position of their id in the
    processor's catalogue); unmatched rows keep `randn * 0.01` (drawn FIRST, from torch's global generator, :138), row 0
    = 0.  A dict keyed by str(id) is built in export order, so of a repeated id the LAST row wins (:147-148).
    `pretrained` None = files missing: random init only (:157-158)."""
    aligned = torch.randn(len(item_ids) + 1, dim) * 0.01
    aligned[0] = 0.0
    if pretrained is None:
        return aligned
    if isinstance(pretrained, dict):
        pretrained = pretrained.get("weight", pretrained.get("item_content_emb.weight"))
    table = {str(i.item()) if isinstance(i, torch.Tensor) else str(i): r for r, i in enumerate(pretrained_ids)}
    for i, cur in enumerate(item_ids):
        r = table.get(cur)
        if r is not None:
            aligned[i + 1] = pretrained[r]
    return aligned


# --------------------------------------------------------------------------------------------- N4
def candidate_union(user_a, items_a, user_b, items_b, pool_k):
    """mined_inference.py:1103-1133: two global top-M, side by side; both models re-score the union."""
    ia = torch.topk(user_a @ items_a.T, k=pool_k, dim=1).indices
    ib = torch.topk(user_b @ items_b.T, k=pool_k, dim=1).indices
    comb = torch.cat([ia, ib], dim=1)
    sa = (user_a.unsqueeze(1) * items_a[comb]).sum(-1)
    sb = (user_b.unsqueeze(1) * items_b[comb]).sum(-1)
    return comb, sa, sb


def min_max_norm(t):
    """:1139-1142"""
    lo, hi = t.min(dim=1, keepdim=True)[0], t.max(dim=1, keepdim=True)[0]
    return (t - lo) / (hi - lo + 1e-9)


def reciprocal_ranks(s, k_rrf):
    """:1358-1380: rank by a descending sort (ties: position ascending, what a stable sort gives), 1 / (k + rank + 1)."""
    order = torch.sort(s, dim=1, descending=True, stable=True).indices
    rank = torch.zeros_like(s)
    rank.scatter_(1, order, torch.arange(s.shape[1]).expand(s.shape[0], -1).float())
    return 1.0 / (k_rrf + rank + 1.0), rank


def dedup_keep_order(ids_row):
    """:1182-1183: np.unique(return_index) + sort of the first occurrences"""
    _, first = np.unique(ids_row, return_index=True)
    return ids_row[np.sort(first)]


def blend_and_rank(comb, n1, n2, alpha, k_sel):
    """:1162-1183 / :1393-1407 for one alpha: list (per user) of distinct global ids in ranking order."""
    final = alpha * n1 + (1.0 - alpha) * n2
    local = torch.topk(final, k=k_sel, dim=1).indices.numpy()
    c = comb.numpy()
    return final, [dedup_keep_order(c[i][local[i]]) for i in range(c.shape[0])]


# --------------------------------------------------------------------------------------------- N3
def lightgcl_bpr(local_emb, users, pos_items, neg_items):
    """gnn_model/v1_lightgcl.py:188-195"""
    u, p, n = local_emb[users], local_emb[pos_items], local_emb[neg_items]
    return -torch.mean(torch.log(torch.sigmoid((u * p).sum(1) - (u * n).sum(1)) + 1e-10))


def lightgcl_ssl(local_emb, global_emb, users, items, temp):
    """gnn_model/v1_lightgcl.py:197-213: InfoNCE between the two views over the distinct users and the distinct items of
    the batch, logits clamped at 100."""
    a, b = F.normalize(local_emb, dim=1), F.normalize(global_emb, dim=1)

    def nce(idx):
        logits = torch.clamp(a[idx] @ b[idx].t() / temp, max=100.0)
        return F.cross_entropy(logits, torch.arange(logits.shape[0]))
    return nce(torch.unique(users)) + nce(torch.unique(items))


def lightgcl_reg(user_w, item_w, users, pos_items, neg_items):
    """gnn_model/v1_lightgcl.py:215-219"""
    return 0.5 * (user_w[users].norm(2).pow(2) + item_w[pos_items].norm(2).pow(2) + item_w[neg_items].norm(2).pow(2))
