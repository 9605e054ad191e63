"""Oracle: in-batch contrastive losses (fp32, CPU), logits materialised.

TEST INFRASTRUCTURE -- see oracle/__init__.py for who may import this.

Rows of SURVEY.md section 8a covered: C1, C2, C3, C4, C5.  Every function
builds the full [N, N] logits the way the reference does; the CUDA path never
materialises them, which is exactly what these functions check.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

NEG_INF = float("-inf")


def _diag_ce(logits: torch.Tensor) -> torch.Tensor:
    """Mean cross-entropy with label i on row i (`F.cross_entropy(S, arange)`)."""
    n = logits.shape[0]
    return F.cross_entropy(logits, torch.arange(n, device=logits.device))


def _offdiag(mask: torch.Tensor) -> torch.Tensor:
    n = mask.shape[0]
    return mask & ~torch.eye(n, dtype=torch.bool, device=mask.device)


def simcse_loss(emb1: torch.Tensor, emb2: torch.Tensor, temperature: float = 0.08) -> torch.Tensor:
    """C1: symmetric InfoNCE -- item_tower.py:1075-1082.

    S = emb1 @ emb2.T / tau ; (CE(S, diag) + CE(S.T, diag)) / 2.
    """
    s = emb1 @ emb2.T / temperature
    return (_diag_ce(s) + _diag_ce(s.T)) / 2


def inbatch_corrected_logq_loss(user_emb, item_tower_emb, target_ids, user_ids, log_q_tensor,
                                temperature: float = 0.1, lambda_logq: float = 1.0) -> torch.Tensor:
    """C2: the EFFECTIVE definition, tower_code/v1_refine_usertower.py:826-861
    (the :520 definition without `user_ids` is shadowed -- invariant 3).

    S = U @ V.T / tau - lambda * logq[tgt][None, :]       (:836-842)
    S[i, j] = -inf where i != j and (tgt_i == tgt_j or uid_i == uid_j)   (:846-857)
    loss = CE(S, diag)                                     (:860-861)
    """
    v = item_tower_emb[target_ids]
    s = (user_emb @ v.T) / temperature
    if lambda_logq > 0.0:
        s = s - log_q_tensor[target_ids].view(1, -1) * lambda_logq
    same_item = target_ids.unsqueeze(1) == target_ids.unsqueeze(0)
    same_user = user_ids.unsqueeze(1) == user_ids.unsqueeze(0)
    s = s.masked_fill(_offdiag(same_item | same_user), NEG_INF)
    return _diag_ce(s)


def inbatch_corrected_logq_loss_columns(user_emb, col_rows, col_item_ids, col_counts, target_ids, pos_col,
                                        own_cols, log_q_tensor, temperature: float = 0.1,
                                        lambda_logq: float = 1.0) -> torch.Tensor:
    """C2 (tower_code/v1_refine_usertower.py:826-861) restated over a MULTISET of columns: in-batch columns
    that share a target item have the same item row and the same logQ, hence the same logit, so the [N, N]
    softmax collapses to [N, U] over the distinct items with their batch multiplicities m_c:

        Z_i = e^{s_i,t_i}                              (the diagonal; all other copies of t_i are masked, :846)
            + sum_{c != t_i} m_c e^{s_ic}              (every in-batch column, grouped by item)
            - sum_{j in user(i), t_j != t_i} e^{s_i,t_j}   (same-user columns are masked too, :848)
        loss = mean_i( log Z_i - s_i,t_i )

    `col_rows[U, D]` rows of the column items, `col_item_ids[U]`, `col_counts[U]` (m_c, may be 0 = absent),
    `pos_col[N]` column of each row's own target, `own_cols[N, K]` columns of the same user's targets
    (-1 = none).  Checked against `inbatch_corrected_logq_loss` in tests/test_oracle_golden.py."""
    s = (user_emb @ col_rows.T) / temperature
    if lambda_logq > 0.0:
        s = s - log_q_tensor[col_item_ids].view(1, -1) * lambda_logq
    n = user_emb.shape[0]
    r = torch.arange(n, device=s.device)
    s_pos = s[r, pos_col]
    w = col_counts.to(s.dtype).view(1, -1).expand(n, -1).clone()
    w[col_item_ids.view(1, -1) == target_ids.view(-1, 1)] = 0.0         # every copy of the row's own target
    ok = own_cols >= 0
    oc = own_cols.clamp(min=0)
    ok = ok & (col_item_ids[oc] != target_ids.view(-1, 1))
    w.scatter_add_(1, oc, -ok.to(s.dtype))                               # same-user columns
    mx = torch.maximum(s.max(dim=1).values, s_pos).detach()
    z = (w * torch.exp(s - mx.view(-1, 1))).sum(1) + torch.exp(s_pos - mx)
    return (mx + torch.log(z) - s_pos).mean()


def inbatch_logq_loss_no_user(user_emb, item_tower_emb, target_ids, log_q_tensor,
                              temperature: float = 0.1, lambda_logq: float = 1.0) -> torch.Tensor:
    """The shadowed first definition (tower_code/v1_refine_usertower.py:520-573):
    C2 with the same-item mask only.  Kept because `train_user_tower`
    (v1_usertower_train.py:479) was written against it."""
    n = user_emb.shape[0]
    return inbatch_corrected_logq_loss(user_emb, item_tower_emb, target_ids,
                                       torch.arange(n, device=user_emb.device), log_q_tensor, temperature, lambda_logq)


def duorec_loss_refined(user_emb_1, user_emb_2, target_ids,
                        temperature: float = 0.1, lambda_sup: float = 0.1) -> torch.Tensor:
    """C3: unsupervised InfoNCE + SupCon -- tower_code/v1_refine_usertower.py:576-627.

    z = normalize(.)                                        (:584-585)
    unsup = CE(z1 @ z2.T / tau, diag)                       (:588-590)
    pos[i, j] = tgt_i == tgt_j, tgt_i != 0, i != j          (:599-606)
    sup_i = -sum_j pos[i,j] * log_softmax(z1 @ z1.T / tau with diag=-inf)[i,j] / sum_j pos[i,j]
            over rows with at least one positive, then mean (:609-625)
    Rows whose target is 0 (padding) have no positives (invariant 8); because
    the pad test multiplies ROW i (`mask * (1 - pad_mask)` with pad_mask [B,1]),
    a pad target only clears its own row -- but a pad target can only equal
    another pad target, whose row is cleared too, so the mask stays symmetric.
    """
    z1 = F.normalize(user_emb_1, dim=1)
    z2 = F.normalize(user_emb_2, dim=1)
    unsup = _diag_ce(z1 @ z2.T / temperature)
    sup = torch.zeros((), dtype=unsup.dtype, device=unsup.device)
    if lambda_sup > 0:
        t = target_ids.view(-1, 1)
        pos = (t == t.T) & (t != 0)
        pos = _offdiag(pos).to(z1.dtype)
        if pos.sum() > 0:
            n = z1.shape[0]
            eye = torch.eye(n, dtype=torch.bool, device=z1.device)
            s = (z1 @ z1.T / temperature).masked_fill(eye, NEG_INF)
            logp = F.log_softmax(s, dim=1).masked_fill(eye, 0.0)
            cnt = pos.sum(1)
            rows = cnt > 0
            if rows.sum() > 0:
                sup = (-(pos[rows] * logp[rows]).sum(1) / cnt[rows]).mean()
    return unsup + lambda_sup * sup


def _hnm_common(user_emb, item_tower_emb, target_ids, hnm_threshold):
    """Shared head of the hard-negative family (:775-783, :643-657, :707-719)."""
    u = F.normalize(user_emb, p=2, dim=1)
    v = F.normalize(item_tower_emb[target_ids], p=2, dim=1)
    cos = u @ v.T
    same_item = target_ids.unsqueeze(1) == target_ids.unsqueeze(0)
    too_similar = _offdiag((v @ v.T) > hnm_threshold)
    return u, v, cos, same_item, same_item | too_similar


def full_batch_hard_emphasis_loss(user_emb, item_tower_emb, target_ids, log_q_tensor,
                                  top_k_percent: float = 0.01, hard_margin: float = 0.2,
                                  hnm_threshold: float = 0.90, temperature: float = 0.1,
                                  lambda_logq: float = 1.0):
    """C4: tower_code/v1_refine_usertower.py:762-822.

    Mining (:786-791): per row, top-k (k = max(1, int((N-1)*pct))) of cos with
    ignore-mask (same item incl. the diagonal, or item-item cos > thr) at -inf.
    Loss (:794-815): cos/tau - lambda*logq[tgt] + margin/tau at mined positions,
    off-diagonal same-item at -inf, CE on the diagonal.
    Returns (loss, {"avg_hn_similarity", "num_hard"}) like the reference.
    NB a mined position may itself be ignore-masked (-inf ties in topk when a
    row has fewer than k finite entries); the reference adds the margin there
    too, and so does this.
    """
    n = user_emb.shape[0]
    _, _, cos, same_item, ignore = _hnm_common(user_emb, item_tower_emb, target_ids, hnm_threshold)
    k = max(1, int((n - 1) * top_k_percent))
    with torch.no_grad():
        mined = torch.topk(cos.detach().masked_fill(ignore, NEG_INF), k=k, dim=1).indices
    s = cos / temperature
    if lambda_logq > 0.0:
        s = s - log_q_tensor[target_ids].view(1, -1) * lambda_logq
    emph = torch.zeros_like(s, dtype=torch.bool).scatter_(1, mined, True)
    s = s + emph.to(s.dtype) * (hard_margin / temperature)
    s = s.masked_fill(_offdiag(same_item), NEG_INF)
    loss = _diag_ce(s)
    with torch.no_grad():
        avg = torch.gather(cos, 1, mined).mean().item()
    return loss, {"avg_hn_similarity": avg, "num_hard": k}


def inbatch_hnm_corrected_loss_with_stats(user_emb, item_tower_emb, target_ids, log_q_tensor,
                                          top_k_percent: float = 0.01, hnm_threshold: float = 0.90,
                                          temperature: float = 0.1, lambda_logq: float = 0.7,
                                          lambda_cl: float = 0.2):
    """C5: sampled hard-negative CE -- tower_code/v1_refine_usertower.py:632-692.

    logits row i = [S_ii, S_i,mined_1 .. S_i,mined_k], label 0; k additionally
    capped by the smallest count of non-ignored columns over rows (:665-666).
    """
    n = user_emb.shape[0]
    _, _, cos, _, ignore = _hnm_common(user_emb, item_tower_emb, target_ids, hnm_threshold)
    avail = (~ignore).sum(dim=1)
    k = max(1, min(int((n - 1) * top_k_percent), int(avail.min().item())))
    mined = torch.topk((cos / temperature).detach().masked_fill(ignore, NEG_INF), k=k, dim=1).indices
    s = cos / temperature
    if lambda_logq > 0.0:
        s = s - log_q_tensor[target_ids].view(1, -1) * lambda_logq
    final = torch.cat([torch.diagonal(s).unsqueeze(1), torch.gather(s, 1, mined)], dim=1)
    loss = F.cross_entropy(final, torch.zeros(n, dtype=torch.long))
    with torch.no_grad():
        avg = torch.gather(cos, 1, mined).mean().item()
    return loss, {"avg_hn_similarity": avg, "num_active_hard_negs": k}


def logq_correction_loss(user_emb, item_emb, pos_item_ids, item_probs,
                         temperature: float = 0.07, lambda_logq: float = 0.0) -> torch.Tensor:
    """C5: tower_code/mined_inference.py:738-749.  NB `item_emb` is already the
    per-row positive matrix [N, D] (no gather), logq is applied BEFORE the
    division by tau, and the collision mask value is -1e4, not -inf."""
    s = user_emb @ item_emb.T
    if lambda_logq > 0.0:
        s = s - lambda_logq * torch.log(item_probs[pos_item_ids] + 1e-4).view(1, -1)
    s = s / temperature
    coll = _offdiag(pos_item_ids.unsqueeze(1) == pos_item_ids.unsqueeze(0))
    return _diag_ce(s.masked_fill(coll, -1e4))


def efficient_corrected_logq_loss(user_emb, item_emb, pos_item_ids, precomputed_log_q,
                                  temperature: float = 0.1, lambda_logq: float = 0.1) -> torch.Tensor:
    """C5: tower_code/mined_inference.py:751-789.  LogQ is subtracted from every
    column, then the DIAGONAL is restored to the raw u.v/tau ("positive
    recovery" :774-775); collisions are filled with -1e9 in fp32 (-3e4 in fp16)."""
    s = (user_emb @ item_emb.T) / temperature
    if lambda_logq > 0.0:
        s = s - precomputed_log_q[pos_item_ids].view(1, -1) * lambda_logq
        raw = (user_emb * item_emb).sum(dim=1) / temperature
        n = s.shape[0]
        eye = torch.eye(n, dtype=torch.bool)
        s = torch.where(eye, raw.unsqueeze(1).expand(n, n), s)
    coll = _offdiag(pos_item_ids.unsqueeze(1) == pos_item_ids.unsqueeze(0))
    fill = -30000.0 if s.dtype == torch.float16 else -1e9
    return _diag_ce(s.masked_fill(coll, fill))


def inbatch_mixed_hnm_loss_with_stats(user_emb, item_tower_emb, target_ids, log_q_tensor, random_indices,
                                      top_k_percent: float = 0.01, hnm_threshold: float = 0.90,
                                      temperature: float = 0.1, lambda_logq: float = 0.7):
    """C5: hard (top-k) + random negatives -- tower_code/v1_refine_usertower.py:695-757.

    The reference draws `random_indices = torch.randint(0, N, (N, M))` inside the
    function (:722); here they are an argument so that the result is a pure
    function.  Random picks that fall on an ignore-masked column get -1e9 (:741-742).
    """
    n = user_emb.shape[0]
    _, _, cos, _, ignore = _hnm_common(user_emb, item_tower_emb, target_ids, hnm_threshold)
    k = max(1, int((n - 1) * top_k_percent))
    mined = torch.topk((cos / temperature).detach().masked_fill(ignore, NEG_INF), k=k, dim=1).indices
    s = cos / temperature
    if lambda_logq > 0.0:
        s = s - log_q_tensor[target_ids].view(1, -1) * lambda_logq
    rnd = torch.gather(s, 1, random_indices).masked_fill(torch.gather(ignore, 1, random_indices), -1e9)
    final = torch.cat([torch.diagonal(s).unsqueeze(1), torch.gather(s, 1, mined), rnd], dim=1)
    loss = F.cross_entropy(final, torch.zeros(n, dtype=torch.long))
    with torch.no_grad():
        avg = torch.gather(cos, 1, mined).mean().item()
    return loss, {"avg_hn_similarity": avg, "num_hard": k, "num_random": random_indices.shape[1]}
