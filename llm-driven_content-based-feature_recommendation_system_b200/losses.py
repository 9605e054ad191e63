"""In-batch contrastive losses with the reference's signatures, computed by the fused
tcgen05 softmax kernel (csrc/infonce.cu): the [N,N] logits, the three [N,N] masks and
the log-softmax temporaries of the reference never exist in memory.

Reference functions mirrored (paths relative to the reference checkout):
    simcse_loss                        item_tower.py:1075-1082 (inline)
    inbatch_corrected_logq_loss        tower_code/v1_refine_usertower.py:826-861 (effective def)
    inbatch_logq_loss_no_user          tower_code/v1_refine_usertower.py:520-573 (shadowed def)
    duorec_loss_refined                tower_code/v1_refine_usertower.py:576-627
    logq_correction_loss               tower_code/mined_inference.py:738-749
    efficient_corrected_logq_loss      tower_code/mined_inference.py:751-789

Precision: the similarity contraction runs on the tensor cores in `COMPUTE_DTYPE`
(bf16 by default -- BASELINE.json; the reference autocasts its matmul to fp16) with
fp32 accumulation; softmax, logsumexp and the loss are fp32 like the reference's
autocast policy (SURVEY.md 8c invariant 5).
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F
from torch import Tensor

from . import _lib as L
from . import ops

_lib = L.load()

COMPUTE_DTYPE = torch.bfloat16
UNIT_NORM_BOUND = 1.0 + 2.0 ** -6        # |<a, b>| of 16-bit-rounded unit rows
NEG_INF = float("-inf")
# True: when the row operand needs a gradient the forward pass also accumulates its unnormalised gradient on the tensor
# cores (ops.ce_fwd_grad, "flash" form) and the backward runs one tensor-core pass (dB side) instead of two.
FUSE_ROW_GRAD = True
# the per-row tail of logq_infonce_columns in one kernel each way (False: the torch expression, kept as the checker)
FUSE_ROW_COMBINE = True


def _operand_dtype(*ts) -> torch.dtype:
    if torch.is_autocast_enabled("cuda"):
        d = torch.get_autocast_dtype("cuda")
        if d in (torch.float16, torch.bfloat16):
            return d
    for t in ts:
        if t.dtype in (torch.float16, torch.bfloat16):
            return t.dtype
    return COMPUTE_DTYPE


# one-entry-per-tensor memo of 16-bit operand casts inside one loss call: the fused softmax and the user-block kernel
# of losses.logq_infonce_columns read the SAME rows and columns, each through its own autograd Function (so that the
# gradients come back in fp32): the cast is done once.  Keyed by storage + version; holds no tensor across calls.
_cast_memo: dict = {}


def _cast16(t: Tensor, dtype: torch.dtype) -> Tensor:
    t = t.detach()
    if t.dtype == dtype and t.is_contiguous():
        return t
    key = (t.data_ptr(), t._version, tuple(t.shape), t.dtype, dtype)
    hit = _cast_memo.get(key)
    if hit is not None:
        return hit
    out = t.to(dtype).contiguous()
    if _cast_memo_on:
        _cast_memo[key] = out
    return out


_cast_memo_on = False


class _FusedSoftmax(torch.autograd.Function):
    """(lse, diag, pos_sum, pos_cnt) = rows of softmax statistics of S = scale*A@B^T - bias (+masks)."""

    @staticmethod
    def forward(ctx, a, b, scale, col_bias, key_a_row, key_a_col, key_b_row, key_b_col, diag_offset, mask_value,
                flags, dtype, logit_bound):
        a16, b16 = _cast16(a, dtype), _cast16(b, dtype)
        g_parts = g_info = None
        if FUSE_ROW_GRAD and ctx.needs_input_grad[0] and ops.ce_fwd_grad_supported(
                a16, flags, mask_value, logit_bound, key_a_row is not None or key_b_row is not None):
            # the forward also accumulates G = sum_j e_ij b_j: the backward's dA is then a row scaling (rs_ce_fwd_grad)
            lse, diag, g_parts, g_info = L.direct.ce_fwd_grad(a16, b16, scale, col_bias, key_a_row, key_a_col,
                                                              key_b_row, key_b_col, diag_offset, mask_value, flags,
                                                              logit_bound)
            pos_sum, pos_cnt = lse.new_empty(0), lse.new_empty(0)
        else:
            lse, diag, pos_sum, pos_cnt = L.direct.ce_fwd(a16, b16, scale, col_bias, key_a_row, key_a_col, key_b_row,
                                                          key_b_col, diag_offset, mask_value, flags, logit_bound)
        ctx.save_for_backward(a16, b16, col_bias, key_a_row, key_a_col, key_b_row, key_b_col, lse, g_parts, g_info)
        ctx.meta = (scale, diag_offset, mask_value, flags, a.dtype, b.dtype, logit_bound)
        ctx.mark_non_differentiable(pos_cnt)
        return lse, diag, pos_sum, pos_cnt

    @staticmethod
    def backward(ctx, g_lse, g_diag, g_pos, _g_cnt):
        a16, b16, col_bias, kar, kac, kbr, kbc, lse, g_parts, g_info = ctx.saved_tensors
        scale, diag_offset, mask_value, flags, adt, bdt, logit_bound = ctx.meta
        M = a16.shape[0]
        zeros = None
        if g_lse is None:
            zeros = torch.zeros(M, dtype=torch.float32, device=a16.device)
            g_lse = zeros
        w_pos = g_pos.float().contiguous() if (g_pos is not None and flags & L.RS_CE_SUPCON) else None
        w_diag = None if g_diag is None else g_diag.float().contiguous()
        dA, dB = L.direct.ce_bwd(a16, b16, scale, col_bias, kar, kac, kbr, kbc, diag_offset, mask_value, flags,
                                     lse, g_lse.float().contiguous(), w_diag, w_pos, logit_bound, g_parts, g_info)
        return (dA.to(adt), dB.to(bdt)) + (None,) * 11


def fused_softmax_stats(a: Tensor, b: Tensor, scale: float, col_bias: Optional[Tensor] = None,
                        key_a_row: Optional[Tensor] = None, key_a_col: Optional[Tensor] = None,
                        key_b_row: Optional[Tensor] = None, key_b_col: Optional[Tensor] = None,
                        diag_offset: int = 0, mask_value: float = NEG_INF, flags: int = 0,
                        dtype: Optional[torch.dtype] = None, unit_norm: bool = False):
    """`unit_norm`: the caller vouches that the rows of a and b have norm <= 1 (|logit| <= scale + bias range): the
    kernels then use a fixed softmax offset instead of a running maximum (rs_ce_problem.logit_bound)."""
    dtype = dtype or _operand_dtype(a, b)
    if col_bias is not None:
        col_bias = col_bias.detach().float().contiguous()
    return _FusedSoftmax.apply(a, b, float(scale), col_bias, key_a_row, key_a_col, key_b_row, key_b_col,
                               int(diag_offset), float(mask_value), int(flags), dtype, UNIT_NORM_BOUND if unit_norm else 0.0)


def info_nce(a: Tensor, b: Tensor, temperature: float, **kw) -> Tensor:
    """mean_i( logsumexp_j S_ij - S_i,i+off ) == F.cross_entropy(S, arange + off)."""
    lse, diag, _, _ = fused_softmax_stats(a, b, 1.0 / temperature, **kw)
    return (lse - diag).mean()


# --------------------------------------------------------------------------------------------- C1
def simcse_loss(emb1: Tensor, emb2: Tensor, temperature: float = 0.08) -> Tensor:
    """(CE(S, diag) + CE(S^T, diag)) / 2 with S = emb1 @ emb2.T / temperature  -- item_tower.py:1075-1082."""
    return 0.5 * (info_nce(emb1, emb2, temperature) + info_nce(emb2, emb1, temperature))


# --------------------------------------------------------------------------------------------- C2
def inbatch_corrected_logq_loss(user_emb: Tensor, item_tower_emb: Tensor, target_ids: Tensor, user_ids: Tensor,
                                log_q_tensor: Tensor, temperature: float = 0.1, lambda_logq: float = 1.0) -> Tensor:
    """tower_code/v1_refine_usertower.py:826-861.  `item_tower_emb` is the (normalised) item table; the
    rows of the batch targets are gathered here (bit-exact gather, dense scatter-add backward)."""
    v = ops.gather_rows(item_tower_emb, target_ids)
    return logq_infonce_rows(user_emb, v, target_ids, user_ids, log_q_tensor, temperature, lambda_logq)


def logq_infonce_rows(user_emb: Tensor, item_rows: Tensor, target_ids: Tensor, user_ids: Optional[Tensor],
                      log_q_tensor: Tensor, temperature: float = 0.1, lambda_logq: float = 1.0,
                      col_rows: Optional[Tensor] = None, col_target_ids: Optional[Tensor] = None,
                      col_user_ids: Optional[Tensor] = None, diag_offset: int = 0, unit_norm: bool = False) -> Tensor:
    """C2 on already-gathered item rows.  With `col_*` given the columns are a larger (all-gathered)
    set of negatives and `diag_offset` locates this rank's positives inside it."""
    cols = item_rows if col_rows is None else col_rows
    ct = target_ids if col_target_ids is None else col_target_ids
    cu = user_ids if col_user_ids is None else col_user_ids
    bias = (log_q_tensor[ct] * lambda_logq) if lambda_logq > 0.0 else None
    return info_nce(user_emb, cols, temperature, col_bias=bias, key_a_row=target_ids, key_a_col=ct,
                    key_b_row=user_ids, key_b_col=cu if user_ids is not None else None,
                    diag_offset=diag_offset, mask_value=NEG_INF, unit_norm=unit_norm)


def item_columns(target_ids: Tensor, num_items: Optional[int] = None):
    """Column multiset of a batch of targets for `logq_infonce_columns` (pure index arithmetic; run it where the
    ids live -- the loader does it on the host so that the data-dependent column count is known without a
    device synchronisation).  Returns (col_item_ids[U], col_counts[U], pos_col[N]).
      num_items None : the DISTINCT targets of the batch (U = number of distinct items)
      num_items int  : every item of the catalogue (U = num_items, static; absent items get count 0)"""
    if num_items is None:
        ids, pos_col, counts = torch.unique(target_ids, return_inverse=True, return_counts=True)
        return ids, counts, pos_col
    return torch.arange(num_items, device=target_ids.device), count_ids(target_ids, num_items), target_ids


def count_ids(ids: Tensor, n: int) -> Tensor:
    """bincount(ids, minlength=n) as fp32, without torch.bincount's device->host read of max(ids)."""
    return torch.zeros(n, dtype=torch.float32, device=ids.device).scatter_add_(
        0, ids, torch.ones(1, dtype=torch.float32, device=ids.device).expand(ids.numel()))


def logq_infonce_columns(user_emb: Tensor, col_rows: Tensor, col_item_ids: Tensor, col_counts: Tensor,
                         target_ids: Tensor, pos_col: Tensor, own_cols: Optional[Tensor], log_q_tensor: Tensor,
                         temperature: float = 0.1, lambda_logq: float = 1.0, row_cu: Optional[Tensor] = None,
                         max_rows_per_user: int = 0, unit_norm: bool = False,
                         row_weight: Optional[Tensor] = None) -> Tensor:
    """C2 (tower_code/v1_refine_usertower.py:826-861) over the DISTINCT items of the batch.  In-batch columns with
    the same target item share the item row and the logQ, hence the logit, so the reference's [N, N] softmax
    equals an [N, U] softmax over distinct items with the batch multiplicities m_c folded into the column bias:

        Z_i = e^{s_i,t_i}                                  (the label; every other copy of t_i is masked, :846)
            + sum_{c != t_i} m_c e^{s_ic}                  (fused tensor-core pass, bias_c = lambda logq_c - log m_c)
            - sum_{j in user(i), t_j != t_i} e^{s_i,t_j}   (same-user columns are masked too, :848; <= L per row)

    U is ~5x smaller than N on an H&M-shaped batch (Zipf targets), and it is bounded by the catalogue size
    however many ranks contribute columns.  `col_counts` may hold zeros (absent items: bias = +inf).
    The same-user term: either `row_cu` (int32 [n_users+1], rows grouped by user, at most `max_rows_per_user` <= 64
    each -> one small dense block per user, ops.user_block_logits) or `own_cols[N, K]` = columns of the row's
    user's targets, -1 = none (generic sparse path); both None: no same-user mask.
    `row_weight` [N] (fp32): the loss is sum_i row_weight_i * loss_i instead of the mean -- the bucketed, device-built
    batch index (ops.batch_index_build) pads rows and columns to static capacities: padding rows carry weight 0 (real
    rows 1/N_true) and padding columns count 0 (bias +inf: no softmax mass)."""
    dtype = _operand_dtype(user_emb, col_rows)
    scale = 1.0 / temperature
    lq = (log_q_tensor[col_item_ids] * lambda_logq).float() if lambda_logq > 0.0 else None
    bias = -torch.log(col_counts.float())
    if lq is not None:
        bias = bias + lq
    global _cast_memo_on
    _cast_memo_on = True
    try:
        return _columns_body(user_emb, col_rows, col_item_ids, target_ids, pos_col, own_cols, lq, bias, scale, dtype,
                             row_cu, max_rows_per_user, unit_norm, row_weight)
    finally:
        _cast_memo_on = False
        _cast_memo.clear()


class _RowCombine(torch.autograd.Function):
    """sum_i w_i (log(e^{lse0_i} + e^{pos_i} - e^{own_i}) - pos_i) and its three gradient vectors in one pass
    (rs_ce_row_combine) -- the per-row tail of logq_infonce_columns."""

    @staticmethod
    def forward(ctx, lse0, pos, own, row_weight):
        lse0, pos = lse0.float().contiguous(), pos.float().contiguous()
        own = own.float().contiguous() if own is not None else None
        rw = row_weight.float().contiguous() if row_weight is not None else None
        n = lse0.numel()
        loss = torch.empty((), dtype=torch.float32, device=lse0.device)
        c = torch.empty(3 if own is not None else 2, n, dtype=torch.float32, device=lse0.device)
        ws = L.workspace(_lib.rs_ce_row_combine_workspace_bytes(n), lse0.device)
        L.check(_lib.rs_ce_row_combine(L.ptr(lse0), L.ptr(pos), L.ptr(own), L.ptr(rw), n, L.ptr(loss), L.ptr(c[0]),
                                       L.ptr(c[1]), L.ptr(c[2]) if own is not None else None, L.ptr(ws), ws.numel(),
                                       L.stream()), "rs_ce_row_combine")
        ctx.save_for_backward(c)
        ctx.has_own = own is not None
        return loss

    @staticmethod
    def backward(ctx, g):
        (c,) = ctx.saved_tensors
        d = c * g                                   # one launch for all three vectors
        return d[0], d[1], (d[2] if ctx.has_own else None), None


def row_combine(lse0: Tensor, pos: Tensor, own: Optional[Tensor], row_weight: Optional[Tensor]) -> Tensor:
    L.require_cuda(lse0, pos)
    return _RowCombine.apply(lse0, pos, own, row_weight)


def _columns_body(user_emb, col_rows, col_item_ids, target_ids, pos_col, own_cols, lq, bias, scale, dtype, row_cu,
                  max_rows_per_user, unit_norm, row_weight):
    lse0 = fused_softmax_stats(user_emb, col_rows, scale, col_bias=bias, key_a_row=target_ids, key_a_col=col_item_ids,
                               mask_value=NEG_INF, flags=L.RS_CE_NO_DIAG, dtype=dtype, unit_norm=unit_norm)[0]
    if row_cu is not None:
        s_pos, own_lse = ops.user_block_logits(user_emb, col_rows, pos_col, row_cu, max_rows_per_user, scale, lq,
                                               compute_dtype=dtype)
        if FUSE_ROW_COMBINE:
            return row_combine(lse0, s_pos, own_lse, row_weight)
        mx = torch.maximum(lse0, s_pos).detach()
        z = torch.exp(lse0 - mx) + torch.exp(s_pos - mx) - torch.exp(own_lse - mx)
    else:
        s_pos = ops.sparse_logits(user_emb, col_rows, pos_col.view(-1, 1), scale, lq, compute_dtype=dtype).squeeze(1)
        mx = torch.maximum(lse0, s_pos).detach()
        z = torch.exp(lse0 - mx) + torch.exp(s_pos - mx)
        if own_cols is not None and own_cols.shape[1] > 0:
            s_own = ops.sparse_logits(user_emb, col_rows, own_cols, scale, lq, target_ids, col_item_ids,
                                      compute_dtype=dtype)
            z = z - torch.exp(s_own - mx.unsqueeze(1)).sum(dim=1)
    lse = mx + torch.log(z.clamp_min(1e-30))
    if row_weight is not None:
        return ((lse - s_pos) * row_weight).sum()
    return (lse - s_pos).mean()


def inbatch_logq_loss_no_user(user_emb, item_tower_emb, target_ids, log_q_tensor, temperature=0.1, lambda_logq=1.0):
    """The shadowed first definition (tower_code/v1_refine_usertower.py:520-573): same-item mask only."""
    v = ops.gather_rows(item_tower_emb, target_ids)
    return logq_infonce_rows(user_emb, v, target_ids, None, log_q_tensor, temperature, lambda_logq)


# --------------------------------------------------------------------------------------------- C3
def duorec_loss_refined(user_emb_1: Tensor, user_emb_2: Tensor, target_ids: Tensor, temperature: float = 0.1,
                        lambda_sup: float = 0.1) -> Tensor:
    """tower_code/v1_refine_usertower.py:576-627: InfoNCE(z1, z2) + lambda_sup * SupCon(z1; same target).
    No host synchronisation: the reference's `if mask.sum() > 0` / `valid_rows.sum() > 0` tests (:608,:623)
    become arithmetic on device scalars."""
    from . import encoder                      # (late: encoder imports ops, which this module shares)
    z1 = encoder.l2_normalize(user_emb_1)
    z2 = encoder.l2_normalize(user_emb_2)
    loss = info_nce(z1, z2, temperature, unit_norm=True)
    if lambda_sup > 0:
        lse, _, pos_sum, pos_cnt = fused_softmax_stats(z1, z1, 1.0 / temperature, key_a_row=target_ids,
                                                       key_a_col=target_ids, mask_value=NEG_INF,
                                                       flags=L.RS_CE_DIAG_MASK | L.RS_CE_SUPCON, unit_norm=True)
        valid = pos_cnt > 0
        per_row = torch.where(valid, lse - pos_sum / pos_cnt.clamp(min=1.0), torch.zeros_like(lse))
        sup = per_row.sum() / valid.sum().clamp(min=1).to(per_row.dtype)
        loss = loss + lambda_sup * sup
    return loss


# --------------------------------------------------------------------------------------------- C5
def logq_correction_loss(user_emb, item_emb, pos_item_ids, item_probs, temperature=0.07, lambda_logq=0.0):
    """tower_code/mined_inference.py:738-749: logQ applied before the division by tau, collisions at -1e4."""
    bias = None
    if lambda_logq > 0.0:
        bias = lambda_logq * torch.log(item_probs[pos_item_ids] + 1e-4) / temperature
    return info_nce(user_emb, item_emb, temperature, col_bias=bias, key_a_row=pos_item_ids, key_a_col=pos_item_ids,
                    mask_value=-1e4)


def efficient_corrected_logq_loss(user_emb, item_emb, pos_item_ids, precomputed_log_q, temperature=0.1,
                                  lambda_logq=0.1):
    """tower_code/mined_inference.py:751-789: logQ on every column, raw positive on the diagonal
    ("positive recovery" :774-775), collisions at -1e9 (-3e4 when the logits are fp16)."""
    dtype = _operand_dtype(user_emb, item_emb)
    bias, flags = None, 0
    if lambda_logq > 0.0:
        bias = precomputed_log_q[pos_item_ids] * lambda_logq
        flags = L.RS_CE_DIAG_RAW
    return info_nce(user_emb, item_emb, temperature, col_bias=bias, key_a_row=pos_item_ids, key_a_col=pos_item_ids,
                    mask_value=-30000.0 if dtype == torch.float16 else -1e9, flags=flags, dtype=dtype)


# --------------------------------------------------------------------------------------------- C4 / C5 (hard negatives)
def _hnm_common(user_emb, item_tower_emb, target_ids, top_k_percent, hnm_threshold):
    """Shared head of the hard-negative family (:775-791, :643-668, :707-719): normalise, gather, mine.
    Mining is fused (no [N,N] cos / item-item matrices): ops.mine_hard_negatives."""
    n = user_emb.size(0)
    u = F.normalize(user_emb, p=2, dim=1)
    v = F.normalize(ops.gather_rows(item_tower_emb, target_ids), p=2, dim=1)
    k0 = max(1, int((n - 1) * top_k_percent))
    k0 = min(k0, n)
    if k0 > 1024:
        raise NotImplementedError(
            f"hard-negative mining keeps at most 1024 candidates per row (rs_mine_hard_negatives); N = {n} rows with "
            f"top_k_percent = {top_k_percent} asks for k = {k0}.  Lower top_k_percent below {1024.0 / max(n - 1, 1):.4f} "
            f"or split the batch (the reference runs torch.topk over the materialised [N, N] matrix at this size).")
    scores, idx, avail = ops.mine_hard_negatives(u, v, target_ids, k0, hnm_threshold)
    dtype = _operand_dtype(user_emb, item_tower_emb)
    return u, v, u.to(dtype), v.to(dtype), scores, idx, avail, k0


def _avg_sim(scores, k):
    s = scores[:, :k]
    fin = torch.isfinite(s)
    return (torch.where(fin, s, torch.zeros_like(s)).sum() / fin.sum().clamp(min=1)).item()


def full_batch_hard_emphasis_loss(user_emb, item_tower_emb, target_ids, log_q_tensor, top_k_percent=0.01,
                                  hard_margin=0.2, hnm_threshold=0.90, temperature=0.1, lambda_logq=1.0):
    """tower_code/v1_refine_usertower.py:762-822.  The full [N,N] softmax runs in the fused kernel; the margin
    on the mined positions is an exact sparse correction of each row's log-sum-exp:
        lse_i = lse0_i + log1p( (e^{m/T} - 1) * sum_{j in mined_i} exp(s_ij - lse0_i) ).
    (When a row has fewer than k non-ignored columns the reference's topk returns arbitrary ignored positions;
    here only real candidates are emphasised.)  Returns (loss, stats) like the reference (one host sync for
    the stats' .item())."""
    u, v, u16, v16, scores, idx, avail, k = _hnm_common(user_emb, item_tower_emb, target_ids, top_k_percent,
                                                        hnm_threshold)
    scale = 1.0 / temperature
    bias = (log_q_tensor[target_ids] * lambda_logq) if lambda_logq > 0.0 else None
    lse0, diag, _, _ = fused_softmax_stats(u16, v16, scale, col_bias=bias, key_a_row=target_ids,
                                           key_a_col=target_ids, mask_value=NEG_INF, dtype=u16.dtype)
    s_m = ops.sparse_logits(u16, v16, idx, scale, bias, target_ids, target_ids)
    corr = torch.log1p(math.expm1(hard_margin / temperature) * torch.exp(s_m - lse0.unsqueeze(1)).sum(dim=1))
    loss = (lse0 + corr - diag).mean()
    return loss, {"avg_hn_similarity": _avg_sim(scores, k), "num_hard": k}


def inbatch_hnm_corrected_loss_with_stats(user_emb, item_tower_emb, target_ids, log_q_tensor, top_k_percent=0.01,
                                          hnm_threshold=0.90, temperature=0.1, lambda_logq=0.7, lambda_cl=0.2):
    """tower_code/v1_refine_usertower.py:632-692: CE over [positive, top-k mined negatives]; k is capped by the
    smallest number of non-ignored columns over the rows (:665-666, a host sync in the reference too)."""
    n = user_emb.size(0)
    u, v, u16, v16, scores, idx, avail, k0 = _hnm_common(user_emb, item_tower_emb, target_ids, top_k_percent,
                                                         hnm_threshold)
    k = max(1, min(int((n - 1) * top_k_percent), int(avail.min().item())))
    scale = 1.0 / temperature
    bias = (log_q_tensor[target_ids] * lambda_logq) if lambda_logq > 0.0 else None
    cols = torch.cat([torch.arange(n, device=idx.device).unsqueeze(1), idx[:, :k]], dim=1)
    final = ops.sparse_logits(u16, v16, cols, scale, bias)
    loss = F.cross_entropy(final, torch.zeros(n, dtype=torch.long, device=final.device))
    return loss, {"avg_hn_similarity": _avg_sim(scores, k), "num_active_hard_negs": k}


def inbatch_mixed_hnm_loss_with_stats(user_emb, item_tower_emb, target_ids, log_q_tensor, top_k_percent=0.01,
                                      random_sample_size=100, hnm_threshold=0.90, temperature=0.1, lambda_logq=0.7,
                                      random_indices=None):
    """tower_code/v1_refine_usertower.py:695-757: hard (top-k) + random negatives.  `random_indices` may be
    passed for reproducibility; by default they are drawn like the reference does (:722)."""
    n = user_emb.size(0)
    u, v, u16, v16, scores, idx, avail, k = _hnm_common(user_emb, item_tower_emb, target_ids, top_k_percent,
                                                        hnm_threshold)
    if random_indices is None:
        random_indices = torch.randint(0, n, (n, random_sample_size), device=user_emb.device)
    scale = 1.0 / temperature
    bias = (log_q_tensor[target_ids] * lambda_logq) if lambda_logq > 0.0 else None
    cols = torch.cat([torch.arange(n, device=idx.device).unsqueeze(1), idx, random_indices], dim=1)
    logits = ops.sparse_logits(u16, v16, cols, scale, bias)
    with torch.no_grad():      # ignore mask of the random picks: same item, or item-item cosine above the threshold
        vv = L.direct.sparse_logits(v.detach().float(), v.detach().float(), random_indices, 1.0, None, None, None)
        rows = torch.arange(n, device=idx.device).unsqueeze(1)
        ign = (target_ids[random_indices] == target_ids.unsqueeze(1)) | ((vv > hnm_threshold) & (random_indices != rows))
    rnd = logits[:, 1 + k:].masked_fill(ign, -1e9)
    final = torch.cat([logits[:, :1 + k], rnd], dim=1)
    loss = F.cross_entropy(final, torch.zeros(n, dtype=torch.long, device=final.device))
    return loss, {"avg_hn_similarity": _avg_sim(scores, k), "num_hard": k, "num_random": random_indices.shape[1]}
