"""Oracle: embedding gathers and the fused "fronts" of the towers (fp32, CPU).

TEST INFRASTRUCTURE -- see oracle/__init__.py for who may import this.

Rows of SURVEY.md section 8a covered here: U1, U2, U4 (gather side), I1, I2, H1.
Backward passes (U3, I3) are obtained by running torch autograd over these
functions: `F.embedding(..., padding_idx=p)` is the same ATen op the reference
uses, so "row p never receives gradient" is inherited, not re-implemented.
"""
from __future__ import annotations

from typing import Sequence

import torch
import torch.nn.functional as F


def gather_rows(table: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    """`table[ids]` -- the bare row gather.

    Reference call sites: `pretrained_lookup[item_ids.cpu()]`
    tower_code/v1_usertower_train.py:760; `item_tower_emb[target_ids]`
    tower_code/v1_refine_usertower.py:833; `gnn_user_emb(u_idx)`
    tower_code/mined_inference.py:670.  The forward returns the STORED row
    even for a padding index (SURVEY.md 8c invariant 1).
    """
    return table[ids]


def seq_front(base: torch.Tensor,
              ids: Sequence[torch.Tensor],
              tables: Sequence[torch.Tensor],
              gates: torch.Tensor,
              pos_table: torch.Tensor,
              padding_idx: int = 0) -> torch.Tensor:
    """U1: `base + sum_t gates[t] * tables[t][ids[t]] + pos_table[arange(L)]`.

    Follows tower_code/v1_refine_usertower.py:447-456 including its order of
    operations (each gathered tensor is multiplied by its gate, rounded, then
    added in place, left to right; the positional row goes last), so an fp32
    kernel that keeps the same order is bit-exact against this function.
    `gates` is `sigmoid(seq_gate) * s_mask` (:434-438) -- computed by the caller.
    `base` is `item_proj(pretrained_vecs)` (:447).
    """
    out = base.clone()
    for t, (idx, tab) in enumerate(zip(ids, tables)):
        out += F.embedding(idx, tab, padding_idx=padding_idx) * gates[t]
    L = base.shape[1]
    out += pos_table[torch.arange(L)].unsqueeze(0)
    return out


def layer_norm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """`nn.LayerNorm` over the last dim (emb_ln :458; std_ln item_tower.py:241)."""
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def static_front(ids: Sequence[torch.Tensor],
                 tables: Sequence[torch.Tensor],
                 cont_feats: torch.Tensor,
                 cont_w: torch.Tensor,
                 cont_b: torch.Tensor,
                 gates: torch.Tensor,
                 padding_idx: int = 0) -> torch.Tensor:
    """U2: nine gated tiny gathers + gated relu(Linear(4->16)), concatenated.

    Follows tower_code/v1_refine_usertower.py:472-491.  `gates` is
    `sigmoid(static_gate) * u_mask` with u_mask == 1 (:440-442), length 10.
    Output [B, 100] in the reference's column order (age, price, cnt, recency,
    channel, club, news, fn, active, cont).
    """
    cols = [F.embedding(idx, tab, padding_idx=padding_idx) * gates[i]
            for i, (idx, tab) in enumerate(zip(ids, tables))]
    cols.append(F.relu(F.linear(cont_feats, cont_w, cont_b)) * gates[len(cols)])
    return torch.cat(cols, dim=1)


def normalized_rows(table: torch.Tensor, ids: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    """U4: `F.normalize(table, p=2, dim=1)[ids]`.

    Follows tower_code/v1_usertower_train.py:810-811 + v1_refine_usertower.py:833:
    the reference normalises the WHOLE table each step and then gathers.
    """
    return F.normalize(table, p=2, dim=1, eps=eps)[ids]


def std_front(std_input: torch.Tensor,
              std_table: torch.Tensor,
              field_emb: torch.Tensor,
              ln_w: torch.Tensor,
              ln_b: torch.Tensor,
              eps: float = 1e-5,
              padding_idx: int = 0) -> torch.Tensor:
    """I1: `LN(E_std[std_input] + std_field_emb)` -- item_tower.py:239-241."""
    x = F.embedding(std_input, std_table, padding_idx=padding_idx) + field_emb
    return layer_norm(x, ln_w, ln_b, eps)


def bert_embeddings_eval(input_ids: torch.Tensor,
                         word: torch.Tensor,
                         pos: torch.Tensor,
                         tok_type: torch.Tensor,
                         ln_w: torch.Tensor,
                         ln_b: torch.Tensor,
                         eps: float = 1e-12) -> torch.Tensor:
    """I2 (first half): HF `BertEmbeddings` in eval mode (dropout off).

    Called at item_tower.py:249 as `self.bert_model.embeddings(input_ids=ids)`:
    LN(word[ids] + token_type[0] + pos[arange(T)]) with eps = 1e-12
    (transformers `BertEmbeddings.forward`; absolute positions, all-zero
    token types).  Add order follows HF: (word + type) + pos.
    """
    T = input_ids.shape[-1]
    x = word[input_ids] + tok_type[0]
    x = x + pos[torch.arange(T)]
    return layer_norm(x, ln_w, ln_b, eps)


def masked_mean_pool(feats: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """I2 (second half): mask-weighted mean over tokens -- item_tower.py:254-257.

    feats [R, T, D], mask [R, T] (any numeric dtype) -> [R, D];
    count is clamped at 1e-9 so an all-zero mask gives a zero vector.
    """
    m = mask.unsqueeze(-1).to(feats.dtype)
    return (feats * m).sum(dim=1) / m.sum(dim=1).clamp(min=1e-9)


def hybrid_time_rows(time_table: torch.Tensor, seq_deltas: torch.Tensor) -> torch.Tensor:
    """H1: `time_emb(seq_deltas.clamp(max=1000))` -- tower_code/mined_inference.py:695."""
    return time_table[seq_deltas.clamp(max=1000)]


def re_front(re_input_ids, re_attn_mask, word, pos, tok_type, bert_ln_w, bert_ln_b,
             proj_w, proj_b, proj_ln_w, proj_ln_b, field_pos, re_ln_w, re_ln_b, bert_eps=1e-12):
    """I2 whole: item_tower.py:246-261 in eval mode.

    BERT embeddings (no grad) -> re_proj = Linear(768->128), LayerNorm, GELU (:160-164)
    -> mask-weighted mean over tokens -> + re_field_position -> re_ln.  Returns
    (word_embs [B*9,T,768], re_vectors [B,9,128])."""
    B, Fn, T = re_input_ids.shape
    flat = re_input_ids.reshape(-1, T)
    with torch.no_grad():
        we = bert_embeddings_eval(flat, word, pos, tok_type, bert_ln_w, bert_ln_b, bert_eps)
    h = F.gelu(layer_norm(F.linear(we, proj_w, proj_b), proj_ln_w, proj_ln_b))
    pooled = masked_mean_pool(h, re_attn_mask.reshape(-1, T))
    rv = pooled.view(B, Fn, -1) + field_pos
    return we, layer_norm(rv, re_ln_w, re_ln_b)
