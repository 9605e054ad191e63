"""Oracle: FM second-order interaction and the DeepFM logit (fp32, CPU).

TEST INFRASTRUCTURE -- see oracle/__init__.py for who may import this.

PARITY UNPINNED.  Row F1 of SURVEY.md section 8a: the reference contains no
FM / DeepFM code at all (`temp_model/ranker_skelet.py` is a
CatBoost re-ranker).  The module it would come from, deepctr-torch 0.2.9
(requirements.txt:41), is neither vendored nor installed and has no call
site, so there is nothing to run or to take golden vectors from.  This file
restates the PUBLISHED algorithm of `deepctr_torch.layers.interaction.FM`
and `deepctr_torch.models.DeepFM`:

    FM(x[B,F,k])  = 0.5 * sum_d( (sum_f x[b,f,d])^2 - sum_f x[b,f,d]^2 )      -> [B,1]
    DeepFM logit  = sum_f w_f[id_f] + FM(E[ids]) + DNN(flatten(E[ids])) (+ bias)

and is reviewed as the specification; tests check the CUDA kernel against it
and check the FM identity against the explicit pairwise sum.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def fm_second_order(x: torch.Tensor) -> torch.Tensor:
    """x [B, F, k] -> [B]  (deepctr FM: square-of-sum minus sum-of-square)."""
    s = x.sum(dim=1)
    return 0.5 * (s * s - (x * x).sum(dim=1)).sum(dim=1)


def fm_pairwise(x: torch.Tensor) -> torch.Tensor:
    """The definition the identity is derived from: sum_{f<g} <x_f, x_g>."""
    g = torch.einsum("bfd,bgd->bfg", x, x)
    iu = torch.triu_indices(x.shape[1], x.shape[1], offset=1)
    return g[:, iu[0], iu[1]].sum(dim=1)


def field_rows(ids: torch.Tensor, table: torch.Tensor, offsets: torch.Tensor) -> torch.Tensor:
    """ids [B,F] per-field local ids; one concatenated table [sum_vocab, k];
    offsets [F] first row of each field -> [B,F,k]."""
    return table[ids + offsets.unsqueeze(0)]


def deepfm_logit(ids, emb_table, lin_table, offsets, mlp, bias=None):
    """DeepFM logit [B]: linear + FM + DNN.  `mlp` is any callable [B,F*k]->[B,1]."""
    e = field_rows(ids, emb_table, offsets)
    lin = lin_table[ids + offsets.unsqueeze(0)].squeeze(-1).sum(dim=1)
    y = lin + fm_second_order(e) + mlp(e.flatten(1)).squeeze(-1)
    return y if bias is None else y + bias
