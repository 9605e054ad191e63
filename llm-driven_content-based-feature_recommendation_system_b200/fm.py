"""F1: FM second-order interaction / DeepFM reranker on the fused gather kernel (csrc/fm.cu).

The reference has no FM or DeepFM (temp_model/ranker_skelet.py is CatBoost); the
module follows the published deepctr-torch 0.2.9 `DeepFM` (requirements.txt:41): one embedding
vector of size k and one linear weight per (field, id); logit = linear + FM + DNN(concat).
All field tables live in ONE concatenated [sum_vocab, k] parameter (row offsets per field), so a
sample's 39 rows are fetched by one warp in one pass and the (sum, sum of squares) pair never
leaves registers.
"""
from __future__ import annotations

from typing import Sequence

import torch
import torch.nn as nn

from . import ops


class FM(nn.Module):
    """deepctr_torch.layers.interaction.FM on already-gathered [B,F,k] inputs is replaced by
    `fm_interaction(ids)`: gather + 0.5*sum((sum v)^2 - sum v^2) fused."""

    def __init__(self, vocab_sizes: Sequence[int], k: int = 16, init_std: float = 1e-4, linear: bool = True):
        super().__init__()
        offs = torch.zeros(len(vocab_sizes), dtype=torch.int64)
        offs[1:] = torch.cumsum(torch.tensor(list(vocab_sizes[:-1]), dtype=torch.int64), 0)
        self.register_buffer("offsets", offs)
        total = int(sum(vocab_sizes))
        self.embedding = nn.Parameter(torch.randn(total, k) * init_std)
        self.linear = nn.Parameter(torch.randn(total, 1) * init_std) if linear else None
        self.k, self.num_fields = k, len(vocab_sizes)

    def forward(self, ids: torch.Tensor, want_concat: bool = True):
        """ids [B,F] per-field local ids -> (linear + fm)[B], concat[B,F*k]"""
        return ops.fm_interaction(ids, self.offsets, self.embedding, self.linear, want_concat)


class DeepFM(nn.Module):
    def __init__(self, vocab_sizes: Sequence[int], k: int = 16, dnn_hidden_units=(256, 128), init_std: float = 1e-4,
                 dnn_dropout: float = 0.0):
        super().__init__()
        self.fm = FM(vocab_sizes, k, init_std, linear=True)
        dims = [len(vocab_sizes) * k, *dnn_hidden_units]
        layers = []
        for i in range(len(dims) - 1):
            layers += [nn.Linear(dims[i], dims[i + 1]), nn.ReLU()]
            if dnn_dropout > 0:
                layers.append(nn.Dropout(dnn_dropout))
        self.dnn = nn.Sequential(*layers)
        self.dnn_linear = nn.Linear(dims[-1], 1, bias=False)
        self.bias = nn.Parameter(torch.zeros(1))

    def logits(self, ids):
        y, concat = self.fm(ids, want_concat=True)
        return y + self.dnn_linear(self.dnn(concat)).squeeze(-1).float() + self.bias

    def forward(self, ids):
        return torch.sigmoid(self.logits(ids))
