"""Synthetic H&M-shaped inputs: the same generator feeds the tests, bench.py
(both arms) and smoke().  Pure torch on the CPU, fixed seed -> identical batches everywhere."""
from __future__ import annotations

from types import SimpleNamespace

import torch
import torch.nn.functional as F

N_ITEMS = 105_542          # H&M articles (BASELINE.json: ~105k)
N_CUSTOMERS = 1_371_980    # H&M customers


def tower_args(num_items=N_ITEMS, max_len=50, d_model=128, side=1000):
    """PipelineConfig fields used by SASRecUserTower.__init__ (tower_code/v1_usertower_train.py:21-60);
    side-info ids are MD5-hashed into [1, 1000] (:211-218, :910)."""
    return SimpleNamespace(d_model=d_model, max_len=max_len, dropout=0.2, pretrained_dim=128, nhead=4, num_layers=2,
                           num_items=num_items, num_prod_types=side, num_colors=side, num_graphics=side,
                           num_sections=side)


def zipf_ids(n, num_items, alpha, g):
    """ids in [1, num_items] with P(id=r) ~ r^-alpha (inverse-CDF sampling on a fixed table)."""
    w = torch.arange(1, num_items + 1, dtype=torch.float64).pow(-alpha)
    cdf = torch.cumsum(w / w.sum(), 0)
    u = torch.rand(n, generator=g, dtype=torch.float64)
    return (torch.searchsorted(cdf, u).clamp(max=num_items - 1) + 1).to(torch.int64)


def log_q(num_items, alpha=1.05):
    """log of normalised Zipf counts + 1e-6, [0] = -20 (tower_code/v1_refine_usertower.py:124-137)."""
    w = torch.arange(1, num_items + 1, dtype=torch.float64).pow(-alpha)
    q = torch.empty(num_items + 1, dtype=torch.float32)
    q[1:] = torch.log(w / w.sum() + 1e-6).float()
    q[0] = -20.0
    return q


def pretrained_table(num_items, dim=128, seed=42):
    g = torch.Generator().manual_seed(seed)
    t = F.normalize(torch.randn(num_items + 1, dim, generator=g), dim=1)
    t[0] = 0
    return t


def make_batch(B, L=50, num_items=N_ITEMS, seed=42, side=1000):
    """One training batch as SASRecDataset collates it (tower_code/v1_refine_usertower.py:204-306):
    left-padded sequences, next-item targets, static buckets, 4 continuous features."""
    g = torch.Generator().manual_seed(seed)
    lens = (1 + torch.empty(B).geometric_(1.0 / 12.0, generator=g)).clamp(max=L).to(torch.int64)
    pad = torch.arange(L).unsqueeze(0) < (L - lens).unsqueeze(1)               # True = padding (left)
    seq = zipf_ids(B * (L + 1), num_items, 1.05, g).view(B, L + 1)
    item_ids = seq[:, :L].masked_fill(pad, 0)
    target_ids = seq[:, 1:].masked_fill(pad, 0)                                  # shifted by one step

    def r(lo, hi, shape=(B, L)):
        return torch.randint(lo, hi + 1, shape, generator=g)

    batch = dict(
        item_ids=item_ids, target_ids=target_ids, padding_mask=pad,
        time_bucket_ids=r(1, 9).masked_fill(pad, 0),
        type_ids=r(1, side).masked_fill(pad, 0), color_ids=r(1, side).masked_fill(pad, 0),
        graphic_ids=r(1, side).masked_fill(pad, 0), section_ids=r(1, side).masked_fill(pad, 0),
        age_bucket=r(0, 10, (B,)), price_bucket=r(0, 10, (B,)), cnt_bucket=r(0, 10, (B,)),
        recency_bucket=r(0, 10, (B,)), channel_ids=r(0, 3, (B,)), club_status_ids=r(0, 3, (B,)),
        news_freq_ids=r(0, 2, (B,)), fn_ids=r(0, 2, (B,)), active_ids=r(0, 2, (B,)),
        cont_feats=torch.randn(B, 4, generator=g))
    return batch


FORWARD_KEYS = ("item_ids", "time_bucket_ids", "type_ids", "color_ids", "graphic_ids", "section_ids", "age_bucket",
                "price_bucket", "cnt_bucket", "recency_bucket", "channel_ids", "club_status_ids", "news_freq_ids",
                "fn_ids", "active_ids", "cont_feats", "padding_mask")


def criteo_vocab_sizes(seed=42):
    """Config 3 (builder-chosen): 26 categorical fields with log-uniform vocab in
    [10, 1e6] + 13 bucketised dense fields of 64 buckets."""
    g = torch.Generator().manual_seed(seed)
    cat = (10 ** (1 + 5 * torch.rand(26, generator=g))).to(torch.int64).tolist()
    return cat + [64] * 13


def make_fm_batch(B, vocab_sizes, seed=42):
    g = torch.Generator().manual_seed(seed)
    cols = [zipf_ids(B, v, 1.1, g) - 1 for v in vocab_sizes]
    return torch.stack(cols, dim=1)
