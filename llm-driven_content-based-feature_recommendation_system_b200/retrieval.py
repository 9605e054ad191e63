"""R1: catalog retrieval.  `retrieve_topk` replaces the reference's
`scores = user @ items.T; topk(scores, k)` (tower_code/v1_usertower_train.py:672-675,
mined_inference.py:901-909,1103-1108,1536-1543, temp_model/ranker_skelet.py:193-196) with the
fused scoring + running top-k kernel; users are processed in chunks so that the partial lists
stay small."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import ops


def retrieve_topk(user_emb: torch.Tensor, item_emb: torch.Tensor, k: int, mask_index0: bool = False,
                  chunk: int = 262144):
    """-> (scores [b,k] fp32, ids [b,k] int64), sorted by score desc, ties by ascending id."""
    if user_emb.shape[0] <= chunk:
        return ops.retrieve_topk(user_emb, item_emb, k, mask_index0)
    ss, ii = [], []
    for lo in range(0, user_emb.shape[0], chunk):
        s, i = ops.retrieve_topk(user_emb[lo:lo + chunk], item_emb, k, mask_index0)
        ss.append(s)
        ii.append(i)
    return torch.cat(ss), torch.cat(ii)


def recall_candidates(user_tower_out: torch.Tensor, item_tower, k: int):
    """evaluate_model's retrieval step (tower_code/v1_usertower_train.py:566-567,672-675)."""
    items = F.normalize(item_tower.get_all_embeddings(), p=2, dim=1)
    return retrieve_topk(user_tower_out.float(), items, k)
