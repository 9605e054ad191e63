"""N2: the on-disk contract between the ItemTower export and UserTower training, and the alignment of
the exported vectors to the training catalogue.

    export  utils/inference_utils.py:84-85,200-202   pretrained_item_matrix.pt  [N, 128] fp32, L2-normalised rows
                                                     item_ids.pt                list[str], sorted by str(product_id)
    import  tower_code/v1_usertower_train.py:131-160 load_aligned_pretrained_embeddings(processor, model_dir, dim)

The reference aligns with a Python dict over N ids and one tensor row copy per item.  Here the join is vectorised
(sort + binary search over the id strings on the host -- ids are strings, that is host work) and the row movement is
ONE device gather (`rs_gather_rows`, bit-exact rows); with `shard=(rank, world)` a rank materialises only the rows it
owns (row % world == rank).  Same results as the reference, including its quirks: unmatched rows keep `randn * 0.01`
drawn from torch's global CPU generator before anything else, row 0 is zero, and of a repeated exported id the LAST
row wins (dict semantics)."""
from __future__ import annotations

import os
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops

MATRIX_FILE = "pretrained_item_matrix.pt"
IDS_FILE = "item_ids.pt"


def save_item_vectors(final_tensor: torch.Tensor, ordered_ids: Sequence, save_dir: str = "models") -> Tuple[str, str]:
    """utils/inference_utils.py:195-202: the two files, under the reference's names."""
    os.makedirs(save_dir, exist_ok=True)
    mp, ip = os.path.join(save_dir, MATRIX_FILE), os.path.join(save_dir, IDS_FILE)
    torch.save(final_tensor.detach().float().cpu(), mp)
    torch.save([str(i) for i in ordered_ids], ip)
    return mp, ip


def _id_strings(ids) -> np.ndarray:
    """str(id) of every entry as the reference's dict keys spell it (:147: tensors through .item())."""
    if isinstance(ids, torch.Tensor):
        return np.asarray([str(v) for v in ids.tolist()], dtype=str)
    return np.asarray([str(i.item()) if isinstance(i, torch.Tensor) else str(i) for i in ids], dtype=str)


def id_join(pretrained_ids, item_ids) -> torch.Tensor:
    """src[i] = row of `pretrained_ids` whose id equals item_ids[i] (the LAST such row), -1 if none.  int64 [len(item_ids)].
    O((N + M) log N) in numpy's C loops instead of an N-entry Python dict and an M-step Python loop."""
    pre, cur = _id_strings(pretrained_ids), _id_strings(item_ids)
    if pre.size == 0 or cur.size == 0:
        return torch.full((cur.size,), -1, dtype=torch.int64)
    # unique over the REVERSED export gives, for every distinct id, its last occurrence
    uniq, first_rev = np.unique(pre[::-1], return_index=True)
    last = pre.size - 1 - first_rev
    pos = np.searchsorted(uniq, cur)
    pos_c = np.minimum(pos, uniq.size - 1)
    hit = uniq[pos_c] == cur
    return torch.from_numpy(np.where(hit, last[pos_c], -1).astype(np.int64))


def align_pretrained(pretrained: Optional[torch.Tensor], pretrained_ids, item_ids, pretrained_dim: int, device,
                     shard: Optional[Tuple[int, int]] = None) -> torch.Tensor:
    """The aligned [len(item_ids) + 1, dim] table on `device` (or its rows rank::world with `shard`)."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("rs_twotower ops run on CUDA tensors only (sm_100a); there is no CPU fallback")
    n = len(item_ids) + 1
    aligned = torch.randn(n, pretrained_dim) * 0.01          # the reference's draw, same generator, same order (:138)
    aligned[0] = 0.0
    rows = torch.arange(n) if shard is None else torch.arange(shard[0], n, shard[1])
    out = aligned[rows].to(device)
    if pretrained is None:
        return out
    if isinstance(pretrained, dict):
        pretrained = pretrained.get("weight", pretrained.get("item_content_emb.weight"))
    src = torch.cat([torch.full((1,), -1, dtype=torch.int64), id_join(pretrained_ids, item_ids)])[rows].to(device)
    hit = src >= 0
    got = ops.gather_rows(pretrained.to(device=device, dtype=torch.float32), src.clamp(min=0))
    return torch.where(hit.unsqueeze(1), got, out)


def load_aligned_pretrained_embeddings(processor, model_dir: str, pretrained_dim: int, device="cuda",
                                       shard: Optional[Tuple[int, int]] = None) -> torch.Tensor:
    """Drop-in for tower_code/v1_usertower_train.py:131-160 (`processor.item_ids`: the catalogue in training order).
    Missing / unreadable files fall back to the random initialisation, like the reference's try/except (:157-158)."""
    pretrained = ids = None
    try:
        pretrained = torch.load(os.path.join(model_dir, MATRIX_FILE), map_location="cpu")
        ids = torch.load(os.path.join(model_dir, IDS_FILE), map_location="cpu")
    except Exception as e:          # noqa: BLE001 -- mirrors the reference's blanket except
        print(f"[warning] failed to load pretrained files: {e}. Using random init.")
        pretrained = ids = None
    return align_pretrained(pretrained, ids, processor.item_ids, pretrained_dim, device, shard)
