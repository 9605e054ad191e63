"""Oracle: whole-tower CPU restatements with the reference's parameter names.

TEST INFRASTRUCTURE -- see oracle/__init__.py for who may import this.

`UserTowerOracle` restates `SASRecUserTower` (tower_code/v1_refine_usertower.py:312-510)
and `ItemMatrixOracle` restates `SASRecItemTower` (tower_code/v1_usertower_train.py:266-293)
on top of oracle.embed, keeping the state-dict keys (SURVEY.md 8b "Parameter
naming") so reference / oracle / product checkpoints are interchangeable.
The transformer encoder and the MLPs are stock torch.nn modules in the
reference too (third-party ATen arithmetic), so they are instantiated, not
re-derived.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import embed

SEQ_TABLES = ("item_id_emb", "time_emb", "type_emb", "color_emb", "graphic_emb", "section_emb")
SEQ_INPUTS = ("item_ids", "time_bucket_ids", "type_ids", "color_ids", "graphic_ids", "section_ids")
SEQ_GATE_MASK = (1.0, 1.0, 0.0, 0.0, 0.0, 0.0)          # v1_refine_usertower.py:437
STATIC_TABLES = (("age_emb", 11, 16), ("price_emb", 11, 16), ("cnt_emb", 11, 16), ("recency_emb", 11, 16),
                 ("channel_emb", 4, 4), ("club_status_emb", 4, 4), ("news_freq_emb", 3, 4),
                 ("fn_emb", 3, 4), ("active_emb", 3, 4))     # :361-372
STATIC_INPUTS = ("age_bucket", "price_bucket", "cnt_bucket", "recency_bucket", "channel_ids",
                 "club_status_ids", "news_freq_ids", "fn_ids", "active_ids")


class UserTowerOracle(nn.Module):
    def __init__(self, args):
        super().__init__()
        d = args.d_model
        self.d_model, self.max_len = d, args.max_len
        self.item_proj = nn.Linear(args.pretrained_dim, d)
        sizes = {"item_id_emb": args.num_items + 1, "type_emb": args.num_prod_types + 1,
                 "color_emb": args.num_colors + 1, "graphic_emb": args.num_graphics + 1,
                 "section_emb": args.num_sections + 1, "time_emb": 12}
        for name in SEQ_TABLES:
            setattr(self, name, nn.Embedding(sizes[name], d, padding_idx=0))
        self.pos_emb = nn.Embedding(args.max_len, d)
        self.seq_gate = nn.Parameter(torch.ones(6))
        self.static_gate = nn.Parameter(torch.ones(10))
        self.emb_ln = nn.LayerNorm(d)
        self.emb_dropout = nn.Dropout(args.dropout)
        layer = nn.TransformerEncoderLayer(d_model=d, nhead=args.nhead, dim_feedforward=2 * d,
                                           dropout=args.dropout, activation="gelu",
                                           norm_first=True, batch_first=True)
        self.transformer_encoder = nn.TransformerEncoder(layer, num_layers=args.num_layers)
        for name, rows, dim in STATIC_TABLES:
            setattr(self, name, nn.Embedding(rows, dim, padding_idx=0))
        self.cont_proj = nn.Linear(4, 16)
        self.static_mlp = nn.Sequential(nn.Linear(100, d), nn.LayerNorm(d), nn.GELU(), nn.Dropout(args.dropout))
        self.output_proj = nn.Sequential(nn.Linear(2 * d, d), nn.LayerNorm(d), nn.GELU(), nn.Linear(d, d))

    def seq_gates(self):
        return torch.sigmoid(self.seq_gate) * torch.tensor(SEQ_GATE_MASK, device=self.seq_gate.device)

    def embed_front(self, pretrained_vecs, **seq_ids):
        """Pre-LayerNorm sequence embedding (v1_refine_usertower.py:447-456)."""
        return embed.seq_front(self.item_proj(pretrained_vecs),
                               [seq_ids[k] for k in SEQ_INPUTS],
                               [getattr(self, n).weight for n in SEQ_TABLES],
                               self.seq_gates(), self.pos_emb.weight)

    def static_front(self, cont_feats, **static_ids):
        return embed.static_front([static_ids[k] for k in STATIC_INPUTS],
                                  [getattr(self, n).weight for n, _, _ in STATIC_TABLES],
                                  cont_feats, self.cont_proj.weight, self.cont_proj.bias,
                                  torch.sigmoid(self.static_gate))

    def forward(self, pretrained_vecs, item_ids, time_bucket_ids, type_ids, color_ids, graphic_ids,
                section_ids, age_bucket, price_bucket, cnt_bucket, recency_bucket, channel_ids,
                club_status_ids, news_freq_ids, fn_ids, active_ids, cont_feats,
                padding_mask=None, training_mode=True):
        L = item_ids.size(1)
        x = self.embed_front(pretrained_vecs, item_ids=item_ids, time_bucket_ids=time_bucket_ids,
                             type_ids=type_ids, color_ids=color_ids, graphic_ids=graphic_ids,
                             section_ids=section_ids)
        x = self.emb_dropout(self.emb_ln(x))                                   # :458-459
        causal = torch.triu(torch.ones(L, L, dtype=torch.bool, device=item_ids.device), diagonal=1)    # :413-415
        h = self.transformer_encoder(x, mask=causal, src_key_padding_mask=padding_mask)
        prof = self.static_mlp(self.static_front(
            cont_feats, age_bucket=age_bucket, price_bucket=price_bucket, cnt_bucket=cnt_bucket,
            recency_bucket=recency_bucket, channel_ids=channel_ids, club_status_ids=club_status_ids,
            news_freq_ids=news_freq_ids, fn_ids=fn_ids, active_ids=active_ids))
        if training_mode:                                                      # :499-504
            fused = torch.cat([h, prof.unsqueeze(1).expand(-1, L, -1)], dim=-1)
        else:                                                                  # :506-510
            fused = torch.cat([h[:, -1, :], prof], dim=-1)
        return F.normalize(self.output_proj(fused), p=2, dim=-1)


class ItemMatrixOracle(nn.Module):
    def __init__(self, num_items, d_model, log_q_tensor=None):
        super().__init__()
        self.item_matrix = nn.Embedding(num_items + 1, d_model, padding_idx=0)
        self.register_buffer("log_q", log_q_tensor if log_q_tensor is not None
                             else torch.zeros(num_items + 1))

    def get_all_embeddings(self):
        return self.item_matrix.weight

    def get_log_q(self):
        return self.log_q


def all_timestep_step_loss(user_tower, item_tower, batch, lambda_logq=1.0, lambda_sup=0.1, lambda_cl=0.2):
    """Loss of ONE train step as `train_user_tower_all_time` builds it
    (tower_code/v1_usertower_train.py:787-845), fp32, no autocast:
    two forward views, C2 over all valid timesteps with batch-row user ids,
    C3 on the last valid step, total = main + lambda_cl * cl.
    `batch` holds the forward kwargs plus 'target_ids' and 'padding_mask'.
    Returns (total, main, cl)."""
    from . import losses
    kw = {k: v for k, v in batch.items() if k != "target_ids"}
    out1 = user_tower(**kw, training_mode=True)
    out2 = user_tower(**kw, training_mode=True)
    valid = ~batch["padding_mask"]
    B, L = batch["item_ids"].shape
    rows = torch.arange(B, device=valid.device).unsqueeze(1).expand(-1, L)
    u = F.normalize(out1[valid], p=2, dim=1)
    v_all = F.normalize(item_tower.get_all_embeddings(), p=2, dim=1)
    main = losses.inbatch_corrected_logq_loss(u, v_all, batch["target_ids"][valid], rows[valid],
                                              item_tower.get_log_q(), 0.1, lambda_logq)
    last = (valid.sum(dim=1) - 1).clamp(min=0)
    br = torch.arange(B, device=valid.device)
    cl = losses.duorec_loss_refined(out1[br, last], out2[br, last], batch["target_ids"][br, last],
                                    lambda_sup=lambda_sup)
    return main + lambda_cl * cl, main, cl
