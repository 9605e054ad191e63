"""Drop-in towers: same constructor arguments, forward signatures, parameter names and shapes as the
reference's nn.Modules (so its .pth checkpoints load with strict=True), with the embedding fronts
running on the hand-written sm_100a kernels.  Everything that is stock torch.nn in the reference
(nn.TransformerEncoder, the MLPs, HF BERT) stays stock torch.nn here -- it is third-party ATen
arithmetic on both sides and outside the hot path.

    SASRecUserTower      tower_code/v1_refine_usertower.py:312-510
    SASRecItemTower      tower_code/v1_usertower_train.py:266-293
    HybridItemTower      item_tower.py:131-286   (fronts I1/I2; encoder body stock)
    OptimizedItemTower   item_tower.py:289-305
    SimCSEModelWrapper   item_tower.py:308-323
    HybridUserEmbeddings tower_code/mined_inference.py:614-616,634,643 + forward gathers :670,687-688,695,705
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import encoder as enc
from . import ops

SEQ_TABLES = ("item_id_emb", "time_emb", "type_emb", "color_emb", "graphic_emb", "section_emb")
SEQ_GATE_MASK = (1.0, 1.0, 0.0, 0.0, 0.0, 0.0)       # s_mask, v1_refine_usertower.py:437
STATIC_TABLES = ("age_emb", "price_emb", "cnt_emb", "recency_emb", "channel_emb", "club_status_emb",
                 "news_freq_emb", "fn_emb", "active_emb")


import os as _os
FUSE_EMB_LN = _os.environ.get('RS_FUSE_EMB_LN', '1') == '1'      # (A/B switch for encoder.emb_layer_norm2)


class _ZeroGradTables(torch.autograd.Function):
    """identity on x; the listed tables get explicit all-zero gradients (they take part with a gate of exactly 0)."""

    @staticmethod
    def forward(ctx, x, *tables):
        ctx.save_for_backward(*tables)
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return (g,) + tuple(ops.zeros_many([tuple(t.shape) for t in ctx.saved_tensors], g.device))


class SASRecUserTower(nn.Module):
    def __init__(self, args):
        super().__init__()
        d = self.d_model = args.d_model
        self.max_len = args.max_len
        self.dropout_rate = args.dropout
        self.item_proj = nn.Linear(args.pretrained_dim, d)
        self.item_id_emb = nn.Embedding(args.num_items + 1, d, padding_idx=0)
        self.type_emb = nn.Embedding(args.num_prod_types + 1, d, padding_idx=0)
        self.color_emb = nn.Embedding(args.num_colors + 1, d, padding_idx=0)
        self.graphic_emb = nn.Embedding(args.num_graphics + 1, d, padding_idx=0)
        self.section_emb = nn.Embedding(args.num_sections + 1, d, padding_idx=0)
        self.pos_emb = nn.Embedding(self.max_len, d)
        self.seq_gate = nn.Parameter(torch.ones(6))
        self.static_gate = nn.Parameter(torch.ones(10))
        self.time_emb = nn.Embedding(12, d, padding_idx=0)
        self.emb_ln = nn.LayerNorm(d)
        self.emb_dropout = nn.Dropout(self.dropout_rate)
        layer = nn.TransformerEncoderLayer(d_model=d, nhead=args.nhead, dim_feedforward=2 * d,
                                           dropout=self.dropout_rate, activation="gelu", norm_first=True,
                                           batch_first=True)
        self.transformer_encoder = nn.TransformerEncoder(layer, num_layers=args.num_layers)
        for name, rows, dim in (("age_emb", 11, 16), ("price_emb", 11, 16), ("cnt_emb", 11, 16),
                                ("recency_emb", 11, 16), ("channel_emb", 4, 4), ("club_status_emb", 4, 4),
                                ("news_freq_emb", 3, 4), ("fn_emb", 3, 4), ("active_emb", 3, 4)):
            setattr(self, name, nn.Embedding(rows, dim, padding_idx=0))
        self.num_cont_feats = 4
        self.cont_proj = nn.Linear(4, 16)
        self.static_mlp = nn.Sequential(nn.Linear(100, d), nn.LayerNorm(d), nn.GELU(), nn.Dropout(self.dropout_rate))
        self.output_proj = nn.Sequential(nn.Linear(2 * d, d), nn.LayerNorm(d), nn.GELU(), nn.Linear(d, d))
        self.register_buffer("_seq_gate_mask", torch.tensor(SEQ_GATE_MASK), persistent=False)
        self.apply(self._init_weights)

    @staticmethod
    def _init_weights(m):
        # same initialisation as the reference (:399-410); note it overwrites the padding rows (invariant 1)
        if isinstance(m, nn.Linear):
            nn.init.kaiming_normal_(m.weight, mode="fan_in", nonlinearity="relu")
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.Embedding):
            nn.init.normal_(m.weight, mean=0.0, std=0.02)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def get_causal_mask(self, seq_len, device):
        return torch.triu(torch.ones(seq_len, seq_len, device=device, dtype=torch.bool), diagonal=1)

    def embed_front(self, pretrained_vecs, item_ids, time_bucket_ids, type_ids, color_ids, graphic_ids, section_ids,
                    item_id_rows=None):
        """U1 (:447-456): one fused kernel instead of 6 gathers + 13 elementwise passes.  Tables whose gate
        is hard-masked to zero (type/color/graphic/section) contribute exactly 0 and are not read.
        `item_id_rows` [1 + B*L, 128]: the rows `item_id_emb.weight[item_ids]` already fetched from their owner
        ranks (row-sharded table, sharded.py) behind one unused leading row; the kernel then reads row 1 + b*L + l
        for position (b, l) -- same arithmetic, and the gradient of the buffer goes back through the exchange."""
        s_g = torch.sigmoid(self.seq_gate) * self._seq_gate_mask
        base = enc.linear(self.item_proj, pretrained_vecs)
        ids = [item_ids, time_bucket_ids, type_ids, color_ids, graphic_ids, section_ids]
        tables = [getattr(self, n).weight for n in SEQ_TABLES]
        if item_id_rows is not None:
            ids[0] = torch.arange(1, item_ids.numel() + 1, device=item_ids.device).view_as(item_ids)
            tables[0] = item_id_rows
        n_live = sum(1 for m in SEQ_GATE_MASK if m != 0.0)        # live tables come first in SEQ_TABLES
        return ops.seq_front(base, ids, tables, s_g, self.pos_emb.weight, padding_idx=0, n_live=n_live)

    def embed_front_packed(self, pretrained_vecs, item_ids, time_bucket_ids, pos_ids, item_id_rows=None):
        """U1 on PACKED tokens: the same kernel and the same order of additions, with the position embedding passed
        as a third gathered table (gate 1.0, ids = l + 1 behind a zero row so that the padding convention of the
        other tables -- id 0 never receives gradient -- leaves position 0 alone).  `item_ids` / `time_bucket_ids` /
        `pos_ids` are [R, 64] grids holding the batch's valid tokens (train.add_host_index), zero-padded at the end."""
        s_g = torch.sigmoid(self.seq_gate) * self._seq_gate_mask
        base = enc.linear(self.item_proj, pretrained_vecs)
        w = self.pos_emb.weight
        pos_ext = torch.cat([w.new_zeros(1, w.shape[1]), w])
        gates = torch.cat([s_g[:2], s_g.new_ones(1)])
        ids = [item_ids, time_bucket_ids, pos_ids]
        tables = [self.item_id_emb.weight, self.time_emb.weight, pos_ext]
        if isinstance(item_id_rows, (tuple, list)):
            # (buffer, slots): the DISTINCT item rows of the batch fetched from their owners (sharded.dedup_lookup) and
            # the buffer slot of every token; slot 0 is the padding id's row (train.ShardedDeviceStep forces it there), so
            # the kernels' "id 0 never receives gradient" rule keeps its meaning
            tables[0], ids[0] = item_id_rows[0], item_id_rows[1].view_as(item_ids)
        elif item_id_rows is not None:
            ids[0] = torch.arange(1, item_ids.numel() + 1, device=item_ids.device).view_as(item_ids)
            tables[0] = item_id_rows
        out = ops.seq_front(base, ids, tables, gates, None, padding_idx=0, n_live=3)
        # the hard-masked tables contribute exactly 0 and receive exactly-0 gradients in the reference (invariant 2)
        return _ZeroGradTables.apply(out, *[getattr(self, n).weight for n in SEQ_TABLES[2:]])

    def static_front(self, age_bucket, price_bucket, cnt_bucket, recency_bucket, channel_ids, club_status_ids,
                     news_freq_ids, fn_ids, active_ids, cont_feats):
        """U2 (:472-491): one kernel instead of ~25."""
        u_g = torch.sigmoid(self.static_gate)
        ids = [age_bucket, price_bucket, cnt_bucket, recency_bucket, channel_ids, club_status_ids, news_freq_ids,
               fn_ids, active_ids]
        tables = [getattr(self, n).weight for n in STATIC_TABLES]
        return ops.static_front(ids, tables, cont_feats.float(), self.cont_proj.weight, self.cont_proj.bias, u_g,
                                padding_idx=0)

    def forward(self, pretrained_vecs, item_ids, time_bucket_ids, type_ids, color_ids, graphic_ids, section_ids,
                age_bucket, price_bucket, cnt_bucket, recency_bucket, channel_ids, club_status_ids, news_freq_ids,
                fn_ids, active_ids, cont_feats, padding_mask=None, training_mode=True, select_index=None,
                item_id_rows=None, packed_index=None, cu_seqlens=None, packed_zero_tail=0, views=1, packed_fold=None,
                select_users=None, packed_inputs=None, packed_fold_inv=None, select_prefix=None):
        """Reference signature (:417-429) plus one optional extension: `select_index` (flat b*L+l positions).
        When given (training_mode only) the late-fusion head runs on those rows alone and [len(index),128] is
        returned -- the train step only ever consumes the valid / last time steps (v1_usertower_train.py:794-842),
        so the other ~75 % of the [B,L] grid need not go through output_proj."""
        seq_len = item_ids.size(1)
        if packed_inputs is not None:      # U1 on the packed tokens too (pretrained_vecs is gathered for that grid)
            seq_emb = self.embed_front_packed(pretrained_vecs, packed_inputs["item_ids"],
                                              packed_inputs["time_bucket_ids"], packed_inputs["pos_ids"], item_id_rows)
        else:
            seq_emb = self.embed_front(pretrained_vecs, item_ids, time_bucket_ids, type_ids, color_ids, graphic_ids,
                                       section_ids, item_id_rows)
        static_input = self.static_front(age_bucket, price_bucket, cnt_bucket, recency_bucket, channel_ids,
                                         club_status_ids, news_freq_ids, fn_ids, active_ids, cont_feats)
        if packed_index is not None:
            # `views` dropout views in ONE pass: the deterministic fronts are computed once, the index / offsets carry
            # every token `views` times (train.add_host_index), and each copy draws its own dropout masks
            user_profile_vec = enc.sequential(self.static_mlp, static_input if views == 1 else static_input.repeat(views, 1))
            return self._forward_packed(seq_emb, user_profile_vec, seq_len, training_mode, select_index, packed_index,
                                        cu_seqlens, packed_zero_tail, select_users, packed_fold, packed_fold_inv,
                                        select_prefix, views)
        user_profile_vec = self.static_mlp(static_input)
        seq_emb = self.emb_dropout(self.emb_ln(seq_emb))
        # is_causal=True only tells nn.TransformerEncoder not to PROBE the mask: with is_causal=None it compares the
        # mask with a generated causal one and reads the verdict back (`bool((mask == causal).all())`,
        # torch/nn/modules/transformer.py:_detect_is_causal_mask) -- one device->host synchronisation per forward.
        # The arithmetic is unchanged: with a key-padding mask present, multi_head_attention_forward merges both
        # masks and drops the hint (torch/nn/functional.py), exactly the path the reference's call takes.
        output = self.transformer_encoder(seq_emb, mask=self.get_causal_mask(seq_len, item_ids.device),
                                          src_key_padding_mask=padding_mask,
                                          is_causal=True if padding_mask is not None else None)
        if training_mode and select_index is not None:
            rows = ops.gather_rows(output.reshape(-1, output.shape[-1]), select_index)
            prof = ops.gather_rows(user_profile_vec, select_index // seq_len)
            final_vec = self.output_proj(torch.cat([rows, prof.to(rows.dtype)], dim=-1))
        elif training_mode:
            expanded = user_profile_vec.unsqueeze(1).expand(-1, seq_len, -1)
            final_vec = self.output_proj(torch.cat([output, expanded], dim=-1))
        else:
            final_vec = self.output_proj(torch.cat([output[:, -1, :], user_profile_vec], dim=-1))
        return F.normalize(final_vec, p=2, dim=-1)

    def _forward_packed(self, seq_emb, user_profile_vec, seq_len, training_mode, select_index, packed_index, cu_seqlens,
                        zero_tail=0, select_users=None, packed_fold=None, packed_fold_inv=None, select_prefix=None,
                        views=1):
        """The encoder on the packed valid tokens (encoder.py): `packed_index` [T] = flat b*L+l positions of the
        valid time steps in batch-major order, `cu_seqlens` int32 their per-sequence offsets (every sequence
        non-empty).  `select_index` then indexes PACKED rows; returns [len(select_index), 128] (all T rows when
        None) in training mode, the last valid step of every sequence [B, 128] otherwise -- the values the reference
        computes at those positions of its padded grid.  The last `zero_tail` entries of `cu_seqlens` may be
        one-token pseudo-sequences at PADDED positions (attention output 0 there, train.add_host_index).
        `select_users` [len(select_index)]: the row of `user_profile_vec` that goes with every selected row (needed
        when several dropout views share the pass; default: the token's own batch row)."""
        tr = self.training
        e = seq_emb.reshape(-1, seq_emb.shape[-1])
        first_h = None
        if (FUSE_EMB_LN and packed_fold_inv is not None and packed_index is not None
                and packed_index.numel() == 2 * e.shape[0] and enc.first_layer_fusable(self.transformer_encoder, self.emb_ln, e)):
            # embedding LayerNorm + dropout and the first layer's LayerNorm in one pass each way; the backward folds the two
            # dropout views before the embedding LayerNorm's backward (one warp per U1 row)
            ad = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else torch.float32
            x, first_h = enc.emb_layer_norm2(e, packed_index, packed_fold_inv, self.emb_ln,
                                             self.emb_dropout.p if tr else 0.0, self.transformer_encoder.layers[0].norm1, ad)
        else:
            x = enc.layer_norm(e, self.emb_ln.weight, self.emb_ln.bias, self.emb_ln.eps, index=packed_index,
                               dropout_p=self.emb_dropout.p if tr else 0.0, out_dtype=torch.float32,
                               index_fold=packed_fold, index_inv=packed_fold_inv)
        if select_prefix is not None:
            # device-built index (ops.batch_index_build): the first `select_prefix` selected rows ARE the first packed
            # rows (identity), their users ascend (batch-major), and the remaining selected rows are one DuoRec row per
            # (view, user) in the order of `user_profile_vec` -- no gather for the main rows, no sort in any backward.
            # Only those rows are read from the encoder, so its last layer runs its position-wise half on them alone
            # (half the tokens: the second dropout view feeds nothing but its DuoRec rows).  A DuoRec row that lies
            # inside the prefix is NOT computed a second time (it would draw other dropout masks than the main-loss row
            # it is, v1_usertower_train.py:788-842): it is picked from the prefix afterwards.
            n = int(select_prefix)
            tail = select_index[n:]
            m = tail.numel()
            inside = tail < n
            # With two views the sequences of the second one (index B on) feed only their DuoRec rows, the second half of
            # `tail` (one row per user, some token in the middle of the sequence -- or a zero-tail row: none); the prefix
            # rows that belong to them are padding rows of weight 0: single-row attention for them in the last layer.
            n_users = user_profile_vec.shape[0] // max(views, 1)
            one = dict(one_row_from=n_users, one_rows=tail[m - n_users:].contiguous()) if views == 2 and m == 2 * n_users else {}
            output = enc.packed_encoder(self.transformer_encoder, x, cu_seqlens, seq_len, zero_tail,
                                        last_rows=(n, torch.where(inside, -1, tail)), first_h=first_h, **one)
            pick = torch.where(inside, tail, n + torch.arange(m, device=tail.device))
            # the head's first Linear autocasts its input: emit the rows in that dtype right away
            ad = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else output.dtype
            output = ops.select_prefix_rows(output, n, pick, out_dtype=ad)
            # Linear(256 -> 128) -> LayerNorm -> GELU without the [rows, 256] concatenation (encoder.fused_head): the
            # profile half of the Linear runs once per (view, user) row of `user_profile_vec`, not once per time step
            op = self.output_proj
            if (len(op) == 4 and isinstance(op[0], nn.Linear) and op[0].in_features == 256 and op[0].out_features == 128
                    and op[0].bias is not None and isinstance(op[1], nn.LayerNorm) and isinstance(op[2], nn.GELU)
                    and op[2].approximate == "none" and output.dtype in (torch.bfloat16, torch.float16)):
                h = enc.fused_head(output, user_profile_vec, select_users, n, op[0], op[1])
                return enc.l2_normalize(enc.linear(op[3], h))
            prof = torch.cat([ops.gather_rows_sorted(user_profile_vec, select_users[:n], out_dtype=ad),
                              user_profile_vec.to(ad)])
            final_vec = enc.sequential(self.output_proj, torch.cat([output, prof], dim=-1))
            return enc.l2_normalize(final_vec)
        output = enc.packed_encoder(self.transformer_encoder, x, cu_seqlens, seq_len, zero_tail, first_h=first_h)
        if not training_mode:
            select_index = cu_seqlens[1:user_profile_vec.shape[0] + 1].to(torch.int64) - 1
        users = select_users
        if users is None:
            users = packed_index // seq_len
            if select_index is not None:
                users = users[select_index]
        if select_index is not None:
            output = ops.gather_rows(output, select_index)
        prof = ops.gather_rows(user_profile_vec, users)
        final_vec = enc.sequential(self.output_proj, torch.cat([output, prof.to(output.dtype)], dim=-1))
        return enc.l2_normalize(final_vec)


class SASRecItemTower(nn.Module):
    def __init__(self, num_items, d_model, log_q_tensor=None):
        super().__init__()
        self.item_matrix = nn.Embedding(num_items + 1, d_model, padding_idx=0)
        self.register_buffer("log_q", log_q_tensor if log_q_tensor is not None else torch.zeros(num_items + 1))

    def get_all_embeddings(self):
        return self.item_matrix.weight

    def get_log_q(self):
        return self.log_q

    def set_freeze_state(self, freeze: bool):
        for p in self.parameters():
            p.requires_grad = not freeze

    def init_from_pretrained(self, pretrained_vecs):
        with torch.no_grad():
            self.item_matrix.weight.copy_(pretrained_vecs)

    def normalized_rows(self, target_ids, out_dtype=None):
        """`F.normalize(self.item_matrix.weight, p=2, dim=1)[target_ids]` (the loop's lines
        tower_code/v1_usertower_train.py:810-811 + the gather at v1_refine_usertower.py:833) without
        normalising -- or differentiating through -- the rows that are not in the batch."""
        return ops.normalized_rows(self.item_matrix.weight, target_ids, 1e-12, out_dtype)


# ---------------------------------------------------------------------------------------------------
class _SEBlock(nn.Module):                      # item_tower.py:41-76 (stock layers; parameter names kept)
    def __init__(self, dim, dropout=0.2, expansion_factor=4):
        super().__init__()
        h = dim * expansion_factor
        self.block = nn.Sequential(nn.Linear(dim, h), nn.LayerNorm(h), nn.GELU(), nn.Dropout(dropout),
                                   nn.Linear(h, dim), nn.LayerNorm(dim))
        self.se_block = nn.Sequential(nn.Linear(dim, dim // 4), nn.ReLU(), nn.Linear(dim // 4, dim), nn.Sigmoid())

    def forward(self, x):
        y = self.block(x)
        return x + y * self.se_block(y)


class DeepResidualHead(nn.Module):              # item_tower.py:78-128
    def __init__(self, input_dim, output_dim=128):
        super().__init__()
        mid, hid = input_dim * 2, input_dim * 4
        self.expand_layer1 = nn.Sequential(nn.Linear(input_dim, mid), nn.LayerNorm(mid), nn.GELU(), nn.Dropout(0.1))
        self.expand_layer2 = nn.Sequential(nn.Linear(mid, hid), nn.LayerNorm(hid), nn.GELU(), nn.Dropout(0.1))
        self.res_blocks = nn.Sequential(_SEBlock(hid, dropout=0.2), _SEBlock(hid, dropout=0.2))
        self.final_proj = nn.Linear(hid, output_dim)
        self.input_skip = nn.Linear(input_dim, output_dim)

    def forward(self, x):
        return self.final_proj(self.res_blocks(self.expand_layer2(self.expand_layer1(x)))) + self.input_skip(x)


class HybridItemTower(nn.Module):
    """item_tower.py:131-286.  `bert_model` may be injected (any HF BertModel); by default it is loaded
    like the reference does (`AutoModel.from_pretrained("bert-base-uncased")`, needs the HF cache)."""

    def __init__(self, std_vocab_size: int, num_std_fields: int, embed_dim: int = 128, output_dim: int = 128,
                 bert_model=None, pad_id: int = 0):
        super().__init__()
        self.std_embedding = nn.Embedding(std_vocab_size, embed_dim, padding_idx=pad_id)
        self.std_field_emb = nn.Parameter(torch.randn(1, num_std_fields, embed_dim))
        self.std_ln = nn.LayerNorm(embed_dim)
        self.re_ln = nn.LayerNorm(embed_dim)
        if bert_model is None:
            from transformers import AutoModel
            bert_model = AutoModel.from_pretrained("bert-base-uncased")
        self.bert_model = bert_model
        self.bert_config = bert_model.config
        bert_dim = self.bert_config.hidden_size
        self.re_proj = nn.Sequential(nn.Linear(bert_dim, embed_dim), nn.LayerNorm(embed_dim), nn.GELU())
        self.re_field_position = nn.Parameter(torch.randn(1, 9, embed_dim))
        self.text_proj = nn.Sequential(nn.Linear(bert_dim, embed_dim), nn.LayerNorm(embed_dim), nn.GELU())
        layer = nn.TransformerEncoderLayer(d_model=embed_dim, nhead=4, dim_feedforward=embed_dim * 4,
                                           batch_first=True, dropout=0.1, activation="gelu", norm_first=True)
        self.transformer = nn.TransformerEncoder(layer, num_layers=2, enable_nested_tensor=False)
        self.head = DeepResidualHead(input_dim=embed_dim, output_dim=output_dim)
        self._pad_id = pad_id

    # I1 -- item_tower.py:239-241
    def std_front(self, std_input):
        if torch.is_grad_enabled() and (self.std_embedding.weight.requires_grad or self.std_field_emb.requires_grad):
            x = ops.gather_rows(self.std_embedding.weight, std_input, padding_idx=self._pad_id) + self.std_field_emb
            return self.std_ln(x)
        return torch.ops.rs.std_front(self.std_embedding.weight, std_input, self.std_field_emb, self.std_ln.weight,
                                      self.std_ln.bias, self.std_ln.eps, 0)

    # I2 -- item_tower.py:246-261
    def bert_word_embeddings(self, flat_ids):
        """`self.bert_model.embeddings(input_ids=...)` under no_grad (:248-249), fused gather+add+LN(+dropout)."""
        e = self.bert_model.embeddings
        p = e.dropout.p if (e.training and e.dropout.p > 0) else 0.0
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if p > 0 else 0
        with torch.no_grad():
            return torch.ops.rs.bert_embed(e.word_embeddings.weight, e.position_embeddings.weight,
                                           e.token_type_embeddings.weight, e.LayerNorm.weight, e.LayerNorm.bias,
                                           e.LayerNorm.eps, flat_ids, p, seed, 0)

    def re_front(self, re_input_ids, re_attn_mask):
        B = re_input_ids.shape[0]
        T = re_input_ids.size(-1)
        word_embs = self.bert_word_embeddings(re_input_ids.reshape(-1, T))
        re_feats = self.re_proj(word_embs)
        pooled = ops.masked_mean(re_feats, re_attn_mask.reshape(-1, T))
        re_vectors = pooled.view(B, 9, -1).to(re_feats.dtype) + self.re_field_position
        return self.re_ln(re_vectors)

    def forward(self, std_input, re_input_ids, re_attn_mask, text_input_ids, text_attn_mask):
        std_emb = self.std_front(std_input)
        re_vectors = self.re_front(re_input_ids, re_attn_mask)
        bert_out = self.bert_model(input_ids=text_input_ids, attention_mask=text_attn_mask)
        text_vec = self.text_proj(bert_out.last_hidden_state[:, 0, :]).unsqueeze(1)
        combined = torch.cat([std_emb.to(text_vec.dtype), re_vectors.to(text_vec.dtype), text_vec], dim=1)
        out = self.head(self.transformer(combined).mean(dim=1))
        return F.normalize(out, p=2, dim=1)


class OptimizedItemTower(nn.Module):            # item_tower.py:289-305
    def __init__(self, input_dim=128, output_dim=128):
        super().__init__()
        self.layer = nn.Sequential(nn.Linear(input_dim, input_dim), nn.LayerNorm(input_dim), nn.GELU(),
                                   nn.Linear(input_dim, output_dim))

    def forward(self, x):
        return F.normalize(self.layer(x), p=2, dim=1)


class SimCSEModelWrapper(nn.Module):            # item_tower.py:308-323
    def __init__(self, encoder: nn.Module, projector: nn.Module):
        super().__init__()
        self.encoder = encoder
        self.projector = projector

    def forward(self, std, re_ids, re_mask, txt_ids, txt_mask):
        return self.projector(self.encoder(std, re_ids, re_mask, txt_ids, txt_mask))


class HybridUserEmbeddings(nn.Module):
    """The embedding tables of `HybridUserTower` (tower_code/mined_inference.py:614-616,634,643) under
    their reference names, and its forward-pass gathers (:670,687-688,695,705) -- row H1 of SURVEY.md 8a.
    None of these tables has a padding_idx: row 0 is a zero row that DOES receive gradient."""

    def __init__(self, gnn_user_init, gnn_item_init, item_content_init):
        super().__init__()
        self.gnn_user_emb = nn.Embedding.from_pretrained(gnn_user_init, freeze=False)
        self.gnn_item_emb = nn.Embedding.from_pretrained(gnn_item_init, freeze=False)
        self.item_content_emb = nn.Embedding.from_pretrained(item_content_init, freeze=False)
        self.time_emb = nn.Embedding(1001, 128)
        self.channel_emb = nn.Embedding(2, 32)

    def forward(self, u_idx, seq_ids, seq_deltas, u_cat):
        return dict(
            gnn_user_emb=ops.gather_rows(self.gnn_user_emb.weight, u_idx),
            item_content_emb=ops.gather_rows(self.item_content_emb.weight, seq_ids),
            gnn_item_emb=ops.gather_rows(self.gnn_item_emb.weight, seq_ids),
            time_emb=ops.gather_rows(self.time_emb.weight, seq_deltas, clamp_max=1000),
            channel_emb=ops.gather_rows(self.channel_emb.weight, u_cat))
