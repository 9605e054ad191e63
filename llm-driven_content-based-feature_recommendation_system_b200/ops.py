"""torch.library custom ops (`torch.ops.rs.*`) over the C ABI, plus their autograd.

Every op is a thin marshalling layer: it allocates outputs with torch, passes raw
device pointers + sizes + the current CUDA stream to librs_twotower.so, and
returns.  No arithmetic happens in Python and nothing here touches the CPU
oracle.  CPU tensors are rejected (no fallback).
"""
from __future__ import annotations

import collections
from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _lib as L

_lib = L.load()


def _c(t: Optional[Tensor]) -> Optional[Tensor]:
    return None if t is None else t.contiguous()


def _ids(t: Tensor) -> Tensor:
    if t.dtype != torch.int64:
        t = t.to(torch.int64)
    return t.contiguous()


def _f32(t: Tensor, what: str) -> Tensor:
    if t.dtype != torch.float32:
        raise TypeError(f"{what} must be float32 (master weights), got {t.dtype}")
    return t.contiguous()


# ---------------------------------------------------------------------------------------------
# primitive ops
# ---------------------------------------------------------------------------------------------
@torch.library.custom_op("rs::gather_rows", mutates_args=())
def gather_rows_op(table: Tensor, ids: Tensor, clamp_max: int, out_dtype: int) -> Tensor:
    L.require_cuda(table, ids)
    table, ids = _c(table), _ids(ids)
    rows, dim = table.shape
    out = torch.empty(*ids.shape, dim, dtype=L.torch_dtype(out_dtype), device=table.device)
    L.check(_lib.rs_gather_rows(L.ptr(table), L.dt(table), rows, dim, L.ptr(ids), ids.numel(), clamp_max,
                                L.ptr(out), out_dtype, L.ptr(L.oob_flag(table.device)), L.stream()), "rs_gather_rows")
    return out


@gather_rows_op.register_fake
def _(table, ids, clamp_max, out_dtype):
    return table.new_empty(*ids.shape, table.shape[1], dtype=L.torch_dtype(out_dtype))


_SortEntry = collections.namedtuple("_SortEntry", "ids version rows clamp skeys spos")
_sort_cache: "collections.OrderedDict[tuple, _SortEntry]" = collections.OrderedDict()


def sorted_ids(ids: Tensor, rows: int, clamp_max: int = -1) -> Tuple[Tensor, Tensor]:
    """Stable (id, position) sort, cached per ids tensor (the two dropout views of a step and the
    forward/backward of one table share the same ids)."""
    # keyed by storage address + version counter (custom-op dispatch hands us fresh wrappers of the same
    # tensor, so object identity is useless); the entry keeps a reference, so the address cannot be recycled
    key = (ids.data_ptr(), ids._version, ids.numel(), rows, clamp_max)
    e = _sort_cache.get(key)
    if e is not None:
        _sort_cache.move_to_end(key)
        return e.skeys, e.spos
    n = ids.numel()
    sk = torch.empty(n, dtype=torch.int32, device=ids.device)
    sp = torch.empty(n, dtype=torch.int32, device=ids.device)
    wsb = _lib.rs_sort_ids_workspace_bytes(n)
    ws = L.workspace(wsb, ids.device)
    L.check(_lib.rs_sort_ids(L.ptr(ids), n, rows, clamp_max, L.ptr(sk), L.ptr(sp), L.ptr(ws), ws.numel(),
                             L.ptr(L.oob_flag(ids.device)), L.stream()), "rs_sort_ids")
    _sort_cache[key] = _SortEntry(ids, ids._version, rows, clamp_max, sk, sp)
    while len(_sort_cache) > 32:
        _sort_cache.popitem(last=False)
    return sk, sp


@torch.library.custom_op("rs::embedding_dense_bwd", mutates_args=())
def embedding_dense_bwd_op(grad: Tensor, ids: Tensor, rows: int, padding_idx: int, clamp_max: int,
                           deterministic: bool) -> Tensor:
    """Dense [rows, dim] fp32 gradient of gather_rows (== aten::embedding_dense_backward)."""
    L.require_cuda(grad, ids)
    ids = _ids(ids)
    dim = grad.shape[-1]
    grad = grad.reshape(-1, dim).contiguous()
    d_table = torch.zeros(rows, dim, dtype=torch.float32, device=grad.device)
    n = ids.numel()
    if n == 0:
        return d_table
    if deterministic:
        sk, sp = sorted_ids(ids, rows, clamp_max)
        wsb = _lib.rs_segment_reduce_workspace_bytes(n, dim)
        ws = L.workspace(wsb, grad.device)
        L.check(_lib.rs_segment_reduce_rows(L.ptr(grad), L.dt(grad), L.ptr(sk), L.ptr(sp), n, dim, rows, padding_idx,
                                            None, None, L.ptr(d_table), None, L.ptr(ws), ws.numel(), L.stream()),
                "rs_segment_reduce_rows")
    else:
        L.check(_lib.rs_scatter_add_rows(L.ptr(grad), L.dt(grad), L.ptr(ids), n, dim, rows, padding_idx, clamp_max,
                                         1.0, L.ptr(d_table), L.ptr(L.oob_flag(grad.device)), L.stream()),
                "rs_scatter_add_rows")
    return d_table


@embedding_dense_bwd_op.register_fake
def _(grad, ids, rows, padding_idx, clamp_max, deterministic):
    return grad.new_empty(rows, grad.shape[-1], dtype=torch.float32)


@torch.library.custom_op("rs::seq_front", mutates_args=())
def seq_front_op(base: Optional[Tensor], ids: Sequence[Tensor], tables: Sequence[Tensor], gates: Tensor,
                 pos_table: Optional[Tensor], seq_len: int, out_dtype: int) -> Tensor:
    L.require_cuda(gates, *ids, *tables)
    n = len(tables)
    ids = [_ids(i) for i in ids]
    tables = [_f32(t, "embedding table") for t in tables]
    base = _c(base)
    pos_table = None if pos_table is None else _f32(pos_table, "pos table")
    if pos_table is not None and pos_table.shape[0] < seq_len:
        # the reference's pos_emb(arange(seq_len)) raises here too (the C ABI carries no row count for this table)
        raise IndexError(f"index out of range in self: sequence length {seq_len} exceeds the {pos_table.shape[0]} rows "
                         f"of the position table")
    dim = tables[0].shape[1] if n else (base.shape[-1] if base is not None else pos_table.shape[1])
    shape = ids[0].shape if n else base.shape[:-1]
    P = 1
    for s in shape:
        P *= s
    out = torch.empty(*shape, dim, dtype=L.torch_dtype(out_dtype), device=gates.device)
    gates = _f32(gates, "gates")
    L.check(_lib.rs_seq_front_fwd(L.ptr(base), L.dt(base) if base is not None else 0, L.ptr_array(ids),
                                  L.ptr_array(tables), L.i64_array([t.shape[0] for t in tables]), n, L.ptr(gates),
                                  L.ptr(pos_table), seq_len, P, dim, L.ptr(out), out_dtype,
                                  L.ptr(L.oob_flag(gates.device)), L.stream()), "rs_seq_front_fwd")
    return out


@seq_front_op.register_fake
def _(base, ids, tables, gates, pos_table, seq_len, out_dtype):
    dim = tables[0].shape[1]
    return tables[0].new_empty(*ids[0].shape, dim, dtype=L.torch_dtype(out_dtype))


# a table this small is reduced in per-warp private shared-memory copies (deterministic, no atomics)
SMALL_TABLE_ROWS = 24


@torch.library.custom_op("rs::seq_front_bwd", mutates_args=())
def seq_front_bwd_op(dx: Tensor, ids: Sequence[Tensor], tables: Sequence[Tensor], gates: Tensor, seq_len: int,
                     padding_idx: int, deterministic: bool) -> List[Tensor]:
    """Returns [d_table_0 .. d_table_{n-1}, d_gates, d_pos]."""
    L.require_cuda(dx, gates)
    n = len(tables)
    ids = [_ids(i) for i in ids]
    tables = [_f32(t, "embedding table") for t in tables]
    gates = _f32(gates, "gates")
    dim = dx.shape[-1]
    dx = dx.reshape(-1, dim).contiguous()
    P = dx.shape[0]
    dev = dx.device
    rows = [t.shape[0] for t in tables]
    small_budget = SMALL_TABLE_ROWS
    mode = []
    for r in rows:
        if r <= small_budget:
            mode.append(2)
            small_budget -= r
        else:
            mode.append(0 if deterministic else 1)
    # (outputs of a registered custom op must not share storage: one allocation each)
    d_tables = [torch.zeros(r, dim, dtype=torch.float32, device=dev) for r in rows]
    d_gates = torch.zeros(max(n, 1), dtype=torch.float32, device=dev)
    d_pos = torch.empty(seq_len, dim, dtype=torch.float32, device=dev)
    rows_a, mode_a = L.i64_array(rows), L.i32_array(mode)
    wsb = _lib.rs_seq_front_bwd_workspace_bytes(P, seq_len, dim, n, rows_a, mode_a)
    if wsb == 0:
        raise RuntimeError("rs_seq_front_bwd: unsupported shape")
    ws = L.workspace(wsb, dev)
    L.check(_lib.rs_seq_front_bwd(L.ptr(dx), L.dt(dx), L.ptr_array(ids), L.ptr_array(tables), rows_a, mode_a, n,
                                  L.ptr(gates), seq_len, P, dim, padding_idx, L.ptr_array(d_tables), L.ptr(d_gates),
                                  L.ptr(d_pos), L.ptr(ws), ws.numel(), L.stream()), "rs_seq_front_bwd")
    for t in range(n):
        if mode[t] != 0:
            continue
        sk, sp = sorted_ids(ids[t], rows[t])
        wsb = _lib.rs_segment_reduce_workspace_bytes(P, dim)
        ws2 = L.workspace(wsb, dev)
        L.check(_lib.rs_segment_reduce_rows(L.ptr(dx), L.dt(dx), L.ptr(sk), L.ptr(sp), P, dim, rows[t], padding_idx,
                                            L.ptr(gates[t:t + 1]), L.ptr(tables[t]), L.ptr(d_tables[t]),
                                            L.ptr(d_gates[t:t + 1]), L.ptr(ws2), ws2.numel(), L.stream()),
                "rs_segment_reduce_rows")
    return d_tables + [d_gates, d_pos]


@seq_front_bwd_op.register_fake
def _(dx, ids, tables, gates, seq_len, padding_idx, deterministic):
    dim = dx.shape[-1]
    return [t.new_empty(t.shape) for t in tables] + [gates.new_empty(len(tables)), gates.new_empty(seq_len, dim)]


@torch.library.custom_op("rs::static_front", mutates_args=())
def static_front_op(ids: Sequence[Tensor], tables: Sequence[Tensor], cont: Tensor, cont_w: Tensor, cont_b: Tensor,
                    gates: Tensor) -> Tensor:
    L.require_cuda(cont, gates)
    assert len(ids) == 9 and len(tables) == 9
    ids = [_ids(i) for i in ids]
    tables = [_f32(t, "static table") for t in tables]
    cont, cont_w, cont_b, gates = _f32(cont, "cont"), _f32(cont_w, "cont_w"), _f32(cont_b, "cont_b"), _f32(gates, "g")
    B = cont.shape[0]
    out = torch.empty(B, 100, dtype=torch.float32, device=cont.device)
    L.check(_lib.rs_static_front_fwd(L.ptr_array(ids), L.ptr_array(tables), L.i64_array([t.shape[0] for t in tables]),
                                     L.ptr(cont), L.ptr(cont_w), L.ptr(cont_b), L.ptr(gates), B, L.ptr(out),
                                     L.ptr(L.oob_flag(cont.device)), L.stream()), "rs_static_front_fwd")
    return out


@static_front_op.register_fake
def _(ids, tables, cont, cont_w, cont_b, gates):
    return cont.new_empty(cont.shape[0], 100)


@torch.library.custom_op("rs::static_front_bwd", mutates_args=())
def static_front_bwd_op(d_out: Tensor, ids: Sequence[Tensor], tables: Sequence[Tensor], cont: Tensor,
                        cont_w: Tensor, cont_b: Tensor, gates: Tensor, padding_idx: int) -> List[Tensor]:
    """Returns [d_table_0..8, d_gates[10], d_cont_w[16,4], d_cont_b[16]]."""
    ids = [_ids(i) for i in ids]
    tables = [_f32(t, "static table") for t in tables]
    d_out = _f32(d_out, "d_out")
    cont, cont_w, cont_b, gates = _f32(cont, "cont"), _f32(cont_w, "cont_w"), _f32(cont_b, "cont_b"), _f32(gates, "g")
    B, dev = cont.shape[0], cont.device
    d_tables = [torch.zeros_like(t) for t in tables]
    d_gates = torch.empty(10, dtype=torch.float32, device=dev)
    d_w = torch.empty(16, 4, dtype=torch.float32, device=dev)
    d_b = torch.empty(16, dtype=torch.float32, device=dev)
    ws = L.workspace(_lib.rs_static_front_bwd_workspace_bytes(B), dev)
    L.check(_lib.rs_static_front_bwd(L.ptr(d_out), L.ptr_array(ids), L.ptr_array(tables),
                                     L.i64_array([t.shape[0] for t in tables]), L.ptr(cont), L.ptr(cont_w),
                                     L.ptr(cont_b), L.ptr(gates), B, padding_idx, L.ptr_array(d_tables),
                                     L.ptr(d_gates), L.ptr(d_w), L.ptr(d_b), L.ptr(ws), ws.numel(), L.stream()),
            "rs_static_front_bwd")
    return d_tables + [d_gates, d_w, d_b]


@static_front_bwd_op.register_fake
def _(d_out, ids, tables, cont, cont_w, cont_b, gates, padding_idx):
    return [torch.empty_like(t) for t in tables] + [gates.new_empty(10), gates.new_empty(16, 4), gates.new_empty(16)]


@torch.library.custom_op("rs::normalized_rows", mutates_args=())
def normalized_rows_op(table: Tensor, ids: Tensor, eps: float, out_dtype: int) -> Tensor:
    L.require_cuda(table, ids)
    table, ids = _f32(table, "item matrix"), _ids(ids)
    rows, dim = table.shape
    out = torch.empty(*ids.shape, dim, dtype=L.torch_dtype(out_dtype), device=table.device)
    L.check(_lib.rs_normalized_rows_fwd(L.ptr(table), rows, dim, L.ptr(ids), ids.numel(), eps, L.ptr(out), out_dtype,
                                        None, L.ptr(L.oob_flag(table.device)), L.stream()), "rs_normalized_rows_fwd")
    return out


@normalized_rows_op.register_fake
def _(table, ids, eps, out_dtype):
    return table.new_empty(*ids.shape, table.shape[1], dtype=L.torch_dtype(out_dtype))


@torch.library.custom_op("rs::normalized_rows_bwd", mutates_args=())
def normalized_rows_bwd_op(grad: Tensor, table: Tensor, ids: Tensor, eps: float) -> Tensor:
    table, ids = _f32(table, "item matrix"), _ids(ids)
    rows, dim = table.shape
    grad = grad.reshape(-1, dim).contiguous()
    d_table = torch.zeros_like(table)
    L.check(_lib.rs_normalized_rows_bwd(L.ptr(grad), L.dt(grad), L.ptr(table), rows, dim, L.ptr(ids), ids.numel(),
                                        eps, L.ptr(d_table), L.stream()), "rs_normalized_rows_bwd")
    return d_table


@normalized_rows_bwd_op.register_fake
def _(grad, table, ids, eps):
    return torch.empty_like(table)


@torch.library.custom_op("rs::std_front", mutates_args=())
def std_front_op(table: Tensor, ids: Tensor, field_emb: Tensor, ln_w: Tensor, ln_b: Tensor, eps: float,
                 out_dtype: int) -> Tensor:
    L.require_cuda(table, ids)
    table, ids = _f32(table, "std_embedding"), _ids(ids)
    field_emb = _f32(field_emb, "std_field_emb").reshape(-1, table.shape[1])
    rows, dim = table.shape
    out = torch.empty(*ids.shape, dim, dtype=L.torch_dtype(out_dtype), device=table.device)
    L.check(_lib.rs_std_front_fwd(L.ptr(table), rows, dim, L.ptr(ids), ids.numel(), L.ptr(field_emb),
                                  field_emb.shape[0], L.ptr(_f32(ln_w, "ln_w")), L.ptr(_f32(ln_b, "ln_b")), eps,
                                  L.ptr(out), out_dtype, None, None, L.ptr(L.oob_flag(table.device)), L.stream()),
            "rs_std_front_fwd")
    return out


@std_front_op.register_fake
def _(table, ids, field_emb, ln_w, ln_b, eps, out_dtype):
    return table.new_empty(*ids.shape, table.shape[1], dtype=L.torch_dtype(out_dtype))


@torch.library.custom_op("rs::bert_embed", mutates_args=())
def bert_embed_op(word: Tensor, pos: Tensor, tok_type: Tensor, ln_w: Tensor, ln_b: Tensor, eps: float, ids: Tensor,
                  dropout_p: float, seed: int, out_dtype: int) -> Tensor:
    L.require_cuda(word, ids)
    word, ids = _f32(word, "word_embeddings"), _ids(ids)
    vocab, dim = word.shape
    T = ids.shape[-1]
    pos = _f32(pos, "position_embeddings")
    if pos.shape[0] < T:
        raise IndexError("sequence longer than the position table")
    out = torch.empty(*ids.shape, dim, dtype=L.torch_dtype(out_dtype), device=word.device)
    L.check(_lib.rs_bert_embed_fwd(L.ptr(word), vocab, L.ptr(pos), L.ptr(_f32(tok_type, "token_type")[0].contiguous()),
                                   L.ptr(_f32(ln_w, "ln_w")), L.ptr(_f32(ln_b, "ln_b")), eps, L.ptr(ids),
                                   ids.numel() // T, T, dim, dropout_p, seed, L.ptr(out), out_dtype,
                                   L.ptr(L.oob_flag(word.device)), L.stream()), "rs_bert_embed_fwd")
    return out


@bert_embed_op.register_fake
def _(word, pos, tok_type, ln_w, ln_b, eps, ids, dropout_p, seed, out_dtype):
    return word.new_empty(*ids.shape, word.shape[1], dtype=L.torch_dtype(out_dtype))


@torch.library.custom_op("rs::masked_mean", mutates_args=())
def masked_mean_op(feats: Tensor, mask: Tensor) -> Tensor:
    L.require_cuda(feats, mask)
    feats, mask = _c(feats), _ids(mask)
    R, T, D = feats.shape
    out = torch.empty(R, D, dtype=torch.float32, device=feats.device)
    L.check(_lib.rs_masked_mean_fwd(L.ptr(feats), L.dt(feats), L.ptr(mask), R, T, D, L.ptr(out), L.stream()),
            "rs_masked_mean_fwd")
    return out


@masked_mean_op.register_fake
def _(feats, mask):
    return feats.new_empty(feats.shape[0], feats.shape[2], dtype=torch.float32)


@torch.library.custom_op("rs::masked_mean_bwd", mutates_args=())
def masked_mean_bwd_op(d_out: Tensor, mask: Tensor, out_dtype: int) -> Tensor:
    d_out, mask = _f32(d_out, "d_out"), _ids(mask)
    R, T = mask.shape
    D = d_out.shape[1]
    d_feats = torch.empty(R, T, D, dtype=L.torch_dtype(out_dtype), device=d_out.device)
    L.check(_lib.rs_masked_mean_bwd(L.ptr(d_out), L.ptr(mask), R, T, D, L.ptr(d_feats), out_dtype, L.stream()),
            "rs_masked_mean_bwd")
    return d_feats


@masked_mean_bwd_op.register_fake
def _(d_out, mask, out_dtype):
    return d_out.new_empty(mask.shape[0], mask.shape[1], d_out.shape[1], dtype=L.torch_dtype(out_dtype))


@torch.library.custom_op("rs::retrieve_topk", mutates_args=())
def retrieve_topk_op(users: Tensor, items: Tensor, k: int, mask_index0: bool) -> Tuple[Tensor, Tensor]:
    L.require_cuda(users, items)
    users, items = _f32(users, "user_emb"), _f32(items, "item_emb")
    nu, dim = users.shape
    ni = items.shape[0]
    if k > ni:
        raise RuntimeError("selected index k out of range")          # torch.topk's message
    ids = torch.empty(nu, k, dtype=torch.int64, device=users.device)
    scores = torch.empty(nu, k, dtype=torch.float32, device=users.device)
    ws = L.workspace(_lib.rs_topk_workspace_bytes(nu, ni, dim, k), users.device)
    L.check(_lib.rs_retrieve_topk(L.ptr(users), nu, L.ptr(items), ni, dim, k, int(mask_index0), L.ptr(ids),
                                  L.ptr(scores), L.ptr(ws), ws.numel(), L.stream()), "rs_retrieve_topk")
    return scores, ids


@retrieve_topk_op.register_fake
def _(users, items, k, mask_index0):
    return users.new_empty(users.shape[0], k), users.new_empty(users.shape[0], k, dtype=torch.int64)


TC_RETRIEVAL_CHUNK = 65536      # users per call of the tensor-core path (bounds its candidate workspace to ~0.5 GB)


@torch.library.custom_op("rs::retrieve_topk_tc", mutates_args=())
def retrieve_topk_tc_op(users: Tensor, items: Tensor, k: int, mask_index0: bool) -> Tuple[Tensor, Tensor]:
    """R1 with a tcgen05 bf16 candidate pass + exact fp32 re-scoring (rs_retrieve_topk_tc): same results contract as
    retrieve_topk (fp32 ranking, ties by ascending id); dim == 128, k <= 32."""
    L.require_cuda(users, items)
    users, items = _f32(users, "user_emb"), _f32(items, "item_emb")
    nu, dim = users.shape
    ni = items.shape[0]
    if k > ni:
        raise RuntimeError("selected index k out of range")
    ids = torch.empty(nu, k, dtype=torch.int64, device=users.device)
    scores = torch.empty(nu, k, dtype=torch.float32, device=users.device)
    for u0 in range(0, nu, TC_RETRIEVAL_CHUNK):
        n = min(TC_RETRIEVAL_CHUNK, nu - u0)
        ws = L.workspace(_lib.rs_retrieve_topk_tc_workspace_bytes(n, ni, dim, k), users.device)
        L.check(_lib.rs_retrieve_topk_tc(L.ptr(users[u0:u0 + n]), n, L.ptr(items), ni, dim, k, int(mask_index0),
                                         L.ptr(ids[u0:u0 + n]), L.ptr(scores[u0:u0 + n]), L.ptr(ws), ws.numel(),
                                         L.stream()), "rs_retrieve_topk_tc")
    return scores, ids


@retrieve_topk_tc_op.register_fake
def _(users, items, k, mask_index0):
    return users.new_empty(users.shape[0], k), users.new_empty(users.shape[0], k, dtype=torch.int64)


@torch.library.custom_op("rs::fm_fwd", mutates_args=())
def fm_fwd_op(ids: Tensor, offsets: Tensor, emb: Tensor, lin: Optional[Tensor], want_concat: bool,
              concat_dtype: int) -> Tuple[Tensor, Tensor]:
    L.require_cuda(ids, emb)
    ids, offsets, emb = _ids(ids), _ids(offsets), _f32(emb, "fm embedding table")
    B, F = ids.shape
    total, k = emb.shape
    fm = torch.empty(B, dtype=torch.float32, device=emb.device)
    concat = torch.empty((B, F * k) if want_concat else (0,), dtype=L.torch_dtype(concat_dtype), device=emb.device)
    lin_ = None if lin is None else _f32(lin, "fm linear table").reshape(-1)
    L.check(_lib.rs_fm_fwd(L.ptr(ids), L.ptr(offsets), B, F, L.ptr(emb), k, L.ptr(lin_), total, L.ptr(fm),
                           L.ptr(concat) if want_concat else None, concat_dtype, L.ptr(L.oob_flag(emb.device)),
                           L.stream()), "rs_fm_fwd")
    return fm, concat


@fm_fwd_op.register_fake
def _(ids, offsets, emb, lin, want_concat, concat_dtype):
    B, F = ids.shape
    return emb.new_empty(B), emb.new_empty((B, F * emb.shape[1]) if want_concat else (0,),
                                           dtype=L.torch_dtype(concat_dtype))


@torch.library.custom_op("rs::fm_bwd", mutates_args=())
def fm_bwd_op(ids: Tensor, offsets: Tensor, emb: Tensor, d_fm: Optional[Tensor], d_concat: Optional[Tensor],
              want_lin: bool) -> Tuple[Tensor, Tensor]:
    ids, offsets, emb = _ids(ids), _ids(offsets), _f32(emb, "fm embedding table")
    B, F = ids.shape
    total, k = emb.shape
    d_emb = torch.zeros_like(emb)
    d_lin = torch.zeros(total if want_lin else 0, dtype=torch.float32, device=emb.device)
    d_fm_ = None if d_fm is None else _f32(d_fm, "d_fm")
    d_concat = _c(d_concat)
    L.check(_lib.rs_fm_bwd(L.ptr(ids), L.ptr(offsets), B, F, L.ptr(emb), k, L.ptr(d_fm_), L.ptr(d_concat),
                           L.dt(d_concat) if d_concat is not None else 0, total, L.ptr(d_emb),
                           L.ptr(d_lin) if want_lin else None, L.stream()), "rs_fm_bwd")
    return d_emb, d_lin


@fm_bwd_op.register_fake
def _(ids, offsets, emb, d_fm, d_concat, want_lin):
    return torch.empty_like(emb), emb.new_empty(emb.shape[0] if want_lin else 0)


@torch.library.custom_op("rs::mine_hard_negatives", mutates_args=())
def mine_hard_negatives_op(u: Tensor, v: Tensor, key: Tensor, k: int, hnm_threshold: float) -> List[Tensor]:
    """[scores[n,k], ids[n,k] (-1 padded), avail[n] int32] -- no gradient (the reference mines under no_grad)."""
    L.require_cuda(u, v, key)
    u, v, key = _f32(u, "u"), _f32(v, "v"), _ids(key)
    n, dim = u.shape
    ids = torch.empty(n, k, dtype=torch.int64, device=u.device)
    scores = torch.empty(n, k, dtype=torch.float32, device=u.device)
    avail = torch.empty(n, dtype=torch.int32, device=u.device)
    ws = L.workspace(_lib.rs_mine_workspace_bytes(n, dim, k), u.device)
    L.check(_lib.rs_mine_hard_negatives(L.ptr(u), L.ptr(v), L.ptr(key), n, dim, k, hnm_threshold, L.ptr(ids),
                                        L.ptr(scores), L.ptr(avail), L.ptr(ws), ws.numel(), L.stream()),
            "rs_mine_hard_negatives")
    return [scores, ids, avail]


@mine_hard_negatives_op.register_fake
def _(u, v, key, k, hnm_threshold):
    n = u.shape[0]
    return [u.new_empty(n, k), u.new_empty(n, k, dtype=torch.int64), u.new_empty(n, dtype=torch.int32)]


@torch.library.custom_op("rs::sparse_logits", mutates_args=())
def sparse_logits_op(a: Tensor, b: Tensor, idx: Tensor, scale: float, bias: Optional[Tensor],
                     key_row: Optional[Tensor], key_col: Optional[Tensor]) -> Tensor:
    L.require_cuda(a, b, idx)
    a, b, idx = _c(a), _c(b), _ids(idx)
    n, dim = a.shape
    k = idx.shape[1]
    out = torch.empty(n, k, dtype=torch.float32, device=a.device)
    bias_ = None if bias is None else _f32(bias, "bias")
    kr = None if key_row is None else _ids(key_row)
    kc = None if key_col is None else _ids(key_col)
    L.check(_lib.rs_sparse_logits_fwd(L.ptr(a), L.ptr(b), L.dt(a), L.ptr(idx), n, b.shape[0], k, dim, scale,
                                      L.ptr(bias_), L.ptr(kr), L.ptr(kc), L.ptr(out), L.stream()),
            "rs_sparse_logits_fwd")
    return out


@sparse_logits_op.register_fake
def _(a, b, idx, scale, bias, key_row, key_col):
    return a.new_empty(a.shape[0], idx.shape[1], dtype=torch.float32)


@torch.library.custom_op("rs::sparse_logits_bwd", mutates_args=())
def sparse_logits_bwd_op(a: Tensor, b: Tensor, idx: Tensor, scale: float, key_row: Optional[Tensor],
                         key_col: Optional[Tensor], g: Tensor) -> List[Tensor]:
    a, b, idx = _c(a), _c(b), _ids(idx)
    n, dim = a.shape
    d_a = torch.empty(n, dim, dtype=torch.float32, device=a.device)
    d_b = torch.zeros(b.shape[0], dim, dtype=torch.float32, device=a.device)
    kr = None if key_row is None else _ids(key_row)
    kc = None if key_col is None else _ids(key_col)
    L.check(_lib.rs_sparse_logits_bwd(L.ptr(a), L.ptr(b), L.dt(a), L.ptr(idx), n, b.shape[0], idx.shape[1], dim,
                                      scale, L.ptr(kr), L.ptr(kc), L.ptr(_f32(g, "g")), L.ptr(d_a), L.ptr(d_b),
                                      L.stream()), "rs_sparse_logits_bwd")
    return [d_a, d_b]


@sparse_logits_bwd_op.register_fake
def _(a, b, idx, scale, key_row, key_col, g):
    return [a.new_empty(a.shape, dtype=torch.float32), b.new_empty(b.shape, dtype=torch.float32)]


# ---- N1: on-device batch assembly ---------------------------------------------------------------
def zeros_many(shapes, device, dtype=torch.float32):
    """One zero-filled allocation, returned as one contiguous tensor per shape (256-byte aligned views): a single fill
    launch instead of one per tensor.  Only for tensors built inside autograd Functions: the outputs of a registered
    custom op must not share storage."""
    esz = torch.empty(0, dtype=dtype).element_size()
    q = 256 // esz
    sizes = [int(torch.Size(sh).numel()) for sh in shapes]
    offs, total = [], 0
    for n in sizes:
        offs.append(total)
        total += (n + q - 1) // q * q
    flat = torch.zeros(max(total, 1), dtype=dtype, device=device)
    return [flat[o:o + n].view(*sh) for o, n, sh in zip(offs, sizes, shapes)]


def round_up(n: int, q: int) -> int:
    return (int(n) + q - 1) // q * q


def batch_index_counts(padding_mask: Tensor, target_ids: Tensor, n_item_rows: int, meta: Optional[Tensor] = None) -> Tensor:
    """int32[8] device tensor: [0] valid time steps T, [1] extras E, [2] distinct valid targets U (rs_batch_index_counts):
    what the loader needs to pick the shape bucket of a batch.  Stream-ordered, no synchronisation."""
    L.require_cuda(padding_mask, target_ids)
    if padding_mask.dtype != torch.bool:
        raise TypeError("padding_mask must be a bool tensor (True = padding)")
    pm, tg = padding_mask.contiguous(), _ids(target_ids)
    B, SL = pm.shape
    meta = torch.empty(8, dtype=torch.int32, device=pm.device) if meta is None else meta
    ws = L.workspace(_lib.rs_batch_index_workspace_bytes(B, SL, n_item_rows), pm.device)
    L.check(_lib.rs_batch_index_counts(L.ptr(pm), L.ptr(tg), B, SL, n_item_rows, L.ptr(meta), L.ptr(ws), ws.numel(),
                                       L.stream()), "rs_batch_index_counts")
    return meta


def batch_index_alloc(B: int, SL: int, tok_cap: int, col_cap: int, device) -> dict:
    """Static-shape output arrays of rs_batch_index_build for one shape bucket (tok_cap main rows, col_cap columns)."""
    grid_cap = round_up(tok_cap + B, 64)                 # E <= B extras behind the T <= tok_cap valid tokens
    i64 = lambda *s: torch.empty(*s, dtype=torch.int64, device=device)
    return dict(pk_item_ids=i64(grid_cap // 64, 64), pk_time_ids=i64(grid_cap // 64, 64), pk_pos_ids=i64(grid_cap // 64, 64),
                pk_index_2v=i64(2 * grid_cap), fold_inv1=i64(grid_cap), fold_inv2=i64(grid_cap),
                cu_seqlens_2v=torch.empty(2 * B + 2, dtype=torch.int32, device=device),
                row_cu=torch.empty(B + 1, dtype=torch.int32, device=device),
                select_2v=i64(tok_cap + 2 * B), users_2v=i64(tok_cap + 2 * B), main_tgt=i64(tok_cap), last_tgt=i64(B),
                row_weight=torch.empty(tok_cap, dtype=torch.float32, device=device), col_item_ids=i64(col_cap),
                col_counts=torch.empty(col_cap, dtype=torch.float32, device=device), pos_col=i64(tok_cap),
                meta=torch.empty(8, dtype=torch.int32, device=device))


def batch_index_build(padding_mask: Tensor, item_ids: Tensor, time_ids: Tensor, target_ids: Tensor, n_item_rows: int,
                      out: dict) -> dict:
    """Fill `out` (from batch_index_alloc) for one collated batch: every index the packed two-view step consumes
    (rs_batch_index_build; layout in include/rs_twotower.h).  Stream-ordered, static shapes: graph-capturable.
    `out["meta"]` carries the true counts and the overflow flag."""
    L.require_cuda(padding_mask, item_ids, time_ids, target_ids)
    if padding_mask.dtype != torch.bool:
        raise TypeError("padding_mask must be a bool tensor (True = padding)")
    pm, it, tm, tg = padding_mask.contiguous(), _ids(item_ids), _ids(time_ids), _ids(target_ids)
    B, SL = pm.shape
    d = L.BatchIndex()
    d.padding_mask, d.item_ids, d.time_ids, d.target_ids = pm.data_ptr(), it.data_ptr(), tm.data_ptr(), tg.data_ptr()
    d.B, d.L, d.n_item_rows = B, SL, n_item_rows
    d.tok_cap, d.col_cap, d.grid_cap = out["main_tgt"].numel(), out["col_item_ids"].numel(), out["fold_inv1"].numel()
    for name in L.BATCH_INDEX_OUTPUTS:
        setattr(d, name, out[name].data_ptr())
    ws = L.workspace(_lib.rs_batch_index_workspace_bytes(B, SL, n_item_rows), pm.device)
    L.check(_lib.rs_batch_index_build(d, L.ptr(ws), ws.numel(), L.stream()), "rs_batch_index_build")
    return out


# ---- row-sharded tables: device-side routing ------------------------------------------------------
def id_histogram(ids: Tensor, n_bins: int, n_valid: Optional[Tensor] = None, force_bin0: bool = False) -> Tensor:
    """int32 [n_bins] occurrences of every id among the first `n_valid` (device int32 scalar; default all) entries."""
    L.require_cuda(ids)
    ids = _ids(ids).reshape(-1)
    cnt = torch.empty(n_bins, dtype=torch.int32, device=ids.device)
    L.check(_lib.rs_id_histogram(L.ptr(ids), ids.numel(), L.ptr(n_valid), n_bins, int(force_bin0), L.ptr(cnt),
                                 L.ptr(L.oob_flag(ids.device)), L.stream()), "rs_id_histogram")
    return cnt


def owner_compact(cnt: Tensor, world: int, rows_per_owner: int, cap: int, want_ids: bool = False):
    """Owner-major compaction of the present ids of a histogram (rs_owner_compact):
    (rows [world*cap] int64 local rows / -1, ids [world*cap] int64 | None, counts [world*cap] fp32 | None,
     slot_of [n_ids] int32, meta int32[4] = (largest per-owner count, overflow flag, present ids, -))."""
    L.require_cuda(cnt)
    dev, n_ids = cnt.device, cnt.numel()
    rows = torch.empty(world * cap, dtype=torch.int64, device=dev)
    ids = torch.empty(world * cap, dtype=torch.int64, device=dev) if want_ids else None
    cf = torch.empty(world * cap, dtype=torch.float32, device=dev) if want_ids else None
    slot_of = torch.empty(n_ids, dtype=torch.int32, device=dev)
    meta = torch.empty(4, dtype=torch.int32, device=dev)
    ws = L.workspace(_lib.rs_owner_compact_workspace_bytes(world, rows_per_owner), dev)
    L.check(_lib.rs_owner_compact(L.ptr(cnt), world, rows_per_owner, n_ids, cap, L.ptr(rows), L.ptr(ids), L.ptr(cf),
                                  L.ptr(slot_of), L.ptr(meta), L.ptr(ws), ws.numel(), L.stream()), "rs_owner_compact")
    return rows, ids, cf, slot_of, meta


def lookup_i32(table: Tensor, ids: Tensor, fill: int = 0) -> Tensor:
    """table[ids] (int32 table) as int64, `fill` for ids outside the table or negative entries."""
    L.require_cuda(table, ids)
    ids = _ids(ids)
    out = torch.empty(ids.shape, dtype=torch.int64, device=ids.device)
    L.check(_lib.rs_lookup_i32(L.ptr(table), table.numel(), L.ptr(ids), ids.numel(), fill, L.ptr(out), L.stream()),
            "rs_lookup_i32")
    return out


@torch.library.custom_op("rs::gather_add2", mutates_args=())
def gather_add2_op(x: Tensor, i1: Tensor, i2: Tensor) -> Tensor:
    """out[u] = x[i1[u]] + x[i2[u]] (rows; an index outside [0, len(x)) contributes 0)."""
    L.require_cuda(x, i1, i2)
    x, i1, i2 = x.contiguous(), _ids(i1), _ids(i2)
    out = torch.empty(i1.numel(), x.shape[1], dtype=x.dtype, device=x.device)
    L.check(_lib.rs_gather_add2(L.ptr(x), L.dt(x), L.ptr(i1), L.ptr(i2), i1.numel(), x.shape[0], x.shape[1], L.ptr(out),
                                L.stream()), "rs_gather_add2")
    return out


@gather_add2_op.register_fake
def _(x, i1, i2):
    return x.new_empty(i1.numel(), x.shape[1])


# ---- fused in-batch softmax -------------------------------------------------------------------
def _ce_problem(a, b, scale, col_bias, key_a_row, key_a_col, key_b_row, key_b_col, diag_offset, mask_value, flags,
                logit_bound=0.0):
    p = L.CEProblem()
    p.a, p.b = a.data_ptr(), b.data_ptr()
    p.ab_dtype = L.dt(a)
    p.M, p.N, p.K = a.shape[0], b.shape[0], a.shape[1]
    p.scale = scale
    p.col_bias = col_bias.data_ptr() if col_bias is not None else None
    p.key_a_row = key_a_row.data_ptr() if key_a_row is not None else None
    p.key_a_col = key_a_col.data_ptr() if key_a_col is not None else None
    p.key_b_row = key_b_row.data_ptr() if key_b_row is not None else None
    p.key_b_col = key_b_col.data_ptr() if key_b_col is not None else None
    p.diag_offset, p.mask_value, p.flags = diag_offset, mask_value, flags
    p.logit_bound = logit_bound
    return p


def _ce_prepare(a, b, col_bias, keys):
    if a.dtype not in (torch.bfloat16, torch.float16) or b.dtype != a.dtype:
        raise TypeError("fused softmax operands must both be bfloat16 or float16")
    a, b = a.contiguous(), b.contiguous()
    col_bias = None if col_bias is None else _f32(col_bias, "col_bias")
    keys = [None if k is None else _ids(k) for k in keys]
    return a, b, col_bias, keys


@torch.library.custom_op("rs::ce_fwd", mutates_args=())
def ce_fwd_op(a: Tensor, b: Tensor, scale: float, col_bias: Optional[Tensor], key_a_row: Optional[Tensor],
              key_a_col: Optional[Tensor], key_b_row: Optional[Tensor], key_b_col: Optional[Tensor],
              diag_offset: int, mask_value: float, flags: int, logit_bound: float = 0.0) -> List[Tensor]:
    """Returns [lse[M], diag[M], pos_sum[M], pos_cnt[M]] (the last two only meaningful with RS_CE_SUPCON)."""
    L.require_cuda(a, b)
    a, b, col_bias, keys = _ce_prepare(a, b, col_bias, [key_a_row, key_a_col, key_b_row, key_b_col])
    M, dev = a.shape[0], a.device
    lse = torch.empty(M, dtype=torch.float32, device=dev)
    diag = torch.empty(M, dtype=torch.float32, device=dev)
    sup = bool(flags & L.RS_CE_SUPCON)
    pos_sum = torch.empty(M if sup else 0, dtype=torch.float32, device=dev)
    pos_cnt = torch.empty(M if sup else 0, dtype=torch.float32, device=dev)
    p = _ce_problem(a, b, scale, col_bias, *keys, diag_offset, mask_value, flags, logit_bound)
    ws = L.workspace(_lib.rs_ce_workspace_bytes(p), dev)
    L.check(_lib.rs_ce_fwd(p, L.ptr(lse), L.ptr(diag), L.ptr(pos_sum) if sup else None,
                           L.ptr(pos_cnt) if sup else None, L.ptr(ws), ws.numel(), L.stream()), "rs_ce_fwd")
    return [lse, diag, pos_sum, pos_cnt]


@ce_fwd_op.register_fake
def _(a, b, scale, col_bias, key_a_row, key_a_col, key_b_row, key_b_col, diag_offset, mask_value, flags, logit_bound=0.0):
    M = a.shape[0]
    n = M if flags & L.RS_CE_SUPCON else 0
    f = lambda k: a.new_empty(k, dtype=torch.float32)
    return [f(M), f(M), f(n), f(n)]


@torch.library.custom_op("rs::ce_fwd_grad", mutates_args=())
def ce_fwd_grad_op(a: Tensor, b: Tensor, scale: float, col_bias: Optional[Tensor], key_a_row: Optional[Tensor],
                   key_a_col: Optional[Tensor], key_b_row: Optional[Tensor], key_b_col: Optional[Tensor],
                   diag_offset: int, mask_value: float, flags: int, logit_bound: float) -> List[Tensor]:
    """Forward that also accumulates the row side of the backward (rs_ce_fwd_grad): [lse[M], diag[M], g_parts, g_info].
    `g_parts` / `g_info` go to ce_bwd unchanged."""
    L.require_cuda(a, b)
    a, b, col_bias, keys = _ce_prepare(a, b, col_bias, [key_a_row, key_a_col, key_b_row, key_b_col])
    M, dev = a.shape[0], a.device
    lse = torch.empty(M, dtype=torch.float32, device=dev)
    diag = torch.empty(M, dtype=torch.float32, device=dev)
    p = _ce_problem(a, b, scale, col_bias, *keys, diag_offset, mask_value, flags, logit_bound)
    g_parts = torch.empty(_lib.rs_ce_fwd_grad_bytes(p) // 4, dtype=torch.float32, device=dev)
    g_info = torch.empty(8, dtype=torch.float32, device=dev)
    ws = L.workspace(_lib.rs_ce_workspace_bytes(p), dev)
    L.check(_lib.rs_ce_fwd_grad(p, L.ptr(lse), L.ptr(diag), L.ptr(g_parts), L.ptr(g_info), L.ptr(ws), ws.numel(),
                                L.stream()), "rs_ce_fwd_grad")
    return [lse, diag, g_parts, g_info]


@ce_fwd_grad_op.register_fake
def _(a, b, scale, col_bias, key_a_row, key_a_col, key_b_row, key_b_col, diag_offset, mask_value, flags, logit_bound):
    M = a.shape[0]
    f = lambda k: a.new_empty(k, dtype=torch.float32)
    return [f(M), f(M), f(M * a.shape[1]), f(8)]


def ce_fwd_grad_supported(a: Tensor, flags: int, mask_value: float, logit_bound: float, has_keys: bool) -> bool:
    """Host-side preconditions of rs_ce_fwd_grad (the device decides the rest, see g_info[1])."""
    if a.dtype != torch.bfloat16 or logit_bound <= 0.0 or (flags & L.RS_CE_SUPCON):
        return False
    return mask_value == float("-inf") or not (has_keys or (flags & L.RS_CE_DIAG_MASK))


@torch.library.custom_op("rs::ce_bwd", mutates_args=())
def ce_bwd_op(a: Tensor, b: Tensor, scale: float, col_bias: Optional[Tensor], key_a_row: Optional[Tensor],
              key_a_col: Optional[Tensor], key_b_row: Optional[Tensor], key_b_col: Optional[Tensor],
              diag_offset: int, mask_value: float, flags: int, lse: Tensor, w_lse: Tensor,
              w_diag: Optional[Tensor], w_pos: Optional[Tensor], logit_bound: float = 0.0,
              g_parts: Optional[Tensor] = None, g_info: Optional[Tensor] = None) -> List[Tensor]:
    """Returns [dA[M,K], dB[N,K]] in fp32 for dS = w_lse*softmax + w_diag*[diag] + w_pos*[positive].
    With `g_parts` / `g_info` (outputs of ce_fwd_grad for the same problem) dA is a row scaling of the forward's G and
    only the dB side runs on the tensor cores (rs_ce_bwd_from_grad)."""
    a, b, col_bias, keys = _ce_prepare(a, b, col_bias, [key_a_row, key_a_col, key_b_row, key_b_col])
    dev = a.device
    dA = torch.empty(a.shape, dtype=torch.float32, device=dev)
    dB = torch.empty(b.shape, dtype=torch.float32, device=dev)
    p = _ce_problem(a, b, scale, col_bias, *keys, diag_offset, mask_value, flags, logit_bound)
    ws = L.workspace(_lib.rs_ce_workspace_bytes(p), dev)
    w_diag_ = None if w_diag is None else _f32(w_diag, "w_diag")
    w_pos_ = None if w_pos is None else _f32(w_pos, "w_pos")
    if g_parts is not None:
        L.check(_lib.rs_ce_bwd_from_grad(p, L.ptr(_f32(lse, "lse")), L.ptr(_f32(w_lse, "w_lse")), L.ptr(w_diag_),
                                         L.ptr(g_parts), L.ptr(g_info), L.ptr(dA), L.ptr(dB), L.ptr(ws), ws.numel(),
                                         L.stream()), "rs_ce_bwd_from_grad")
        return [dA, dB]
    L.check(_lib.rs_ce_bwd(p, L.ptr(_f32(lse, "lse")), L.ptr(_f32(w_lse, "w_lse")), L.ptr(w_diag_), L.ptr(w_pos_),
                           L.ptr(dA), L.ptr(dB), L.ptr(ws), ws.numel(), L.stream()), "rs_ce_bwd")
    return [dA, dB]


@ce_bwd_op.register_fake
def _(a, b, scale, col_bias, key_a_row, key_a_col, key_b_row, key_b_col, diag_offset, mask_value, flags, lse, w_lse,
      w_diag, w_pos, logit_bound=0.0, g_parts=None, g_info=None):
    return [a.new_empty(a.shape, dtype=torch.float32), b.new_empty(b.shape, dtype=torch.float32)]


# ---------------------------------------------------------------------------------------------
# autograd
# ---------------------------------------------------------------------------------------------
DETERMINISTIC = True        # sorted segment-reduce backward (default); False -> vector atomics


def _out_dtype(*float_inputs) -> int:
    """Output dtype: the autocast dtype when autocast is on (what nn.Linear would hand on), else fp32."""
    if torch.is_autocast_enabled("cuda"):
        return L.dt(torch.get_autocast_dtype("cuda"))
    return L.RS_F32


class _GatherRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, table, ids, padding_idx, clamp_max, out_dtype):
        ctx.save_for_backward(ids)
        ctx.meta = (table.shape[0], padding_idx, clamp_max, table.dtype)
        return L.direct.gather_rows(table, ids, clamp_max, out_dtype)

    @staticmethod
    def backward(ctx, g):
        (ids,) = ctx.saved_tensors
        rows, padding_idx, clamp_max, tdt = ctx.meta
        d = L.direct.embedding_dense_bwd(g, ids, rows, padding_idx, clamp_max, DETERMINISTIC)
        return d.to(tdt), None, None, None, None


def gather_rows(table: Tensor, ids: Tensor, padding_idx: int = -1, clamp_max: int = -1,
                out_dtype: Optional[torch.dtype] = None) -> Tensor:
    """`table[ids]` / `F.embedding(ids, table, padding_idx)`; rows are copied bit-exactly."""
    od = L.dt(out_dtype) if out_dtype is not None else L.dt(table)
    return _GatherRows.apply(table, ids, padding_idx, clamp_max, od)


class _SplitRows(torch.autograd.Function):
    """torch.split(x, sizes) along dim 0 whose backward is ONE concatenation of the pieces' gradients (the stock slices
    each allocate a zero tensor of the whole shape, copy their piece into it, and autograd then adds the results)."""

    @staticmethod
    def forward(ctx, x, *sizes):
        ctx.sizes = sizes
        ctx.meta = (x.shape[1:], x.dtype, x.device)
        return tuple(torch.split(x, list(sizes)))

    @staticmethod
    def backward(ctx, *gs):
        tail, dt, dev = ctx.meta
        parts = [g if g is not None else torch.zeros(n, *tail, dtype=dt, device=dev) for g, n in zip(gs, ctx.sizes)]
        return (torch.cat([p.to(dt) for p in parts]),) + (None,) * len(ctx.sizes)


def split_rows(x: Tensor, *sizes: int):
    """`torch.split(x, sizes)` (views of x; do not modify them in place) with a single-pass backward."""
    assert sum(sizes) == x.shape[0]
    return _SplitRows.apply(x, *[int(n) for n in sizes])


class _SelectPrefix(torch.autograd.Function):
    """cat([x[:n_prefix], x[idx]]) for row matrices.  Backward: the prefix's gradient is COPIED into place and the
    gathered rows' gradients are added on top (idx must not repeat a row: every row then has at most one addend after
    the copy, so the result is deterministic although the add is an atomic) -- no sort, no segment reduce.
    `disjoint`: the caller vouches that every idx entry is either -1 (the null id: a zero row, no gradient) or a row
    OUTSIDE the prefix; the backward is then pure data movement in x's own dtype (copy, clear, row scatter)."""

    @staticmethod
    def forward(ctx, x, n_prefix, idx, out_dtype, disjoint):
        L.require_cuda(x, idx)
        x, idx = x.contiguous(), _ids(idx)
        P, D = x.shape
        m = idx.numel()
        out = torch.empty(n_prefix + m, D, dtype=out_dtype or x.dtype, device=x.device)
        out[:n_prefix].copy_(x[:n_prefix])                       # (converts on the way when out_dtype differs)
        L.check(_lib.rs_gather_rows(L.ptr(x), L.dt(x), P, D, L.ptr(idx), m, -1,
                                    L.C.c_void_p(out.data_ptr() + n_prefix * D * out.element_size()), L.dt(out),
                                    L.ptr(L.oob_flag(x.device)), L.stream()), "rs_gather_rows")
        if disjoint:
            idx = torch.where(idx < 0, P, idx)                   # null ids -> a scratch row behind the matrix
        ctx.save_for_backward(idx)
        ctx.meta = (P, n_prefix, x.dtype, disjoint)
        return out

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        P, n_prefix, xdt, disjoint = ctx.meta
        g = g.contiguous()
        D = g.shape[1]
        if disjoint:
            d = torch.empty(P + 1, D, dtype=xdt, device=g.device)
            d[:n_prefix].copy_(g[:n_prefix])
            d[n_prefix:].zero_()
            d.index_copy_(0, idx, g[n_prefix:].to(xdt))
            return d[:P], None, None, None, None
        d = torch.empty(P, D, dtype=torch.float32, device=g.device)
        d[:n_prefix].copy_(g[:n_prefix])
        d[n_prefix:].zero_()
        tail = g[n_prefix:]
        L.check(_lib.rs_scatter_add_rows(L.ptr(tail), L.dt(tail), L.ptr(idx), idx.numel(), D, P, -1, -1, 1.0, L.ptr(d),
                                         L.ptr(L.oob_flag(g.device)), L.stream()), "rs_scatter_add_rows")
        return d.to(xdt), None, None, None, None


def select_prefix_rows(x: Tensor, n_prefix: int, idx: Tensor, out_dtype: Optional[torch.dtype] = None,
                       disjoint: bool = False) -> Tensor:
    """cat([x[:n_prefix], x[idx]]) with a sort-free backward; `idx` must hold distinct rows.  `out_dtype`: emit the rows
    in this dtype (the cast a consuming autocast Linear would apply anyway, folded into the copy).  `disjoint`: see
    _SelectPrefix."""
    return _SelectPrefix.apply(x, int(n_prefix), idx, out_dtype, bool(disjoint))


def segment_sum_sorted(g: Tensor, ids: Tensor, rows: int) -> Tensor:
    """out[k] = sum of the rows g[i] with ids[i] == k, fp32 [rows, dim], for ASCENDING ids: the segment-reduce kernel on
    (ids, arange) directly, without the radix sort (fixed summation order)."""
    g = g.reshape(-1, g.shape[-1]).contiguous()
    n, dim = g.shape
    sk = ids.reshape(-1).to(torch.int32)
    sp = torch.arange(n, dtype=torch.int32, device=g.device)
    out = torch.zeros(rows, dim, dtype=torch.float32, device=g.device)
    ws = L.workspace(_lib.rs_segment_reduce_workspace_bytes(n, dim), g.device)
    L.check(_lib.rs_segment_reduce_rows(L.ptr(g), L.dt(g), L.ptr(sk), L.ptr(sp), n, dim, rows, -1, None, None,
                                        L.ptr(out), None, L.ptr(ws), ws.numel(), L.stream()),
            "rs_segment_reduce_rows")
    return out


def scatter_add_distinct_(dst: Tensor, g: Tensor, ids: Tensor) -> Tensor:
    """dst[ids[i]] += g[i] in place (dst fp32) for DISTINCT ids: one addend per row, so the atomic adds cannot reorder
    anything and the result is deterministic."""
    g = g.contiguous()
    ids = _ids(ids)
    L.check(_lib.rs_scatter_add_rows(L.ptr(g), L.dt(g), L.ptr(ids), ids.numel(), g.shape[1], dst.shape[0], -1, -1, 1.0,
                                     L.ptr(dst), L.ptr(L.oob_flag(g.device)), L.stream()), "rs_scatter_add_rows")
    return dst


class _GatherRowsSorted(torch.autograd.Function):
    """table[ids] for ASCENDING ids (e.g. the user of every packed row, batch-major): the backward is a segment sum over
    runs of equal ids (segment_sum_sorted)."""

    @staticmethod
    def forward(ctx, table, ids, out_dtype):
        ctx.save_for_backward(ids)
        ctx.meta = (table.shape[0], table.dtype)
        return L.direct.gather_rows(table, ids, -1, L.dt(out_dtype or table.dtype))

    @staticmethod
    def backward(ctx, g):
        (ids,) = ctx.saved_tensors
        rows, tdt = ctx.meta
        return segment_sum_sorted(g, ids, rows).to(tdt), None, None


def gather_rows_sorted(table: Tensor, ids: Tensor, out_dtype: Optional[torch.dtype] = None) -> Tensor:
    """`table[ids]` where the caller vouches that `ids` is ascending (sort-free segment-sum backward)."""
    return _GatherRowsSorted.apply(table, _ids(ids), out_dtype)


class _SeqFront(torch.autograd.Function):
    @staticmethod
    def forward(ctx, base, gates, pos_table, padding_idx, out_dtype, n_live, *rest):
        # rest = ids[0..n) + tables[0..n); the first n_live tables are gathered, the others have a gate
        # that is hard-masked to 0 (s_mask, v1_refine_usertower.py:437): exact zeros, never read
        n = len(rest) // 2
        ids, tables = list(rest[:n]), list(rest[n:])
        L_ = ids[0].shape[-1]
        out = L.direct.seq_front(base, ids[:n_live], tables[:n_live], gates[:n_live].contiguous(), pos_table, L_,
                                     out_dtype)
        ctx.save_for_backward(gates, *ids[:n_live], *tables)
        ctx.meta = (n, n_live, L_, padding_idx, None if base is None else base.dtype,
                    None if pos_table is None else pos_table.shape[0])
        return out

    @staticmethod
    def backward(ctx, dx):
        n, n_live, L_, padding_idx, base_dtype, pos_rows = ctx.meta
        gates = ctx.saved_tensors[0]
        ids = list(ctx.saved_tensors[1:1 + n_live])
        tables = list(ctx.saved_tensors[1 + n_live:])
        res = L.direct.seq_front_bwd(dx, ids, tables[:n_live], gates[:n_live].contiguous(), L_, padding_idx,
                                         DETERMINISTIC)
        *dead, d_gates = zeros_many([tuple(t.shape) for t in tables[n_live:]] + [tuple(gates.shape)], dx.device)
        d_tables = list(res[:n_live]) + dead
        d_gates[:n_live] = res[n_live][:n_live]
        d_pos = None
        if pos_rows is not None:
            d_pos = res[n_live + 1]                  # [seq_len, dim]: rows arange(seq_len) of the position table
            if pos_rows > L_:                        # shorter batch than max_len: the other rows got no gradient
                d_pos = torch.cat([d_pos, d_pos.new_zeros(pos_rows - L_, d_pos.shape[1])])
        d_base = None if base_dtype is None else dx.to(base_dtype)
        return (d_base, d_gates, d_pos, None, None, None, *([None] * n), *d_tables)


def seq_front(base, ids, tables, gates, pos_table, padding_idx=0, n_live=None, out_dtype=None):
    """U1: base + sum_t gates[t]*tables[t][ids[t]] + pos_table[arange(L)]  (v1_refine_usertower.py:447-456)."""
    n_live = len(tables) if n_live is None else n_live
    od = L.dt(out_dtype) if out_dtype is not None else _out_dtype()
    return _SeqFront.apply(base, gates, pos_table, padding_idx, od, n_live, *ids, *tables)


class _StaticFront(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cont, cont_w, cont_b, gates, padding_idx, *rest):
        ids, tables = list(rest[:9]), list(rest[9:])
        ctx.save_for_backward(cont, cont_w, cont_b, gates, *ids, *tables)
        ctx.padding_idx = padding_idx
        return L.direct.static_front(ids, tables, cont, cont_w, cont_b, gates)

    @staticmethod
    def backward(ctx, g):
        cont, cont_w, cont_b, gates = ctx.saved_tensors[:4]
        ids, tables = list(ctx.saved_tensors[4:13]), list(ctx.saved_tensors[13:])
        res = L.direct.static_front_bwd(g.float().contiguous(), ids, tables, cont, cont_w, cont_b, gates,
                                            ctx.padding_idx)
        return (None, res[10], res[11], res[9], None, *([None] * 9), *res[:9])


def static_front(ids, tables, cont, cont_w, cont_b, gates, padding_idx=0):
    """U2: nine gated tiny gathers + gated relu(Linear(4->16)) -> [B,100]  (v1_refine_usertower.py:472-491)."""
    return _StaticFront.apply(cont, cont_w, cont_b, gates, padding_idx, *ids, *tables)


class _NormalizedRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, table, ids, eps, out_dtype):
        ctx.save_for_backward(table, ids)
        ctx.eps = eps
        return L.direct.normalized_rows(table, ids, eps, out_dtype)

    @staticmethod
    def backward(ctx, g):
        table, ids = ctx.saved_tensors
        return L.direct.normalized_rows_bwd(g, table, ids, ctx.eps), None, None, None


def normalized_rows(table, ids, eps: float = 1e-12, out_dtype=None):
    """U4: F.normalize(table, p=2, dim=1)[ids] without touching the rows that are not asked for."""
    od = L.dt(out_dtype) if out_dtype is not None else L.RS_F32
    return _NormalizedRows.apply(table, ids, eps, od)


class _MaskedMean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, mask):
        ctx.save_for_backward(mask)
        ctx.dt = L.dt(feats)
        return L.direct.masked_mean(feats, mask)

    @staticmethod
    def backward(ctx, g):
        (mask,) = ctx.saved_tensors
        return L.direct.masked_mean_bwd(g.float().contiguous(), mask, ctx.dt), None


def masked_mean(feats, mask):
    """I2 tail: sum_t feats*mask / clamp(sum_t mask, 1e-9)  (item_tower.py:254-257)."""
    return _MaskedMean.apply(feats, mask)


class _FM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ids, offsets, emb, lin, want_concat, concat_dtype):
        fm, concat = L.direct.fm_fwd(ids, offsets, emb, lin, want_concat, concat_dtype)
        ctx.save_for_backward(ids, offsets, emb)
        ctx.meta = (lin is not None, want_concat, None if lin is None else lin.shape)
        return fm, concat

    @staticmethod
    def backward(ctx, d_fm, d_concat):
        ids, offsets, emb = ctx.saved_tensors
        has_lin, want_concat, lin_shape = ctx.meta
        d_emb, d_lin = L.direct.fm_bwd(ids, offsets, emb, d_fm, d_concat if want_concat else None, has_lin)
        return None, None, d_emb, (d_lin.reshape(lin_shape) if has_lin else None), None, None


def fm_interaction(ids, offsets, emb, lin=None, want_concat=True, concat_dtype=None):
    """F1: (fm[B] (+ linear term), concat[B,F*k]) from one pass over the gathered field rows."""
    cd = L.dt(concat_dtype) if concat_dtype is not None else _out_dtype()
    return _FM.apply(ids, offsets, emb, lin, want_concat, cd)


RETRIEVAL_TENSOR_CORES = True      # route 128-dim, k <= 32 retrievals through the tcgen05 candidate pass


def retrieve_topk(user_emb: Tensor, item_emb: Tensor, k: int, mask_index0: bool = False, tensor_cores=None):
    """R1: topk(user_emb @ item_emb.T, k) -> (scores, ids); score desc, ties by ascending id.  The ranking is the fp32
    one on both paths: `tensor_cores` (default: RETRIEVAL_TENSOR_CORES when the shape allows -- 128-dim rows, k <= 32,
    at least 1024 items) uses the bf16 tcgen05 pass only as a candidate filter in front of exact fp32 re-scoring."""
    use_tc = RETRIEVAL_TENSOR_CORES if tensor_cores is None else tensor_cores
    if use_tc and user_emb.shape[-1] == 128 and k <= 32 and item_emb.shape[0] >= 1024:
        return L.direct.retrieve_topk_tc(user_emb.float(), item_emb.float(), k, mask_index0)
    return L.direct.retrieve_topk(user_emb.float(), item_emb.float(), k, mask_index0)


class _SparseLogits(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, idx, scale, bias, key_row, key_col, compute_dtype):
        # operands are rounded to `compute_dtype` here (not by the caller) so that the gradients come back in the
        # callers' dtype (fp32) instead of being rounded to 16 bits on their way through a cast node
        ac = a.detach() if compute_dtype is None else a.detach().to(compute_dtype)
        bc = b.detach() if compute_dtype is None else b.detach().to(compute_dtype)
        ctx.save_for_backward(ac, bc, idx, key_row, key_col)
        ctx.meta = (scale, a.dtype, b.dtype)
        return L.direct.sparse_logits(ac, bc, idx, scale, bias, key_row, key_col)

    @staticmethod
    def backward(ctx, g):
        a, b, idx, key_row, key_col = ctx.saved_tensors
        scale, adt, bdt = ctx.meta
        d_a, d_b = L.direct.sparse_logits_bwd(a, b, idx, scale, key_row, key_col, g.float().contiguous())
        return d_a.to(adt), d_b.to(bdt), None, None, None, None, None, None


def sparse_logits(a, b, idx, scale, bias=None, key_row=None, key_col=None, compute_dtype=None):
    """out[i,c] = scale*<a_i, b_idx[i,c]> - bias[idx[i,c]]  (-inf for idx < 0 or equal keys); differentiable in a, b.
    `compute_dtype`: round the operands to this dtype first (to match a tensor-core pass over the same operands)."""
    if bias is not None:
        bias = bias.detach().float().contiguous()
    return _SparseLogits.apply(a, b, idx, float(scale), bias, key_row, key_col, compute_dtype)


@torch.library.custom_op("rs::user_block_logits", mutates_args=())
def user_block_logits_op(u: Tensor, cols: Tensor, pos_col: Tensor, row_cu: Tensor, max_len: int, scale: float,
                         bias: Optional[Tensor]) -> List[Tensor]:
    """[s_pos[n], own_lse[n]] of the per-user blocks (see rs_user_block_logits_fwd)."""
    L.require_cuda(u, cols, pos_col, row_cu)
    u, cols, pos_col = _c(u), _c(cols), _ids(pos_col)
    n = u.shape[0]
    if row_cu.dtype != torch.int32 or not row_cu.is_contiguous():
        raise TypeError("row_cu must be a contiguous int32 tensor")
    # rows that belong to no user block (bucket padding behind row_cu[-1]) keep 0 instead of uninitialised memory
    s_pos = torch.zeros(n, dtype=torch.float32, device=u.device)
    own = torch.zeros(n, dtype=torch.float32, device=u.device)
    L.check(_lib.rs_user_block_logits_fwd(L.ptr(u), L.ptr(cols), L.dt(u), L.ptr(pos_col), L.ptr(row_cu),
                                          row_cu.numel() - 1, cols.shape[0], u.shape[1], max_len, scale, L.ptr(bias),
                                          L.ptr(s_pos), L.ptr(own), L.stream()), "rs_user_block_logits_fwd")
    return [s_pos, own]


@user_block_logits_op.register_fake
def _(u, cols, pos_col, row_cu, max_len, scale, bias):
    return [u.new_empty(u.shape[0], dtype=torch.float32), u.new_empty(u.shape[0], dtype=torch.float32)]


@torch.library.custom_op("rs::user_block_logits_bwd", mutates_args=())
def user_block_logits_bwd_op(u: Tensor, cols: Tensor, pos_col: Tensor, row_cu: Tensor, max_len: int, scale: float,
                             bias: Optional[Tensor], own_lse: Tensor, g_pos: Tensor, g_own: Tensor) -> List[Tensor]:
    u, cols, pos_col = _c(u), _c(cols), _ids(pos_col)
    d_u = torch.zeros(u.shape, dtype=torch.float32, device=u.device)
    d_c = torch.zeros(cols.shape, dtype=torch.float32, device=u.device)
    L.check(_lib.rs_user_block_logits_bwd(L.ptr(u), L.ptr(cols), L.dt(u), L.ptr(pos_col), L.ptr(row_cu),
                                          row_cu.numel() - 1, cols.shape[0], u.shape[1], max_len, scale, L.ptr(bias),
                                          L.ptr(own_lse), L.ptr(_f32(g_pos, "g_pos")), L.ptr(_f32(g_own, "g_own")),
                                          L.ptr(d_u), L.ptr(d_c), L.stream()), "rs_user_block_logits_bwd")
    return [d_u, d_c]


@user_block_logits_bwd_op.register_fake
def _(u, cols, pos_col, row_cu, max_len, scale, bias, own_lse, g_pos, g_own):
    return [u.new_empty(u.shape, dtype=torch.float32), cols.new_empty(cols.shape, dtype=torch.float32)]


class _UserBlockLogits(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u, cols, pos_col, row_cu, max_len, scale, bias, compute_dtype):
        from .losses import _cast16              # (shares the cast with the fused softmax of the same loss call)
        uc = u.detach() if compute_dtype is None else _cast16(u, compute_dtype)
        cc = cols.detach() if compute_dtype is None else _cast16(cols, compute_dtype)
        s_pos, own = L.direct.user_block_logits(uc, cc, pos_col, row_cu, max_len, scale, bias)
        ctx.save_for_backward(uc, cc, pos_col, row_cu, bias, own)
        ctx.meta = (max_len, scale, u.dtype, cols.dtype)
        return s_pos, own

    @staticmethod
    def backward(ctx, g_pos, g_own):
        uc, cc, pos_col, row_cu, bias, own = ctx.saved_tensors
        max_len, scale, udt, cdt = ctx.meta
        d_u, d_c = L.direct.user_block_logits_bwd(uc, cc, pos_col, row_cu, max_len, scale, bias, own,
                                                      g_pos.float().contiguous(), g_own.float().contiguous())
        return d_u.to(udt), d_c.to(cdt), None, None, None, None, None, None


def user_block_logits(u, cols, pos_col, row_cu, max_len, scale, bias=None, compute_dtype=None):
    """(s_pos[n], own_lse[n]): the label logit of every row and the log-sum-exp of its logits against the OTHER
    target items of the same user (rows grouped by user, `row_cu` offsets); differentiable in u and cols."""
    if bias is not None:
        bias = bias.detach().float().contiguous()
    return _UserBlockLogits.apply(u, cols, pos_col, row_cu, int(max_len), float(scale), bias, compute_dtype)


def mine_hard_negatives(u, v, key, k, hnm_threshold):
    """(scores, ids, avail): per-row top-k of <u_i, v_j> over non-ignored columns (C4/C5 mining, no gradient)."""
    with torch.no_grad():
        return L.direct.mine_hard_negatives(u.detach().float(), v.detach().float(), key, int(k),
                                                float(hnm_threshold))
