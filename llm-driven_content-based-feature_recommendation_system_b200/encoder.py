"""The sequence-encoder body of SASRecUserTower on PACKED valid tokens.

Reference: `nn.TransformerEncoder(nn.TransformerEncoderLayer(d_model=128, nhead=4, dim_feedforward=256,
dropout=0.2, activation="gelu", norm_first=True, batch_first=True), num_layers=2)` called with a causal mask and
`src_key_padding_mask` over the left-padded [B, L=50] grid (tower_code/v1_refine_usertower.py:343-352, :458-466).

A valid position attends only to valid positions of its own sequence and the rest of the layer is position-wise, so
the valid rows of the output depend on the valid rows alone.  `packed_encoder` therefore runs the SAME layers (same
parameters -- the nn.TransformerEncoder module stays the owner, state-dict names unchanged) on the packed valid
tokens [T, 128] (T ~ 25 % of B*L on H&M-shaped batches), delimited by `cu_seqlens`:

    h   = LN1(x)                                    rs::ln            (fp32 stats, activation-dtype output)
    qkv = in_proj(h)                                library GEMM (nn.Linear in the reference as well)
    o   = causal softmax(q k^T / sqrt(32)) v        rs::attn_varlen   (short-sequence kernel, dropout inside)
    x   = x + dropout(out_proj(o))                  GEMM + rs::dropout_add   (residual stream stays fp32)
    (the Linear biases are added inside the consuming kernel; their gradients are rs::colsum of its backward)
    f   = dropout(gelu(linear1(LN2(x))))            rs::ln, GEMM, rs::gelu_dropout
    x   = x + dropout(linear2(f))                   GEMM + rs::dropout_add

Values at valid positions equal the reference's in eval mode (tests/test_gpu_encoder.py); in train mode the dropout
masks come from a counter-based hash instead of Philox (same Bernoulli(1-p)/(1-p) law, different stream).
"""
from __future__ import annotations

import math
import weakref
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F
from torch import Tensor

from . import _lib as L
from . import ops

_lib = L.load()


def _seed() -> int:
    """A fresh 62-bit seed from torch's CPU generator (host-side: no device synchronisation; reproducible under
    torch.manual_seed)."""
    return int(torch.randint(0, 2 ** 62, (1,)).item())


def rng_advance() -> None:
    """Advance the device-side dropout epoch (stream-ordered; call once per train step)."""
    L.check(_lib.rs_rng_advance(L.stream()), "rs_rng_advance")


# ------------------------------------------------------------------------------------------------ custom ops
def _colsum(x: Tensor) -> Tensor:
    x = x.contiguous()
    n_cols = x.shape[-1]
    n_rows = x.numel() // n_cols
    out = torch.empty(n_cols, dtype=torch.float32, device=x.device)
    ws = L.workspace(_lib.rs_colsum_workspace_bytes(n_rows, n_cols), x.device)
    L.check(_lib.rs_colsum(L.ptr(x), L.dt(x), n_rows, n_cols, L.ptr(out), L.ptr(ws), ws.numel(), L.stream()), "rs_colsum")
    return out


def _fused_colsum_ok(n_cols: int) -> bool:
    """the elementwise backward kernels can fold the column sums in when a CTA's 256 threads tile the column groups"""
    return n_cols > 0 and n_cols % 4 == 0 and 256 % (n_cols // 4) == 0


@torch.library.custom_op("rs::colsum", mutates_args=())
def colsum_op(x: Tensor) -> Tensor:
    """fp32 column sums of a [rows, cols] matrix in a fixed order (bias gradients)."""
    L.require_cuda(x)
    return _colsum(x)


@colsum_op.register_fake
def _(x):
    return x.new_empty(x.shape[-1], dtype=torch.float32)


def _one_rows(one_rows: Optional[Tensor], cu_seqlens: Tensor, zero_tail: int, one_row_from: int) -> Optional[Tensor]:
    if one_rows is None or one_row_from < 0:
        return None
    assert one_rows.dtype == torch.int64 and one_rows.is_contiguous() and one_rows.is_cuda
    assert one_rows.numel() == cu_seqlens.numel() - 1 - zero_tail - one_row_from, "one row index per single-row sequence"
    return one_rows


@torch.library.custom_op("rs::attn_varlen", mutates_args=())
def attn_varlen_op(qkv: Tensor, bias: Optional[Tensor], cu_seqlens: Tensor, n_heads: int, max_len: int, zero_tail: int,
                   scale: float, dropout_p: float, seed: int, one_row_from: int = -1,
                   one_rows: Optional[Tensor] = None) -> List[Tensor]:
    L.require_cuda(qkv, cu_seqlens)
    qkv = qkv.contiguous()
    T = qkv.shape[0]
    hd = qkv.shape[1] // (3 * n_heads)
    out = torch.empty(T, n_heads * hd, dtype=qkv.dtype, device=qkv.device)
    lse = torch.empty(T, n_heads, dtype=torch.float32, device=qkv.device)
    L.check(_lib.rs_attn_varlen_fwd(L.ptr(qkv), L.dt(qkv), L.ptr(bias), L.ptr(cu_seqlens), cu_seqlens.numel() - 1, T,
                                    n_heads, hd, max_len, zero_tail, one_row_from, L.ptr(_one_rows(one_rows, cu_seqlens, zero_tail, one_row_from)),
                                    scale, dropout_p, seed, L.ptr(out), L.ptr(lse), L.stream()), "rs_attn_varlen_fwd")
    return [out, lse]


@attn_varlen_op.register_fake
def _(qkv, bias, cu_seqlens, n_heads, max_len, zero_tail, scale, dropout_p, seed, one_row_from=-1, one_rows=None):
    return [qkv.new_empty(qkv.shape[0], qkv.shape[1] // 3), qkv.new_empty(qkv.shape[0], n_heads, dtype=torch.float32)]


@torch.library.custom_op("rs::attn_varlen_bwd", mutates_args=())
def attn_varlen_bwd_op(qkv: Tensor, bias: Optional[Tensor], d_out: Tensor, out: Tensor, lse: Tensor, cu_seqlens: Tensor,
                       n_heads: int, max_len: int, zero_tail: int, scale: float, dropout_p: float,
                       seed: int, one_row_from: int = -1, one_rows: Optional[Tensor] = None) -> List[Tensor]:
    """[d_qkv, d_bias] (d_bias empty when there is no bias)."""
    d_out = d_out.to(qkv.dtype).contiguous()
    T = qkv.shape[0]
    hd = qkv.shape[1] // (3 * n_heads)
    d_qkv = torch.empty_like(qkv)
    L.check(_lib.rs_attn_varlen_bwd(L.ptr(qkv), L.ptr(d_out), L.ptr(out), L.dt(qkv), L.ptr(bias), L.ptr(lse),
                                    L.ptr(cu_seqlens), cu_seqlens.numel() - 1, T, n_heads, hd, max_len, zero_tail,
                                    one_row_from, L.ptr(_one_rows(one_rows, cu_seqlens, zero_tail, one_row_from)), scale,
                                    dropout_p, seed, L.ptr(d_qkv), L.stream()), "rs_attn_varlen_bwd")
    d_bias = _colsum(d_qkv) if bias is not None else qkv.new_empty(0, dtype=torch.float32)
    return [d_qkv, d_bias]


@attn_varlen_bwd_op.register_fake
def _(qkv, bias, d_out, out, lse, cu_seqlens, n_heads, max_len, zero_tail, scale, dropout_p, seed, one_row_from=-1,
      one_rows=None):
    return [torch.empty_like(qkv), qkv.new_empty(qkv.shape[1] if bias is not None else 0, dtype=torch.float32)]


@torch.library.custom_op("rs::ln", mutates_args=())
def ln_op(x: Tensor, index: Optional[Tensor], w: Tensor, b: Tensor, eps: float, dropout_p: float, seed: int,
          out_dtype: int) -> List[Tensor]:
    L.require_cuda(x, w, b)
    x = x.contiguous()
    n = x.shape[0] if index is None else index.numel()
    y = torch.empty(n, x.shape[1], dtype=L.torch_dtype(out_dtype), device=x.device)
    mean = torch.empty(n, dtype=torch.float32, device=x.device)
    rstd = torch.empty(n, dtype=torch.float32, device=x.device)
    L.check(_lib.rs_ln_fwd(L.ptr(x), L.dt(x), L.ptr(index), n, x.shape[1], L.ptr(w), L.ptr(b), eps, dropout_p, seed,
                           L.ptr(y), out_dtype, L.ptr(mean), L.ptr(rstd), L.stream()), "rs_ln_fwd")
    return [y, mean, rstd]


@ln_op.register_fake
def _(x, index, w, b, eps, dropout_p, seed, out_dtype):
    n = x.shape[0] if index is None else index.numel()
    return [x.new_empty(n, x.shape[1], dtype=L.torch_dtype(out_dtype)), x.new_empty(n, dtype=torch.float32),
            x.new_empty(n, dtype=torch.float32)]


@torch.library.custom_op("rs::ln_bwd", mutates_args=())
def ln_bwd_op(dy: Tensor, x: Tensor, index: Optional[Tensor], w: Tensor, mean: Tensor, rstd: Tensor, dropout_p: float,
              seed: int, res: Optional[Tensor] = None) -> List[Tensor]:
    """`res` (fp32, shaped like the packed dx): gradient reaching the LN input through the residual path, added in."""
    dy = dy.contiguous()
    if res is not None:
        res = res.float().contiguous()
    n = dy.shape[0]
    dx = torch.empty(n, x.shape[1], dtype=x.dtype, device=x.device)       # packed like dy
    dw = torch.empty_like(w)
    db = torch.empty_like(w)
    ws = L.workspace(_lib.rs_ln_bwd_workspace_bytes(n), x.device)
    L.check(_lib.rs_ln_bwd(L.ptr(dy), L.dt(dy), L.ptr(x), L.dt(x), L.ptr(index), n, x.shape[1], L.ptr(w), L.ptr(mean),
                           L.ptr(rstd), dropout_p, seed, L.ptr(dx), L.ptr(res), L.ptr(dw), L.ptr(db), L.ptr(ws), ws.numel(),
                           L.stream()), "rs_ln_bwd")
    return [dx, dw, db]


@ln_bwd_op.register_fake
def _(dy, x, index, w, mean, rstd, dropout_p, seed, res=None):
    return [x.new_empty(dy.shape[0], x.shape[1]), torch.empty_like(w), torch.empty_like(w)]


@torch.library.custom_op("rs::ln_act", mutates_args=())
def ln_act_op(x: Tensor, add: Optional[Tensor], add_rows: Optional[Tensor], lin_bias: Optional[Tensor], w: Tensor,
              b: Tensor, eps: float, act: int, dropout_p: float, seed: int, out_dtype: int) -> List[Tensor]:
    """dropout(act(LayerNorm(x + add[add_rows] + lin_bias))) for 128-wide rows -> [y, mean, rstd] (rs_ln_act_fwd)."""
    L.require_cuda(x, w, b)
    x = x.contiguous()
    n = x.shape[0]
    if add is not None:
        assert add.dtype == torch.float32 and add.is_contiguous() and add.shape[1] == x.shape[1]
        assert add_rows is None or (add_rows.dtype == torch.int64 and add_rows.numel() == n and add_rows.is_contiguous())
    y = torch.empty(n, x.shape[1], dtype=L.torch_dtype(out_dtype), device=x.device)
    mean = torch.empty(n, dtype=torch.float32, device=x.device)
    rstd = torch.empty(n, dtype=torch.float32, device=x.device)
    L.check(_lib.rs_ln_act_fwd(L.ptr(x), L.dt(x), L.ptr(add), L.ptr(add_rows), L.ptr(lin_bias), n, x.shape[1], L.ptr(w),
                               L.ptr(b), eps, act, dropout_p, seed, L.ptr(y), out_dtype, L.ptr(mean), L.ptr(rstd),
                               L.stream()), "rs_ln_act_fwd")
    return [y, mean, rstd]


@ln_act_op.register_fake
def _(x, add, add_rows, lin_bias, w, b, eps, act, dropout_p, seed, out_dtype):
    n = x.shape[0]
    return [x.new_empty(n, x.shape[1], dtype=L.torch_dtype(out_dtype)), x.new_empty(n, dtype=torch.float32),
            x.new_empty(n, dtype=torch.float32)]


@torch.library.custom_op("rs::ln_act_bwd", mutates_args=())
def ln_act_bwd_op(dy: Tensor, x: Tensor, add: Optional[Tensor], add_rows: Optional[Tensor], lin_bias: Optional[Tensor],
                  w: Tensor, b: Tensor, mean: Tensor, rstd: Tensor, act: int, dropout_p: float, seed: int) -> List[Tensor]:
    """-> [dx (x's dtype: gradient of the LN input of every row), d gamma, d beta] (rs_ln_act_bwd)."""
    dy = dy.contiguous()
    n = dy.shape[0]
    dx = torch.empty(n, x.shape[1], dtype=x.dtype, device=x.device)
    dw = torch.empty_like(w)
    db = torch.empty_like(w)
    ws = L.workspace(_lib.rs_ln_bwd_workspace_bytes(n), x.device)
    L.check(_lib.rs_ln_act_bwd(L.ptr(dy), L.dt(dy), L.ptr(x), L.dt(x), L.ptr(add), L.ptr(add_rows), L.ptr(lin_bias), n,
                               x.shape[1], L.ptr(w), L.ptr(b), L.ptr(mean), L.ptr(rstd), act, dropout_p, seed, L.ptr(dx),
                               L.ptr(dw), L.ptr(db), L.ptr(ws), ws.numel(), L.stream()), "rs_ln_act_bwd")
    return [dx, dw, db]


@ln_act_bwd_op.register_fake
def _(dy, x, add, add_rows, lin_bias, w, b, mean, rstd, act, dropout_p, seed):
    return [x.new_empty(dy.shape[0], x.shape[1]), torch.empty_like(w), torch.empty_like(w)]


@torch.library.custom_op("rs::emb_ln2", mutates_args=())
def emb_ln2_op(x: Tensor, index: Tensor, w0: Tensor, b0: Tensor, eps0: float, dropout_p: float, seed: int, w1: Tensor,
               b1: Tensor, eps1: float, out_dtype: int) -> List[Tensor]:
    """x0 = dropout(LN_emb(x[index])), h = LN_1(x0) -> [x0 fp32, h, mean0, rstd0, mean1, rstd1] (rs_emb_ln2_fwd)."""
    L.require_cuda(x, index, w0, w1)
    x = x.contiguous()
    n = index.numel()
    x0 = torch.empty(n, 128, dtype=torch.float32, device=x.device)
    h = torch.empty(n, 128, dtype=L.torch_dtype(out_dtype), device=x.device)
    st = [torch.empty(n, dtype=torch.float32, device=x.device) for _ in range(4)]
    L.check(_lib.rs_emb_ln2_fwd(L.ptr(x), L.dt(x), L.ptr(index), n, x.shape[1], L.ptr(w0), L.ptr(b0), eps0, dropout_p, seed,
                                L.ptr(w1), L.ptr(b1), eps1, L.ptr(x0), L.ptr(h), out_dtype, L.ptr(st[0]), L.ptr(st[1]),
                                L.ptr(st[2]), L.ptr(st[3]), L.stream()), "rs_emb_ln2_fwd")
    return [x0, h] + st


@emb_ln2_op.register_fake
def _(x, index, w0, b0, eps0, dropout_p, seed, w1, b1, eps1, out_dtype):
    n = index.numel()
    return [x.new_empty(n, 128, dtype=torch.float32), x.new_empty(n, 128, dtype=L.torch_dtype(out_dtype))] + \
           [x.new_empty(n, dtype=torch.float32) for _ in range(4)]


@torch.library.custom_op("rs::emb_ln2_bwd", mutates_args=())
def emb_ln2_bwd_op(dh: Tensor, x: Tensor, x0: Tensor, res: Optional[Tensor], inv1: Tensor, inv2: Tensor, w0: Tensor,
                   w1: Tensor, mean0: Tensor, rstd0: Tensor, mean1: Tensor, rstd1: Tensor, dropout_p: float,
                   seed: int) -> List[Tensor]:
    """-> [d x (x's dtype, one row per source row), d gamma_emb, d beta_emb, d gamma_1, d beta_1] (rs_emb_ln2_bwd)."""
    dh = dh.contiguous()
    if res is not None:
        res = res.float().contiguous()
    n_src = x.shape[0]
    dx = torch.empty_like(x)
    g = [torch.empty_like(w0) for _ in range(4)]
    ws = L.workspace(_lib.rs_emb_ln2_bwd_workspace_bytes(n_src), x.device)
    L.check(_lib.rs_emb_ln2_bwd(L.ptr(dh), L.dt(dh), L.ptr(x), L.dt(x), L.ptr(x0), L.ptr(res), L.ptr(inv1), L.ptr(inv2),
                                n_src, x.shape[1], L.ptr(w0), L.ptr(w1), L.ptr(mean0), L.ptr(rstd0), L.ptr(mean1),
                                L.ptr(rstd1), dropout_p, seed, L.ptr(dx), L.ptr(g[0]), L.ptr(g[1]), L.ptr(g[2]), L.ptr(g[3]),
                                L.ptr(ws), ws.numel(), L.stream()), "rs_emb_ln2_bwd")
    return [dx] + g


@emb_ln2_bwd_op.register_fake
def _(dh, x, x0, res, inv1, inv2, w0, w1, mean0, rstd0, mean1, rstd1, dropout_p, seed):
    return [torch.empty_like(x)] + [torch.empty_like(w0) for _ in range(4)]


@torch.library.custom_op("rs::dropout_add_ln", mutates_args=())
def dropout_add_ln_op(x: Tensor, y: Tensor, lin_bias: Optional[Tensor], w: Tensor, b: Tensor, eps: float,
                      dropout_p: float, seed: int, out_dtype: int) -> List[Tensor]:
    """x1 = x + dropout(y + lin_bias); h = LayerNorm(x1) -> [x1 fp32, h, mean, rstd] (rs_dropout_add_ln_fwd)."""
    L.require_cuda(x, y, w, b)
    assert x.dtype == torch.float32 and x.shape == y.shape and x.shape[1] == 128
    x, y = x.contiguous(), y.contiguous()
    n = x.shape[0]
    x1 = torch.empty_like(x)
    h = torch.empty(n, 128, dtype=L.torch_dtype(out_dtype), device=x.device)
    mean = torch.empty(n, dtype=torch.float32, device=x.device)
    rstd = torch.empty(n, dtype=torch.float32, device=x.device)
    L.check(_lib.rs_dropout_add_ln_fwd(L.ptr(x), L.ptr(y), L.dt(y), L.ptr(lin_bias), n, 128, dropout_p, seed, L.ptr(w),
                                       L.ptr(b), eps, L.ptr(x1), L.ptr(h), out_dtype, L.ptr(mean), L.ptr(rstd),
                                       L.stream()), "rs_dropout_add_ln_fwd")
    return [x1, h, mean, rstd]


@dropout_add_ln_op.register_fake
def _(x, y, lin_bias, w, b, eps, dropout_p, seed, out_dtype):
    n = x.shape[0]
    return [torch.empty_like(x), x.new_empty(n, 128, dtype=L.torch_dtype(out_dtype)), x.new_empty(n), x.new_empty(n)]


@torch.library.custom_op("rs::ln_bwd_dropout", mutates_args=())
def ln_bwd_dropout_op(dh: Tensor, x1: Tensor, res: Optional[Tensor], w: Tensor, mean: Tensor, rstd: Tensor,
                      dropout_p: float, seed: int, dy_dtype: int) -> List[Tensor]:
    """-> [dx fp32, dy, d gamma, d beta, d lin_bias] (rs_ln_bwd_dropout)."""
    dh = dh.contiguous()
    if res is not None:
        res = res.float().contiguous()
    n = dh.shape[0]
    dx = torch.empty_like(x1)
    dy = torch.empty(n, 128, dtype=L.torch_dtype(dy_dtype), device=x1.device)
    dw, db, dl = torch.empty_like(w), torch.empty_like(w), torch.empty_like(w)
    ws = L.workspace(_lib.rs_ln_bwd_dropout_workspace_bytes(n), x1.device)
    L.check(_lib.rs_ln_bwd_dropout(L.ptr(dh), L.dt(dh), L.ptr(x1), L.ptr(res), n, 128, L.ptr(w), L.ptr(mean), L.ptr(rstd),
                                   dropout_p, seed, L.ptr(dx), L.ptr(dy), dy_dtype, L.ptr(dw), L.ptr(db), L.ptr(dl),
                                   L.ptr(ws), ws.numel(), L.stream()), "rs_ln_bwd_dropout")
    return [dx, dy, dw, db, dl]


@ln_bwd_dropout_op.register_fake
def _(dh, x1, res, w, mean, rstd, dropout_p, seed, dy_dtype):
    return [torch.empty_like(x1), x1.new_empty(x1.shape, dtype=L.torch_dtype(dy_dtype)), torch.empty_like(w),
            torch.empty_like(w), torch.empty_like(w)]


@torch.library.custom_op("rs::dropout_add", mutates_args=())
def dropout_add_op(x: Tensor, y: Tensor, bias: Optional[Tensor], dropout_p: float, seed: int) -> Tensor:
    L.require_cuda(x, y)
    x, y = x.contiguous(), y.contiguous()
    out = torch.empty_like(x)
    L.check(_lib.rs_dropout_add_fwd(L.ptr(x), L.dt(x), L.ptr(y), L.dt(y), L.ptr(bias), x.shape[-1], x.numel(),
                                    dropout_p, seed, L.ptr(out), L.stream()), "rs_dropout_add_fwd")
    return out


@dropout_add_op.register_fake
def _(x, y, bias, dropout_p, seed):
    return torch.empty_like(x)


@torch.library.custom_op("rs::dropout_bwd", mutates_args=())
def dropout_bwd_op(g: Tensor, dropout_p: float, seed: int, out_dtype: int, want_colsum: bool) -> List[Tensor]:
    """[dy, colsum(dy)] -- the gradient of the dropped branch and (optionally) of the bias folded into it."""
    g = g.contiguous()
    dy = torch.empty(g.shape, dtype=L.torch_dtype(out_dtype), device=g.device)
    n_cols = g.shape[-1]
    if want_colsum and _fused_colsum_ok(n_cols):
        db = torch.empty(n_cols, dtype=torch.float32, device=g.device)
        ws = L.workspace(_lib.rs_ew_colsum_workspace_bytes(g.numel(), n_cols), g.device)
        L.check(_lib.rs_dropout_bwd_bias(L.ptr(g), L.dt(g), g.numel(), n_cols, dropout_p, seed, L.ptr(dy), out_dtype,
                                         L.ptr(db), L.ptr(ws), ws.numel(), L.stream()), "rs_dropout_bwd_bias")
        return [dy, db]
    L.check(_lib.rs_dropout_bwd(L.ptr(g), L.dt(g), g.numel(), dropout_p, seed, L.ptr(dy), out_dtype, L.stream()),
            "rs_dropout_bwd")
    return [dy, _colsum(dy) if want_colsum else g.new_empty(0, dtype=torch.float32)]


@dropout_bwd_op.register_fake
def _(g, dropout_p, seed, out_dtype, want_colsum):
    return [g.new_empty(g.shape, dtype=L.torch_dtype(out_dtype)),
            g.new_empty(g.shape[-1] if want_colsum else 0, dtype=torch.float32)]


@torch.library.custom_op("rs::gelu_dropout", mutates_args=())
def gelu_dropout_op(z: Tensor, bias: Optional[Tensor], dropout_p: float, seed: int) -> Tensor:
    L.require_cuda(z)
    z = z.contiguous()
    out = torch.empty_like(z)
    L.check(_lib.rs_gelu_dropout_fwd(L.ptr(z), L.dt(z), L.ptr(bias), z.shape[-1], z.numel(), dropout_p, seed,
                                     L.ptr(out), L.stream()), "rs_gelu_dropout_fwd")
    return out


@gelu_dropout_op.register_fake
def _(z, bias, dropout_p, seed):
    return torch.empty_like(z)


@torch.library.custom_op("rs::gelu_dropout_bwd", mutates_args=())
def gelu_dropout_bwd_op(z: Tensor, bias: Optional[Tensor], g: Tensor, dropout_p: float, seed: int,
                        want_colsum: bool = False) -> List[Tensor]:
    """[dz, column sums of dz] (the sums when a bias was folded in, or `want_colsum`: the bias then sits in z already)."""
    g = g.to(z.dtype).contiguous()
    dz = torch.empty_like(z)
    n_cols = z.shape[-1]
    want_colsum = want_colsum or bias is not None
    if want_colsum and _fused_colsum_ok(n_cols):
        db = torch.empty(n_cols, dtype=torch.float32, device=z.device)
        ws = L.workspace(_lib.rs_ew_colsum_workspace_bytes(z.numel(), n_cols), z.device)
        L.check(_lib.rs_gelu_dropout_bwd_bias(L.ptr(z), L.ptr(g), L.dt(z), L.ptr(bias), n_cols, z.numel(), dropout_p, seed,
                                              L.ptr(dz), L.ptr(db), L.ptr(ws), ws.numel(), L.stream()),
                "rs_gelu_dropout_bwd_bias")
        return [dz, db]
    L.check(_lib.rs_gelu_dropout_bwd(L.ptr(z), L.ptr(g), L.dt(z), L.ptr(bias), z.shape[-1], z.numel(), dropout_p, seed,
                                     L.ptr(dz), L.stream()), "rs_gelu_dropout_bwd")
    return [dz, _colsum(dz) if want_colsum else z.new_empty(0, dtype=torch.float32)]


@gelu_dropout_bwd_op.register_fake
def _(z, bias, g, dropout_p, seed, want_colsum=False):
    return [torch.empty_like(z), z.new_empty(z.shape[-1] if (bias is not None or want_colsum) else 0, dtype=torch.float32)]


# ------------------------------------------------------------------------------------------------ autograd
class _AttnVarlen(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, bias, cu_seqlens, n_heads, max_len, zero_tail, scale, dropout_p, seed, one_row_from, one_rows):
        out, lse = L.direct.attn_varlen(qkv, bias, cu_seqlens, n_heads, max_len, zero_tail, scale, dropout_p, seed,
                                        one_row_from, one_rows)
        ctx.save_for_backward(qkv, bias, out, lse, cu_seqlens, one_rows)
        ctx.meta = (n_heads, max_len, zero_tail, scale, dropout_p, seed, one_row_from)
        return out

    @staticmethod
    def backward(ctx, g):
        qkv, bias, out, lse, cu, one_rows = ctx.saved_tensors
        d_qkv, d_bias = L.direct.attn_varlen_bwd(qkv, bias, g, out, lse, cu, *ctx.meta, one_rows)
        return (d_qkv, d_bias if bias is not None else None) + (None,) * 9


def attn_varlen(qkv: Tensor, cu_seqlens: Tensor, n_heads: int, max_len: int, dropout_p: float = 0.0,
                scale: Optional[float] = None, zero_tail: int = 0, bias: Optional[Tensor] = None,
                one_row_from: int = -1, one_rows: Optional[Tensor] = None) -> Tensor:
    """Causal self-attention over packed sequences: qkv [T, 3*H*32] (in_proj output; its `bias` [3*H*32] may be
    passed separately and is added on load) -> [T, H*32].  The last `zero_tail` sequences are queries at padded
    positions (every key masked): output 0, no gradient.  `one_row_from` >= 0: the caller reads only ONE row of each
    sequence from that index on (before the zero tail) -- `one_rows` [that many] int64 packed row indices (None: the last
    token; an index outside its sequence: none); their other output rows are zeros and only that row's gradient is
    propagated (a matrix-vector kernel instead of the tile kernel)."""
    hd = qkv.shape[1] // (3 * n_heads)
    scale = 1.0 / math.sqrt(hd) if scale is None else scale
    return _AttnVarlen.apply(qkv, bias, cu_seqlens, n_heads, max_len, int(zero_tail), float(scale), float(dropout_p),
                             _seed() if dropout_p > 0 else 0, int(one_row_from), one_rows)


class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, index, w, b, eps, dropout_p, seed, out_dtype, fold=None, inv1=None, inv2=None):
        y, mean, rstd = L.direct.ln(x, index, w, b, eps, dropout_p, seed, out_dtype)
        ctx.save_for_backward(x, index, w, mean, rstd, inv1, inv2)
        ctx.meta = (dropout_p, seed)
        ctx.fold = fold
        return y

    @staticmethod
    def backward(ctx, g):
        x, index, w, mean, rstd, inv1, inv2 = ctx.saved_tensors
        dx, dw, db = L.direct.ln_bwd(g, x, index, w, mean, rstd, *ctx.meta)
        if index is not None and inv1 is not None:
            # every row of x is read by exactly two packed rows (inv1 / inv2: the two dropout views): the scatter-add
            # is a gather of two rows (static shapes, deterministic)
            dx = L.direct.gather_add2(dx, inv1, inv2)
        elif index is not None and ctx.fold is not None:
            # index == [0..T) twice, then [T..T+E) twice (two dropout views of the same packed rows): the scatter-add
            # is two elementwise sums (deterministic, no atomics)
            T, E = ctx.fold
            full = torch.empty_like(x)
            torch.add(dx[:T], dx[T:2 * T], out=full[:T])
            if E:
                torch.add(dx[2 * T:2 * T + E], dx[2 * T + E:2 * T + 2 * E], out=full[T:T + E])
            full[T + E:].zero_()
            dx = full
        elif index is not None:        # an index may repeat rows (one copy per dropout view): scatter-ADD
            dx = torch.zeros_like(x).index_add_(0, index, dx)
        return dx, None, dw, db, None, None, None, None, None, None, None


def layer_norm(x: Tensor, weight: Tensor, bias: Tensor, eps: float = 1e-5, index: Optional[Tensor] = None,
               dropout_p: float = 0.0, out_dtype: Optional[torch.dtype] = None, index_fold=None,
               index_inv=None) -> Tensor:
    """dropout(LayerNorm(x[index])) for 128-wide rows; `index` (int64) packs rows on the way in.  `index_fold` = (T, E):
    the caller vouches that index == cat(arange(T), arange(T), T + arange(E), T + arange(E)) (train.add_host_index's
    two-view layout), which turns the backward's scatter-add into two elementwise sums.  `index_inv` = (inv1, inv2)
    [len(x)] each: the caller vouches that row r of x is read by exactly the packed rows inv1[r] and inv2[r]
    (ops.batch_index_build's fold_inv1 / fold_inv2): the scatter-add becomes a two-row gather, shapes static."""
    od = L.dt(out_dtype) if out_dtype is not None else L.dt(x)
    if index_fold is not None:
        T, E = index_fold
        assert index is not None and index.numel() == 2 * (T + E) and x.shape[0] >= T + E
    inv1, inv2 = index_inv if index_inv is not None else (None, None)
    if inv1 is not None:
        assert index is not None and index.numel() == 2 * x.shape[0] and inv1.numel() == x.shape[0] == inv2.numel()
    return _LayerNorm.apply(x, index, weight, bias, float(eps), float(dropout_p), _seed() if dropout_p > 0 else 0, od,
                            index_fold, inv1, inv2)


class _LayerNormAct(torch.autograd.Function):
    """dropout(act(LayerNorm(x))) in one pass each way (static_mlp's LayerNorm -> GELU -> Dropout)."""

    @staticmethod
    def forward(ctx, x, w, b, eps, act, dropout_p, seed, out_dtype):
        y, mean, rstd = L.direct.ln_act(x, None, None, None, w, b, eps, act, dropout_p, seed, out_dtype)
        ctx.save_for_backward(x, w, b, mean, rstd)
        ctx.meta = (act, dropout_p, seed)
        return y

    @staticmethod
    def backward(ctx, g):
        x, w, b, mean, rstd = ctx.saved_tensors
        dx, dw, db = L.direct.ln_act_bwd(g, x, None, None, None, w, b, mean, rstd, *ctx.meta)
        return dx, dw, db, None, None, None, None, None


def layer_norm_act(x: Tensor, weight: Tensor, bias: Tensor, eps: float = 1e-5, act: str = "gelu",
                   dropout_p: float = 0.0, out_dtype: Optional[torch.dtype] = None) -> Tensor:
    """dropout(act(LayerNorm(x))) for 128-wide rows; act: "gelu" (exact) or "none"."""
    od = L.dt(out_dtype) if out_dtype is not None else L.dt(x)
    return _LayerNormAct.apply(x, weight, bias, float(eps), 1 if act == "gelu" else 0, float(dropout_p),
                               _seed() if dropout_p > 0 else 0, od)


# 16-bit copies of the GEMM weights for one train step.  Autocast casts every fp32 weight where it is used (one small
# kernel per weight per step) and autograd casts every weight gradient back (another one): ~35 launches of ~3 us in
# profiles/r02d_small_launches.txt.  `prepare_weights` makes all the copies with ONE multi-tensor launch at the start of
# the step; the GEMM wrappers below pick them up and hand fp32 weight gradients straight out of the GEMM.
_wcache: Dict[int, Tensor] = {}
_wlists = weakref.WeakKeyDictionary()        # module -> its GEMM weights / biases (collected once)


def prepare_weights(module: torch.nn.Module, dtype: Optional[torch.dtype]) -> None:
    """Refresh the 16-bit copies of `module`'s Linear / attention-projection weights and biases (call once per step,
    after the optimizer has updated them; `release_weights()` drops them)."""
    _wcache.clear()
    if dtype is None or dtype == torch.float32:
        return
    ws = _wlists.get(module)
    if ws is None:
        ws = []
        for m in module.modules():
            if isinstance(m, torch.nn.Linear):
                ws += [m.weight] + ([m.bias] if m.bias is not None else [])
            elif isinstance(m, torch.nn.MultiheadAttention) and m.in_proj_weight is not None:
                ws += [m.in_proj_weight] + ([m.in_proj_bias] if m.in_proj_bias is not None else [])
        ws = [w for w in ws if w.dtype == torch.float32 and w.is_cuda]
        _wlists[module] = ws
    if not ws:
        return
    with torch.no_grad():
        outs = [torch.empty(w.shape, dtype=dtype, device=w.device) for w in ws]
        torch._foreach_copy_(outs, ws)
    for w, o in zip(ws, outs):
        _wcache[id(w)] = o


def release_weights() -> None:
    _wcache.clear()


def _w16(weight: Tensor, dtype: torch.dtype) -> Tensor:
    """`weight` in `dtype`: the step's prepared copy when there is one, else a fresh cast (no autograd either way)."""
    if weight.dtype == dtype:
        return weight.detach()
    c = _wcache.get(id(weight))
    if c is not None and c.dtype == dtype:
        return c
    return weight.detach().to(dtype)


class _MatmulW(torch.autograd.Function):
    """x @ W^T for an fp32 parameter W and a 16-bit (or fp32) activation x: the GEMM runs in x's dtype on W's prepared
    copy; the weight gradient comes out of its GEMM in fp32 (no 16-bit rounding of a sum over ~10^5 rows, no cast)."""

    @staticmethod
    def forward(ctx, x, weight):
        wb = _w16(weight, x.dtype)
        ctx.save_for_backward(x, wb)
        ctx.wdt = weight.dtype
        return x @ wb.t()

    @staticmethod
    def backward(ctx, g):
        x, wb = ctx.saved_tensors
        g2 = g.reshape(-1, g.shape[-1])
        dx = (g2 @ wb).view(x.shape) if ctx.needs_input_grad[0] else None
        dw = _mm32(g2.t(), x.reshape(-1, x.shape[-1])).to(ctx.wdt) if ctx.needs_input_grad[1] else None
        return dx, dw


def matmul_w(x: Tensor, weight: Tensor) -> Tensor:
    """F.linear(x, weight) (no bias) -- see _MatmulW."""
    return _MatmulW.apply(x, weight)


def _mm32(a: Tensor, b: Tensor) -> Tensor:
    """a @ b with an fp32 result from 16-bit operands (weight gradients: no 16-bit rounding of the sum over ~10^5 rows)."""
    if a.dtype == torch.float32:
        return a @ b
    return torch.mm(a, b, out_dtype=torch.float32)


class _FusedHead(torch.autograd.Function):
    """GELU(LayerNorm(Linear(256 -> 128)(cat([rows, prof[users]], -1)))) -- the first three layers of the late-fusion head
    (v1_refine_usertower.py:394-399 on :499-502) -- without the concatenation: the Linear splits into W[:, :128] on the
    sequence rows and W[:, 128:] on the DISTINCT profile rows (one per user instead of one per time step; fp32 result),
    and rs::ln_act adds the two halves + the bias, normalises, applies the GELU and emits the next Linear's operand dtype.
    `users` [R] int64: the first `n_sorted` entries ascend (batch-major valid steps), the rest name distinct rows."""

    @staticmethod
    def forward(ctx, rows, prof, users, n_sorted, W, bias, ln_w, ln_b, eps):
        cd = rows.dtype
        Wc = _w16(W, cd)
        profc = prof.to(cd)
        h1 = rows @ Wc[:, :128].t()
        p2 = _mm32(profc, Wc[:, 128:].t())
        y, mean, rstd = L.direct.ln_act(h1, p2, users, bias, ln_w, ln_b, eps, 1, 0.0, 0, L.dt(rows))
        ctx.save_for_backward(rows, profc, users, Wc, bias, ln_w, ln_b, h1, p2, mean, rstd)
        ctx.meta = (n_sorted, prof.dtype, W.dtype)
        return y

    @staticmethod
    def backward(ctx, g):
        rows, profc, users, Wc, bias, ln_w, ln_b, h1, p2, mean, rstd = ctx.saved_tensors
        n, pdt, wdt = ctx.meta
        dpre, dlw, dlb = L.direct.ln_act_bwd(g, h1, p2, users, bias, ln_w, ln_b, mean, rstd, 1, 0.0, 0)
        d_rows = dpre @ Wc[:, :128] if ctx.needs_input_grad[0] else None
        # d p2 = per-user sums of dpre: a segment sum over the ascending prefix + one distinct row each for the rest
        dp2 = ops.segment_sum_sorted(dpre[:n], users[:n], p2.shape[0])
        ops.scatter_add_distinct_(dp2, dpre[n:], users[n:])
        dbias = _colsum(dp2)
        dp2c = dp2.to(rows.dtype)
        dW = torch.cat([_mm32(dpre.t(), rows), _mm32(dp2c.t(), profc)], dim=1).to(wdt)
        d_prof = (dp2c @ Wc[:, 128:]).to(pdt) if ctx.needs_input_grad[1] else None
        return d_rows, d_prof, None, None, dW, dbias.to(bias.dtype), dlw, dlb, None


def fused_head(rows: Tensor, prof: Tensor, users: Tensor, n_sorted: int, linear: torch.nn.Linear,
               norm: torch.nn.LayerNorm) -> Tensor:
    """see _FusedHead; `rows` [R, 128] in the operand dtype of the GEMMs (the autocast dtype), `prof` [U, 128]."""
    L.require_cuda(rows, prof)
    return _FusedHead.apply(rows.contiguous(), prof.contiguous(), users.contiguous(), int(n_sorted), linear.weight,
                            linear.bias, norm.weight, norm.bias, float(norm.eps))


class _ResidualLN(torch.autograd.Function):
    """(x, LN(x)) for a pre-norm residual block: x is handed through so that the gradient arriving on the residual path
    and the gradient of the LN branch meet in ONE backward kernel (ln_bwd adds the former to its dx) instead of an
    autograd accumulation pass over the [T, 128] fp32 stream."""

    @staticmethod
    def forward(ctx, x, w, b, eps, out_dtype):
        y, mean, rstd = L.direct.ln(x, None, w, b, eps, 0.0, 0, out_dtype)
        ctx.save_for_backward(x, w, mean, rstd)
        return x.view_as(x), y

    @staticmethod
    def backward(ctx, gx, gy):
        x, w, mean, rstd = ctx.saved_tensors
        if gy is None:
            return gx, None, None, None, None
        dx, dw, db = L.direct.ln_bwd(gy, x, None, w, mean, rstd, 0.0, 0, gx)
        return dx, dw, db, None, None


def residual_layer_norm(x: Tensor, weight: Tensor, bias: Tensor, eps: float, out_dtype: torch.dtype):
    """returns (x, LayerNorm(x)); use the returned x for the residual add (see _ResidualLN)."""
    return _ResidualLN.apply(x, weight, bias, float(eps), L.dt(out_dtype))


class _EmbLN2(torch.autograd.Function):
    """(x0, h): x0 = dropout(LayerNorm_emb(x[index])) -- the encoder's input stream -- and h = LayerNorm_1(x0), the first
    layer's normalised input, in one pass (rs_emb_ln2_fwd).  Every row of x is read by exactly the packed rows inv1 / inv2
    (two dropout views): the backward folds the two views BEFORE LayerNorm_emb's (linear) backward, one warp per source
    row (rs_emb_ln2_bwd)."""

    @staticmethod
    def forward(ctx, x, index, inv1, inv2, w0, b0, eps0, dropout_p, seed, w1, b1, eps1, out_dtype):
        x0, h, m0, r0, m1, r1 = L.direct.emb_ln2(x, index, w0, b0, eps0, dropout_p, seed, w1, b1, eps1, out_dtype)
        ctx.save_for_backward(x, x0, inv1, inv2, w0, w1, m0, r0, m1, r1)
        ctx.meta = (dropout_p, seed)
        return x0, h

    @staticmethod
    def backward(ctx, gx0, gh):
        x, x0, inv1, inv2, w0, w1, m0, r0, m1, r1 = ctx.saved_tensors
        if gh is None:
            gh = torch.zeros(x0.shape, dtype=torch.bfloat16, device=x0.device)
        dx, dw0, db0, dw1, db1 = L.direct.emb_ln2_bwd(gh, x, x0, gx0, inv1, inv2, w0, w1, m0, r0, m1, r1, *ctx.meta)
        return dx, None, None, None, dw0, db0, None, None, None, dw1, db1, None, None


def emb_layer_norm2(x: Tensor, index: Tensor, index_inv, emb_ln: torch.nn.LayerNorm, dropout_p: float,
                    norm1: torch.nn.LayerNorm, out_dtype: torch.dtype):
    """returns (x0, h) -- see _EmbLN2.  `index` [2 * len(x)], `index_inv` = (inv1, inv2) as for layer_norm(index_inv=...)."""
    inv1, inv2 = index_inv
    assert index.numel() == 2 * x.shape[0] and inv1.numel() == x.shape[0] == inv2.numel()
    return _EmbLN2.apply(x, index, inv1, inv2, emb_ln.weight, emb_ln.bias, float(emb_ln.eps), float(dropout_p),
                         _seed() if dropout_p > 0 else 0, norm1.weight, norm1.bias, float(norm1.eps), L.dt(out_dtype))


class _DropoutAddLN(torch.autograd.Function):
    """(x1, h) with x1 = x + dropout(y + bias), h = LayerNorm(x1): the end of one pre-norm block and the start of the next
    in one pass each way (rs_dropout_add_ln_fwd / rs_ln_bwd_dropout).  The gradient reaching x1 through the residual
    path and the gradient of the LN branch meet in the one backward kernel, which also emits dy and the three parameter
    gradients."""

    @staticmethod
    def forward(ctx, x, y, bias, w, b, eps, dropout_p, seed, out_dtype):
        x1, h, mean, rstd = L.direct.dropout_add_ln(x, y, bias, w, b, eps, dropout_p, seed, out_dtype)
        ctx.save_for_backward(x1, w, mean, rstd)
        ctx.meta = (dropout_p, seed, L.dt(y), bias is not None)
        return x1, h

    @staticmethod
    def backward(ctx, gx1, gh):
        x1, w, mean, rstd = ctx.saved_tensors
        p, seed, ydt, has_bias = ctx.meta
        if gh is None:                              # the LN branch is unused: only the residual path carries a gradient
            dy, dl = L.direct.dropout_bwd(gx1, p, seed, ydt, has_bias)
            return gx1, dy, (dl if has_bias else None), None, None, None, None, None, None
        dx, dy, dw, db, dl = L.direct.ln_bwd_dropout(gh, x1, gx1, w, mean, rstd, p, seed, ydt)
        return dx, dy, (dl if has_bias else None), dw, db, None, None, None, None


def dropout_add_layer_norm(x: Tensor, y: Tensor, dropout_p: float, bias: Optional[Tensor], weight: Tensor,
                           ln_bias: Tensor, eps: float, out_dtype: torch.dtype):
    """returns (x + dropout(y + bias), LayerNorm of that); x: the fp32 residual stream [T, 128]."""
    return _DropoutAddLN.apply(x, y, bias, weight, ln_bias, float(eps), float(dropout_p),
                               _seed() if dropout_p > 0 else 0, L.dt(out_dtype))


class _DropoutAdd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, bias, dropout_p, seed):
        ctx.meta = (dropout_p, seed, L.dt(y), bias is not None)
        return L.direct.dropout_add(x, y, bias, dropout_p, seed)

    @staticmethod
    def backward(ctx, g):
        p, seed, ydt, has_bias = ctx.meta
        dy, db = L.direct.dropout_bwd(g, p, seed, ydt, has_bias)
        return g, dy, (db if has_bias else None), None, None


def dropout_add(x: Tensor, y: Tensor, dropout_p: float = 0.0, bias: Optional[Tensor] = None) -> Tensor:
    """x + dropout(y + bias) in x's dtype (the fp32 residual stream)."""
    return _DropoutAdd.apply(x, y, bias, float(dropout_p), _seed() if dropout_p > 0 else 0)


class _GeluDropout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, bias, dropout_p, seed):
        ctx.save_for_backward(z, bias)
        ctx.meta = (dropout_p, seed)
        return L.direct.gelu_dropout(z, bias, dropout_p, seed)

    @staticmethod
    def backward(ctx, g):
        z, bias = ctx.saved_tensors
        dz, db = L.direct.gelu_dropout_bwd(z, bias, g, *ctx.meta)
        return dz, (db if bias is not None else None), None, None


class _LinearGeluDropout(torch.autograd.Function):
    """dropout(gelu(x @ W^T + b)) for an fp32 nn.Linear under a 16-bit activation x: the bias rides in the GEMM's epilogue
    (one rounding of acc + b, as the reference's Linear), the GELU / dropout kernel has no bias work, and the bias gradient
    is still the column sum that the GELU backward kernel accumulates while it writes dz."""

    @staticmethod
    def forward(ctx, x, weight, bias, dropout_p, seed):
        wb, bb = _w16(weight, x.dtype), _w16(bias, x.dtype)
        with torch.autocast("cuda", enabled=False):
            z = torch.addmm(bb, x, wb.t())
        ctx.save_for_backward(x, wb, z)
        ctx.meta = (dropout_p, seed, weight.dtype, bias.dtype)
        return L.direct.gelu_dropout(z, None, dropout_p, seed)

    @staticmethod
    def backward(ctx, g):
        x, wb, z = ctx.saved_tensors
        p, seed, wdt, bdt = ctx.meta
        dz, db = L.direct.gelu_dropout_bwd(z, None, g, p, seed, True)
        dx = dz @ wb if ctx.needs_input_grad[0] else None
        dw = _mm32(dz.t(), x).to(wdt) if ctx.needs_input_grad[1] else None
        return dx, dw, db.to(bdt), None, None


def linear_gelu_dropout(x: Tensor, lin: torch.nn.Linear, dropout_p: float = 0.0) -> Tensor:
    """dropout(gelu(lin(x))) -- see _LinearGeluDropout; x [rows, in_features] in the GEMM's operand dtype."""
    return _LinearGeluDropout.apply(x, lin.weight, lin.bias, float(dropout_p), _seed() if dropout_p > 0 else 0)


def gelu_dropout(z: Tensor, dropout_p: float = 0.0, bias: Optional[Tensor] = None) -> Tensor:
    """dropout(gelu(z + bias))"""
    return _GeluDropout.apply(z, bias, float(dropout_p), _seed() if dropout_p > 0 else 0)


# ------------------------------------------------------------------------------------------------ composition
def _act_dtype(x: Tensor) -> torch.dtype:
    if torch.is_autocast_enabled("cuda"):
        return torch.get_autocast_dtype("cuda")
    return x.dtype


@torch.library.custom_op("rs::l2_normalize", mutates_args=())
def l2_normalize_op(x: Tensor, eps: float) -> List[Tensor]:
    """[y fp32, inv_norm] for 128-wide rows."""
    L.require_cuda(x)
    x = x.contiguous()
    n = x.numel() // 128
    y = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    inv = torch.empty(n, dtype=torch.float32, device=x.device)
    L.check(_lib.rs_l2_normalize_fwd(L.ptr(x), L.dt(x), n, 128, eps, L.ptr(y), L.RS_F32, L.ptr(inv), L.stream()),
            "rs_l2_normalize_fwd")
    return [y, inv]


@l2_normalize_op.register_fake
def _(x, eps):
    return [x.new_empty(x.shape, dtype=torch.float32), x.new_empty(x.numel() // 128, dtype=torch.float32)]


@torch.library.custom_op("rs::l2_normalize_bwd", mutates_args=())
def l2_normalize_bwd_op(g: Tensor, y: Tensor, inv: Tensor, dx_dtype: int) -> Tensor:
    g = g.contiguous()
    dx = torch.empty(y.shape, dtype=L.torch_dtype(dx_dtype), device=y.device)
    L.check(_lib.rs_l2_normalize_bwd(L.ptr(g), L.dt(g), L.ptr(y), L.RS_F32, L.ptr(inv), inv.numel(), 128, L.ptr(dx),
                                     dx_dtype, L.stream()), "rs_l2_normalize_bwd")
    return dx


@l2_normalize_bwd_op.register_fake
def _(g, y, inv, dx_dtype):
    return y.new_empty(y.shape, dtype=L.torch_dtype(dx_dtype))


class _L2Normalize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, eps):
        y, inv = L.direct.l2_normalize(x, eps)
        ctx.save_for_backward(y, inv)
        ctx.xdt = L.dt(x)
        return y

    @staticmethod
    def backward(ctx, g):
        y, inv = ctx.saved_tensors
        return L.direct.l2_normalize_bwd(g, y, inv, ctx.xdt), None


def l2_normalize(x: Tensor, eps: float = 1e-12) -> Tensor:
    """F.normalize(x, p=2, dim=-1) for 128-wide rows in one pass each way (fp32 output, as under autocast where the
    norm runs in fp32).  Other widths / dtypes are not on the path: the stock op serves them (on the device).  CPU tensors
    are rejected like by every other entry point of this package (no CPU path)."""
    L.require_cuda(x)
    if x.shape[-1] != 128 or x.dtype not in (torch.float32, torch.bfloat16, torch.float16):
        return F.normalize(x, p=2, dim=-1, eps=eps)
    return _L2Normalize.apply(x, float(eps))


class _LinearColsumBias(torch.autograd.Function):
    """y = x W^T + b as nn.Linear computes it (one library GEMM with the bias in its epilogue; under autocast in the
    autocast dtype); the only difference is the bias gradient: rs::colsum of the incoming gradient (fixed order, one
    pass) instead of the generic reduction autograd launches behind every Linear."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.xdt, ctx.wdt = x.dtype, weight.dtype
        ctx.k = x.shape[-1]
        ad = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else None
        if x.shape[-1] % 8 and ad in (torch.bfloat16, torch.float16):
            # a 16-bit operand whose rows are not 16-byte aligned (static_mlp's Linear(100 -> 128)) sends the library to
            # an sm_80 `align2` GEMM, and the fp32 detour of round 1 to an sm_80 SIMT sgemm (profiles/r01e_launches.md,
            # r02h_launches.md: 0.05 ms for a 0.4 GFLOP weight gradient).  Zero-pad the contraction dimension to a multiple
            # of 8 instead (100 -> 104: a 3 MB copy): all three GEMMs of the layer run on the tensor cores, same values.
            kp = (x.shape[-1] + 7) // 8 * 8
            lead = x.shape[:-1]
            x = F.pad(x.reshape(-1, x.shape[-1]).to(ad), (0, kp - ctx.k))
            weight = F.pad(_w16(weight, ad), (0, kp - ctx.k))
            with torch.autocast("cuda", enabled=False):
                y = F.linear(x, weight, _w16(bias, ad)).view(*lead, weight.shape[0])
        elif x.shape[-1] % 8:
            with torch.autocast("cuda", enabled=False):
                x = x.float()
                y = F.linear(x, weight.float(), bias.float())
        else:
            if ad is not None:
                x = x.to(ad)                                 # the cast autocast applies; kept for the backward (one cast, not two)
            weight = _w16(weight, x.dtype)                   # the step's prepared 16-bit copies (no per-use cast kernels)
            with torch.autocast("cuda", enabled=False):
                y = F.linear(x, weight, _w16(bias, x.dtype))
        ctx.save_for_backward(x, weight)
        ctx.bias_dtype = bias.dtype
        return y

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        k = ctx.k
        g2 = g.reshape(-1, g.shape[-1])
        x2 = x.reshape(-1, x.shape[-1]).to(g2.dtype)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = (g2 @ weight.to(g2.dtype))[:, :k]
            dx = dx.reshape(*g.shape[:-1], k).to(ctx.xdt)
        if ctx.needs_input_grad[1]:
            dw = _mm32(g2.t(), x2)[:, :k].to(ctx.wdt)
        db = _colsum(g2).to(ctx.bias_dtype) if ctx.needs_input_grad[2] else None
        return dx, dw, db


def linear(module: torch.nn.Linear, x: Tensor) -> Tensor:
    """`module(x)` for a stock nn.Linear with its bias gradient by rs::colsum (training); the plain module call when
    there is no bias gradient to take or the shape is outside rs::colsum's range.  CUDA tensors only."""
    L.require_cuda(x)
    if module.bias is None or not torch.is_grad_enabled() or x.shape[-1] % 4 or module.out_features % 4 \
            or module.out_features > 1024:
        return module(x)
    return _LinearColsumBias.apply(x, module.weight, module.bias)


def _is_ln128(m, x) -> bool:
    return (isinstance(m, torch.nn.LayerNorm) and tuple(m.normalized_shape) == (128,) and m.elementwise_affine
            and m.bias is not None and x.dtype in (torch.float32, torch.bfloat16, torch.float16))


def sequential(seq: torch.nn.Sequential, x: Tensor) -> Tensor:
    """`seq(x)` with its nn.Linear members routed through `linear` and LayerNorm(128) [-> GELU [-> Dropout]] runs through
    the row kernels (same modules, same parameters).  CUDA tensors only."""
    L.require_cuda(x)
    mods = list(seq)
    i = 0
    while i < len(mods):
        m = mods[i]
        if isinstance(m, torch.nn.Linear):
            x = linear(m, x)
        elif _is_ln128(m, x):
            nxt = mods[i + 1] if i + 1 < len(mods) else None
            if isinstance(nxt, torch.nn.GELU) and nxt.approximate == "none":
                # LayerNorm -> GELU [-> Dropout] in one pass (fp32 statistics, GELU in fp32 as under autocast).  When a
                # Linear follows under autocast the result is emitted in its operand dtype (the cast it would apply)
                i += 1
                p = 0.0
                after = mods[i + 1] if i + 1 < len(mods) else None
                if isinstance(after, torch.nn.Dropout):
                    p = after.p if after.training else 0.0
                    i += 1
                    after = mods[i + 1] if i + 1 < len(mods) else None
                od = torch.float32
                if isinstance(after, torch.nn.Linear) and torch.is_autocast_enabled("cuda"):
                    od = torch.get_autocast_dtype("cuda")
                x = layer_norm_act(x.reshape(-1, 128), m.weight, m.bias, m.eps, "gelu", p, od).view(*x.shape)
            else:
                # the row kernel (fp32 statistics and fp32 output, as torch's layer_norm under autocast; its backward
                # folds d_gamma / d_beta into the same pass)
                x = layer_norm(x.reshape(-1, 128), m.weight, m.bias, m.eps, out_dtype=torch.float32).view(*x.shape)
        else:
            x = m(x)
        i += 1
    return x


import os as _os
# in_proj / linear1: the bias in the library GEMM's epilogue (True) or folded into the kernel that consumes the GEMM output
# (False; the round-1 layout, kept for A/B runs: RS_BIAS_IN_GEMM=0).  The consumers of these two GEMMs -- the attention
# tiles and the GELU kernel -- are instruction-bound, so the add belongs in the epilogue: 7.88 -> 7.75 ms per step.
BIAS_IN_GEMM = _os.environ.get('RS_BIAS_IN_GEMM', '1') == '1'


def first_layer_fusable(encoder: torch.nn.TransformerEncoder, emb_ln: torch.nn.LayerNorm, x: Tensor) -> bool:
    """can the embedding LayerNorm be fused with the first layer's LayerNorm (emb_layer_norm2)?"""
    l0 = encoder.layers[0] if len(encoder.layers) else None
    return (l0 is not None and l0.norm_first and tuple(emb_ln.normalized_shape) == (128,) and emb_ln.elementwise_affine
            and emb_ln.bias is not None and l0.norm1.elementwise_affine and l0.norm1.bias is not None
            and x.shape[-1] == 128 and x.dtype in (torch.float32, torch.bfloat16, torch.float16))


def _add_norm(x: Tensor, pending, norm: torch.nn.LayerNorm, ad: torch.dtype):
    """(residual stream, LayerNorm of it); `pending` = (y, bias, p): a block's closing  x + dropout(y + bias)  that has not
    been applied yet -- it is then folded into the same pass as the LayerNorm (fp32 residual stream only)."""
    if pending is None:
        return residual_layer_norm(x, norm.weight, norm.bias, norm.eps, ad)
    y, bias, p = pending
    if x.dtype != torch.float32:
        return residual_layer_norm(dropout_add(x, y, p, bias=bias), norm.weight, norm.bias, norm.eps, ad)
    return dropout_add_layer_norm(x, y, p, bias, norm.weight, norm.bias, norm.eps, ad)


def packed_encoder_layer(layer: torch.nn.TransformerEncoderLayer, x: Tensor, cu_seqlens: Tensor, max_len: int,
                         zero_tail: int = 0, rows=None, one_row_from: int = -1, one_rows=None, pending=None,
                         defer_close: bool = False, first_h: Optional[Tensor] = None):
    """One pre-norm layer on the packed tokens.  `rows` = (n_prefix, idx): only the rows cat([arange(n_prefix), idx]) are
    wanted from this layer (idx: rows outside the prefix, -1 = none -> a zero row).  Attention still sees every token --
    the wanted rows attend to the others -- but everything behind it is position-wise, so the out-projection, the residual
    adds, the second LayerNorm and the feed-forward block run on the wanted rows alone; returns [n_prefix + len(idx), 128].
    `pending`: the previous layer's closing residual add, not yet applied (see _add_norm); `defer_close`: return
    (x, pending) instead of applying this layer's own closing add, for the next layer to fold into its first LayerNorm."""
    if not layer.norm_first or layer.activation_relu_or_gelu != 2:
        raise NotImplementedError("packed encoder: pre-norm GELU layers only (the reference's configuration)")
    tr = layer.training
    attn = layer.self_attn
    ad = _act_dtype(x)
    if first_h is not None:            # LayerNorm_1(x) came with x out of the fused embedding pass (emb_layer_norm2)
        h = first_h
    else:
        x, h = _add_norm(x, pending, layer.norm1, ad)
    # the four biases are folded into the kernels that consume the GEMM outputs: plain matmuls, and the bias
    # gradients are column sums of tensors those kernels' backward passes produce (no reduction behind each GEMM)
    # (in_proj: the bias rides in the GEMM's epilogue -- free there, one add per fragment load in the attention kernels
    # otherwise -- and its gradient is still the column sum of d_qkv, taken by the Linear wrapper's backward)
    if BIAS_IN_GEMM and attn.in_proj_bias is not None:
        qkv, qkv_bias = _LinearColsumBias.apply(h, attn.in_proj_weight, attn.in_proj_bias), None
    else:
        qkv, qkv_bias = matmul_w(h, attn.in_proj_weight), attn.in_proj_bias
    o = attn_varlen(qkv, cu_seqlens, attn.num_heads, max_len, attn.dropout if tr else 0.0, zero_tail=zero_tail,
                    bias=qkv_bias, one_row_from=one_row_from if rows is not None else -1,
                    one_rows=one_rows if rows is not None else None)
    if rows is not None:
        o = ops.select_prefix_rows(o, rows[0], rows[1], disjoint=True)
        x = ops.select_prefix_rows(x, rows[0], rows[1], disjoint=True)
    x, h = _add_norm(x, (matmul_w(o, attn.out_proj.weight), attn.out_proj.bias, layer.dropout1.p if tr else 0.0),
                     layer.norm2, ad)
    if BIAS_IN_GEMM and layer.linear1.bias is not None and h.dtype != torch.float32:
        f = linear_gelu_dropout(h, layer.linear1, layer.dropout.p if tr else 0.0)
    else:
        f = gelu_dropout(matmul_w(h, layer.linear1.weight), layer.dropout.p if tr else 0.0, bias=layer.linear1.bias)
    close = (matmul_w(f, layer.linear2.weight), layer.linear2.bias, layer.dropout2.p if tr else 0.0)
    if defer_close:
        return x, close
    return dropout_add(x, close[0], close[2], bias=close[1])


def packed_encoder(encoder: torch.nn.TransformerEncoder, x: Tensor, cu_seqlens: Tensor, max_len: int,
                   zero_tail: int = 0, last_rows=None, one_row_from: int = -1, one_rows=None,
                   first_h: Optional[Tensor] = None) -> Tensor:
    """x: [T, 128] fp32 residual stream of the packed valid tokens -> same shape; with `last_rows` = (n_prefix, idx) only
    those rows of the LAST layer's output, [n_prefix + len(idx), 128] (see packed_encoder_layer).  `one_row_from` /
    `one_rows` (with last_rows): the caller vouches that of the sequences from that index on, `last_rows` names nothing
    but the token `one_rows` gives (or tokens whose value does not matter) -- the last layer's attention then computes
    only that row of them (attn_varlen)."""
    n_layers = len(encoder.layers)
    pending = None
    for i, layer in enumerate(encoder.layers):
        last = i == n_layers - 1
        out = packed_encoder_layer(layer, x, cu_seqlens, max_len, zero_tail, last_rows if last else None,
                                   one_row_from if last else -1, one_rows if last else None, pending=pending,
                                   defer_close=not last, first_h=first_h if i == 0 else None)
        x, pending = out if not last else (out, None)
    if encoder.norm is not None:
        x = encoder.norm(x)
    return x
