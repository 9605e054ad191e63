"""GPU parity of the packed sequence encoder (encoder.py / csrc/encoder.cu) against stock PyTorch and against the
reference-generated user-tower fixture.  fp32 cases: 1e-4 / 1e-3 (different summation order); bf16: stated per test."""
import math
from types import SimpleNamespace

import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _cu(lens):
    cu = torch.zeros(len(lens) + 1, dtype=torch.int32)
    cu[1:] = torch.cumsum(torch.tensor(lens), 0)
    return cu


def _ref_attention(qkv, lens, H):
    """dense causal attention per sequence, fp32"""
    outs, t0 = [], 0
    for n in lens:
        x = qkv[t0:t0 + n].view(n, 3, H, -1)
        q, k, v = x[:, 0].transpose(0, 1), x[:, 1].transpose(0, 1), x[:, 2].transpose(0, 1)      # [H, n, hd]
        s = q @ k.transpose(1, 2) / math.sqrt(q.shape[-1])
        s = s.masked_fill(torch.triu(torch.ones(n, n, dtype=torch.bool, device=s.device), 1), float("-inf"))
        outs.append((torch.softmax(s, -1) @ v).transpose(0, 1).reshape(n, -1))
        t0 += n
    return torch.cat(outs)


@pytest.mark.parametrize("lens", [[1], [1, 2, 3], [50, 7, 33, 32, 64, 1, 17], [13] * 40])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_attn_varlen_vs_dense(rs, lens, dtype):
    H, hd = 4, 32
    g = torch.Generator().manual_seed(sum(lens))
    T = sum(lens)
    qkv32 = torch.randn(T, 3 * H * hd, generator=g)
    w = torch.randn(T, H * hd, generator=g)
    a = qkv32.to(dtype).float().to(DEV).requires_grad_(True)               # same rounded operands on both sides
    want = _ref_attention(a, lens, H)
    (want * w.to(DEV)).sum().backward()
    b = qkv32.to(dtype).to(DEV).requires_grad_(True)
    got = rs.encoder.attn_varlen(b, _cu(lens).to(DEV), H, max(lens))
    (got.float() * w.to(DEV)).sum().backward()
    tol = dict(rtol=1e-4, atol=1e-4) if dtype == torch.float32 else dict(rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(got.float(), want.detach(), **tol)
    torch.testing.assert_close(b.grad.float(), a.grad, **(tol if dtype == torch.float32 else dict(rtol=3e-2, atol=6e-2)))


def test_folded_biases_and_colsum(rs):
    """biases folded into attention / dropout_add / gelu_dropout: values and bias gradients vs torch."""
    H, hd, lens = 4, 32, [7, 20, 3, 41]
    g = torch.Generator().manual_seed(5)
    T = sum(lens)
    qkv = torch.randn(T, 3 * H * hd, generator=g).to(DEV)
    bias = torch.randn(3 * H * hd, generator=g).to(DEV)
    w = torch.randn(T, H * hd, generator=g).to(DEV)
    a, ba = qkv.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    (_ref_attention(a + ba, lens, H) * w).sum().backward()
    b, bb = qkv.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    got = rs.encoder.attn_varlen(b, _cu(lens).to(DEV), H, max(lens), bias=bb)
    (got * w).sum().backward()
    torch.testing.assert_close(got.detach(), _ref_attention(qkv + bias, lens, H), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(b.grad, a.grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(bb.grad, ba.grad, rtol=1e-4, atol=1e-3)
    for n_rows, n_cols in ((1, 128), (1000, 256), (70001, 384)):
        x = torch.randn(n_rows, n_cols, generator=g).to(DEV)
        torch.testing.assert_close(torch.ops.rs.colsum(x), x.sum(0), rtol=1e-4, atol=1e-3)
        torch.testing.assert_close(torch.ops.rs.colsum(x.bfloat16()), x.bfloat16().float().sum(0), rtol=1e-4, atol=1e-3)
    x = torch.randn(500, 128, generator=g).to(DEV).requires_grad_(True)
    y = torch.randn(500, 128, generator=g).to(DEV).requires_grad_(True)
    b1 = torch.randn(128, generator=g).to(DEV).requires_grad_(True)
    out = rs.encoder.dropout_add(x, y, 0.0, bias=b1)
    torch.testing.assert_close(out.detach(), (x + y + b1).detach())
    w2 = torch.randn(500, 128, generator=g).to(DEV)
    (out * w2).sum().backward()
    torch.testing.assert_close(b1.grad, w2.sum(0), rtol=1e-4, atol=1e-3)
    z = torch.randn(500, 256, generator=g).to(DEV).requires_grad_(True)
    b2 = torch.randn(256, generator=g).to(DEV).requires_grad_(True)
    f = rs.encoder.gelu_dropout(z, 0.0, bias=b2)
    torch.testing.assert_close(f.detach(), F.gelu((z + b2).detach()), rtol=1e-5, atol=1e-6)
    f.sum().backward()
    z2, b3 = z.detach().clone().requires_grad_(True), b2.detach().clone().requires_grad_(True)
    F.gelu(z2 + b3).sum().backward()
    torch.testing.assert_close(z.grad, z2.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(b2.grad, b3.grad, rtol=1e-4, atol=1e-3)


def test_attn_zero_tail_sequences(rs):
    """pseudo-sequences at padded positions: output 0, gradient 0, the others untouched."""
    H, hd, lens = 4, 32, [5, 9, 1, 1]
    g = torch.Generator().manual_seed(3)
    qkv = torch.randn(sum(lens), 3 * H * hd, generator=g).to(DEV).requires_grad_(True)
    out = rs.encoder.attn_varlen(qkv, _cu(lens).to(DEV), H, 16, zero_tail=2)
    ref = _ref_attention(qkv.detach(), lens[:2], H)
    torch.testing.assert_close(out[:14].detach(), ref, rtol=1e-4, atol=1e-4)
    assert (out[14:] == 0).all()
    out.sum().backward()
    assert (qkv.grad[14:] == 0).all() and qkv.grad[:14].abs().sum() > 0


def test_attn_dropout_forward_backward_share_the_mask(rs):
    """with an explicit seed the op is a deterministic function: its analytic gradient must match central differences
    of the SAME seeded forward (fp32), i.e. forward, dQ pass and dK/dV pass all draw the same keep mask."""
    H, hd, lens = 2, 32, [5, 37, 2]
    g = torch.Generator().manual_seed(0)
    T = sum(lens)
    qkv = (0.5 * torch.randn(T, 3 * H * hd, generator=g)).to(DEV)
    w = torch.randn(T, H * hd, generator=g).to(DEV)
    cu, seed, p, scale = _cu(lens).to(DEV), 1234567, 0.3, 1 / math.sqrt(hd)
    out, lse = torch.ops.rs.attn_varlen(qkv, None, cu, H, 64, 0, scale, p, seed)
    out0, _ = torch.ops.rs.attn_varlen(qkv, None, cu, H, 64, 0, scale, 0.0, 0)
    assert not torch.allclose(out, out0)                                   # dropout did something
    out_again, _ = torch.ops.rs.attn_varlen(qkv, None, cu, H, 64, 0, scale, p, seed)
    assert torch.equal(out, out_again)
    dq, _ = torch.ops.rs.attn_varlen_bwd(qkv, None, w, out, lse, cu, H, 64, 0, scale, p, seed)
    idx = torch.randint(0, qkv.numel(), (40,), generator=g)
    eps = 1e-2
    for flat in idx.tolist():
        d = torch.zeros_like(qkv).view(-1)
        d[flat] = eps
        fp, _ = torch.ops.rs.attn_varlen(qkv + d.view_as(qkv), None, cu, H, 64, 0, scale, p, seed)
        fm, _ = torch.ops.rs.attn_varlen(qkv - d.view_as(qkv), None, cu, H, 64, 0, scale, p, seed)
        num = ((fp - fm) * w).sum().item() / (2 * eps)
        assert abs(num - dq.view(-1)[flat].item()) < 2e-2 * max(1.0, abs(num)), (flat, num, dq.view(-1)[flat].item())


@pytest.mark.parametrize("xdt,ydt", [(torch.float32, torch.float32), (torch.float32, torch.bfloat16),
                                     (torch.bfloat16, torch.float32)])
def test_layer_norm_vs_torch(rs, xdt, ydt):
    g = torch.Generator().manual_seed(1)
    n_in, n = 700, 333
    x = (torch.randn(n_in, 128, generator=g) * 2 + 0.5).to(xdt)
    w, b = torch.rand(128, generator=g) + 0.5, torch.randn(128, generator=g)
    index = torch.randperm(n_in, generator=g)[:n]
    cot = torch.randn(n, 128, generator=g)
    xr, wr, br = x.float().to(DEV).requires_grad_(True), w.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
    want = F.layer_norm(xr[index.to(DEV)], (128,), wr, br, 1e-5)
    (want * cot.to(DEV)).sum().backward()
    xp, wp, bp = x.to(DEV).requires_grad_(True), w.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
    got = rs.encoder.layer_norm(xp, wp, bp, 1e-5, index=index.to(DEV), out_dtype=ydt)
    assert got.dtype == ydt
    (got.float() * cot.to(DEV)).sum().backward()
    lo = xdt != torch.float32 or ydt != torch.float32
    tol = dict(rtol=2e-2, atol=3e-2) if lo else dict(rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(got.float(), want.detach(), **tol)
    torch.testing.assert_close(xp.grad.float(), xr.grad, **tol)
    torch.testing.assert_close(wp.grad, wr.grad, **(dict(rtol=2e-2, atol=0.3) if lo else dict(rtol=1e-4, atol=1e-3)))
    torch.testing.assert_close(bp.grad, br.grad, **(dict(rtol=2e-2, atol=0.3) if lo else dict(rtol=1e-4, atol=1e-3)))
    assert (xp.grad.float()[torch.ones(n_in, dtype=torch.bool).index_fill_(0, index, False)] == 0).all()


def test_elementwise_ops_vs_torch_and_dropout_law(rs):
    g = torch.Generator().manual_seed(2)
    x = torch.randn(1000, 128, generator=g).to(DEV).requires_grad_(True)
    y = torch.randn(1000, 128, generator=g).to(torch.bfloat16).to(DEV).requires_grad_(True)
    out = rs.encoder.dropout_add(x, y, 0.0)
    torch.testing.assert_close(out, x.detach() + y.detach().float())
    z = torch.randn(1000, 256, generator=g).to(DEV).requires_grad_(True)
    f = rs.encoder.gelu_dropout(z, 0.0)
    torch.testing.assert_close(f, F.gelu(z.detach()), rtol=1e-5, atol=1e-6)
    f.sum().backward()
    z2 = z.detach().clone().requires_grad_(True)
    F.gelu(z2).sum().backward()
    torch.testing.assert_close(z.grad, z2.grad, rtol=1e-4, atol=1e-5)
    # dropout: Bernoulli(1-p)/(1-p), the same mask in forward and backward
    p = 0.2
    torch.manual_seed(7)
    out = rs.encoder.dropout_add(x, y, p)
    out.sum().backward()
    kept = (out.detach() - x.detach()).abs() > 0
    frac = kept.float().mean().item()
    assert abs(frac - (1 - p)) < 0.01, frac
    torch.testing.assert_close(y.grad.float(), kept.float() / (1 - p), rtol=1e-2, atol=1e-2)
    torch.testing.assert_close((out.detach() - x.detach())[kept], (y.detach().float() / (1 - p))[kept], rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(x.grad, torch.ones_like(x))
    # two calls draw different masks; column-wise keep rates are unbiased too
    out2 = rs.encoder.dropout_add(x, y, p)
    assert not torch.equal(out2, out)
    assert (kept.float().mean(0) - (1 - p)).abs().max() < 0.06


def _tower(rs, ut):
    m = rs.SASRecUserTower(SimpleNamespace(**ut["args"]))
    m.load_state_dict(ut["state"], strict=True)
    return m.to(DEV).eval()


def _packing(pad):
    valid = ~pad
    B, Lq = valid.shape
    idx = torch.nonzero(valid.reshape(-1)).squeeze(1)
    return idx, _cu(valid.sum(1).tolist())


def test_packed_tower_vs_reference_fixture(rs):
    """the packed encoder reproduces the reference's outputs at every valid position (train-mode signature, eval-mode
    layers) and its last-step output, and the parameter gradients for a cotangent supported on the valid steps."""
    ut = load_golden("user_tower.pt")
    m = _tower(rs, ut)
    inp = {k: v.to(DEV) for k, v in ut["inputs"].items()}
    pad = ut["inputs"]["padding_mask"]
    if not bool(((~pad).sum(1) > 0).all()):
        pytest.skip("fixture has an empty sequence")
    idx, cu = _packing(pad)
    out = m(**inp, training_mode=True, packed_index=idx.to(DEV), cu_seqlens=cu.to(DEV))
    want = ut["out_train"].reshape(-1, 128)[idx]
    torch.testing.assert_close(out.detach().cpu(), want, rtol=1e-3, atol=1e-4)
    ev = m(**inp, training_mode=False, packed_index=idx.to(DEV), cu_seqlens=cu.to(DEV))
    torch.testing.assert_close(ev.detach().cpu(), ut["out_eval"], rtol=1e-3, atol=1e-4)
    # gradients: dense path vs packed path with the same (valid-only) cotangent
    cot = ut["cotangent"].reshape(-1, 128)[idx].to(DEV)
    (out * cot).sum().backward()
    got = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
    m.zero_grad(set_to_none=True)
    dense = m(**inp, training_mode=True)
    (dense.reshape(-1, 128)[idx.to(DEV)] * cot).sum().backward()
    for k, p in m.named_parameters():
        if p.grad is None:
            continue
        torch.testing.assert_close(got[k], p.grad, rtol=2e-3, atol=2e-4, msg=k)


def test_packed_step_matches_dense_step_bf16(rs):
    """full train step, eval-mode towers (no dropout), bf16 autocast: packed vs padded-grid encoder."""
    syn = rs.synthetic
    n_items, B, SL = 3000, 64, 50
    torch.manual_seed(0)
    model = rs.SASRecUserTower(syn.tower_args(num_items=n_items, max_len=SL)).to(DEV).eval()
    item = rs.SASRecItemTower(n_items, 128, syn.log_q(n_items)).to(DEV)
    lookup = syn.pretrained_table(n_items).to(DEV)
    item.init_from_pretrained(lookup)
    batch = rs.train.prepare_batch(rs.train.add_host_index(syn.make_batch(B, SL, n_items, seed=9)), DEV)
    res = {}
    for packed in (False, True):
        model.zero_grad(set_to_none=True)
        item.zero_grad(set_to_none=True)
        opt = torch.optim.SGD(list(model.parameters()) + list(item.parameters()), lr=0.0)
        t, mn, c = rs.train.two_tower_step(model, item, batch, lookup, opt, packed=packed)
        res[packed] = (mn.item(), c.item(), {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None})
    assert abs(res[True][0] - res[False][0]) < 2e-2 and abs(res[True][1] - res[False][1]) < 2e-2, (res[True][:2], res[False][:2])
    for k, gref in res[False][2].items():
        g = res[True][2][k]
        assert (g - gref).abs().max() <= 0.1 * gref.abs().max() + 1e-6, (k, (g - gref).abs().max(), gref.abs().max())


def test_graphed_step_matches_eager_and_resorts_after_load(rs):
    """train.GraphedStep: the whole step replayed from one CUDA graph gives the eager step's losses and gradients,
    also after `load` refilled the static batch with DIFFERENT ids of the same shapes (the users of the batch in
    another order: same token / distinct-item counts) -- the sparse backward's sorts must be inside the graph."""
    syn = rs.synthetic
    n_items, B, SL = 3000, 64, 50
    torch.manual_seed(0)
    model = rs.SASRecUserTower(syn.tower_args(num_items=n_items, max_len=SL)).to(DEV).eval()      # no dropout
    item = rs.SASRecItemTower(n_items, 128, syn.log_q(n_items)).to(DEV)
    lookup = syn.pretrained_table(n_items).to(DEV)
    item.init_from_pretrained(lookup)
    host = syn.make_batch(B, SL, n_items, seed=9)
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(1))
    host_b = {k: v[perm].clone() for k, v in host.items()}
    ha, hb = rs.train.add_host_index(host), rs.train.add_host_index(host_b)
    assert all(ha[k].shape == hb[k].shape for k in ha)
    params = list(model.parameters()) + list(item.parameters())
    opt = torch.optim.AdamW(params, lr=0.0, fused=True, capturable=True)
    step = lambda b: rs.train.two_tower_step(model, item, b, lookup, opt)

    def grads():
        return {k: p.grad.clone() for k, p in list(model.named_parameters()) + list(item.named_parameters())
                if p.grad is not None}

    want = {}
    for name, h in (("a", ha), ("b", hb)):
        out = step(rs.train.prepare_batch(h, DEV))
        want[name] = ([x.item() for x in out], grads())
    g = rs.train.GraphedStep(step, rs.train.prepare_batch(ha, DEV))
    for name, h in (("a", ha), ("b", hb), ("a", ha)):
        g.load(h)
        out = g.replay()
        got = ([x.item() for x in out], grads())
        for x, y in zip(got[0], want[name][0]):
            assert abs(x - y) <= 1e-4 * max(1.0, abs(y)), (name, got[0], want[name][0])
        for k, gref in want[name][1].items():
            torch.testing.assert_close(got[1][k], gref, rtol=1e-3, atol=1e-5 + 1e-3 * gref.abs().max().item(), msg=k)


def test_graphed_step_draws_new_dropout_masks_per_replay(rs):
    """seeds are by-value kernel arguments frozen at capture; the device-side epoch (rs_rng_advance, inside the step)
    makes every replay an independent draw: consecutive replays of the same batch give different losses (train mode,
    lr = 0 so that nothing else changes)."""
    syn = rs.synthetic
    n_items, B, SL = 3000, 64, 50
    torch.manual_seed(0)
    model = rs.SASRecUserTower(syn.tower_args(num_items=n_items, max_len=SL)).to(DEV).train()
    item = rs.SASRecItemTower(n_items, 128, syn.log_q(n_items)).to(DEV)
    lookup = syn.pretrained_table(n_items).to(DEV)
    item.init_from_pretrained(lookup)
    batch = rs.train.prepare_batch(rs.train.add_host_index(syn.make_batch(B, SL, n_items, seed=9)), DEV)
    opt = torch.optim.AdamW(list(model.parameters()) + list(item.parameters()), lr=0.0, fused=True, capturable=True)
    g = rs.train.GraphedStep(lambda b: rs.train.two_tower_step(model, item, b, lookup, opt), batch)
    vals = [g.replay()[1].item() for _ in range(4)]
    assert len(set(vals)) == 4, vals
    assert max(vals) - min(vals) < 0.5, vals                   # same loss up to the dropout noise


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_attn_tensor_core_path_matches_simt_path(rs, dtype):
    """16-bit operands run on the m16n8k16 tiles (attn_mma.cuh), fp32 on the SIMT kernel: same in_proj bias handling,
    same zero-tail rule and -- with a seed -- the SAME dropout mask (both hash (row id, key)), forward and backward.
    Tolerance: 16-bit rounding of the biased operands, of P / dS before the second product and of the outputs."""
    H, hd, lens = 4, 32, [5, 37, 2, 16, 17, 50, 33, 1, 1]
    g = torch.Generator().manual_seed(11)
    T = sum(lens)
    qkv = (0.7 * torch.randn(T, 3 * H * hd, generator=g)).to(dtype).to(DEV)
    bias = (0.3 * torch.randn(3 * H * hd, generator=g)).to(DEV)
    w = torch.randn(T, H * hd, generator=g).to(dtype).to(DEV)
    cu, scale = _cu(lens).to(DEV), 1 / math.sqrt(hd)
    for p, seed in ((0.0, 0), (0.25, 987654321)):
        args = (cu, H, 64, 2, scale, p, seed)
        out16, lse16 = torch.ops.rs.attn_varlen(qkv, bias, *args)
        out32, lse32 = torch.ops.rs.attn_varlen(qkv.float(), bias, *args)
        torch.testing.assert_close(out16.float(), out32, rtol=2e-2, atol=2e-2)
        torch.testing.assert_close(lse16, lse32, rtol=1e-2, atol=2e-2)
        assert (out16[-2:] == 0).all()
        d16, _ = torch.ops.rs.attn_varlen_bwd(qkv, bias, w, out16, lse16, *args)
        d32, _ = torch.ops.rs.attn_varlen_bwd(qkv.float(), bias, w.float(), out32, lse32, *args)
        assert (d16[-2:] == 0).all()
        err = (d16.float() - d32).abs().max().item()
        assert err <= 3e-2 * d32.abs().max().item() + 1e-3, (p, err, d32.abs().max().item())


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_l2_normalize_vs_torch(rs, dtype):
    """rs::l2_normalize == F.normalize(x.float(), p=2, dim=-1) forward and backward, zero rows included (clamp active:
    output 0, gradient g / eps)."""
    g = torch.Generator().manual_seed(2)
    x = torch.randn(1001, 128, generator=g).to(dtype)
    x[17] = 0
    w = torch.randn(1001, 128, generator=g)
    a = x.float().to(DEV).requires_grad_(True)
    want = F.normalize(a, p=2, dim=-1)
    (want * w.to(DEV)).sum().backward()
    b = x.to(DEV).requires_grad_(True)
    got = rs.encoder.l2_normalize(b)
    assert got.dtype == torch.float32
    (got * w.to(DEV)).sum().backward()
    torch.testing.assert_close(got, want.detach(), rtol=1e-6, atol=1e-6)
    assert (got[17] == 0).all()
    rows = torch.arange(1001) != 17
    tol = dict(rtol=1e-5, atol=1e-5) if dtype == torch.float32 else dict(rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(b.grad.float()[rows], a.grad[rows], **tol)
    assert torch.isfinite(b.grad.float()).all()


@pytest.mark.parametrize("xdt,ydt", [(torch.float32, torch.float32), (torch.bfloat16, torch.bfloat16)])
def test_layer_norm_gelu_vs_torch(rs, xdt, ydt):
    """rs::ln_act (LayerNorm -> exact GELU in one pass) against F.layer_norm + F.gelu in fp32."""
    g = torch.Generator().manual_seed(11)
    n = 517
    x = (torch.randn(n, 128, generator=g) * 2 + 0.5).to(xdt)
    w, b = torch.rand(128, generator=g) + 0.5, torch.randn(128, generator=g)
    cot = torch.randn(n, 128, generator=g).to(DEV)
    xr, wr, br = x.float().to(DEV).requires_grad_(True), w.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
    want = F.gelu(F.layer_norm(xr, (128,), wr, br, 1e-5))
    (want * cot).sum().backward()
    xp, wp, bp = x.to(DEV).requires_grad_(True), w.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
    got = rs.encoder.layer_norm_act(xp, wp, bp, 1e-5, "gelu", 0.0, ydt)
    assert got.dtype == ydt
    (got.float() * cot).sum().backward()
    lo = xdt != torch.float32
    tol = dict(rtol=2e-2, atol=3e-2) if lo else dict(rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(got.float(), want.detach(), **tol)
    torch.testing.assert_close(xp.grad.float(), xr.grad, **tol)
    torch.testing.assert_close(wp.grad, wr.grad, **(dict(rtol=2e-2, atol=0.3) if lo else dict(rtol=1e-4, atol=1e-3)))
    torch.testing.assert_close(bp.grad, br.grad, **(dict(rtol=2e-2, atol=0.3) if lo else dict(rtol=1e-4, atol=1e-3)))


def test_sequential_fuses_ln_gelu_dropout_like_the_stock_modules(rs):
    """encoder.sequential on static_mlp's layout (Linear -> LayerNorm -> GELU -> Dropout): eval mode equals the stock
    modules; in train mode the dropout keeps ~1-p of the entries, scales the kept ones by 1/(1-p), and the backward uses
    the same mask."""
    torch.manual_seed(5)
    seq = torch.nn.Sequential(torch.nn.Linear(100, 128), torch.nn.LayerNorm(128), torch.nn.GELU(), torch.nn.Dropout(0.25)).to(DEV)
    x = torch.randn(4096, 100, device=DEV)
    seq.eval()
    torch.testing.assert_close(rs.encoder.sequential(seq, x), seq(x), rtol=1e-4, atol=1e-4)
    seq.train()
    rs.encoder.rng_advance()
    xg = x.clone().requires_grad_(True)
    y = rs.encoder.sequential(seq, xg)
    seq.eval()
    full = seq(x)
    kept = y != 0
    frac = kept.float().mean().item()
    assert abs(frac - 0.75 * (full != 0).float().mean().item()) < 0.01
    torch.testing.assert_close(y[kept], (full / 0.75)[kept], rtol=1e-4, atol=1e-4)
    # the gradient flows only through kept entries: d/dy sum(y) with the mask == the stock modules' gradient of sum(mask * y / 0.75)
    y.sum().backward()
    xr = x.clone().requires_grad_(True)
    (seq(xr) * kept / 0.75).sum().backward()
    torch.testing.assert_close(xg.grad, xr.grad, rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_fused_head_matches_concat_linear_ln_gelu(rs, dtype):
    """encoder.fused_head (split Linear(256 -> 128), no concatenation, LN + GELU in one pass) against the stock modules in
    fp32 on cat([rows, prof[users]]): values and every gradient (rows, profile rows, Linear weight / bias, LN gamma /
    beta).  Tolerance: 16-bit operands (rel 2^-8) through a K = 256 contraction and a LayerNorm."""
    g = torch.Generator().manual_seed(21)
    U, n_sorted, extra = 96, 1500, 64
    lin, ln = torch.nn.Linear(256, 128).to(DEV), torch.nn.LayerNorm(128).to(DEV)
    with torch.no_grad():
        ln.weight.add_(torch.randn(128, generator=g).to(DEV) * 0.1)
        ln.bias.add_(torch.randn(128, generator=g).to(DEV) * 0.1)
    users = torch.cat([torch.sort(torch.randint(0, U, (n_sorted,), generator=g)).values,
                       torch.randperm(U, generator=g)[:extra]]).to(DEV)
    rows = torch.randn(n_sorted + extra, 128, generator=g).to(DEV)
    prof = torch.randn(U, 128, generator=g).to(DEV)
    cot = torch.randn(n_sorted + extra, 128, generator=g).to(DEV)
    rr, pr = rows.clone().requires_grad_(True), prof.clone().requires_grad_(True)
    want = F.gelu(ln(lin(torch.cat([rr, pr[users]], -1))))
    (want * cot).sum().backward()
    ref = [rr.grad, pr.grad, lin.weight.grad.clone(), lin.bias.grad.clone(), ln.weight.grad.clone(), ln.bias.grad.clone()]
    for p in list(lin.parameters()) + list(ln.parameters()):
        p.grad = None
    rp, pp = rows.to(dtype).requires_grad_(True), prof.clone().requires_grad_(True)
    got = rs.encoder.fused_head(rp, pp, users, n_sorted, lin, ln)
    assert got.dtype == dtype
    (got.float() * cot).sum().backward()
    torch.testing.assert_close(got.float(), want.detach(), rtol=3e-2, atol=3e-2)
    mine = [rp.grad.float(), pp.grad, lin.weight.grad, lin.bias.grad, ln.weight.grad, ln.bias.grad]
    for a, b in zip(mine, ref):
        rel = (a - b).norm() / b.norm()
        assert rel < 2e-2, rel


def test_split_rows_backward_is_one_concatenation(rs):
    x = torch.randn(100, 8, device=DEV, requires_grad=True)
    a, b, c = rs.ops.split_rows(x, 70, 20, 10)
    assert a.shape[0] == 70 and b.shape[0] == 20 and c.shape[0] == 10
    (a.sum() * 2 + c.sum() * 3).backward()                 # b unused: its gradient is zeros
    want = torch.cat([torch.full((70, 8), 2.0), torch.zeros(20, 8), torch.full((10, 8), 3.0)]).to(DEV)
    torch.testing.assert_close(x.grad, want)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_attn_last_only_rows_match_full_attention(rs, dtype):
    """attn_varlen(one_row_from=k): for the sequences from k on only the LAST row is computed (matrix-vector kernel);
    that row and the d_qkv it induces must equal full attention with a gradient that is zero on the other rows; the
    sequences before k are untouched; the other rows of the last-only sequences are zeros."""
    g = torch.Generator().manual_seed(17)
    lens = [5, 50, 17, 1, 33, 16, 2, 64, 9, 31]
    k, H = 4, 4
    cu = _cu(lens).to(DEV)
    T = sum(lens)
    qkv = (torch.randn(T, 3 * H * 32, generator=g) * 0.7).to(dtype).to(DEV)
    bias = (torch.randn(3 * H * 32, generator=g) * 0.1).to(DEV)
    cot = torch.randn(T, H * 32, generator=g).to(DEV)
    last = torch.tensor([sum(lens[:i + 1]) - 1 for i in range(len(lens))])
    t_k = sum(lens[:k])
    keep = torch.zeros(T, dtype=torch.bool)
    keep[:t_k] = True
    keep[last[k:]] = True
    keep = keep.to(DEV)
    cot = cot * keep.unsqueeze(1)
    a = qkv.clone().requires_grad_(True)
    full = rs.encoder.attn_varlen(a, cu, H, 64, bias=bias)
    (full.float() * cot).sum().backward()
    b = qkv.clone().requires_grad_(True)
    part = rs.encoder.attn_varlen(b, cu, H, 64, bias=bias, one_row_from=k)
    (part.float() * cot).sum().backward()
    tol = dict(rtol=1e-4, atol=1e-5) if dtype == torch.float32 else dict(rtol=3e-2, atol=3e-2)
    torch.testing.assert_close(part[keep].float(), full[keep].float(), **tol)
    assert (part[~keep] == 0).all()
    if dtype == torch.float32:
        torch.testing.assert_close(b.grad, a.grad, rtol=1e-4, atol=1e-5)
    else:
        rel = (b.grad.float() - a.grad.float()).norm() / a.grad.float().norm()
        assert rel < 2e-2, rel
    # dropout: forward and backward of the last-row kernel agree on the mask (gradient of sum(out) w.r.t. V rows is
    # the dropped probability row: it is zero exactly where the forward dropped a key)
    rs.encoder.rng_advance()
    torch.manual_seed(3)
    c = qkv.float().clone().requires_grad_(True)
    out = rs.encoder.attn_varlen(c, cu, H, 64, dropout_p=0.5, one_row_from=0)
    out[last].sum().backward()
    dv = c.grad.view(T, 3, H, 32)[:, 2]                       # [T, H, 32]: p_drop[j] * 1
    dropped = (dv.abs().sum(-1) == 0)                          # key j dropped for head h
    frac = dropped.float().mean().item()
    assert 0.3 < frac < 0.7, frac


def test_attn_one_row_in_the_middle_of_a_sequence(rs):
    """attn_varlen(one_row_from=k, one_rows=...): the single row may be ANY token of its sequence (DuoRec reads position
    len-1 of the left-padded grid), or none (a row index outside the sequence)."""
    g = torch.Generator().manual_seed(18)
    lens = [7, 40, 12, 1, 33, 20]
    k, H = 2, 4
    cu = _cu(lens).to(DEV)
    T = sum(lens)
    starts = [sum(lens[:i]) for i in range(len(lens))]
    one = torch.tensor([starts[2] + 5, starts[3] + 0, T + 3, starts[5] + 19])          # seq 4: none
    qkv = (torch.randn(T, 3 * H * 32, generator=g) * 0.7).to(DEV)
    bias = (torch.randn(3 * H * 32, generator=g) * 0.1).to(DEV)
    keep = torch.zeros(T, dtype=torch.bool)
    keep[:starts[k]] = True
    keep[one[one < T]] = True
    keep = keep.to(DEV)
    cot = torch.randn(T, H * 32, generator=g).to(DEV) * keep.unsqueeze(1)
    a = qkv.clone().requires_grad_(True)
    full = rs.encoder.attn_varlen(a, cu, H, 64, bias=bias)
    (full * cot).sum().backward()
    b = qkv.clone().requires_grad_(True)
    part = rs.encoder.attn_varlen(b, cu, H, 64, bias=bias, one_row_from=k, one_rows=one.to(DEV))
    (part * cot).sum().backward()
    torch.testing.assert_close(part[keep], full[keep], rtol=1e-4, atol=1e-5)
    assert (part[~keep] == 0).all()
    torch.testing.assert_close(b.grad, a.grad, rtol=1e-4, atol=1e-5)


def test_dropout_add_layer_norm_fused_pass(rs):
    """encoder.dropout_add_layer_norm == (x + dropout(y + bias), LayerNorm of it): against torch in fp32 with p = 0 (values
    and all five gradients, both outputs used); with p > 0 the forward and backward share one mask and the bias gradient
    is the column sum of dy."""
    g = torch.Generator().manual_seed(31)
    n = 777
    x = torch.randn(n, 128, generator=g).to(DEV)
    y = torch.randn(n, 128, generator=g).to(DEV)
    bias, w, b = (torch.randn(128, generator=g).to(DEV) * 0.3 for _ in range(3))
    w = w + 1.0
    c1, c2 = torch.randn(n, 128, generator=g).to(DEV), torch.randn(n, 128, generator=g).to(DEV)
    ref = [t.clone().requires_grad_(True) for t in (x, y, bias, w, b)]
    x1 = ref[0] + ref[1] + ref[2]
    h = F.layer_norm(x1, (128,), ref[3], ref[4], 1e-5)
    ((x1 * c1).sum() + (h * c2).sum()).backward()
    mine = [t.clone().requires_grad_(True) for t in (x, y, bias, w, b)]
    x1m, hm = rs.encoder.dropout_add_layer_norm(mine[0], mine[1], 0.0, mine[2], mine[3], mine[4], 1e-5, torch.float32)
    ((x1m * c1).sum() + (hm * c2).sum()).backward()
    torch.testing.assert_close(x1m, x1.detach(), rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(hm, h.detach(), rtol=1e-4, atol=1e-4)
    for a, r in zip(mine, ref):
        torch.testing.assert_close(a.grad, r.grad, rtol=1e-4, atol=1e-3)
    # bf16 branch + dropout: kept entries of x1 - x equal (y + bias) / keep; dy is zero exactly where the forward dropped
    p = 0.3
    rs.encoder.rng_advance()
    torch.manual_seed(5)
    xb = x.clone().requires_grad_(True)
    yb = y.to(torch.bfloat16).requires_grad_(True)
    bb = bias.clone().requires_grad_(True)
    x1d, hd = rs.encoder.dropout_add_layer_norm(xb, yb, p, bb, w, b, 1e-5, torch.bfloat16)
    assert hd.dtype == torch.bfloat16
    delta = x1d.detach() - x
    kept = delta != 0
    assert abs(kept.float().mean().item() - (1 - p)) < 0.01
    torch.testing.assert_close(delta[kept], ((yb.detach().float() + bias) / (1 - p))[kept], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(hd.float(), F.layer_norm(x1d.detach(), (128,), w, b, 1e-5), rtol=2e-2, atol=2e-2)
    (x1d * c1).sum().backward()                              # LN branch unused
    torch.testing.assert_close(yb.grad.float(), (c1 * kept / (1 - p)).to(torch.bfloat16).float(), rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(bb.grad, (c1 * kept / (1 - p)).to(torch.bfloat16).float().sum(0), rtol=1e-2, atol=2e-1)
    torch.testing.assert_close(xb.grad, c1)
    xb.grad = yb.grad = bb.grad = None
    rs.encoder.rng_advance()
    torch.manual_seed(6)
    x1d, hd = rs.encoder.dropout_add_layer_norm(xb, yb, p, bb, w, b, 1e-5, torch.bfloat16)
    kept = (x1d.detach() - x) != 0
    ((x1d * c1).sum() + (hd.float() * c2).sum()).backward()   # both branches
    assert ((yb.grad.float() == 0) | kept).all() and (yb.grad.float()[kept] != 0).float().mean() > 0.99
    torch.testing.assert_close(yb.grad.float()[kept], (xb.grad / (1 - p))[kept], rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(bb.grad, yb.grad.float().sum(0), rtol=1e-2, atol=3e-1)


def test_gelu_dropout_mask_law_and_fwd_bwd_agreement(rs):
    """gelu_dropout with p > 0 (two 32-bit hashes -> four 16-bit keep decisions): keep rate, independence of the four
    decisions of one hash pair, per-column rates, and the backward uses the forward's mask."""
    p = 0.25
    rs.encoder.rng_advance()
    torch.manual_seed(11)
    z = (torch.randn(4096, 256, device=DEV) + 3.0).bfloat16().requires_grad_(True)      # gelu(z) != 0 almost surely
    f = rs.encoder.gelu_dropout(z, p)
    kept = f != 0
    assert abs(kept.float().mean().item() - (1 - p)) < 0.005
    assert (kept.float().mean(0) - (1 - p)).abs().max() < 0.04
    k4 = kept.view(-1, 4).float()
    for a in range(4):
        for b in range(a + 1, 4):                               # pairwise joint keep rate == (1-p)^2
            assert abs((k4[:, a] * k4[:, b]).mean().item() - (1 - p) ** 2) < 0.01, (a, b)
    ref = torch.nn.functional.gelu(z.detach().float()).bfloat16().float() / (1 - p)
    torch.testing.assert_close(f.float()[kept], ref[kept], rtol=2e-2, atol=2e-2)
    f.float().sum().backward()
    assert ((z.grad != 0) == kept).float().mean().item() > 0.999
    f2 = rs.encoder.gelu_dropout(z, p)
    assert not torch.equal(f2 != 0, kept)


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("p_drop", [0.0, 0.25])
def test_attn_one_tile_backward_matches_two_phase_backward(rs, p_drop, dt):
    """16-bit operands without an in_proj bias: sequences of <= 16 tokens take the one-tile backward kernel (S and dP
    computed once, P^T / dS^T by movmatrix).  With an all-zero bias the same call takes the two-phase kernel for every
    sequence: same seed -> same dropout mask -> the two d_qkv must agree to 16-bit rounding, and both must match the
    fp32 SIMT kernels."""
    g = torch.Generator().manual_seed(23)
    lens = [1, 2, 3, 5, 8, 9, 13, 15, 16, 17, 30, 16, 4, 50, 7, 11]
    H = 4
    cu = _cu(lens).to(DEV)
    T = sum(lens)
    qkv = (torch.randn(T, 3 * H * 32, generator=g) * 0.7).to(dt).to(DEV)
    w = torch.randn(T, H * 32, generator=g).to(dt).to(DEV)
    zero_bias = torch.zeros(3 * H * 32, device=DEV)
    scale, seed = 1 / math.sqrt(32), 987654321
    res = []
    for bias in (None, zero_bias):
        out, lse = torch.ops.rs.attn_varlen(qkv, bias, cu, H, 64, 0, scale, p_drop, seed)
        dq, _ = torch.ops.rs.attn_varlen_bwd(qkv, bias, w, out, lse, cu, H, 64, 0, scale, p_drop, seed)
        res.append((out, dq))
    assert torch.equal(res[0][0], res[1][0])
    a, b = res[0][1].float(), res[1][1].float()
    assert (a - b).norm() / b.norm() < 5e-3
    torch.testing.assert_close(a, b, rtol=3e-2, atol=3e-2)
    if p_drop == 0.0:                                       # (the fp32 kernels draw the same mask only per dtype path)
        out32, lse32 = torch.ops.rs.attn_varlen(qkv.float(), None, cu, H, 64, 0, scale, 0.0, 0)
        d32, _ = torch.ops.rs.attn_varlen_bwd(qkv.float(), None, w.float(), out32, lse32, cu, H, 64, 0, scale, 0.0, 0)
        assert (a - d32).norm() / d32.norm() < 1e-2


def test_emb_layer_norm2_matches_the_two_separate_passes(rs):
    """encoder.emb_layer_norm2 (embedding LayerNorm + dropout + first-layer LayerNorm in one pass; backward folds the two
    dropout views before the embedding LayerNorm's backward) against layer_norm(index, index_inv) + residual_layer_norm:
    p = 0 -> values and all gradients; p > 0 -> same seed, same hash -> identical masks, so everything still matches."""
    g = torch.Generator().manual_seed(41)
    n_src = 600
    e = (torch.randn(n_src, 128, generator=g) * 1.5 + 0.3).bfloat16().to(DEV)
    perm = torch.randperm(2 * n_src, generator=g)
    inv1, inv2 = perm[:n_src].to(DEV), perm[n_src:].to(DEV)            # packed slots of every source row
    index = torch.empty(2 * n_src, dtype=torch.int64, device=DEV)
    index[inv1] = torch.arange(n_src, device=DEV)
    index[inv2] = torch.arange(n_src, device=DEV)
    ln0, ln1 = torch.nn.LayerNorm(128).to(DEV), torch.nn.LayerNorm(128).to(DEV)
    with torch.no_grad():
        for ln in (ln0, ln1):
            ln.weight.add_(torch.randn(128, generator=g).to(DEV) * 0.2)
            ln.bias.add_(torch.randn(128, generator=g).to(DEV) * 0.2)
    c0 = torch.randn(2 * n_src, 128, generator=g).to(DEV)
    c1 = torch.randn(2 * n_src, 128, generator=g).to(DEV)
    for p in (0.0, 0.3):
        res = []
        for fused in (False, True):
            for ln in (ln0, ln1):
                ln.weight.grad = ln.bias.grad = None
            ee = e.clone().requires_grad_(True)
            torch.manual_seed(77)                                       # same host seed -> same dropout stream
            if fused:
                x0, h = rs.encoder.emb_layer_norm2(ee, index, (inv1, inv2), ln0, p, ln1, torch.bfloat16)
            else:
                x0 = rs.encoder.layer_norm(ee, ln0.weight, ln0.bias, ln0.eps, index=index, dropout_p=p,
                                           out_dtype=torch.float32, index_inv=(inv1, inv2))
                x0, h = rs.encoder.residual_layer_norm(x0, ln1.weight, ln1.bias, ln1.eps, torch.bfloat16)
            ((x0 * c0).sum() + (h.float() * c1).sum()).backward()
            res.append([x0.detach(), h.detach().float(), ee.grad.float(), ln0.weight.grad.clone(), ln0.bias.grad.clone(),
                        ln1.weight.grad.clone(), ln1.bias.grad.clone()])
        for a, b in zip(res[1], res[0]):
            rel = (a - b).norm() / b.norm()
            assert rel < 5e-3, (p, rel)
        torch.testing.assert_close(res[1][0], res[0][0], rtol=1e-5, atol=1e-5)
