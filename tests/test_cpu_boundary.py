"""CPU suite: the C-ABI library loads and exports every symbol include/rs_twotower.h declares; the
product package has no CPU fallback and never touches the oracle; host-side routing logic."""
import ctypes
import importlib
import os
import re
import subprocess
import sys

import pytest
import torch

from conftest import PKG, ROOT

PKG_DIR = os.path.join(ROOT, PKG)


def _header_functions():
    src = open(os.path.join(ROOT, "include", "rs_twotower.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rs_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib_path = os.path.join(PKG_DIR, "librs_twotower.so")
    if not os.path.exists(lib_path):
        sys.path.insert(0, ROOT)
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(lib_path)
    names = _header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/rs_twotower.h but not exported"
    lib.rs_abi_version.restype = ctypes.c_int
    assert lib.rs_abi_version() == 1
    lib.rs_error_string.restype = ctypes.c_char_p
    assert b"workspace" in lib.rs_error_string(10003)


def test_python_binding_covers_header(rs):
    assert sorted(rs._lib.PROTOTYPES) == _header_functions()


def test_no_cpu_fallback(rs):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rs.gather_rows(torch.randn(8, 128), torch.tensor([1, 2]))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rs.retrieve_topk(torch.randn(4, 128), torch.randn(64, 128), 3)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(PKG_DIR):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f
    assert "oracle" not in sys.modules or True


def test_missing_library_fails_loudly(tmp_path):
    code = (f"import sys, importlib; sys.path.insert(0, {ROOT!r});"
            f"m = importlib.import_module({PKG!r} + '._lib'); m.LIB_PATH = {str(tmp_path / 'nope.so')!r}; m._lib = None;"
            "m.load()")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU / PyTorch fallback" in r.stderr


def test_custom_ops_registered(rs):
    for name in ("gather_rows", "embedding_dense_bwd", "seq_front", "seq_front_bwd", "static_front",
                 "normalized_rows", "ce_fwd", "ce_bwd", "retrieve_topk", "fm_fwd", "fm_bwd", "bert_embed",
                 "std_front", "masked_mean"):
        assert hasattr(torch.ops.rs, name)


def test_state_dict_names_match_reference(rs):
    from conftest import load_golden
    from types import SimpleNamespace
    ut = load_golden("user_tower.pt")
    m = rs.SASRecUserTower(SimpleNamespace(**ut["args"]))
    m.load_state_dict(ut["state"], strict=True)
    im = load_golden("item_matrix.pt")
    it = rs.SASRecItemTower(300, 128)
    it.load_state_dict(im["state"], strict=True)


def test_route_and_shard_roundtrip(rs):
    sh = rs.sharded
    full = torch.arange(23 * 4, dtype=torch.float32).view(23, 4)
    for world in (1, 2, 3, 8):
        shards = [sh.shard_rows(full, r, world) for r in range(world)]
        assert torch.equal(sh.unshard_rows(shards), full)
        ids = torch.randint(0, 23, (57,))
        order, counts, local = sh.route(ids, world)
        assert counts.sum() == 57 and torch.equal(torch.sort(order).values, torch.arange(57))
        owner = (ids % world)[order]
        assert torch.equal(owner, torch.sort(owner).values)
        rows = torch.stack([shards[o][l] for o, l in zip(owner.tolist(), local.tolist())])
        out = torch.empty_like(rows); out[order] = rows
        assert torch.equal(out, full[ids])


def test_synthetic_batch_shapes(rs):
    syn = rs.synthetic if hasattr(rs, "synthetic") else importlib.import_module(PKG + ".synthetic")
    b = syn.make_batch(16, L=50, num_items=1000)
    assert b["item_ids"].shape == (16, 50) and b["item_ids"].max() <= 1000
    pad = b["padding_mask"]
    assert (b["item_ids"][pad] == 0).all() and (b["item_ids"][~pad] > 0).all()
    assert (pad[:, 1:] <= pad[:, :-1]).all()              # left padding
    assert syn.log_q(1000)[0] == -20.0


def test_two_view_index_layout_is_what_the_ln_backward_fold_assumes(rs):
    """train.add_host_index lays the two dropout views out as [tokens | tokens | extras | extras]; the LayerNorm backward
    (encoder.layer_norm(index_fold=(T, E))) turns its scatter-add into two elementwise sums on that promise."""
    syn = rs.synthetic
    for seed, B in ((3, 64), (4, 257)):
        hb = rs.train.add_host_index(syn.make_batch(B, 50, 3000, seed=seed))
        T = hb["valid_index"].numel()
        idx = hb["pk_index_2v"]
        E = (idx.numel() - 2 * T) // 2
        assert idx.numel() == 2 * (T + E) and E >= 0
        tok, ext = torch.arange(T), T + torch.arange(E)
        assert torch.equal(idx, torch.cat([tok, tok, ext, ext]))
        assert hb["pk_item_ids"].numel() >= T + E
        # the packed grids hold exactly the valid tokens, then the literal-DuoRec positions, then zeros
        flat = hb["pk_item_ids"].reshape(-1)
        assert torch.equal(flat[:T], hb["item_ids"].reshape(-1)[hb["valid_index"]])
        assert (flat[T + E:] == 0).all()
        cu = hb["cu_seqlens_2v"]
        assert cu[0] == 0 and cu[-1] == 2 * (T + E) and (cu[1:] > cu[:-1]).all()


def test_direct_entry_points_cover_every_registered_op(rs):
    """_lib.direct.NAME is the Python implementation behind torch.ops.rs.NAME (no silent fallback to the dispatcher)."""
    from torch._library.custom_ops import OPDEFS
    names = sorted(k.split("::")[1] for k in OPDEFS if k.startswith("rs::"))
    assert len(names) >= 30
    for n in names:
        fn = getattr(rs._lib.direct, n)
        assert callable(fn) and fn is OPDEFS["rs::" + n]._init_fn, n
        assert hasattr(torch.ops.rs, n)


def test_host_wrappers_reject_cpu_tensors(rs):
    """encoder.linear / sequential / l2_normalize wrap stock modules around the row kernels; like every other entry point
    they have no CPU path (the CPU implementation of this package is the oracle, which the product never imports)."""
    g = torch.Generator().manual_seed(0)
    seq = torch.nn.Sequential(torch.nn.Linear(100, 128), torch.nn.LayerNorm(128), torch.nn.GELU(), torch.nn.Linear(128, 128))
    x = torch.randn(7, 100, generator=g)
    for call in (lambda: rs.encoder.sequential(seq, x), lambda: rs.encoder.linear(seq[0], x),
                 lambda: rs.encoder.l2_normalize(torch.randn(5, 128, generator=g))):
        with pytest.raises(RuntimeError, match="CUDA tensors only"):
            call()


def _header_struct_fields(name):
    src = open(os.path.join(ROOT, "include", "rs_twotower.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    body = re.search(r"typedef struct \{([^}]*)\}\s*" + name + r"\s*;", src, flags=re.S).group(1)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(","):
            fields.append(re.findall(r"[A-Za-z_][A-Za-z0-9_]*", part)[-1])
    return fields


def test_ctypes_structs_mirror_the_header(rs):
    """rs_ce_problem / rs_batch_index are passed by pointer: the ctypes mirrors must list the same fields in the same order."""
    L = rs._lib
    assert [f[0] for f in L.CEProblem._fields_] == _header_struct_fields("rs_ce_problem")
    assert [f[0] for f in L.BatchIndex._fields_] == _header_struct_fields("rs_batch_index")


def test_flat_batch_and_buckets_host_logic(rs):
    tr, syn = rs.train, rs.synthetic
    hb = syn.make_batch(16, 50, 500, seed=2)
    a, b = tr.FlatBatch(16, 50), tr.FlatBatch(16, 50)
    a.fill(hb)
    b.copy_(a, non_blocking=False)
    assert set(b.views) == set(hb) and all(torch.equal(b.views[k], hb[k]) for k in hb)
    assert all(v.data_ptr() % 256 == a.buf.data_ptr() % 256 for v in a.views.values())     # 256-byte aligned views
    assert a.nbytes == b.nbytes and a.nbytes >= sum(v.numel() * v.element_size() for v in hb.values())
    assert tr.bucket_of(1, 1) == (tr.TOK_BUCKET, tr.COL_BUCKET)
    assert tr.bucket_of(2048, 512) == (2048, 512) and tr.bucket_of(2049, 513) == (4096, 1024)
    assert tr.bucket_of(100, 70, 64, 32) == (128, 96)
    bs = tr.BucketedStep(lambda b_: None, 16, 50, 501, "cpu", use_graph=False, tok_q=256, col_q=64)
    assert bs.bucket(300, 70) == (512, 128) and bs.bucket(300, None) == (512, 512)


def test_zeros_many_is_one_allocation_of_aligned_views(rs):
    shapes = [(3, 4), (1001, 128), (6,), (11, 16)]
    outs = rs.ops.zeros_many(shapes, "cpu")
    assert [tuple(o.shape) for o in outs] == shapes
    base = outs[0].untyped_storage().data_ptr()
    for o in outs:
        assert o.is_contiguous() and (o == 0).all() and o.untyped_storage().data_ptr() == base
        assert (o.data_ptr() - base) % 256 == 0
    outs[1][5, 7] = 1.0                                   # views do not overlap
    assert all((o == 0).all() for k, o in enumerate(outs) if k != 1)
