"""ctypes binding of librs_twotower.so -- the C ABI declared in include/rs_twotower.h.

There is NO fallback: if the library is missing or a symbol cannot be bound,
importing the product package's ops raises.  (The CPU oracle under oracle/ is
test infrastructure and is never imported from here.)
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RS_TWOTOWER_LIB") or os.path.join(_HERE, "librs_twotower.so")   # override: kernel-variant experiments only

RS_F32, RS_F16, RS_BF16 = 0, 1, 2
RS_CE_DIAG_MASK, RS_CE_DIAG_RAW, RS_CE_SUPCON, RS_CE_NO_DIAG = 1, 2, 4, 8
RS_MAX_TABLES = 8

_DT = {torch.float32: RS_F32, torch.float16: RS_F16, torch.bfloat16: RS_BF16}
_DT_INV = {v: k for k, v in _DT.items()}

vp, i64, i32, f32, u64, sz = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_uint64, C.c_size_t


# output arrays of rs_batch_index, in struct order
BATCH_INDEX_OUTPUTS = ("pk_item_ids", "pk_time_ids", "pk_pos_ids", "pk_index_2v", "fold_inv1", "fold_inv2", "cu_seqlens_2v",
                       "row_cu", "select_2v", "users_2v", "main_tgt", "last_tgt", "row_weight", "col_item_ids", "col_counts",
                       "pos_col", "meta")


class CEProblem(C.Structure):
    """Mirror of rs_ce_problem."""
    _fields_ = [("a", vp), ("b", vp), ("ab_dtype", i32), ("M", i64), ("N", i64), ("K", i64), ("scale", f32),
                ("col_bias", vp), ("key_a_row", vp), ("key_a_col", vp), ("key_b_row", vp), ("key_b_col", vp),
                ("diag_offset", i64), ("mask_value", f32), ("flags", i32), ("logit_bound", f32)]


class BatchIndex(C.Structure):
    """Mirror of rs_batch_index."""
    _fields_ = ([("padding_mask", vp), ("item_ids", vp), ("time_ids", vp), ("target_ids", vp), ("B", i64), ("L", i64),
                 ("n_item_rows", i64), ("tok_cap", i64), ("col_cap", i64), ("grid_cap", i64)] +
                [(n, vp) for n in BATCH_INDEX_OUTPUTS])


# name -> (restype, argtypes); must list every function of include/rs_twotower.h
PROTOTYPES = {
    "rs_abi_version": (i32, []),
    "rs_error_string": (C.c_char_p, [i32]),
    "rs_launch_count": (C.c_ulonglong, []),
    "rs_count_launches": (None, [i32]),
    "rs_gather_rows": (i32, [vp, i32, i64, i64, vp, i64, i64, vp, i32, vp, vp]),
    "rs_scatter_add_rows": (i32, [vp, i32, vp, i64, i64, i64, i64, i64, f32, vp, vp, vp]),
    "rs_sort_ids_workspace_bytes": (sz, [i64]),
    "rs_sort_ids": (i32, [vp, i64, i64, i64, vp, vp, vp, sz, vp, vp]),
    "rs_segment_reduce_workspace_bytes": (sz, [i64, i64]),
    "rs_segment_reduce_rows": (i32, [vp, i32, vp, vp, i64, i64, i64, i64, vp, vp, vp, vp, vp, sz, vp]),
    "rs_seq_front_fwd": (i32, [vp, i32, vp, vp, vp, i32, vp, vp, i64, i64, i64, vp, i32, vp, vp]),
    "rs_seq_front_bwd_workspace_bytes": (sz, [i64, i64, i64, i32, vp, vp]),
    "rs_seq_front_bwd": (i32, [vp, i32, vp, vp, vp, vp, i32, vp, i64, i64, i64, i64, vp, vp, vp, vp, sz, vp]),
    "rs_static_front_fwd": (i32, [vp, vp, vp, vp, vp, vp, vp, i64, vp, vp, vp]),
    "rs_static_front_bwd_workspace_bytes": (sz, [i64]),
    "rs_static_front_bwd": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, i64, i64, vp, vp, vp, vp, vp, sz, vp]),
    "rs_normalized_rows_fwd": (i32, [vp, i64, i64, vp, i64, f32, vp, i32, vp, vp, vp]),
    "rs_normalized_rows_bwd": (i32, [vp, i32, vp, i64, i64, vp, i64, f32, vp, vp]),
    "rs_std_front_fwd": (i32, [vp, i64, i64, vp, i64, vp, i64, vp, vp, f32, vp, i32, vp, vp, vp, vp]),
    "rs_bert_embed_fwd": (i32, [vp, i64, vp, vp, vp, vp, f32, vp, i64, i64, i64, f32, u64, vp, i32, vp, vp]),
    "rs_masked_mean_fwd": (i32, [vp, i32, vp, i64, i64, i64, vp, vp]),
    "rs_masked_mean_bwd": (i32, [vp, vp, i64, i64, i64, vp, i32, vp]),
    "rs_attn_varlen_fwd": (i32, [vp, i32, vp, vp, i64, i64, i32, i32, i32, i64, i64, vp, f32, f32, u64, vp, vp, vp]),
    "rs_attn_varlen_bwd": (i32, [vp, vp, vp, i32, vp, vp, vp, i64, i64, i32, i32, i32, i64, i64, vp, f32, f32, u64, vp, vp]),
    "rs_rng_advance": (i32, [vp]),
    "rs_l2_normalize_fwd": (i32, [vp, i32, i64, i64, f32, vp, i32, vp, vp]),
    "rs_l2_normalize_bwd": (i32, [vp, i32, vp, i32, vp, i64, i64, vp, i32, vp]),
    "rs_colsum_workspace_bytes": (sz, [i64, i64]),
    "rs_colsum": (i32, [vp, i32, i64, i64, vp, vp, sz, vp]),
    "rs_ln_fwd": (i32, [vp, i32, vp, i64, i64, vp, vp, f32, f32, u64, vp, i32, vp, vp, vp]),
    "rs_ln_act_fwd": (i32, [vp, i32, vp, vp, vp, i64, i64, vp, vp, f32, i32, f32, u64, vp, i32, vp, vp, vp]),
    "rs_ln_act_bwd": (i32, [vp, i32, vp, i32, vp, vp, vp, i64, i64, vp, vp, vp, vp, i32, f32, u64, vp, vp, vp, vp, sz, vp]),
    "rs_ln_bwd_workspace_bytes": (sz, [i64]),
    "rs_ln_bwd": (i32, [vp, i32, vp, i32, vp, i64, i64, vp, vp, vp, f32, u64, vp, vp, vp, vp, vp, sz, vp]),
    "rs_dropout_add_fwd": (i32, [vp, i32, vp, i32, vp, i64, i64, f32, u64, vp, vp]),
    "rs_dropout_bwd": (i32, [vp, i32, i64, f32, u64, vp, i32, vp]),
    "rs_gelu_dropout_fwd": (i32, [vp, i32, vp, i64, i64, f32, u64, vp, vp]),
    "rs_gelu_dropout_bwd": (i32, [vp, vp, i32, vp, i64, i64, f32, u64, vp, vp]),
    "rs_emb_ln2_fwd": (i32, [vp, i32, vp, i64, i64, vp, vp, f32, f32, u64, vp, vp, f32, vp, vp, i32, vp, vp, vp, vp, vp]),
    "rs_emb_ln2_bwd_workspace_bytes": (sz, [i64]),
    "rs_emb_ln2_bwd": (i32, [vp, i32, vp, i32, vp, vp, vp, vp, i64, i64, vp, vp, vp, vp, vp, vp, f32, u64, vp, vp, vp, vp, vp,
                             vp, sz, vp]),
    "rs_dropout_add_ln_fwd": (i32, [vp, vp, i32, vp, i64, i64, f32, u64, vp, vp, f32, vp, vp, i32, vp, vp, vp]),
    "rs_ln_bwd_dropout_workspace_bytes": (sz, [i64]),
    "rs_ln_bwd_dropout": (i32, [vp, i32, vp, vp, i64, i64, vp, vp, vp, f32, u64, vp, vp, i32, vp, vp, vp, vp, sz, vp]),
    "rs_ew_colsum_workspace_bytes": (sz, [i64, i64]),
    "rs_dropout_bwd_bias": (i32, [vp, i32, i64, i64, f32, u64, vp, i32, vp, vp, sz, vp]),
    "rs_gelu_dropout_bwd_bias": (i32, [vp, vp, i32, vp, i64, i64, f32, u64, vp, vp, vp, sz, vp]),
    "rs_ce_workspace_bytes": (sz, [C.POINTER(CEProblem)]),
    "rs_ce_fwd": (i32, [C.POINTER(CEProblem), vp, vp, vp, vp, vp, sz, vp]),
    "rs_ce_bwd": (i32, [C.POINTER(CEProblem), vp, vp, vp, vp, vp, vp, vp, sz, vp]),
    "rs_ce_fwd_grad_bytes": (sz, [C.POINTER(CEProblem)]),
    "rs_ce_row_combine_workspace_bytes": (sz, [i64]),
    "rs_ce_row_combine": (i32, [vp, vp, vp, vp, i64, vp, vp, vp, vp, vp, sz, vp]),
    "rs_ce_fwd_grad": (i32, [C.POINTER(CEProblem), vp, vp, vp, vp, vp, sz, vp]),
    "rs_ce_bwd_from_grad": (i32, [C.POINTER(CEProblem), vp, vp, vp, vp, vp, vp, vp, vp, sz, vp]),
    "rs_batch_index_workspace_bytes": (sz, [i64, i64, i64]),
    "rs_batch_index_build": (i32, [C.POINTER(BatchIndex), vp, sz, vp]),
    "rs_batch_index_counts": (i32, [vp, vp, i64, i64, i64, vp, vp, sz, vp]),
    "rs_gather_add2": (i32, [vp, i32, vp, vp, i64, i64, i64, vp, vp]),
    "rs_id_histogram": (i32, [vp, i64, vp, i64, i32, vp, vp, vp]),
    "rs_owner_compact_workspace_bytes": (sz, [i32, i64]),
    "rs_owner_compact": (i32, [vp, i32, i64, i64, i64, vp, vp, vp, vp, vp, vp, sz, vp]),
    "rs_lookup_i32": (i32, [vp, i64, vp, i64, i64, vp, vp]),
    "rs_ensemble_merge": (i32, [vp, vp, vp, i64, i64, i32, f32, vp, i32, i64, vp, vp, vp, vp, vp]),
    "rs_topk_workspace_bytes": (sz, [i64, i64, i64, i64]),
    "rs_retrieve_topk": (i32, [vp, i64, vp, i64, i64, i64, i32, vp, vp, vp, sz, vp]),
    "rs_retrieve_topk_tc_workspace_bytes": (sz, [i64, i64, i64, i64]),
    "rs_retrieve_topk_tc": (i32, [vp, i64, vp, i64, i64, i64, i32, vp, vp, vp, sz, vp]),
    "rs_mine_workspace_bytes": (sz, [i64, i64, i64]),
    "rs_mine_hard_negatives": (i32, [vp, vp, vp, i64, i64, i64, f32, vp, vp, vp, vp, sz, vp]),
    "rs_sparse_logits_fwd": (i32, [vp, vp, i32, vp, i64, i64, i64, i64, f32, vp, vp, vp, vp, vp]),
    "rs_sparse_logits_bwd": (i32, [vp, vp, i32, vp, i64, i64, i64, i64, f32, vp, vp, vp, vp, vp, vp]),
    "rs_user_block_logits_fwd": (i32, [vp, vp, i32, vp, vp, i64, i64, i64, i32, f32, vp, vp, vp, vp]),
    "rs_user_block_logits_bwd": (i32, [vp, vp, i32, vp, vp, i64, i64, i64, i32, f32, vp, vp, vp, vp, vp, vp, vp]),
    "rs_fm_fwd": (i32, [vp, vp, i64, i64, vp, i64, vp, i64, vp, vp, i32, vp, vp]),
    "rs_fm_bwd": (i32, [vp, vp, i64, i64, vp, i64, vp, vp, i32, i64, vp, vp, vp]),
}

_lib = None
PROFILE = None          # set to a list to record (name, start_event, end_event) around every C call


def _algorithmic_work(name, args):
    """Algorithmic flops (fused softmax) or bytes (segment reduce) of one call, for the roofline report."""
    if name == "rs_ce_fwd":
        p = args[0]
        return 2.0 * p.M * p.N * p.K
    if name == "rs_ce_bwd":                     # per requested side: S recompute + dS@X
        p = args[0]
        sides = (args[5] is not None) + (args[6] is not None)
        return 2.0 * p.M * p.N * p.K * 2 * sides
    if name == "rs_ce_fwd_grad":                # S + P@B
        p = args[0]
        return 2.0 * p.M * p.N * p.K * 2
    if name == "rs_ce_bwd_from_grad":           # the dB side only (S recompute + dS^T@A); dA is a row scaling of G
        p = args[0]
        return 2.0 * p.M * p.N * p.K * 2 * (args[7] is not None)
    if name == "rs_segment_reduce_rows":        # one read of the gradient rows + the sorted (id, pos) pairs
        n, dim, esz = args[4], args[5], (4 if args[1] == RS_F32 else 2)
        return float(n) * (dim * esz + 8)
    return 0.0


class _Proxy:
    """Attribute access like the ctypes library; optionally brackets each call with CUDA events recorded
    on the current stream (the stream the kernels are launched on) for bench.py's roofline numbers."""

    def __init__(self, lib):
        self._lib = lib

    def __getattr__(self, name):
        raw = getattr(self._lib, name)

        def call(*args):
            if PROFILE is None:
                return raw(*args)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            rc = raw(*args)
            e.record()
            PROFILE.append((name, s, e, _algorithmic_work(name, args)))
            return rc

        call.__name__ = name
        setattr(self, name, call)
        return call


class _DirectOps:
    """`direct.NAME` = the Python implementation behind `torch.ops.rs.NAME`, called WITHOUT the dispatcher.  The ops are
    registered with torch.library (schema, fake kernels, `torch.ops.rs.*` for users and tests); the autograd Functions
    of this package, which already run with autograd off and real CUDA tensors, call the implementation directly: a
    dispatcher round trip costs ~20 us of host time and a train step makes ~140 of them (eager N > 1 steps are
    host-bound).  Falls back to the registered op if torch's registry layout changes."""

    def __getattr__(self, name):
        try:
            from torch._library.custom_ops import OPDEFS
            fn = OPDEFS["rs::" + name]._init_fn
        except Exception:
            fn = getattr(torch.ops.rs, name)
        setattr(self, name, fn)
        return fn


direct = _DirectOps()


def load():
    """Load the shared library (once) and bind every prototype.  Raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python {os.path.join(_HERE, 'build.py')}` "
            "(or __graft_entry__.build()).  This package has no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the .so is stale
        fn.restype, fn.argtypes = res, args
    _lib = _Proxy(lib)
    return _lib


def check(code: int, what: str):
    if code != 0:
        msg = load().rs_error_string(code).decode()
        raise RuntimeError(f"{what} failed: {msg} (code {code})")


def dt(t) -> int:
    d = t if isinstance(t, torch.dtype) else t.dtype
    try:
        return _DT[d]
    except KeyError:
        raise TypeError(f"unsupported dtype {d}; expected float32/float16/bfloat16") from None


def torch_dtype(code: int) -> torch.dtype:
    return _DT_INV[code]


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def ptr_array(ts):
    return (C.c_void_p * len(ts))(*[t.data_ptr() if t is not None else 0 for t in ts])


def i64_array(xs):
    return (C.c_int64 * len(xs))(*xs)


def i32_array(xs):
    return (C.c_int * len(xs))(*xs)


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("rs_twotower ops run on CUDA tensors only (sm_100a); there is no CPU fallback")


_oob = {}


def oob_flag(device):
    """Per-device int32 word set by kernels that meet an out-of-range id."""
    key = torch.device(device).index or 0
    f = _oob.get(key)
    if f is None:
        f = torch.zeros(1, dtype=torch.int32, device=device)
        _oob[key] = f
    return f


def check_ids(device=None):
    """Raise IndexError if any kernel since the last call saw an id outside its table
    (the reference raises IndexError on CPU / a device-side assert on CUDA).  Synchronises."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    f = oob_flag(dev)
    if int(f.item()) != 0:
        f.zero_()
        raise IndexError("index out of range in an rs_twotower embedding lookup")


def workspace(nbytes: int, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
