"""Kernel-level timings of the SURVEY 8a rows that are NOT inside bench.py's train step, at BASELINE.json's sizes,
each against its roofline (MEASURED_PEAKS.json): F1 (DeepFM config 3), R1 (retrieval config 5, a user sample),
H1 (1.37 M-row table gather + its sorted scatter-add backward), I2 (768-wide BERT row gather).  CUDA events, L2
flushed between iterations, median of n.  Prints one JSON line per row."""
import importlib, json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rs = importlib.import_module("llm-driven_content-based-feature_recommendation_system_b200")
syn = rs.synthetic
dev = "cuda"
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
HBM = peaks["hbm_gbs"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=9):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def report(row, what, ms, work, unit, peak, **extra):
    ach = work / (ms * 1e-3) / 1e9
    print(json.dumps(dict(row=row, what=what, ms=round(ms, 4), achieved=round(ach, 1), unit=unit, peak=peak,
                          frac=round(ach / peak, 3), **extra)), flush=True)


g = torch.Generator().manual_seed(0)
# ---- F1: DeepFM config 3, B = 65536, F = 39, k = 16
vocab = syn.criteo_vocab_sizes()
B, F_, k = 65536, len(vocab), 16
ids = syn.make_fm_batch(B, vocab, seed=1).to(dev)
fm = rs.FM(vocab, k=k, init_std=0.01).to(dev)
fwd_bytes = B * F_ * (k * 4 + 4 + 8) + B * 4 + B * F_ * k * 2           # rows + linear + ids, fm out, bf16 concat out
ms = timeit(lambda: torch.ops.rs.fm_fwd(ids, fm.offsets, fm.embedding, fm.linear.reshape(-1), True, 2))
report("F1", f"fm_fwd B={B} F={F_} k={k} (gather + FM + concat for the DNN)", ms, fwd_bytes, "GB/s", HBM)
d_fm, d_cat = torch.randn(B, device=dev), torch.randn(B, F_ * k, device=dev).bfloat16()
bwd_bytes = B * F_ * (k * 4 + 8) + B * 4 + B * F_ * k * 2 + 2 * B * F_ * (k + 1) * 4   # re-read rows/ids/dY + RMW of touched rows
ms = timeit(lambda: torch.ops.rs.fm_bwd(ids, fm.offsets, fm.embedding, d_fm, d_cat, True))
report("F1", "fm_bwd (atomic scatter into the concatenated table, incl. its memset)", ms, bwd_bytes, "GB/s", HBM)

# ---- R1: retrieval config 5: 105 542 items, top-12, a 65 536-user sample of the 1.37 M users
ni, nu, kk = syn.N_ITEMS, 65536, 12
I = torch.nn.functional.normalize(torch.randn(ni, 128, generator=g), dim=1).to(dev)
U = torch.nn.functional.normalize(torch.randn(nu, 128, generator=g), dim=1).to(dev)
ms = timeit(lambda: rs.retrieve_topk(U, I, kk), n=3)
flops = 2.0 * nu * ni * 128
print(json.dumps(dict(row="R1", what=f"retrieve_topk {nu} users x {ni} items, k={kk}, fp32 FMA (exact ids)", ms=round(ms, 2),
                      achieved=round(flops / ms / 1e9, 1), unit="TFLOP/s fp32", users_per_s=round(nu / ms * 1e3),
                      full_1p37M_users_s=round(1371980 / (nu / ms * 1e3), 2))), flush=True)

# ---- H1: 1.37 M x 64 user table (mined_inference.py:670), B = 65536 gathers + dense backward
tab = torch.randn(syn.N_CUSTOMERS + 1, 64, generator=g).to(dev)
uidx = torch.randint(0, syn.N_CUSTOMERS, (65536,), generator=g).to(dev)
ms = timeit(lambda: torch.ops.rs.gather_rows(tab, uidx, -1, 0))
report("H1", "gather_rows 65536 of [1.37M, 64] fp32", ms, 65536 * (64 * 4 * 2 + 8), "GB/s", HBM)
seq = rs.synthetic.zipf_ids(8192 * 50, syn.N_ITEMS, 1.05, g).to(dev)
tab2 = torch.randn(syn.N_ITEMS + 1, 128, generator=g).to(dev)
ms = timeit(lambda: torch.ops.rs.gather_rows(tab2, seq, -1, 0))
report("H1", "gather_rows 409600 (Zipf ids) of [105543, 128] fp32", ms, 409600 * (128 * 4 * 2 + 8), "GB/s", HBM)
gsd = torch.randn(409600, 128, device=dev)
rs.ops._sort_cache.clear()
ms = timeit(lambda: (rs.ops._sort_cache.clear(), torch.ops.rs.embedding_dense_bwd(gsd, seq, syn.N_ITEMS + 1, 0, -1, True)))
report("U3/I3", "embedding_dense_bwd sorted (sort + segment reduce + 54 MB memset), 409600 x 128 fp32 grads", ms,
       409600 * (128 * 4 + 8) + 2 * 54e6, "GB/s", HBM)
ms = timeit(lambda: torch.ops.rs.embedding_dense_bwd(gsd, seq, syn.N_ITEMS + 1, 0, -1, False))
report("U3/I3", "embedding_dense_bwd atomic (red.v4.f32), same", ms, 409600 * (128 * 4 + 8) + 2 * 54e6, "GB/s", HBM)
# ---- I2: BERT word rows, B = 512 items x 9 fields x 32 tokens, 768 wide
word = torch.randn(30522, 768, generator=g).to(dev)
tok = torch.randint(0, 30522, (512 * 9 * 32,), generator=g).to(dev)
ms = timeit(lambda: torch.ops.rs.gather_rows(word, tok, -1, 0))
report("I2", "gather_rows 147456 of [30522, 768] fp32", ms, 147456 * (768 * 4 * 2 + 8), "GB/s", HBM)
