"""Driver for ncu captures / timing of the HBM-bound front kernels at BASELINE config-2 size."""
import importlib, os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rs = importlib.import_module("llm-driven_content-based-feature_recommendation_system_b200")
syn = rs.synthetic
dev = "cuda"
B, L, D, NI = 8192, 50, 128, syn.N_ITEMS
b = syn.make_batch(B, L, NI)
g = torch.Generator().manual_seed(0)
item_tab = (torch.randn(NI + 1, D, generator=g) * 0.02).to(dev)
time_tab = (torch.randn(12, D, generator=g) * 0.02).to(dev)
pos = (torch.randn(L, D, generator=g) * 0.02).to(dev)
ids = [b["item_ids"].to(dev), b["time_bucket_ids"].to(dev)]
gates = torch.tensor([0.7, 0.4], device=dev)
base = torch.randn(B, L, D, generator=g).to(dev).bfloat16()
cot = torch.randn(B, L, D, generator=g).to(dev).bfloat16()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(name, fn, bytes_, n=10):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()                                   # evict L2 between iterations
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print(f"{name:28s} {ms:7.3f} ms  {bytes_ / ms / 1e6:8.0f} GB/s algorithmic", flush=True)
P = B * L
timeit("seq_front_fwd (bf16 io)", lambda: torch.ops.rs.seq_front(base, ids, [item_tab, time_tab], gates, pos, L, 2), P * (256 + 1024 + 256 + 16))
timeit("seq_front_bwd det (dense)", lambda: torch.ops.rs.seq_front_bwd(cot, ids, [item_tab, time_tab], gates, L, 0, True), P * (256 + 8) * 2)
timeit("seq_front_bwd atomics", lambda: torch.ops.rs.seq_front_bwd(cot, ids, [item_tab, time_tab], gates, L, 0, False), P * (256 + 16 + 512))
timeit("embedding_dense_bwd det", lambda: torch.ops.rs.embedding_dense_bwd(cot.view(-1, D), ids[0].view(-1), NI + 1, 0, -1, True), P * (256 + 8))
timeit("gather_rows fp32", lambda: torch.ops.rs.gather_rows(item_tab, ids[0], -1, 0), P * (512 + 512 + 8))
big = torch.randn(1371981, 64, device=dev)
uidx = torch.randint(0, 1371981, (1 << 20,), device=dev)
timeit("gather_rows 1.37M x64", lambda: torch.ops.rs.gather_rows(big, uidx, -1, 0), (1 << 20) * (256 + 256 + 8))
vocab = syn.criteo_vocab_sizes()
fm = rs.FM(vocab, k=16).to(dev)
fids = syn.make_fm_batch(65536, vocab).to(dev)
timeit("fm_fwd 65536x39 k16", lambda: torch.ops.rs.fm_fwd(fids, fm.offsets, fm.embedding, fm.linear.view(-1), True, 0), 65536 * 39 * (64 + 8 + 64 + 4))
