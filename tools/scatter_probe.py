"""Per-kernel breakdown of the deterministic scatter (U3) at config 2's full [8192, 50] grid; run under
`ncu --metrics gpu__time_duration.sum` to list the launches."""
import importlib, sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rs = importlib.import_module("llm-driven_content-based-feature_recommendation_system_b200")
syn = rs.synthetic
dev = torch.device("cuda:0")
B, SL, D, NI = 8192, 50, 128, syn.N_ITEMS
b = syn.make_batch(B, SL, NI)
ids = b["item_ids"].to(dev)
cot = torch.randn(B, SL, D, device=dev).bfloat16()
for det in (True, False, True, False):
    rs.ops._sort_cache.clear()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g = torch.ops.rs.embedding_dense_bwd(cot.view(-1, D), ids.view(-1), NI + 1, 0, -1, det)
    e1.record()
    torch.cuda.synchronize()
    print("deterministic" if det else "atomic", round(e0.elapsed_time(e1), 4), "ms", float(g.abs().sum()))
