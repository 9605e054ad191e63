"""CUDA-event timing of the user-block kernels (fwd, bwd) at the bench step's shapes (B = 8192 users, ~104k rows, 22k
columns).  (used to compare the CTA-per-user kernels with a warp-per-user variant for short users, which lost)."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rs = importlib.import_module("llm-driven_content-based-feature_recommendation_system_b200")
syn, tr = rs.synthetic, rs.train
dev = torch.device("cuda:0")
B, SL = 8192, 50
hb = syn.make_batch(B, SL, syn.N_ITEMS, seed=3)
d = {k: v.to(dev) for k, v in hb.items()}
m = rs.ops.batch_index_counts(d["padding_mask"], d["target_ids"], syn.N_ITEMS + 1).cpu()
tok_cap, col_cap = tr.bucket_of(int(m[0]), int(m[2]))
idx = tr.device_index(d, syn.N_ITEMS + 1, tok_cap, col_cap)
n = idx["main_tgt"].numel()
u = torch.nn.functional.normalize(torch.randn(n, 128, device=dev), dim=1).bfloat16().requires_grad_(True)
c = torch.nn.functional.normalize(torch.randn(col_cap, 128, device=dev), dim=1).bfloat16().requires_grad_(True)
bias = torch.randn(col_cap, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
tf, tb = [], []
for it in range(9):
    flush.zero_()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    torch.cuda.synchronize()
    e[0].record()
    s_pos, own = torch.ops.rs.user_block_logits(u, c, idx["pos_col"], idx["row_cu"], SL, 10.0, bias)
    e[1].record()
    g1, g2 = torch.randn_like(s_pos), torch.randn_like(own)
    flush.zero_()
    e[2].record()
    du, dc = torch.ops.rs.user_block_logits_bwd(u, c, idx["pos_col"], idx["row_cu"], SL, 10.0, bias, own, g1, g2)
    e[3].record()
    torch.cuda.synchronize()
    tf.append(e[0].elapsed_time(e[1])); tb.append(e[2].elapsed_time(e[3]))
tf.sort(); tb.sort()
print(f"RS_UB_SHORT_SPLIT={os.environ.get('RS_UB_SHORT_SPLIT', '1')}: fwd {tf[4]*1e3:.1f} us  bwd {tb[4]*1e3:.1f} us  (incl. zero fills)")
