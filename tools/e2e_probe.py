"""Per-step wall times of the end-to-end loop (host batches in, losses out) + allocator statistics."""
import importlib, os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rs = importlib.import_module("llm-driven_content-based-feature_recommendation_system_b200")
syn = rs.synthetic
dev = torch.device("cuda", 0)
torch.manual_seed(42)
B, SL = 8192, 50
model = rs.SASRecUserTower(syn.tower_args(max_len=SL)).to(dev).train()
item = rs.SASRecItemTower(syn.N_ITEMS, 128, syn.log_q(syn.N_ITEMS)).to(dev)
lookup = syn.pretrained_table(syn.N_ITEMS).to(dev)
item.init_from_pretrained(lookup)
opt = torch.optim.AdamW(list(model.parameters()) + list(item.parameters()), lr=5e-4, weight_decay=1e-4, fused=True)
pool = 3
host = [{k: v.pin_memory() for k, v in rs.train.add_host_index(syn.make_batch(B, SL, syn.N_ITEMS, seed=42 + i)).items()}
        for i in range(pool)]
print("tensors per batch", len(host[0]), "bytes", sum(v.numel() * v.element_size() for v in host[0].values()))
def one(i, read=True):
    t0 = time.perf_counter()
    b = rs.train.prepare_batch(host[i % pool], dev, non_blocking=True)
    t1 = time.perf_counter()
    t, m, c = rs.train.two_tower_step(model, item, b, lookup, opt)
    t2 = time.perf_counter()
    if read:
        x = (t.item(), m.item(), c.item())
    t3 = time.perf_counter()
    return (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3
for i in range(6):
    one(i)
torch.cuda.synchronize()
s0 = torch.cuda.memory_stats()
for i in range(12):
    print("step %d: copy-enqueue %.2f ms, step-enqueue %.2f ms, wait %.2f ms" % ((i,) + one(i)))
s1 = torch.cuda.memory_stats()
for k in ("num_alloc_retries", "num_device_alloc", "num_device_free", "num_sync_all_streams"):
    print(k, s0.get(k), "->", s1.get(k))
print("reserved GB", torch.cuda.memory_reserved() / 2**30, "allocated GB", torch.cuda.max_memory_allocated() / 2**30)
