"""Who launches the small torch kernels of a step?  torch.profiler with Python stacks, aten ops that own a CUDA kernel
grouped by (op, calling source line inside this package).  Run on the GPU box."""
import collections, importlib, os, sys
import torch
from torch.profiler import ProfilerActivity, profile

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
batch = rs.train.prepare_batch(rs.train.add_host_index(syn.make_batch(B, SL, syn.N_ITEMS, seed=42)), dev)
for _ in range(3):
    rs.train.two_tower_step(model, item, batch, lookup, opt)
torch.cuda.synchronize()
N = 2
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    for _ in range(N):
        rs.train.two_tower_step(model, item, batch, lookup, opt)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type != torch.autograd.DeviceType.CPU or not ev.name.startswith("aten::"):
        continue
    ct = getattr(ev, "self_device_time_total", 0) or 0
    if ct <= 0:
        continue
    where = "?"
    for fr in (ev.stack or []):
        if "recommendation_system_b200" in fr and "site-packages" not in fr:
            where = fr.split("recommendation_system_b200/")[-1][:70]
            break
    else:
        for fr in (ev.stack or []):
            if "torch/autograd" in fr or "optim" in fr or "clip_grad" in fr:
                where = fr.split("site-packages/")[-1][:70]
                break
    a = agg[(ev.name, where)]
    a[0] += 1
    a[1] += ct
tot = sum(v[1] for v in agg.values())
print(f"aten ops with own CUDA time: {tot / N / 1e3:.3f} ms/step")
for (name, where), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print(f"{t / N / 1e3:8.3f} ms  x{c / N:6.1f}  {name:34s} {where}")
