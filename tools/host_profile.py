"""Where does the HOST time of one train step go?  torch.profiler over a few steps of the bench workload
(CPU-side op times, both the forward thread and the autograd thread).  Run on the GPU box."""
import importlib
import os
import sys

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
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        rs.train.two_tower_step(model, item, batch, lookup, opt)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=45, max_name_column_width=60))
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=80))
