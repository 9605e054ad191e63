"""Where do the torch-native (non rs::) kernels of one eager train step come from?  torch.profiler with stacks; prints,
per aten op that launches a kernel, count / total us / shapes / innermost repo frame."""
import importlib, os, sys, collections
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rs = importlib.import_module("llm-driven_content-based-feature_recommendation_system_b200")
syn, tr = rs.synthetic, rs.train
dev = torch.device("cuda:0")
B, SL = 8192, 50
n_rows = syn.N_ITEMS + 1
torch.manual_seed(42)
model = rs.SASRecUserTower(syn.tower_args(max_len=SL)).to(dev).train()
item = rs.SASRecItemTower(syn.N_ITEMS, 128, syn.log_q(syn.N_ITEMS)).to(dev)
lookup = syn.pretrained_table(syn.N_ITEMS).to(dev)
item.init_from_pretrained(lookup)
params = list(model.parameters()) + list(item.parameters())
opt = torch.optim.AdamW(params, lr=5e-4, weight_decay=1e-4, fused=True, capturable=True)

def step(b):
    return tr.two_tower_step(model, item, b, lookup, opt, loss_scope="all", amp_dtype=torch.bfloat16, columns="unique")

bs = tr.BucketedStep(step, B, SL, n_rows, dev, use_graph=False)
fb = tr.FlatBatch(B, SL, pin=True)
for i in range(3):
    fb.fill(syn.make_batch(B, SL, syn.N_ITEMS, seed=i))
    bs.raw.copy_(fb)
    m = bs.counts(bs.raw).cpu()
    key = bs.bucket(int(m[0]), int(m[2]))
    bs.run(key)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True, record_shapes=True,
             experimental_config=torch._C._profiler._ExperimentalConfig(verbose=True)) as prof:
    bs.run(key)
    torch.cuda.synchronize()
ev = prof.events()
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
agg = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    if e.device_type != torch.autograd.DeviceType.CPU or not e.name.startswith("aten::"):
        continue
    kern = [k for k in e.kernels] if hasattr(e, "kernels") else []
    if not kern:
        continue
    # only leaf aten ops (those whose children did not launch the same kernels)
    if any(c.name.startswith("aten::") and getattr(c, "kernels", []) for c in e.cpu_children):
        continue
    real = os.path.realpath(root)
    frames = [s for s in (e.stack or []) if "_b200/" in s or "bench.py" in s]
    site = frames[0].split("_b200/")[-1] if frames else (e.stack[0] if e.stack else "?")
    bwd = "" if frames else " [autograd engine]"
    key_ = (e.name, str(e.input_shapes)[:70], " <- ".join(f.split("_b200/")[-1][:48] for f in frames[:3]) + bwd if frames else site[:110] + bwd)
    agg[key_][0] += 1
    agg[key_][1] += sum(k.duration for k in kern)
tot = 0.0
for k_, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{us:8.1f} us x{c:3d}  {k_[0]:28s} {k_[1]:70s} {k_[2]}")
    tot += us
print("total us", round(tot, 1))
