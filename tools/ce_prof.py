"""Short driver for ncu captures of the fused softmax kernels (bench-like key structure: sorted user ids)."""
import importlib, os, sys, torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rs = importlib.import_module("llm-driven_content-based-feature_recommendation_system_b200")
dev = "cuda"
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
a = F.normalize(torch.randn(N, 128), dim=1).to(dev).bfloat16()
b = F.normalize(torch.randn(N, 128), dim=1).to(dev).bfloat16()
tgt = torch.randint(0, 100000, (N,), device=dev)
uid = torch.arange(N, device=dev) // 12
cb = torch.randn(N, device=dev)
args = (10.0, cb, tgt, tgt, uid, uid, 0, float("-inf"), 0)
wl = torch.full((N,), 1.0 / N, device=dev)
import time
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = torch.ops.rs.ce_fwd(a, b, *args)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    g = torch.ops.rs.ce_bwd(a, b, *args, out[0], wl, -wl, None)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"N={N} fwd {1e3*(t1-t0):.3f} ms {2.0*N*N*128/(t1-t0)/1e12:.0f} TF/s | bwd {1e3*(t2-t1):.3f} ms {8.0*N*N*128/(t2-t1)/1e12:.0f} TF/s")
