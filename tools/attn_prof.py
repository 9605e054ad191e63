"""Attention kernels at the bench step's shape (two dropout views of the B=8192 batch, 4 heads x 32, mean length 12.7
+ the literal-DuoRec one-token pseudo-sequences): CUDA-event timing, or a driver for ncu."""
import importlib, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rs = importlib.import_module("llm-driven_content-based-feature_recommendation_system_b200")
dev = "cuda"
syn = rs.synthetic
hb = rs.train.add_host_index(syn.make_batch(8192, 50, syn.N_ITEMS, seed=42))
cu = hb["cu_seqlens_2v"].to(dev)
T = int(cu[-1])
B = 8192
zero_tail = cu.numel() - 1 - 2 * B
print("tokens", T, "seqs", cu.numel() - 1, "zero_tail", zero_tail)
qkv = (torch.randn(T, 384, device=dev) * 0.5).bfloat16()
bias = torch.randn(384, device=dev) * 0.1
seed = 1234
args = (4, 50, zero_tail, 32 ** -0.5, 0.2, seed)
out, lse = torch.ops.rs.attn_varlen(qkv, bias, cu, *args)
g = torch.randn_like(out)
dq, db = torch.ops.rs.attn_varlen_bwd(qkv, bias, g, out, lse, cu, *args)
torch.cuda.synchronize()
if len(sys.argv) > 1 and sys.argv[1] == "once":
    sys.exit(0)


def timeit(fn, reps=7):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


tf = timeit(lambda: torch.ops.rs.attn_varlen(qkv, bias, cu, *args))
tb = timeit(lambda: _lib_bwd())  if False else timeit(lambda: rs.encoder._lib.rs_attn_varlen_bwd(
    rs._lib.ptr(qkv), rs._lib.ptr(g), rs._lib.ptr(out), rs._lib.dt(qkv), rs._lib.ptr(bias), rs._lib.ptr(lse), rs._lib.ptr(cu),
    cu.numel() - 1, T, 4, 32, 50, zero_tail, 32 ** -0.5, 0.2, seed, rs._lib.ptr(dq), rs._lib.stream()))
byt_f = T * 384 * 2 + T * 128 * 2
byt_b = T * 384 * 2 * 2 + T * 128 * 2 * 2
print(f"attn fwd {tf:.3f} ms ({byt_f / tf / 1e6:.0f} GB/s algorithmic) | bwd {tb:.3f} ms ({byt_b / tb / 1e6:.0f} GB/s)")
