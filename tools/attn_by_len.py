"""Where does the tile attention spend its time?  fwd / bwd CUDA-event times of attn_varlen on the step's sequences (two
views of 8192 users), split by sequence length class."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rs = importlib.import_module("llm-driven_content-based-feature_recommendation_system_b200")
syn = rs.synthetic
dev = torch.device("cuda:0")
b = syn.make_batch(8192, 50, syn.N_ITEMS, seed=1)
lens_all = (~b["padding_mask"]).sum(1)
lens_all = torch.cat([lens_all, lens_all])          # two views
H = 4
bias = torch.randn(3 * H * 32, device=dev) * 0.1


def run(lens, label):
    cu = torch.zeros(lens.numel() + 1, dtype=torch.int32)
    cu[1:] = torch.cumsum(lens, 0)
    cu = cu.to(dev)
    T = int(lens.sum())
    qkv = (torch.randn(T, 3 * H * 32, device=dev) * 0.5).bfloat16().requires_grad_(True)
    cot = torch.randn(T, H * 32, device=dev).bfloat16()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    tf, tb = [], []
    for it in range(7):
        flush.zero_()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        torch.manual_seed(it)
        e[0].record()
        out = rs.encoder.attn_varlen(qkv, cu, H, 50, dropout_p=0.1, bias=bias)
        e[1].record()
        out.backward(cot)
        e[2].record()
        torch.cuda.synchronize()
        tf.append(e[0].elapsed_time(e[1])); tb.append(e[1].elapsed_time(e[2]))
        qkv.grad = None
    tf.sort(); tb.sort()
    t = ((lens + 15) // 16)
    pairs = int((t * (t + 1) // 2).sum())
    print(f"{label:12s} seqs {lens.numel():6d} tokens {T:7d} tile-pairs {pairs:6d}  fwd {tf[3]*1e3:7.1f} us  bwd(+colsum) {tb[3]*1e3:7.1f} us")


run(lens_all, "all")
for lo, hi in [(1, 8), (9, 16), (17, 32), (33, 64)]:
    m = (lens_all >= lo) & (lens_all <= hi)
    run(lens_all[m], f"len {lo}-{hi}")
