"""Which stock SDPA backend is fastest for the reference's encoder shape (B=8192, L=50, 4 heads x 32)?"""
import torch, time
import torch.nn as nn
from torch.nn.attention import sdpa_kernel, SDPBackend
dev = "cuda"
B, L, D = 8192, 50, 128
layer = nn.TransformerEncoderLayer(d_model=D, nhead=4, dim_feedforward=2 * D, dropout=0.2, activation="gelu",
                                   norm_first=True, batch_first=True)
enc = nn.TransformerEncoder(layer, num_layers=2).to(dev).train()
x = torch.randn(B, L, D, device=dev, dtype=torch.bfloat16, requires_grad=True)
lens = torch.randint(1, L + 1, (B,), device=dev)
pad = torch.arange(L, device=dev).unsqueeze(0) < (L - lens).unsqueeze(1)
causal = torch.triu(torch.ones(L, L, device=dev, dtype=torch.bool), diagonal=1)
def run():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = enc(x, mask=causal, src_key_padding_mask=pad)
    y.float().nan_to_num().sum().backward()
for name, be in (("default", None), ("efficient", [SDPBackend.EFFICIENT_ATTENTION]), ("cudnn", [SDPBackend.CUDNN_ATTENTION]),
                 ("math", [SDPBackend.MATH]), ("eff+math", [SDPBackend.EFFICIENT_ATTENTION, SDPBackend.MATH])):
    try:
        ctx = sdpa_kernel(be) if be else torch.autocast("cuda", enabled=False)
        with ctx:
            for _ in range(3): run()
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(5): run()
            torch.cuda.synchronize()
        print(f"{name:10s} {(time.perf_counter() - t0) / 5 * 1e3:8.2f} ms fwd+bwd", flush=True)
    except Exception as e:
        print(name, "failed:", str(e)[:200], flush=True)
