"""Diagnostics for the tcgen05 fused softmax kernel (development aid; compares against a torch fp32
computation on the same GPU so that it runs at any size)."""
import importlib
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rs = importlib.import_module("llm-driven_content-based-feature_recommendation_system_b200")
L = rs._lib
dev = "cuda"


def ref(a, b, scale, bias=None, ka=None, kb=None, off=0, maskv=float("-inf"), flags=0):
    a32, b32 = a.float(), b.float()
    s = a32 @ b32.T * scale
    raw = s.clone()
    if bias is not None:
        s = s - bias.view(1, -1)
    M, N = s.shape
    lab = torch.arange(M, device=a.device) + off
    isdiag = torch.zeros_like(s, dtype=torch.bool)
    isdiag[torch.arange(M, device=a.device), lab] = True
    if flags & L.RS_CE_DIAG_RAW:
        s = torch.where(isdiag, raw, s)
    mask = torch.zeros_like(s, dtype=torch.bool)
    pos = None
    if flags & L.RS_CE_SUPCON:
        pos = (ka[0].view(-1, 1) == ka[1].view(1, -1)) & (ka[0].view(-1, 1) != 0) & ~isdiag
    else:
        if ka is not None:
            mask |= ka[0].view(-1, 1) == ka[1].view(1, -1)
        if kb is not None:
            mask |= kb[0].view(-1, 1) == kb[1].view(1, -1)
        mask &= ~isdiag
    if flags & L.RS_CE_DIAG_MASK:
        mask |= isdiag
    s = s.masked_fill(mask, maskv)
    lse = torch.logsumexp(s, dim=1)
    diag = s[torch.arange(M, device=a.device), lab]
    return s, lse, diag, pos


def check(name, M, N, dtype=torch.bfloat16, bias=False, keys=False, off=0, flags=0, scale=10.0, bwd=True):
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N)
    a = F.normalize(torch.randn(M, 128, generator=g), dim=1).to(dev).to(dtype)
    b = F.normalize(torch.randn(N, 128, generator=g), dim=1).to(dev).to(dtype)
    cb = (torch.randn(N, generator=g) * 2).to(dev) if bias else None
    ka = kb = None
    if keys or flags & L.RS_CE_SUPCON:
        kc = torch.randint(0, max(2, N // 4), (N,), generator=g).to(dev)
        kr = kc[torch.arange(M) + off] if M + off <= N else torch.randint(0, max(2, N // 4), (M,), generator=g).to(dev)
        ka = (kr, kc)
        if keys and not flags & L.RS_CE_SUPCON:
            uc = torch.randint(0, max(2, N // 2), (N,), generator=g).to(dev)
            ur = uc[torch.arange(M) + off] if M + off <= N else uc[:M]
            kb = (ur, uc)
    args = (float(scale), cb, ka[0] if ka else None, ka[1] if ka else None, kb[0] if kb else None,
            kb[1] if kb else None, off, float("-inf"), flags)
    lse, diag, ps, pc = torch.ops.rs.ce_fwd(a, b, *args)
    torch.cuda.synchronize()
    s, lse_r, diag_r, pos = ref(a, b, scale, cb, ka, kb, off, float("-inf"), flags)
    e1 = (lse - lse_r).abs().max().item()
    fin = torch.isfinite(diag_r)
    e2 = (diag - diag_r)[fin].abs().max().item() if fin.any() else 0.0
    msg = f"{name:28s} M={M:6d} N={N:6d} {str(dtype)[6:]:9s} lse_err={e1:.2e} diag_err={e2:.2e}"
    if flags & L.RS_CE_SUPCON:
        ps_r = (s.masked_fill(~pos, 0.0)).sum(1)
        msg += f" pos_sum_err={(ps - ps_r).abs().max().item():.2e} cnt_err={(pc - pos.sum(1).float()).abs().max().item():.1f}"
    if bwd:
        wl = torch.rand(M, device=dev) / M
        wd = -torch.rand(M, device=dev) / M if not flags & L.RS_CE_DIAG_MASK else None
        wp = torch.rand(M, device=dev) / M if flags & L.RS_CE_SUPCON else None
        dA, dB = torch.ops.rs.ce_bwd(a, b, *args, lse, wl, wd, wp)
        torch.cuda.synchronize()
        P = torch.exp(s - lse_r.view(-1, 1))
        dS = wl.view(-1, 1) * P
        if wd is not None:
            dS[torch.arange(M, device=dev), torch.arange(M, device=dev) + off] += wd
        if wp is not None:
            dS = dS + wp.view(-1, 1) * pos.float()
        dA_r = scale * dS @ b.float()
        dB_r = scale * dS.T @ a.float()
        msg += f" dA_rel={(dA - dA_r).abs().max().item() / dA_r.abs().max().item():.2e}"
        msg += f" dB_rel={(dB - dB_r).abs().max().item() / dB_r.abs().max().item():.2e}"
    print(msg, flush=True)


if __name__ == "__main__":
    bwd = "--nobwd" not in sys.argv
    check("plain 128", 128, 128, bwd=bwd)
    check("plain 128 fp16", 128, 128, dtype=torch.float16, bwd=bwd)
    check("plain 256x384", 256, 384, bwd=bwd)
    check("plain ragged", 200, 300, bwd=bwd)
    check("bias", 256, 256, bias=True, bwd=bwd)
    check("bias+keys", 384, 384, bias=True, keys=True, bwd=bwd)
    check("bias+keys ragged", 1000, 1000, bias=True, keys=True, bwd=bwd)
    check("rect offset", 200, 600, bias=True, keys=True, off=200, bwd=bwd)
    check("diag raw", 300, 300, bias=True, keys=True, flags=L.RS_CE_DIAG_RAW, bwd=bwd)
    check("supcon", 500, 500, flags=L.RS_CE_DIAG_MASK | L.RS_CE_SUPCON, bwd=bwd)
    check("big", 8192, 8192, bias=True, keys=True, bwd=bwd)
    check("tall", 20000, 1000, bias=True, keys=True, off=0, bwd=bwd) if False else None
    import time
    for (M, N) in ((8192, 8192), (32768, 32768)):
        a = F.normalize(torch.randn(M, 128), dim=1).to(dev).bfloat16()
        b = F.normalize(torch.randn(N, 128), dim=1).to(dev).bfloat16()
        kk = torch.randint(0, 50000, (N,), device=dev)
        cb = torch.randn(N, device=dev)
        for label, extra in (("plain", (None, None, None, None, None)), ("general", (cb, kk[:M], kk, kk[:M], kk))):
            args = (10.0, *extra, 0, float("-inf"), 0)
            for _ in range(2):
                out = torch.ops.rs.ce_fwd(a, b, *args)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                out = torch.ops.rs.ce_fwd(a, b, *args)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 5
            print(f"fwd {label:8s} {M}x{N}: {dt * 1e3:.3f} ms  {2.0 * M * N * 128 / dt / 1e12:.1f} TFLOP/s", flush=True)
            if bwd:
                wl = torch.full((M,), 1.0 / M, device=dev)
                for _ in range(2):
                    g = torch.ops.rs.ce_bwd(a, b, *args, out[0], wl, -wl, None)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(5):
                    g = torch.ops.rs.ce_bwd(a, b, *args, out[0], wl, -wl, None)
                torch.cuda.synchronize()
                dt = (time.perf_counter() - t0) / 5
                print(f"bwd {label:8s} {M}x{N}: {dt * 1e3:.3f} ms  {8.0 * M * N * 128 / dt / 1e12:.1f} TFLOP/s", flush=True)
