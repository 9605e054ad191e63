"""Microbenchmark + quick parity check of the fused softmax kernels on the main-loss shape of the bench step
(distinct-item columns: M = 103,976 rows x N = 21,435 sorted column keys, column bias, own column masked by key,
RS_CE_NO_DIAG, bounded logits) and on the plain DuoRec shape (8192 x 8192).  CUDA-event timing, L2 flushed."""
import importlib, os, sys, torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rs = importlib.import_module("llm-driven_content-based-feature_recommendation_system_b200")
L = rs._lib
dev = "cuda"
NO_DIAG = L.RS_CE_NO_DIAG
BOUND = rs.losses.UNIT_NORM_BOUND


def problem(M, N, seed=0, dtype=torch.bfloat16):
    g = torch.Generator().manual_seed(seed)
    a = F.normalize(torch.randn(M, 128, generator=g), dim=1).to(dev).to(dtype)
    b = F.normalize(torch.randn(N, 128, generator=g), dim=1).to(dev).to(dtype)
    col_ids = torch.sort(torch.randperm(5 * N, generator=g)[:N]).values.to(dev)
    w = torch.arange(1, N + 1, dtype=torch.float64).pow(-1.05)
    own = torch.multinomial(w, M, replacement=True, generator=g).to(dev)
    tgt = col_ids[own]
    bias = (torch.randn(N, generator=g) * 2 - 8).to(dev)
    return a, b, bias, tgt, col_ids, own


def check(M=777, N=1500):
    a, b, bias, tgt, col_ids, own = problem(M, N, 1)
    args = (10.0, bias, tgt, col_ids, None, None, 0, float("-inf"), NO_DIAG)
    lse = torch.ops.rs.ce_fwd(a, b, *args, BOUND)[0]
    S = 10.0 * a.float() @ b.float().T - bias[None, :]
    S[torch.arange(M, device=dev), own] = float("-inf")
    ref = torch.logsumexp(S, 1)
    e_f = (lse - ref).abs().max().item()
    wl = torch.rand(M, device=dev) / M
    dA, dB = torch.ops.rs.ce_bwd(a, b, *args, lse, wl, None, None, BOUND)
    P = torch.softmax(S, 1) * wl[:, None]
    rA, rB = 10.0 * P @ b.float(), 10.0 * P.T @ a.float()
    e_a = ((dA - rA).abs().max() / rA.abs().max()).item()
    e_b = ((dB - rB).abs().max() / rB.abs().max()).item()
    print(f"check general: lse err {e_f:.2e}  dA rel {e_a:.2e}  dB rel {e_b:.2e}")
    # plain
    args = (10.0, None, None, None, None, None, 0, float("-inf"), 0)
    out = torch.ops.rs.ce_fwd(a, b[:M], *args, BOUND)
    S = 10.0 * a.float() @ b[:M].float().T
    ref = torch.logsumexp(S, 1)
    e_f = (out[0] - ref).abs().max().item()
    dA, dB = torch.ops.rs.ce_bwd(a, b[:M], *args, out[0], wl, -wl, None, BOUND)
    P = torch.softmax(S, 1) * wl[:, None] - torch.diag(wl)
    rA, rB = 10.0 * P @ b[:M].float(), 10.0 * P.T @ a.float()
    e_a = ((dA - rA).abs().max() / rA.abs().max()).item()
    e_b = ((dB - rB).abs().max() / rB.abs().max()).item()
    print(f"check plain  : lse err {e_f:.2e}  dA rel {e_a:.2e}  dB rel {e_b:.2e}")


def timeit(fn, reps=5):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def bench(M, N, general=True, label=""):
    a, b, bias, tgt, col_ids, own = problem(M, N)
    if general:
        args = (10.0, bias, tgt, col_ids, None, None, 0, float("-inf"), NO_DIAG)
    else:
        args = (10.0, None, None, None, None, None, 0, float("-inf"), 0)
    wl = torch.full((M,), 1.0 / M, device=dev)
    out = torch.ops.rs.ce_fwd(a, b, *args, BOUND)
    tf = timeit(lambda: torch.ops.rs.ce_fwd(a, b, *args, BOUND))
    tb = timeit(lambda: torch.ops.rs.ce_bwd(a, b, *args, out[0], wl, None if general else -wl, None, BOUND))
    fl = 2.0 * M * N * 128
    print(f"{label} M={M} N={N}: fwd {tf:.3f} ms {fl / tf / 1e9:.0f} TF/s ({fl / tf / 1e9 / 1361:.2f}) | "
          f"bwd {tb:.3f} ms {4 * fl / tb / 1e9:.0f} TF/s ({4 * fl / tb / 1e9 / 1361:.2f})")


if __name__ == "__main__":
    check()
    bench(103976, 21435, True, "main ")
    bench(8192, 8192, False, "duorec")
    bench(32768, 32768, False, "plain ")
