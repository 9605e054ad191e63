#!/usr/bin/env python
"""bench.py -- two-tower train samples/sec (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): UserTower + ItemTower in-batch InfoNCE training, batch 8192 per
GPU, L=50, d=128, bf16 autocast, synthetic H&M-shaped data (105,542 articles).  One step = the body of
the reference's `train_user_tower_all_time` (tower_code/v1_usertower_train.py:729-859): two dropout
views of SASRecUserTower, logQ-corrected InfoNCE over ALL valid time steps (N ~ 98k rows) with same-item /
same-user masks, DuoRec on the last step, backward, grad clip, AdamW on both towers.

  value : samples/s with the batch already resident in HBM (device-timed, max over ranks)
  e2e   : the same step through the public API with HOST (pinned) batches: H2D copies of every input and
          a D2H read of the losses inside the timed region
  --impl reference : the reference's algorithm for the same step on the host CPU (oracle port, all host
          threads) on a bounded sample of the workload (B=256 users of the same batch)
"""
import argparse
import os
# one-time allocator choice for a workload whose per-batch shapes differ (valid tokens, distinct items): growing the
# pool by mapping pages instead of cudaMalloc-ing new blocks avoids multi-ms allocation stalls inside timed steps
if int(os.environ.get("WORLD_SIZE", "1")) == 1:
    # (not for N > 1: replaying captured graphs that contain NCCL collectives deadlocked on this stack with expandable
    # segments -- both ranks stuck behind the first replays, gpurun_out/r2_b2_graphA.log -- and runs fine without)
    os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "llm-driven_content-based-feature_recommendation_system_b200"
METRIC = "two_tower_train_samples_per_sec"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch_gpu"])
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--seq-len", type=int, default=50)
    ap.add_argument("--loss-scope", default="all", choices=["all", "last"])
    ap.add_argument("--columns", default="unique", choices=["unique", "catalog", "batch"],
                    help="column enumeration of the main in-batch softmax (same loss): distinct batch items with "
                         "multiplicities / whole catalogue / one column per row as the reference materialises it")
    ap.add_argument("--parallelism", default="sharded", choices=["sharded", "dp"],
                    help="N>1: row-sharded item tables (all-to-all lookups) + negatives spanning the box [default], "
                         "or plain data-parallel replicas with all-reduced gradients and rank-local negatives")
    ap.add_argument("--cuda-graph", type=int, default=1,
                    help="1: capture the whole step in one CUDA graph per pooled batch and replay it (default, 1 GPU); "
                         "0: eager launches")
    ap.add_argument("--cuda-graph-multi", type=int, default=1,
                    help="N > 1: capture the sharded step (NCCL collectives included) per shape bucket [1], or launch eagerly [0]")
    ap.add_argument("--cpu-batch", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--minimal", action="store_true", help="only the resident-input timed loop (for ncu runs)")
    ap.add_argument("--tok-bucket", type=int, default=2048, help="shape bucket of the main-loss rows (1 GPU)")
    ap.add_argument("--col-bucket", type=int, default=512, help="shape bucket of the distinct-target columns (1 GPU)")
    ap.add_argument("--no-extra", action="store_true", help="skip the other BASELINE configs' kernel-level numbers")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_step_rate(args, steps, warmup, syn):
    """The reference's step on the host cores: oracle port (fp32, autocast off -- what the reference runs
    on a CPU), B=cpu_batch users cut from the same synthetic batch, all-timestep loss."""
    from oracle import towers as otowers
    torch.manual_seed(42)
    nthreads = os.cpu_count() or 1
    torch.set_num_threads(nthreads)
    targs = syn.tower_args(max_len=args.seq_len)
    model = otowers.UserTowerOracle(targs).train()
    item = otowers.ItemMatrixOracle(syn.N_ITEMS, 128, syn.log_q(syn.N_ITEMS))
    lookup = syn.pretrained_table(syn.N_ITEMS)
    with torch.no_grad():
        item.item_matrix.weight.copy_(lookup)
    opt = torch.optim.AdamW(list(model.parameters()) + list(item.parameters()), lr=5e-4, weight_decay=1e-4)
    full = syn.make_batch(args.batch, args.seq_len, syn.N_ITEMS, seed=42)
    Bc = min(args.cpu_batch, args.batch)
    b = {k: v[:Bc] for k, v in full.items()}
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        batch = {k: b[k] for k in syn.FORWARD_KEYS}
        batch["pretrained_vecs"] = lookup[b["item_ids"]]             # v1_usertower_train.py:760
        batch["target_ids"] = b["target_ids"]
        total, _, _ = otowers.all_timestep_step_loss(model, item, batch)
        total.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=5.0)
        opt.step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    return dict(value=Bc / dt, unit="samples/s", cores=nthreads, kind="port",
                sample=f"B={Bc} users of the B={args.batch} synthetic batch, all-timestep loss, fp32, "
                       f"{len(times)} timed steps after {warmup} warm-up, {dt * 1e3:.0f} ms/step"), dt


def gpu_eager_step_rate(syn, dev, B, SL, steps=3, warmup=2):
    """Stock PyTorch eager on the same B200 (the bar SURVEY.md section 2 names: the reference ships no kernels of its
    own): the oracle-port nn.Modules on `cuda`, bf16 autocast, the unfused [N, N] loss as the reference materialises it,
    per-step host-side pretrained lookup as the reference does it (v1_usertower_train.py:760), stock AdamW.
    Baseline leg only -- nothing of the product runs here."""
    from oracle import towers as otowers
    torch.manual_seed(42)
    model = otowers.UserTowerOracle(syn.tower_args(max_len=SL)).to(dev).train()
    item = otowers.ItemMatrixOracle(syn.N_ITEMS, 128, syn.log_q(syn.N_ITEMS)).to(dev)
    lookup = syn.pretrained_table(syn.N_ITEMS)
    with torch.no_grad():
        item.item_matrix.weight.copy_(lookup)
    opt = torch.optim.AdamW(list(model.parameters()) + list(item.parameters()), lr=5e-4, weight_decay=1e-4)
    batches = [syn.make_batch(B, SL, syn.N_ITEMS, seed=900 + i) for i in range(warmup + steps)]
    ts, n_rows = [], 0
    for i, hb in enumerate(batches):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        b = {k: hb[k].to(dev) for k in syn.FORWARD_KEYS}
        b["pretrained_vecs"] = lookup[hb["item_ids"]].to(dev)                       # :760
        b["target_ids"] = hb["target_ids"].to(dev)
        opt.zero_grad()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            total, _, _ = otowers.all_timestep_step_loss(model, item, b)
        total.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=5.0)
        opt.step()
        total.item()                                                                # the loop's .item() reads (:857-859)
        torch.cuda.synchronize()
        if i >= warmup:
            ts.append(time.perf_counter() - t0)
        n_rows = int((~hb["padding_mask"]).sum())
    dt = sum(ts) / len(ts)
    return dict(value=B / dt, unit="samples/s", batch=B, loss_rows=n_rows, ms_per_step=dt * 1e3,
                what="stock PyTorch eager on this GPU: oracle-port modules, bf16 autocast, unfused [N, N] loss, end to end "
                     "from host batches")


def run_torch_gpu(args):
    """`--impl torch_gpu`: the stock-PyTorch-eager arm alone, at the reference's batch (768) and at the largest batch
    whose unfused [N, N] logits fit comfortably (2048: N ~ 26k rows)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    syn = _load_synthetic()
    dev = torch.device("cuda", 0)
    rows = [gpu_eager_step_rate(syn, dev, b, args.seq_len) for b in (768, 2048)]
    best = max(rows, key=lambda r: r["value"])
    print(json.dumps(dict(metric=METRIC, value=best["value"], unit="samples/s", n_gpus=1, steps=3, warmup=2,
                          ms_per_step=best["ms_per_step"], higher_is_better=True, scaling="weak", vs_baseline=None,
                          dtype="bf16", data="synthetic", impl="torch_gpu",
                          config=dict(workload="two_tower_infonce_train_step, stock PyTorch eager", seq_len=args.seq_len,
                                      batches=[r["batch"] for r in rows]), gpu_eager_baseline=rows)))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rs_syn = _load_synthetic()
    steps = max(1, min(args.steps, 20))
    cb, dt = cpu_reference_step_rate(args, steps, min(args.warmup, 1) if args.warmup else 0, rs_syn)
    line = dict(metric=METRIC, value=cb["value"], unit="samples/s", n_gpus=args.gpus, steps=steps, warmup=args.warmup,
                ms_per_step=dt * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", impl="reference",
                config=dict(workload="two_tower_infonce_train_step", batch_per_gpu=args.batch, seq_len=args.seq_len,
                            loss_scope="all", note="reference algorithm (CPU oracle port) on a bounded sample"),
                cpu_baseline=cb,
                e2e=dict(value=cb["value"], unit="samples/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def _load_synthetic():
    """synthetic.py is pure torch: load it without importing the package (the reference arm must not need
    the CUDA library)."""
    spec = importlib.util.spec_from_file_location("_rs_synthetic", os.path.join(ROOT, PKG, "synthetic.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock, throttle reasons and power sampled every 200 ms DURING the timed region.  NVML in-process (a thread
    that sleeps between two ~50 us queries): an `nvidia-smi -lms` child per rank re-initialises the driver interface on
    every sample and slowed the eager (N > 1) step's kernel launches from 15 to 42 ms/step.  Falls back to one
    nvidia-smi child if pynvml is missing."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.thread, self.stop = index, [], None, None, threading.Event()

    def _nvml_loop(self, nv, h):
        names = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap))
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        while True:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                rs_ = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                self.rows.append([str(sm), str(mx)] + ["Active" if rs_ & bit else "Not Active" for _, bit in names] + [f"{pw:.2f}"])
            except Exception:
                pass
            if self.stop.wait(0.2):
                return

    def __enter__(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else self.index
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.thread.start()
            return self
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.thread is not None:
            time.sleep(0.25)                  # make sure at least one sample falls after the last step
            self.stop.set()
            self.thread.join(timeout=2)
            return
        if self.proc is None:
            return
        time.sleep(0.25)                      # make sure at least one sample falls after the last step
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        self.rows = [[x.strip() for x in line.split(",")] for line in out.splitlines() if line.strip()]

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        pw = [float(r[6]) for r in self.rows if len(r) > 6 and r[6].replace(".", "").isdigit()]
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=reasons, samples=len(sm), power_w_max=max(pw) if pw else None)


# ------------------------------------------------------------------------------------------------ our arm
def _peaks():
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return pk.get("hbm_gbs", 6650.0), pk.get("bf16_tflops_sustained", 1400.0), "measured"
    except Exception:
        return 6650.0, 1400.0, "fallback"


def _kernel_table(prof, nprof, P, D, step_ms):
    """per-C-ABI-call CUDA-event times of `nprof` eager steps -> {name: {...}}, plus the fused-softmax aggregate.
    Algorithmic work: SURVEY.md 8(d) -- 2*M*N*128 flops per contraction, FOUR contractions per loss (S forward, one S
    recompute, dU, dV); the fused forward executes two of them (S, P@B), the backward the other two (S, dS^T@A)."""
    hbm_peak, tf_peak, peak_src = _peaks()
    agg = {}
    for name, s, e, extra in prof:
        a = agg.setdefault(name, dict(ms=0.0, calls=0, work=0.0))
        a["ms"] += s.elapsed_time(e)
        a["calls"] += 1
        a["work"] += extra
    kernels = {}
    ce = dict(ms=0.0, flops=0.0)
    for name, a in agg.items():
        per_step_ms = a["ms"] / nprof
        k = dict(ms_per_step=round(per_step_ms, 4), calls_per_step=a["calls"] / nprof)
        if name in ("rs_ce_fwd", "rs_ce_bwd", "rs_ce_fwd_grad", "rs_ce_bwd_from_grad"):
            k.update(bound="tensor", unit="TFLOP/s", achieved=a["work"] / nprof / (per_step_ms * 1e-3) / 1e12, peak=tf_peak)
            ce["ms"] += per_step_ms
            ce["flops"] += a["work"] / nprof
        elif name == "rs_seq_front_fwd":
            # per position: base bf16 + the item-id row fp32 + out bf16 + 3 ids (time / position rows: 12- and
            # 51-row tables, cache resident, not counted)
            by = P * (D * 2 + D * 4 + D * 2 + 3 * 8) * a["calls"] / nprof
            k.update(bound="hbm", unit="GB/s", achieved=by / (per_step_ms * 1e-3) / 1e9, peak=hbm_peak)
        elif name == "rs_seq_front_bwd":
            by = P * (D * 2 + 8) * a["calls"] / nprof          # one pass over dX (bf16) + time ids
            k.update(bound="hbm", unit="GB/s", achieved=by / (per_step_ms * 1e-3) / 1e9, peak=hbm_peak)
        elif name == "rs_segment_reduce_rows":
            k.update(bound="hbm", unit="GB/s", achieved=a["work"] / nprof / (per_step_ms * 1e-3) / 1e9, peak=hbm_peak)
        if "achieved" in k:
            k["frac"] = k["achieved"] / k["peak"]
        kernels[name] = k
    roof = None
    if ce["ms"] > 0:
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("fused_softmax")
        except Exception:
            pass
        ach = ce["flops"] / (ce["ms"] * 1e-3) / 1e12
        roof = dict(kernel="fused softmax (ce_fwdg_kernel + ce_bwd_kernel + ce_fwd_kernel: all rs_ce_* calls of a step)",
                    bound="tensor", achieved=ach, peak=tf_peak, unit="TFLOP/s", frac=ach / tf_peak, traffic=traffic,
                    peak_source=peak_src, ms_per_step=round(ce["ms"], 4), share_of_step=ce["ms"] / step_ms,
                    accounting="4 contractions of 2*M*N*128 flops per loss (SURVEY 8d); executed = algorithmic")
    return kernels, roof


def _cuda_timed(n, fn, barrier, dev, world):
    import torch.distributed as dist
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    h0 = time.perf_counter()
    for i in range(n):
        fn(i)
    host_ms = (time.perf_counter() - h0) * 1e3 / max(n, 1)       # time to ENQUEUE a step (launch-bound if ~= ms)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.item(), host_ms


def sharded_parity_check(rs, dev, rank, world):
    """Small sharded-vs-single-process check, run by every `bench.py --gpus N` (the pytest for it needs 2 GPUs and is
    skipped on 1-GPU boxes): the row-sharded step on `world` ranks against the single-process step on the concatenated
    batch -- losses and gradients (eval-mode towers: no dropout; lr 0).  Returns a short verdict string."""
    import torch.distributed as dist
    syn, tr = rs.synthetic, rs.train
    n_items, B, SL = 3000, 64, 50
    n_rows = n_items + 1

    def towers():
        torch.manual_seed(7)
        m = rs.SASRecUserTower(syn.tower_args(num_items=n_items, max_len=SL)).to(dev).eval()
        it = rs.SASRecItemTower(n_items, 128, syn.log_q(n_items)).to(dev)
        lk = syn.pretrained_table(n_items).to(dev)
        it.init_from_pretrained(lk)
        return m, it, lk
    hbs = [syn.make_batch(B, SL, n_items, seed=500 + r) for r in range(world)]
    # single process, concatenated batch
    m0, it0, lk = towers()
    cat = {k: torch.cat([hb[k] for hb in hbs]).to(dev) for k in hbs[0]}
    full = tr.device_index(cat, n_rows, *tr.bucket_of(int((~cat["padding_mask"]).sum()), 4096, 256, 4096))
    opt0 = torch.optim.SGD(list(m0.parameters()) + list(it0.parameters()), lr=0.0)
    want = tr.two_tower_step(m0, it0, full, lk, opt0)
    # sharded
    m1, it1, _ = towers()
    trainer = tr.ShardedDeviceStep(m1, it1)
    mine = {k: v.to(dev) for k, v in hbs[rank].items()}
    t_cap = rs.ops.round_up(int((~mine["padding_mask"]).sum()), 256)
    idx = tr.device_index(mine, n_rows, t_cap, t_cap)
    trainer.calibrate([idx], q=64)
    opt1 = torch.optim.SGD(list(m1.parameters()) + list(it1.parameters()), lr=0.0)
    got = trainer.step(idx, lk, opt1)
    trainer.check()
    worst = 0.0
    for a, b in zip(got, want):
        worst = max(worst, abs(a.item() - b.item()) / max(1.0, abs(b.item())))
    sd = tr.gathered_state_dicts(trainer)          # (also exercises the gather-to-full state_dict hook)
    assert sd["user_tower"]["item_id_emb.weight"].shape == m0.item_id_emb.weight.shape
    assert torch.equal(sd["user_tower"]["item_id_emb.weight"], m0.item_id_emb.weight.detach())
    g_rel = 0.0
    ref = dict(m0.named_parameters())
    for k, p in m1.named_parameters():
        if p.grad is None or k == "item_id_emb.weight":
            continue
        g0 = ref[k].grad
        g_rel = max(g_rel, ((p.grad - g0).norm() / g0.norm().clamp_min(1e-12)).item())
    for shard_p, full_p in ((m1.item_id_emb.weight, m0.item_id_emb.weight), (it1.item_matrix.weight, it0.item_matrix.weight)):
        parts = [torch.empty_like(shard_p.grad) for _ in range(world)]
        dist.all_gather(parts, shard_p.grad.contiguous())
        g1 = rs.sharded.unshard_rows(parts)[:full_p.shape[0]]
        g_rel = max(g_rel, ((g1 - full_p.grad).norm() / full_p.grad.norm().clamp_min(1e-12)).item())
    t = torch.tensor([worst, g_rel], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    worst, g_rel = t.tolist()
    ok = worst < 2e-3 and g_rel < 3e-2
    return f"{'ok' if ok else 'MISMATCH'} (loss rel err {worst:.1e}, gradient rel-Frobenius err {g_rel:.1e}; {world} ranks x B={B} vs one process)"


def run_bucketed(args, rs, dev, rank, world):
    """Device-indexed batches, one captured graph per shape bucket, every timed step on a batch no replay has seen
    before.  N = 1: the plain step.  N > 1: the row-sharded step (train.ShardedDeviceStep) -- every rank runs the same
    bucket per step (the largest row count over the ranks picks it) so that captures and replays stay in lockstep."""
    import torch.distributed as dist
    syn, L, tr = rs.synthetic, rs._lib, rs.train
    lib = L.load()
    B, SL = args.batch, args.seq_len
    n_rows = syn.N_ITEMS + 1
    parity = sharded_parity_check(rs, dev, rank, world) if world > 1 else None
    torch.manual_seed(42)
    model = rs.SASRecUserTower(syn.tower_args(max_len=SL)).to(dev).train()
    item = rs.SASRecItemTower(syn.N_ITEMS, 128, syn.log_q(syn.N_ITEMS)).to(dev)
    lookup = syn.pretrained_table(syn.N_ITEMS).to(dev)
    item.init_from_pretrained(lookup)
    trainer = tr.ShardedDeviceStep(model, item) if world > 1 else None      # re-shards the two item tables in place
    params = list(model.parameters()) + list(item.parameters())
    use_graph = bool(args.cuda_graph) if world == 1 else bool(args.cuda_graph_multi)
    opt = torch.optim.AdamW(params, lr=5e-4, weight_decay=1e-4, fused=True, capturable=use_graph)
    meta_group = dist.new_group() if world > 1 else None      # own communicator for the loader-stage row-count exchange

    def step(b):
        if trainer is not None:
            return trainer.step(b, lookup, opt, amp_dtype=torch.bfloat16)
        return tr.two_tower_step(model, item, b, lookup, opt, loss_scope="all", amp_dtype=torch.bfloat16, columns="unique")

    # N > 1: every rank pads to the LARGEST row count over the ranks, which sits ~1.4 sigma above the mean: finer buckets
    tok_q = args.tok_bucket if world == 1 else min(args.tok_bucket, 1024)
    bs = tr.BucketedStep(step, B, SL, n_rows, dev, use_graph=use_graph, tok_q=tok_q, col_q=args.col_bucket)

    def key_of(t_, u_):
        return bs.bucket(t_, u_ if world == 1 else None)

    # ---- the batches: all distinct (seed 42 + 1000 * rank + i).  value run: warm-up + steps; e2e run: as many again.
    W = max(args.warmup, 3)
    n_val, n_e2e = W + args.steps, (0 if args.minimal else W + args.steps)
    n_all = n_val + n_e2e
    t0 = time.perf_counter()
    seed0 = 42 + 1000 * rank
    host = [tr.FlatBatch(B, SL, pin=True).fill(syn.make_batch(B, SL, syn.N_ITEMS, seed=seed0 + i)) for i in range(n_all)]
    gen_s = time.perf_counter() - t0
    resident = [tr.FlatBatch(B, SL, device=dev).copy_(host[i]) for i in range(n_val)]
    # loader metadata of the resident batches: their shape bucket (device pre-pass, read back before the timed region)
    metas = torch.stack([bs.counts(fb) for fb in resident])
    if world > 1:
        dist.all_reduce(metas, op=dist.ReduceOp.MAX)
    metas = metas.cpu()
    keys = [key_of(int(m[0]), int(m[2])) for m in metas]

    # (the e2e run's batches stay on the host until their step; their buckets are computed here on the host only to
    # capture the graphs up front -- a long run has met each of its few buckets within its first steps -- the timed
    # pipeline itself picks the bucket from the device counts)
    def host_counts(fb):
        valid = ~fb.views["padding_mask"]
        return [int(valid.sum()), 0, int(torch.unique(fb.views["target_ids"][valid]).numel()) if world == 1 else 0]
    e2e_meta = torch.tensor([host_counts(host[n_val + j]) for j in range(n_e2e)], device=dev).reshape(-1, 3)
    if world > 1 and n_e2e:
        dist.all_reduce(e2e_meta, op=dist.ReduceOp.MAX)
    e2e_keys = [key_of(int(m[0]), int(m[2])) for m in e2e_meta.cpu()]

    if trainer is not None:           # per-owner capacities of the two exchanges, from the first resident batches
        cal = []
        for i in range(min(3, n_val)):
            cal.append(tr.device_index(resident[i].views, n_rows, keys[i][0], keys[i][1]))
        caps = trainer.calibrate(cal)
        del cal

    # capture every bucket the run needs.  N = 1: on batches that no timed step uses (extra seeds are drawn until each
    # needed bucket has been met, or 48 tries).  N > 1: in lockstep on the first batch of each bucket.
    need, tries, captured_on_timed = set(keys) | set(e2e_keys), 0, 0
    stage_fb = tr.FlatBatch(B, SL, pin=True)
    while world == 1 and need - set(bs.graphs) and use_graph and tries < 48:
        stage_fb.fill(syn.make_batch(B, SL, syn.N_ITEMS, seed=seed0 + n_all + tries))
        tries += 1
        bs.raw.copy_(stage_fb, non_blocking=False)
        m = bs.counts(bs.raw).cpu()
        k = key_of(int(m[0]), int(m[2]))
        if k in need and k not in bs.graphs:
            bs.ensure(k)
    for i, k in enumerate(keys):
        if use_graph and k not in bs.graphs:
            bs.raw.copy_(resident[i])
            bs.ensure(k)
            captured_on_timed += i >= W
    for j, k in enumerate(e2e_keys):
        if use_graph and k not in bs.graphs:
            bs.raw.copy_(host[n_val + j], non_blocking=False)
            bs.ensure(k)
            captured_on_timed += j >= W
    per_step_launches = max((g_.launches for g_ in bs.graphs.values()), default=0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    last = {}

    def run_resident(i):
        bs.raw.copy_(resident[i])                    # device -> the graphs' static inputs (one copy of the batch)
        last["loss"] = bs.run(keys[i])

    for i in range(W):
        run_resident(i)
    launches0 = lib.rs_launch_count()
    with ClockSampler(dev.index or 0) as clk:
        ms, host_ms = _cuda_timed(args.steps, lambda i: run_resident(W + i), barrier, dev, world)
    launches = (per_step_launches * args.steps) if use_graph else (lib.rs_launch_count() - launches0)
    total, main, cl = [float(x) for x in last["loss"]]
    assert all(map(lambda v: v == v and abs(v) < 1e6, (total, main, cl))), f"non-finite loss {total, main, cl}"
    chk = tr.check_index(bs.index[keys[W + args.steps - 1]])          # overflow / malformed-batch flags of the last step
    if trainer is not None:
        trainer.check()
    value = world * B * args.steps / (ms * 1e-3)
    if args.minimal:
        if rank == 0:
            print(json.dumps(dict(metric=METRIC, value=value, ms_per_step=ms / args.steps, minimal=True)), flush=True)
        return

    # ---- end to end: pinned host batch -> device (one copy, on a copy stream, one step ahead) -> bucket choice from the
    # device counts (N > 1: largest row count over the ranks, exchanged on the copy stream) -> replay -> losses read
    # back, EVERY step, each on a batch nothing has seen before
    copy_stream = torch.cuda.Stream()
    staging = [tr.FlatBatch(B, SL, device=dev) for _ in range(2)]
    meta_dev = [torch.zeros(8, dtype=torch.int32, device=dev) for _ in range(2)]
    meta_host = [torch.zeros(8, dtype=torch.int32).pin_memory() for _ in range(2)]
    staged, consumed = {}, {}
    late_captures = []

    def stage(j):                 # H2D of batch j + its counts, on the copy stream
        slot = j % 2
        if slot in consumed:
            copy_stream.wait_event(consumed[slot])               # the step that read this staging slot has copied it out
        with torch.cuda.stream(copy_stream):
            staging[slot].copy_(host[n_val + j])
            bs.counts(staging[slot], meta_dev[slot])
            if world > 1:
                dist.all_reduce(meta_dev[slot], op=dist.ReduceOp.MAX, group=meta_group)
            meta_host[slot].copy_(meta_dev[slot], non_blocking=True)
            staged[j] = torch.cuda.Event()
            staged[j].record(copy_stream)

    def e2e_step(j):
        if j not in staged:
            stage(j)
        ev = staged.pop(j)
        ev.synchronize()                                         # issued one step ago: complete in steady state
        slot = j % 2
        key = key_of(int(meta_host[slot][0]), int(meta_host[slot][2]))
        torch.cuda.current_stream().wait_event(ev)
        bs.raw.copy_(staging[slot])
        consumed[slot] = torch.cuda.Event()
        consumed[slot].record()
        if use_graph and key not in bs.graphs:                   # a bucket met for the first time: capture (then replay)
            late_captures.append(j)
            bs.ensure(key)
        t, m, c = bs.run(key)
        # D2H read of the step's result, every step: the three losses go to pinned memory behind the replay; the host
        # reads them one step later, after it has enqueued the next step, so that the device never waits for the host
        # (the graphs share one memory pool: the losses must leave it before the next replay anyway)
        loss_pin[slot].copy_(torch.stack([t, m, c]), non_blocking=True)
        loss_ev[slot] = torch.cuda.Event()
        loss_ev[slot].record()
        if j + 1 < n_e2e:
            stage(j + 1)
        prev = 1 - slot
        if loss_ev[prev] is not None:
            loss_ev[prev].synchronize()
            last["host_loss"] = tuple(loss_pin[prev].tolist())
            loss_ev[prev] = None

    loss_pin = [torch.zeros(3).pin_memory() for _ in range(2)]
    loss_ev = [None, None]
    for j in range(W):
        e2e_step(j)
    n_late_warm = len(late_captures)
    ms_e2e, _ = _cuda_timed(args.steps, lambda i: e2e_step(W + i), barrier, dev, world)
    for ev_, pin_ in zip(loss_ev, loss_pin):                     # the last step's losses
        if ev_ is not None:
            ev_.synchronize()
            last["host_loss"] = tuple(pin_.tolist())
    assert all(v == v for v in last["host_loss"]), last["host_loss"]
    if trainer is not None:
        trainer.check()
    h2d_bytes = host[0].nbytes
    e2e = dict(value=world * B * args.steps / (ms_e2e * 1e-3), unit="samples/s", h2d_bytes_per_step=h2d_bytes,
               d2h_bytes_per_step=12 + 32, ms_per_step=ms_e2e / args.steps,
               graph_captures_inside_timed_region=len(late_captures) - n_late_warm)

    # ---- loader-stage cost on the device, per batch: counts pre-pass and the index build (the latter runs INSIDE every
    # timed step); and the same index on the host (torch nonzero / unique / cumsum: what round 1 did outside the timed region)
    def ev_time(fn, n=10):
        fn()
        torch.cuda.synchronize()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b_.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b_) / n
    k0 = keys[0]
    bs.raw.copy_(resident[0])
    idx0 = rs.ops.batch_index_alloc(B, SL, k0[0], k0[1], dev)
    counts_ms = ev_time(lambda: bs.counts(bs.raw))
    build_ms = ev_time(lambda: tr.device_index(bs.raw.views, n_rows, k0[0], k0[1], out=idx0))
    th = time.perf_counter()
    tr.add_host_index({k: v for k, v in host[0].views.items()})
    host_index_ms = (time.perf_counter() - th) * 1e3
    loader = dict(device_counts_ms_per_batch=round(counts_ms, 4), device_index_build_ms_per_batch=round(build_ms, 4),
                  host_index_ms_per_batch_for_comparison=round(host_index_ms, 2),
                  note="index build is inside every timed step (value and e2e); counts pre-pass is inside e2e")

    # ---- per-kernel pass: CUDA events around every C-ABI call of two EAGER steps, on the launching stream
    L.PROFILE = []
    nprof = 2
    for i in range(nprof):
        bs.raw.copy_(resident[i])
        bs._run_eager(keys[i])
    torch.cuda.synchronize()
    prof, L.PROFILE = L.PROFILE, None
    grid_cap = int(bs.index[keys[0]]["fold_inv1"].numel())
    kernels, roof = _kernel_table(prof, nprof, grid_cap, 128, ms / args.steps)

    extra = None if (args.no_extra or world > 1) else bench_extra(rs, dev)
    cpu = gpu_eager = None
    n_graphs = len(bs.graphs)
    if not args.no_cpu_baseline and world == 1:
        cpu, _ = cpu_reference_step_rate(args, 2, 1, syn)
        n_graphs = len(bs.graphs)
        bs.graphs.clear()                                      # free the graphs' pool before the unfused [N, N] baseline
        bs.index.clear()
        resident.clear()
        torch.cuda.empty_cache()
        try:
            gpu_eager = [gpu_eager_step_rate(syn, dev, b_, SL) for b_ in (768, 2048)]
        except Exception as ex:          # noqa: BLE001 -- a baseline leg must not take the bench line down
            gpu_eager = dict(error=repr(ex)[:200])
    if rank != 0:
        return
    buckets = {}
    for k in keys[W:]:
        buckets[f"{k[0]}x{k[1]}"] = buckets.get(f"{k[0]}x{k[1]}", 0) + 1
    if world == 1:
        par = "1 GPU"
        n_cols = chk["columns"]
    else:
        n_cols = world * caps[1]
        par = (f"{world} ranks: item_id_emb + item_matrix row-sharded (owner = row % {world}); U1 rows by a de-duplicated "
               f"equal-split exchange built on the device ({caps[0]} request slots per owner), negatives = distinct targets "
               f"of ALL ranks from the all-reduced target histogram ({caps[1]} column slots per owner, all-gathered "
               f"segments), DuoRec columns all-gathered, other parameters replicated + all-reduced; the id exchange, "
               f"the routing and the index build are inside every timed step")
    line = dict(metric=METRIC, value=value, unit="samples/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16",
                data="synthetic",
                config=dict(workload="two_tower_infonce_train_step (BASELINE configs[1])", batch_per_gpu=B,
                            global_batch=B * world, seq_len=SL, d_model=128, n_items=syn.N_ITEMS, loss_scope="all",
                            loss_rows=chk["tokens"], loss_cols=n_cols,
                            loss_columns="unique" if world == 1 else "unique (box-wide, capacity-padded)", parallelism=par,
                            batches=f"every step (value and e2e) consumes a batch no earlier step has seen "
                                    f"({n_all} distinct synthetic batches per rank, seeds {seed0}..{seed0 + n_all - 1}; "
                                    f"{captured_on_timed} timed batches doubled as graph-capture batches)",
                            index="built on the device inside every step (rs_batch_index_build); host sends the collated "
                                  "[B, L] tensors only",
                            cuda_graph=(f"one captured graph per shape bucket (rows % {tok_q}"
                                        f"{', columns % ' + str(args.col_bucket) if world == 1 else ''}): {n_graphs} "
                                        f"graphs, buckets of the timed steps {buckets}"
                                        if use_graph else "off (eager launches)"),
                            l2="inputs larger than L2 (tables 2x54 MB + >2 GB activations per step), fresh batch per step",
                            full_step="forward of both dropout views, main loss on every valid step + DuoRec, backward, "
                                      "gradient clip, fused AdamW on both towers (dense table gradients) -- nothing the "
                                      "reference step computes and a loss reads is skipped; the second dropout view's "
                                      "last-layer rows that feed no loss are not computed (exact: DESIGN 3.7-3.8, "
                                      "tests/test_gpu_batch_index.py::test_step_on_device_index_matches_step_on_host_index)"),
                e2e=e2e, gpu_launches=int(launches), host_enqueue_ms_per_step=host_ms, clocks=clk.summary(),
                roofline=roof, cpu_baseline=cpu, gpu_eager_baseline=gpu_eager, loader=loader, kernels=kernels, extra=extra,
                loss=dict(total=total, main=main, cl=cl), host_batch_generation_s=round(gen_s, 1))
    if parity is not None:
        line["sharded_parity"] = parity
        # how much of the step is the loss's column growth (inherent to "negatives span the box"): the fused-softmax time
        # scales with the column count; had the columns stayed at this rank's own distinct targets, the step would be
        # shorter by extra_ms -- (ms - extra_ms) / ms is the efficiency this run would show if column growth were free
        ce_ms = sum(kernels[n]["ms_per_step"] for n in ("rs_ce_fwd_grad", "rs_ce_bwd_from_grad") if n in kernels)
        extra_ms = ce_ms * (1.0 - chk["columns"] / max(n_cols, 1))
        step_ms = ms / args.steps
        line["column_growth"] = dict(columns_this_rank_alone=chk["columns"], columns_box_wide_padded=n_cols,
                                     fused_softmax_ms=round(ce_ms, 3), extra_ms_from_growth=round(extra_ms, 3),
                                     efficiency_bound_if_only_columns_grew=round(1.0 / (1.0 + extra_ms / max(step_ms - extra_ms, 1e-9)), 3))
    if extra:
        line["gather_frac"] = extra.get("gather_frac")
        line["scatter_frac"] = extra.get("scatter_frac")
    print(json.dumps(line), flush=True)


def bench_extra(rs, dev):
    """Driver-visible kernel-level numbers for the BASELINE configs that are not the train step, each against its
    roofline (MEASURED_PEAKS.json), CUDA events, L2 flushed between iterations, median of n:
      fronts     config 2's U1 gather-sum / U3 scatter at the FULL [B=8192, L=50] grid and the plain row gathers (the
                 metric's "gather GB/s vs HBM peak"; inside the step these kernels run on the 4x smaller packed grid)
      fm         config 3: DeepFM front, B=65536, F=39, k=16
      retrieval  config 5: 105,542 items, top-12, a 65,536-user sample of the 1.37 M users (ids checked against a
                 chunked fp32 torch matmul+topk on the GPU for the first 4096 users)
      ensemble   N4: union of two top-1000 lists + blend + de-duplicated top-520 for 4096 users x 11 alphas"""
    syn = rs.synthetic
    hbm, _, src = _peaks()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timeit(fn, n=7):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(n):
            flush.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return sorted(ts)[len(ts) // 2]

    def hb(what, ms, nbytes):
        a = nbytes / (ms * 1e-3) / 1e9
        return dict(what=what, ms=round(ms, 4), bound="hbm", achieved=round(a, 1), unit="GB/s", peak=hbm, frac=round(a / hbm, 3),
                    algorithmic_bytes=int(nbytes))
    out = dict(peak_source=src)
    g = torch.Generator().manual_seed(0)
    B, SL, D, NI = 8192, 50, 128, syn.N_ITEMS
    P = B * SL
    b = syn.make_batch(B, SL, NI)
    item_tab = (torch.randn(NI + 1, D, generator=g) * 0.02).to(dev)
    time_tab = (torch.randn(12, D, generator=g) * 0.02).to(dev)
    pos = (torch.randn(SL, D, generator=g) * 0.02).to(dev)
    ids = [b["item_ids"].to(dev), b["time_bucket_ids"].to(dev)]
    gates = torch.tensor([0.7, 0.4], device=dev)
    base = torch.randn(B, SL, D, generator=g).to(dev).bfloat16()
    cot = torch.randn(B, SL, D, generator=g).to(dev).bfloat16()
    fr = {}
    # U1: per position base bf16 + item row fp32 + out bf16 + 2 ids (time / position rows are cache resident)
    fr["seq_front_fwd"] = hb("U1 gather-sum, 409,600 positions, bf16 in/out", timeit(
        lambda: torch.ops.rs.seq_front(base, ids, [item_tab, time_tab], gates, pos, SL, 2)), P * (256 + 512 + 256 + 16))
    fr["gather_rows_item"] = hb("gather 409,600 (Zipf ids) rows of [105543, 128] fp32", timeit(
        lambda: torch.ops.rs.gather_rows(item_tab, ids[0], -1, 0)), P * (512 + 512 + 8))
    big = torch.randn(syn.N_CUSTOMERS + 1, 64, generator=g).to(dev)
    uidx = torch.randint(0, syn.N_CUSTOMERS, (1 << 20,), generator=g).to(dev)
    fr["gather_rows_customers"] = hb("gather 1,048,576 rows of the [1.37M, 64] fp32 customer table", timeit(
        lambda: torch.ops.rs.gather_rows(big, uidx, -1, 0)), (1 << 20) * (256 + 256 + 8))
    word = torch.randn(30522, 768, generator=g).to(dev)
    tok = torch.randint(0, 30522, (512 * 9 * 32,), generator=g).to(dev)
    fr["gather_rows_bert"] = hb("I2: gather 147,456 rows of [30522, 768] fp32", timeit(
        lambda: torch.ops.rs.gather_rows(word, tok, -1, 0)), 147456 * (768 * 4 * 2 + 8))
    # U3: one read of dX (bf16) + ids, RMW of the rows that are hit (unique rows x 512 B x 2) + the dense table's clear
    uniq = int(torch.unique(ids[0]).numel())
    sc_bytes = P * (256 + 8) + 2 * uniq * 512 + (NI + 1) * 512
    rs.ops._sort_cache.clear()
    fr["scatter_sorted"] = hb("U3 deterministic: sort + segment reduce into the dense [105543, 128] grad (incl. its clear)", timeit(
        lambda: (rs.ops._sort_cache.clear(), torch.ops.rs.embedding_dense_bwd(cot.view(-1, D), ids[0].view(-1), NI + 1, 0, -1, True))),
        sc_bytes)
    fr["scatter_atomic"] = hb("U3 red.global.add.v4.f32 variant, same", timeit(
        lambda: torch.ops.rs.embedding_dense_bwd(cot.view(-1, D), ids[0].view(-1), NI + 1, 0, -1, False)), sc_bytes)
    out["fronts"] = fr
    out["gather_frac"] = max(fr["seq_front_fwd"]["frac"], fr["gather_rows_item"]["frac"])
    out["scatter_frac"] = max(fr["scatter_sorted"]["frac"], fr["scatter_atomic"]["frac"])
    del big, word, base, cot
    # ---- config 3
    vocab = syn.criteo_vocab_sizes()
    Bf, F_, k = 65536, len(vocab), 16
    fids = syn.make_fm_batch(Bf, vocab, seed=1).to(dev)
    fm = rs.FM(vocab, k=k, init_std=0.01).to(dev)
    fwd_bytes = Bf * F_ * (k * 4 + 4 + 8) + Bf * 4 + Bf * F_ * k * 2
    d_fm, d_cat = torch.randn(Bf, device=dev), torch.randn(Bf, F_ * k, device=dev).bfloat16()
    bwd_bytes = Bf * F_ * (k * 4 + 8) + Bf * 4 + Bf * F_ * k * 2 + 2 * Bf * F_ * (k + 1) * 4
    out["fm"] = dict(
        fwd=hb(f"fm_fwd B={Bf} F={F_} k={k}: gather + FM + bf16 concat for the DNN", timeit(
            lambda: torch.ops.rs.fm_fwd(fids, fm.offsets, fm.embedding, fm.linear.reshape(-1), True, 2)), fwd_bytes),
        bwd=hb("fm_bwd: atomic scatter into the concatenated table (incl. its clear)", timeit(
            lambda: torch.ops.rs.fm_bwd(fids, fm.offsets, fm.embedding, d_fm, d_cat, True)), bwd_bytes))
    # ---- config 5
    nu, kk = 65536, 12
    I = torch.nn.functional.normalize(torch.randn(NI, D, generator=g), dim=1).to(dev)
    U = torch.nn.functional.normalize(torch.randn(nu, D, generator=g), dim=1).to(dev)
    ms32 = timeit(lambda: rs.ops.retrieve_topk(U, I, kk, tensor_cores=False), n=3)
    ms = timeit(lambda: rs.retrieve_topk(U, I, kk), n=5)              # default path: tcgen05 candidate pass + fp32 re-scoring
    sc, ix = rs.retrieve_topk(U[:4096], I, kk)
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    sc0, ix0 = torch.topk(U[:4096] @ I.T, kk + 1, dim=1)
    torch.backends.cuda.matmul.allow_tf32 = old
    gap = (sc0[:, :-1] - sc0[:, 1:]).min(dim=1).values > 1e-5          # rows whose top-13 are separated beyond fp32 noise
    exact = bool(torch.equal(ix[gap], ix0[gap][:, :kk]))
    flops = 2.0 * nu * NI * D
    _, tfp, _ = _peaks()
    out["retrieval"] = dict(what=f"retrieve_topk {nu} users x {NI} items, k={kk}: bf16 tcgen05 candidate pass + exact fp32 "
                                 f"re-scoring (ids = the fp32 ranking's)", ms=round(ms, 2), bound="tensor",
                            achieved=round(flops / ms / 1e9, 1), unit="TFLOP/s", peak=tfp, frac=round(flops / ms / 1e9 / tfp, 3),
                            users_per_s=round(nu / ms * 1e3),
                            full_1p37M_users_s=round(syn.N_CUSTOMERS / (nu / ms * 1e3), 3),
                            fp32_kernel=dict(ms=round(ms32, 2), tflops_fp32=round(flops / ms32 / 1e9, 1),
                                             users_per_s=round(nu / ms32 * 1e3)),
                            speedup_vs_fp32_kernel=round(ms32 / ms, 1),
                            ids_exact_vs_fp32_torch=f"{exact} on {int(gap.sum())}/4096 rows with a top-13 gap > 1e-5")
    # ---- N4
    ua = torch.nn.functional.normalize(torch.randn(4096, 64, generator=g), dim=1).to(dev)
    ia = torch.nn.functional.normalize(torch.randn(NI, 64, generator=g), dim=1).to(dev)
    alphas = [i / 10 for i in range(11)]
    comb, sa, sb = rs.ensemble.candidate_union(ua, ia, U[:4096], I, 1000)
    ms_m = timeit(lambda: rs.ensemble.merge(comb, sa, sb, alphas, 520, "minmax"), n=3)
    ms_all = timeit(lambda: rs.ensemble.weighted_score_ensemble(ua, ia, U[:4096], I, alphas, 1000, 500), n=3)
    out["ensemble"] = dict(what="4096 users: two top-1000 retrievals over 105,542 items, re-score, min-max blend for 11 alphas, "
                                "de-duplicated top-520 each", merge_kernel_ms=round(ms_m, 3), whole_batch_ms=round(ms_all, 2),
                           users_per_s=round(4096 / ms_all * 1e3))
    return out


def run_ours(args):
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rs = importlib.import_module(PKG)          # raises if librs_twotower.so is missing: no fallback
    if args.loss_scope == "all" and args.columns == "unique" and (world == 1 or args.parallelism == "sharded"):
        run_bucketed(args, rs, dev, rank, world)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    return run_multi(args, rs, dev, rank, world, local)


def run_multi(args, rs, dev, rank, world, local):
    import torch.distributed as dist
    syn, L = rs.synthetic, rs._lib
    lib = L.load()
    torch.manual_seed(42)
    B, SL = args.batch, args.seq_len

    model = rs.SASRecUserTower(syn.tower_args(max_len=SL)).to(dev).train()
    item = rs.SASRecItemTower(syn.N_ITEMS, 128, syn.log_q(syn.N_ITEMS)).to(dev)
    lookup = syn.pretrained_table(syn.N_ITEMS).to(dev)
    item.init_from_pretrained(lookup)
    sharded = world > 1 and args.parallelism == "sharded"
    trainer = rs.train.ShardedTwoTower(model, item) if sharded else None      # re-shards the two item tables in place
    params = list(model.parameters()) + list(item.parameters())
    use_graph = bool(args.cuda_graph) and world == 1
    opt = torch.optim.AdamW(params, lr=5e-4, weight_decay=1e-4, fused=True, capturable=use_graph)

    def sync_grads():
        if world == 1:
            return
        gs = [p.grad for p in params if p.grad is not None]
        flat = torch.cat([g.reshape(-1) for g in gs])
        dist.all_reduce(flat)
        flat.div_(world)
        o = 0
        for g in gs:
            g.copy_(flat[o:o + g.numel()].view_as(g))
            o += g.numel()

    # a small pool of distinct batches (different seeds per rank), pinned on the host + resident copies
    pool = 3
    host = [{k: v.pin_memory() for k, v in rs.train.add_host_index(
        syn.make_batch(B, SL, syn.N_ITEMS, seed=42 + 1000 * rank + i)).items()} for i in range(pool)]
    resident = [rs.train.prepare_batch(hb, dev) for hb in host]
    if sharded:
        resident = [trainer.plan(b, "catalog" if args.columns == "catalog" else "unique") for b in resident]
        plan_keys = [k for k in ("lookup_plan", "col_plan", "col_item_ids", "col_counts", "pos_col") if k in resident[0]]
    h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values())
    n_valid = int(host[0]["valid_index"].numel())
    n_cols = {"unique": int(host[0]["col_item_ids"].numel()), "catalog": syn.N_ITEMS + 1}.get(args.columns, n_valid)
    if args.loss_scope != "all":
        n_cols = B if args.columns != "catalog" else n_cols
    if sharded:
        n_cols = int(resident[0]["col_item_ids"].numel()) if "col_plan" in resident[0] else trainer.cols.n_cols

    def step(b):
        if sharded:
            return trainer.step(b, lookup, opt, amp_dtype=torch.bfloat16)
        return rs.train.two_tower_step(model, item, b, lookup, opt, loss_scope=args.loss_scope,
                                       amp_dtype=torch.bfloat16, grad_hook=sync_grads, columns=args.columns)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    last = {}
    graphs = None
    if use_graph:
        graphs = [rs.train.GraphedStep(step, rb) for rb in resident]
        per_step_launches = max(g_.launches for g_ in graphs)
        run = lambda i: graphs[i % pool].replay()
    else:
        run = lambda i: step(resident[i % pool])
    for i in range(max(args.warmup, 3 * pool) if world > 1 else args.warmup):
        last["loss"] = run(i)
    launches0 = lib.rs_launch_count()
    with ClockSampler(local) as clk:
        ms, host_ms = _cuda_timed(args.steps, lambda i: last.__setitem__("loss", run(i)), barrier, dev, world)
    launches = (per_step_launches * args.steps) if use_graph else (lib.rs_launch_count() - launches0)
    total, main, cl = [float(x) for x in last["loss"]]
    assert all(map(lambda v: v == v and abs(v) < 1e6, (total, main, cl))), f"non-finite loss {total, main, cl}"
    value = world * B * args.steps / (ms * 1e-3)

    copy_stream = torch.cuda.Stream() if use_graph else None
    staged, replayed = {}, {}

    def stage(i):
        g_ = graphs[i % pool]
        if i % pool in replayed:
            copy_stream.wait_event(replayed[i % pool])
        with torch.cuda.stream(copy_stream):
            g_.load(host[i % pool])
            staged[i] = torch.cuda.Event()
            staged[i].record(copy_stream)

    def e2e_step(i):
        if use_graph:
            if i not in staged:
                stage(i)
            torch.cuda.current_stream().wait_event(staged.pop(i))
            t, m, c = graphs[i % pool].replay()
            replayed[i % pool] = torch.cuda.Event()
            replayed[i % pool].record()
            stage(i + 1)
        else:
            b = rs.train.prepare_batch(host[i % pool], dev, non_blocking=True)
            if sharded:
                b.update({k: resident[i % pool][k] for k in plan_keys})
            t, m, c = step(b)
        last["host_loss"] = (t.item(), m.item(), c.item())

    if args.minimal:
        if rank == 0:
            print(json.dumps(dict(metric=METRIC, value=value, ms_per_step=ms / args.steps, minimal=True)), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    for i in range(max(args.warmup, 3 * pool)):
        e2e_step(i)
    ms_e2e, _ = _cuda_timed(args.steps, e2e_step, barrier, dev, world)
    e2e = dict(value=world * B * args.steps / (ms_e2e * 1e-3), unit="samples/s", h2d_bytes_per_step=h2d_bytes,
               d2h_bytes_per_step=12, ms_per_step=ms_e2e / args.steps)

    kernels, roof = {}, None
    nprof = 2
    L.PROFILE = []
    for i in range(nprof):
        step(resident[i % pool])
    torch.cuda.synchronize()
    prof, L.PROFILE = L.PROFILE, None
    if rank == 0:
        P = int(resident[0]["pk_item_ids"].numel()) if "pk_item_ids" in resident[0] else B * SL
        kernels, roof = _kernel_table(prof, nprof, P, 128, ms / args.steps)

    if rank == 0:
        line = dict(metric=METRIC, value=value, unit="samples/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16",
                    data="synthetic",
                    config=dict(workload="two_tower_infonce_train_step (BASELINE configs[1])", batch_per_gpu=B,
                                global_batch=B * world, seq_len=SL, d_model=128, n_items=syn.N_ITEMS,
                                loss_scope=args.loss_scope, loss_rows=n_valid if args.loss_scope == "all" else B,
                                loss_columns=("catalog" if args.columns == "catalog" else "unique (box-wide)") if sharded else args.columns,
                                loss_cols=n_cols,
                                parallelism=("1 GPU" if world == 1 else
                                             f"{world} ranks: item_id_emb + item_matrix row-sharded (owner = row % {world}, "
                                             f"all-to-all lookups), negatives = distinct targets of all ranks with box-wide counts, "
                                             f"DuoRec columns all-gathered, other parameters replicated + all-reduced"
                                             if sharded else f"dp{world} (replicated tables, gradients all-reduced, "
                                                             f"rank-local negatives)"),
                                cuda_graph=("one captured graph per pooled batch (per-batch shapes), replayed" if use_graph
                                            else "off (eager launches)"),
                                l2="inputs larger than L2 (tables 2x54 MB + >2 GB activations per step), 3 rotating batches"),
                    e2e=e2e, gpu_launches=int(launches), host_enqueue_ms_per_step=host_ms, clocks=clk.summary(), roofline=roof,
                    cpu_baseline=None, kernels=kernels, loss=dict(total=total, main=main, cl=cl))
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if os.environ.get("RS_BENCH_FAULT_DUMP"):      # debugging aid: dump every thread's Python stack after N seconds and exit
        import faulthandler
        faulthandler.dump_traceback_later(float(os.environ["RS_BENCH_FAULT_DUMP"]), exit=True)
    if a.impl == "reference":
        run_reference(a)
    elif a.impl == "torch_gpu":
        run_torch_gpu(a)
    else:
        run_ours(a)
