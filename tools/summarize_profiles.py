"""Turn gpurun_out/launches.csv (ncu --metrics gpu__time_duration.sum) and an ncu report's raw page into the
small tracked summaries under profiles/."""
import collections
import csv
import json
import subprocess
import sys


def launches(path, steps, out_md):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    if steps < 0:
        # keep only the last -steps train steps: a step starts with the dropout-epoch kernel (rs_rng_advance)
        starts = [i for i, r in enumerate(rows) if "rng_advance" in r["Kernel Name"]]
        steps = -steps
        rows = rows[starts[-steps]:]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
        a = agg[r["Kernel Name"]]
        a[0] += 1
        a[1] += ms
    tot = sum(v[1] for v in agg.values())
    ours = sum(v[1] for k, v in agg.items() if "rs::" in k)
    with open(out_md, "w") as f:
        f.write(f"# Kernel launch list (ncu gpu__time_duration.sum, cold-cache, serialised) -- {len(rows)} launches, "
                f"{steps} steps\n\n")
        f.write(f"total {tot / steps:.2f} ms/step; kernels of this library (rs::*) {ours / steps:.2f} ms/step "
                f"({100 * ours / tot:.1f} %)\n\n| ms/step | share | launches/step | kernel |\n|---:|---:|---:|---|\n")
        for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
            f.write(f"| {ms / steps:.3f} | {100 * ms / tot:.1f}% | {c / steps:.1f} | `{k[:120]}` |\n")


def raw(rep, out_json):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size"]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        out.append({w: (r[idx[w]] + " " + units[idx[w]]).strip() for w in want if w in idx})
    json.dump(out, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], int(sys.argv[3]), sys.argv[4])
    else:
        raw(sys.argv[2], sys.argv[3])
