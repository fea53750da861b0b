"""Turn the ncu captures brought back in gpurun_out/ into the tracked summaries
under profiles/ (run here, no GPU needed):

    python tools/summarize_profiles.py r01

Inputs (gpurun_out/): launches_<tag>.csv  -- `ncu --metrics gpu__time_duration.sum` launch list
                      prof_<tag>_k_jacobian.ncu-rep, prof_<tag>_k_chi2.ncu-rep -- `ncu --set full`
Outputs (profiles/):  <tag>_launches_c2.csv/.md, <tag>_<kernel>_raw.csv, <tag>_summary.md/.json
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")

METRICS = [
    "gpu__time_duration.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread",
    "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def launches(tag):
    src = os.path.join(SRC, f"launches_{tag}.csv")
    if not os.path.exists(src):
        return None
    shutil.copy(src, os.path.join(OUT, f"{tag}_launches_c2.csv"))
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = r[ki].split("(")[0][:70]
        agg[name][0] += 1
        agg[name][1] += v
    # k_dfma / k_dmma are bench.py's FP64 peak micro-benchmarks (csrc/microbench.cu),
    # launched before the timed region: not part of the step
    micro = {k: v for k, v in agg.items() if "k_dfma" in k or "k_dmma" in k}
    solver = {k: v for k, v in agg.items() if k not in micro}
    total = sum(v for _, v in solver.values())
    lines = [f"# ncu launch list, `bench.py --steps 2 --warmup 1 --repeats 1` on C2 ({tag})", "",
             "Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.",
             "Not listed: the FP64 peak micro-benchmarks bench.py runs before the timed region ("
             + ", ".join(f"`{k.split('::')[-1]}` x {n}" for k, (n, _) in micro.items()) + ").", "",
             "| kernel | launches | total us | share of solver kernels |", "|---|---|---|---|"]
    table = {}
    for k, (n, v) in sorted(solver.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k}` | {n} | {v / 1e3:.1f} | {100 * v / total:.1f} % |")
        table[k] = dict(launches=n, total_us=v / 1e3, share=v / total)
    open(os.path.join(OUT, f"{tag}_launches_c2.md"), "w").write("\n".join(lines) + "\n")
    return table


def kernel(tag, name):
    rep = os.path.join(SRC, f"prof_{tag}_{name}.ncu-rep")
    if not os.path.exists(rep):
        return None
    raw = os.path.join(OUT, f"{tag}_{name}_c2_raw.csv")
    with open(raw, "w") as f:
        subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=f, stderr=subprocess.DEVNULL, check=True)
    rows = list(csv.reader(open(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = {}
    for m in METRICS:
        if m in hdr:
            i = hdr.index(m)
            out[m] = dict(value=vals[i], unit=units[i])
    stalls = {h.split("issue_stalled_")[1].split("_per_")[0]: float(vals[i] or 0)
              for i, h in enumerate(hdr) if "smsp__average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio")}
    out["stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:6])
    # DRAM traffic of this launch in bytes (bench.py reports it as roofline.traffic)
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    traffic = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        if m in out:
            traffic += float(out[m]["value"]) * scale.get(out[m]["unit"], 1.0)
    out["dram_bytes_per_launch"] = traffic
    return out


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(OUT, exist_ok=True)
    summary = dict(tag=tag, launches=launches(tag),
                   k_jacobian=kernel(tag, "k_jacobian"), k_chi2=kernel(tag, "k_chi2"))
    bench = os.path.join(SRC, f"bench_{tag}.json")
    if os.path.exists(bench):
        summary["bench"] = json.load(open(bench))
        shutil.copy(bench, os.path.join(OUT, f"{tag}_bench_c2.json"))
    json.dump(summary, open(os.path.join(OUT, f"{tag}_summary.json"), "w"), indent=1)
    lines = [f"# Profile summary {tag} (B200, BASELINE config 2: 64 x 1M points, 20 shared parameters)", ""]
    for kname in ("k_jacobian", "k_chi2"):
        k = summary[kname]
        if not k:
            continue
        lines += [f"## {kname} (`ncu --set full --clock-control none`, one launch)", "",
                  "| metric | value |", "|---|---|"]
        for m in METRICS:
            if m in k:
                lines.append(f"| `{m}` | {k[m]['value']} {k[m]['unit']} |")
        lines.append("| top stalls (warps per issue) | " + ", ".join(f"{a} {b:.2f}" for a, b in k["stalls_per_issue"].items()) + " |")
        lines.append("")
    if summary.get("bench"):
        b = summary["bench"]
        r = b["roofline"]
        lines += ["## bench.py line of the same build", "",
                  f"* value {b['value']:.4g} {b['unit']}, {b['ms_per_step']:.3f} ms per LM iteration, "
                  f"{b['gpu_launches']} kernel launches in the timed fit",
                  f"* k_jacobian {r['seconds_per_launch'] * 1e3:.3f} ms per launch (CUDA events, warm), "
                  f"{r['achieved']:.2f} of {r['peak']:.2f} TFLOP/s = {r['frac']:.3f}; "
                  f"HBM {r['hbm']['achieved']:.0f} of {r['hbm']['peak']:.0f} GB/s = {r['hbm']['frac']:.3f}",
                  f"* phases [s]: {b['phase_seconds']}", ""]
    open(os.path.join(OUT, f"{tag}_summary.md"), "w").write("\n".join(lines))
    print("\n".join(lines))


if __name__ == "__main__":
    main()
