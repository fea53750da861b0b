"""Round 2: turn what the final GPU passes brought back in gpurun_out/ into the
tracked summaries under profiles/ (run here, no GPU needed):

    python tools/summarize_r02.py r02z

Inputs (gpurun_out/, written by tools/gpu_r02_final.sh and tools/gpu_r02_multi.sh):
    <tag>_bench_C2.json (default line incl. `workloads`), <tag>_bench_C3|C4|C5.json,
    <tag>_bench_reference.json, r02_bench_C2_n<N>.json (multi-GPU),
    <tag>_launches_C2.csv, <tag>_launches_C3.csv,
    <tag>_prof_<workload>_<kernel>.ncu-rep
Outputs (profiles/): r02_bench_*.json, r02_launches_c2|c3.csv/.md, r02_<W>_<kernel>.md +
    cost_table.json (tools/cost_table.py), r02_*_stalls.md (tools/stall_summary.py), r02_summary.md
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
POINTS = {"C2": 64_000_000, "C3": 4_096_000, "C4": 64 * 65536, "C5": 25000 * 256}
CAPTURES = [("C2", "k_jacobian"), ("C2", "k_chi2"), ("C3", "k_jacobian"), ("C3", "k_chi2"),
            ("C3", "k_solve_small"), ("C4", "k_jacobian"), ("C4", "k_chi2"), ("C4", "k_omega"),
            ("C5", "k_jacobian"), ("C5", "k_chi2")]


def load(path):
    try:
        return json.loads(open(path).read().strip().splitlines()[-1])
    except (OSError, ValueError, IndexError):
        return None


def launch_list(tag, workload, command):
    src = os.path.join(SRC, f"{tag}_launches_{workload}.csv")
    if not os.path.exists(src):
        return None
    low = workload.lower()
    shutil.copy(src, os.path.join(OUT, f"r02_launches_{low}.csv"))
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = r[ki].split("(")[0][:80]
        agg[name][0] += 1
        agg[name][1] += v
    micro = {k: v for k, v in agg.items() if "k_dfma" in k or "k_dmma" in k}
    solver = {k: v for k, v in agg.items() if k not in micro}
    total = sum(v for _, v in solver.values())
    lines = [f"# ncu launch list on BASELINE config {workload[1]} (r02), `{command}`", "",
             "Per-launch times are cold-cache and serialised (ncu flushes the caches before every "
             "launch): compare SHARES and COUNTS, not absolutes.  Not listed: the FP64 peak "
             "micro-benchmarks bench.py runs before the timed region.", "",
             "| kernel | launches | mean us | share of solver kernels |", "|---|---|---|---|"]
    table = {}
    for k, (n, v) in sorted(solver.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k}` | {n} | {v / 1e3 / n:.1f} | {100 * v / total:.1f} % |")
        table[k] = dict(launches=n, mean_us=v / 1e3 / n, share=v / total)
    open(os.path.join(OUT, f"r02_launches_{low}.md"), "w").write("\n".join(lines) + "\n")
    return table


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02z"
    os.makedirs(OUT, exist_ok=True)
    # ---- cost table v2 + one summary per capture
    args = []
    for w, k in CAPTURES:
        rep = os.path.join(SRC, f"{tag}_prof_{w}_{k}.ncu-rep")
        if os.path.exists(rep) and k != "k_solve_small":
            args.append(f"{w}:{k}:{rep}:{POINTS[w]}")
    if args:
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "cost_table.py"), "r02", *args], check=True)
    for w, k in (("C2", "k_jacobian"), ("C2", "k_chi2"), ("C3", "k_jacobian"), ("C3", "k_solve_small")):
        rep = os.path.join(SRC, f"{tag}_prof_{w}_{k}.ncu-rep")
        if os.path.exists(rep):
            subprocess.run([sys.executable, os.path.join(ROOT, "tools", "stall_summary.py"), rep,
                            os.path.join(OUT, f"r02_{w}_{k}_stalls.md")], check=False,
                           stdout=subprocess.DEVNULL)
    # ---- bench lines
    lines = {}
    for name, src in [("c2", f"{tag}_bench_C2.json"), ("c3", f"{tag}_bench_C3.json"),
                      ("c4", f"{tag}_bench_C4.json"), ("c5", f"{tag}_bench_C5.json"),
                      ("reference", f"{tag}_bench_reference.json"),
                      ("c2_n2", "r02_bench_C2_n2.json"), ("c2_n4", "r02_bench_C2_n4.json"),
                      ("c2_n8", "r02_bench_C2_n8.json")]:
        d = load(os.path.join(SRC, src))
        if d is not None:
            lines[name] = d
            json.dump(d, open(os.path.join(OUT, f"r02_bench_{name}.json"), "w"), indent=1)
    l2 = launch_list(tag, "C2", "bench.py --workload C2 --steps 3 --warmup 1 --repeats 1")
    l3 = launch_list(tag, "C3", "bench.py --workload C3 --steps 10 --warmup 1 --repeats 1")
    # ---- summary
    out = ["# Profile summary, round 2 (one B200 unless stated; FP64)", "",
           "Sources: `profiles/r02_bench_*.json` (bench.py lines of the final build), "
           "`profiles/r02_<config>_<kernel>.md` (one `ncu --set full --clock-control none "
           "--import-source on` capture each, counted by `tools/cost_table.py`), "
           "`profiles/r02_launches_c2|c3.md`, `profiles/r02_k1_experiments.md`.", ""]
    if "c2" in lines:
        b = lines["c2"]
        r = b["roofline"]
        per = {k: b["phase_seconds"][k] / max(b["phase_calls"][k], 1) for k in b["phase_seconds"]}
        out += ["## Config 2 (headline): 64 data sets x 1M points, 20 shared parameters", "",
                f"* {b['ms_per_step']:.3f} ms per LM iteration, {b['value']:.4g} {b['unit']}, "
                f"{b['gpu_launches']} kernel launches per {b['steps']} iterations; clocks {b['clocks']}",
                f"* K1 `{r['kernel']}` {r['seconds_per_launch'] * 1e3:.3f} ms per launch (CUDA events inside the "
                f"timed fit): {r['achieved']:.2f} of {r['peak']:.2f} TFLOP/s = **{r['frac']:.3f}** "
                f"({r['flops_per_point']:.0f} flop/point; frac_v1 {r['frac_v1']:.3f}); DRAM traffic "
                f"{(r['traffic'] or 0) / 1e9:.3f} GB vs {r['algorithmic_bytes'] / 1e9:.3f} GB algorithmic",
                f"* K2 chi2 sweep {per['chi2'] * 1e3:.3f} ms per call ({b['phase_calls']['chi2']} calls), "
                f"linear algebra {per['linalg'] * 1e6:.0f} us per trial",
                f"* e2e {b['e2e']['value']:.4g} entries/s ({b['e2e']['seconds'] * 1e3:.1f} ms for upload + "
                f"{b['e2e']['steps']} iterations + read-back; {b['e2e']['h2d_bytes_per_step']} B H2D per step)",
                f"* cpu_baseline: {b['cpu_baseline']}", ""]
        k1_share = None
        if l2:
            for k, v in l2.items():
                if "k_jacobian" in k:
                    k1_share = v["share"]
            jac = b["phase_seconds"]["jacobian"]
            tot = sum(b["phase_seconds"][k] for k in ("jacobian", "chi2", "linalg", "omega"))
            out += [f"* K1 share of the step: {100 * jac / tot:.1f} % of the phase times (bench, CUDA events) vs "
                    f"{100 * (k1_share or 0):.1f} % of the ncu launch list (`r02_launches_c2.md`)", ""]
        out += ["## The other BASELINE shapes (full size, same bench line, `workloads`)", "",
                "| config | ms / iteration | K1 us | chi2 us | omega us | linear algebra us | launches | K1 frac (executed work) | frac_v1 |",
                "|---|---|---|---|---|---|---|---|---|"]
        for name, v in (b.get("workloads") or {}).items():
            s = v["seconds_per_call"]
            out.append(f"| {name} | {v['ms_per_step']:.4f} | {s['k1_jacobian_kernel'] * 1e6:.1f} | "
                       f"{s['k2_chi2_phase'] * 1e6:.1f} | {s['k3_omega_phase'] * 1e6:.1f} | "
                       f"{s['linear_algebra_phase'] * 1e6:.1f} | {v['gpu_launches']} / {v['steps']} it. | "
                       f"{v['roofline']['frac']:.3f} | {v['roofline']['frac_v1']:.3f} |")
        out.append("")
    for name in ("c3", "c4", "c5"):
        if name in lines:
            d = lines[name]
            out.append(f"* `bench.py --workload {name.upper()}` alone: {d['ms_per_step']:.4f} ms per iteration, "
                       f"K1 {d['roofline']['seconds_per_launch'] * 1e6:.1f} us, frac {d['roofline']['frac']:.3f}, "
                       f"e2e {d['e2e']['value']:.4g} entries/s" if d.get("e2e") else
                       f"* `bench.py --workload {name.upper()}` alone: {d['ms_per_step']:.4f} ms per iteration")
    out.append("")
    if "reference" in lines:
        d = lines["reference"]
        out += ["## Reference arm (`bench.py --impl reference`, unmodified reference on the host cores)", "",
                f"* {d['config']['workload']}",
                f"* {d['value']:.4g} {d['unit']}, {d['ms_per_step'] / 1e3:.1f} s per iteration, steps run: {d['steps']}, "
                f"same config as the GPU arm: {d.get('same_config_as_gpu_arm')}", ""]
    multi = [(n, lines[f"c2_n{n}"]) for n in (2, 4, 8) if f"c2_n{n}" in lines]
    if multi and "c2" in lines:
        base = lines["c2"]
        out += ["## Multi-GPU (config 2 per GPU, weak scaling; builder-run under `gpurun --gpus N`)", "",
                "| GPUs | entries/s | ms / iteration | device-timed efficiency | e2e entries/s | parity record |", "|---|---|---|---|---|---|"]
        for n, d in multi:
            eff = d["value"] / (n * base["value"])
            par = d.get("parity") or {}
            out.append(f"| {n} | {d['value']:.4g} | {d['ms_per_step']:.3f} | {100 * eff:.1f} % | "
                       f"{d['e2e']['value']:.4g} | ok = {par.get('ok')} ({len(par.get('cases', []))} cases, "
                       f"worst parameter diff {max([c['parameter_max_rel_diff'] for c in par.get('cases', [])] or [0]):.1e}) |")
        out.append("")
    if l3:
        out += ["## Config 3 launch list (device-driven loop)", ""]
        for k, v in list(l3.items())[:6]:
            out.append(f"* `{k}`: {v['launches']} launches, mean {v['mean_us']:.1f} us, {100 * v['share']:.1f} %")
        out.append("")
    open(os.path.join(OUT, "r02_summary.md"), "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main()
