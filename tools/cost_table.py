"""Cost table v2 from committed `ncu --set full --import-source on` captures
(run here, no GPU needed):

    python tools/cost_table.py r02 C4:k_jacobian:gpurun_out/r02a_prof_C4_k_jacobian.ncu-rep:4194304 ...

Each argument is  workload:kernel:report:points_per_launch.  For every capture
the FP64 work the kernel EXECUTES is counted from the source page (thread-level
executed instructions per opcode): DFMA / DMUL / DADD / DSETP / DMNMX occupy one
FP64-pipe slot per thread (2 flop at peak), a warp-level DMMA.8x8x4 is 256 FMA
(512 flop).  Written to profiles/cost_table.json (read by bench.py for the
roofline numerator of the quadrature configs) and, per capture, a summary
profiles/<tag>_<workload>_<kernel>.md with the headline counters.
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
FP64_OPS = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")
RAW = [
    "gpu__time_duration.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
]
BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TIME = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "second": 1.0}


def ncu(rep, page, *extra):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra],
                          capture_output=True, text=True, check=True).stdout


def raw_metrics(rep):
    rows = list(csv.reader(io.StringIO(ncu(rep, "raw"))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = {}
    for h, u, v in zip(hdr, units, vals):
        if h in RAW or h == "Kernel Name":
            try:
                x = float(v.replace(",", ""))
            except ValueError:
                out[h] = v
                continue
            if u in BYTES:
                x *= BYTES[u]
            elif u in TIME:
                x *= TIME[u]
            out[h] = x
            out[h + " [unit]"] = u
    return out


def source_counts(rep):
    rows = list(csv.reader(io.StringIO(ncu(rep, "source", "--print-source", "sass"))))
    hdr, data = rows[1], rows[2:]
    i_src = hdr.index("Source")
    i_warp = hdr.index("Instructions Executed")
    i_thr = hdr.index("Predicated-On Thread Instructions Executed")
    i_smp = hdr.index("Warp Stall Sampling (All Samples)")
    reasons = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    warp, thr, stall = collections.Counter(), collections.Counter(), collections.Counter()
    for r in data:
        toks = r[i_src].split()
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        warp[op] += int(r[i_warp])
        thr[op] += int(r[i_thr])
        for i, h in reasons:
            if r[i]:
                stall[h] += int(r[i])
    return warp, thr, stall, sum(int(r[i_smp]) for r in data)


def main():
    tag = sys.argv[1]
    table_path = os.path.join(OUT, "cost_table.json")
    table = json.load(open(table_path)) if os.path.exists(table_path) else {}
    for spec in sys.argv[2:]:
        workload, kernel, rep, points = spec.split(":")
        points = int(points)
        raw = raw_metrics(rep)
        warp, thr, stall, samples = source_counts(rep)
        fp64_thread = sum(thr[o] for o in FP64_OPS)
        fp64_warp = sum(warp[o] for o in FP64_OPS)
        dmma = warp["DMMA"]
        flop_pt = (2.0 * fp64_thread + 512.0 * dmma) / points
        t = raw["gpu__time_duration.sum"]
        # pipe time: 2 cycles per FP64 warp instruction and 16 per DMMA on each of 148 x 4 sub-partitions
        clock = 1.965e9
        pipe_s = (2.0 * fp64_warp + 16.0 * dmma) / 592 / clock
        entry = {
            "fp64_flop_per_point": flop_pt,
            "fp64_thread_instructions_per_point": fp64_thread / points,
            "dmma_warp_instructions_per_point": dmma / points,
            "points_per_launch": points,
            "seconds_per_launch_under_ncu": t,
            "fp64_datapath_seconds": pipe_s,
            "datapath_utilisation": pipe_s / t,
            "dram_bytes_per_launch": raw["dram__bytes_read.sum"] + raw["dram__bytes_write.sum"],
            "registers_per_thread": raw["launch__registers_per_thread"],
            "source": f"profiles/{tag}_{workload}_{kernel}.md (ncu --set full --import-source on, one launch)",
        }
        table.setdefault(workload, {})[kernel] = entry
        name = str(raw.get("Kernel Name", kernel))
        lines = [f"# {workload} `{kernel}` ({tag}): `{name[:100]}`", "",
                 f"Capture: `{os.path.relpath(rep, ROOT)}`, {points} points in the launch.", "",
                 "| metric | value |", "|---|---|"]
        for k in RAW:
            if k in raw:
                lines.append(f"| `{k}` | {raw[k]:.6g} {'' if k + ' [unit]' not in raw else '(' + raw[k + ' [unit]'] + ' -> SI)'} |")
        lines += ["", "## FP64 work executed (source page, predicated-on thread instructions)", "",
                  "| opcode | warp instructions | thread instructions | per point |", "|---|---|---|---|"]
        for op in FP64_OPS + ("DMMA", "MUFU"):
            if warp[op]:
                lines.append(f"| {op} | {warp[op]} | {thr[op]} | {thr[op] / points:.2f} |")
        lines += ["",
                  f"* FP64-pipe instruction slots per point: **{fp64_thread / points:.1f}** "
                  f"(+ {dmma / points * 32:.2f} DMMA per 32 points) => cost table v2: **{flop_pt:.0f} flop/point**",
                  f"* FP64 datapath time of that work at 1.965 GHz: {pipe_s * 1e3:.3f} ms of {t * 1e3:.3f} ms "
                  f"=> datapath utilisation **{pipe_s / t:.3f}** (this launch, under ncu)",
                  f"* DRAM traffic {entry['dram_bytes_per_launch'] / 1e6:.1f} MB vs {24 * points / 1e6:.1f} MB algorithmic "
                  f"(24 B/point)", "", "## Warp stalls (samples)", "", "| reason | share |", "|---|---|"]
        for h, v in stall.most_common(7):
            lines.append(f"| {h} | {100 * v / max(samples, 1):.1f} % |")
        path = os.path.join(OUT, f"{tag}_{workload}_{kernel}.md")
        open(path, "w").write("\n".join(lines) + "\n")
        print(f"{workload} {kernel}: {flop_pt:.0f} flop/pt, {fp64_thread / points:.1f} FP64 instr/pt, "
              f"utilisation {pipe_s / t:.3f}, {t * 1e3:.3f} ms, regs {raw['launch__registers_per_thread']:.0f}")
    json.dump(table, open(table_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
