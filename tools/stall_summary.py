"""Source-level view of one `ncu --set full --import-source on` capture: where
the warps of a kernel wait, by opcode class and by instruction (run here, no
GPU needed):

    python tools/stall_summary.py gpurun_out/prof_r01_k_jacobian.ncu-rep profiles/r01_k_jacobian_stalls.md
"""
import collections
import csv
import io
import subprocess
import sys


def main(rep: str, out: str) -> None:
    text = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                          capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    kernel = rows[0][1]
    hdr, data = rows[1], rows[2:]
    i_src, i_smp, i_exe = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), \
        hdr.index("Instructions Executed")
    reasons = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    total = sum(int(r[i_smp]) for r in data)
    by_op_samples, by_op_exec = collections.Counter(), collections.Counter()
    by_reason = collections.Counter()
    for r in data:
        toks = r[i_src].split()
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        by_op_samples[op] += int(r[i_smp])
        by_op_exec[op] += int(r[i_exe])
        for i, h in reasons:
            if r[i]:
                by_reason[h] += int(r[i])
    fp64 = sum(by_op_exec[o] for o in ("DFMA", "DMUL", "DADD"))
    dmma = by_op_exec["DMMA"]
    lines = [f"# Warp-stall samples by instruction: `{kernel[:90]}`", "",
             f"Source: `{rep}` (`ncu --set full --import-source on`, one launch); {total} samples, "
             f"{len(data)} SASS instructions.", "",
             f"Executed FP64 work: {fp64 / 1e6:.0f} M DFMA/DMUL/DADD warp instructions (2 issue cycles each "
             f"per SM sub-partition) + {dmma / 1e6:.0f} M DMMA.8x8x4 (16 cycles each) = "
             f"{(2 * fp64 + 16 * dmma) / 592 / 1.965e9 * 1e3:.2f} ms of FP64 datapath time on 148 x 4 "
             "sub-partitions at 1.965 GHz.", "",
             "## By stall reason", "", "| reason | samples | share |", "|---|---|---|"]
    for h, v in by_reason.most_common(8):
        lines.append(f"| {h} | {v} | {100 * v / total:.1f} % |")
    lines += ["", "## By opcode", "", "| opcode | samples | share | executed (warp instructions) |", "|---|---|---|---|"]
    for op, v in by_op_samples.most_common(14):
        lines.append(f"| {op} | {v} | {100 * v / total:.1f} % | {by_op_exec[op]} |")
    lines += ["", "## Top instructions", "", "| samples | share | executed | instruction | main reasons |", "|---|---|---|---|---|"]
    for r in sorted(data, key=lambda r: -int(r[i_smp]))[:16]:
        st = sorted(((h, int(r[i])) for i, h in reasons if r[i] and int(r[i]) > 0), key=lambda kv: -kv[1])[:2]
        lines.append(f"| {r[i_smp]} | {100 * int(r[i_smp]) / total:.1f} % | {r[i_exe]} | `{r[i_src].strip()[:60]}` | "
                     + ", ".join(f"{h} {v}" for h, v in st) + " |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
