#!/usr/bin/env python3
"""Turn ncu captures from gpurun_out/ into the small text/CSV summaries committed under
profiles/ (the .ncu-rep files themselves are large and stay in gpurun_out/).

usage: python profiles/summarize.py <tag> <kernel.ncu-rep> [launches.csv]
"""
import csv
import io
import json
import os
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2:]
    return hdr, units, vals


def main():
    tag, rep = sys.argv[1], sys.argv[2]
    here = os.path.dirname(os.path.abspath(__file__))
    hdr, units, vals = raw(rep)
    lines = [f"# ncu --set full --clock-control none summary of {os.path.basename(rep)} (captured under gpurun; cold-cache, serialised)"]
    name_i = hdr.index("Kernel Name")
    summary = {}
    for v in vals:
        lines.append(f"kernel: {v[name_i]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                lines.append(f"  {k:75s} {v[i]:>16s} {units[i]}")
                summary[k] = (v[i], units[i])
        lines.append("  -- warp stall reasons (warps stalled per issue-active cycle) --")
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                lines.append(f"  {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:30s} {v[i]}")
    open(os.path.join(here, f"{tag}_summary.txt"), "w").write("\n".join(lines) + "\n")
    if len(sys.argv) > 3:
        rows = [r for r in csv.reader(open(sys.argv[3])) if r and not r[0].startswith("==")]
        h = rows[0]
        ni, mi, vi, ui = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
        per = {}
        order = []
        for r in rows[1:]:
            if r[mi] != "gpu__time_duration.sum":
                continue
            nm = r[ni].split("(")[0]
            val = float(r[vi].replace(",", ""))
            if r[ui] == "ns":
                val /= 1000.0
            elif r[ui] == "ms":
                val *= 1000.0
            per.setdefault(nm, []).append(val)
            order.append((nm, val))
        tot = sum(sum(v) for v in per.values())
        with open(os.path.join(here, f"{tag}_launches.csv"), "w") as f:
            f.write("# ncu --metrics gpu__time_duration.sum --clock-control none launch list (serialised, cold cache): compare SHARES\n")
            f.write("kernel,launches,mean_us,total_us,share_of_step\n")
            for nm, v in per.items():
                f.write(f"{nm},{len(v)},{sum(v)/len(v):.2f},{sum(v):.2f},{sum(v)/tot:.4f}\n")
            f.write("# per-launch sequence\nkernel,us\n")
            for nm, val in order:
                f.write(f"{nm},{val:.2f}\n")
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()
