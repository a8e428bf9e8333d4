#!/usr/bin/env python
"""Turn gpurun_out/ ncu artefacts into the small text summaries committed under profiles/.

    python profiles/summarize.py <tag> [--launches gpurun_out/launches.csv] [--rep NAME=gpurun_out/x.ncu-rep ...]

Writes profiles/<tag>_launches.md (per-kernel launch counts, device time and SHARE of the command) and
profiles/<tag>_<NAME>.md (selected raw metrics per captured launch + stall breakdown of the source page).
"""
import argparse
import collections
import csv
import io
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))

RAW_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
]
STALLS = ["stall_long_sb", "stall_short_sb", "stall_wait", "stall_mio", "stall_lg", "stall_math", "stall_barrier",
          "stall_branch_resolving", "stall_no_inst", "stall_not_selected", "stall_selected", "stall_dispatch",
          "stall_drain", "stall_membar", "stall_sleep", "stall_tex", "stall_misc"]


def launches_summary(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")))
    unit = rows[1][ui]
    total = sum(sum(v) for v in agg.values())
    out = [f"source: {path} (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised "
           "launches: compare SHARES, not absolutes)", "",
           f"| kernel | launches | total {unit} | avg {unit} | share |", "|---|---:|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"| `{k[:110]}` | {len(v)} | {sum(v):.0f} | {sum(v) / len(v):.0f} | {sum(v) / total:.3f} |")
    return "\n".join(out) + "\n"


def ncu_csv(rep, page):
    p = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True)
    return list(csv.reader(io.StringIO(p.stdout)))


def rep_summary(rep):
    rows = ncu_csv(rep, "raw")
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = [f"source: {rep} (ncu --set full --clock-control none --import-source on)", ""]
    ki = hdr.index("Kernel Name")
    out.append("kernels captured: " + "; ".join(sorted({r[ki][:90] for r in data})))
    out += ["", "| metric | unit | per captured launch |", "|---|---|---|"]
    for m in RAW_METRICS:
        if m in hdr:
            i = hdr.index(m)
            out.append(f"| {m} | {units[i]} | {', '.join(r[i] for r in data)} |")
    src = ncu_csv(rep, "source")
    if len(src) > 2:
        h = src[1]
        block = []
        for r in src[2:]:
            if len(r) < len(h):
                break
            block.append(r)
        si = h.index("# Samples")
        total = sum(int(r[si]) for r in block) or 1
        out += ["", f"warp-stall samples of the first captured launch ({total} samples, {len(block)} SASS instructions):", "",
                "| stall reason | share |", "|---|---:|"]
        for s in STALLS:
            if s in h:
                v = sum(int(r[h.index(s)]) for r in block) / total
                if v >= 0.005:
                    out.append(f"| {s} | {v:.3f} |")
        ops, opn = collections.Counter(), collections.Counter()
        ii, srci = h.index("Instructions Executed"), h.index("Source")
        for r in block:
            tok = r[srci].split()
            op = (tok[1] if tok[0].startswith("@") else tok[0]).split(".")[0]
            ops[op] += int(r[si])
            opn[op] += int(r[ii])
        out += ["", "| SASS opcode | stall-sample share | warp instructions executed |", "|---|---:|---:|"]
        for op, v in ops.most_common(14):
            out.append(f"| {op} | {v / total:.3f} | {opn[op]} |")
        out.append(f"| (all) | 1.000 | {sum(opn.values())} |")
    return "\n".join(out) + "\n"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--launches")
    ap.add_argument("--rep", action="append", default=[])
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    if a.launches:
        with open(os.path.join(HERE, f"{a.tag}_launches.md"), "w") as f:
            f.write(f"# {a.tag}: launch list\n\n{a.note}\n\n" + launches_summary(a.launches))
    for spec in a.rep:
        name, path = spec.split("=", 1)
        with open(os.path.join(HERE, f"{a.tag}_{name}.md"), "w") as f:
            f.write(f"# {a.tag}: {name}\n\n{a.note}\n\n" + rep_summary(path))


if __name__ == "__main__":
    sys.exit(main())
