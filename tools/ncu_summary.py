#!/usr/bin/env python
"""Summarise Nsight Compute artefacts into the markdown tables kept under profiles/.

    python tools/ncu_summary.py raw  gpurun_out/prof.ncu-rep          # one row per captured launch (--set full capture)
    python tools/ncu_summary.py list gpurun_out/launches.csv          # per-kernel totals of a gpu__time_duration launch list
    python tools/ncu_summary.py hot  gpurun_out/prof.ncu-rep [N]      # the N most-sampled SASS instructions with stall reasons
Runs here (no GPU needed): it only reads reports brought back by gpurun."""
import collections
import csv
import io
import subprocess
import sys

RAW = [
    ("time us", "gpu__time_duration.sum", 1.0),
    ("dram rd MB", "dram__bytes_read.sum", 1.0),
    ("dram wr MB", "dram__bytes_write.sum", 1.0),
    ("dram %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("l1tex lsu wavefronts %", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("smem wavefronts M", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", 1e-6),
    ("smem bank conflicts M", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", 1e-6),
    ("fp64 pipe %", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", 1.0),
    ("alu pipe %", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", 1.0),
    ("fma pipe %", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", 1.0),
    ("issue active %", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1.0),
    ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1.0),
    ("warp inst M", "smsp__inst_executed.sum", 1e-6),
    ("regs", "launch__registers_per_thread", 1.0),
    ("dyn smem KB", "launch__shared_mem_per_block_dynamic", 1.0),
    ("stall long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", 1.0),
    ("stall short_sb", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", 1.0),
    ("stall wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", 1.0),
    ("stall barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", 1.0),
    ("stall math", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", 1.0),
]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def short(name):
    return name.split("(")[0].replace("void ", "").replace("mmd::", "")


def raw(rep):
    rows = page(rep, "raw")
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("| kernel | " + " | ".join(k for k, m, s in RAW if m in idx) + " |")
    print("|---|" + "---|" * sum(1 for k, m, s in RAW if m in idx))
    for r in rows[2:]:
        cells = []
        for k, m, sc in RAW:
            if m not in idx:
                continue
            try:
                v = float(r[idx[m]].replace(",", "")) * sc
                u = units[idx[m]]
                if k == "time us" and u == "ms":
                    v *= 1000
                if k.endswith("MB") and u.startswith("G"):
                    v *= 1000
                if k.endswith("MB") and u.startswith("K"):
                    v /= 1000
                cells.append(f"{v:.1f}" if abs(v) < 1000 else f"{v:.0f}")
            except ValueError:
                cells.append(r[idx[m]])
        print(f"| `{short(r[idx['Kernel Name']])}` | " + " | ".join(cells) + " |")


def launch_list(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, data = r, rows[i + 1:]
            break
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        if len(r) <= vi:
            continue
        t = float(r[vi].replace(",", ""))
        t = {"ns": t / 1000, "us": t, "ms": t * 1000, "nsecond": t / 1000, "usecond": t, "msecond": t * 1000}.get(r[ui], t / 1000)
        a = agg[short(r[ki])]
        a[0] += 1
        a[1] += t
    tot = sum(v[1] for v in agg.values())
    print(f"{sum(v[0] for v in agg.values())} launches, {tot / 1000:.2f} ms of kernel time\n")
    print("| share | launches | avg us | kernel |\n|---|---|---|---|")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {100 * t / tot:.2f} % | {c} | {t / c:.1f} | `{n}` |")


def hot(rep, n=20):
    rows = page(rep, "source")
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) > idx["# Samples"] and (r[idx["# Samples"]] or "0").isdigit()]
    tot = sum(int(r[idx["# Samples"]] or 0) for r in data)
    print(f"{len(data)} SASS instructions, {tot} samples\n")
    print("| samples | executed M | instruction | dominant stalls |\n|---|---|---|---|")
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]] or 0))[:n]:
        st = sorted(((int(r[idx[k]] or 0), k[6:]) for k in stalls), reverse=True)[:2]
        print(f"| {100 * int(r[idx['# Samples']]) / tot:.1f} % | {int(r[idx['Instructions Executed']] or 0) / 1e6:.1f} | "
              f"`{r[idx['Source']].strip()[:70]}` | {', '.join(f'{k} {v}' for v, k in st if v)} |")


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "raw":
        raw(sys.argv[2])
    elif mode == "list":
        launch_list(sys.argv[2])
    else:
        hot(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 20)
