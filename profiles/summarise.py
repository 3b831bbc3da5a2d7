#!/usr/bin/env python
"""Turn the raw ncu outputs that came back in gpurun_out/ into the small tracked summaries under profiles/.

    python profiles/summarise.py launches gpurun_out/launches_r01b.csv profiles/r01b_launches.md
    python profiles/summarise.py kernel   gpurun_out/eval_r01b.ncu-rep   profiles/r01b_evaluate_kernel.md

`launches`: per-kernel launch count / total / mean device time and SHARE of the captured launches
            (ncu --metrics gpu__time_duration.sum; cold-cache, serialised: shares are meaningful, absolutes are not).
`kernel`:   the headline counters of one `ncu --set full` capture (needs `ncu` on PATH to read the .ncu-rep).
`traffic`:  DRAM bytes (read + written) per launch of every per-call kernel, grouped by the phases bench.py reports, as JSON
            (bench.py reads the newest profiles/*_dram_traffic.json for `roofline.traffic`):
                python profiles/summarise.py traffic gpurun_out/r02_step.ncu-rep profiles/r02_dram_traffic.json <git hash>
            The LAST launch of each kernel in the capture is used (the steady-state step of the default workload).
"""
from __future__ import annotations

import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "sm__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "sm__inst_issued.avg.per_cycle_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_not_selected_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
    "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct", "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct",
    "smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_membar_per_warp_active.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]


def short(name: str) -> str:
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"<.*", "", name)
    name = re.sub(r"\(.*", "", name)
    return name.split("::")[-1] if "at::" not in name else "torch:" + name.split("::")[-1]


def launches(src: str, dst: str) -> None:
    with open(src) as f:
        text = f.read()
    text = text[text.index('"ID"'):]
    rows = list(csv.DictReader(io.StringIO(text)))
    agg: "OrderedDict[str, list[float]]" = OrderedDict()
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
        agg.setdefault(short(r["Kernel Name"]), []).append(us)
    total = sum(sum(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list: {src}\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` "
                f"(cold-cache, serialised; compare shares).  {sum(len(v) for v in agg.values())} launches, "
                f"{total/1e3:.3f} ms in total.\n\n| kernel | launches | total us | mean us | share |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| `{k}` | {len(v)} | {sum(v):.1f} | {sum(v)/len(v):.1f} | {100*sum(v)/total:.1f}% |\n")
    print(open(dst).read())


PHASE_OF = {"bin_count_kernel": "bin", "scan_reduce_kernel": "scan", "scan_blocks_kernel": "scan", "scan_apply_kernel": "scan",
            "bin_scatter_kernel": "scatter", "evaluate_kernel": "evaluate", "exact_kernel": "evaluate", "direct_kernel": "evaluate",
            "ring_kernel": "tree", "bvh_kernel": "tree", "brute_cull_kernel": "exhaustive", "pending_winner_kernel": "pending",
            "finalize_rows_kernel": "epilogue"}


def traffic(src: str, dst: str, git: str) -> None:
    import json
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    header, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(header)}

    def num(row, name):
        v = float(row[col[name]].replace(",", ""))
        u = units[col[name]]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6,
                    "nsecond": 1e-3, "usecond": 1, "msecond": 1e3, "second": 1e6}.get(u, 1)
    last = OrderedDict()
    for row in rows[2:]:
        last[short(row[col["Kernel Name"]])] = row
    kernels, per_kernel = {}, {}
    for name, row in last.items():
        phase = PHASE_OF.get(name)
        if phase is None:
            continue
        b = num(row, "dram__bytes_read.sum") + num(row, "dram__bytes_write.sum")
        kernels[phase] = kernels.get(phase, 0.0) + b
        per_kernel[name] = {"dram_bytes": b, "us": num(row, "gpu__time_duration.sum")}
    with open(dst, "w") as f:
        json.dump({"git": git, "capture": src, "workload": "bench.py default: 10M points x 50k cylinders, variant A, one B200",
                   "kernels": kernels, "per_kernel": per_kernel}, f, indent=1)
    print(open(dst).read())


def kernel(src: str, dst: str) -> None:
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    header, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full: {src}\n\n")
        for row in rows[2:]:
            rec = dict(zip(header, row))
            f.write(f"## `{short(rec.get('Kernel Name', '?'))}`  grid {rec.get('Grid Size')} block {rec.get('Block Size')}\n\n"
                    "| metric | value | unit |\n|---|---:|---|\n")
            for name in KEEP:
                if name in rec:
                    f.write(f"| {name} | {rec[name]} | {units[header.index(name)]} |\n")
            try:
                rd = float(rec["dram__bytes_read.sum"].replace(",", ""))
                wr = float(rec["dram__bytes_write.sum"].replace(",", ""))
                ur, uw = units[header.index("dram__bytes_read.sum")], units[header.index("dram__bytes_write.sum")]
                scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                f.write(f"\nDRAM traffic per launch: {(rd*scale[ur] + wr*scale[uw])/1e6:.2f} MB\n\n")
            except (KeyError, ValueError):
                pass
    print(open(dst).read())


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2], sys.argv[3])
