#!/usr/bin/env python3
"""Summarise one kernel of an `ncu --set full` report as JSON (the files under profiles/).

    python tools/ncu_summary.py REPORT.ncu-rep "note" [units_per_launch] > profiles/NAME.json

units_per_launch (e.g. pixels) adds per-unit figures: warp instructions x 32 / units, DRAM bytes / unit.
"""
import csv
import io
import json
import subprocess
import sys

rep, note = sys.argv[1], sys.argv[2]
units = float(sys.argv[3]) if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
names, unit_row, vals = rows[hdr], rows[hdr + 1], rows[hdr + 2]
col = {n: i for i, n in enumerate(names)}
keep = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active",
]
keep += [n for n in names if n.startswith("smsp__average_warps_issue_stalled") and n.endswith("per_issue_active.ratio")]
m = {k: (vals[col[k]] + (" " + unit_row[col[k]] if unit_row[col[k]] else "")).strip() for k in keep if k in col}
out = {"kernel": vals[col["Kernel Name"]], "note": note, "metrics": m}
if units:
    f = lambda k: float(vals[col[k]].replace(",", ""))
    mult = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    dram = sum(f(k) * mult[unit_row[col[k]]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    tmul = {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}[unit_row[col["gpu__time_duration.sum"]]]
    out["per_unit"] = {"units_per_launch": units, "thread_instr_per_unit": f("smsp__inst_executed.sum") * 32 / units,
                       "dram_bytes_per_unit": dram / units,
                       "units_per_s_under_ncu": units / (f("gpu__time_duration.sum") * tmul)}
print(json.dumps(out, indent=1))
