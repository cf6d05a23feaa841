#!/usr/bin/env python
"""figures of the step kernel from an `ncu --set full` capture for bench.py's roofline object:
dram bytes per MD step, fp64-pipe and issue-slot utilisation, duration.
usage: ncu_traffic.py rep workload md_steps_in_the_captured_launch  -> updates profiles/r02_kernel_metrics.json"""
import csv, io, json, subprocess, sys
from pathlib import Path

rep, workload, nsteps = sys.argv[1], sys.argv[2], float(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, r = rows[0], rows[1], rows[2]
d = dict(zip(hdr, r))
u = dict(zip(hdr, units))


def val(k):
    v = float(d[k].replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u[k], 1)


def pct(k):
    return float(d[k].replace(",", "")) / 100.0 if k in d else None


rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
f = Path(__file__).resolve().parent.parent / "profiles" / "r02_kernel_metrics.json"
j = json.loads(f.read_text()) if f.exists() else {}
j[workload] = {"kernel": d["Kernel Name"], "capture": Path(rep).name, "md_steps_in_launch": nsteps, "dram_read_bytes": rd,
               "dram_write_bytes": wr, "dram_bytes_per_md_step": (rd + wr) / nsteps,
               "fp64_pipe_frac": pct("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
               "issue_slot_frac": pct("smsp__issue_active.avg.pct_of_peak_sustained_active"),
               "warps_active_frac": pct("sm__warps_active.avg.pct_of_peak_sustained_active"),
               "gpu_time_us": float(d["gpu__time_duration.sum"].replace(",", "")) *
               {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3, "second": 1e6, "s": 1e6}.get(u["gpu__time_duration.sum"], 1.0)}
f.write_text(json.dumps(j, indent=1))
print(j[workload])
