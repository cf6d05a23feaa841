#!/usr/bin/env python
"""dram traffic per MD step of the persistent step kernel from an `ncu --set full` capture.
usage: ncu_traffic.py rep workload md_steps_in_the_captured_launch  -> updates profiles/r01_traffic.json"""
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
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[k]]


rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
f = Path(__file__).resolve().parent.parent / "profiles" / "r01_traffic.json"
j = json.loads(f.read_text()) if f.exists() else {}
j[workload] = {"kernel": d["Kernel Name"], "capture": Path(rep).name, "md_steps_in_launch": nsteps, "dram_read_bytes": rd,
               "dram_write_bytes": wr, "dram_bytes_per_md_step": (rd + wr) / nsteps,
               "gpu_time_us": float(d["gpu__time_duration.sum"].replace(",", "")) *
               {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3, "second": 1e6, "s": 1e6}.get(u["gpu__time_duration.sum"], 1.0)}
f.write_text(json.dumps(j, indent=1))
print(j[workload])
