#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples of a kernel in an .ncu-rep
(needs -lineinfo and --import-source on).  usage: ncu_lines.py rep [--top N]"""
import csv, io, subprocess, sys, collections

rep = sys.argv[1]
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
agg = collections.defaultdict(lambda: [0, 0, ""])
fpath = "?"
hdr = None
for row in csv.reader(io.StringIO(out)):
    if not row:
        continue
    if row[0] == "File Path":
        fpath = row[1].split("/")[-1]; hdr = None; continue
    if row[0] == "Function Name":
        continue
    if row[0] == "Line No":
        hdr = row; continue
    if hdr is None or len(row) < len(hdr):
        continue
    try:
        line = int(row[0])
    except ValueError:
        continue
    # columns repeat "Source" (cuda text, then sass); use positions
    i_samp = hdr.index("# Samples"); i_inst = hdr.index("Instructions Executed")
    try:
        s = int(row[i_samp] or 0); n = int(row[i_inst] or 0)
    except ValueError:
        continue
    a = agg[(fpath, line)]
    a[0] += s; a[1] += n
    if not a[2]:
        a[2] = row[1].strip()
tot_s = sum(a[0] for a in agg.values()) or 1
tot_n = sum(a[1] for a in agg.values()) or 1
print(f"total samples {tot_s}, total warp instructions {tot_n}")
# by file
byfile = collections.defaultdict(lambda: [0, 0])
for (f, l), a in agg.items():
    byfile[f][0] += a[0]; byfile[f][1] += a[1]
for f, a in byfile.items():
    print(f"  {f:22s} samples {100*a[0]/tot_s:5.1f}%  instr {100*a[1]/tot_n:5.1f}%")
print("top lines by instructions executed:")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"  {100*a[1]/tot_n:5.1f}% instr {100*a[0]/tot_s:5.1f}% smp  {f}:{l:<5d} {a[2][:90]}")
