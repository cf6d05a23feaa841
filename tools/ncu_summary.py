#!/usr/bin/env python
"""Summarise an .ncu-rep here (no GPU needed): per-kernel headline metrics from the raw page and
the stall-reason totals + hottest SASS lines from the source page.  Output is the text that gets
committed under profiles/.

usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [--top 12] > profiles/rNN_x.txt
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

RAW_KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
            "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
            "sm__cycles_elapsed.max"]


def run(args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True).stdout


def raw(rep):
    out = run(["-i", rep, "--page", "raw", "--csv"])
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        return
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"== {d.get('Kernel Name', '?')[:70]}  (id {d.get('ID')})")
        for k in RAW_KEYS:
            if k in d:
                print(f"   {k:70s} {d[k]:>16s} {units[hdr.index(k)]}")


def source(rep, top):
    out = run(["-i", rep, "--page", "source", "--csv"])
    blocks = out.split('"Kernel Name",')
    for b in blocks[1:]:
        lines = b.splitlines()
        name = lines[0].strip().strip(',').strip('"')
        rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
        hdr = rows[0]
        idx = {h: i for i, h in enumerate(hdr)}
        stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        tot = defaultdict(int)
        per_line = []
        for r in rows[1:]:
            if len(r) < len(hdr):
                continue
            s = int(r[idx["# Samples"]] or 0)
            per_line.append((s, r[idx["Source"]].strip(), r))
            for c in stall_cols:
                tot[c] += int(r[idx[c]] or 0)
        total = sum(x[0] for x in per_line) or 1
        print(f"== source page: {name[:70]}  total samples {total}")
        print("   stall reasons: " + ", ".join(f"{k[6:]} {100 * v / total:.1f}%" for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]))
        for s, src, r in sorted(per_line, key=lambda x: -x[0])[:top]:
            why = max(stall_cols, key=lambda c: int(r[idx[c]] or 0))
            print(f"   {100 * s / total:5.1f}%  {src[:60]:60s} {why[6:]}")
        break   # first captured launch is enough


if __name__ == "__main__":
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 12
    raw(rep)
    source(rep, top)
