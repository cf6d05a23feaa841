#!/bin/bash
# ncu launch list (gpu__time_duration) of a command; usage: tools/ncu_launches.sh out.csv <command...>
out=$1; shift
ncu --metrics gpu__time_duration.sum --clock-control none -c ${NCU_COUNT:-600} --csv --log-file "$out" "$@" > /dev/null 2>&1
python - "$out" <<'PY'
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1], errors='ignore')))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID']
if not hdr:
    print('no launches captured'); sys.exit(0)
h = rows[hdr[0]]
ki, vi = h.index('Kernel Name'), h.index('Metric Value')
d = collections.OrderedDict()
for r in rows[hdr[0] + 2:]:
    if len(r) > vi:
        d.setdefault(r[ki][:70], []).append(float(r[vi].replace(',', '')))
tot = sum(sum(v) for v in d.values())
for k, v in d.items():
    print(f"{k:72s} n={len(v):4d} mean={sum(v)/len(v)/1000:9.1f} us share={100*sum(v)/tot:5.1f}%")
PY
