SIS_CTA_PROFILE=gpurun_out/s7_cta.npy python bench.py --steps 5 --warmup 3 --profile --no-cpu "$@" > gpurun_out/s7_prof.json 2> gpurun_out/s7_prof.err
python - <<'PY'
import numpy as np, json
d = json.load(open('gpurun_out/s7_prof.json'))
print(d['value'], d['timers_ms'])
c = np.load('gpurun_out/s7_cta.npy')
np.set_printoptions(linewidth=220, suppress=True, precision=0)
print("cols: pair, bar3wait, integ, gridbar, bp_total, c1b-c1, nact, units*1000+beads")
print("mean", c.mean(0)); print("max", c.max(0)); print("min", c.min(0))
units = (c[:, 7] // 1000); beads = c[:, 7] % 1000
o = np.argsort(-c[:, 0])
print("slowest pair CTAs: idx, pair clk, units, beads, bp_total, nact")
for i in o[:12]: print(i, c[i, 0], units[i], beads[i], c[i, 4], c[i, 6])
print("fastest:")
for i in o[-6:]: print(i, c[i, 0], units[i], beads[i], c[i, 4], c[i, 6])
print("units: mean %.1f max %.0f min %.0f ; beads min %d max %d" % (units.mean(), units.max(), units.min(), beads.min(), beads.max()))
print("corr pair clk vs units", np.corrcoef(c[:, 0], units)[0, 1])
PY
