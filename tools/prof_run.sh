SIS_CTA_PROFILE=gpurun_out/s7_cta.npy python bench.py --steps 5 --warmup 3 --profile --no-cpu > gpurun_out/s7_prof.json 2> gpurun_out/s7_prof.err
python - <<'PY'
import numpy as np, json
d = json.load(open('gpurun_out/s7_prof.json'))
print(d['value'], d['timers_ms'])
c = np.load('gpurun_out/s7_cta.npy')
np.set_printoptions(linewidth=200, suppress=True, precision=0)
print("cols: pair, bar3wait(?), integ, gridbar, bp_total, c1b-c1, nact, -")
print("mean", c.mean(0)); print("max", c.max(0)); print("min", c.min(0))
PY
