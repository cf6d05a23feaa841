import sys, time
sys.path.insert(0, '.')
from sis_b200 import synth
from sis_b200.api import SisGpu
T = float(sys.argv[1]) if len(sys.argv) > 1 else 312.15
spec = synth.c4(); spec.tempK = T
g = SisGpu(0); spec.configure(g); g.init_velocities(0); g.neighbor_list(0)
g.set_option("timing", 1)
for c in range(int(sys.argv[2]) if len(sys.argv) > 2 else 12):
    g.get_timers(reset=True)
    t0 = time.perf_counter(); nreb = g.md_steps(1000); dt = time.perf_counter() - t0
    t = g.get_timers()
    print(c, f"wall {dt*1e3:.2f} ms", {k: round(v, 2) for k, v in t.items() if isinstance(v, float)}, "rebuilds", nreb, flush=True)
