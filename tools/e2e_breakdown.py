import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from sis_b200 import synth
from sis_b200.api import SisGpu
spec = synth.c4()
g = SisGpu(0); spec.configure(g); g.init_velocities(0); g.neighbor_list(0)
for _ in range(3): g.md_steps(1000)
x, v, a = g.get_state(0)
T = {k: 0.0 for k in "set md get ene".split()}
for _ in range(5):
    t0 = time.perf_counter(); g.set_state(0, x, v, a); t1 = time.perf_counter()
    g.md_steps(1000); t2 = time.perf_counter()
    x, v, a = g.get_state(0); t3 = time.perf_counter()
    e = g.energy_sumup(0, spec.tempK, False); t4 = time.perf_counter()
    T["set"] += t1 - t0; T["md"] += t2 - t1; T["get"] += t3 - t2; T["ene"] += t4 - t3
print({k: round(1e3 * v / 5, 3) for k, v in T.items()}, "ms per chunk")
t0 = time.perf_counter()
for _ in range(5): g.md_steps(1000)
print("md only", 1e3 * (time.perf_counter() - t0) / 5)
