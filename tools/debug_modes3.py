import sys, numpy as np
sys.path.insert(0, "/root/repo")
from sis_b200 import fixtures as fx
from sis_b200.api import SisGpu

def mk(margin, **opts):
    spec = fx.cag47(4, pbc=True)
    spec.nl_margin = margin
    g = SisGpu(0, **opts)
    spec.configure(g)
    g.init_velocities(0)
    g.neighbor_list(0)
    return g

for n in range(10, 22):
    ga, gb = mk(2.0, launch_mode=0), mk(2.0, launch_mode=2)
    ra, rb = ga.md_steps(n), gb.md_steps(n)
    xa, va, aa = ga.get_state(0); xb, vb, ab = gb.get_state(0)
    d = np.abs(xa - xb).max(axis=1)
    print(f"n {n} nreb {ra} {rb} x {d.max():.3e} at {d.argmax()} cnt {(d>1e-9).sum()} v {np.abs(va-vb).max():.3e} a {np.abs(aa-ab).max():.3e} worst {np.argsort(-d)[:8]}", flush=True)
