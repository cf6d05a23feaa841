import sys, numpy as np
sys.path.insert(0, "/root/repo")
from sis_b200 import fixtures as fx
from sis_b200.api import SisGpu

def run(n, margin, **opts):
    spec = fx.cag47(4, pbc=True)
    spec.nl_margin = margin
    g = SisGpu(0, **opts)
    spec.configure(g)
    g.init_velocities(0)
    g.neighbor_list(0)
    nreb = g.md_steps(n)
    x, v, a = g.get_state(0)
    return nreb, x, v, a

for margin in (10.0, 2.0):
    for n in (1, 2, 3, 5, 10, 20, 40, 100):
        r0, x0, v0, a0 = run(n, margin, launch_mode=0)
        r1, x1, v1, a1 = run(n, margin, launch_mode=2, fast_kernel=1)
        r2, x2, v2, a2 = run(n, margin, launch_mode=2, fast_kernel=0)
        d1 = np.abs(x1 - x0).max(axis=1); d2 = np.abs(x2 - x0).max(axis=1)
        print(f"margin {margin} n {n:3d} nreb {r0} {r1} {r2}  fast: max {d1.max():.3e} at bead {d1.argmax()} (#>1e-9: {(d1>1e-9).sum()})  fused: max {d2.max():.3e} at {d2.argmax()}", flush=True)
