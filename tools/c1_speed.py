"""us per MD step of BASELINE config 1 (T2S2HP, 27 beads, nstep_MC = 1): stream / graph path vs the
Monte-Carlo pruning inside the replica-local persistent kernel."""
import sys, time
sys.path.insert(0, '.')
from sis_b200 import fixtures as fx
from sis_b200.api import SisGpu
for opts in ({}, {"mc_persistent": 1}):
    spec = fx.t2s2hp("T2S2HP_random", nstep_MC=1)
    g = SisGpu(0, **opts)
    spec.configure(g); g.init_velocities(0); g.neighbor_list(0)
    g.md_steps(2000)
    t0 = time.perf_counter(); nreb = g.md_steps(20000); dt = time.perf_counter() - t0
    print(opts, g.last_step_kernel(), f"{1e6*dt/20000:.2f} us/step", f"{27*20000/dt/1e6:.2f} M particle-steps/s", "rebuilds", nreb, flush=True)
