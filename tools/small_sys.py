"""us per MD step of small single systems (C2: one CAG47 chain; 4 chains PBC) against the k_md_fast grid size."""
import sys, time
sys.path.insert(0, '.')
from sis_b200 import fixtures as fx
from sis_b200.api import SisGpu
for name, spec in (("cag47x1", fx.cag47(1)), ("cag47x4_pbc", fx.cag47(4, pbc=True))):
    for ncta in (0,):
        g = SisGpu(0)
        if ncta:
            g.set_option("fast_max_cta", ncta)
        spec.configure(g); g.init_velocities(0); g.neighbor_list(0)
        g.md_steps(2000)
        g.set_option("timing", 1); g.get_timers(reset=True)
        t0 = time.perf_counter(); nreb = g.md_steps(10000); dt = time.perf_counter() - t0
        t = g.get_timers()
        print(name, "ncta", ncta, f"{1e6*dt/10000:.2f} us/step wall", f"step {1e3*t['step_ms']/10000:.2f} us nl {1e3*t['nl_ms']/10000:.2f} us", "rebuilds", nreb, flush=True)
