"""Small md_steps runs of both k_md_fast modes and the general kernel, for compute-sanitizer."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from sis_b200 import fixtures as fx
from sis_b200.api import SisGpu

which = sys.argv[1] if len(sys.argv) > 1 else "fast"
if which == "fast":
    spec = fx.cag47(4, pbc=True)
    spec.nl_margin = 3.0
    g = SisGpu(0)
    spec.configure(g)
    g.init_velocities(0); g.neighbor_list(0)
    print("rebuilds", g.md_steps(40), "step", g.get_step())
elif which == "fused":
    spec = fx.cag47(4, pbc=True)
    spec.nl_margin = 3.0
    g = SisGpu(0, fast_kernel=0)
    spec.configure(g)
    g.init_velocities(0); g.neighbor_list(0)
    print("rebuilds", g.md_steps(30), "step", g.get_step())
elif which == "mc":
    spec = fx.t2s2hp("T2S2HP_folded", nstep_MC=1)
    spec.nl_margin = 0.5
    g = SisGpu(0)
    spec.configure(g)
    g.init_velocities(0); g.neighbor_list(0)
    print("rebuilds", g.md_steps(40), "step", g.get_step(), g.last_step_kernel())
else:
    temps = [290.0, 320.0, 350.0, 380.0]
    spec = fx.t2s2hp("T2S2HP_folded", nstep_MC=0).replicated(4, temps=temps)
    spec.nl_margin = 0.4
    g = SisGpu(0)
    spec.configure(g)
    g.remd_setup([4, 1, 1, 1], [temps, [], [], []], 10)
    for r in range(4):
        g.init_velocities(r); g.neighbor_list(r)
    print("rebuilds", g.md_steps(60), "step", g.get_step(), g.remd_stats())
