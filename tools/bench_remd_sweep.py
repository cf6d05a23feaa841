#!/usr/bin/env python
"""replica-steps/s of the force- and concentration-REMD inputs of the reference (BASELINE config 5,
"force/concentration REMD (T2S2HP_REMD_F/C) sweep"), one GPU, the in-tree force field substituted:
  C-REMD  check/T2S2HP_REMD_C/input_REMD.toml: 8 ionic strengths 0.02 .. 0.80 M, exchange every 10 steps, nstep_MC 0
  F-REMD  check/T2S2HP_REMD_F/input_REMD.toml: 8 forces 0 .. 18 pN on beads (1, 27), exchange every 100 steps,
          nstep_MC 10
`--copies k` runs k independent copies of the 8 labels side by side is NOT what the reference does; instead the
sweep is widened the way a user would: more values of the same variable (8 k replicas, interpolated)."""
import argparse, json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from sis_b200 import fixtures as fx
from sis_b200.api import SisGpu
from sis_b200.system import PN2KCALMOLA

ap = argparse.ArgumentParser()
ap.add_argument("--md-steps", type=int, default=2000)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--nrep", type=int, default=8)
a = ap.parse_args()
ION8 = [0.02, 0.05, 0.10, 0.15, 0.20, 0.30, 0.50, 0.80]
F8 = [0.0, 4.0, 8.0, 10.0, 12.0, 14.0, 16.0, 18.0]


def widen(v, n):
    return list(np.interp(np.linspace(0, len(v) - 1, n), np.arange(len(v)), v))


out = {}
for kind in ("C", "F"):
    n = a.nrep
    if kind == "C":
        vals = widen(ION8, n)
        spec = fx.t2s2hp("T2S2HP_folded", nstep_MC=0).replicated(n, ionic=vals)
        nrep, values, nex = [1, 1, n, 1], [[], [], vals, []], 10
    else:
        vals = [f * PN2KCALMOLA for f in widen(F8, n)]
        spec = fx.t2s2hp("T2S2HP_folded", nstep_MC=10).replicated(n)
        spec.twz_dcf_pairs = [(1, 27)]
        spec.twz_dcf_forces = [(1.0, 0.0, 0.0)]      # direction; the replica variable is the magnitude
        nrep, values, nex = [1, n, 1, 1], [[], vals, [], []], 100
    g = SisGpu(0)
    spec.configure(g)
    g.remd_setup(nrep, values, nex)
    for r in range(n):
        g.init_velocities(r)
        g.neighbor_list(r)
    g.md_steps(a.md_steps)
    l0 = g.launch_count()
    g.mark(0)
    for _ in range(a.steps):
        g.md_steps(a.md_steps)
    g.mark(1)
    dt = 1e-3 * g.elapsed_ms()
    out[kind] = {"nrep": n, "nstep_exchange": nex, "replica_steps_per_s": n * a.md_steps * a.steps / dt,
                 "us_per_md_step": 1e6 * dt / (a.md_steps * a.steps), "kernel": g.last_step_kernel(),
                 "launches": g.launch_count() - l0, "exchange": g.remd_stats()}
print(json.dumps(out))
