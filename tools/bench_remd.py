#!/usr/bin/env python
"""replica-steps/s of T-REMD on one GPU: BASELINE config 3 (T2S2HP, 27 beads, nrep replicas,
temperatures 293.15..363.15 K, exchange attempt every 10 steps, nstep_MC 0)."""
import argparse, json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from sis_b200 import fixtures as fx
from sis_b200.api import SisGpu

ap = argparse.ArgumentParser()
ap.add_argument("--nrep", type=int, default=64)
ap.add_argument("--nex", type=int, default=10)
ap.add_argument("--md-steps", type=int, default=2000)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--opt", action="append", default=[])
ap.add_argument("--profile", action="store_true")
a = ap.parse_args()
n = a.nrep
temps = [293.15 + 70.0 / (n - 1) * k for k in range(n)]
spec = fx.t2s2hp("T2S2HP_folded", nstep_MC=0).replicated(n, temps=temps)
g = SisGpu(0)
for kv in a.opt:
    k, v = kv.split("=")
    g.set_option(k, float(v))
spec.configure(g)
g.remd_setup([n, 1, 1, 1], [temps, [], [], []], a.nex)
for r in range(n):
    g.init_velocities(r)
    g.neighbor_list(r)
for _ in range(a.warmup):
    g.md_steps(a.md_steps)
if a.profile:
    g.set_option("profile", 1)
g.set_option("timing", 1)
g.get_timers(reset=True)
l0 = g.launch_count()
t0 = time.perf_counter()
nreb = 0
for _ in range(a.steps):
    nreb += g.md_steps(a.md_steps)
g.get_state(0)
dt = time.perf_counter() - t0
att, acc = g.remd_stats()
print(json.dumps({"metric": "replica-steps/s", "value": n * a.md_steps * a.steps / dt, "nrep": n, "nmp": spec.nmp,
                  "us_per_md_step": 1e6 * dt / (a.md_steps * a.steps), "rebuilds": nreb, "launches": g.launch_count() - l0,
                  "timers_ms": g.get_timers(), "exchange": {"attempts": att, "accepted": acc}}))
