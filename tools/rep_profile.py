#!/usr/bin/env python
"""phase clocks of k_md_rep (option "profile") on BASELINE config 3: cycles per MD step and per phase, averaged
over the replicas' CTAs.  usage: rep_profile.py [--nrep N] [--nex N] [--mc N] [--opt k=v]
The per-warp / per-chunk-kind / exchange-section lines need the instrumented build:
    python __graft_entry__.py --prof ;  SISGPU_LIB=sis_b200/libsisgpu_prof.so python tools/rep_profile.py ..."""
import argparse, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from sis_b200 import fixtures as fx
from sis_b200.api import SisGpu

ap = argparse.ArgumentParser()
ap.add_argument("--nrep", type=int, default=64)
ap.add_argument("--nex", type=int, default=10)
ap.add_argument("--mc", type=int, default=0)
ap.add_argument("--md-steps", type=int, default=2000)
ap.add_argument("--opt", action="append", default=[])
a = ap.parse_args()
n = a.nrep
temps = [293.15 + 70.0 / max(1, n - 1) * k for k in range(n)]
spec = fx.t2s2hp("T2S2HP_folded", nstep_MC=a.mc).replicated(n, temps=temps)
g = SisGpu(0)
for kv in a.opt:
    k, v = kv.split("=")
    g.set_option(k, float(v))
spec.configure(g)
if a.nex > 0:
    g.remd_setup([n, 1, 1, 1], [temps, [], [], []], a.nex)
for r in range(n):
    g.init_velocities(r)
    g.neighbor_list(r)
g.md_steps(a.md_steps)
g.set_option("profile", 1)
g.get_timers(reset=True)
g.get_state(0)
t0 = time.perf_counter()
nreb = g.md_steps(a.md_steps)
g.get_state(0)
dt = time.perf_counter() - t0
import ctypes as C
raw = np.zeros(16)
g._call("get_timers", raw.ctypes.data_as(C.POINTER(C.c_double)), C.c_int(0))
pall = g.cta_profile(1024)
p = pall[:min(n, 1000)]
steps = p[:, 7].mean()
names = ["P0 gate/mc", "P1 bp", "P1 static", "P2 integrate", "verdict", "exchange", "(xch wait)"]
print(f"{g.last_step_kernel()} nrep {n} nex {a.nex} mc {a.mc}: {1e6 * dt / a.md_steps:.2f} us/step wall, "
      f"{n * a.md_steps / dt / 1e6:.2f} M replica-steps/s, rebuilds {nreb}")
for i, nm in enumerate(names):
    print(f"  {nm:14s} {p[:, i].mean() / steps:9.0f} cycles/step  (min {p[:, i].min() / steps:.0f}, max {p[:, i].max() / steps:.0f})")
print(f"  total          {p[:, :6].sum(axis=1).mean() / steps:9.0f} cycles/step")
print("  block 0:", [round(v / steps) for v in p[0, :6]], " warps 0-3 {B2 wait, P2, B3 wait}:", [round(v / steps) for v in raw[4:16]])
tq = pall[1000:1002].reshape(-1)
names = ["bp angle roles", "bp dihedral roles", "bp combine", "gauss", "dihedral", "bond+angle", "pairs (128)", "owner sums"]
print("  block 0 chunks: " + ", ".join(f"{nm} {tq[2 * i] / max(1, tq[2 * i + 1]):.0f} cyc x {tq[2 * i + 1] / steps:.1f}/step" for i, nm in enumerate(names)))
xq = pall[1003]
if xq[5] > 0:
    print("  block 0 exchange (cycles per attempt): energies %.0f, publish %.0f, wait %.0f, replay %.0f, refresh %.0f" % tuple(xq[:5] / xq[5]))
