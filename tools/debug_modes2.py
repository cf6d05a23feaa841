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

def cmp(tag, ga, gb):
    xa, va, aa = ga.get_state(0); xb, vb, ab = gb.get_state(0)
    d = np.abs(xa - xb).max(axis=1)
    print(f"{tag}: x {d.max():.3e} at {d.argmax()} n>1e-9 {(d>1e-9).sum()}  v {np.abs(va-vb).max():.3e}  a {np.abs(aa-ab).max():.3e}", flush=True)
    return d

# (a) two launches, no rebuild
ga, gb = mk(10.0, launch_mode=0), mk(10.0, launch_mode=2)
for g in (ga, gb):
    g.md_steps(5); g.md_steps(5)
cmp("two launches", ga, gb)
# (b) explicit rebuild between launches
ga, gb = mk(10.0, launch_mode=0), mk(10.0, launch_mode=2)
for g in (ga, gb):
    g.md_steps(5); g.neighbor_list(0); g.md_steps(1)
cmp("rebuild then 1 step", ga, gb)
# forces right after the rebuild on identical state
ga, gb = mk(10.0, launch_mode=0), mk(10.0, launch_mode=2)
for g in (ga, gb):
    g.md_steps(30); g.neighbor_list(0)
cmp("30 steps + rebuild", ga, gb)
x, v, a = ga.get_state(0)
gb.set_state(0, x, v, a)
ga.md_steps(1); gb.md_steps(1)
d = cmp("one step from identical state", ga, gb)
print("worst beads", np.argsort(-d)[:12], np.sort(d)[::-1][:12])
wca, ele, bp, coef = gb.get_lists(0)
print("lists", len(wca), len(ele), len(bp))
print("x range", x.min(axis=0), x.max(axis=0))
