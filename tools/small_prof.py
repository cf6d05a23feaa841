import sys
sys.path.insert(0, '.')
from sis_b200 import fixtures as fx
from sis_b200.api import SisGpu
spec = fx.cag47(1)
g = SisGpu(0); g.set_option("profile", 1)
spec.configure(g); g.init_velocities(0); g.neighbor_list(0)
g.md_steps(2000); g.set_option("timing", 1); g.get_timers(reset=True)
g.md_steps(10000)
print(g.get_timers())
