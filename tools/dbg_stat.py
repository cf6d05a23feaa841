import sys, json, importlib.util, numpy as np
sys.path.insert(0, '/root/repo')
spec = importlib.util.spec_from_file_location("m", "/root/repo/tests/golden/make_stat_fixture.py"); m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
from sis_b200.api import SisGpu
ref = json.load(open('/root/repo/tests/golden/stat_fixture.json'))["remd"]
print("oracle acc", (np.array(ref["acceptance"]["accepted"])/np.array(ref["acceptance"]["attempts"])).round(3))
print("oracle E", [round(ref[f"E_pot_T{k}"]["mean"],2) for k in range(8)], [round(ref[f"n_bp_T{k}"]["mean"],2) for k in range(8)])
m.NPROD = 100000
for opts in ({}, {"remd_inkernel": 0}, {"launch_mode": 0}):
    g = SisGpu(0, **opts)
    got = m.run_remd(g, seed=777, oracle=False)
    a = got["acceptance"]
    print(opts, "acc", (np.array(a["accepted"])/np.array(a["attempts"])).round(3), g.remd_stats())
    print("   E", [round(got[f"E_pot_T{k}"]["mean"],2) for k in range(8)], [round(got[f"n_bp_T{k}"]["mean"],2) for k in range(8)])
