#!/usr/bin/env python
"""Multi-rank REMD with the exchange step inside the persistent kernels over peer-mapped buffers
(sisgpu_peer_export / _import) against the same replicas held by ONE handle: run under torchrun
on 2+ GPUs of one box.  Every rank also runs the single-handle reference on its own GPU, so the
comparison is local: label tables and acceptance counts must be identical, the states of the
rank's replicas equal to rounding.  Also checks the ncclAllGather chain path (no peer import)."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from sis_b200 import fixtures as fx
from sis_b200.api import SisGpu

world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
R = int(os.environ.get("SIS_R", "4"))
n, nex, nsteps = R * world, 10, 130
temps = list(np.linspace(290.0, 380.0, n))


def run(nrank, myrank, peers, inkernel=1):
    Rl = n // nrank
    spec = fx.t2s2hp("T2S2HP_folded", nstep_MC=0).replicated(Rl, temps=temps[myrank * Rl:(myrank + 1) * Rl])
    spec.nl_margin = 1.5
    g = SisGpu(local, remd_inkernel=inkernel)
    spec.configure(g)
    g.remd_setup([n, 1, 1, 1], [temps, [], [], []], nex, myrank, nrank)
    if nrank > 1:
        g.comm_init_torch()
        if peers:
            g.peer_init_torch()
    for r in range(Rl):
        # the Gaussian streams are keyed by the global replica id, so the split run draws the same numbers
        g.init_velocities(r)
        g.neighbor_list(r)
    for k in (45, 55, 30):
        g.md_steps(k)
    return g, [g.get_state(r) for r in range(Rl)]


ref, xref = run(1, 0, False)
out = {}
for name, peers, ink in (("peer", True, 1), ("nccl_chain", False, 1)):
    g, x = run(world, rank, peers, ink)
    same_lab = bool(np.array_equal(g.remd_state()[0], ref.remd_state()[0]))
    dx = max(np.abs(x[r][0] - xref[rank * R + r][0]).max() for r in range(R))
    dv = max(np.abs(x[r][1] - xref[rank * R + r][1]).max() for r in range(R))
    out[name] = dict(labels_equal=same_lab, stats=g.remd_stats(), ref_stats=ref.remd_stats(), dx=float(dx), dv=float(dv),
                     launches=g.launch_count(), kernel=g.last_step_kernel())
    ok = same_lab and g.remd_stats() == ref.remd_stats() and dx < 1e-8 and dv < 1e-8
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    out[name]["ok_all_ranks"] = bool(t.item())
    g.close()
    dist.barrier()
print(f"[rank {rank}] " + json.dumps(out), flush=True)
ok = all(v["ok_all_ranks"] for v in out.values())
dist.destroy_process_group()
sys.exit(0 if ok else 1)
