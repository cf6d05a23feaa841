#!/usr/bin/env python
"""replica-steps/s of T-REMD: BASELINE config 3 (T2S2HP, 27 beads, temperatures 293.15..363.15 K,
exchange attempt every 10 steps, nstep_MC 0), `--nrep` replicas PER GPU (weak scaling along
replicas, the axis the reference shards on).  One process per GPU under torchrun; the exchange
all-gather runs over NCCL inside the library."""
import argparse, json, os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from sis_b200 import fixtures as fx
from sis_b200.api import SisGpu

ap = argparse.ArgumentParser()
ap.add_argument("--nrep", type=int, default=64, help="replicas per GPU")
ap.add_argument("--nex", type=int, default=10)
ap.add_argument("--md-steps", type=int, default=2000)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--opt", action="append", default=[])
ap.add_argument("--profile", action="store_true")
a = ap.parse_args()
world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
import torch
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
R, n = a.nrep, a.nrep * world
temps = [293.15 + 70.0 / (n - 1) * k for k in range(n)]
spec = fx.t2s2hp("T2S2HP_folded", nstep_MC=0).replicated(R, temps=temps[rank * R:(rank + 1) * R])
g = SisGpu(local)
for kv in a.opt:
    k, v = kv.split("=")
    g.set_option(k, float(v))
spec.configure(g)
g.remd_setup([n, 1, 1, 1], [temps, [], [], []], a.nex, rank, world)
if world > 1:
    g.comm_init_torch()
for r in range(R):
    g.init_velocities(r)
    g.neighbor_list(r)


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for _ in range(a.warmup):
    g.md_steps(a.md_steps)
if a.profile:
    g.set_option("profile", 1)
g.set_option("timing", 1)
g.get_timers(reset=True)
l0 = g.launch_count()
barrier()
t0 = time.perf_counter()
nreb = 0
for _ in range(a.steps):
    nreb += g.md_steps(a.md_steps)
g.get_state(0)
barrier()
dt = time.perf_counter() - t0
if world > 1:
    tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt.item())
att, acc = g.remd_stats()
if rank == 0:
    print(json.dumps({"metric": "replica-steps/s", "value": n * a.md_steps * a.steps / dt, "n_gpus": world, "nrep_all": n,
                      "nrep_per_gpu": R, "nmp": spec.nmp, "nstep_exchange": a.nex,
                      "us_per_md_step": 1e6 * dt / (a.md_steps * a.steps), "rebuilds_rank0": nreb,
                      "launches_rank0": g.launch_count() - l0, "timers_ms": g.get_timers(),
                      "exchange": {"attempts": att, "accepted": acc}}))
if world > 1:
    dist.destroy_process_group()
