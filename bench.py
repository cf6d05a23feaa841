#!/usr/bin/env python
"""bench.py -- particle-steps/s of the SIS per-step hot path (BASELINE.json metric).

Workload at every N: 64 x CAG47 chains (9 024 beads) in the periodic 809.972 A box, GJF-2GJ
Langevin dynamics, fp64, synthetic random coils (seed 1) -- BASELINE config 4.  One "step" of
the contract is one chunk of `--md-steps` MD steps.  With N > 1 GPUs every rank runs one
temperature replica of that system (T-REMD, exchange attempt every chunk through the NCCL
all-gather): weak scaling along replicas, the only axis the reference shards on.

  value : device-resident throughput (state stays in HBM across chunks)
  e2e   : through the C ABI with HOST buffers: every chunk uploads xyz/velos/accels, runs the
          steps and downloads xyz/velos/accels + the 11 energies (what the Fortran driver does
          at each save step)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))


def clocks_sampler(stop, out):
    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    dev = os.environ.get("LOCAL_RANK", "0")
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", dev],
                               capture_output=True, text=True, timeout=5)
            f = [x.strip() for x in r.stdout.strip().split(",")]
            if len(f) >= 6:
                out.append(f)
        except Exception:
            pass
        stop.wait(0.2)


def summarize_clocks(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
    sm = sorted(int(s[0]) for s in samples if s[0].isdigit())
    reasons = set()
    for s in samples:
        for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], s[2:6]):
            if v.lower().startswith("active"):
                reasons.add(name)
    return {"sm_mhz": sm[len(sm) // 2] if sm else None,
            "sm_max_mhz": int(samples[0][1]) if samples[0][1].isdigit() else None, "reasons": sorted(reasons)}


def build_system(args):
    from sis_b200 import synth
    if args.workload == "c4":
        return synth.c4(), "64xCAG47 PBC box 809.972 A, 9024 beads (BASELINE config 4)"
    if args.workload == "c5":
        return synth.c5(), "1024xCAG47 PBC box 2041.0 A, 144384 beads (BASELINE config 5)"
    raise SystemExit(f"unknown workload {args.workload}")


def list_stats(g):
    nw, ne, nb = g.get_list_sizes(0)
    return {"P_wca": nw, "P_ele": ne, "P_bp": nb}


def algorithmic_bytes_per_step(nmp, st):
    # B_step = 24N + 24N + 224N + 8 P_wca + 8 P_ele + 53 P_bp
    return 272 * nmp + 8 * st["P_wca"] + 8 * st["P_ele"] + 53 * st["P_bp"]


def ncu_traffic(workload):
    """dram bytes (read + write) per MD step of the step kernel from the committed ncu --set full
    capture of this workload (profiles/r01_traffic.json, written by tools/ncu_traffic.py); None if
    there is no capture for it."""
    f = ROOT / "profiles" / "r01_traffic.json"
    if not f.exists():
        return None
    return json.loads(f.read_text()).get(workload, {}).get("dram_bytes_per_md_step")


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (the oracle port: no Fortran compiler exists
    here, so oracle/_ref cannot be built) on the host cores, bounded sample of the same workload."""
    from oracle.oracle import Oracle
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    spec, wl = build_system(args)
    cores = os.cpu_count() or 1
    o = Oracle(omp=True)
    spec.configure(o)
    o.init_velocities(0)
    o.neighbor_list(0)
    nstep = max(2, args.ref_md_steps)
    for _ in range(args.warmup):
        o.md_run(nstep)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.md_run(nstep)
    dt = time.perf_counter() - t0
    value = spec.nmp * nstep * args.steps / dt
    line = {"impl": "reference", "metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl, "md_steps_per_step": nstep},
            "cpu_baseline": {"value": value, "unit": "particle-steps/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps} x {nstep} MD steps of the same system (oracle, OpenMP)"},
            "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="sis_b200")
    ap.add_argument("--workload", default="c4")
    ap.add_argument("--md-steps", type=int, default=1000, help="MD steps per contract step (chunk)")
    ap.add_argument("--ref-md-steps", type=int, default=20)
    ap.add_argument("--cpu-baseline-steps", type=int, default=200)
    ap.add_argument("--launch-mode", type=int, default=None, help="0 stream launches, 1 CUDA-graph chunks")
    ap.add_argument("--chunk", type=int, default=None)
    ap.add_argument("--tpb", type=int, default=None, help="lanes per bead in the pair task")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="library option name=value (repeatable)")
    ap.add_argument("--profile", action="store_true", help="phase clocks of the persistent kernel")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from sis_b200.api import SisGpu

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        # one block of host cores per rank: the enqueueing thread, the NCCL proxy and the clock sampler of
        # a rank must not share cores with the spinning threads of the other ranks
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = max(1, len(cores) // world)
            if not os.environ.get("SIS_NO_PIN"):
                os.sched_setaffinity(0, set(cores[local * per:(local + 1) * per]))
        except (AttributeError, OSError):
            pass
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    spec, wl = build_system(args)
    g = SisGpu(local)
    for name, val in (("launch_mode", args.launch_mode), ("chunk", args.chunk), ("threads_per_bead", args.tpb)):
        if val is not None:
            g.set_option(name, val)
    for kv in args.opt:
        name, val = kv.split("=")
        g.set_option(name, float(val))
    spec.configure(g)
    if args.profile:
        g.set_option("profile", 1)
    if world > 1:
        temps = [spec.tempK + 2.0 * k for k in range(world)]
        g.remd_setup([world, 1, 1, 1], [temps, [], [], []], args.md_steps, rank, world)
        g.comm_init_torch()
    g.init_velocities(0)
    g.neighbor_list(0)
    st = list_stats(g)
    nmp = spec.nmp

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        g.md_steps(args.md_steps)
    samples, stop = [], threading.Event()
    th = threading.Thread(target=clocks_sampler, args=(stop, samples), daemon=True)
    if not os.environ.get("SIS_NO_SAMPLER"):
        th.start()

    # ---- end to end through the C ABI with host buffers (the headline) ------------------
    # every chunk: upload xyz / velos / accels, run the steps, download them + the 11 energies.
    # Measured first: the synthetic coils relax over the run and the step slowly gets more
    # expensive (more formed base pairs), so both loops are kept close to the initial state.
    host = tuple(torch.empty((nmp, 3), dtype=torch.float64, pin_memory=True).numpy() for _ in range(3))   # pinned host buffers
    x, v, a = g.get_state(0, out=host)
    barrier()
    g.mark(0)
    for _ in range(args.steps):
        g.set_state(0, x, v, a)
        g.md_steps(args.md_steps)
        x, v, a = g.get_state(0, out=host)
        e = g.energy_sumup(0, spec.tempK, False)
    g.mark(1)
    dte = 1e-3 * g.elapsed_ms()
    barrier()
    tt = torch.tensor([dte], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dte = float(tt.item())

    # ---- device-resident throughput ---------------------------------------------------
    g.set_option("timing", 1)
    g.get_timers(reset=True)
    l0 = g.launch_count()
    barrier()
    g.mark(0)                      # CUDA events on the library's stream: device-side time of the region
    nreb = 0
    for _ in range(args.steps):
        nreb += g.md_steps(args.md_steps)
    g.mark(1)
    dt = 1e-3 * g.elapsed_ms()
    barrier()
    launches = g.launch_count() - l0
    timers = g.get_timers(reset=False)
    if args.profile and os.environ.get("SIS_CTA_PROFILE"):
        cp = g.cta_profile(148)
        nst = max(1.0, timers.get("persistent_clk_per_step") and args.md_steps * args.steps or 1.0)
        np.save(os.environ["SIS_CTA_PROFILE"], cp / nst)
    g.get_timers(reset=True)
    g.set_option("timing", 0)
    tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt.item())
    value = world * nmp * args.md_steps * args.steps / dt
    stop.set()
    if th.is_alive():
        th.join(timeout=2)
    if os.environ.get("SIS_RANK_DEBUG"):
        print(f"[rank {rank}] value-loop {dt * 1e3 / args.steps:.2f} ms/chunk, e2e {dte * 1e3 / args.steps:.2f} ms/chunk, "
              f"timers {timers}, rebuilds {nreb}", file=sys.stderr, flush=True)
    e2e = world * nmp * args.md_steps * args.steps / dte

    if rank == 0:
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            peaks = json.loads(pk.read_text())
        hbm = peaks.get("hbm_gbs", 6650.0)
        step_ms = timers["step_ms"]
        nsteps_total = args.md_steps * args.steps
        bytes_step = algorithmic_bytes_per_step(nmp, st)
        achieved = bytes_step * nsteps_total / (step_ms * 1e-3) / 1e9 if step_ms > 0 else None
        line = {
            "metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl, "md_steps_per_step": args.md_steps, "lists": st,
                       "parallelism": f"replicas x{world}" if world > 1 else "single system",
                       "l2": "state (<2 MB) is L2-resident by design; no flush between MD steps of one trajectory"},
            "e2e": {"value": e2e, "unit": "particle-steps/s", "h2d_bytes_per_step": 3 * nmp * 24,
                    "d2h_bytes_per_step": 3 * nmp * 24 + 88},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s",
                         "frac": (achieved / hbm) if achieved else None, "traffic": ncu_traffic(args.workload),
                         "kernel": g.last_step_kernel() + " (one launch = one MD step of the persistent kernel)",
                         "algorithmic_bytes_per_launch": bytes_step,
                         "us_per_launch": 1e3 * step_ms / nsteps_total if step_ms else None,
                         "peak_source": "measured" if pk.exists() else "fallback"},
            "clocks": summarize_clocks(samples),
            "rebuilds": nreb, "timers_ms": timers,
        }
        # CPU baseline (oracle port, bounded sample) on this box's host cores
        try:
            if args.no_cpu:
                raise RuntimeError("skipped (--no-cpu)")
            if world > 1:
                raise RuntimeError("reported at N = 1 only")
            from oracle.oracle import Oracle
            o = Oracle(omp=True)
            spec.configure(o)
            o.init_velocities(0)
            o.neighbor_list(0)
            n = args.cpu_baseline_steps
            o.md_run(5)
            t0 = time.perf_counter()
            o.md_run(n)
            dtc = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": nmp * n / dtc, "unit": "particle-steps/s", "cores": os.cpu_count(),
                                    "kind": "port", "sample": f"{n} MD steps of the same system (oracle, OpenMP)"}
        except Exception as ex:  # pragma: no cover
            line["cpu_baseline"] = {"value": None, "error": str(ex)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
