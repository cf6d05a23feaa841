#!/usr/bin/env python
"""bench.py -- particle-steps/s of the SIS per-step hot path (BASELINE.json metric).

Workload at every N (default): 64 x CAG47 chains (9 024 beads) in the periodic 809.972 A box,
GJF-2GJ Langevin dynamics, fp64, synthetic random coils (seed 1) -- BASELINE config 4.  One
"step" of the contract is one chunk of `--md-steps` MD steps.  With N > 1 GPUs every rank runs
one temperature replica of that system (T-REMD, one exchange attempt per chunk): weak scaling
along replicas, the only axis the reference shards on.

  value : device-resident throughput (state stays in HBM across chunks)
  e2e   : through the C ABI with HOST buffers: every chunk uploads xyz/velos/accels, runs the
          steps and downloads xyz/velos/accels + the 11 energies (what the Fortran driver does
          at each save step)
Both loops start from the SAME saved state (positions, velocities, beta, step counter, fresh
lists), so they time the same trajectory; the end-to-end loop runs first.

`extra.c3_remd` (every N): BASELINE config 3 -- T2S2HP T-REMD, 64 replicas in total split over
the N ranks (strong scaling along replicas), exchange attempt every 10 steps, replica-steps/s.
`--workload c5` / `--workload c3` make those the main line instead.
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

PROFILE_JSON = ROOT / "profiles" / "r02_kernel_metrics.json"   # ncu --set full figures per workload (tools/ncu_traffic.py)


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


def build_system(workload):
    from sis_b200 import synth
    if workload == "c4":
        return synth.c4(), "64xCAG47 PBC box 809.972 A, 9024 beads (BASELINE config 4)"
    if workload == "c5":
        return synth.c5(), "1024xCAG47 PBC box 2041.0 A, 144384 beads (BASELINE config 5)"
    raise SystemExit(f"unknown workload {workload}")


def list_stats(g):
    nw, ne, nb = g.get_list_sizes(0)
    return {"P_wca": nw, "P_ele": ne, "P_bp": nb}


def algorithmic_bytes_per_step(nmp, st):
    # B_step = 24N + 24N + 224N + 8 P_wca + 8 P_ele + 53 P_bp
    return 272 * nmp + 8 * st["P_wca"] + 8 * st["P_ele"] + 53 * st["P_bp"]


def kernel_metrics(workload):
    """figures of the step kernel of this workload from the committed `ncu --set full` capture
    (profiles/r02_kernel_metrics.json, written by tools/ncu_traffic.py); {} if there is none."""
    for f in (PROFILE_JSON, ROOT / "profiles" / "r01_traffic.json"):
        if f.exists():
            d = json.loads(f.read_text()).get(workload)
            if d:
                return d
    return {}


def cpu_sample(spec, nsteps, warm=5):
    """the CPU port of the reference's algorithm (oracle, OpenMP build: the loops the reference
    threads, serial O(N^2) neighbor_list as in the reference) on this box's host cores"""
    from oracle.oracle import Oracle
    o = Oracle(omp=True)
    spec.configure(o)
    o.init_velocities(0)
    o.neighbor_list(0)
    if warm:
        o.md_run(warm)
    o.nl_seconds(reset=True)
    t0, c0 = time.perf_counter(), time.process_time()
    nreb = o.md_run(nsteps)
    dt, cpu = time.perf_counter() - t0, time.process_time() - c0
    tnl = o.nl_seconds()
    return {"value": spec.nmp * nsteps / dt, "unit": "particle-steps/s", "cores": o.num_threads(), "kind": "port",
            "threads_busy": round(cpu / dt, 2), "rebuilds": int(nreb), "rebuild_share": round(tnl / dt, 3),
            "value_without_rebuilds": spec.nmp * nsteps / max(dt - tnl, 1e-9), "seconds": round(dt, 2)}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm on the host cores.  oracle/_ref cannot be
    built (the reference is Fortran 2008, the image has no Fortran compiler), so this is the oracle
    port with the reference's OpenMP threading; each step is a bounded sample of `--ref-md-steps` MD
    steps of the same workload, list rebuilds (serial O(N^2), as in the reference) included."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.oracle import Oracle
    spec, wl = build_system(args.workload if args.workload in ("c4", "c5") else "c4")
    o = Oracle(omp=True)
    spec.configure(o)
    o.init_velocities(0)
    o.neighbor_list(0)
    nstep = max(2, args.ref_md_steps)
    for _ in range(args.warmup):
        o.md_run(nstep)
    o.nl_seconds(reset=True)
    t0, c0 = time.perf_counter(), time.process_time()
    nreb = 0
    for _ in range(args.steps):
        nreb += o.md_run(nstep)
    dt, cpu = time.perf_counter() - t0, time.process_time() - c0
    tnl = o.nl_seconds()
    value = spec.nmp * nstep * args.steps / dt
    threads = o.num_threads()
    sample = (f"{args.steps} x {nstep} MD steps of the same system (oracle port, OpenMP {threads} threads, "
              f"{nreb} serial O(N^2) list rebuilds = {100 * tnl / dt:.0f} % of the time)")
    line = {"impl": "reference", "metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl, "md_steps_per_step": nstep},
            "cpu_baseline": {"value": value, "unit": "particle-steps/s", "cores": threads, "kind": "port",
                             "sample": sample, "threads_busy": round(cpu / dt, 2), "rebuilds": int(nreb),
                             "rebuild_share": round(tnl / dt, 3),
                             "value_without_rebuilds": spec.nmp * nstep * args.steps / max(dt - tnl, 1e-9)},
            "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


class Ctx:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        if self.world > 1:
            # one block of host cores per rank: the enqueueing thread, the NCCL proxy and the clock sampler of
            # a rank must not share cores with the spinning threads of the other ranks
            try:
                cores = sorted(os.sched_getaffinity(0))
                per = max(1, len(cores) // self.world)
                if not os.environ.get("SIS_NO_PIN"):
                    os.sched_setaffinity(0, set(cores[self.local * per:(self.local + 1) * per]))
            except (AttributeError, OSError):
                pass
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        tt = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(tt, op=self.dist.ReduceOp.MAX)
        return float(tt.item())

    def new_handle(self):
        from sis_b200.api import SisGpu
        g = SisGpu(self.local)
        a = self.args
        for name, val in (("launch_mode", a.launch_mode), ("chunk", a.chunk), ("threads_per_bead", a.tpb)):
            if val is not None:
                g.set_option(name, val)
        for kv in a.opt:
            name, val = kv.split("=")
            g.set_option(name, float(val))
        return g


def run_remd_c3(cx, nrep_all=64, nex=int(os.environ.get("SIS_NEX", "10")), md_steps=2000, steps=5, warmup=2, scaling="strong"):
    """BASELINE config 3: T2S2HP T-REMD (check/T2S2HP_REMD_T/input_REMD.toml with 64 temperatures
    293.15 .. 363.15 K), nrep_all replicas split over the ranks, exchange attempt every `nex` steps"""
    from sis_b200 import fixtures as fx
    world, rank = cx.world, cx.rank
    if nrep_all % world:
        return {"error": f"{nrep_all} replicas do not split over {world} ranks"}
    R = nrep_all // world
    # up to 256 temperatures (MAX_REP_PER_DIM of the reference, src/const.F90:26-28); beyond that a second replica
    # dimension: a few ionic strengths per temperature (labels: the first variable varies slowest, src/init_replica.F90:70-92)
    nion = next(d for d in range(1, nrep_all + 1) if nrep_all % d == 0 and nrep_all // d <= 256)
    ntemp = nrep_all // nion
    tgrid = [293.15 + 70.0 / (ntemp - 1) * k for k in range(ntemp)]
    igrid = [0.15 + 0.01 * k for k in range(nion)]
    temps = [tgrid[l // nion] for l in range(nrep_all)]
    ions = [igrid[l % nion] for l in range(nrep_all)]
    spec = fx.t2s2hp("T2S2HP_folded", nstep_MC=0).replicated(R, temps=temps[rank * R:(rank + 1) * R],
                                                              ionic=ions[rank * R:(rank + 1) * R] if nion > 1 else None)
    g = cx.new_handle()
    spec.configure(g)
    g.remd_setup([ntemp, 1, nion, 1], [tgrid, [], igrid if nion > 1 else [], []], nex, rank, world)
    if world > 1:
        g.comm_init_torch()
        if not os.environ.get("SIS_NO_PEER"):
            g.peer_init_torch()      # exchange steps inside the kernels over NVLink peer memory
    for r in range(R):
        g.init_velocities(r)
        g.neighbor_list(r)
    for _ in range(warmup):
        g.md_steps(md_steps)
    if cx.args.profile:
        g.set_option("profile", 1)
    g.set_option("timing", 1)
    g.get_timers(reset=True)
    l0 = g.launch_count()
    cx.barrier()
    g.mark(0)
    nreb = 0
    for _ in range(steps):
        nreb += g.md_steps(md_steps)
    g.mark(1)
    dt = 1e-3 * g.elapsed_ms()
    cx.barrier()
    dt = cx.max_over_ranks(dt)
    att, acc = g.remd_stats()
    out = {"metric": "replica-steps/s", "value": nrep_all * md_steps * steps / dt, "unit": "replica-steps/s",
           "workload": "T2S2HP T-REMD, 27 beads x %d replicas, exchange every %d steps (BASELINE config 3)" % (nrep_all, nex),
           "nrep_all": nrep_all, "nrep_per_gpu": R, "n_gpus": world, "scaling": scaling, "md_steps": md_steps * steps,
           "us_per_md_step": 1e6 * dt / (md_steps * steps), "kernel": g.last_step_kernel(),
           "launches_rank0": g.launch_count() - l0, "rebuilds_rank0": nreb, "timers_ms": g.get_timers(),
           "exchange": {"attempts": att, "accepted": acc}}
    g.close()
    return out


def run_mixed(cx, workload, md_steps, steps, warmup):
    """the same system in the mixed-precision mode (option precision = 1: fp32 pair scalars, everything else fp64;
    stated force tolerance 2e-5, tests/test_gpu_mixed.py), device-resident loop only.  Reported beside the fp64
    headline, never instead of it."""
    spec, wl = build_system(workload)
    g = cx.new_handle()
    g.set_option("precision", 1)
    spec.configure(g)
    g.init_velocities(0)
    g.neighbor_list(0)
    for _ in range(warmup):
        g.md_steps(md_steps)
    cx.barrier()
    g.mark(0)
    nreb = 0
    for _ in range(steps):
        nreb += g.md_steps(md_steps)
    g.mark(1)
    dt = 1e-3 * g.elapsed_ms()
    out = {"metric": "particle-steps/sec", "value": spec.nmp * md_steps * steps / dt, "unit": "particle-steps/s", "dtype": "f64 + f32 pair scalars",
           "force_tolerance": 2e-5, "workload": wl, "us_per_md_step": 1e6 * dt / (md_steps * steps), "kernel": g.last_step_kernel(),
           "rebuilds": nreb}
    g.close()
    return out


def run_md(cx, workload):
    args, world, rank, torch = cx.args, cx.world, cx.rank, cx.torch
    spec, wl = build_system(workload)
    g = cx.new_handle()
    spec.configure(g)
    if args.profile:
        g.set_option("profile", 1)
    if world > 1:
        temps = [spec.tempK + 2.0 * k for k in range(world)]
        g.remd_setup([world, 1, 1, 1], [temps, [], [], []], args.md_steps, rank, world)
        g.comm_init_torch()
    g.init_velocities(0)
    g.neighbor_list(0)
    nmp = spec.nmp
    for _ in range(args.warmup):
        g.md_steps(args.md_steps)
    # the state both timed loops start from (pinned host copies)
    host0 = tuple(torch.empty((nmp, 3), dtype=torch.float64, pin_memory=True).numpy() for _ in range(3))
    x0, v0, a0 = g.get_state(0, out=host0)
    step0 = g.get_step()

    def restart():
        g.set_state(0, x0, v0, a0)
        g.set_step(step0)
        g.neighbor_list(0)

    samples, stop = [], threading.Event()
    th = threading.Thread(target=clocks_sampler, args=(stop, samples), daemon=True)
    if not os.environ.get("SIS_NO_SAMPLER"):
        th.start()

    # ---- end to end through the C ABI with host buffers (the headline) ------------------
    # every chunk: upload xyz / velos / accels, run the steps, download them + the 11 energies;
    # same start state as the loop above
    restart()
    host = tuple(torch.empty((nmp, 3), dtype=torch.float64, pin_memory=True).numpy() for _ in range(3))
    x, v, a = g.get_state(0, out=host)
    g.set_option("timing", 1)      # the per-kernel event pairs of the device-resident loop, so that both loops carry them
    cx.barrier()
    g.mark(0)
    for _ in range(args.steps):
        g.set_state(0, x, v, a)
        g.md_steps(args.md_steps)
        x, v, a = g.get_state(0, out=host)
        g.energy_sumup(0, spec.tempK, False)
    g.mark(1)
    dte = 1e-3 * g.elapsed_ms()
    cx.barrier()
    dte = cx.max_over_ranks(dte)
    e2e = world * nmp * args.md_steps * args.steps / dte
    # ---- device-resident throughput ---------------------------------------------------
    restart()
    st = list_stats(g)
    g.set_option("timing", 1)
    g.get_timers(reset=True)
    l0 = g.launch_count()
    cx.barrier()
    g.mark(0)                      # CUDA events on the library's stream: device-side time of the region
    nreb = 0
    for _ in range(args.steps):
        nreb += g.md_steps(args.md_steps)
    g.mark(1)
    dt = 1e-3 * g.elapsed_ms()
    cx.barrier()
    launches = g.launch_count() - l0
    timers = g.get_timers(reset=False)
    if args.profile and os.environ.get("SIS_CTA_PROFILE"):
        cp = g.cta_profile(148)
        np.save(os.environ["SIS_CTA_PROFILE"], cp / max(1.0, args.md_steps * args.steps))
    g.get_timers(reset=True)
    g.set_option("timing", 0)
    dt = cx.max_over_ranks(dt)
    value = world * nmp * args.md_steps * args.steps / dt

    stop.set()
    if th.is_alive():
        th.join(timeout=2)
    if os.environ.get("SIS_RANK_DEBUG"):
        print(f"[rank {rank}] value-loop {dt * 1e3 / args.steps:.2f} ms/chunk, e2e {dte * 1e3 / args.steps:.2f} ms/chunk, "
              f"timers {timers}, rebuilds {nreb}", file=sys.stderr, flush=True)

    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    hbm = peaks.get("hbm_gbs", 6650.0)
    step_ms = timers["step_ms"]
    nsteps_total = args.md_steps * args.steps
    bytes_step = algorithmic_bytes_per_step(nmp, st)
    achieved = bytes_step * nsteps_total / (step_ms * 1e-3) / 1e9 if step_ms > 0 else None
    km = kernel_metrics(workload)
    kernel = g.last_step_kernel()
    line = {
        "metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl, "md_steps_per_step": args.md_steps, "lists": st,
                   "parallelism": f"replicas x{world}" if world > 1 else "single system",
                   "l2": "state + lists are L2-resident by design (no flush between the MD steps of one trajectory); "
                         "value and e2e are timed on the same trajectory segment"},
        "e2e": {"value": e2e, "unit": "particle-steps/s", "h2d_bytes_per_step": 3 * nmp * 24,
                "d2h_bytes_per_step": 3 * nmp * 24 + 88},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s",
                     "frac": (achieved / hbm) if achieved else None, "traffic": km.get("dram_bytes_per_md_step"),
                     "kernel": kernel + " (one launch = one MD step)",
                     "algorithmic_bytes_per_launch": bytes_step,
                     "us_per_launch": 1e3 * step_ms / nsteps_total if step_ms else None,
                     "peak_source": "measured" if pk.exists() else "fallback",
                     # secondary bound: fp64 pipe utilisation of that kernel from the committed ncu capture
                     "fp64_pipe_frac": km.get("fp64_pipe_frac"), "issue_slot_frac": km.get("issue_slot_frac"),
                     "hbm_floor_us": 1e6 * bytes_step / (hbm * 1e9),
                     "ncu_kernel": km.get("kernel")},
        "clocks": summarize_clocks(samples),
        "rebuilds": nreb, "timers_ms": timers,
    }
    g.close()
    return line, spec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="sis_b200")
    ap.add_argument("--workload", default="c4", help="c4 (default), c5, c3 (T-REMD replica-steps/s)")
    ap.add_argument("--md-steps", type=int, default=1000, help="MD steps per contract step (chunk)")
    ap.add_argument("--ref-md-steps", type=int, default=100, help="MD steps per step of the reference arm")
    ap.add_argument("--cpu-baseline-steps", type=int, default=1000)
    ap.add_argument("--launch-mode", type=int, default=None, help="0 stream launches, 1 CUDA-graph chunks")
    ap.add_argument("--chunk", type=int, default=None)
    ap.add_argument("--tpb", type=int, default=None, help="lanes per bead in the pair task")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip extra.c3_remd")
    ap.add_argument("--nrep", type=int, default=64, help="replicas in total of the c3 workload")
    ap.add_argument("--opt", action="append", default=[], help="library option name=value (repeatable)")
    ap.add_argument("--profile", action="store_true", help="phase clocks of the persistent kernel")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    cx = Ctx(args)
    if args.workload == "c3":
        r = run_remd_c3(cx, nrep_all=args.nrep, md_steps=max(10, args.md_steps), steps=args.steps, warmup=args.warmup)
        if cx.rank == 0:
            line = {"metric": "replica-steps/sec", "value": r["value"], "unit": "replica-steps/s", "n_gpus": cx.world,
                    "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e-3 * r["us_per_md_step"] * max(10, args.md_steps),
                    "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                    "config": {"workload": r["workload"], "md_steps_per_step": max(10, args.md_steps)},
                    "gpu_launches": r["launches_rank0"], "detail": r}
            print(json.dumps(line))
    else:
        line, spec = run_md(cx, args.workload)
        extra = {}
        if not args.no_extra:
            try:
                extra["c3_remd"] = run_remd_c3(cx)
            except Exception as ex:  # pragma: no cover
                extra["c3_remd"] = {"error": str(ex)}
            if cx.world == 1:
                try:
                    extra["mixed_precision"] = run_mixed(cx, args.workload, args.md_steps, max(2, args.steps // 2), args.warmup)
                except Exception as ex:  # pragma: no cover
                    extra["mixed_precision"] = {"error": str(ex)}
            if cx.world == 1:        # the GPU filled with replicas (k_md_rep: 4 CTAs of 128 threads per SM): 296 temperatures x 2 ionic strengths
                try:
                    extra["c3_remd_filled"] = run_remd_c3(cx, nrep_all=592, md_steps=1000, steps=3, warmup=1)
                except Exception as ex:  # pragma: no cover
                    extra["c3_remd_filled"] = {"error": str(ex)}
            if cx.world > 1:         # the same replicas per GPU as the N = 1 line: weak scaling along replicas
                try:
                    extra["c3_remd_weak"] = run_remd_c3(cx, nrep_all=64 * cx.world, scaling="weak")
                except Exception as ex:  # pragma: no cover
                    extra["c3_remd_weak"] = {"error": str(ex)}
        if cx.rank == 0:
            line["extra"] = extra
            # CPU baseline (oracle port with the reference's OpenMP threading, bounded sample) on this box's host cores
            try:
                if args.no_cpu:
                    raise RuntimeError("skipped (--no-cpu)")
                if cx.world > 1:
                    raise RuntimeError("reported at N = 1 only")
                if args.workload == "c5":
                    raise RuntimeError("the reference cannot run C5: 2 N x N int32 maps (167 GB) and a serial "
                                       "1e10-pair neighbor_list (SURVEY 8a row NL)")
                n = args.cpu_baseline_steps if args.workload == "c4" else max(2, args.cpu_baseline_steps // 500)
                cb = cpu_sample(spec, n)
                cb["sample"] = (f"{n} MD steps of the same system (oracle port, OpenMP {cb['cores']} threads, "
                                f"{cb['rebuilds']} serial O(N^2) list rebuilds = {100 * cb['rebuild_share']:.0f} % of the time)")
                line["cpu_baseline"] = cb
            except Exception as ex:  # pragma: no cover
                line["cpu_baseline"] = {"value": None, "error": str(ex)}
            print(json.dumps(line))
    if cx.world > 1:
        cx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
