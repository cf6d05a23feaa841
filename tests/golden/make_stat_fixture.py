"""Long-run statistics of the CPU oracle for the statistical parity test of the CUDA path
(north_star check 3: "trajectories agree statistically ... in energy distributions, base-pair
fraction and REMD acceptance ratios over long runs, since RNG streams differ").

Writes tests/golden/stat_fixture.json.  Run once in the build container:
    python tests/golden/make_stat_fixture.py
The protocol (shared with tests/test_gpu_statistics.py through `protocol()` / `summarise()`):
  C1   check/input_T2S2HP.toml (27-nt hairpin, model 5, DH, Monte-Carlo base-pair limit every step),
       folded start, 20 000 equilibration steps, then 200 000 steps sampled every 100:
       E_pot and its terms (energy_sumup with the cached bp energies, as write_out), number of base
       pairs write_bp would record (bp_status and ene_bp <= -kT);
  REMD 8 temperatures 293.15 .. 363.15 K (check/T2S2HP_REMD_T/input_REMD.toml), exchange every 10
       steps, nstep_MC 0: acceptance ratio of every neighbouring label pair, mean E_pot and mean
       base-pair count per temperature label.
The hairpin has slow modes (an extra terminal base pair that lives for ~1e5 steps at the lowest
temperatures), so a single run's block-averaged error underestimates the true one: the fixture holds
the oracle's summaries for SEVERAL seeds, and the test compares the CUDA runs (other seeds) with the
oracle's mean over seeds using the seed-to-seed spread.  Only distributions are compared."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from sis_b200 import fixtures as fx  # noqa: E402

KB = 0.0019872042586408316
NEQ, NPROD, NSAMP = 20000, 200000, 100
NPROD_REMD = 100000
ORACLE_SEEDS = [123, 777, 4242, 9001]
REMD_T = [293.15 + 10.0 * k for k in range(8)]
NEX = 10


def block_se(x, nblock=20):
    x = np.asarray(x, dtype=float)
    n = len(x) // nblock * nblock
    b = x[:n].reshape(nblock, -1).mean(axis=1)
    return float(b.std(ddof=1) / np.sqrt(nblock))


def summarise(samples):
    out = {}
    for k, v in samples.items():
        v = np.asarray(v, dtype=float)
        out[k] = {"mean": float(v.mean()), "std": float(v.std(ddof=1)), "se": block_se(v), "n": int(len(v))}
    return out


def run_c1(be, seed):
    spec = fx.t2s2hp("T2S2HP_folded", nstep_MC=1)
    spec.rng_seed = seed
    spec.configure(be)
    be.init_velocities(0)
    be.neighbor_list(0)
    step = getattr(be, "md_steps", None) or be.md_run
    step(NEQ)
    kT = spec.tempK * KB
    s = {"E_pot": [], "E_bp": [], "E_ele": [], "E_bond": [], "n_bp": []}
    for _ in range(NPROD // NSAMP):
        step(NSAMP)
        e = be.energy_sumup(0, spec.tempK, False)
        st, en = be.get_bp_state(0)
        s["E_pot"].append(e[0]); s["E_bp"].append(e[4]); s["E_ele"].append(e[6]); s["E_bond"].append(e[1])
        s["n_bp"].append(int(((st != 0) & (en <= -kT)).sum()))
    return summarise(s)


def run_remd(be, seed, oracle):
    n = len(REMD_T)
    spec = fx.t2s2hp("T2S2HP_folded", nstep_MC=0).replicated(n, temps=REMD_T)
    spec.rng_seed = seed
    spec.configure(be)
    be.remd_setup([n, 1, 1, 1], [REMD_T, [], [], []], NEX)
    if oracle:
        be.set_save_bp(1)     # every exchange step taken as a save step: bp_status / ene_bp restored (src/energy_replica.F90:64-70,157-162)
    for r in range(n):
        be.init_velocities(r)
        be.neighbor_list(r)
    counter = 0
    att = np.zeros(n - 1)
    acc = np.zeros(n - 1)
    s = {f"E_pot_T{k}": [] for k in range(n)}
    s.update({f"n_bp_T{k}": [] for k in range(n)})
    ntot = (NEQ + NPROD_REMD) // NEX
    for it in range(ntot):
        before = be.remd_state()[0].copy()
        if oracle:
            be.md_run(NEX)
            _, re = be.energy_replica()
            u = np.array([be.uniform(seed, 0x7fffffff, counter + i) for i in range((n + 1) // 2)])
            counter += (n + 1) // 2
            be.replica_exchange(re, u)
        else:
            be.md_steps(NEX)
        after = be.remd_state()[0]
        if (it + 1) * NEX <= NEQ:
            continue
        # attempts of this round: odd / even neighbour pairs alternate (src/var_replica.F90:162-259), first round odd
        first = 1 if it % 2 == 0 else 2
        lab_of = {int(l): r for r, l in enumerate(before)}
        for l in range(first, n, 2):
            att[l - 1] += 1
            if after[lab_of[l]] == l + 1:
                acc[l - 1] += 1
        if ((it + 1) * NEX) % NSAMP == 0:
            for r in range(n):
                lab = int(after[r]) - 1
                T = REMD_T[lab]
                e = be.energy_sumup(r, T, True)       # bp term re-evaluated at the label's temperature on both sides
                st, en = be.get_bp_state(r)
                s[f"E_pot_T{lab}"].append(e[0])
                s[f"n_bp_T{lab}"].append(int(((st != 0) & (en <= -T * KB)).sum()))
    out = summarise(s)
    out["acceptance"] = {"attempts": att.tolist(), "accepted": acc.tolist()}
    return out


def main():
    from oracle.oracle import Oracle
    res = {"protocol": {"neq": NEQ, "nprod": NPROD, "nprod_remd": NPROD_REMD, "nsamp": NSAMP, "remd_T": REMD_T, "nex": NEX,
                        "seeds": ORACLE_SEEDS},
           "c1": {}, "remd": {}}
    for seed in ORACLE_SEEDS:
        res["c1"][str(seed)] = run_c1(Oracle(), seed)
        res["remd"][str(seed)] = run_remd(Oracle(), seed, True)
        print(seed, {k: round(v["mean"], 3) for k, v in res["c1"][str(seed)].items()}, res["remd"][str(seed)]["acceptance"], flush=True)
    (Path(__file__).resolve().parent / "stat_fixture.json").write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
