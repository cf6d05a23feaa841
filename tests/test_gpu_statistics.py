"""Long-run statistical parity (north_star check 3): the CUDA path with its own seed against the
committed statistics of the CPU oracle (tests/golden/stat_fixture.json, written by
tests/golden/make_stat_fixture.py with the protocol imported below): potential-energy distribution
and its terms, number of recorded base pairs, per-pair REMD acceptance ratios, mean energy and
base-pair count per temperature.  Different RNG streams, 2.2e5 steps each: agreement within
4 standard errors of the difference (block-averaged; binomial for the acceptance ratios)."""
import importlib.util
import json
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"
NSIGMA = 4.0


def protocol():
    spec = importlib.util.spec_from_file_location("make_stat_fixture", GOLDEN / "make_stat_fixture.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def gpu():
    from sis_b200.api import SisGpu
    return SisGpu(0)


def compare(got, ref, keys):
    for k in keys:
        g, r = got[k], ref[k]
        se = np.hypot(g["se"], r["se"])
        # a quantity that never fluctuated in either run (e.g. a base-pair count pinned at its maximum)
        se = max(se, 1e-3 * max(1.0, abs(r["mean"])))
        assert abs(g["mean"] - r["mean"]) <= NSIGMA * se, (k, g, r)
        if r["std"] > 10 * se and g["n"] >= 1000:          # width of the distribution, where it is resolved
            assert 0.8 < g["std"] / r["std"] < 1.25, (k, g["std"], r["std"])


def test_c1_energy_distribution_and_base_pair_count():
    m = protocol()
    ref = json.loads((GOLDEN / "stat_fixture.json").read_text())
    g = gpu()
    got = m.run_c1(g, seed=777)
    assert g.last_step_kernel() == "k_md_fast<1,true>"
    compare(got, ref["c1"], ["E_pot", "E_bp", "E_ele", "E_bond", "n_bp"])


def test_remd_acceptance_ratios_and_per_temperature_means():
    m = protocol()
    ref = json.loads((GOLDEN / "stat_fixture.json").read_text())["remd"]
    g = gpu()
    got = m.run_remd(g, seed=777, oracle=False)
    att_g, acc_g = np.array(got["acceptance"]["attempts"]), np.array(got["acceptance"]["accepted"])
    att_r, acc_r = np.array(ref["acceptance"]["attempts"]), np.array(ref["acceptance"]["accepted"])
    assert (att_g == att_r).all() and att_g.min() >= 5000
    pg, pr = acc_g / att_g, acc_r / att_r
    # successive attempts of a pair are correlated (the energies relax over ~100 steps): allow 3x the binomial error
    se = 3.0 * np.sqrt(pg * (1 - pg) / att_g + pr * (1 - pr) / att_r)
    assert (np.abs(pg - pr) <= NSIGMA * se).all(), (pg, pr, se)
    assert 0.5 < pg.min() and pg.max() < 0.9
    n = len(m.REMD_T)
    compare(got, ref, [f"E_pot_T{k}" for k in range(n)] + [f"n_bp_T{k}" for k in range(n)])
    # hotter labels have higher mean potential energy
    means = [got[f"E_pot_T{k}"]["mean"] for k in range(n)]
    assert means[-1] > means[0]
