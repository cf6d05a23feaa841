"""Mixed-precision mode (option "precision" = 1, north_star: "stated looser bound in mixed mode"): the WCA /
Debye-Hueckel pair scalars are evaluated in fp32 from the fp64 squared distance; pair vectors, sums and every
other term stay fp64.  Stated bound: every force component within MIXED_TOL of the largest force component of
the fp64 oracle, per term and in total; everything that is not a pair scalar keeps the fp64 bound."""
import numpy as np
import pytest

from oracle.oracle import Oracle
from sis_b200 import fixtures as fx
from sis_b200 import synth
from tests.test_host_parity import SPECS, TERMS

pytestmark = pytest.mark.gpu

MIXED_TOL = 2e-5      # norm-relative, mixed mode
FP64_TOL = 1e-10      # the terms the mode does not touch


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-12)


def pair(spec):
    from sis_b200.api import SisGpu
    g, o = SisGpu(0, precision=1), Oracle()
    spec.configure(g)
    spec.configure(o)
    g.neighbor_list(0); o.neighbor_list(0)
    return g, o


@pytest.mark.parametrize("name", ["t2s2hp", "cag47_4_pbc", "circ"])
def test_mixed_mode_forces_within_the_stated_bound(name):
    g, o = pair(SPECS[name]())
    errs = {}
    for term, bit in TERMS.items():
        fg, fo = g.force_terms(0, 1 << bit), o.force_terms(0, 1 << bit)
        errs[term] = rel(fg, fo)
        assert errs[term] < (MIXED_TOL if term in ("exv", "ele") else FP64_TOL), (term, errs)
    assert rel(g.force(0), o.force(0)) < MIXED_TOL
    # the mode is really on: the pair terms are NOT at the fp64 level
    assert max(errs["exv"], errs["ele"]) > 1e-9, errs


@pytest.mark.parametrize("name,kernel", [("t2s2hp", "k_md_rep"), ("cag47_4_pbc", "k_md_fast"), ("cag47_4_pbc", "k_md_big")])
def test_mixed_mode_short_trajectory_stays_close_to_fp64(name, kernel):
    """20 steps of each step kernel in mixed mode against the same kernel in fp64: the positions differ by the
    accumulated fp32 rounding of the pair forces only (no list or image decision may flip in 20 steps)"""
    from sis_b200.api import SisGpu
    from tests.conftest import is_replica_local
    spec = SPECS[name]()
    xs = []
    for prec in (0, 1):
        opts = dict(precision=prec)
        if kernel == "k_md_big":
            opts["fast_kernel"] = 0
        g = SisGpu(0, **opts)
        spec.configure(g)
        g.init_velocities(0)
        g.neighbor_list(0)
        g.md_steps(20)
        k = g.last_step_kernel()
        assert k.startswith(kernel) or (kernel == "k_md_rep" and (is_replica_local(k) or k.startswith("k_md_fast"))), k
        xs.append(g.get_state(0)[0])
    d = np.abs(xs[0] - xs[1]).max()
    assert 0.0 < d < 1e-4, d


def test_mixed_mode_at_c4_size():
    """per-term bound on the benchmark system (64 x CAG47, PBC): pair terms of 9 024 beads against the oracle"""
    spec = synth.c4()
    g, o = pair(spec)
    for term in ("exv", "ele"):
        fg, fo = g.force_terms(0, 1 << TERMS[term]), o.force_terms(0, 1 << TERMS[term])
        assert rel(fg, fo) < MIXED_TOL, term
