"""k_md_big, the one-launch-per-step kernel of large systems (sis_b200/csrc/sis_big.cuh): parity with
the oracle through the C ABI on small inputs forced onto it (option fast_kernel = 0), with biases,
several replicas, odd step counts and device-driven rebuilds; and at BASELINE size C5 (144 384
beads) through size-independent checks: sampled rows of the cell-grid lists against the O(N^2)
predicate of src/neighbor_list.F90:119-180, the stream-path forces against the oracle on a
256-chain box, and k_md_big against the per-phase stream path step by step."""
import numpy as np
import pytest

from oracle.oracle import Oracle
from sis_b200 import fixtures as fx
from sis_b200 import synth
from tests.test_gpu_terms import BIAS, both, rel, run_both, twz_fr_spec
from tests.test_host_parity import circ_spec

pytestmark = pytest.mark.gpu
FORCE_TOL = 1e-10


def gpu(**opts):
    from sis_b200.api import SisGpu
    return SisGpu(0, **opts)


@pytest.mark.parametrize("name", ["t2s2hp", "cag47_4_pbc", "circ", "cag47_1"])
def test_big_kernel_tracks_the_oracle_step_by_step(name):
    spec = {"t2s2hp": lambda: fx.t2s2hp("T2S2HP_folded", nstep_MC=0), "cag47_4_pbc": lambda: fx.cag47(4, pbc=True),
            "circ": circ_spec, "cag47_1": lambda: fx.cag47(1)}[name]()
    spec.nl_margin = 3.0
    g, o = both(spec, fast_kernel=0)
    g.init_velocities(0); o.init_velocities(0)
    g.neighbor_list(0); o.neighbor_list(0)
    nrg = nro = 0
    for n in (1, 2, 7, 40, 33):          # odd and even chunks: the bookkeeping alternates between S.ctl[0] and [1]
        nrg += g.md_steps(n)
        nro += o.md_run(n)
        assert g.last_step_kernel() == "k_md_big"
        xg, vg, ag = g.get_state(0)
        xo, vo, ao = o.get_state(0)
        vmax = max(1.0, np.abs(vo).max())     # the circRNA input starts with forces > 100: |v| ~ 1e3
        assert np.abs(xg - xo).max() < 1e-8 * vmax and np.abs(vg - vo).max() < 1e-9 * vmax and np.abs(ag - ao).max() < 1e-12
    assert g.get_step() == 83 and nrg == nro
    sg, eg = g.get_bp_state(0)
    so, eo = o.get_bp_state(0)
    assert np.array_equal(sg, so) and np.allclose(eg, eo, rtol=1e-7, atol=1e-10)


@pytest.mark.parametrize("name", ["stage", "twz_dcf", "twz_fr", "restraint"])
def test_big_kernel_biases(name):
    mk = twz_fr_spec if name == "twz_fr" else BIAS[name][0]
    g, o = both(mk(), fast_kernel=0)
    run_both(g, o, 25)
    assert g.last_step_kernel() == "k_md_big"


def test_big_kernel_multi_replica_and_remd():
    """several replicas in one launch (a CTA's bead range may straddle two replicas) and the
    device-side exchange step between the graph chunks"""
    temps = [290.0, 310.0, 330.0, 350.0]
    spec = fx.t2s2hp("T2S2HP_folded", nstep_MC=0).replicated(4, temps=temps)
    res = []
    for opts in (dict(launch_mode=0), dict(fast_kernel=0)):
        g = gpu(**opts)
        spec.configure(g)
        g.remd_setup([4, 1, 1, 1], [temps, [], [], []], 10)
        for r in range(4):
            g.init_velocities(r)
            g.neighbor_list(r)
        g.md_steps(60)
        res.append(([g.get_state(r)[0] for r in range(4)], g.remd_state()[0], g.remd_stats(), g.last_step_kernel()))
    assert res[1][3] == "k_md_big"
    assert np.array_equal(res[0][1], res[1][1]) and res[0][2] == res[1][2]
    for r in range(4):
        assert np.abs(res[0][0][r] - res[1][0][r]).max() < 1e-7


# ---- BASELINE size ------------------------------------------------------------------------------------
def _predicate_rows(spec, x, beads, wca_cut, ele_cut):
    """rows of `beads` by the O(N^2) predicate of src/neighbor_list.F90:119-180 (linear chains of 141, PBC)"""
    L = np.array(spec.box)
    n = len(x)
    chain, loc = np.arange(n) // 141, np.arange(n) % 141
    out = {}
    for i in beads:
        d = x[i] - x
        d = np.where(d > 0.5 * L, d - L, np.where(d < -0.5 * L, d + L, d))
        d2 = (d * d).sum(axis=1)
        same = chain == chain[i]
        sep = np.abs(loc - loc[i])
        wca = (d2 <= wca_cut ** 2) & ~(same & (sep < 3))
        ele = (d2 <= ele_cut ** 2) & ~(same & (sep <= 2))
        wca[i] = ele[i] = False
        out[i] = (set(np.nonzero(wca)[0]), set(np.nonzero(ele)[0]))
    return out


def test_c5_lists_sampled_rows_and_big_kernel_vs_stream_path():
    spec = synth.c5()
    g = gpu()
    spec.configure(g)
    g.init_velocities(0)
    g.neighbor_list(0)
    x, v, a = g.get_state(0)              # wrapped by the list build
    wca, ele, bp, _ = g.get_lists(0)
    rng = np.random.default_rng(5)
    beads = rng.choice(spec.nmp, size=2000, replace=False)
    want = _predicate_rows(spec, x, beads, spec.ff.wca_sigma + spec.nl_margin, spec.ele_cutoff + spec.nl_margin)
    for name, lst, which in (("wca", wca, 0), ("ele", ele, 1)):
        lst = lst - 1
        order = np.argsort(lst[:, 0], kind="stable")
        lo = np.searchsorted(lst[order, 0], beads, side="left")
        hi = np.searchsorted(lst[order, 0], beads, side="right")
        order2 = np.argsort(lst[:, 1], kind="stable")
        lo2 = np.searchsorted(lst[order2, 1], beads, side="left")
        hi2 = np.searchsorted(lst[order2, 1], beads, side="right")
        for q, i in enumerate(beads):
            got = set(lst[order[lo[q]:hi[q]], 1]) | set(lst[order2[lo2[q]:hi2[q]], 0])
            assert got == want[i][which], (name, i)
    # k_md_big against the per-phase stream kernels, same start: a few steps, then across a rebuild
    s = gpu(launch_mode=0)
    spec.configure(s)
    s.set_state(0, x, v, a)
    s.neighbor_list(0)
    assert s.get_list_sizes(0) == g.get_list_sizes(0)
    for n in (3, 80):
        rg, rs = g.md_steps(n), s.md_steps(n)
        assert g.last_step_kernel() == "k_md_big" and rg == rs
        xg, vg, ag = g.get_state(0)
        xs, vs, as_ = s.get_state(0)
        assert np.abs(xg - xs).max() < 1e-8 and np.abs(vg - vs).max() < 1e-8 and np.abs(ag - as_).max() < 1e-12
    assert rg >= 1 and g.get_step() == 83
    f = g.force(0)
    assert np.isfinite(f).all() and np.abs(f.sum(axis=0)).max() < 1e-6


def test_256_chain_box_forces_vs_oracle():
    """36 096 beads (a quarter of C5, same density): cell-grid lists as index sets and per-term forces /
    energies against the oracle's O(N^2) list and serial loops"""
    spec = synth.cag47_box(256, 1285.7, seed=3)
    g, o = both(spec)
    g.neighbor_list(0); o.neighbor_list(0)
    assert g.get_list_sizes(0) == o.get_list_sizes(0)
    for a, b in zip(g.get_lists(0)[:3], o.get_lists(0)[:3]):
        assert np.array_equal(a, b)
    for t in (1, 2, 4, 5, 6):
        assert rel(g.force_terms(0, 1 << t), o.force_terms(0, 1 << t)) < FORCE_TOL
    assert rel(g.force(0), o.force(0)) < FORCE_TOL
    assert np.allclose(g.energy_sumup(0, spec.tempK, True), o.energy_sumup(0, spec.tempK, True), rtol=1e-10, atol=1e-8)
    # and the big kernel on it, a few steps against the oracle
    gb = gpu(fast_kernel=0)
    spec.configure(gb)
    gb.init_velocities(0); o.init_velocities(0)
    gb.neighbor_list(0)
    gb.md_steps(4); o.md_run(4)
    assert gb.last_step_kernel() == "k_md_big"
    assert np.abs(gb.get_state(0)[0] - o.get_state(0)[0]).max() < 1e-9
