"""GPU parity tests proper: the CUDA path, through the C ABI, against the CPU oracle on the same
seeded inputs (bit-exact lists, 1e-10 forces / energies), plus size-independent properties at
the benchmark sizes."""
import numpy as np
import pytest

from oracle.oracle import Oracle
from sis_b200 import fixtures as fx
from sis_b200 import synth
from tests.test_host_parity import SPECS, TERMS, pairset

pytestmark = pytest.mark.gpu

FORCE_TOL = 1e-10   # norm-relative, fp64 mode (north_star)


def gpu():
    from sis_b200.api import SisGpu
    return SisGpu(0)


def both(spec):
    g, o = gpu(), Oracle()
    spec.configure(g)
    spec.configure(o)
    return g, o


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-12)


@pytest.mark.parametrize("name", list(SPECS))
@pytest.mark.parametrize("full", [False, True])
def test_lists_bit_exact(name, full):
    spec = SPECS[name]()
    g, o = both(spec)
    if full:
        g.full_lists(); o.full_lists()
    else:
        g.neighbor_list(0); o.neighbor_list(0)
    wg, eg, bg, cg = g.get_lists(0)
    wo, eo, bo, co = o.get_lists(0)
    # identical index sets AND identical order (the reference's loop order)
    assert np.array_equal(wg, wo)
    assert np.array_equal(eg, eo)
    assert np.array_equal(bg, bo)
    assert np.array_equal(cg, co)
    xg, _, _ = g.get_state(0)
    xo, _, _ = o.get_state(0)
    assert np.array_equal(xg, xo)   # pbc_wrap is bit-exact


@pytest.mark.parametrize("name", list(SPECS))
def test_forces_and_energies(name):
    spec = SPECS[name]()
    g, o = both(spec)
    g.neighbor_list(0); o.neighbor_list(0)
    for term, bit in TERMS.items():
        assert rel(g.force_terms(0, 1 << bit), o.force_terms(0, 1 << bit)) < FORCE_TOL, term
    assert rel(g.force(0), o.force(0)) < FORCE_TOL
    eg = g.energy_sumup(0, spec.tempK, True)
    eo = o.energy_sumup(0, spec.tempK, True)
    assert np.allclose(eg, eo, rtol=1e-10, atol=1e-9), (eg, eo)
    # cached bp energies from the force pass (flg_bp_energy = .True.)
    g.force(0); o.force(0)
    assert np.allclose(g.energy_sumup(0, spec.tempK, False), o.energy_sumup(0, spec.tempK, False), rtol=1e-10, atol=1e-9)
    sg, ng = g.get_bp_state(0)
    so, no = o.get_bp_state(0)
    assert np.array_equal(sg, so)
    assert np.allclose(ng, no, rtol=1e-10, atol=1e-12)


def test_c4_lists_forces_vs_oracle():
    """64 x CAG47 in the 809.972 A box (BASELINE config 4): cell-grid lists == O(N^2) oracle."""
    spec = synth.c4()
    g, o = both(spec)
    g.neighbor_list(0); o.neighbor_list(0)
    assert g.get_list_sizes(0) == o.get_list_sizes(0)
    wg, eg, bg, cg = g.get_lists(0)
    wo, eo, bo, co = o.get_lists(0)
    assert np.array_equal(wg, wo) and np.array_equal(eg, eo) and np.array_equal(bg, bo)
    assert rel(g.force(0), o.force(0)) < FORCE_TOL
    assert np.allclose(g.energy_sumup(0, spec.tempK, True), o.energy_sumup(0, spec.tempK, True), rtol=1e-10, atol=1e-8)


@pytest.mark.parametrize("name", ["t2s2hp", "cag47_4_pbc"])
def test_trajectory_matches_oracle_short(name):
    """Same counter-based Gaussians on both sides: the GJF-2GJ update (src/job_md.F90:415-444) must
    track the oracle over a short run (divergence grows chaotically, so the window is short)."""
    spec = SPECS[name]()
    g, o = both(spec)
    g.init_velocities(0); o.init_velocities(0)
    xg, vg, ag = g.get_state(0)
    xo, vo, ao = o.get_state(0)
    assert np.allclose(vg, vo, rtol=1e-12, atol=1e-14) and np.allclose(ag, ao, rtol=1e-12, atol=1e-16)
    g.neighbor_list(0); o.neighbor_list(0)
    g.md_steps(25)
    o.md_run(25)
    xg, vg, ag = g.get_state(0)
    xo, vo, ao = o.get_state(0)
    assert np.abs(xg - xo).max() < 1e-8
    assert np.abs(vg - vo).max() < 1e-9
    assert g.get_step() == 25


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_launch_modes_track_oracle(mode):
    """stream launches, CUDA-graph chunks and the persistent cooperative kernel run the same step."""
    from sis_b200.api import SisGpu
    spec = fx.cag47(4, pbc=True)
    spec.nl_margin = 2.0
    g, o = SisGpu(0, launch_mode=mode, chunk=8), Oracle()
    spec.configure(g); spec.configure(o)
    g.init_velocities(0); o.init_velocities(0)
    g.neighbor_list(0); o.neighbor_list(0)
    nreb_g = g.md_steps(100)
    nreb_o = o.md_run(100)
    assert g.get_step() == 100
    assert nreb_g == nreb_o and nreb_g >= 1
    xg, vg, _ = g.get_state(0)
    xo, vo, _ = o.get_state(0)
    assert np.abs(xg - xo).max() < 1e-6


def test_rebuild_trigger_matches_oracle():
    """Displacement top-2 test (src/job_md.F90:376-395): same rebuild count as the oracle over a
    run long enough to cross the margin several times (small margin to force rebuilds)."""
    spec = fx.cag47(4, pbc=True)
    spec.nl_margin = 1.0
    g, o = both(spec)
    g.init_velocities(0); o.init_velocities(0)
    g.neighbor_list(0); o.neighbor_list(0)
    nreb_g = g.md_steps(120)
    nreb_o = o.md_run(120)
    assert nreb_g == nreb_o and nreb_g >= 2
    xg, _, _ = g.get_state(0)
    xo, _, _ = o.get_state(0)
    assert np.abs(xg - xo).max() < 1e-6


def test_check_force_job():
    """job_check_force through the C ABI: |f_energy - f_force| <= 1e-3 (src/job_check_force.F90:21,78)."""
    spec = fx.t2s2hp("T2S2HP_folded", nstep_MC=0)
    rng = np.random.default_rng(5)
    spec.xyz = spec.xyz + 0.001 * rng.random(spec.xyz.shape)
    g, o = both(spec)
    g.full_lists(); o.full_lists()
    fe, ff, md = g.check_force(spec.tempK)
    feo, ffo, mdo = o.check_force(spec.tempK)
    assert md <= 1e-3
    assert rel(ff, ffo) < FORCE_TOL
    assert np.abs(fe - feo).max() < 1e-6


def test_c4_properties_long_run():
    """Benchmark-size properties: finite coordinates, momentum-free forces, bounded bonds, lists
    stay supersets of the exact cutoff sets after rebuild-driven dynamics."""
    spec = synth.c4()
    g = gpu()
    spec.configure(g)
    g.init_velocities(0)
    g.neighbor_list(0)
    nreb = g.md_steps(600)
    x, v, a = g.get_state(0)
    assert np.isfinite(x).all() and np.isfinite(v).all()
    f = g.force(0)
    assert np.abs(f.sum(axis=0)).max() < 1e-7
    dv = np.diff(x.reshape(64, 141, 3), axis=1)
    dv -= 809.972 * np.round(dv / 809.972)      # minimum image: chains are wrapped bead by bead
    d = np.linalg.norm(dv, axis=2)
    assert d.min() > 4.0 and d.max() < 8.0
    # temperature from kinetic energy (GJF-2GJ half-step velocity): within 10 % of 310.15 K
    ek = g.energy_kinetic(0)
    T = 2.0 * ek / (3 * spec.nmp * 0.0019872042586408316)
    assert abs(T / spec.tempK - 1.0) < 0.1, T
    # after the run: rebuild on both sides and compare exactly
    o = Oracle()
    spec.configure(o)
    o.set_state(0, x)
    g.neighbor_list(0); o.neighbor_list(0)
    assert g.get_list_sizes(0) == o.get_list_sizes(0)
    assert rel(g.force(0), o.force(0)) < FORCE_TOL
    assert nreb >= 1


def _run_mode(spec, nsteps, **opts):
    from sis_b200.api import SisGpu
    g = SisGpu(0, **opts)
    spec.configure(g)
    g.init_velocities(0)
    g.neighbor_list(0)
    nreb = g.md_steps(nsteps)
    x, v, a = g.get_state(0)
    return g, nreb, x, v, a


@pytest.mark.parametrize("name", ["t2s2hp", "cag47_4_pbc", "circ"])
def test_persistent_kernels_match_stream_path(name):
    """k_md_fast (on-chip rows / tables, role-split base pairs) and the general k_md_fused run the
    same trajectory as the per-phase stream launches (dihedral tables, circRNA closure terms and
    PBC image flags included)."""
    spec = SPECS[name]()
    spec.nl_margin = 3.0
    _, r0, x0, v0, a0 = _run_mode(spec, 60, launch_mode=0)
    g1, r1, x1, v1, a1 = _run_mode(spec, 60, launch_mode=2, fast_kernel=1)
    g2, r2, x2, v2, a2 = _run_mode(spec, 60, launch_mode=2, fast_kernel=0)                  # k_md_big
    g3, r3, x3, v3, a3 = _run_mode(spec, 60, launch_mode=2, fast_kernel=0, big_kernel=0)    # k_md_fused
    assert g1.last_step_kernel().startswith(("k_md_fast", "k_md_rep")) and g2.last_step_kernel() == "k_md_big"
    assert g3.last_step_kernel() == "k_md_fused"
    assert r0 == r1 == r2 == r3
    for x, v, a in ((x1, v1, a1), (x2, v2, a2), (x3, v3, a3)):
        assert np.abs(x - x0).max() < 1e-7
        assert np.abs(v - v0).max() < 1e-7   # rounding differences grow along the trajectory
        assert np.abs(a - a0).max() < 1e-12
    assert g2.get_step() == 60 and g3.get_step() == 60
    s_big, e_big = g2.get_bp_state(0)
    assert g1.get_step() == 60
    # cached bp state of the last force evaluation (write_bp / flg_bp_energy semantics)
    s_ref, e_ref = _run_mode(spec, 60, launch_mode=0)[0].get_bp_state(0)
    s_fast, e_fast = g1.get_bp_state(0)
    assert np.array_equal(s_ref, s_fast) and np.array_equal(s_ref, s_big)
    assert np.allclose(e_ref, e_fast, rtol=1e-7, atol=1e-10)
    assert np.allclose(e_ref, e_big, rtol=1e-7, atol=1e-10)


def test_persistent_kernel_multi_replica():
    """Several replicas batched into one persistent launch give the stream-path trajectories."""
    from sis_b200.api import SisGpu
    spec = fx.t2s2hp("T2S2HP_folded", nstep_MC=0).replicated(4, temps=[290.0, 300.0, 310.0, 320.0])
    outs = []
    for mode in (0, 2):
        g = SisGpu(0, launch_mode=mode)
        spec.configure(g)
        for r in range(4):
            g.init_velocities(r)
            g.neighbor_list(r)
        g.md_steps(80)
        outs.append([g.get_state(r)[0] for r in range(4)])
    for r in range(4):
        assert np.abs(outs[0][r] - outs[1][r]).max() < 1e-7
    assert np.abs(outs[0][0] - outs[0][3]).max() > 1e-3   # the replicas really differ


def test_replica_local_kernel_rebuilds_in_place():
    """Replica-local mode of k_md_fast (one CTA per replica, lists rebuilt inside the kernel, no
    grid barrier): same trajectories, same number of rebuilds (src/job_md.F90:388-395) and the
    same lists as the stream path with host-driven rebuilds, replicas rebuilding at different steps."""
    from sis_b200.api import SisGpu
    temps = [290.0, 320.0, 350.0, 380.0]
    spec = fx.t2s2hp("T2S2HP_folded", nstep_MC=0).replicated(4, temps=temps)
    spec.nl_margin = 1.5
    res = []
    for opts in (dict(launch_mode=0), dict(launch_mode=2, local_kernel=1), dict(launch_mode=2, local_kernel=0)):
        g = SisGpu(0, **opts)
        spec.configure(g)
        for r in range(4):
            g.init_velocities(r)
            g.neighbor_list(r)
        nreb = g.md_steps(150)
        nreb += g.md_steps(50)
        res.append((nreb, [g.get_state(r) for r in range(4)], [g.get_lists(r) for r in range(4)], g.get_step()))
    assert res[0][0] == res[1][0] == res[2][0] and res[0][0] >= 4
    assert res[0][3] == res[1][3] == 200
    for other in (res[1], res[2]):
        for r in range(4):
            for a, b in zip(res[0][1][r], other[1][r]):
                assert np.abs(a - b).max() < 1e-6
            for la, lb in zip(res[0][2][r][:3], other[2][r][:3]):   # wca, ele, bp as index sets
                ka = np.lexsort(la.T[::-1]) if len(la) else []
                kb = np.lexsort(lb.T[::-1]) if len(lb) else []
                assert np.array_equal(la[ka], lb[kb])


def test_device_side_rebuild_recovers_from_invalid_lists():
    """A device-side rebuild that cannot fit its rows marks the lists invalid: the persistent kernels
    enqueued behind it fall through and the host grows the capacities, rebuilds every replica and
    carries on.  The recovery is exercised with the `inject_invalid` test hook (a real overflow
    needs a collapsing system); the trajectory must not notice."""
    from sis_b200.api import SisGpu
    spec = SPECS["cag47_4_pbc"]()
    spec.nl_margin = 3.0
    out = []
    for inject in (0, 2):
        g = SisGpu(0)
        spec.configure(g)
        g.init_velocities(0)
        g.neighbor_list(0)
        if inject:
            g.set_option("inject_invalid", inject)
        nreb = g.md_steps(120)
        out.append((g.get_state(0), nreb, g.get_step()))
    assert out[0][2] == out[1][2] == 120
    assert out[1][1] >= 2
    for a, b in zip(out[0][0], out[1][0]):
        assert np.abs(a - b).max() < 1e-7


def test_row_overflow_during_device_side_rebuilds_grows_the_lists():
    """Rows that outgrow their capacity at a device-side rebuild (x1.2 growth of
    src/neighbor_list.F90:247-298): with a capacity equal to the longest row of the first build the
    following rebuilds overflow sooner or later; the run must equal the one with ample capacity."""
    from sis_b200.api import SisGpu
    spec = SPECS["cag47_4_pbc"]()
    spec.nl_margin = 3.0
    g0 = SisGpu(0)
    spec.configure(g0)
    g0.neighbor_list(0)
    wca, ele, bp, _ = g0.get_lists(0)
    both = np.unique(np.concatenate([wca, ele]), axis=0)          # a pair listed for WCA and DH is one row entry
    longest = np.bincount(both.ravel() - 1, minlength=spec.nmp).max()
    out = []
    for cap in (0, int(longest)):
        g = SisGpu(0)
        spec.configure(g)
        if cap:
            g.set_option("list_capacity", cap)
        g.init_velocities(0)
        g.neighbor_list(0)
        nreb = g.md_steps(400)
        out.append((g.get_state(0), nreb, g.get_step(), g.get_list_sizes(0)))
    assert out[0][2] == out[1][2] == 400 and out[1][1] >= 10
    assert out[0][3] == out[1][3]
    for a, b in zip(out[0][0], out[1][0]):
        assert np.abs(a - b).max() < 1e-6


def test_gpu_dynamics_match_the_reference_trajectory_statistically():
    """Same distributional check as tests/test_oracle_golden.py against the reference binary's own
    trajectory check/test_md.dcd, for the persistent CUDA path (north_star: trajectories agree
    statistically since the RNG streams differ)."""
    from tests.test_oracle_golden import chain_statistics
    from sis_b200.api import SisGpu
    ref = np.array([chain_statistics(f) for f in fx.array("test_md_dcd")])
    spec = fx.cag47(4, pbc=True)
    g = SisGpu(0)
    spec.configure(g)
    g.init_velocities(0)
    g.neighbor_list(0)
    g.md_steps(3000)
    ours = []
    for _ in range(40):
        g.md_steps(200)
        ours.append(chain_statistics(g.get_state(0)[0]))
    ours = np.array(ours)
    assert abs(ours[:, 0].mean() - ref[:, 0].mean()) < 0.03
    assert abs(ours[:, 1].mean() - ref[:, 1].mean()) < 0.008
    assert abs(ours[:, 2].mean() - ref[-4:, 2].mean()) < 0.04
    ek = g.energy_kinetic(0)
    assert abs(2.0 * ek / (3 * spec.nmp * 0.0019872042586408316) / spec.tempK - 1.0) < 0.1
