"""Round-2 REMD / restart checks on the GPU: the exchange step inside the replica-local persistent
kernel (fast_remd_exchange) against the device-side launch chain and against the oracle, the base-pair
state around an exchange step that is also a save step (src/energy_replica.F90:64-70,157-162), and the
restart payload of the counter-based generators."""
import numpy as np
import pytest

from tests.conftest import is_replica_local

from oracle.oracle import Oracle
from sis_b200 import fixtures as fx

pytestmark = pytest.mark.gpu


def gpu(**opts):
    from sis_b200.api import SisGpu
    return SisGpu(0, **opts)


def remd_gpu(spec, n, temps, nex, **opts):
    g = gpu(**opts)
    spec.configure(g)
    g.remd_setup([n, 1, 1, 1], [temps, [], [], []], nex)
    for r in range(n):
        g.init_velocities(r)
        g.neighbor_list(r)
    return g


@pytest.mark.parametrize("n", [4, 7, 16])
def test_inkernel_exchange_equals_the_launch_chain(n):
    """one launch per md_steps call with the exchange inside the kernel == one launch per exchange
    interval + the exchange chain, across launch boundaries that do and do not fall on exchange steps, with
    in-kernel list rebuilds in between.  Same energies bit for bit (same summation order); the trajectories
    agree to rounding only: every launch re-partitions the on-chip rows by distance, so the two paths add the
    pair forces in different orders"""
    temps = list(np.linspace(290.0, 380.0, n))
    spec = fx.t2s2hp("T2S2HP_folded", nstep_MC=0).replicated(n, temps=temps)
    spec.nl_margin = 1.0
    res = []
    for ink in (1, 0):
        g = remd_gpu(spec, n, temps, 10, remd_inkernel=ink)
        nreb = 0
        for nsteps in (25, 15, 40, 7, 13):
            nreb += g.md_steps(nsteps)
        assert is_replica_local(g.last_step_kernel()) and g.get_step() == 100
        res.append(([g.get_state(r) for r in range(n)], g.remd_state(), g.remd_stats(), nreb,
                    [g.get_bp_state(r) for r in range(n)], g.launch_count()))
    a, b = res
    assert a[2] == b[2] and a[2][0] >= 10 * ((n - 1) // 2) and a[2][1] > 0   # odd / even pair tables alternate
    assert np.array_equal(a[1][0], b[1][0]) and np.array_equal(a[1][1], b[1][1])
    assert a[3] == b[3]
    for r in range(n):
        for u, v in zip(a[0][r], b[0][r]):
            assert np.abs(u - v).max() < 1e-9
        assert np.array_equal(a[4][r][0], b[4][r][0]) and np.allclose(a[4][r][1], b[4][r][1], rtol=1e-9, atol=1e-12)
    assert a[5] < b[5]          # far fewer launches


def test_inkernel_exchange_follows_the_oracle():
    n, nex = 8, 10
    temps = list(np.linspace(293.15, 363.15, n))
    spec = fx.t2s2hp("T2S2HP_folded", nstep_MC=0).replicated(n, temps=temps)
    g = remd_gpu(spec, n, temps, nex)
    o = Oracle()
    spec.configure(o)
    o.remd_setup([n, 1, 1, 1], [temps, [], [], []], nex)
    o.set_save_bp(1)                  # every exchange step taken as a save step (src/energy_replica.F90:64-70,157-162)
    for r in range(n):
        o.init_velocities(r)
        o.neighbor_list(r)
    g.md_steps(60)
    counter = 0
    for _ in range(6):
        o.md_run(nex)
        _, re = o.energy_replica()
        u = np.array([o.uniform(spec.rng_seed, 0x7fffffff, counter + i) for i in range(n // 2)])
        counter += n // 2
        o.replica_exchange(re, u)
    assert np.array_equal(g.remd_state()[0], o.remd_state()[0])
    for r in range(n):
        xg, vg, _ = g.get_state(r)
        xo, vo, _ = o.get_state(r)
        assert np.abs(xg - xo).max() < 1e-6 and np.abs(vg - vo).max() < 1e-6
        # bp_status / ene_bp after an exchange step are the replica's own (src/energy_replica.F90:157-162)
        sg, eg = g.get_bp_state(r)
        so, eo = o.get_bp_state(r)
        assert np.array_equal(sg, so) and np.allclose(eg, eo, rtol=1e-7, atol=1e-10)


@pytest.mark.parametrize("opts", [dict(remd_inkernel=0), dict(launch_mode=0)], ids=["device-chain", "host-driven"])
def test_cross_energy_pass_leaves_the_own_base_pair_state(opts):
    """the partner-parameter pass of the exchange step must not leave its bp_status / ene_bp
    behind (a save step that is also an exchange step writes them)"""
    n, nex = 4, 10
    temps = [290.0, 320.0, 350.0, 380.0]
    spec = fx.t2s2hp("T2S2HP_folded", nstep_MC=0).replicated(n, temps=temps)
    g = remd_gpu(spec, n, temps, nex, **opts)
    o = Oracle()
    spec.configure(o)
    o.remd_setup([n, 1, 1, 1], [temps, [], [], []], nex)
    o.set_save_bp(1)
    for r in range(n):
        o.init_velocities(r)
        o.neighbor_list(r)
    g.md_steps(nex)
    o.md_run(nex)
    _, re = o.energy_replica()
    o.replica_exchange(re, np.array([o.uniform(spec.rng_seed, 0x7fffffff, i) for i in range(n // 2)]))
    for r in range(n):
        sg, eg = g.get_bp_state(r)
        so, eo = o.get_bp_state(r)
        assert np.array_equal(sg, so) and np.allclose(eg, eo, rtol=1e-7, atol=1e-10)


def test_restart_payload_continues_the_trajectory():
    """stop -> write (XYZ, VELO, ACCEL, STEP, PRNG counters, REPLICA labels) -> new handle -> read -> continue
    == the run that was never stopped but does at the same step what job_md does when it starts
    (src/job_md.F90:196-198 neighbor_list, :270-348 initial energies with flg_bp_MC set, bp_status_MC = bp_status);
    T-REMD with the Monte-Carlo base-pair limit so that every generator is in play"""
    n, nex = 4, 10
    temps = [290.0, 310.0, 330.0, 350.0]
    spec = fx.t2s2hp("T2S2HP_folded", nstep_MC=5).replicated(n, temps=temps)

    def fresh():
        return remd_gpu(spec, n, temps, nex)

    def job_md_start(g):
        lab = g.remd_state()[0]
        for r in range(n):
            g.neighbor_list(r)
        g.set_bp_mc(1)
        for r in range(n):
            g.energy_sumup(r, temps[lab[r] - 1], True)
        g.set_bp_mc(0)
        g.bp_status_to_mc()

    a = fresh()
    a.md_steps(30)
    job_md_start(a)
    a.md_steps(30)
    ref = [a.get_state(r) for r in range(n)]

    b = fresh()
    b.md_steps(30)
    saved = dict(state=[tuple(np.copy(t) for t in b.get_state(r)) for r in range(n)], rng=b.get_rng_state(),
                 r2l=b.remd_state()[0].copy(), idx=b.remd_idx_pair(), stats=b.remd_stats())
    assert saved["rng"][1] == 30 and saved["rng"][3] == n and (saved["rng"][4:] > 0).all() and saved["rng"][2] == 3 * (n // 2)
    b.close()
    c = gpu()
    spec.configure(c)
    c.remd_setup([n, 1, 1, 1], [temps, [], [], []], nex)
    c.remd_set_labels(saved["r2l"], saved["idx"])
    for r in range(n):
        c.set_state(r, *saved["state"][r])
    c.set_rng_state(saved["rng"])
    assert c.get_step() == 30
    job_md_start(c)
    c.md_steps(30)
    assert np.array_equal(c.remd_state()[0], a.remd_state()[0]) and c.remd_idx_pair() == a.remd_idx_pair()
    assert np.array_equal(c.get_rng_state(), a.get_rng_state())        # every draw counter where the other run has it
    for r in range(n):
        for u, v in zip(ref[r], c.get_state(r)):
            # same arithmetic; the base-pair forces of a bead are added with shared-memory atomics whose order
            # is not fixed, so two runs agree to the last bits only
            assert np.abs(u - v).max() < 1e-11


def test_set_step_keeps_a_pending_rebuild_trigger():
    """moving the step counter must not lose the displacement top-2 (src/job_md.F90:376-388)"""
    spec = fx.cag47(4, pbc=True)
    spec.nl_margin = 3.0
    out = []
    for shift in (0, 1001):
        g = gpu()
        spec.configure(g)
        g.init_velocities(0)
        g.neighbor_list(0)
        nreb = g.md_steps(30)
        g.set_step(g.get_step() + shift)
        nreb += g.md_steps(60)
        out.append(nreb)
    assert out[0] >= 2
    # the Gaussians of the shifted run differ (the step is their counter), the rebuild cadence must survive
    assert out[1] >= 2


@pytest.mark.parametrize("nstep_MC", [0, 10])
def test_force_remd_in_the_replica_local_kernel(nstep_MC):
    """F-REMD (check/T2S2HP_REMD_F: dual constant-force tweezers on beads (1, 27), the force magnitude is the
    replica variable): the replica-local persistent kernel carries the tweezers force; without the Monte-Carlo
    limit the exchange step runs inside the kernel (tweezers energy -F.(x_i - x_j) with own and partner force,
    src/energy_tweezers.F90:21-31), with it (the input's nstep_MC = 10) through the launch chain"""
    from sis_b200.system import PN2KCALMOLA
    n, nex = 4, 10
    forces = [f * PN2KCALMOLA for f in (0.0, 6.0, 12.0, 18.0)]
    spec = fx.t2s2hp("T2S2HP_folded", nstep_MC=nstep_MC).replicated(n)
    spec.twz_dcf_pairs = [(1, 27)]
    spec.twz_dcf_forces = [(1.0, 0.0, 0.0)]
    res = []
    for opts in (dict(launch_mode=0), dict()):
        g = gpu(**opts)
        spec.configure(g)
        g.remd_setup([1, n, 1, 1], [[], forces, [], []], nex)
        for r in range(n):
            g.init_velocities(r)
            g.neighbor_list(r)
        g.md_steps(60)
        res.append(([g.get_state(r) for r in range(n)], g.remd_state()[0].copy(), g.remd_stats(), g.last_step_kernel(), g.launch_count()))
    assert is_replica_local(res[1][3])
    assert np.array_equal(res[0][1], res[1][1]) and res[0][2] == res[1][2] and res[0][2][0] > 0
    for r in range(n):
        for a, b in zip(res[0][0][r], res[1][0][r]):
            assert np.abs(a - b).max() < 1e-7
    if nstep_MC == 0:
        assert res[1][4] < res[0][4] // 4      # one launch per md_steps call: the exchange ran inside the kernel
        o = Oracle()
        spec.configure(o)
        o.remd_setup([1, n, 1, 1], [[], forces, [], []], nex)
        for r in range(n):
            o.init_velocities(r)
            o.neighbor_list(r)
        counter = 0
        for _ in range(6):
            o.md_run(nex)
            _, re = o.energy_replica()
            u = np.array([o.uniform(spec.rng_seed, 0x7fffffff, counter + i) for i in range(n // 2)])
            counter += n // 2
            o.replica_exchange(re, u)
        assert np.array_equal(res[1][1], o.remd_state()[0])
        for r in range(n):
            assert np.abs(res[1][0][r][0] - o.get_state(r)[0]).max() < 1e-6
