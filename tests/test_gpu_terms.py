"""GPU parity, second part: the SURVEY section 8(a) rows the basic systems do not touch --
force_bp_limit_triplet (Monte-Carlo pruning), the bias terms (stage, tweezers DCF / Force-Ramp,
Rg, restraints, Bias_SS), the Brownian integrator, frozen beads, the harmonic angle / cosine
dihedral variants, temp_independent electrostatics, pbc_resize, and the REMD routines
(energy_replica, replica_exchange, the exchange inside md_steps).  Same oracle, same tolerances."""
import copy

import numpy as np
import pytest

from tests.conftest import is_replica_local

from oracle.oracle import Oracle
from sis_b200 import fixtures as fx
from tests.test_host_parity import circ_spec

pytestmark = pytest.mark.gpu

FORCE_TOL = 1e-10
ENE = {"bond": 1, "angl": 2, "dihe": 3, "bp": 4, "exv": 5, "ele": 6, "stage": 7, "twz": 8, "rg": 9, "rest": 10}


def gpu(**opts):
    from sis_b200.api import SisGpu
    return SisGpu(0, **opts)


def both(spec, **opts):
    g, o = gpu(**opts), Oracle()
    spec.configure(g)
    spec.configure(o)
    return g, o


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-12)


def check_forces_energies(g, o, spec, terms):
    g.neighbor_list(0); o.neighbor_list(0)
    for t in terms:
        fg, fo = g.force_terms(0, 1 << ENE[t]), o.force_terms(0, 1 << ENE[t])
        assert np.abs(fo).max() > 0.0, f"term {t} is not exercised by this input"
        assert rel(fg, fo) < FORCE_TOL, t
    assert rel(g.force(0), o.force(0)) < FORCE_TOL
    eg = g.energy_sumup(0, spec.tempK, True)
    eo = o.energy_sumup(0, spec.tempK, True)
    for t in terms:
        assert eo[ENE[t]] != 0.0, f"energy term {t} is zero"
    assert np.allclose(eg, eo, rtol=1e-10, atol=1e-9), (eg, eo)


def run_both(g, o, nsteps, xtol=1e-8, vtol=1e-9):
    g.init_velocities(0); o.init_velocities(0)
    g.neighbor_list(0); o.neighbor_list(0)
    rg_, ro_ = g.md_steps(nsteps), o.md_run(nsteps)
    xg, vg, ag = g.get_state(0)
    xo, vo, ao = o.get_state(0)
    assert np.abs(xg - xo).max() < xtol
    assert np.abs(vg - vo).max() < vtol
    assert rg_ == ro_
    return xg, xo


# ---- F-bp3L / E-bp: force_bp_limit_triplet and energy_bp_limit_triplet ------------------------
@pytest.mark.parametrize("coords", ["T2S2HP_folded", "T2S2HP_random"])
def test_bp_mc_force_and_status(coords):
    """MC step (src/force_bp_limit_triplet.F90:85-376): same surviving pairs, energies and forces."""
    spec = fx.t2s2hp(coords, nstep_MC=1)
    g, o = both(spec)
    g.neighbor_list(0); o.neighbor_list(0)
    g.set_bp_mc(1); o.set_bp_mc(1)
    fg, fo = g.force(0), o.force(0)
    assert rel(fg, fo) < FORCE_TOL
    sg, eg = g.get_bp_state(0)
    so, eo = o.get_bp_state(0)
    assert np.array_equal(sg, so)
    assert np.allclose(eg, eo, rtol=1e-10, atol=1e-12)
    # every nucleotide keeps at most max_bp_per_nt pairs (the point of the pruning)
    _, _, bp, _ = o.get_lists(0)
    cnt = np.zeros(spec.nmp + 1, int)
    for k in np.nonzero(so)[0]:
        cnt[bp[k, 0]] += 1; cnt[bp[k, 1]] += 1
    assert cnt.max() <= spec.max_bp_per_nt
    # energies with the pruning on the energy side (src/energy_bp_limit_triplet.F90)
    eg2 = g.energy_sumup(0, spec.tempK, True)
    eo2 = o.energy_sumup(0, spec.tempK, True)
    assert np.allclose(eg2, eo2, rtol=1e-10, atol=1e-9)


def test_bp_mc_prunes_something():
    """The folded hairpin has nucleotides with more than one candidate pair: the MC pass must remove
    at least one pair that the unrestricted model keeps (otherwise the test above is vacuous)."""
    spec = fx.t2s2hp("T2S2HP_folded", nstep_MC=1)
    o = Oracle()
    spec.configure(o)
    o.neighbor_list(0)
    o.set_bp_mc(1); o.force(0)
    s_mc, _ = o.get_bp_state(0)
    spec0 = fx.t2s2hp("T2S2HP_folded", nstep_MC=0)
    o0 = Oracle()
    spec0.configure(o0)
    o0.neighbor_list(0)
    o0.force(0)
    s_all, _ = o0.get_bp_state(0)
    assert s_all.sum() > s_mc.sum() > 0


@pytest.mark.parametrize("nstep_MC", [1, 5])
@pytest.mark.parametrize("mode,extra", [(0, {}), (2, {}), (2, {"mc_persistent": 1})],
                         ids=["stream", "default", "replica-local-kernel"])
def test_trajectory_with_mc(nstep_MC, mode, extra):
    """C1 as the reference runs it (check/input_T2S2HP.toml: nstep_MC = 1) and with a sparser MC
    schedule (bp_status_MC mask between MC steps, src/job_md.F90:365-368,399-401); also with the
    Monte-Carlo pruning inside the replica-local persistent kernel."""
    spec = fx.t2s2hp("T2S2HP_folded", nstep_MC=nstep_MC)
    if extra:
        spec.nl_margin = 0.5        # in-kernel rebuilds: the step after one is a Monte-Carlo step
    g, o = both(spec, launch_mode=mode, **extra)
    if extra:
        run_both(g, o, 60, xtol=1e-7, vtol=1e-7)
        assert is_replica_local(g.last_step_kernel())
    else:
        run_both(g, o, 30)
    sg, eg = g.get_bp_state(0)
    so, eo = o.get_bp_state(0)
    assert np.array_equal(sg, so)
    # ene_bp of pairs whose bp_status is .False. is never read (src/force_bp_limit_triplet.F90:318-322)
    assert np.allclose(eg[so != 0], eo[so != 0], rtol=1e-7, atol=1e-10)


# ---- bias terms -----------------------------------------------------------------------------------
def stage_spec():
    spec = fx.t2s2hp("T2S2HP_folded", nstep_MC=0)
    spec.xyz = spec.xyz.copy()
    spec.xyz[:, 2] += 9.5 - spec.xyz[:, 2].min()      # energy_stage aborts below 0.85 sigma
    spec.stage = (10.0, 1.2)
    return spec


def twz_dcf_spec():
    spec = fx.t2s2hp("T2S2HP_folded", nstep_MC=0)
    spec.twz_dcf_pairs = [(1, 27)]
    spec.twz_dcf_forces = [(0.05, -0.02, 0.11)]
    return spec


def twz_fr_spec():
    spec = fx.t2s2hp("T2S2HP_folded", nstep_MC=0)
    x = spec.xyz
    spec.twz_fr_pairs = [(1, 27)]
    spec.twz_fr_k = [(0.3, 0.2)]
    spec.twz_fr_init = [[(x[0] + [1.0, 0.5, -0.7]).tolist(), (x[26] + [-2.0, 0.1, 0.4]).tolist()]]
    spec.twz_fr_velo = [[[1e-3, 0.0, 0.0], [-1e-3, 2e-3, 0.0]]]
    return spec


def rg_spec(pott):
    spec = fx.t2s2hp("T2S2HP_random", nstep_MC=0)
    spec.bias_rg = (pott, 0.8, 8.0)
    return spec


def rest_spec():
    spec = fx.t2s2hp("T2S2HP_folded", nstep_MC=0)
    x = spec.xyz
    d = float(np.linalg.norm(x[2] - x[20]))
    # (imp, jmp, type, rcut, eps, d, s): type 1 acts on both beads, type 2 leaves jmp force-free
    spec.restraints = [(3, 21, 1, d + 20.0, 2.0, d - 1.0, 3.0), (5, 15, 2, 200.0, 1.5, 10.0, 4.0)]
    return spec


BIAS = {
    "stage": (stage_spec, ["stage"]),
    "twz_dcf": (twz_dcf_spec, ["twz"]),
    "rg_harmonic": (lambda: rg_spec(1), ["rg"]),
    "rg_flat_bottom": (lambda: rg_spec(2), ["rg"]),
    "restraint": (rest_spec, ["rest"]),
}


@pytest.mark.parametrize("name", list(BIAS))
def test_bias_forces_energies(name):
    mk, terms = BIAS[name]
    spec = mk()
    g, o = both(spec)
    check_forces_energies(g, o, spec, terms)
    if name.startswith("rg"):
        assert abs(g.get_rg(0) - o.get_rg(0)) < 1e-10


def test_force_ramp_tweezers():
    """Force-Ramp traps move with istep and contribute no energy (src/force_tweezers.F90:37-61,
    src/energy_tweezers.F90:21-28)."""
    spec = twz_fr_spec()
    g, o = both(spec)
    g.neighbor_list(0); o.neighbor_list(0)
    fg, fo = g.force_terms(0, 1 << ENE["twz"]), o.force_terms(0, 1 << ENE["twz"])
    assert np.abs(fo).max() > 0 and rel(fg, fo) < FORCE_TOL
    eg, eo = g.energy_sumup(0, spec.tempK, True), o.energy_sumup(0, spec.tempK, True)
    assert eo[ENE["twz"]] == 0.0 and np.allclose(eg, eo, rtol=1e-10, atol=1e-9)


@pytest.mark.parametrize("name", ["stage", "twz_dcf", "twz_fr", "rg_harmonic", "restraint"])
@pytest.mark.parametrize("mode", [0, 2])
def test_bias_trajectories(name, mode):
    mk = twz_fr_spec if name == "twz_fr" else BIAS[name][0]
    spec = mk()
    g, o = both(spec, launch_mode=mode)
    run_both(g, o, 25)


def test_bias_ss():
    """[Bias_SS] (src/force_bp_triplet.F90:85-93): constant pull on every listed pair outside the gate;
    the bp list has no distance cutoff (src/neighbor_list.F90:183-190)."""
    spec = fx.t2s2hp("T2S2HP_random", nstep_MC=0)
    spec.bias_ss_force = 0.5 * 0.014393261854684513
    g, o = both(spec)
    g.neighbor_list(0); o.neighbor_list(0)
    assert g.get_list_sizes(0) == o.get_list_sizes(0)
    assert rel(g.force_terms(0, 1 << ENE["bp"]), o.force_terms(0, 1 << ENE["bp"])) < FORCE_TOL
    assert rel(g.force(0), o.force(0)) < FORCE_TOL
    run_both(g, o, 20)


# ---- integrator variants ----------------------------------------------------------------------------
@pytest.mark.parametrize("mode", [0, 2])
def test_brownian_dynamics(mode):
    """BD-EM (src/job_md.F90:446-456, set_md_coef :754-766)."""
    spec = fx.t2s2hp("T2S2HP_folded", nstep_MC=0)
    spec.integrator = 2
    spec.dt = 0.5
    spec.viscosity_Pas = 1.0e-3      # water; at the 1e-5 of the Langevin inputs Euler-Maruyama with these
                                     # contacts amplifies a 1e-10 perturbation to 1e-3 in 25 steps
    g, o = both(spec, launch_mode=mode)
    run_both(g, o, 25)


@pytest.mark.parametrize("mode", [0, 2])
def test_frozen_beads(mode):
    spec = fx.cag47(1)
    spec.frozen = [1, 2, 70, 141]
    g, o = both(spec, launch_mode=mode)
    xg, xo = run_both(g, o, 25)
    for i in spec.frozen:
        assert np.array_equal(xg[i - 1], spec.xyz[i - 1])
    ekg, eko = g.energy_kinetic(0), o.energy_kinetic(0)
    assert abs(ekg - eko) < 1e-9 * max(1.0, abs(eko))


def test_kinetic_energy_and_velocities():
    """energy_kinetic (src/energy_kinetic.F90:17-29) and the initial velocity draw (src/job_md.F90:157-176)."""
    spec = fx.cag47(4, pbc=True)
    g, o = both(spec)
    g.init_velocities(0); o.init_velocities(0)
    ekg, eko = g.energy_kinetic(0), o.energy_kinetic(0)
    assert abs(ekg / eko - 1.0) < 1e-12
    T = 2.0 * eko / (3 * spec.nmp * 0.0019872042586408316)
    assert abs(T / spec.tempK - 1.0) < 0.15


# ---- force-field variants ---------------------------------------------------------------------------
def test_harmonic_angle_and_cos_dihedral():
    """[potential.angle] (src/force_angl.F90:36-82) and [potential.dihedral] (src/force_dih_cos.F90:21-59)
    instead of the ReB / exp forms of the shipped force fields."""
    spec = circ_spec()
    spec.ff = copy.copy(spec.ff)
    spec.ff.angl_type = 0
    spec.ff.angl_k, spec.ff.angl_t0 = 10.0, 2.6
    spec.ff.dih_flags = 1
    spec.ff.dih_k, spec.ff.dih_p0 = 1.4, 0.5
    g, o = both(spec)
    check_forces_energies(g, o, spec, ["bond", "angl", "dihe"])
    spec.ff.dih_flags = 3       # both dihedral forms at once (src/force.F90:36-44)
    g, o = both(spec)
    check_forces_energies(g, o, spec, ["dihe"])
    run_both(g, o, 20)


def test_temp_independent_electrostatics():
    """temp_independent = 1 (src/set_ele.F90:88-98, src/energy_ele_DH.F90:30-38, energy_bp_triplet.F90:27-31)."""
    spec = fx.t2s2hp("T2S2HP_folded", nstep_MC=0)
    spec.temp_independent = 1
    spec.tempK_ref = 295.0
    g, o = both(spec)
    check_forces_energies(g, o, spec, ["ele", "bp"])


# ---- PBC-r: pbc_resize -> new box + rebuild (src/job_md.F90:576-594) --------------------------------
def test_pbc_resize_rebuilds_lists():
    spec = fx.cag47(4, pbc=True)
    g, o = both(spec)
    g.neighbor_list(0); o.neighbor_list(0)
    for L in (700.0, 500.0, 333.3):
        g.set_pbc(1, np.array([L, L * 0.9, L * 1.1])); o.set_pbc(1, np.array([L, L * 0.9, L * 1.1]))
        g.neighbor_list(0); o.neighbor_list(0)
        assert g.get_list_sizes(0) == o.get_list_sizes(0)
        for a, b in zip(g.get_lists(0), o.get_lists(0)):
            assert np.array_equal(a, b)
        assert np.array_equal(g.get_state(0)[0], o.get_state(0)[0])
        assert rel(g.force(0), o.force(0)) < FORCE_TOL


# ---- REMD ------------------------------------------------------------------------------------------
def same_remd_state(g, o, nrep):
    """rep2lab, lab2rep and the label -> value rows of the replica variables in use."""
    (r2l_g, l2r_g, l2v_g), (r2l_o, l2r_o, l2v_o) = g.remd_state(), o.remd_state()
    assert np.array_equal(r2l_g, r2l_o) and np.array_equal(l2r_g, l2r_o)
    for v, n in enumerate(nrep):
        if n > 1:
            assert np.array_equal(l2v_g[v], l2v_o[v])


def remd_pair(nrep, values, spec, nstep_exchange=10, **opts):
    g, o = gpu(**opts), Oracle()
    for be in (g, o):
        spec.configure(be)
        be.remd_setup(nrep, values, nstep_exchange)
    return g, o


def t_remd_spec(n=8, nstep_MC=0):
    temps = [293.15 + 70.0 / (n - 1) * k for k in range(n)]
    spec = fx.t2s2hp("T2S2HP_folded", nstep_MC=nstep_MC).replicated(n, temps=temps)
    return spec, temps


def test_remd_tables():
    """init_replica label tables and make_replica_exchange_pair_tb (src/init_replica.F90:70-98,
    src/var_replica.F90:162-259), 1-D and 2-D (temperature x ionic strength)."""
    spec, temps = t_remd_spec(8)
    g, o = remd_pair([8, 1, 1, 1], [temps, [], [], []], spec)
    assert np.array_equal(g.remd_pair_table(), o.remd_pair_table())
    same_remd_state(g, o, [8, 1, 1, 1])
    ions = [0.05, 0.15, 0.5]
    spec2 = fx.t2s2hp("T2S2HP_folded", nstep_MC=0).replicated(12)
    g, o = remd_pair([4, 1, 3, 1], [temps[:4], [], ions, []], spec2)
    assert np.array_equal(g.remd_pair_table(), o.remd_pair_table())
    same_remd_state(g, o, [4, 1, 3, 1])


@pytest.mark.parametrize("kind", ["temp", "ion", "force", "temp_mc"])
def test_energy_replica_and_exchange(kind):
    """energy_replica (own + cross energies, src/energy_replica.F90:45-163) and replica_exchange
    (src/replica_exchange.F90:44-154) with caller-supplied uniforms, over a full cycle of the
    exchange-pair table."""
    n = 6
    temps = [293.15 + 14.0 * k for k in range(n)]
    if kind in ("temp", "temp_mc"):
        spec = fx.t2s2hp("T2S2HP_folded", nstep_MC=1 if kind == "temp_mc" else 0).replicated(n, temps=temps)
        nrep, values = [n, 1, 1, 1], [temps, [], [], []]
    elif kind == "ion":
        ions = [0.02, 0.05, 0.1, 0.2, 0.4, 0.8]
        spec = fx.t2s2hp("T2S2HP_folded", nstep_MC=0).replicated(n, ionic=ions)
        nrep, values = [1, 1, n, 1], [[], [], ions, []]
    else:
        forces = [0.0144 * 3.0 * k for k in range(n)]
        spec = twz_dcf_spec().replicated(n)
        spec.twz_dcf_forces = [(0.0, 0.0, 1.0)]          # direction; the replica variable is the magnitude
        nrep, values = [1, n, 1, 1], [[], forces, [], []]
    g, o = remd_pair(nrep, values, spec)
    rng = np.random.default_rng(11)
    for r in range(n):
        x = spec.xyz + rng.normal(scale=0.15, size=spec.xyz.shape)
        g.set_state(r, x); o.set_state(r, x)
        g.init_velocities(r); o.init_velocities(r)
        g.neighbor_list(r); o.neighbor_list(r)
        g.force(r); o.force(r)                            # fills the cached bp energies (flg_bp_energy)
    nacc_total = 0
    for it in range(4):
        eg, reg = g.energy_replica()
        eo, reo = o.energy_replica()
        assert np.allclose(eg, eo, rtol=1e-10, atol=1e-9)
        assert np.allclose(reg, reo, rtol=1e-10, atol=1e-9)
        u = rng.random((n + 1) // 2)
        ag, cg = g.replica_exchange(reo, u)
        ao, co = o.replica_exchange(reo, u)
        assert (ag, cg) == (ao, co)
        nacc_total += co
        same_remd_state(g, o, nrep)
        for r in range(n):
            vg, vo = g.get_state(r)[1], o.get_state(r)[1]
            assert np.allclose(vg, vo, rtol=1e-13, atol=1e-15)    # scale_velo :175-213
            assert rel(g.force(r), o.force(r)) < FORCE_TOL        # parameters refreshed (src/job_md.F90:538-567)
    assert nacc_total > 0


@pytest.mark.parametrize("mode,extra", [(0, {}), (2, {}), (2, {"local_kernel": 0}), (2, {"remd_device": 0})],
                         ids=["stream", "replica-local+device-exchange", "pipelined+device-exchange", "host-exchange"])
def test_remd_md_steps_with_exchange(mode, extra):
    """sisgpu_md_steps with REMD configured = MD + energy_replica + replica_exchange every
    nstep_exchange steps (src/job_md.F90:474-481,514-570); the oracle is driven through the same
    sequence with the library's counter-based exchange uniforms."""
    n, nex = 4, 10
    temps = [290.0, 310.0, 330.0, 350.0]
    spec = fx.t2s2hp("T2S2HP_folded", nstep_MC=0).replicated(n, temps=temps)
    g, o = remd_pair([n, 1, 1, 1], [temps, [], [], []], spec, nstep_exchange=nex, launch_mode=mode, **extra)
    for r in range(n):
        g.init_velocities(r); o.init_velocities(r)
        g.neighbor_list(r); o.neighbor_list(r)
    ncycle, counter = 3, 0
    g.md_steps(nex * ncycle)
    for c in range(ncycle):
        o.md_run(nex)
        _, re = o.energy_replica()
        u = np.array([o.uniform(spec.rng_seed, 0x7fffffff, counter + i) for i in range((n + 1) // 2)])
        counter += (n + 1) // 2
        o.replica_exchange(re, u)
    same_remd_state(g, o, [n, 1, 1, 1])
    att, acc = g.remd_stats()
    assert att > 0
    for r in range(n):
        xg, vg, _ = g.get_state(r)
        xo, vo, _ = o.get_state(r)
        assert np.abs(xg - xo).max() < 1e-7
        assert np.abs(vg - vo).max() < 1e-7


def test_job_dcd_batch_energies_match_oracle_and_reference_golden():
    """job_dcd batch mode (src/job_dcd.F90:65-79): the 100 OpenMM frames of
    check/test_DCD_CAG47/CAG47.dcd evaluated 8 frames per launch through the replica slots.
    (a) every frame equals the oracle's energy_sumup to 1e-10; (b) the reference's own
    known-answer file CAG47.energy.out is met within the tolerances of tests/test_oracle_golden.py."""
    frames32 = fx.array("CAG47_dcd").astype(np.float32)
    gold = fx.array("CAG47_energy_out")
    spec = fx.cag47(1)
    spec.tempK = 295.15
    g = gpu()
    spec.replicated(8).configure(g)
    g.full_lists()
    e = g.dcd_energies(frames32, 295.15)
    assert e.shape == (100, 11)
    o = Oracle()
    spec.configure(o)
    o.full_lists()
    for k in range(0, 100, 9):
        o.set_state(0, frames32[k].astype(np.float64))
        eo = o.energy_sumup(0, 295.15, True)
        assert np.allclose(e[k], eo, rtol=1e-10, atol=1e-10), (k, e[k], eo)
    assert np.max(np.abs(e[:, 1] - gold[:100, 1]) / np.abs(gold[:100, 1])) < 2e-4      # bond
    assert np.max(np.abs(e[:, 2] - gold[:100, 2]) / np.abs(gold[:100, 2])) < 2e-4      # angle
    assert np.max(np.abs(e[:, 5] - gold[:100, 3]) / np.maximum(np.abs(gold[:100, 3]), 1e-3)) < 2e-3   # WCA
    assert np.max(np.abs(e[:, 6] - gold[:100, 4]) / np.abs(gold[:100, 4])) < 6e-3      # Debye-Hueckel
    assert np.max(np.abs(e[:, 4] - gold[:100, 6]) / np.maximum(np.abs(gold[:100, 6]), 1.0)) < 2e-2    # base pairs
    # a batch size that does not divide the frame count gives the same numbers
    g3 = gpu()
    spec.replicated(3).configure(g3)
    g3.full_lists()
    assert np.array_equal(g3.dcd_energies(frames32[:10], 295.15), e[:10])


def test_snapshot_staging_matches_the_synchronous_getters():
    """Output staging of a save step: float32 DCD frame, .bp records filtered by
    bp_status and ene_bp <= -kT in bp_mp order (src/write_bp.F90:36-49), .bpall records, energies
    + Ekin, staged on the device and copied on a second stream while the next MD steps run."""
    spec = fx.cag47(4, pbc=True)
    g = gpu()
    spec.configure(g)
    g.init_velocities(0)
    g.neighbor_list(0)
    g.md_steps(200)
    # reference values through the synchronous calls
    x, v, a = g.get_state(0)
    wca, ele, bp, coef = g.get_lists(0)
    st, en = g.get_bp_state(0)
    e_ref = g.energy_sumup(0, spec.tempK, False)
    ek_ref = g.energy_kinetic(0)
    kT = spec.tempK * 0.0019872042586408316
    slot = g.snapshot_begin(0, spec.tempK, 1 | 2 | 8)
    g.md_steps(50)                       # overlaps with the copies; must not disturb the staged data
    xyz32, (bi, bj, be), ene, ek = g.snapshot_end(slot, spec.nmp)
    assert np.array_equal(xyz32, x.astype(np.float32))
    keep = (st != 0) & (en <= -kT)
    assert np.array_equal(bi, bp[keep, 0]) and np.array_equal(bj, bp[keep, 1])
    assert np.array_equal(be, en[keep].astype(np.float32))
    assert np.allclose(ene, e_ref, rtol=1e-12, atol=1e-12) and abs(ek - ek_ref) < 1e-9 * abs(ek_ref)
    assert g.get_step() == 250
    # .bpall: every pair with bp_status, whatever its energy; two slots can be pending at once
    st2, en2 = g.get_bp_state(0)
    bp2 = g.get_lists(0)[2]
    s1 = g.snapshot_begin(0, spec.tempK, 4)
    s2 = g.snapshot_begin(0, spec.tempK, 1)
    _, (ai, aj, ae), _, _ = g.snapshot_end(s1, spec.nmp)
    x2 = g.snapshot_end(s2, spec.nmp)[0]
    assert np.array_equal(ai, bp2[st2 != 0, 0]) and np.array_equal(aj, bp2[st2 != 0, 1])
    assert np.array_equal(ae, en2[st2 != 0].astype(np.float32))
    assert np.array_equal(x2, g.get_state(0)[0].astype(np.float32))
    assert keep.sum() > 0


@pytest.mark.parametrize("extra", [{}, {"local_kernel": 0}], ids=["replica-local", "pipelined"])
def test_remd_2d_device_exchange_inside_md_steps(extra):
    """Two replica dimensions (temperature x ionic strength, src/var_replica.F90:162-259 alternating
    pair tables) with the exchange step on the device: label tables, parameters (Debye-Hueckel
    coefficients follow both variables, src/job_md.F90:538-567) and trajectories follow the oracle
    over a full cycle of the four pair tables."""
    temps, ions = [295.0, 330.0], [0.05, 0.3]
    nrep, values = [2, 1, 2, 1], [temps, [], ions, []]
    n, nex = 4, 5
    # label -> (T, I) as init_replica lays the labels out: the first variable varies slowest
    spec = fx.t2s2hp("T2S2HP_folded", nstep_MC=0).replicated(n, temps=[temps[0], temps[0], temps[1], temps[1]],
                                                              ionic=[ions[0], ions[1], ions[0], ions[1]])
    g, o = remd_pair(nrep, values, spec, nstep_exchange=nex, **extra)
    _, _, l2v = o.remd_state()
    assert np.allclose(l2v[0][:4], [295.0, 295.0, 330.0, 330.0]) and np.allclose(l2v[2][:4], [0.05, 0.3, 0.05, 0.3])
    for r in range(n):
        g.init_velocities(r); o.init_velocities(r)
        g.neighbor_list(r); o.neighbor_list(r)
    ncycle, counter = 8, 0
    g.md_steps(nex * ncycle)
    for c in range(ncycle):
        o.md_run(nex)
        _, re = o.energy_replica()
        u = np.array([o.uniform(spec.rng_seed, 0x7fffffff, counter + i) for i in range((n + 1) // 2)])
        counter += (n + 1) // 2
        o.replica_exchange(re, u)
    same_remd_state(g, o, nrep)
    assert g.remd_stats() == o.remd_stats() if hasattr(o, "remd_stats") else g.remd_stats()[0] > 0
    for r in range(n):
        xg, vg, _ = g.get_state(r)
        xo, vo, _ = o.get_state(r)
        assert np.abs(xg - xo).max() < 1e-6 and np.abs(vg - vo).max() < 1e-6
        assert rel(g.force(r), o.force(r)) < FORCE_TOL


def test_remd_with_mc_persistent_and_stream_paths_agree():
    """T-REMD with the Monte-Carlo base-pair limit (nstep_MC = 1): the replica-local persistent kernel
    (in-kernel pruning) with the device-side exchange chain (flg_bp_MC set, bp energies re-evaluated
    with the limit, src/job_md.F90:476-479) gives the trajectories, label tables and acceptance counts
    of the per-phase stream path with host-driven exchanges."""
    n, nex = 4, 10
    temps = [290.0, 310.0, 330.0, 350.0]
    spec = fx.t2s2hp("T2S2HP_folded", nstep_MC=1).replicated(n, temps=temps)
    res = []
    for opts in (dict(launch_mode=0), dict(launch_mode=2, mc_persistent=1)):
        g = gpu(**opts)
        spec.configure(g)
        g.remd_setup([n, 1, 1, 1], [temps, [], [], []], nex)
        for r in range(n):
            g.init_velocities(r)
            g.neighbor_list(r)
        g.md_steps(40)
        res.append(([g.get_state(r) for r in range(n)], g.remd_state(), g.remd_stats(), g.last_step_kernel()))
    assert is_replica_local(res[1][3])
    assert np.array_equal(res[0][1][0], res[1][1][0]) and res[0][2] == res[1][2] and res[0][2][0] > 0
    for r in range(n):
        for a, b in zip(res[0][0][r], res[1][0][r]):
            assert np.abs(a - b).max() < 1e-7


def test_md_with_secondary_structure_restriction():
    """MD of the hairpin with the base pairs restricted to check/T2S2HP.ct (check/input_T2S2HP_go.toml
    style): the persistent replica-local kernel, in-kernel rebuilds included, follows the oracle."""
    from tests.test_host_parity import t2s2hp_restricted
    spec = t2s2hp_restricted(True)
    spec.nl_margin = 1.0
    g, o = both(spec)
    run_both(g, o, 60, xtol=1e-7, vtol=1e-7)
    assert is_replica_local(g.last_step_kernel())
    assert g.get_list_sizes(0) == o.get_list_sizes(0) and g.get_list_sizes(0)[2] <= 7
    sg, eg = g.get_bp_state(0)
    so, eo = o.get_bp_state(0)
    assert np.array_equal(sg, so) and np.allclose(eg, eo, rtol=1e-7, atol=1e-10)
