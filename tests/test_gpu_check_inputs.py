"""The reference's own check/ inputs on the GPU (north_star: "GPU forces/lists matching
job_check_force on all check/ inputs"), each through the C ABI against the oracle:

  check/nick_restraint/*          nicked dumbbell, base pairs from the .bpl, two Sigmoid-to-bead restraints
  check/input_T2S2HP_stage.toml   T2S2HP_random.xyz on the [stage] wall, Monte-Carlo base-pair limit every step
  check/input_T2S2HP.toml         T2S2HP_random.xyz: cutoff-free lists + CHECK_FORCE
  check/T2S2HP_folded.pdb, check/CAG47_4.pdb (+PBC), one CAG47 chain: CHECK_FORCE
plus the save-step records against the oracle's base-pair state in the reference's .bp layout
(src/write_bp.F90:36-49) and a system of unequal chain lengths (src/neighbor_list.F90:206-216).
The force fields the inputs name (htv23_5.0_ee-17.ff, htv23-19.ff) are not in the reference tree:
check/sis_dih_ReB_NN.ff, the alternative the inputs themselves comment in, is used."""
import json
import struct
from pathlib import Path

import numpy as np
import pytest

from oracle.oracle import Oracle
from sis_b200 import fixtures as fx
from sis_b200._capi import SisError
from sis_b200.system import SystemSpec
from tests.test_gpu_terms import ENE, both, check_forces_energies, rel, run_both

pytestmark = pytest.mark.gpu
FORCE_TOL = 1e-10
GOLDEN = Path(__file__).resolve().parent / "golden"
KB = 0.0019872042586408316


def nick_spec():
    d = json.loads((GOLDEN / "ss_pairs.json").read_text())
    spec = SystemSpec(seqs=fx.seqs("nick_dumbbell"), ff=fx.force_field("sis_dih_ReB_NN"), xyz=fx.array("nick_dumbbell"),
                      tempK=300.0, rng_seed=123, dt=0.2, bp_model=5, bp_min_loop=3, max_bp_per_nt=1, nstep_MC=0)
    spec.bp_restrict = d["nick_dumbbell_bpl"]
    spec.restraints = [tuple(r) for r in d["nick_dumbbell_restraints"]]
    return spec


def stage_input_spec(lift=True):
    spec = fx.t2s2hp("T2S2HP_random", nstep_MC=1)       # [Basepair] of the input has no nstep_MC: default 1
    spec.stage = (10.0, 1.2)
    if lift:
        spec.xyz = spec.xyz.copy()
        spec.xyz[:, 2] += 10.5 - spec.xyz[:, 2].min()
    return spec


def test_nick_dumbbell_with_restraints():
    spec = nick_spec()
    g, o = both(spec)
    g.neighbor_list(0); o.neighbor_list(0)
    assert g.get_list_sizes(0) == o.get_list_sizes(0)
    nbp = g.get_list_sizes(0)[2]
    assert 0 < nbp <= 14                                  # only pairs of the .bpl (read_ss filters applied)
    for a, b in zip(g.get_lists(0), o.get_lists(0)):
        assert np.array_equal(a, b)
    for t in ("bond", "angl", "dihe", "bp", "exv", "rest"):
        fg, fo = g.force_terms(0, 1 << ENE[t]), o.force_terms(0, 1 << ENE[t])
        assert rel(fg, fo) < FORCE_TOL, t
    fo = o.force_terms(0, 1 << ENE["rest"])
    assert np.abs(fo[0]).max() > 0 and np.abs(fo[46]).max() == 0.0      # Sigmoid-to-bead: the reference bead is force-free
    assert rel(g.force(0), o.force(0)) < FORCE_TOL
    eg, eo = g.energy_sumup(0, spec.tempK, True), o.energy_sumup(0, spec.tempK, True)
    assert eo[ENE["rest"]] < 0.0 and np.allclose(eg, eo, rtol=1e-10, atol=1e-9)
    # the input's own job: 100 MD steps at dt 0.2
    run_both(g, o, 100, xtol=1e-7, vtol=1e-7)


def test_nick_dumbbell_check_force():
    """CHECK_FORCE on the nick input.  A Sigmoid-to-bead restraint leaves its reference bead force-free by
    design (src/force_restraint.F90:19-41), so the reference's own check reports exactly that missing reaction
    on beads 47 and 2 and nothing else; without the restraints the input passes at 1e-3"""
    spec = nick_spec()
    rng = np.random.default_rng(11)
    spec.xyz = spec.xyz + 0.001 * rng.random(spec.xyz.shape)           # src/job_check_force.F90:46-51
    g, o = both(spec)
    g.full_lists(); o.full_lists()
    fe, ff, md = g.check_force(spec.tempK)
    feo, ffo, mdo = o.check_force(spec.tempK)
    assert rel(ff, ffo) < FORCE_TOL and np.abs(fe - feo).max() < 1e-6 and abs(md - mdo) < 1e-6
    bad = np.nonzero(np.abs(fe - ff).max(axis=1) > 1e-3)[0]
    assert set(bad.tolist()) == {1, 46}                                # 0-based: beads 2 and 47
    spec.restraints = []
    g, o = both(spec)
    g.full_lists()
    assert g.check_force(spec.tempK)[2] <= 1e-3


def test_stage_input():
    """check/input_T2S2HP_stage.toml: T2S2HP_random.xyz above the [stage] wall, Monte-Carlo base-pair limit
    every step; and the hard abort of energy_stage when a bead is below 0.85 sigma (src/energy_stage.F90:22-26)"""
    spec = stage_input_spec(lift=False)
    assert spec.xyz[:, 2].min() > 8.5                                  # the input starts above the abort height
    g, o = both(spec)
    check_forces_energies(g, o, spec, ["stage", "bond", "ele"])
    run_both(g, o, 40)
    sg, eg = g.get_bp_state(0)
    so, eo = o.get_bp_state(0)
    assert np.array_equal(sg, so) and np.allclose(eg[so != 0], eo[so != 0], rtol=1e-7, atol=1e-10)
    low = stage_input_spec(lift=False)
    low.xyz = low.xyz.copy()
    low.xyz[5, 2] = 8.0
    g, o = both(low)
    g.neighbor_list(0); o.neighbor_list(0)
    with pytest.raises(SisError, match="stage"):
        o.energy_sumup(0, low.tempK, True)
    with pytest.raises(SisError, match="stage"):
        g.energy_sumup(0, low.tempK, True)


def _cag47_wide(nchain, pbc, cut):
    # the truncated Debye-Hueckel term is discontinuous at the cutoff (5e-5 kcal/mol at 50 A): a pair that straddles
    # it within +-0.001 A breaks the central difference in the reference too, so the criterion is checked with the
    # cutoff moved beyond every pair (tests/test_oracle_golden.py does the same for the oracle)
    spec = fx.cag47(nchain, pbc=pbc)
    spec.ele_cutoff = cut
    return spec


CHECK = {
    "T2S2HP_random": lambda: fx.t2s2hp("T2S2HP_random", nstep_MC=1),
    "T2S2HP_folded": lambda: fx.t2s2hp("T2S2HP_folded", nstep_MC=1),
    "CAG47": lambda: _cag47_wide(1, False, 2000.0),
    "CAG47_4_pbc": lambda: _cag47_wide(4, True, 300.0),
}


@pytest.mark.parametrize("name", list(CHECK))
def test_check_force_on_the_check_inputs(name):
    """[Job] type = "CHECK_FORCE" (src/job_check_force.F90): cutoff-free lists (src/main.F90:148-154), the
    0.001 A perturbation, analytic force against the central difference of energy_sumup, tolerance 1e-3"""
    spec = CHECK[name]()
    rng = np.random.default_rng(3)
    spec.xyz = spec.xyz + 0.001 * rng.random(spec.xyz.shape)
    g, o = both(spec)
    g.full_lists(); o.full_lists()
    assert g.get_list_sizes(0) == o.get_list_sizes(0)
    for a, b in zip(g.get_lists(0)[:3], o.get_lists(0)[:3]):
        assert np.array_equal(a, b)
    fe, ff, md = g.check_force(spec.tempK)
    feo, ffo, mdo = o.check_force(spec.tempK)
    assert md <= 1e-3, md
    assert rel(ff, ffo) < FORCE_TOL
    assert np.abs(fe - feo).max() < 1e-6


def bp_file_bytes(i, j, e):
    """one frame of a .bp file: (int16 i, int16 j, real32 e) records + the (0, 0, 0.0) terminator
    (src/write_bp.F90:36-49, kinds of src/var_io.F90:25-26)"""
    out = b"".join(struct.pack("<hhf", int(a), int(b), float(c)) for a, b, c in zip(i, j, e))
    return out + struct.pack("<hhf", 0, 0, 0.0)


@pytest.mark.parametrize("name", ["t2s2hp", "cag47_4_pbc"])
def test_snapshot_records_against_the_oracle_state(name):
    """save-step staging against the ORACLE's bp_status / ene_bp (not the GPU's own getters): the bytes
    write_bp would put into the .bp and .bpall files, the float32 DCD frame, the 11 energies with the cached
    bp energies (flg_bp_energy) and Ekin"""
    spec = fx.t2s2hp("T2S2HP_folded", nstep_MC=0) if name == "t2s2hp" else fx.cag47(4, pbc=True)
    g, o = both(spec)
    for be in (g, o):
        be.init_velocities(0)
        be.neighbor_list(0)
    g.md_steps(40); o.md_run(40)
    kT = spec.tempK * KB
    so, eo = o.get_bp_state(0)
    bp = o.get_lists(0)[2]
    xo, vo, _ = o.get_state(0)
    e_o = o.energy_sumup(0, spec.tempK, False)          # cached bp energies of the last force call
    ek_o = o.energy_kinetic(0)
    slot = g.snapshot_begin(0, spec.tempK, 1 | 2 | 8)
    xyz32, (bi, bj, be_), ene, ek = g.snapshot_end(slot, spec.nmp)
    keep = (so != 0) & (eo <= -kT)
    assert keep.sum() > 0
    assert np.array_equal(bi, bp[keep, 0]) and np.array_equal(bj, bp[keep, 1])
    assert np.allclose(be_, eo[keep].astype(np.float32), rtol=1e-6)
    ref = bp_file_bytes(bp[keep, 0], bp[keep, 1], eo[keep])
    got = bp_file_bytes(bi, bj, be_)
    assert len(got) == len(ref) and got[-8:] == ref[-8:]
    # identical index fields byte for byte; energies to float32 rounding of values equal to 1e-7
    assert all(got[8 * k:8 * k + 4] == ref[8 * k:8 * k + 4] for k in range(len(bi)))
    assert np.abs(xyz32 - xo.astype(np.float32)).max() < 1e-4
    assert np.allclose(ene, e_o, rtol=1e-8, atol=1e-8) and abs(ek - ek_o) < 1e-7 * abs(ek_o)
    s2 = g.snapshot_begin(0, spec.tempK, 4)              # .bpall: every pair with bp_status
    _, (ai, aj, ae), _, _ = g.snapshot_end(s2, spec.nmp)
    assert np.array_equal(ai, bp[so != 0, 0]) and np.array_equal(aj, bp[so != 0, 1])
    assert np.allclose(ae, eo[so != 0].astype(np.float32), rtol=1e-6)


def test_unequal_chain_lengths():
    """Two chains of different length (the hairpin and a truncated copy): lists, base-pair bead sextets and
    forces as the oracle gives them.  src/neighbor_list.F90:212 tests `j == nmp_chain(ichain)` where
    nmp_chain(jchain) is meant; oracle and GPU both implement the intended index (DESIGN.md section 2 lists
    this as the one deliberate deviation: the reference form would pair bead j = nmp_chain(ichain) of the
    longer second chain with a wrapped j+1 neighbour)."""
    seq = fx.seqs("T2S2HP")[0]
    x = fx.array("T2S2HP_folded")
    spec = fx.t2s2hp("T2S2HP_folded", nstep_MC=0)
    spec.seqs = [seq[:21], seq]                            # ichain shorter than jchain: the case :212 mistreats
    spec.xyz = np.concatenate([x[:21] + np.array([0.0, 0.0, 18.0]), x])
    g, o = both(spec)
    g.neighbor_list(0); o.neighbor_list(0)
    assert g.get_list_sizes(0) == o.get_list_sizes(0)
    lg, lo = g.get_lists(0), o.get_lists(0)
    for a, b in zip(lg, lo):
        assert np.array_equal(a, b)
    bp = lo[2]
    inter = bp[(bp[:, 0] <= 21) & (bp[:, 1] > 21)]
    assert len(inter) > 0 and (inter[:, 5] == inter[:, 1] + 1).all()       # j + 1 stays inside chain j
    for t in (1, 2, 3, 4, 5, 6):
        assert rel(g.force_terms(0, 1 << t), o.force_terms(0, 1 << t)) < FORCE_TOL
    run_both(g, o, 30, xtol=1e-7, vtol=1e-7)


def test_bpcoef_table_without_the_n_by_n_map():
    """write_bpcoef's content (src/write_bpcoef.F90:30-83) from the sequence alone: the same pairs and (dH, dS) the
    oracle's bp_lookup / bp3 table give, in (imp, jmp) order; 468-entry invariant of the table behind it"""
    spec = fx.t2s2hp("T2S2HP_folded", nstep_MC=0)
    g, o = both(spec)
    i, j, dH, dS = g.bpcoef()
    n3, _, tH, tS = o.bp3_table()
    assert n3 == 468
    want = []
    for a in range(1, spec.nmp):
        for b in range(a + 1, spec.nmp + 1):
            t, b3 = o.bp_lookup(a, b)          # 1-based, as bp_map(imp, jmp)
            if b3 > 0:
                want.append((a, b, tH[b3 - 1], tS[b3 - 1]))
    assert len(want) == len(i) > 0
    assert [(int(a), int(b)) for a, b in zip(i, j)] == [(w[0], w[1]) for w in want]
    assert np.allclose(dH, [w[2] for w in want], rtol=0, atol=1e-14) and np.allclose(dS, [w[3] for w in want], rtol=0, atol=1e-17)
    # what the driver prints for tempK = 300: pairs with dG < 0
    dG = dH - 300.0 * dS
    assert 0 < (dG < 0).sum() <= len(dG)
