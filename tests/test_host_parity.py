"""CPU-side parity of the PRODUCT's force / list code against the oracle.

The product's term arithmetic, gather roles and list predicates are `__host__ __device__`
code shared by the kernels; tests/csrc/hostcheck.cu compiles them for the host.  These tests
run in the build container (no GPU) so that sign, role and exclusion-rule errors are caught
before GPU time is spent; the GPU parity tests proper are tests/test_gpu_parity.py.
"""
import numpy as np
import pytest

from oracle.oracle import Oracle
from sis_b200 import fixtures as fx
from sis_b200.system import SystemSpec
from tests.hostcheck import HostCheck

TERMS = {"bond": 1, "angl": 2, "dihe": 3, "bp": 4, "exv": 5, "ele": 6}


def pairset(a):
    return set(map(tuple, np.asarray(a)[:, :2].tolist()))


def circ_spec():
    """Two circular RNAs + one linear chain with a dummy bead: exercises the closure terms,
    the circRNA exclusions of src/neighbor_list.F90:131-135,159-166 and SEQT%D handling."""
    rng = np.random.default_rng(7)
    seqs = ["GCGCAAUUGGCCAAGCUU", "AUGCGCAUAUGGCCAU", "GGDACUAGCUUAGCC"]
    n = sum(map(len, seqs))
    # compact random blob, ~6 A spacing along each chain
    xyz = []
    for s in seqs:
        p = rng.normal(scale=15.0, size=3)
        for _ in s:
            step = rng.normal(size=3)
            p = p + 5.8 * step / np.linalg.norm(step)
            xyz.append(p.copy())
    spec = SystemSpec(seqs=seqs, ff=fx.force_field("sis_dih_ReB_NN"), moltypes=[2, 2, 1], xyz=np.array(xyz),
                      tempK=300.0, dt=0.05, bp_model=5, nstep_MC=0, flg_ele=True, ele_cutoff_type=2,
                      ele_cutoff=10.0)
    assert spec.nmp == n
    return spec


def ct_pairs():
    import json
    return json.loads((fx.GOLDEN / "ss_pairs.json").read_text())["T2S2HP_ct_pairs"]


def t2s2hp_restricted(model5=True):
    """check/input_T2S2HP_go.toml / _bpl.toml: base pairs only between the pairs of the secondary
    structure file (src/read_ss.F90); extra listed pairs exercise the filters of read_ss (chain end,
    minimum loop, isolated pair) and, for bp_model /= 5, the pairs init_bp's loop skips but bp3_map keeps."""
    # (2, 19), (2, 21), (9, 26): read_ss accepts them (a complementary neighbour pair exists), the loop
    # of init_bp skips them (that neighbour is a chain end): listed for bp_model 3, not for 5
    pairs = ct_pairs() + [[1, 27], [9, 12], [10, 17], [9, 17], [3, 24], [2, 19], [21, 2], [9, 26]]
    if model5 is True:
        spec = fx.t2s2hp("T2S2HP_folded", nstep_MC=0)
    else:
        spec = SystemSpec(seqs=fx.seqs("T2S2HP"), ff=fx.force_field("rna_cg1"), xyz=fx.array("T2S2HP_folded"),
                          tempK=300.0, dt=0.05, bp_model=model5 if model5 else 1, bp_min_loop=3 if model5 else 4, nstep_MC=0, flg_ele=True,
                          ionic_strength=0.15, ele_cutoff_type=1, ele_cutoff=50.0, length_per_charge=4.38178046)
    spec.bp_restrict = pairs
    return spec


SPECS = {
    "t2s2hp": lambda: fx.t2s2hp("T2S2HP_folded", nstep_MC=0),
    "t2s2hp_ct": lambda: t2s2hp_restricted(True),
    "t2s2hp_ct_model1": lambda: t2s2hp_restricted(False),
    "t2s2hp_ct_model3": lambda: t2s2hp_restricted(3),
    "cag47_4_pbc": lambda: fx.cag47(4, pbc=True),
    "circ": circ_spec,
}


def both(spec):
    o, hcx = Oracle(), HostCheck()
    spec.configure(o)
    spec.configure(hcx)
    return o, hcx


@pytest.mark.parametrize("name", list(SPECS))
@pytest.mark.parametrize("full", [False, True])
def test_lists_bit_exact(name, full):
    spec = SPECS[name]()
    o, hcx = both(spec)
    if full:
        o.full_lists()
    else:
        o.neighbor_list(0)
    hcx.build_lists(all_pairs=full)
    wo, eo, bo, co = o.get_lists(0)
    wh, eh, bh, ch = hcx.lists()
    assert pairset(wo) == pairset(wh)
    assert pairset(eo) == pairset(eh)
    assert pairset(bo) == pairset(bh)
    # bp type and (dH, dS) per pair
    do = {(r[0], r[1]): (r[6], tuple(c)) for r, c in zip(bo.tolist(), co.tolist())}
    dh = {(r[0], r[1]): (r[2], tuple(c)) for r, c in zip(bh.tolist(), ch.tolist())}
    assert do == dh


@pytest.mark.parametrize("name", list(SPECS))
def test_forces_per_term(name):
    spec = SPECS[name]()
    o, hcx = both(spec)
    o.neighbor_list(0)
    hcx.build_lists(False)
    x, _, _ = o.get_state(0)            # wrapped coordinates
    assert np.allclose(x, hcx.xyz, rtol=0, atol=1e-12)
    for term, bit in TERMS.items():
        fo = o.force_terms(0, 1 << bit)
        fh = hcx.force_terms(0, 1 << bit)
        scale = max(np.abs(fo).max(), 1e-12)
        assert np.abs(fo - fh).max() / scale < 1e-10, term
    fo, fh = o.force(0), hcx.force_terms(0, 0x7FE)
    assert np.abs(fo - fh).max() / np.abs(fo).max() < 1e-10


@pytest.mark.parametrize("name", list(SPECS))
def test_bp_role_split_matches_oracle(name):
    """k_md_fused evaluates a base pair with one role per lane (bond, 4 angles, 2 dihedrals):
    the role arithmetic and bead routing summed serially must reproduce force_bp*."""
    spec = SPECS[name]()
    o, hcx = both(spec)
    o.neighbor_list(0)
    hcx.build_lists(False)
    fo = o.force_terms(0, 1 << TERMS["bp"])
    fh = hcx.force_bp_roles()
    assert np.abs(fo).max() > 0
    assert np.abs(fo - fh).max() / np.abs(fo).max() < 1e-10


@pytest.mark.parametrize("name", list(SPECS))
def test_bp_role_major_matches_oracle(name):
    """k_md_rep / k_md_fast evaluate the survivors of the distance gate role-major: seven role items per pair
    (bp_role_eval: the angles through one reciprocal square root each), then the pair energy and the six bead
    forces from the role results (bp_roles_u, bp_roles_bead_force) -- must reproduce force_bp* of the oracle."""
    spec = SPECS[name]()
    o, hcx = both(spec)
    o.neighbor_list(0)
    hcx.build_lists(False)
    fo = o.force_terms(0, 1 << TERMS["bp"])
    fh = hcx.force_bp_role_major()
    assert np.abs(fo).max() > 0
    assert np.abs(fo - fh).max() / np.abs(fo).max() < 1e-10


def test_mixed_mode_pair_scalar_within_the_stated_bound():
    """pair_scalar_f32 (option precision = 1) against the fp64 WCA / Debye-Hueckel scalars over the distances a
    listed pair can have: relative error of the scalar below the bound the GPU tests state for the forces"""
    spec = fx.cag47(4, pbc=True)
    o, hcx = both(spec)
    for wca, ele, lo, hi in ((True, False, 4.0, 9.99), (False, True, 3.0, 50.0), (True, True, 4.0, 9.99)):
        d = np.linspace(lo, hi, 400)
        ref = np.array([hcx.pair_scalar(x * x, wca, ele, False) for x in d])
        got = np.array([hcx.pair_scalar(x * x, wca, ele, True) for x in d])
        assert np.abs(ref).max() > 0
        assert np.abs(got - ref).max() / np.abs(ref).max() < 2e-6, (wca, ele)


def test_bp3_table_matches_oracle():
    spec = fx.t2s2hp()
    o, hcx = both(spec)
    no, ho, dHo, dSo = o.bp3_table()
    nh, hh, dHh, dSh = hcx.bp3_table()
    assert no == nh == 468
    assert (ho == hh).all()
    assert np.array_equal(dHo, dHh) and np.array_equal(dSo, dSh)


def test_rng_and_coefficients_match_oracle():
    spec = fx.t2s2hp()
    o, hcx = both(spec)
    for stream, step, bead in [(0, 1, 0), (3, 12345678901, 26), (63, 2**40 + 5, 9023)]:
        assert np.allclose(o.gauss(123, stream, step, bead), hcx.gauss(123, stream, step, bead), rtol=1e-13, atol=1e-15)
        assert o.uniform(123, stream, step) == hcx.uniform(123, stream, step)
    for imp in (0, 5, 26):
        assert np.allclose(o.md_coef(0, imp), hcx.coef(imp), rtol=1e-14)
    eo, eh = o.get_ele(0), hcx.rep_params()
    for k in ("ele_coef", "lambdaD", "diele"):
        assert eo[k] == pytest.approx(eh[k], rel=1e-15)


def test_gauss_statistics():
    hcx = HostCheck()
    g = np.array([hcx.gauss(1, 0, s, b) for s in range(1, 60) for b in range(60)])
    assert abs(g.mean()) < 0.03 and abs(g.std() - 1.0) < 0.03
    assert abs(np.corrcoef(g[:, 0], g[:, 1])[0, 1]) < 0.03


def test_restriction_keeps_skipped_pairs_for_models_other_than_5():
    """src/init_bp.F90:315-317: with a ct / bpseq / bpl restriction and bp_model /= 5, bp3_map = bp_map keeps
    the pairs read_ss accepted even where the loop of init_bp `cycle`s; bp_model 5 drops them."""
    quirk = {(2, 19), (2, 21), (9, 26)}
    for name, expect in (("t2s2hp_ct_model3", True), ("t2s2hp_ct", False)):
        o, hcx = both(SPECS[name]())
        o.full_lists()
        hcx.build_lists(all_pairs=True)
        po, ph = pairset(o.get_lists(0)[2]), pairset(hcx.lists()[2])
        assert po == ph
        assert quirk.issubset(po) == expect and (not expect) == quirk.isdisjoint(po)
