"""k_md_rep (sis_rep.cuh): cases the shared replica-local tests do not reach -- a small replica in a periodic box
(pbc_wrap at the rebuild events, minimum image on the static pair table), the three CTA widths against each other,
more replicas than the GPU holds at once (waves), and the event bookkeeping (rebuild count, lists left behind)
against the oracle."""
import numpy as np
import pytest

from oracle.oracle import Oracle
from sis_b200 import fixtures as fx

pytestmark = pytest.mark.gpu


def gpu(**opts):
    from sis_b200.api import SisGpu
    return SisGpu(0, **opts)


def pbc_hairpin():
    spec = fx.t2s2hp("T2S2HP_folded", nstep_MC=0)
    spec.box = (200.0, 210.0, 220.0)       # > 2 (cutoff + margin); the hairpin sits across the x and y faces
    spec.nl_margin = 0.4
    spec.xyz = spec.xyz - spec.xyz.min(axis=0) + np.array([190.0, -8.0, 100.0])
    return spec


def test_small_pbc_replica_follows_the_oracle_through_rebuild_events(replica_kernel):
    spec = pbc_hairpin()
    g, o = gpu(), Oracle()
    spec.configure(g); spec.configure(o)
    g.init_velocities(0); o.init_velocities(0)
    g.neighbor_list(0); o.neighbor_list(0)
    nreb_g = nreb_o = 0
    for k in (50, 50):
        nreb_g += g.md_steps(k)
        nreb_o += o.md_run(k)
        xg, vg, _ = g.get_state(0)
        xo, vo, _ = o.get_state(0)
        assert np.abs(xg - xo).max() < 1e-6 and np.abs(vg - vo).max() < 1e-7
    assert nreb_g == nreb_o and nreb_g >= 2
    if replica_kernel == "k_md_rep":
        assert g.last_step_kernel().startswith("k_md_rep")
    # the lists left behind are those of the last rebuild event
    for lg, lo in zip(g.get_lists(0)[:3], o.get_lists(0)[:3]):
        kg = np.lexsort(lg.T[::-1]) if len(lg) else []
        ko = np.lexsort(lo.T[::-1]) if len(lo) else []
        assert np.array_equal(lg[kg], lo[ko])
    eg, eo = g.energy_sumup(0, spec.tempK, False), o.energy_sumup(0, spec.tempK, False)
    assert np.allclose(eg, eo, rtol=1e-7, atol=1e-7)


@pytest.mark.parametrize("mc", [0, 5])
def test_cta_widths_agree(mc, replica_kernel):
    if replica_kernel != "k_md_rep":
        pytest.skip("k_md_rep only")
    temps = [295.0, 315.0, 335.0, 355.0, 375.0, 395.0]
    out = []
    for nt in (128, 256, 512):
        spec = fx.t2s2hp("T2S2HP_folded", nstep_MC=mc).replicated(6, temps=temps)
        spec.nl_margin = 1.0
        g = gpu(rep_threads=nt)
        spec.configure(g)
        if mc == 0:
            g.remd_setup([6, 1, 1, 1], [temps, [], [], []], 10)
        for r in range(6):
            g.init_velocities(r); g.neighbor_list(r)
        nreb = g.md_steps(90)
        assert g.last_step_kernel() == f"k_md_rep<{nt}>"
        out.append((nreb, [g.get_state(r)[0] for r in range(6)], g.remd_state()[0].copy() if mc == 0 else None))
    for other in out[1:]:
        assert other[0] == out[0][0]
        if mc == 0:
            assert np.array_equal(other[2], out[0][2])
        for a, b in zip(out[0][1], other[1]):
            assert np.abs(a - b).max() < 1e-7


def test_more_replicas_than_the_gpu_holds_run_in_waves(replica_kernel):
    """700 independent replicas (no exchange: nothing needs them co-resident) against the first 5 of them run alone."""
    if replica_kernel != "k_md_rep":
        pytest.skip("k_md_rep only")
    n = 700
    temps = list(np.linspace(290.0, 390.0, n))
    res = []
    for m in (n, 5):
        spec = fx.t2s2hp("T2S2HP_folded", nstep_MC=0).replicated(m, temps=temps[:m])
        g = gpu()
        spec.configure(g)
        for r in range(m):
            g.init_velocities(r); g.neighbor_list(r)
        g.md_steps(60)
        assert g.last_step_kernel() == ("k_md_rep<128>" if m == n else "k_md_rep<512>") and g.get_step() == 60
        res.append([g.get_state(r)[0] for r in range(5)])
    for a, b in zip(*res):
        assert np.abs(a - b).max() < 1e-7
