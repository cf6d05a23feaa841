"""Standard systems (C1..C5) built from the committed fixtures and the
synthetic generators.  Used by tests, bench.py and smoke(); nothing here reads /root/reference."""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np

from .inputs import ForceField
from .system import SystemSpec

GOLDEN = Path(__file__).resolve().parent.parent / "tests" / "golden"
_cache: dict = {}


def _load():
    if not _cache:
        _cache["json"] = json.loads((GOLDEN / "fixtures.json").read_text())
        _cache["npz"] = np.load(GOLDEN / "fixtures.npz")
    return _cache["json"], _cache["npz"]


def force_field(name: str) -> ForceField:
    js, _ = _load()
    d = dict(js["ff"][name])
    for k in ("bp_paras", "NN_dH", "NN_dS", "NNend_dH", "NNend_dS"):
        if d.get(k) is not None:
            d[k] = np.array(d[k], dtype=np.float64)
    return ForceField(**d)


def array(name: str) -> np.ndarray:
    return np.array(_load()[1][name])


def seqs(name: str) -> list[str]:
    return list(_load()[0]["seqs"][name])


def t2s2hp(coords: str = "T2S2HP_random", nstep_MC: int = 1, **kw) -> SystemSpec:
    """C1: check/input_T2S2HP.toml -- 27-nt hairpin, model 5, DH cutoff_type 2, MC every step."""
    return SystemSpec(seqs=seqs("T2S2HP"), ff=force_field("sis_dih_ReB_NN"), xyz=array(coords),
                      tempK=300.0, rng_seed=123, dt=0.05, bp_model=5, bp_min_loop=3, max_bp_per_nt=1,
                      nstep_MC=nstep_MC, flg_ele=True, ionic_strength=0.15, ele_cutoff_type=2,
                      ele_cutoff=10.0, length_per_charge=4.38178046, **kw)


def cag47(nchain: int = 1, pbc: bool = False, **kw) -> SystemSpec:
    """C2 (1 chain) / the check/input_md.toml system (4 chains, PBC): (CAG)47, rna_cg1.ff,
    model 1 without MC, DH 0.15 M cutoff 50 A, T 310.15 K, dt 0.5."""
    xyz = array("CAG47_4")[: 141 * nchain]
    return SystemSpec(seqs=seqs("CAG47") * nchain, ff=force_field("rna_cg1"), xyz=xyz, tempK=310.15,
                      dt=0.5, bp_model=1, bp_min_loop=4, nstep_MC=0, flg_ele=True, ionic_strength=0.15,
                      ele_cutoff_type=1, ele_cutoff=50.0, length_per_charge=4.38178046,
                      box=(809.972, 809.972, 809.972) if pbc else None, **kw)
