"""SystemSpec: everything the Fortran driver knows after its `init_*` phase
(`src/main.F90:55-154`), gathered host-side so it can be handed to the C ABI.

`SystemSpec.configure(backend)` issues the same sequence of `*_set_*` calls the Fortran
shim (`sis_b200/fortran/sisgpu_mod.F90`) makes.  `backend` is any object exposing the
set-up entry points of `include/sisgpu.h` (the product library, or -- in tests only -- the
CPU oracle, whose C API mirrors them one to one).
"""
from __future__ import annotations

import copy
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

from .inputs import (SEQ_CODE, ForceField, read_fasta, read_force_field, read_input_toml,
                     read_pdb, read_xyz)

MOLT_RNA, MOLT_CIRCRNA = 1, 2
JOUL2KCAL_MOL = (1.0 / 4184.0) * 6.02214076e23  # src/const_phys.F90:13-17
PN2KCALMOLA = JOUL2KCAL_MOL * 1.0e-22            # src/read_input.F90:552


@dataclass
class SystemSpec:
    seqs: list[str]
    ff: ForceField
    moltypes: list[int] | None = None
    xyz: np.ndarray | None = None             # [nmp][3] initial coordinates
    # [Condition]
    tempK: float = 300.0
    rng_seed: int = 123
    temp_independent: int = 0
    tempK_ref: float = 0.0
    # [MD]
    integrator: int = 1                       # 1 GJF-2GJ, 2 BD-EM
    dt: float = 0.05
    viscosity_Pas: float = 1.0e-5
    nl_margin: float = 10.0
    # [Basepair]
    bp_model: int = 5
    bp_min_loop: int = 3
    max_bp_per_nt: int = 1
    nstep_MC: int = 1
    bp_restrict: object = None      # ct / bpseq / bpl pairs, 1-based (imp, jmp); None = all pairwise
    # [Electrostatic]
    flg_ele: bool = False
    ionic_strength: float = 0.15
    ele_cutoff_type: int = 1
    ele_cutoff: float = 50.0
    length_per_charge: float = 4.38178046
    excl_bond: bool = True
    excl_angle: bool = True
    no_charge: list[int] = field(default_factory=list)   # 1-based ids
    dummy_has_charge: bool = False
    # [PBC_box]
    box: tuple[float, float, float] | None = None
    # biases
    bias_ss_force: float | None = None        # kcal/mol/A
    stage: tuple[float, float] | None = None  # sigma, eps
    twz_dcf_pairs: list[tuple[int, int]] = field(default_factory=list)
    twz_dcf_forces: list[tuple[float, float, float]] = field(default_factory=list)  # kcal/mol/A
    # Force-Ramp traps (src/force_tweezers.F90:37-61): pairs [n][2], k [n][2], init / velo [n][2][3]
    twz_fr_pairs: list[tuple[int, int]] = field(default_factory=list)
    twz_fr_k: list[tuple[float, float]] = field(default_factory=list)
    twz_fr_init: list | None = None
    twz_fr_velo: list | None = None
    bias_rg: tuple[int, float, float] | None = None  # pott, k, rg0
    restraints: list[tuple[int, int, int, float, float, float, float]] = field(default_factory=list)
    frozen: list[int] = field(default_factory=list)       # 1-based ids
    # replicas
    nrep_proc: int = 1
    rep_tempK: list[float] | None = None      # per local replica override
    rep_ionic: list[float] | None = None

    # ---- derived ------------------------------------------------------------
    @property
    def nmp(self) -> int:
        return sum(len(s) for s in self.seqs)

    @property
    def nchains(self) -> int:
        return len(self.seqs)

    def seq_codes(self) -> np.ndarray:
        return np.array([SEQ_CODE[c] for s in self.seqs for c in s], dtype=np.int32)

    def has_charge(self) -> np.ndarray:
        """`src/init_ele.F90:45-62`."""
        hc = np.ones(self.nmp, dtype=np.int32)
        if not self.dummy_has_charge:
            hc[self.seq_codes() == SEQ_CODE["D"]] = 0
        for i in self.no_charge:
            hc[i - 1] = 0
        return hc

    def is_frozen(self) -> np.ndarray:
        fz = np.zeros(self.nmp, dtype=np.int32)
        for i in self.frozen:
            fz[i - 1] = 1
        return fz

    def replicated(self, nrep: int, temps=None, ionic=None) -> "SystemSpec":
        s = copy.copy(self)
        s.nrep_proc = nrep
        s.rep_tempK = list(temps) if temps is not None else None
        s.rep_ionic = list(ionic) if ionic is not None else None
        return s

    # ---- the call sequence of the Fortran shim -----------------------------------
    def configure(self, be) -> None:
        ff = self.ff
        nmp_chain = np.array([len(s) for s in self.seqs], dtype=np.int32)
        molt = np.array(self.moltypes or [MOLT_RNA] * self.nchains, dtype=np.int32)
        be.set_topology(self.nmp, self.nchains, nmp_chain, molt, self.seq_codes(),
                        self.has_charge(), self.is_frozen(), None)
        be.set_replicas(self.nrep_proc)
        be.set_ff_bonded(ff.bond_k, ff.bond_r0, ff.angl_type, ff.angl_k, ff.angl_t0,
                         ff.dih_flags, ff.dih_k, ff.dih_w, ff.dih_p0)
        be.set_ff_wca(ff.wca_sigma, ff.wca_eps)
        be.set_ff_bp(self.bp_model, self.bp_min_loop, self.max_bp_per_nt if self.nstep_MC > 0 else 0,
                     self.nstep_MC, ff.bp_cutoff_energy, ff.bp_paras)
        if self.bp_restrict is not None:
            be.set_bp_restrict(self.bp_restrict)
        if self.bp_model == 5:
            if ff.NN_dH is None or ff.NN_dS is None:
                raise ValueError("[NN.dH]/[NN.dS] required in FF file for bp model 5")
            be.set_ff_nn(ff.NN_dH, ff.NN_dS, ff.dH0, ff.dS0, int(ff.flg_NNend), ff.NNend_dH,
                         ff.NNend_dS, ff.dHend0, ff.dSend0)
        be.set_ele(int(self.flg_ele), self.ele_cutoff_type, self.ele_cutoff, int(self.excl_bond),
                   int(self.excl_angle), self.length_per_charge, self.temp_independent,
                   self.tempK_ref)
        be.set_pbc(int(self.box is not None), np.array(self.box or (0.0, 0.0, 0.0)))
        be.set_md(self.integrator, self.dt, self.viscosity_Pas, self.nl_margin, self.rng_seed)
        if self.bias_ss_force is not None:
            be.set_bias_ss(1, self.bias_ss_force)
        if self.stage is not None:
            be.set_stage(1, *self.stage)
        twz_mag = 0.0
        if self.twz_dcf_pairs:
            f = np.array(self.twz_dcf_forces, dtype=np.float64)
            # the reference stores direction * magnitude; REMD rescales the magnitude
            # (src/job_md.F90:555-558).  Without a force replica the "direction" carries the
            # full vector and the magnitude is 1.
            be.set_tweezers_dcf(len(self.twz_dcf_pairs), np.array(self.twz_dcf_pairs, dtype=np.int32), f)
            twz_mag = 1.0
        if self.twz_fr_pairs:
            n = len(self.twz_fr_pairs)
            be.set_tweezers_fr(n, np.array(self.twz_fr_pairs, dtype=np.int32),
                               np.array(self.twz_fr_k, dtype=np.float64).reshape(n, 2),
                               np.array(self.twz_fr_init, dtype=np.float64).reshape(n, 6),
                               np.array(self.twz_fr_velo, dtype=np.float64).reshape(n, 6))
        rg0 = 0.0
        if self.bias_rg is not None:
            be.set_bias_rg(1, self.bias_rg[0], self.bias_rg[1])
            rg0 = self.bias_rg[2]
        if self.restraints:
            r = self.restraints
            be.set_restraints(len(r), np.array([[a[0], a[1], a[2]] for a in r], dtype=np.int32),
                              np.array([a[3] for a in r]), np.array([[a[4], a[5], a[6]] for a in r]))
        for irep in range(self.nrep_proc):
            T = self.rep_tempK[irep] if self.rep_tempK else self.tempK
            I = self.rep_ionic[irep] if self.rep_ionic else self.ionic_strength
            be.set_replica_params(irep, T, I, twz_mag, rg0)
        be.finalize_setup()
        if self.xyz is not None:
            for irep in range(self.nrep_proc):
                be.set_state(irep, self.xyz, None, None)


def from_input_toml(path: str | Path, ff_override: str | Path | None = None) -> tuple[SystemSpec, dict]:
    """Build a SystemSpec from a current-format SIS input file (`src/read_input.F90:98-1280`;
    only the keys the hot path consumes).  Returns the spec and the raw TOML dict."""
    path = Path(path)
    t = read_input_toml(path)
    base = path.parent
    fin = t["Files"]["In"]
    ff = read_force_field(ff_override if ff_override else base / fin["ff"])
    xyz, seqs = None, None
    if "fasta" in fin:
        seqs = read_fasta(base / fin["fasta"])
    if "pdb_ini" in fin:
        xyz, pdb_seqs = read_pdb(base / fin["pdb_ini"])
        seqs = seqs or pdb_seqs
    elif "xyz_ini" in fin:
        xyz = read_xyz(base / fin["xyz_ini"])
    spec = SystemSpec(seqs=seqs, ff=ff, xyz=xyz)
    for kind in ("ct", "bpseq", "bpl"):      # src/read_input.F90 [Files.In]; read_ss takes the first that is set
        if kind in fin:
            from .inputs import read_ss_pairs
            spec.bp_restrict = read_ss_pairs(base / fin[kind], kind, spec.nmp if kind != "bpl" else None)
            break
    cond = t.get("Condition", {})
    spec.tempK = float(cond.get("tempK", spec.tempK))
    spec.rng_seed = int(cond.get("rng_seed", spec.rng_seed))
    spec.temp_independent = int(cond.get("temp_independent", 0))
    spec.tempK_ref = float(cond.get("tempK_ref", 0.0))
    md = t.get("MD", {})
    spec.integrator = {"GJF-2GJ": 1, "BD-EM": 2}.get(md.get("integrator", "GJF-2GJ"), 1)
    spec.dt = float(md.get("dt", spec.dt))
    spec.viscosity_Pas = float(md.get("viscosity_Pas", 1.0e-5))
    m = float(md.get("neighbor_list_margin", -1.0))
    spec.nl_margin = 10.0 if m < 0 else m  # src/read_input.F90:398-404
    bp = t.get("Basepair")
    if bp is None:  # src/read_input.F90:674-679
        spec.bp_model, spec.nstep_MC, spec.max_bp_per_nt, spec.bp_min_loop = 5, 1, 1, 3
    else:
        spec.bp_model = int(bp["model"])
        spec.nstep_MC = int(bp.get("nstep_MC", 1))
        spec.max_bp_per_nt = int(bp.get("max_bp_per_nt", 1)) if spec.nstep_MC > 0 else 0
        spec.bp_min_loop = int(bp.get("min_loop", 3))
    el = t.get("Electrostatic")
    if el is not None:
        spec.flg_ele = True
        spec.ionic_strength = float(el["ionic_strength"])
        spec.ele_cutoff_type = int(el.get("cutoff_type", 1))
        spec.ele_cutoff = float(el["cutoff"])
        spec.length_per_charge = float(el["length_per_charge"])
        spec.no_charge = [int(i) for i in el.get("no_charge", [])]
        spec.dummy_has_charge = bool(el.get("dummy_has_charge", False))
        spec.excl_bond = bool(el.get("exclude_covalent_bond_pairs", True))
        spec.excl_angle = bool(el.get("exclude_covalent_angle_pairs", True))
    box = t.get("PBC_box")
    if box is not None:
        spec.box = (float(box["x"]), float(box["y"]), float(box["z"]))
    if "Bias_SS" in t:
        spec.bias_ss_force = float(t["Bias_SS"]["force_pN"]) * PN2KCALMOLA
    if "Stage" in t:
        spec.stage = (float(t["Stage"]["sigma"]), float(t["Stage"]["epsilon"]))
    tw = t.get("Tweezers", {}).get("Dual_Constant_Force")
    if tw is not None:  # src/read_input.F90 [Tweezers.Dual_Constant_Force]: forces in pN
        spec.twz_dcf_pairs = [tuple(int(x) for x in p) for p in tw["id_pairs"]]
        spec.twz_dcf_forces = [tuple(float(x) * PN2KCALMOLA for x in f) for f in tw["forces_pN"]]
    return spec, t


def remd_from_toml(t: dict) -> dict | None:
    """The [Replica] block (`src/read_input.F90:443-620`) in the argument form of `remd_setup`:
    {"nrep": [TEMP, TWZDCF, ION, RG], "values": [[...], [...], [...], [...]], "nstep_exchange": n}.
    A variable with nrep_* of 0 or 1 is no replica variable; forces are converted from pN to
    kcal/mol/A as the reference does (`:552`)."""
    rep = t.get("Replica")
    if rep is None:
        return None
    keys = [("nrep_temp", "Temperature", 1.0), ("nrep_force", "Force", PN2KCALMOLA),
            ("nrep_ion", "Ionic_strength", 1.0), ("nrep_rg", "Rg", 1.0)]
    nrep, values = [], []
    for nkey, block, unit in keys:
        n = int(rep.get(nkey, 1))
        if n > 1:
            if block not in rep:
                raise ValueError(f"Error in input file: [Replica.{block}] is required.")
            values.append([float(rep[block][str(i)]) * unit for i in range(1, n + 1)])
            nrep.append(n)
        else:
            values.append([])
            nrep.append(1)
    if all(n == 1 for n in nrep):
        raise ValueError("Error in input file: [Replica] needs any of nrep_temp, nrep_force, nrep_ion, or nrep_rg.")
    return {"nrep": nrep, "values": values, "nstep_exchange": int(rep.get("nstep_exchange", 0)),
            "exchange": bool(rep.get("exchange", True))}
