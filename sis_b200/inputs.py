"""Readers for the reference's input formats (host side, set-up only).

The reference parses these in Fortran (`src/read_input.F90`, `src/read_force_field.F90`,
`src/read_fasta.F90`, `src/read_pdb.F90`, `src/read_xyz.F90`, `src/dcd.F90`).  No Fortran
compiler exists in this image, so the test / bench harness reads the same files in Python and
hands the numbers to the C ABI exactly as the Fortran driver would after its own `init_*`
calls.  Nothing here is on the per-step path.
"""
from __future__ import annotations

import struct
import tomllib
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

SEQ_CODE = {"A": 0, "U": 1, "G": 2, "C": 3, "D": 4}  # src/const_idx.F90:31-40

# NNT order, src/const_idx.F90:96-121
NN_NAMES = [
    "GC_CG", "CC_GG", "GA_CU", "CG_GC", "AC_UG", "CA_GU", "AG_UC", "UA_AU", "AU_UA", "AA_UU",
    "GC_UG", "CU_GG", "GG_CU", "CG_GU", "AU_UG", "GA_UU", "UG_GU", "UA_GU", "GG_UU", "GU_UG",
    "AG_UU",
]
NNEND_NAMES = ["AUonAU", "AUonCG", "AUonGU", "GUonCG", "GUonAU", "GUonGU"]  # const_idx.F90:123-132
BP_KEYS = [
    "U0", "bond_k", "bond_r", "angl_k1", "angl_k2", "angl_k3", "angl_k4",
    "angl_theta1", "angl_theta2", "angl_theta3", "angl_theta4",
    "dihd_k1", "dihd_k2", "dihd_phi1", "dihd_phi2",
]


@dataclass
class ForceField:
    """Contents of a `.ff` file as `src/read_force_field.F90:68-372` reads it."""
    bond_k: float = 0.0
    bond_r0: float = 0.0
    angl_type: int = 0          # 0 [potential.angle], 1 [potential.angle_ReB]
    angl_k: float = 0.0
    angl_t0: float = 0.0
    dih_flags: int = 0          # bit0 cos, bit1 exp
    dih_k: float = 0.0
    dih_w: float = 0.0
    dih_p0: float = 0.0
    wca_sigma: float = 10.0
    wca_eps: float = 2.0
    bp_cutoff_energy: float = 0.001   # default, src/read_force_field.F90:144
    coef_dG: float = 1.0
    dH0: float = 0.0
    dS0: float = 0.0
    flg_NNend: bool = False
    dHend0: float = 0.0
    dSend0: float = 0.0
    bp_paras: np.ndarray = field(default_factory=lambda: np.zeros((3, 15)))
    NN_dH: np.ndarray | None = None
    NN_dS: np.ndarray | None = None
    NNend_dH: np.ndarray | None = None
    NNend_dS: np.ndarray | None = None


def read_force_field(path: str | Path) -> ForceField:
    with open(path, "rb") as fh:
        t = tomllib.load(fh)
    pot = t["potential"]
    ff = ForceField()
    ff.bond_k = float(pot["bond"]["k"])
    ff.bond_r0 = float(pot["bond"]["r0"])
    if "angle" in pot and "angle_ReB" in pot:
        raise ValueError("[angle] and [angle_ReB] cannot be specified together in FF file.")
    if "angle" in pot:
        ff.angl_type, node = 0, pot["angle"]
    elif "angle_ReB" in pot:
        ff.angl_type, node = 1, pot["angle_ReB"]
    else:
        raise ValueError("Either [angle] or [angle_ReB] parameters are required in FF file")
    ff.angl_k, ff.angl_t0 = float(node["k"]), float(node["a0"])
    if "dihedral" in pot:
        ff.dih_flags |= 1
        ff.dih_k, ff.dih_p0 = float(pot["dihedral"]["k"]), float(pot["dihedral"]["phi0"])
    if "dihedral_exp" in pot:
        ff.dih_flags |= 2
        d = pot["dihedral_exp"]
        ff.dih_k, ff.dih_w, ff.dih_p0 = float(d["k"]), float(d["w"]), float(d["phi0"])
    bp = pot["basepair"]
    ff.bp_cutoff_energy = float(bp.get("cutoff_energy", 0.001))
    ff.coef_dG = float(bp.get("coef_dG", 1.0))
    ff.dH0 = float(bp.get("dH0", 0.0))
    ff.dS0 = float(bp.get("dS0", 0.0))
    ff.flg_NNend = bool(bp.get("NNend", False))
    ff.dHend0 = float(bp.get("dHend0", 0.0))
    ff.dSend0 = float(bp.get("dSend0", 0.0))
    for row, name in enumerate(("GC", "AU", "GU")):
        if name in bp:
            ff.bp_paras[row] = [float(bp[name][k]) for k in BP_KEYS]
    ff.wca_sigma = float(pot["wca"]["sigma"])
    ff.wca_eps = float(pot["wca"]["epsilon"])
    if "NN" in t:
        nn = t["NN"]
        if "dH" in nn:
            ff.NN_dH = np.array([float(nn["dH"][k]) for k in NN_NAMES])
        if "dS" in nn:
            ff.NN_dS = np.array([float(nn["dS"][k]) for k in NN_NAMES])
    if "NNend" in t:
        ne = t["NNend"]
        ff.NNend_dH = np.array([float(ne["dH"][k]) for k in NNEND_NAMES])
        ff.NNend_dS = np.array([float(ne["dS"][k]) for k in NNEND_NAMES])
    return ff


def read_fasta(path: str | Path) -> list[str]:
    """`src/read_fasta.F90`: one chain per '>' record; sequence may span lines."""
    seqs: list[str] = []
    cur: list[str] | None = None
    for line in Path(path).read_text().splitlines():
        line = line.strip()
        if not line:
            continue
        if line.startswith(">"):
            if cur is not None:
                seqs.append("".join(cur))
            cur = []
        else:
            if cur is None:
                cur = []
            cur.append(line.upper())
    if cur is not None:
        seqs.append("".join(cur))
    return seqs


def read_xyz(path: str | Path) -> np.ndarray:
    """`src/read_xyz.F90`: count line, comment line, then `name x y z`."""
    lines = Path(path).read_text().splitlines()
    n = int(lines[0].split()[0])
    out = np.array([[float(v) for v in ln.split()[1:4]] for ln in lines[2:2 + n]])
    assert out.shape == (n, 3)
    return out


def read_pdb(path: str | Path) -> tuple[np.ndarray, list[str]]:
    """ATOM records: columns 31-54 coordinates (`src/read_pdb.F90`); chains split at TER
    or chain-id change; returns coordinates and per-chain sequences from the residue names."""
    xyz, seqs, cur, last_chain = [], [], [], None
    three = {"ADE": "A", "URA": "U", "GUA": "G", "CYT": "C", "A": "A", "U": "U", "G": "G", "C": "C",
             "RA": "A", "RU": "U", "RG": "G", "RC": "C"}
    for ln in Path(path).read_text().splitlines():
        if ln.startswith("TER") or ln.startswith("END"):
            if cur:
                seqs.append("".join(cur))
                cur = []
            last_chain = None
            continue
        if not (ln.startswith("ATOM") or ln.startswith("HETATM")):
            continue
        chain = ln[21]
        if last_chain is not None and chain != last_chain and cur:
            seqs.append("".join(cur))
            cur = []
        last_chain = chain
        xyz.append([float(ln[30:38]), float(ln[38:46]), float(ln[46:54])])
        cur.append(three.get(ln[17:20].strip(), "D"))
    if cur:
        seqs.append("".join(cur))
    return np.array(xyz), seqs


def read_dcd(path: str | Path) -> tuple[np.ndarray, np.ndarray | None]:
    """CHARMM DCD as written by `src/dcd.F90:287-355,538-596`: returns frames [nframe][natom][3]
    (float32 -> float64) and boxes [nframe][3] or None."""
    data = Path(path).read_bytes()
    pos = 0

    def rec():
        nonlocal pos
        (n,) = struct.unpack_from("<i", data, pos)
        body = data[pos + 4:pos + 4 + n]
        (m,) = struct.unpack_from("<i", data, pos + 4 + n)
        assert n == m, "corrupt DCD record"
        pos += 8 + n
        return body

    hdr = rec()
    assert hdr[:4] == b"CORD"
    icntrl = struct.unpack_from("<20i", hdr, 4)
    has_cell = icntrl[10] == 1  # control word 11, src/dcd.F90:158-166
    rec()  # title
    (natom,) = struct.unpack("<i", rec())
    frames, boxes = [], []
    while pos < len(data):
        if has_cell:
            c = struct.unpack("<6d", rec())
            boxes.append([c[0], c[2], c[5]])  # (a, 0, b, 0, 0, c): src/dcd.F90:563-572
        x = np.frombuffer(rec(), dtype="<f4")
        y = np.frombuffer(rec(), dtype="<f4")
        z = np.frombuffer(rec(), dtype="<f4")
        frames.append(np.stack([x, y, z], axis=1).astype(np.float64))
    return np.array(frames), (np.array(boxes) if has_cell else None)


def write_dcd(path: str | Path, frames: np.ndarray, box: np.ndarray | None = None) -> None:
    """Minimal CHARMM DCD writer in the layout of `src/dcd.F90:287-355,538-596`."""
    frames = np.asarray(frames, dtype=np.float32)
    nframe, natom, _ = frames.shape

    def rec(b: bytes) -> bytes:
        return struct.pack("<i", len(b)) + b + struct.pack("<i", len(b))

    icntrl = [0] * 20
    icntrl[0], icntrl[1], icntrl[2], icntrl[3] = nframe, 0, 1, nframe
    icntrl[10] = 1 if box is not None else 0
    icntrl[19] = 24
    out = [rec(b"CORD" + struct.pack("<20i", *icntrl))]
    title = b"sis-b200".ljust(80)
    out.append(rec(struct.pack("<i", 1) + title))
    out.append(rec(struct.pack("<i", natom)))
    for f in frames:
        if box is not None:
            out.append(rec(struct.pack("<6d", box[0], 0.0, box[1], 0.0, 0.0, box[2])))
        for k in range(3):
            out.append(rec(np.ascontiguousarray(f[:, k]).tobytes()))
    Path(path).write_bytes(b"".join(out))


def read_input_toml(path: str | Path) -> dict:
    """The TOML input of `src/read_input.F90:98-1280` as a plain dict (keys are
    case-sensitive in the reference; no normalisation is done here)."""
    with open(path, "rb") as fh:
        return tomllib.load(fh)


def read_ss_pairs(path, kind: str, nmp: int | None = None) -> list[tuple[int, int]]:
    """Base pairs of a secondary-structure file as `read_ss` reads it (`src/read_ss.F90:28-148`):
    kind = "ct" (header line with the nucleotide count, then `imp nt _ _ jmp _`), "bpseq"
    (`imp nt jmp`) or "bpl" (`imp jmp` per line).  Returns 1-based (imp, jmp) with imp < jmp; the
    filters of read_ss (chain ends, minimum loop, isolated pairs) are applied by the library."""
    lines = [l.split() for l in Path(path).read_text().splitlines() if l.split()]
    pairs = []
    if kind == "ct":
        n = int(lines[0][0])
        if nmp is not None and n != nmp:
            raise ValueError("CT file format error. The total number of nucleotide does not match.")
        for k, f in enumerate(lines[1:n + 1], start=1):
            imp, jmp = int(f[0]), int(f[4])
            if imp != k:
                raise ValueError(f"CT file format error. Nucleotide number does not match for Nucleotide {k}")
            if jmp != 0:
                pairs.append((min(imp, jmp), max(imp, jmp)))
    elif kind == "bpseq":
        for k, f in enumerate(lines, start=1):
            imp, jmp = int(f[0]), int(f[2])
            if imp != k:
                raise ValueError(f"BPSEQ file format error. Nucleotide number does not match for Nucleotide {k}")
            if jmp != 0:
                pairs.append((min(imp, jmp), max(imp, jmp)))
    elif kind == "bpl":
        for f in lines:
            imp, jmp = int(f[0]), int(f[1])
            if jmp != 0:
                pairs.append((min(imp, jmp), max(imp, jmp)))
    else:
        raise ValueError(f"unknown secondary-structure file kind {kind}")
    return sorted(set(pairs))
