"""Generate tests/golden/fixtures.npz + fixtures.json from the reference tree.

Run once in the build container (`python tests/golden/make_fixtures.py`); the outputs are
committed because /root/reference does not exist on the GPU box.  Only DATA is extracted
(force-field numbers, sequences, coordinates, the reference's golden energies) -- no source.
"""
import dataclasses
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from sis_b200.inputs import read_dcd, read_fasta, read_force_field, read_pdb, read_xyz  # noqa: E402

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


def ff_dict(path):
    d = dataclasses.asdict(read_force_field(path))
    return {k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in d.items()}


def main():
    js = {"ff": {}, "seqs": {}, "source": "Hori-Lab/SIS @ dca93e8 (check/, *.ff)"}
    for name, p in {
        "sis_dih_ReB_NN": REF / "check/sis_dih_ReB_NN.ff",
        "rna_cg1": REF / "rna_cg1.ff",
        "rna_cg1_para2": REF / "rna_cg1_para2.ff",
        "sis": REF / "sis.ff",
        "rna_cg1_para3": REF / "check/rna_cg1_para3.ff",
    }.items():
        try:
            js["ff"][name] = ff_dict(p)
        except Exception as e:  # stale-format files are skipped, not fatal
            print(f"skip {p}: {e}")
    js["seqs"]["T2S2HP"] = read_fasta(REF / "check/T2S2HP.fasta")
    js["seqs"]["CAG47"] = read_fasta(REF / "check/test_DCD_CAG47/CAG47.fasta")
    js["seqs"]["nick_dumbbell"] = read_fasta(REF / "check/nick_restraint/nick_dumbbell.fasta")
    arrays = {}
    arrays["T2S2HP_random"] = read_xyz(REF / "check/T2S2HP_random.xyz")
    arrays["T2S2HP_folded"], _ = read_pdb(REF / "check/T2S2HP_folded.pdb")
    arrays["CAG47_4"], s4 = read_pdb(REF / "check/CAG47_4.pdb")
    js["seqs"]["CAG47_4_pdb"] = s4
    arrays["nick_dumbbell"], _ = read_pdb(REF / "check/nick_restraint/nick_dumbbell.pdb")
    fr, box = read_dcd(REF / "check/test_DCD_CAG47/CAG47.dcd")
    arrays["CAG47_dcd"] = fr.astype(np.float32)
    arrays["CAG47_energy_out"] = np.loadtxt(REF / "check/test_DCD_CAG47/CAG47.energy.out")
    fr, box = read_dcd(REF / "check/test_md.dcd")
    arrays["test_md_dcd"] = fr.astype(np.float32)
    arrays["test_md_box"] = box
    np.savez_compressed(OUT / "fixtures.npz", **arrays)
    (OUT / "fixtures.json").write_text(json.dumps(js, indent=1))
    for k, v in arrays.items():
        print(k, v.shape)


if __name__ == "__main__":
    main()


def make_ss_pairs():
    """tests/golden/ss_pairs.json: the base pairs of check/T2S2HP.ct (1-based i < j)."""
    pairs = []
    for line in (REF / "check/T2S2HP.ct").read_text().splitlines()[1:]:
        f = line.split()
        if len(f) >= 5 and int(f[4]) > int(f[0]):
            pairs.append([int(f[0]), int(f[4])])
    # check/nick_restraint: the base-pair list (bpl) and the two Sigmoid-to-bead restraints of dummy_restraint.txt
    bpl = [[int(f) for f in line.split()[:2]] for line in (REF / "check/nick_restraint/nick_dumbbell.bpl").read_text().splitlines()
           if line.strip() and not line.lstrip().startswith("#")]
    rest = []
    for line in (REF / "check/nick_restraint/dummy_restraint.txt").read_text().splitlines():
        f = line.split()
        if f and f[0] in ("Sigmoid", "Sigmoid-to-bead"):
            # (imp, jmp, type 1 both beads / 2 reference bead force-free, rcut, eps, d, s): src/read_restraint.F90
            rest.append([int(f[1]), int(f[2]), 1 if f[0] == "Sigmoid" else 2, float(f[6]), float(f[3]), float(f[4]), float(f[5])])
    (Path(__file__).resolve().parent / "ss_pairs.json").write_text(json.dumps(
        {"T2S2HP_ct_pairs": pairs, "source": "check/T2S2HP.ct (column 1, column 5), pairs with j > i",
         "nick_dumbbell_bpl": bpl, "nick_dumbbell_restraints": rest,
         "nick_source": "check/nick_restraint/nick_dumbbell.bpl, dummy_restraint.txt"}, indent=1))
