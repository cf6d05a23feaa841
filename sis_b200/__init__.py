"""sis_b200 -- B200-native (sm_100a) implementation of the SIS per-step hot path.

The product is `libsisgpu.so` (CUDA kernels + the C ABI of `include/sisgpu.h`); this
package is the thin host-side mirror used where the reference's Fortran driver is not
available (tests, bench).  There is no CPU fallback.
"""
from .system import SystemSpec, from_input_toml, remd_from_toml  # noqa: F401
from .inputs import read_force_field, read_fasta, read_pdb, read_xyz, read_dcd, read_ss_pairs  # noqa: F401

__all__ = ["SystemSpec", "from_input_toml", "remd_from_toml", "read_force_field", "read_fasta", "read_pdb",
           "read_xyz", "read_dcd", "read_ss_pairs"]
