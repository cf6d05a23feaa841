"""Test-only host build of the product's force/list code (tests/csrc/hostcheck.cu)."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from sis_b200._capi import CApi, _d, _i

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent
SO = HERE / "csrc" / "libhostcheck.so"


def build():
    srcs = [HERE / "csrc" / "hostcheck.cu"] + sorted((ROOT / "sis_b200" / "csrc").glob("*.cu*"))
    if SO.exists() and all(SO.stat().st_mtime >= s.stat().st_mtime for s in srcs):
        return
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-Xcompiler",
                    "-fPIC", "-shared", "-o", str(SO), str(srcs[0]), "-ldl"], check=True, capture_output=True)


class HostCheck(CApi):
    def __init__(self):
        build()
        lib = C.CDLL(str(SO))
        self.hc = C.c_void_p()
        assert lib.hc_create(C.byref(self.hc)) == 0
        lib.hc_handle.restype = C.c_void_p
        super().__init__(lib, "sisgpu_", C.c_void_p(lib.hc_handle(self.hc)))

    def finalize_setup(self):
        if self.lib.hc_prepare(self.hc) != 0:
            raise RuntimeError(self._err().decode())

    def set_state(self, irep, xyz=None, velos=None, accels=None):
        self.xyz = np.array(xyz, dtype=np.float64)

    def build_lists(self, all_pairs=False):
        self.xyz = np.ascontiguousarray(self.xyz)
        assert self.lib.hc_build_lists(self.hc, self.xyz.ctypes.data_as(C.POINTER(C.c_double)), int(all_pairs)) == 0

    def list_sizes(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self.lib.hc_list_sizes(self.hc, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def lists(self):
        nw, ne, nb = self.list_sizes()
        wca, ele, bp = np.zeros((nw, 2), np.int32), np.zeros((ne, 2), np.int32), np.zeros((nb, 3), np.int32)
        coef = np.zeros((nb, 2))
        self.lib.hc_get_lists(self.hc, _i(wca)[0], _i(ele)[0], _i(bp)[0], _d(coef)[0])
        return wca, ele, bp, coef

    def force_terms(self, irep, mask):
        f = np.zeros((self.nmp, 3))
        self.lib.hc_force(self.hc, _d(self.xyz)[0], C.c_uint(mask), _d(f)[0])
        return f

    def force_bp_role_major(self):
        f = np.zeros((self.nmp, 3))
        self.lib.hc_force_bp_role_major(self.hc, _d(self.xyz)[0], _d(f)[0])
        return f

    def pair_scalar(self, d2, wca, ele, mixed):
        self.lib.hc_pair_scalar.restype = C.c_double
        return self.lib.hc_pair_scalar(self.hc, C.c_double(d2), C.c_int(int(wca)), C.c_int(int(ele)), C.c_int(int(mixed)))

    def force_bp_roles(self):
        f = np.zeros((self.nmp, 3))
        self.lib.hc_force_bp_roles(self.hc, _d(self.xyz)[0], _d(f)[0])
        return f

    def gauss(self, seed, stream, step, bead):
        self.lib.hc_gauss.restype = C.c_double
        return np.array([self.lib.hc_gauss(C.c_uint64(seed), C.c_uint(stream), C.c_uint64(step), C.c_uint(bead), c)
                         for c in range(3)])

    def uniform(self, seed, stream, counter):
        self.lib.hc_uniform.restype = C.c_double
        return self.lib.hc_uniform(C.c_uint64(seed), C.c_uint(stream), C.c_uint64(counter))

    def remd_setup(self, nrep, values, nstep_exchange, rank=0, nranks=1):
        nr, vals = self._remd_args(nrep, values)
        if self.lib.hc_remd_setup(self.hc, _i(nr)[0], _d(vals)[0], C.c_int(nstep_exchange), C.c_int(rank), C.c_int(nranks)) != 0:
            raise RuntimeError(self._err().decode())

    def remd_exchange(self, replica_energies, uniforms):
        """-> (n_attempt, n_accept, velocity scale per local replica, [tempK, ionic, force, rg0] per local replica)"""
        na, nc = C.c_int(), C.c_int()
        scale, par = np.zeros(self.nrep_proc), np.zeros((self.nrep_proc, 4))
        re, u = np.ascontiguousarray(replica_energies, dtype=np.float64), np.ascontiguousarray(uniforms, dtype=np.float64)
        assert self.lib.hc_remd_exchange(self.hc, _d(re)[0], _d(u)[0], C.byref(na), C.byref(nc), _d(scale)[0], _d(par)[0]) == 0
        return na.value, nc.value, scale, par

    def coef(self, imp):
        out = np.zeros(4)
        self.lib.hc_coef(self.hc, C.c_int(imp), _d(out)[0])
        return out

    def rep_params(self):
        out = np.zeros(4)
        self.lib.hc_rep_params(self.hc, _d(out)[0])
        return dict(zip(["ele_coef", "lambdaD", "ele_cutoff", "diele"], out))
