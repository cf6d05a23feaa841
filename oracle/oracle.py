"""ctypes wrapper of the CPU oracle (oracle/sis_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package sis_b200 never imports this.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from sis_b200._capi import CApi, SisError, _d, _i

HERE = Path(__file__).resolve().parent


def build(force: bool = False) -> None:
    so = HERE / "libsis_oracle.so"
    src = HERE / "sis_oracle.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE), "all"], check=True, capture_output=True)


class Oracle(CApi):
    def __init__(self, omp: bool = False):
        build()
        lib = C.CDLL(str(HERE / ("libsis_oracle_omp.so" if omp else "libsis_oracle.so")))
        lib.oracle_create.restype = C.c_void_p
        h = C.c_void_p(lib.oracle_create())
        super().__init__(lib, "oracle_", h)

    def close(self):
        if self.h:
            self.lib.oracle_destroy.argtypes = [C.c_void_p]
            self.lib.oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # oracle-only helpers
    def bp_lookup(self, imp, jmp):
        a, b = C.c_int(), C.c_int()
        self._call("bp_lookup", C.c_int(imp), C.c_int(jmp), C.byref(a), C.byref(b))
        return a.value, b.value

    def md_step(self, irep=0, gauss=None):
        reb = C.c_int()
        self._call("md_step", C.c_int(irep), _d(gauss), C.byref(reb))
        return reb.value

    def md_run(self, nsteps):
        n = C.c_int64()
        self._call("md_run", C.c_int(nsteps), C.byref(n))
        return n.value

    def set_save_bp(self, flg):
        self._call("set_save_bp", C.c_int(int(flg)))

    def num_threads(self):
        self.lib.oracle_num_threads.restype = C.c_int
        return self.lib.oracle_num_threads()

    def nl_seconds(self, reset=False):
        self.lib.oracle_nl_seconds.restype = C.c_double
        self.lib.oracle_nl_seconds.argtypes = [C.c_void_p, C.c_int]
        return self.lib.oracle_nl_seconds(self.h, int(reset))

    def md_coef(self, irep, imp):
        c = np.zeros(4)
        self._call("md_coef", C.c_int(irep), C.c_int(imp), _d(c))
        return c

    def set_ele_calc(self, tempK, ionstr):
        out = np.zeros(6)
        self._call("set_ele_calc", C.c_double(tempK), C.c_double(ionstr), _d(out))
        return dict(zip(["diele", "lb", "Zp", "ele_coef", "lambdaD", "diele_dTcoef"], out))

    def get_ele(self, irep=0):
        out = np.zeros(4)
        self._call("get_ele", C.c_int(irep), _d(out))
        return dict(zip(["ele_coef", "lambdaD", "ele_cutoff", "diele"], out))

    def remd_setup(self, nrep, values, nstep_exchange):
        nr, vals = self._remd_args(nrep, values)
        self._call("remd_setup", _i(nr), _d(vals), C.c_int(nstep_exchange))

    def gauss(self, seed, stream, step, bead):
        g = (C.c_double * 3)()
        self.lib.oracle_gauss.restype = None
        self.lib.oracle_gauss(C.c_uint64(seed), C.c_uint32(stream), C.c_uint64(step), C.c_uint32(bead), g)
        return np.array(g[:])

    def uniform(self, seed, stream, counter):
        self.lib.oracle_uniform.restype = C.c_double
        return self.lib.oracle_uniform(C.c_uint64(seed), C.c_uint32(stream), C.c_uint64(counter))
