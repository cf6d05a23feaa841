"""ctypes binding of the C ABI declared in `include/sisgpu.h`.

`CApi` binds a shared library whose entry points are `<prefix><name>`; the product
library uses the prefix `sisgpu_`.  The product path fails loudly if `libsisgpu.so` is
missing: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

ENE_N = 11
ENE_NAMES = ["total", "bond", "angl", "dihe", "bp", "exv", "ele", "stage", "twz", "rg", "rest"]

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def _d(a):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a.ctypes.data_as(_dp), a


def _i(a):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a.ctypes.data_as(_ip), a


class SisError(RuntimeError):
    pass


class CApi:
    """Thin, explicit wrapper: one method per C entry point, numpy in / numpy out."""

    def __init__(self, lib: C.CDLL, prefix: str, handle):
        self.lib, self.prefix, self.h = lib, prefix, handle
        self.nmp = 0
        self.nrep_proc = 1
        self.nrep_all = 1
        self.ndim_replica = 0
        self._err = getattr(lib, prefix + "last_error")
        self._err.restype = C.c_char_p

    def _call(self, name, *args):
        fn = getattr(self.lib, self.prefix + name)
        fn.restype = C.c_int
        keep, conv = [], []
        for a in args:
            if isinstance(a, tuple):       # (pointer, owner)
                conv.append(a[0]); keep.append(a[1])
            else:
                conv.append(a)
        rc = fn(self.h, *conv)
        if rc != 0:
            raise SisError(f"{self.prefix}{name}: {self._err().decode()}")
        return rc

    # ---- set-up ------------------------------------------------------------
    def set_topology(self, nmp, nchains, nmp_chain, moltypes, seq, has_charge, is_frozen, mass):
        self.nmp = int(nmp)
        self._call("set_topology", C.c_int(nmp), C.c_int(nchains), _i(nmp_chain), _i(moltypes), _i(seq),
                   _i(has_charge), _i(is_frozen), _d(mass))

    def set_replicas(self, nrep_proc):
        self.nrep_proc = int(nrep_proc)
        self.nrep_all = max(self.nrep_all, self.nrep_proc)
        self._call("set_replicas", C.c_int(nrep_proc))

    def set_ff_bonded(self, bond_k, bond_r0, angl_type, angl_k, angl_t0, dih_flags, dih_k, dih_w, dih_p0):
        self._call("set_ff_bonded", C.c_double(bond_k), C.c_double(bond_r0), C.c_int(angl_type),
                   C.c_double(angl_k), C.c_double(angl_t0), C.c_int(dih_flags), C.c_double(dih_k),
                   C.c_double(dih_w), C.c_double(dih_p0))

    def set_ff_wca(self, sigma, eps):
        self._call("set_ff_wca", C.c_double(sigma), C.c_double(eps))

    def set_ff_bp(self, model, min_loop, max_bp_per_nt, nstep_MC, cutoff_energy, bp_paras):
        self._call("set_ff_bp", C.c_int(model), C.c_int(min_loop), C.c_int(max_bp_per_nt), C.c_int(nstep_MC),
                   C.c_double(cutoff_energy), _d(bp_paras))

    def set_bp_restrict(self, pairs):
        """ct / bpseq / bpl restriction of the bp map: 1-based (imp, jmp) pairs; None removes it."""
        if pairs is None:
            self._call("set_bp_restrict", C.c_int(-1), None)
            return
        a = np.ascontiguousarray(np.asarray(pairs, dtype=np.int32).reshape(-1, 2))
        self._call("set_bp_restrict", C.c_int(a.shape[0]), _i(a))

    def set_ff_nn(self, NN_dH, NN_dS, dH0, dS0, flg_NNend, NNend_dH, NNend_dS, dHend0, dSend0):
        self._call("set_ff_nn", _d(NN_dH), _d(NN_dS), C.c_double(dH0), C.c_double(dS0), C.c_int(flg_NNend),
                   _d(NNend_dH), _d(NNend_dS), C.c_double(dHend0), C.c_double(dSend0))

    def set_ele(self, flg_ele, cutoff_type, cutoff_inp, excl_bond, excl_angle, length_per_charge,
                temp_independent, tempK_ref):
        self._call("set_ele", C.c_int(flg_ele), C.c_int(cutoff_type), C.c_double(cutoff_inp),
                   C.c_int(excl_bond), C.c_int(excl_angle), C.c_double(length_per_charge),
                   C.c_int(temp_independent), C.c_double(tempK_ref))

    def set_pbc(self, flg_pbc, box):
        self._call("set_pbc", C.c_int(flg_pbc), _d(box))

    def set_md(self, integrator, dt, viscosity_Pas, nl_margin, seed):
        self._call("set_md", C.c_int(integrator), C.c_double(dt), C.c_double(viscosity_Pas),
                   C.c_double(nl_margin), C.c_uint64(seed))

    def set_bias_ss(self, flg, force):
        self._call("set_bias_ss", C.c_int(flg), C.c_double(force))

    def set_stage(self, flg, sigma, eps):
        self._call("set_stage", C.c_int(flg), C.c_double(sigma), C.c_double(eps))

    def set_tweezers_dcf(self, n, pairs, direction):
        self._call("set_tweezers_dcf", C.c_int(n), _i(pairs), _d(direction))

    def set_tweezers_fr(self, n, pairs, k, init, velo):
        self._call("set_tweezers_fr", C.c_int(n), _i(pairs), _d(k), _d(init), _d(velo))

    def set_bias_rg(self, flg, pott, k):
        self._call("set_bias_rg", C.c_int(flg), C.c_int(pott), C.c_double(k))

    def set_restraints(self, n, ids, rcut, para):
        self._call("set_restraints", C.c_int(n), _i(ids), _d(rcut), _d(para))

    def set_replica_params(self, irep, tempK, ionic_strength, twz_force=0.0, rg0=0.0):
        self._call("set_replica_params", C.c_int(irep), C.c_double(tempK), C.c_double(ionic_strength),
                   C.c_double(twz_force), C.c_double(rg0))

    def finalize_setup(self):
        self._call("finalize_setup")

    # ---- state -------------------------------------------------------------
    def set_state(self, irep, xyz=None, velos=None, accels=None):
        self._call("set_state", C.c_int(irep), _d(xyz), _d(velos), _d(accels))

    def get_state(self, irep, out=None):
        """-> xyz, velos, accels ([nmp][3] each); `out` = three caller buffers (e.g. pinned) to fill."""
        x, v, a = out if out is not None else (np.zeros((self.nmp, 3)) for _ in range(3))
        self._call("get_state", C.c_int(irep), _d(x), _d(v), _d(a))
        return x, v, a

    def set_step(self, istep):
        self._call("set_step", C.c_int64(istep))

    # ---- lists -------------------------------------------------------------
    def neighbor_list(self, irep=0):
        self._call("neighbor_list", C.c_int(irep))

    def full_lists(self):
        self._call("full_lists")

    def get_list_sizes(self, irep=0):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self._call("get_list_sizes", C.c_int(irep), C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def get_lists(self, irep=0):
        nw, ne, nb = self.get_list_sizes(irep)
        wca = np.zeros((nw, 2), dtype=np.int32)
        ele = np.zeros((ne, 2), dtype=np.int32)
        bp = np.zeros((nb, 7), dtype=np.int32)
        coef = np.zeros((nb, 2))
        self._call("get_lists", C.c_int(irep), _i(wca), _i(ele), _i(bp), _d(coef))
        return wca, ele, bp, coef

    def get_local_lists(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self._call("get_local_lists", C.byref(a), C.byref(b), C.byref(c), None, None, None)
        bond = np.zeros((a.value, 2), dtype=np.int32)
        angl = np.zeros((b.value, 3), dtype=np.int32)
        dih = np.zeros((c.value, 4), dtype=np.int32)
        self._call("get_local_lists", C.byref(a), C.byref(b), C.byref(c), _i(bond), _i(angl), _i(dih))
        return bond, angl, dih

    def bp3_table(self):
        n = C.c_int()
        hsh = np.zeros(4096, dtype=np.int32)
        dH, dS = np.zeros(468), np.zeros(468)
        self._call("bp3_table", C.byref(n), _i(hsh), _d(dH), _d(dS))
        return n.value, hsh, dH, dS

    # ---- hot path ------------------------------------------------------------
    def set_bp_mc(self, flg):
        self._call("set_bp_mc", C.c_int(flg))

    def force(self, irep=0):
        f = np.zeros((self.nmp, 3))
        self._call("force", C.c_int(irep), _d(f))
        return f

    def force_terms(self, irep, mask):
        f = np.zeros((self.nmp, 3))
        self._call("force_terms", C.c_int(irep), C.c_uint(mask), _d(f))
        return f

    def energy_sumup(self, irep, tempK, recompute_bp=True):
        e = np.zeros(ENE_N)
        self._call("energy_sumup", C.c_int(irep), C.c_double(tempK), C.c_int(int(recompute_bp)), _d(e))
        return e

    def energy_kinetic(self, irep=0):
        e = C.c_double()
        self._call("energy_kinetic", C.c_int(irep), C.byref(e))
        return e.value

    def get_bp_state(self, irep=0):
        nb = self.get_list_sizes(irep)[2]
        st = np.zeros(nb, dtype=np.int32)
        en = np.zeros(nb)
        self._call("get_bp_state", C.c_int(irep), _i(st), _d(en))
        return st, en

    def get_rg(self, irep=0):
        e = C.c_double()
        self._call("get_rg", C.c_int(irep), C.byref(e))
        return e.value

    def init_velocities(self, irep=0):
        self._call("init_velocities", C.c_int(irep))

    def bp_status_to_mc(self):
        """bp_status_MC = bp_status (src/job_md.F90:344-348)"""
        self._call("bp_status_to_mc")

    def bpcoef(self):
        """write_bpcoef's table: (imp, jmp, dH, dS * 1e-3) of every sequence-allowed pair, 1-based"""
        n = C.c_int()
        self._call("bpcoef", C.c_int(0), C.byref(n), None, None, None, None)
        i, j = np.zeros(n.value, dtype=np.int32), np.zeros(n.value, dtype=np.int32)
        dH, dS = np.zeros(n.value), np.zeros(n.value)
        self._call("bpcoef", C.c_int(n.value), C.byref(n), _i(i), _i(j), _d(dH), _d(dS))
        return i, j, dH, dS

    def check_force(self, tempK):
        fe, ff = np.zeros((self.nmp, 3)), np.zeros((self.nmp, 3))
        md = C.c_double()
        self._call("check_force", C.c_double(tempK), _d(fe), _d(ff), C.byref(md))
        return fe, ff, md.value

    # ---- REMD -----------------------------------------------------------------
    def _remd_args(self, nrep, values):
        nr = np.zeros(4, dtype=np.int32)
        vals = np.zeros((4, 256))
        for k, n in enumerate(nrep):
            nr[k] = n
            if n > 1:
                vals[k, :n] = values[k]
        self.nrep_all = int(np.prod([n for n in nr if n > 1])) if any(n > 1 for n in nr) else 1
        self.ndim_replica = int(sum(1 for n in nr if n > 1))
        return nr, vals

    def remd_pair_table(self):
        tb = np.zeros((2 * self.ndim_replica, self.nrep_all), dtype=np.int32)
        self._call("remd_pair_table", _i(tb))
        return tb

    def remd_state(self):
        r2l = np.zeros(self.nrep_all, dtype=np.int32)
        l2r = np.zeros(self.nrep_all, dtype=np.int32)
        l2v = np.zeros((4, self.nrep_all))
        self._call("remd_state", _i(r2l), _i(l2r), _d(l2v))
        return r2l, l2r, l2v

    def energy_replica(self):
        e = np.zeros((self.nrep_proc, ENE_N))
        re = np.zeros((self.nrep_all, 2))
        self._call("energy_replica", _d(e), _d(re))
        return e, re

    def replica_exchange(self, replica_energies, uniforms=None):
        na, nc = C.c_int(), C.c_int()
        self._call("replica_exchange", _d(replica_energies), _d(uniforms), C.byref(na), C.byref(nc))
        return na.value, nc.value


def load_library(path: Path) -> C.CDLL:
    if not path.exists():
        raise SisError(f"{path} not found -- build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
                       "there is no CPU fallback")
    return C.CDLL(str(path), mode=C.RTLD_GLOBAL)
