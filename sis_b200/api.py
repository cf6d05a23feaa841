"""Host-side mirror of the reference's hot-path interface on top of `libsisgpu.so`.

`SisGpu` exposes the same calls the Fortran driver makes (`neighbor_list`, `force`,
`energy_sumup`, `energy_replica`, `replica_exchange`, the `job_md` loop body as `md_steps`,
`check_force`), with the reference's argument meaning, through the C ABI of
`include/sisgpu.h`.  There is no CPU fallback: without the built library or without a CUDA
device construction fails.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

from ._capi import CApi, SisError, _d, _i, load_library

# SISGPU_LIB: another build of the same library (A/B runs of kernel variants on the GPU box)
LIB_PATH = Path(os.environ.get("SISGPU_LIB") or Path(__file__).resolve().parent / "libsisgpu.so")
_lib = None


def library() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = load_library(LIB_PATH)
    return _lib


class SisGpu(CApi):
    # options applied to every new handle before the caller's own (tests: run a suite on another kernel variant)
    default_options: dict = {}

    def __init__(self, device: int = 0, **options):
        lib = library()
        h = C.c_void_p()
        lib.sisgpu_last_error.restype = C.c_char_p
        if lib.sisgpu_create(C.byref(h), C.c_int(device)) != 0:
            raise SisError("sisgpu_create: " + lib.sisgpu_last_error().decode())
        super().__init__(lib, "sisgpu_", h)
        for k, v in {**SisGpu.default_options, **options}.items():
            self.set_option(k, v)

    def close(self):
        if getattr(self, "h", None):
            self.lib.sisgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, name: str, value: float):
        self._call("set_option", C.c_char_p(name.encode()), C.c_double(value))

    # ---- job_md loop body ----------------------------------------------------------
    def md_steps(self, nsteps: int) -> int:
        nreb = C.c_int64()
        self._call("md_steps", C.c_int64(nsteps), C.byref(nreb))
        return nreb.value

    def get_step(self) -> int:
        s = C.c_int64()
        self._call("get_step", C.byref(s))
        return s.value

    def launch_count(self) -> int:
        n = C.c_int64()
        self._call("launch_count", C.byref(n))
        return n.value

    def get_timers(self, reset=False) -> dict:
        out = np.zeros(16)
        self._call("get_timers", _d(out), C.c_int(int(reset)))
        d = dict(zip(["step_ms", "nl_ms", "energy_ms", "exchange_ms"], out[:4].tolist()))
        if out[8] > 0:
            d["persistent_clk_per_step"] = dict(zip(["force", "barrier1", "integrate", "barrier2"],
                                                    (out[4:8] / out[8]).round(1).tolist()))
            d["persistent_clk_per_step"].update(zip(["bp_gate", "bp_eval", "bp_arrive", "loc_tables", "loc_gauss"],
                                                    (out[9:14] / out[8]).round(1).tolist()))
            if out[15] > 0:
                d["persistent_launch"] = {"launches": int(out[15]), "clk_per_launch": round(out[14] / out[15], 1),
                                          "steps_per_launch": round(out[8] / out[15], 2)}
        return d

    def dcd_energies(self, frames32: np.ndarray, tempK: float) -> np.ndarray:
        """job_dcd: the 11 energies of every frame of a float32 trajectory [nframes][nmp][3]."""
        fr = np.ascontiguousarray(frames32, dtype=np.float32)
        out = np.zeros((fr.shape[0], 11))
        self._call("dcd_energies", C.c_int(fr.shape[0]), fr.ctypes.data_as(C.POINTER(C.c_float)), C.c_double(tempK), _d(out))
        return out

    def snapshot_begin(self, irep: int, tempK: float, flags: int = 1 | 2 | 8) -> int:
        slot = C.c_int()
        self._call("snapshot_begin", C.c_int(irep), C.c_double(tempK), C.c_int(flags), C.byref(slot))
        return slot.value

    def snapshot_end(self, slot: int, nmp: int, cap: int = 4096):
        """-> (xyz32 [nmp][3], (imp, jmp, ene32) records, energies[11], Ekin)"""
        xyz = np.zeros((nmp, 3), dtype=np.float32)
        n = C.c_int()
        bi, bj = np.zeros(cap, dtype=np.int32), np.zeros(cap, dtype=np.int32)
        be = np.zeros(cap, dtype=np.float32)
        ene, ek = np.zeros(11), C.c_double()
        self._call("snapshot_end", C.c_int(slot), xyz.ctypes.data_as(C.POINTER(C.c_float)), C.byref(n), _i(bi), _i(bj),
                   be.ctypes.data_as(C.POINTER(C.c_float)), C.c_int(cap), _d(ene), C.byref(ek))
        return xyz, (bi[:n.value], bj[:n.value], be[:n.value]), ene, ek.value

    def last_step_kernel(self) -> str:
        self.lib.sisgpu_last_step_kernel.restype = C.c_char_p
        return self.lib.sisgpu_last_step_kernel(self.h).decode()

    def mark(self, which: int) -> None:
        self._call("mark", C.c_int(which))

    def elapsed_ms(self) -> float:
        ms = C.c_double()
        self._call("elapsed_ms", C.byref(ms))
        return ms.value

    def cta_profile(self, ncta: int) -> np.ndarray:
        out = np.zeros((ncta, 8))
        self._call("get_cta_profile", C.c_int(ncta), _d(out))
        return out

    # ---- REMD ------------------------------------------------------------------------
    def remd_setup(self, nrep, values, nstep_exchange, rank=0, nranks=1):
        nr, vals = self._remd_args(nrep, values)
        self._call("remd_setup", _i(nr), _d(vals), C.c_int(nstep_exchange), C.c_int(rank), C.c_int(nranks))

    def remd_stats(self):
        a, b = C.c_int64(), C.c_int64()
        self._call("remd_stats", C.byref(a), C.byref(b))
        return a.value, b.value

    def get_rng_state(self) -> np.ndarray:
        """restart payload of the counter-based generators (PRNG / PRNGREP blocks of the .rst)"""
        n = C.c_int()
        self._call("get_rng_state", None, C.c_int(0), C.byref(n))
        st = np.zeros(n.value, dtype=np.int64)
        self._call("get_rng_state", st.ctypes.data_as(C.POINTER(C.c_int64)), C.c_int(n.value), C.byref(n))
        return st

    def set_rng_state(self, st: np.ndarray) -> None:
        st = np.ascontiguousarray(st, dtype=np.int64)
        self._call("set_rng_state", st.ctypes.data_as(C.POINTER(C.c_int64)), C.c_int(len(st)))

    def remd_set_labels(self, rep2lab, idx_pair: int) -> None:
        r = np.ascontiguousarray(rep2lab, dtype=np.int32)
        self._call("remd_set_labels", r.ctypes.data_as(C.POINTER(C.c_int)), C.c_int(idx_pair))

    def remd_idx_pair(self) -> int:
        v = C.c_int()
        self._call("remd_get_idx_pair", C.byref(v))
        return v.value

    def comm_init_torch(self):
        """NCCL communicator for the exchange all-gather; the unique id travels over the
        torch.distributed process group (the Fortran driver would MPI_BCAST it)."""
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        buf = (C.c_ubyte * 128)()
        if rank == 0:
            if self.lib.sisgpu_nccl_unique_id(buf) != 0:
                raise SisError(self._err().decode())
        t = torch.tensor(list(buf), dtype=torch.uint8)
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.broadcast(t, 0)
        raw = bytes(t.cpu().tolist())
        self._call("comm_init", C.c_int(world), C.c_int(rank), C.c_char_p(raw))

    def peer_init_torch(self):
        """exchange buffers of all ranks of the node mapped into each other (CUDA IPC handles all-gathered
        over torch.distributed; the Fortran driver would MPI_ALLGATHER them): the exchange steps of a
        multi-rank REMD then run inside the persistent kernels over NVLink peer memory"""
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        buf = (C.c_ubyte * 64)()
        self._call("peer_export", buf)
        t = torch.tensor(list(buf), dtype=torch.uint8)
        if dist.get_backend() == "nccl":
            t = t.cuda()
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        raw = b"".join(bytes(o.cpu().tolist()) for o in out)
        self._call("peer_import", C.c_int(world), C.c_char_p(raw))
