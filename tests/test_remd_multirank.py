"""N > 1 path on CPU: two processes (gloo) run the product's REMD host logic -- block distribution
grep = rank * nrep_proc + irep (src/init_replica.F90:117-119), the all-gather that replaces
MPI_ALLREDUCE (src/job_md.F90:531-533), replica_exchange replayed identically on every rank
(src/replica_exchange.F90:44-154) and the rank-local parameter refresh / velocity scaling -- and are
checked against the single-process oracle.  The device side of the same entry points is covered by
tests/test_gpu_terms.py; here the energies are synthetic, the logic is the product's."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.oracle import Oracle
from sis_b200 import fixtures as fx
from tests.hostcheck import HostCheck

NREP, NRANK, NROUND = 8, 2, 6
TEMPS = [293.15 + 10.0 * k for k in range(NREP)]
IONS = [0.05, 0.1, 0.2, 0.4]


def synthetic_energies(round_, grep, nrep):
    """(own, cross) potential energy of global replica `grep` (1-based) at exchange round `round_`."""
    rng = np.random.default_rng(1000 * round_ + grep)
    e = -40.0 + 3.0 * rng.normal()
    return e, e + 0.8 * rng.normal()


def setup(kind, be, rank=None, nranks=1):
    if kind == "temp":
        nrep, values, n = [NREP, 1, 1, 1], [TEMPS, [], [], []], NREP
    else:   # 2-D: temperature x ionic strength
        nrep, values, n = [2, 1, 4, 1], [TEMPS[:2], [], IONS, []], 8
    nloc = n // nranks
    spec = fx.t2s2hp("T2S2HP_folded", nstep_MC=0).replicated(nloc)
    spec.configure(be)
    if rank is None:
        be.remd_setup(nrep, values, 10)
    else:
        be.remd_setup(nrep, values, 10, rank, nranks)
    return n, nloc


def worker(rank, port, kind, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=NRANK)
    try:
        hc = HostCheck()
        n, nloc = setup(kind, hc, rank, NRANK)
        ref = Oracle() if rank == 0 else None
        if ref is not None:
            setup(kind, ref)
            assert np.array_equal(hc.remd_pair_table(), ref.remd_pair_table())
        counter = 0
        for it in range(NROUND):
            # local (own, cross) energies -> all-gather of 2 * nrep_proc doubles per rank
            send = torch.tensor([synthetic_energies(it, rank * nloc + r + 1, n) for r in range(nloc)], dtype=torch.float64)
            recv = [torch.zeros_like(send) for _ in range(NRANK)]
            dist.all_gather(recv, send)
            E = torch.cat(recv).numpy()
            u = np.array([hc.uniform(123, 0x7fffffff, counter + i) for i in range((n + 1) // 2)])
            counter += (n + 1) // 2
            na, nc, scale, par = hc.remd_exchange(E, u)
            r2l, l2r, _ = hc.remd_state()
            # every rank holds the same tables
            tabs = [torch.zeros(2 * n, dtype=torch.int32) for _ in range(NRANK)]
            dist.all_gather(tabs, torch.tensor(np.concatenate([r2l, l2r]), dtype=torch.int32))
            assert all(torch.equal(tabs[0], t) for t in tabs)
            # rank-local refresh, gathered for the comparison with the oracle
            allpar = [torch.zeros(nloc, 5, dtype=torch.float64) for _ in range(NRANK)]
            dist.all_gather(allpar, torch.tensor(np.column_stack([par, scale])))
            if ref is not None:
                v0 = [ref.get_state(r)[1].copy() for r in range(n)]
                nao, nco = ref.replica_exchange(E, u)
                assert (na, nc) == (nao, nco)
                r2l_o, l2r_o, l2v_o = ref.remd_state()
                assert np.array_equal(r2l, r2l_o) and np.array_equal(l2r, l2r_o)
                P = torch.cat(allpar).numpy()
                for g in range(n):
                    lab = r2l_o[g] - 1
                    if kind == "temp":
                        assert P[g, 0] == l2v_o[0, lab]
                    else:
                        assert P[g, 0] == l2v_o[0, lab] and P[g, 1] == l2v_o[2, lab]
                out.put((it, na, nc))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["temp", "temp_ion"])
def test_remd_two_ranks_match_single_process_oracle(kind):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, port, kind, out)) for r in range(NRANK)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    rounds = [out.get(timeout=5) for _ in range(NROUND)]
    assert sum(r[1] for r in rounds) > 0 and sum(r[2] for r in rounds) > 0   # attempts and acceptances happened


def test_velocity_scale_factors():
    """scale_velo (src/replica_exchange.F90:175-213): an accepted temperature swap scales the two
    replicas by sqrt(T_n / T_m) and its inverse; single rank, forced acceptance."""
    hc = HostCheck()
    n, _ = setup("temp", hc, 0, 1)
    E = np.zeros((n, 2))
    E[:, 0] = 0.0
    E[:, 1] = -100.0          # cross energy far lower: delta < 0 for every pair
    na, nc, scale, par = hc.remd_exchange(E, np.full((n + 1) // 2, 0.999))
    assert na == nc == n // 2
    for k in range(0, n, 2):
        sf = np.sqrt(TEMPS[k + 1] / TEMPS[k])
        assert scale[k] == sf and scale[k + 1] == 1.0 / sf
        assert par[k, 0] == TEMPS[k + 1] and par[k + 1, 0] == TEMPS[k]
