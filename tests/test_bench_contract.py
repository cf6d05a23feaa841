"""bench.py contract on the CPU side: the reference arm (the CPU port of the reference's algorithm;
no Fortran compiler exists in the image, so oracle/_ref cannot be built) prints one JSON line with
the keys the driver reads, and the product arm refuses to run without a GPU instead of falling back."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ref-md-steps", "2"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "particle-steps/sec" and d["unit"] == "particle-steps/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1
    assert "64xCAG47" in d["config"]["workload"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    import os
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "0", "--no-cpu"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0          # no CPU fallback
    assert not any(l.startswith("{") for l in r.stdout.splitlines())


def test_omp_oracle_is_threaded_and_agrees_with_the_serial_build():
    """the CPU baseline of bench.py is the -fopenmp build of the oracle: it must really run on more
    than one thread (the 'OpenMP' baseline was single-threaded) and give the serial
    build's forces / energies up to summation order"""
    import os
    import time
    import numpy as np
    from oracle.oracle import Oracle
    from sis_b200 import fixtures as fx
    lib = (ROOT / "oracle" / "libsis_oracle_omp.so")
    assert lib.exists()
    ldd = subprocess.run(["ldd", str(lib)], capture_output=True, text=True).stdout
    assert "libgomp" in ldd
    spec = fx.cag47(4, pbc=True)
    res = []
    for omp in (False, True):
        o = Oracle(omp=omp)
        spec.configure(o)
        o.neighbor_list(0)
        res.append((o.force(0), o.energy_sumup(0, spec.tempK, True), o))
    assert res[0][2].num_threads() == 1
    f0, e0, _ = res[0]
    f1, e1, o = res[1]
    assert np.abs(f1 - f0).max() <= 1e-12 * np.abs(f0).max()
    assert np.allclose(e1, e0, rtol=1e-12, atol=1e-12)
    if (os.cpu_count() or 1) < 2 or len(os.sched_getaffinity(0)) < 2:
        pytest.skip("one host core")
    assert o.num_threads() > 1
    o.init_velocities(0)
    w0, c0 = time.perf_counter(), time.process_time()
    o.md_run(300)
    w, c = time.perf_counter() - w0, time.process_time() - c0
    assert c / w > 1.5, (c, w)
