import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


# Every GPU test runs twice: with k_md_rep (CTA sized to the replica, static pair tables) and with the
# replica-local mode of k_md_fast behind option rep_kernel = 0.  Systems that are not replica-local
# take the same kernels either way.
def pytest_generate_tests(metafunc):
    if metafunc.definition.get_closest_marker("gpu") and "replica_kernel" in metafunc.fixturenames:
        metafunc.parametrize("replica_kernel", ["k_md_rep", "k_md_fast"], indirect=True)


@pytest.fixture(autouse=True)
def replica_kernel(request):
    kind = getattr(request, "param", None)
    if kind is None:
        yield None
        return
    from sis_b200.api import SisGpu
    saved = dict(SisGpu.default_options)
    SisGpu.default_options = {**saved, "rep_kernel": 1 if kind == "k_md_rep" else 0}
    yield kind
    SisGpu.default_options = saved


def is_replica_local(name: str) -> bool:
    """`name` (sisgpu_last_step_kernel) is the replica-local kernel the current test variant selects."""
    from sis_b200.api import SisGpu
    if SisGpu.default_options.get("rep_kernel", 1):
        return name in ("k_md_rep<128>", "k_md_rep<256>", "k_md_rep<512>")
    return name == "k_md_fast<1,true>"
