import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


# pytest-xdist workers: the 3-D oracle cases use the OpenMP build, and W workers x (all cores) spinning OpenMP threads
# oversubscribe the box several times over (measured: load average 30 on 8 cores, the suite three times slower).  Each worker
# gets its share of the cores and passive waiting; set before libgomp is loaded (the oracle is dlopen'ed by the first test).
if os.environ.get("PYTEST_XDIST_WORKER"):
    _nw = max(1, int(os.environ.get("PYTEST_XDIST_WORKER_COUNT", "4") or 4))
    os.environ.setdefault("OMP_NUM_THREADS", str(max(1, (os.cpu_count() or _nw) // _nw)))
    os.environ.setdefault("OMP_WAIT_POLICY", "passive")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running CPU test")


def pytest_collection_modifyitems(config, items):
    """`gpu`-marked tests skip (instead of failing in mfc5_create) on a box without a CUDA device."""
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device (libmfc5 has no CPU fallback)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.hookimpl(tryfirst=True)
def pytest_cmdline_main(config):
    """CPU suite (no CUDA device): spread the oracle-vs-golden cases over a few worker processes (pytest-xdist, if installed and
    not already requested) so that the whole `-m "not gpu"` run stays within a few minutes.  On a GPU box the suite runs
    in-process, one CUDA context."""
    if getattr(config.option, "numprocesses", None) not in (None, 0) or os.environ.get("MFC5_TEST_SERIAL"):
        return None
    if hasattr(config, "workerinput") or os.environ.get("PYTEST_XDIST_WORKER"):   # an xdist worker never spawns workers of its own
        return None
    if not config.pluginmanager.hasplugin("xdist"):
        return None
    try:
        import torch
        if torch.cuda.is_available():
            return None
    except Exception:
        pass
    config.option.numprocesses = min(4, max(1, (os.cpu_count() or 2) // 2))
    return None
