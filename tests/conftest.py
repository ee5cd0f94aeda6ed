import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (B200); run with -m gpu")
    # the product library is built in-tree by __graft_entry__.build(); build it here too if a fresh checkout runs the
    # tests first (nvcc cross-compiles without a GPU).  A failed build must fail the tests loudly, not skip them.
    from sassena_b200 import build as _b
    if _b.needs_build():
        _b.build()


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.build()
    return o


@pytest.fixture(scope="session")
def gpu_ctx():
    """One ScatterContext on cuda:0 for the whole session (fails loudly if the .so or the GPU is missing)."""
    import sassena_b200
    ctx = sassena_b200.ScatterContext(0)
    yield ctx
    ctx.close()
