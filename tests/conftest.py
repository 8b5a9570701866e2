import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
for p in (REPO, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Shared libraries must exist (built by __graft_entry__.build()); never built lazily on the
    GPU box, where they arrive prebuilt with the snapshot."""
    from iss_b200 import capi
    for p in (capi.cuda_lib_path(), capi.host_lib_path()):
        if not os.path.exists(p):
            import __graft_entry__
            __graft_entry__.build()
            break
    return capi
