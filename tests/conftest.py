import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # Tests that need the reference object code (oracle/_ref) or the drop-in decide at import time whether to skip;
    # build both before collection where that is possible (the f5c tree is present), so that a fresh clone does not
    # silently skip them. Never fatal: on the GPU box the prebuilt files travel with the snapshot.
    try:
        import __graft_entry__ as g
        if os.path.isdir("/root/reference/src"):
            g.build_oracle()
            g.build_cuda()
            g.build_dropin()
    except Exception as e:  # pragma: no cover
        sys.stderr.write(f"conftest: pre-collection build skipped ({e})\n")


@pytest.fixture(scope="session")
def built():
    """Everything compiled once per session (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()
    return True
