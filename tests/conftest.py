import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
for p in (REPO, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Make sure the in-tree libraries exist (they are built by __graft_entry__.build())."""
    import __graft_entry__ as ge
    csrc = os.path.join(REPO, "ocean-bgc_b200", "csrc")
    need = [os.path.join(csrc, n) for n in ("libbgc_b200.so", "libbgc_b200_strict.so", "libbgc_synth.so")]
    need.append(os.path.join(REPO, "oracle", "libbgc_oracle.so"))
    if not all(os.path.exists(p) for p in need):
        ge.build()
    return ge.load_package()


@pytest.fixture(scope="session", autouse=True)
def _libraries_present(built):
    """Every test file may run on its own: build the in-tree libraries first when they are missing
    (a no-op on the GPU box, where the prebuilt files travel with the snapshot)."""
    return built
