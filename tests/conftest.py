import os
import sys

import pytest

# every 16-bit tile-local neighbour row is verified against its chunk's tile after each rebuild (csrc/tile.cu)
os.environ.setdefault("PARM_B200_TILE_CHECK", "1")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.dirname(os.path.abspath(__file__))
for p_ in (HERE, ROOT):
    if p_ not in sys.path:
        sys.path.insert(0, p_)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_built():
    """Build the CPU oracles once (port always; the reference build only where /root/reference exists)."""
    from oracle import cpu
    cpu.build(("port", "ref"))
    return cpu
