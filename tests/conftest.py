import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "support")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU restatement (oracle/liboracle.so) — the checker."""
    import pyoracle
    if not os.path.exists(pyoracle.PORT_LIB):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    return pyoracle.OracleLib(pyoracle.PORT_LIB)


@pytest.fixture(scope="session")
def reflib():
    """The reference's own sources + shims (oracle/_ref/libmygram_ref.so), when it has been built."""
    import pyoracle
    if not os.path.exists(pyoracle.REF_LIB):
        pytest.skip("oracle/_ref/libmygram_ref.so not built (needs /root/reference; `make -C oracle ref`)")
    return pyoracle.OracleLib(pyoracle.REF_LIB)


@pytest.fixture(scope="session")
def mgx():
    """The product: Python mirror over libmgx.so. Fails loudly if the CUDA extension is missing."""
    import mgx_loader
    m = mgx_loader.load()
    m.lib()
    return m
