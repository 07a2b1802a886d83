"""pytest configuration: the `gpu` marker, import paths, and the two checkers (oracle restatement, reference harness)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _build_oracle():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "libmcsim_oracle.so"], check=True)


@pytest.fixture(scope="session")
def oracle():
    """ctypes binding of oracle/libmcsim_oracle.so (pinned math: include/mcsim_math.h)."""
    _build_oracle()
    from oracle.pyoracle import Oracle
    return Oracle(0)


@pytest.fixture(scope="session")
def ref_available():
    from oracle.pyoracle import REF_BIN, REF_SO
    if not (os.path.exists(REF_BIN) and os.path.exists(REF_SO)):
        build = os.path.join(ROOT, "oracle", "ref_harness", "build_ref.sh")
        if os.path.exists("/root/reference/mc_sim_highway/mc_sim_highway.so"):
            subprocess.run([build], check=True, stdout=subprocess.DEVNULL)
    if not (os.path.exists(REF_BIN) and os.path.exists(REF_SO)):
        pytest.skip("oracle/_ref (the reference's machine code under the stub harness) is not built here")
    return True


@pytest.fixture(scope="session")
def backend_lib():
    """libmcsim_b200.so through the product's ctypes layer; built on demand (nvcc cross-compiles without a GPU)."""
    from interactive_highway_simulation_b200 import backend
    if not os.path.exists(backend.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return backend
