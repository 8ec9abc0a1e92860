"""pytest configuration: registers the `gpu` marker and shares the CPU checkers."""
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle_bindings import Oracle

    return Oracle()


@pytest.fixture(scope="session")
def ref():
    from oracle_bindings import RefOracle, ref_available

    if not ref_available():
        pytest.skip("oracle/_ref/libref_oracle.so not built (needs /root/reference; see oracle/build_ref.sh)")
    return RefOracle()
