import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(__file__)))

# Parity flag: zero-knowledge proofs of the drop-in harness draw their masks and salts from the seeded stream the CPU
# reference was run with (libiop_b200/dropin/shims/sodium_shim.c); without it they come from the ChaCha20/getrandom generator.
os.environ.setdefault("B200_DETERMINISTIC_RNG", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle_lib import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(os.path.dirname(__file__), "golden", "hotpath_golden.npz")
    return np.load(path)


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference, only where oracle/_ref was built (never required)."""
    from oracle_lib import Reference
    if not Reference.available():
        pytest.skip("oracle/_ref/libiopref.so not built")
    return Reference()
