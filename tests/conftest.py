import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    """The compiled unmodified reference; skips where neither /root/reference nor a prebuilt
    oracle/_ref/libref_shim.so exists."""
    from oracle.pyoracle import Ref
    if not Ref.available():
        pytest.skip("oracle/_ref not built and /root/reference absent")
    return Ref()


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "hotpath_golden.npz")
    return np.load(path)


@pytest.fixture(scope="session")
def golden_rank3():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "rank3_golden.npz"))


@pytest.fixture(scope="session")
def golden_score_data():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "score_data_golden.npz"))
