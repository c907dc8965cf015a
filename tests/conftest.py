import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    g = np.load(os.path.join(ROOT, "tests", "golden", "golden.npz"))
    names = sorted({k.rsplit("_", 1)[0] for k in g.files})
    return {n: {f: g[n + "_" + f] for f in ("seq", "args", "cp1", "cp2", "rc", "bed")} for n in names}


@pytest.fixture(scope="session")
def built():
    from ribbit_b200 import build
    build.build_cuda()
    build.build_emulator()
    import oracle_util
    oracle_util.port()
    return True
