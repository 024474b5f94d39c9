import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    from oracle.pyoracle import oracle
    return oracle()


@pytest.fixture(scope="session")
def ref():
    from oracle.pyoracle import Reference, reference
    if not Reference.available() and not Path("/root/reference/src/problem.cpp").exists():
        pytest.skip("oracle/_ref/libpagmo_ref.so not built and /root/reference absent")
    return reference()


@pytest.fixture(scope="session")
def ctx():
    from pagmo2_b200 import capi
    c = capi.Context(int(os.environ.get("LOCAL_RANK", "0")))
    yield c
    c.close()
