import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def engine():
    from chessboard_vision_b200.engine import Engine
    e = Engine(0)
    yield e
    # with the debug build (CVB200_LIB=.../libcvb200_dbg.so) the whole session doubles as an index-check run
    import ctypes
    line = ctypes.c_int(0)
    bad = e.lib.cvb_debug_bounds_violations(e.h, ctypes.byref(line))
    e.close()
    if bad >= 0:
        print("\n[debug build] index checks that failed during the session: %d (first at source line %d)" % (bad, line.value))
    assert bad <= 0, "index check failed %d time(s), first at source line %d" % (bad, line.value)
