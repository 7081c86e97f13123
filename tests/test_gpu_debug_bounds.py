"""GPU: the debug build of the library (index checks on the shared-memory tiles, halo rings and global offsets of the
tiled kernels, `make debug`) runs the odd-size whole-path script with zero violations -- for the default kernels and
for the tensor-map (TMA) variants of the fused kernel.  Stands in for compute-sanitizer, which is closed on this pool."""
import os
import re
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DBG = os.path.join(ROOT, "chessboard_vision_b200", "libcvb200_dbg.so")


@pytest.mark.parametrize("fused", ["", "0", "2", "3", "4", "5"])
def test_no_index_violations(fused):
    if not os.path.exists(DBG):
        pytest.skip("libcvb200_dbg.so not built (python -c 'import __graft_entry__ as g; g.build()')")
    env = dict(os.environ, CVB200_LIB=DBG)
    env.pop("CVB_FUSED", None)
    if fused:
        env["CVB_FUSED"] = fused
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sanitize_run.py")], env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    m = re.search(r"bounds violations: (-?\d+) first at line (\d+)", r.stdout)
    assert m, r.stdout[-500:]
    assert int(m.group(1)) == 0, "index check failed %s time(s), first at source line %s" % (m.group(1), m.group(2))
