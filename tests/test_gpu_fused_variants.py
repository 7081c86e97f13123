"""GPU: the opt-in variants of the fused tile kernel (CVB_FUSED=0..5: persistent, tensor-map TMA fed, lane-private weight
table, packed accumulation, two groups, producer / consumer warps -- csrc/cvb_fused2.cu) give the oracle's enhanced frame
bit for bit on eight shapes x {board, noise}, as the default kernel does.  The variable is read when the kernel is
launched, so every variant runs in its own process (tools/prof_run.py with CVB_CHECK=1)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("fused", ["0", "1", "2", "3", "4", "5"])
def test_variant_is_bit_identical_to_the_oracle(fused):
    env = dict(os.environ, CVB_FUSED=fused, CVB_CHECK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "prof_run.py"), "4", "1"], env=env, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("parity")]
    assert len(lines) == 16 and all(l.endswith("OK") for l in lines), "\n".join(lines)
    assert "k_fused" in r.stdout
