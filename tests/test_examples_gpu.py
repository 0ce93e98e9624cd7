"""examples/verify.py (the runnable counterpart of the reference's examples/verify.py:136-156) as a test: the script's
own verdict on a small pair and on the reference's constants (330 x 422, kernel_size 15, full [H,W,W] volume)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("extra", [["--height", "72", "--width", "150", "--kernel-size", "5"],
                                   ["--height", "64", "--width", "120", "--kernel-size", "7", "--seed", "3"],
                                   []])
def test_verify_script(extra):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "verify.py"), *extra], cwd=ROOT,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0 and "VERIFY_OK" in r.stdout, r.stdout[-3000:]
    assert "cuda forward time" in r.stdout and "Cost Volume shape" in r.stdout
