"""Parity run of the EXPERIMENTAL CTA-pair conv kernel (`conv_tc_pair_kernel`, `tcgen05.mma.cta_group::2`, DDN_TC_2CTA=1|2).

The kernel was written after round 1's GPU budget was spent and has never run on hardware, so this test is opt-in
(DDN_TEST_2CTA=1) and runs the existing operator and network parity suites in a SUBPROCESS with the switch set (the library
reads it once per process) under a hard timeout, so that a hang cannot take the box with it:

    DDN_TEST_2CTA=1 python -m pytest tests/test_gpu_pair_kernel.py -m gpu -x -q -s
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(os.environ.get("DDN_TEST_2CTA") != "1", reason="experimental kernel; set DDN_TEST_2CTA=1")
@pytest.mark.parametrize("mode", ["1", "2"])
def test_pair_kernel_passes_the_parity_suites(mode):
    env = dict(os.environ, DDN_TC_2CTA=mode)
    env.pop("DDN_TEST_2CTA", None)
    cmd = ["timeout", "240", sys.executable, "-m", "pytest", "tests/test_gpu_ops.py", "tests/test_gpu_network.py", "-m", "gpu", "-x", "-q"]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True)
    sys.stdout.write(r.stdout[-3000:])
    assert r.returncode == 0, "DDN_TC_2CTA=%s: rc=%d (124 = hang)\n%s" % (mode, r.returncode, r.stderr[-2000:])
