"""CPU: the halo-tile plan of csrc/conv_api.cu (experiment switch VG_HALO=1) regroups the tap
lists of down / up convolutions so that a host evaluation of the halo formulation (one (16 + hy) x (8 + hx) tile per
tap group, windows at shift + (m / 8) * halo_w + m % 8) equals the tap-by-tap formulation exactly.  The native test
includes the launcher source itself; no CUDA call is made."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None or shutil.which("make") is None, reason="needs nvcc + make")
def test_halo_plan_matches_per_tap_formulation():
    build = subprocess.run(["make", "-C", ROOT, "halo_plan_test"], capture_output=True, text=True, timeout=1500)
    assert build.returncode == 0, build.stderr[-3000:]
    run = subprocess.run([os.path.join(ROOT, "build", "halo_plan_test")], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0 and "HALO PLAN OK" in run.stdout, run.stdout[-3000:] + run.stderr[-1000:]
    assert run.stdout.count("0 mismatches") == 7
