"""The bounds-asserting build (libalacgpu_checked.so, -DALACGPU_CHECKED) over the cases that stress the kernels'
indexing: random payload bits, truncated / garbage frames, maximum and tiny frame sizes, degenerate signals, cookie
variants -- through the stream-lane kernels and (ALACGPU_FLAG_FORCE_FRAME_LANES) the frame-lane kernels.

compute-sanitizer is closed on this GPU pool (it answers 86 "closed"), so this build stands in for memcheck: every
index a kernel derives from stream contents -- bitstream ring fills, plane rows, work-list entries, PCM positions,
shared-memory ring levels -- is compared with the extent of its buffer before use; a violation fails the call
with ALACGPU_ERR_STATE (and the test), instead of reading or writing out of bounds."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHECKED = os.path.join(ROOT, "alac", "net_b200", "libalacgpu_checked.so")
CASES = "fuzz or malformed or extreme or degenerate or cookie or mismatch or ragged or added_after"


@pytest.mark.parametrize("suite", ["tests/test_gpu_parity.py", "tests/test_gpu_frame_lanes.py"])
def test_checked_build_runs_the_stress_cases_clean(suite):
    assert os.path.exists(CHECKED), "libalacgpu_checked.so is missing: run __graft_entry__.build()"
    env = dict(os.environ, ALACGPU_LIB=CHECKED)
    r = subprocess.run([sys.executable, "-m", "pytest", suite, "-x", "-q", "-m", "gpu", "-k", CASES, "-p", "no:cacheprovider"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout


def test_checked_build_is_the_one_that_was_loaded():
    """the subprocess really runs the checked library: a deliberately undersized extent trips an assertion"""
    code = (
        "import os, sys; sys.path.insert(0, %r)\n"
        "from alac.net_b200 import _native as N\n"
        "assert N.LIB_PATH.endswith('libalacgpu_checked.so'), N.LIB_PATH\n"
        "L = N.load(); import ctypes as C\n"
        "assert L.alacgpu_abi_version() == 2\n"
        "print('checked lib ok')\n" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, ALACGPU_LIB=CHECKED), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "checked lib ok" in r.stdout, r.stdout + r.stderr
