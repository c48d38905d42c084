"""CPU tier: the host-side staging helpers of libalacgpu (copy pool, stager / issuer / drainer hand-over) built
on their own and run under ThreadSanitizer when the toolchain provides it -- the race detection SURVEY.md
section 5 asks for, for the part of the runtime that has threads."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "staging_test.cpp")
INC = os.path.join(ROOT, "alac", "net_b200", "csrc")


def _cuda_include():
    for p in (os.environ.get("CUDA_HOME", ""), "/usr/local/cuda"):
        if p and os.path.exists(os.path.join(p, "include", "cuda_runtime.h")):
            return os.path.join(p, "include")
    return None


@pytest.mark.parametrize("tsan", [False, True])
def test_copy_pool_and_progress(tsan, tmp_path):
    cuda = _cuda_include()
    if cuda is None:
        pytest.skip("no CUDA headers (host_staging.h includes cuda_runtime.h for its ring type)")
    exe = str(tmp_path / ("staging_tsan" if tsan else "staging"))
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-pthread", "-I", INC, "-I", cuda, SRC, "-o", exe]
    if tsan:
        cmd[1:1] = ["-fsanitize=thread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and tsan:
        pytest.skip("toolchain without ThreadSanitizer: " + r.stderr[-200:])
    assert r.returncode == 0, r.stderr[-2000:]
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=1")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600, env=env)
    if tsan and "FATAL: ThreadSanitizer" in r.stderr and "unexpected memory mapping" in r.stderr:
        pytest.skip("ThreadSanitizer cannot map its shadow memory in this container")
    assert r.returncode == 0 and "staging ok" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]
