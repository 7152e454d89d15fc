"""Both engines of libcassie2d.so against the same parity tests.

The library holds two complete implementations of the step path (csrc/launch.cuh): the quad engine (four lanes per
env) and the thread engine (one env per thread).  By default each control mode runs the engine that measured faster;
CASSIE_ENGINE=quad|thread forces one for every mode.  The choice is read once per process, so the parity files are
re-run here in a subprocess per engine: every mode of every engine is compared with the oracle on a GPU.
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = ["tests/test_gpu_parity.py", "tests/test_gpu_fp32_bar.py", "tests/test_gpu_edges.py", "tests/test_gpu_rollout.py"]


@pytest.mark.parametrize("engine", ["quad", "thread"])
def test_parity_files_with_engine_forced(engine):
    if os.environ.get("CASSIE_ENGINE"):
        pytest.skip("already inside a forced-engine run")
    env = dict(os.environ, CASSIE_ENGINE=engine)
    r = subprocess.run([sys.executable, "-m", "pytest", "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider"] + FILES,
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    tail = "\n".join((r.stdout + r.stderr).strip().splitlines()[-25:])
    assert r.returncode == 0, "engine %s:\n%s" % (engine, tail)
    print("engine %s: %s" % (engine, r.stdout.strip().splitlines()[-1]))
