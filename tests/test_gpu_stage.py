"""GPU suite: the pinned-ring path for pageable host buffers (csrc/hc_stage.cu).  The library reads HC_STAGE_MIN /
HC_STAGE_CHUNK when it is loaded, so the cases run in a child process: with a 1-byte threshold and 4 KB ring buffers
every host<->device copy of hc_store_create, hc_fno1/3 and hc_dedup_edges goes through many ring rounds, and the
results must still equal the oracle's."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env_extra):
    env = dict(os.environ, **env_extra)
    r = subprocess.run([sys.executable] + args, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


def test_staged_copies_small_ring(built_lib):
    env = {"HC_STAGE_MIN": "1", "HC_STAGE_CHUNK": "4096"}
    out = _run(["-c", "import __graft_entry__ as g; g.smoke()"], env)
    assert "smoke ok" in out
    out = _run(["-m", "pytest", "-q", "-x", "-m", "gpu", "tests/test_gpu_fno.py", "tests/test_gpu_dedup.py",
                "tests/test_gpu_fastq.py"], env)
    assert " passed" in out and "failed" not in out


def test_staged_copies_odd_chunk(built_lib):
    out = _run(["-m", "pytest", "-q", "-x", "-m", "gpu", "tests/test_gpu_fno.py", "-k", "random"],
               {"HC_STAGE_MIN": "1", "HC_STAGE_CHUNK": "1000003"})
    assert " passed" in out and "failed" not in out
