"""CPU suite: lib/hc_sfo2overlaps (scripts/sfo2overlaps.py of the reference in C++, all host threads) against what the
reference script wrote for the same SFO text (tests/golden/sfo_*.npz, oracle/make_golden_sfo.py): the overlaps file byte for
byte, and the two counts the script prints."""
import os
import subprocess

import numpy as np
import pytest

from haploconduct_b200 import build as B
from util import GOLDEN

NAMES = sorted(f[len("sfo_"):-4] for f in os.listdir(GOLDEN) if f.startswith("sfo_"))
EXE = os.path.join(B.LIBDIR, "hc_sfo2overlaps")


@pytest.mark.parametrize("threads", [1, 5])
@pytest.mark.parametrize("name", NAMES)
def test_converter_writes_the_scripts_bytes(built_lib, tmp_path, name, threads):
    z = np.load(os.path.join(GOLDEN, "sfo_" + name + ".npz"))
    (tmp_path / "in.sfo").write_bytes(z["sfo"].tobytes())
    env = dict(os.environ, OMP_NUM_THREADS=str(threads))
    out = subprocess.run([EXE, "--in", "in.sfo", "--out", "out.txt", "--num_singles", str(int(z["num_singles"])), "--num_pairs", str(int(z["num_pairs"]))],
                         cwd=str(tmp_path), env=env, check=True, stdout=subprocess.PIPE).stdout
    assert (tmp_path / "out.txt").read_bytes() == z["overlaps"].tobytes()
    assert out == z["stdout"].tobytes()


def test_converter_rejects_what_the_script_asserts_on(built_lib, tmp_path):
    (tmp_path / "bad.sfo").write_text("1\t2\tN\t3\t4\t50\t50\n")             # seven fields: assert len(sfo_line) == 8
    r = subprocess.run([EXE, "--in", "bad.sfo", "--out", "o.txt", "--num_singles", "5", "--num_pairs", "0"], cwd=str(tmp_path), stderr=subprocess.PIPE)
    assert r.returncode == 1 and b"AssertionError" in r.stderr
    (tmp_path / "bad2.sfo").write_text("1\t99\tN\t3\t4\t50\t50\t0\n")          # id beyond num_singles + 2 * num_pairs
    r = subprocess.run([EXE, "--in", "bad2.sfo", "--out", "o.txt", "--num_singles", "5", "--num_pairs", "2"], cwd=str(tmp_path), stderr=subprocess.PIPE)
    assert r.returncode == 1 and b"AssertionError" in r.stderr
