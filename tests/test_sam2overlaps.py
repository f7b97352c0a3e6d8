"""CPU suite: lib/hc_sam2overlaps (scripts/sam2overlaps.py of the reference in C++, all host threads) against what the
reference script wrote for the same SAM / FASTA text (tests/golden/sam_*.npz, oracle/make_golden_sam.py): the overlaps file
byte for byte and, where the fixture ran the script with --verbose, its stdout."""
import os
import subprocess

import numpy as np
import pytest

from haploconduct_b200 import build as B
from util import GOLDEN

NAMES = sorted(f[len("sam_"):-4] for f in os.listdir(GOLDEN) if f.startswith("sam_"))
EXE = os.path.join(B.LIBDIR, "hc_sam2overlaps")


@pytest.mark.parametrize("threads", [1, 6])
@pytest.mark.parametrize("name", NAMES)
def test_converter_writes_the_scripts_bytes(built_lib, tmp_path, name, threads):
    z = np.load(os.path.join(GOLDEN, "sam_" + name + ".npz"))
    for fn, key in (("ref.fasta", "fasta"), ("s.sam", "sam_s"), ("p.sam", "sam_p")):
        (tmp_path / fn).write_bytes(z[key].tobytes())
    cmd = [EXE, "--ref", "ref.fasta", "--out", "out.txt", "--min_overlap_len", str(int(z["min_overlap_len"]))]
    if int(z["use_s"]):
        cmd += ["--sam_s", "s.sam"]
    if int(z["use_p"]):
        cmd += ["--sam_p", "p.sam"]
    if int(z["verbose"]):
        cmd += ["--verbose"]
    out = subprocess.run(cmd, cwd=str(tmp_path), env=dict(os.environ, OMP_NUM_THREADS=str(threads)), check=True, stdout=subprocess.PIPE).stdout
    assert (tmp_path / "out.txt").read_bytes() == z["overlaps"].tobytes()
    assert out == z["stdout"].tobytes()
