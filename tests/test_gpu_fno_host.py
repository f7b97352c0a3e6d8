"""GPU suite: the C++ host binding of the FindNextOverlaps interfaces (haploconduct_b200/host/hcb_fno.h, built as
lib/hc_fno on top of hc_fno1 / hc_fno3) against the overlaps.txt the UNMODIFIED reference wrote.

Input = the state the reference's SRBuilder::findNextOverlaps / findNextOverlaps3 started from (adjacency lists,
branching and inclusion edges, labels, visited, new ids, super-reads; oracle/ref_driver --fno-state, stored in
tests/golden/fnostate_*.npz) + the FASTQ files + nonedge_overlaps.txt; the binary produces the edge stream in the
reference's order, flattens nodes_to_SR, calls the device and writes the file.  Bar: byte-identical files -- including
the contig-shaped case of config 5 (1-10 kb contigs, stage-b flags)."""
import glob
import json
import os
import subprocess

import numpy as np
import pytest

from haploconduct_b200 import build as B, formats as F
from util import GOLDEN

pytestmark = pytest.mark.gpu
EXE = os.path.join(B.LIBDIR, "hc_fno")


def state_names():
    return sorted(os.path.basename(p)[len("fnostate_"):-4] for p in glob.glob(os.path.join(GOLDEN, "fnostate_*.npz")))


@pytest.mark.parametrize("name", state_names())
def test_fno_binding_writes_the_reference_file(built_lib, tmp_path, name):
    z = np.load(os.path.join(GOLDEN, "fnostate_" + name + ".npz"))
    ref = [str(x) for x in np.load(os.path.join(GOLDEN, name + ".npz"))["ref_lines"]]
    rs = F.ReadSet(ids=z["ids"], descs=z["descs"], bases=z["bases"], quals=z["quals"], n_single=int(z["n_single"]))
    d = str(tmp_path)
    F.write_fastq_set(rs, d + "/s.fastq", d + "/p1.fastq", d + "/p2.fastq")
    with open(d + "/state.txt", "wb") as f:
        f.write(z["state"].tobytes())
    with open(d + "/nonedge_overlaps.txt", "wb") as f:
        f.write(z["nonedge"].tobytes())
    cmd = [EXE, "--state", d + "/state.txt", "--output", d + "/", "--FNO", "3" if name.startswith("fno3") else "1"]
    if rs.n_single:
        cmd += ["--singles", d + "/s.fastq"]
    if rs.n_reads > rs.n_single:
        cmd += ["--paired1", d + "/p1.fastq", "--paired2", d + "/p2.fastq"]
    out = subprocess.run(cmd, cwd=d, check=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True).stdout
    summary = json.loads([l for l in out.split("\n") if l.startswith("{")][-1])
    with open(d + "/overlaps.txt") as f:
        got = f.read().split("\n")[:-1]
    assert got == ref, name
    assert summary["lines"] == len(ref)
