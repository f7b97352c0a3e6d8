"""CPU suite: the plain-C oracle against the reference's outputs.

* golden fixtures (tests/golden/*.npz, produced by oracle/make_golden.py from the UNMODIFIED
  reference C++): bit-exact, everywhere;
* the compiled reference itself (oracle/_ref/ref_driver), when it is present: fresh seeded inputs.
"""
import os
import tempfile

import numpy as np
import pytest

from haploconduct_b200 import formats as F, workloads as W
from oracle import oracle as O
from util import golden_names, load_golden


@pytest.mark.parametrize("name", golden_names())
def test_oracle_bit_exact_on_golden(name):
    g = load_golden(name)
    cands = g.scored()
    assert len(cands) == len(g.ref_cands)
    res, _ = O.score_batch(g.rs, g.params(), cands)
    ref = g.ref_cands
    for f in ("score", "mismatch_rate", "pos3", "pos4", "cls"):
        assert np.array_equal(res[f], ref[f]), f
    # the graph the reference built from exactly these edges: every adjacency entry is an accepted candidate
    assert g.ref_counts[0] == len(g.ref_graph)


def test_golden_covers_every_case():
    g = load_golden("synth_all_types")
    c = g.scored()
    seen = set()
    for x in c:
        seen.add((chr(x["type1"]), chr(x["type2"]), int(x["ori1"]), int(x["ori2"]), chr(x["ord"])))
    for t1, t2, ords in (("s", "s", "-"), ("s", "p", "-"), ("p", "s", "-"), ("p", "p", "12")):
        for o1 in (0, 1):
            for o2 in (0, 1):
                for od in ords:
                    assert (t1, t2, o1, o2, od) in seen, (t1, t2, o1, o2, od)
    res = g.ref_cands
    assert set(np.unique(res["cls"])) == {0, 1, 2}


def test_prefilter_matches_c_restatement():
    g = load_golden("synth_mismatch_void")
    L = O.lib()
    pre = g.prefilter()
    for relax in (0, 1):
        for mol, mop in ((120, 0), (60, 50), (200, 90), (0, 0)):
            v = F.prefilter(g.cands, mol, mop, bool(relax))
            c = np.ascontiguousarray(g.cands)
            w = np.array([L.hco_prefilter(c[i:i + 1].ctypes.data, mol, mop, relax) for i in range(len(c))], dtype=np.int8)
            assert np.array_equal(v, w), (relax, mol, mop)
    assert (pre == 1).sum() == len(g.ref_cands)


def test_phred_and_overlap_score_primitive():
    L = O.lib()
    for q in range(0, 94):
        assert L.hco_phred_to_prob(q) == 10 ** (-q / 10.0) or abs(L.hco_phred_to_prob(q) - 10 ** (-q / 10.0)) < 1e-18
    p = F.make_params()
    s, mm = O.overlap_score("ACGTACGTAC", "ACGTAC", "IIIIIIIIII", "IIIIII", 4, p)
    assert mm == 0.0 and 0.999 < s < 1.0
    s, mm = O.overlap_score("ACGTACGTAC", "ACGTAC", "IIIIIIIIII", "IIIIII", 10, p)   # pos >= len: early out
    assert s == 0.0 and mm == 1.0
    s, mm = O.overlap_score("NNNN", "NNNN", "!!!!", "!!!!", 0, p)                     # nothing comparable
    assert s == 0.0 and mm == 1.0


@pytest.mark.skipif(not (O.have_ref() and os.path.isdir(O.REFERENCE_ROOT)), reason="compiled reference not available")
@pytest.mark.parametrize("seed", [1, 2])
def test_oracle_vs_compiled_reference_fresh_inputs(seed):
    ss = W.synth_readset(120, 120, seed=100 + seed, n_rate=0.003)
    c = W.geometry_candidates(ss, 2500, seed=200 + seed)
    d = tempfile.mkdtemp(prefix="hc_t_")
    F.write_fastq_set(ss.rs, d + "/s.fastq", d + "/p1.fastq", d + "/p2.fastq")
    F.write_overlaps(d + "/ov.txt", c, ss.rs.ids)
    out = O.run_ref(d, d + "/ov.txt", d + "/s.fastq", d + "/p1.fastq", d + "/p2.fastq", dump_cands=True, edge_threshold=0.96,
                    min_overlap_len=0, merge_contigs=0.005)
    p = F.make_params(edge_threshold=0.96, merge_contigs=0.005)
    res, _ = O.score_batch(ss.rs, p, c)
    ref = out["cands"]
    assert len(ref) == len(c)
    for f in ("score", "mismatch_rate", "pos3", "pos4", "cls"):
        assert np.array_equal(res[f], ref[f]), f


def test_overlaps_file_roundtrip(tmp_path):
    g = load_golden("synth_all_types")
    path = str(tmp_path / "ov.txt")
    F.write_overlaps(path, g.cands, g.rs.ids)
    c2, lines = F.parse_overlaps(path, g.rs.id_to_index())
    keep = g.cands["idx1"] != g.cands["idx2"]
    assert np.array_equal(c2, g.cands[keep])
    assert len(lines) == len(c2)
