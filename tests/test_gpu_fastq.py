"""GPU suite: hc_store_create_fastq (FASTQ text -> device store: line index, record scan, validation, packing on the
device) against the reference's results on the same reads (tests/golden) and against the array-built store."""
import numpy as np
import pytest

from haploconduct_b200 import capi, formats as F
from util import assert_results_match, golden_names, load_golden

pytestmark = pytest.mark.gpu


def _fastq_text(rs, lower_singles=False, descr=False):
    s, p1, p2 = [], [], []
    for i in range(rs.n_reads):
        head = "@%d%s" % (int(rs.ids[i]), " length=%d extra" % i if descr else "")
        if not rs.is_paired(i):
            seq = rs.seq(i).lower() if lower_singles else rs.seq(i)
            s.append("%s\n%s\n+\n%s\n" % (head, seq, rs.qual(i)))
        else:
            p1.append("%s\n%s\n+\n%s\n" % (head, rs.seq(i, 0), rs.qual(i, 0)))
            p2.append("%s\n%s\n+anything\n%s\n" % (head, rs.seq(i, 1), rs.qual(i, 1)))
    return "".join(s).encode(), "".join(p1).encode(), "".join(p2).encode()


@pytest.mark.parametrize("name", golden_names())
def test_store_from_fastq_text_reproduces_reference(built_lib, name):
    g = load_golden(name)
    s, p1, p2 = _fastq_text(g.rs, lower_singles=True, descr=True)        # singles are upper-cased (:123), descriptions ignored
    cands = g.scored()
    with capi.Store.from_fastq(s, p1, p2) as st:
        ids, lens = st.read_ids()
        assert np.array_equal(ids, g.rs.ids)
        assert np.array_equal(lens, g.rs.descs["seq_len"])
        assert int(capi.lib().hc_store_n_single(st.handle)) == g.rs.n_single
        edges, nonedge, per, stats = st.score_batch(g.params(), cands)
    ref = g.ref_cands
    assert_results_match(per, ref["score"], ref["mismatch_rate"], ref["pos3"], ref["pos4"], ref["cls"], what=name)
    with capi.Store(g.rs) as st2:
        e2, n2, per2, _ = st2.score_batch(g.params(), cands)
    assert per.tobytes() == per2.tobytes() and edges.tobytes() == e2.tobytes() and np.array_equal(nonedge, n2)


def test_fastq_edge_cases(built_lib):
    g = load_golden("synth_all_types")
    s, p1, p2 = _fastq_text(g.rs)
    # no trailing newline, an incomplete fifth record line, max_reads
    with capi.Store.from_fastq(s[:-1], p1, p2 + b"@999\nACGT\n") as st:
        ids, lens = st.read_ids()
        assert np.array_equal(ids, g.rs.ids) and np.array_equal(lens, g.rs.descs["seq_len"])
    with capi.Store.from_fastq(s, p1, p2, max_reads=7) as st:
        ids, _ = st.read_ids()
        n_s = g.rs.n_single
        assert ids.tolist() == g.rs.ids[:7].tolist() + g.rs.ids[n_s:n_s + 7].tolist()
    with capi.Store.from_fastq(b"@0x1F\nACGTN\n+\nIIII!\n@017 x\nacgt\n+\n!!!!\n") as st:       # ids by strtoul(.., 0)
        ids, lens = st.read_ids()
        assert ids.tolist() == [31, 15] and lens.tolist() == [[5, 0], [4, 0]]
    bad = [
        (b"0\nACGT\n+\nIIII\n", b"", b""),                                   # header without '@'
        (b"@0\n\n+\n\n", b"", b""),                                          # empty sequence
        (b"@0\nACGT\n+\nIII\n", b"", b""),                                   # lengths differ
        (b"@0\nACGU\n+\nIIII\n", b"", b""),                                  # invalid nucleotide
        (b"@0\nACGT\n+\nII I\n", b"", b""),                                  # quality below '!'
        (b"", b"@0\nACGT\n+\nIIII\n", b"@1\nACGT\n+\nIIII\n"),               # mate headers differ
        (b"", b"@0\nACGT\n+\nIIII\n", b"@00\nACGT\n+\nIIII\n"),              # ... as strings, not as numbers
        (b"", b"@0\nacgt\n+\nIIII\n", b"@0\nACGT\n+\nIIII\n"),               # mates are not upper-cased (:196-197)
        (b"", b"", b""),                                                      # nothing
        (b"@0\nACGT\n+\n", b"", b""),                                        # no complete record
    ]
    for s_, a_, b_ in bad:
        with pytest.raises(capi.HcError):
            capi.Store.from_fastq(s_, a_, b_)


def test_store_streamed_from_files(built_lib, tmp_path, monkeypatch):
    """hc_store_create_fastq_files: the files go through the pinned ring piece by piece (here with 64 KB ring buffers, so that
    a few hundred reads are many pieces); the store must be the one built from the same text in memory."""
    monkeypatch.setenv("HC_STAGE_CHUNK", "65536")
    g = load_golden("c1_savage_example_full") if "c1_savage_example_full" in golden_names() else load_golden(golden_names()[0])
    s, p1, p2 = _fastq_text(g.rs)
    paths = []
    for name, text in (("s.fastq", s), ("p1.fastq", p1), ("p2.fastq", p2)):
        f = tmp_path / name
        f.write_bytes(text)
        paths.append(str(f) if len(text) else None)
    cands = g.scored()[:20000]
    with capi.Store.from_fastq(s, p1, p2) as a, capi.Store.from_fastq_files(*paths) as b:
        L = capi.lib()
        assert L.hc_store_n_reads(a.handle) == L.hc_store_n_reads(b.handle) and L.hc_store_n_single(a.handle) == L.hc_store_n_single(b.handle)
        ra, rb = a.score_batch(g.params(), cands), b.score_batch(g.params(), cands)
        assert ra[2].tobytes() == rb[2].tobytes() and ra[0].tobytes() == rb[0].tobytes()
        ia, ib = a.read_ids(), b.read_ids()
        assert all(np.array_equal(x, y) for x, y in zip(ia, ib))
    with pytest.raises(capi.HcError):
        capi.Store.from_fastq_files(str(tmp_path / "missing.fastq"))
