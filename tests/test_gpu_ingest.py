"""GPU suite: hc_ingest_overlaps (overlaps-file text -> candidates on the device) against the reference's own output
for irregularly spelled files (tests/golden/ingest_*) and against the pinned restatement on fuzzed text."""
import numpy as np
import pytest

from haploconduct_b200 import capi, formats as F, workloads as W
from oracle import ingest_oracle as IO
from util import IngestGolden, ingest_golden_names, load_golden

pytestmark = pytest.mark.gpu


def _check_against_oracle(text, ids, **kw):
    idmap = {}
    for i, v in enumerate(ids):
        idmap.setdefault(int(v), i)          # std::map::insert keeps the first (src/FastqStorage.h:90-93)
    status, scored, filtered = IO.ingest(text, idmap, **kw)
    m = capi.IdMap(ids)
    cand, cl, filt, fl, st = m.ingest(text, F.make_ingest_params(**kw))
    assert F.candidate_lines(cand, ids) == [IO.rec_line(r[1]) for r in scored]
    assert cand["idx1"].tolist() == [r[2] for r in scored] and cand["idx2"].tolist() == [r[3] for r in scored]
    assert cl.tolist() == [r[0] for r in scored]
    assert F.overlap_rec_lines(filt) == [IO.rec_line(r[1]) for r in filtered]
    assert fl.tolist() == [r[0] for r in filtered]
    assert [int(st["n_lines"]), int(st["n_skipped"]), int(st["n_dropped"])] == [len(status), status.count(IO.SKIPPED), status.count(IO.DROPPED)]
    errs = [i for i, s in enumerate(status) if s in (IO.ERROR, IO.UNKNOWN_ID)]
    if errs:
        assert int(st["first_error_line"]) == errs[0] and int(st["first_error_status"]) == status[errs[0]]
        lines = IO.split_lines(text)
        off, ln = int(st["first_error_offset"]), int(st["first_error_length"])
        assert text[off:off + ln] == lines[errs[0]]
    else:
        assert int(st["first_error_line"]) == 2 ** 64 - 1
    return st


@pytest.mark.parametrize("name", ingest_golden_names())
def test_ingest_reproduces_reference(built_lib, name):
    g = IngestGolden(name)
    m = capi.IdMap(g.ids)
    cand, cl, filt, fl, st = m.ingest(g.text, F.make_ingest_params(**g.kw()))
    assert F.candidate_lines(cand, g.ids) == g.ref_scored          # what process_overlaps received, in order
    assert F.overlap_rec_lines(filt) == g.ref_filtered             # what :654-660 appended, in order
    assert [int(st["n_lines"]), int(st["n_scored"]), int(st["n_filtered"]), int(st["n_dropped"]), int(st["n_skipped"])] == \
        [int(g.ref_counts[0]), int(g.ref_counts[1]), int(g.ref_counts[2]), int(g.ref_counts[3] + g.ref_counts[4]), int(g.ref_counts[5])]
    assert int(st["first_error_line"]) == 2 ** 64 - 1
    _check_against_oracle(g.text, g.ids, **g.kw())


@pytest.mark.parametrize("seed,allow_spaces", [(11, False), (12, True), (13, False)])
def test_ingest_fuzz_against_restatement(built_lib, seed, allow_spaces):
    g = load_golden("synth_all_types" if seed != 13 else "synth_mismatch_void")
    ids = g.rs.ids * np.uint64(2 ** 40 + 12345) + np.uint64(7) if seed == 13 else g.rs.ids       # sparse ids: hash table, not the direct table
    text = W.fuzz_overlap_text(g.cands, ids, seed=seed, allow_spaces=allow_spaces)
    _check_against_oracle(text, ids, min_overlap_len=80, min_overlap_perc=20, relax_PE_edges=seed == 13, allow_spaces=allow_spaces)
    _check_against_oracle(text, ids, min_overlap_len=80, max_overlaps=1000, allow_spaces=allow_spaces)
    _check_against_oracle(text, ids[: len(ids) // 2], min_overlap_len=80, allow_spaces=allow_spaces)     # unknown ids


def test_ingest_error_lines_and_edges_of_the_buffer(built_lib):
    ids = np.array([1, 2, 5, 5], dtype=np.uint64)           # duplicate id: the first index wins
    ok = b"1\t2\t0\t-\t-\t+\t+\t50\t-\t100\t-\ts\ts"
    for text in (ok, ok + b"\n", b"\n" + ok, b"\n\n", b"", b"\n", ok + b"\n" + ok, b"5\t1\t0\t-\t-\t+\t+\t50\t-\t100\t-\ts\ts\n"):
        st = _check_against_oracle(text, ids, min_overlap_len=60)
    bad = [b"1\t2\t-1\t-\t-\t+\t+\t50\t-\t100\t-\ts\ts", b"1\t2\t0\t-\t-\t*\t+\t50\t-\t100\t-\ts\ts",
           b"1\t2\t0\t-\t-\t+\t+\t101\t-\t100\t-\ts\ts", b"1\t2\t0\t-\t-\t+\t+\t50\t-\t-5\t-\ts\ts",
           b"1\t2\t0\t-\t-\t+\t+\t50\t-\t100\t-\ts\tq", b"1\t2\t0\t-\t1\t+\t+\t50\t-\t100\t-\ts\ts",
           b"1\t2\t0\t0\t-\t+\t+\t50\t0\t100\t0\tp\tp", b"1\t2\t0\t-\t-\t+\t+\t50\t-\t100\t-\ts\ts\r",
           b"1\t2\t0\t-\t\t+\t+\t50\t-\t100\t-\ts\ts", b"1\t9\t0\t-\t-\t+\t+\t50\t-\t100\t-\ts\ts",
           b"1\t2\t3000000000\t-\t-\t+\t+\t50\t-\t100\t-\ts\ts", b"1\t2\t0\t-\t-\t+\t+\t50\t-\t99999999999999999999\t-\ts\ts"]
    for b in bad:
        st = _check_against_oracle(ok + b"\n" + b + b"\n" + ok + b"\n", ids, min_overlap_len=60)
        assert int(st["first_error_line"]) == 1
    m = capi.IdMap(ids)
    with pytest.raises(capi.HcError) as e:
        m.ingest((ok + b"\n") * 10, F.make_ingest_params(60), cand_cap=3)
    assert e.value.code == -5 and e.value.required == (10, 0)


def test_ingest_many_tiles(built_lib):
    """A few MB of text: many 4 KB newline tiles and many 1024-line compaction blocks."""
    g = load_golden("synth_all_types")
    text = W.fuzz_overlap_text(np.tile(g.cands, 12), g.rs.ids, seed=21)
    st = _check_against_oracle(text, g.rs.ids, min_overlap_len=70, min_overlap_perc=10)
    assert int(st["n_lines"]) > 60000 and len(text) > 2_500_000


@pytest.mark.parametrize("piece", ["700", "5000", "100000"])
def test_ingest_piece_pipeline(built_lib, monkeypatch, piece):
    """Large buffers are cut into pieces at line ends and pipelined (copy in / kernels / copy out); the piece size must not
    change anything: records, order, line numbers, counts, first error, max_overlaps, capacity report."""
    g = load_golden("synth_all_types")
    text = W.fuzz_overlap_text(np.tile(g.cands, 2), g.rs.ids, seed=31)
    monkeypatch.setenv("HC_INGEST_PIECE", piece)
    _check_against_oracle(text, g.rs.ids, min_overlap_len=80, min_overlap_perc=20)
    _check_against_oracle(text, g.rs.ids, min_overlap_len=80, max_overlaps=3333)
    _check_against_oracle(text, g.rs.ids[: len(g.rs.ids) // 2], min_overlap_len=80)          # unknown ids somewhere in the middle
    lines = text.split(b"\n")
    lines[len(lines) // 2] = b"1\t2\t-1\t-\t-\t+\t+\t50\t-\t100\t-\ts\ts"                      # a line the reference exits on
    _check_against_oracle(b"\n".join(lines), g.rs.ids, min_overlap_len=80)
    _check_against_oracle(text + b"x" * 3000, g.rs.ids, min_overlap_len=80)                   # a last line longer than a piece, no newline
    m = capi.IdMap(g.rs.ids)
    with pytest.raises(capi.HcError) as e:
        m.ingest(text, F.make_ingest_params(80), cand_cap=10, filtered_cap=10)
    cand, _, filt, _, st = m.ingest(text, F.make_ingest_params(80))
    assert e.value.code == -5 and e.value.required == (len(cand), len(filt))
