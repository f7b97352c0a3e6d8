"""CPU suite: the restatement of the overlaps-file text loop (oracle/ingest_oracle.py, src/EdgeCalculator.cpp:581-645
+ src/Overlap.h:39-73) against what the UNMODIFIED reference printed back for irregularly spelled files."""
import pytest

from oracle import ingest_oracle as IO
from util import IngestGolden, ingest_golden_names


@pytest.mark.parametrize("name", ingest_golden_names())
def test_ingest_restatement_matches_reference(name):
    g = IngestGolden(name)
    idmap = {int(v): i for i, v in enumerate(g.ids)}
    status, scored, filtered = IO.ingest(g.text, idmap, **g.kw())
    assert [IO.rec_line(r[1]) for r in scored] == g.ref_scored
    assert [IO.rec_line(r[1]) for r in filtered] == g.ref_filtered
    n_dropped = int(g.ref_counts[3] + g.ref_counts[4])
    assert [len(status), len(scored), len(filtered), status.count(IO.DROPPED), status.count(IO.SKIPPED)] == \
        [int(g.ref_counts[0]), int(g.ref_counts[1]), int(g.ref_counts[2]), n_dropped, int(g.ref_counts[5])]
    assert IO.ERROR not in status and IO.UNKNOWN_ID not in status


def test_c_number_parsers():
    assert IO.c_strtoul0(b"0x1F") == 31 and IO.c_strtoul0(b"017") == 15 and IO.c_strtoul0(b" +12ab") == 12
    assert IO.c_strtoul0(b"0x") == 0 and IO.c_strtoul0(b"08") == 0 and IO.c_strtoul0(b"") == 0
    assert IO.c_strtoul0(b"-1") == 2 ** 64 - 1 and IO.c_strtoul0(b"99999999999999999999999") == 2 ** 64 - 1
    assert IO.c_atoi_u32(b"007") == 7 and IO.c_atoi_u32(b" -3") == 2 ** 32 - 3 and IO.c_atoi_u32(b"12.9") == 12
    assert IO.c_atoi_u32(b"3000000000") == 3000000000 and IO.c_atoi_u32(b"99999999999999999999") == 2 ** 32 - 1
    assert IO.c_atoi_u32(b"-") == 0 and IO.c_atoi_u32(b"") == 0


def test_lines_the_reference_dies_on_are_errors():
    ok = b"1\t2\t0\t-\t-\t+\t+\t50\t-\t100\t-\ts\ts"
    idmap = {1: 0, 2: 1}
    assert IO.ingest(ok, idmap, 60)[0] == [IO.SCORE]
    for bad in (b"1\t2\t-1\t-\t-\t+\t+\t50\t-\t100\t-\ts\ts",       # pos < 0
                b"1\t2\t0\t-\t-\t*\t+\t50\t-\t100\t-\ts\ts",        # ori
                b"1\t2\t0\t-\t-\t+\t+\t101\t-\t100\t-\ts\ts",       # perc > 100
                b"1\t2\t0\t-\t-\t+\t+\t50\t-\t-5\t-\ts\ts",         # len < 0
                b"1\t2\t0\t-\t-\t+\t+\t50\t-\t100\t-\ts\tq",        # type
                b"1\t2\t0\t-\t1\t+\t+\t50\t-\t100\t-\ts\ts",        # ord 1 with single-end types
                b"1\t2\t0\t0\t-\t+\t+\t50\t0\t100\t0\tp\tp",        # ord - with paired types
                b"1\t2\t0\t-\t-\t+\t+\t50\t-\t100\t-\ts\ts\r",      # carriage return stays in TYPE2
                b"1\t2\t0\t-\t\t+\t+\t50\t-\t100\t-\ts\ts"):        # empty ORD
        assert IO.ingest(bad, idmap, 60)[0] == [IO.ERROR], bad
    assert IO.ingest(b"1\t3\t0\t-\t-\t+\t+\t50\t-\t100\t-\ts\ts", idmap, 60)[0] == [IO.UNKNOWN_ID]
    assert IO.ingest(b"1\t3\t0\t-\t-\t+\t+\t50\t-\t10\t-\ts\ts", idmap, 60)[0] == [IO.NONEDGE]   # ids only matter when scored
