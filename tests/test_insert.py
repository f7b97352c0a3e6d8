"""CPU suite: the restatement of the serial graph insert (oracle.graph_insert, src/EdgeCalculator.cpp:441-545)
against what the UNMODIFIED reference built: adjacency lists byte for byte, dup_count, inclusion_count
and -- under --ignore_inclusions -- OverlapGraph::inclusions."""
import numpy as np
import pytest

from oracle import oracle as O
from util import golden_names, load_golden, load_insert_golden, random_insert_edges


@pytest.mark.parametrize("name", golden_names())
def test_graph_insert_restatement_matches_reference(name):
    g = load_golden(name)
    nv, ref_inc = load_insert_golden(name)
    e = O.normalise_ref_edges(g.ref_cands)
    win, inc, dups, incl, adj = O.graph_insert(e, nv, ignore_inclusions=True)
    assert e[adj].tobytes() == g.ref_graph.tobytes()
    assert [int(win.sum()), dups, incl] == g.ref_counts.tolist()
    assert np.array_equal(np.nonzero(inc)[0], ref_inc)
    assert not O.graph_insert(e, nv, ignore_inclusions=False)[1].any()


def test_graph_insert_is_a_per_key_argmax():
    """The property the device kernel relies on: the fold equals 'last maximum per key' of one total order."""
    e = random_insert_edges(3, 4000, 40)
    win, _, dups, _, _ = O.graph_insert(e, 40)
    best = {}
    for i in range(len(e)):
        x = e[i]
        k = (min(int(x["v1"]), int(x["v2"])), max(int(x["v1"]), int(x["v2"])), bool(x["ori1"] == x["ori2"]))
        rank = (float(x["score"]), int(x["len1"]) + int(x["len2"]), -float(x["mismatch_rate"]), -int(x["v1"]), int(x["ori1"]),
                int(x["ori2"]), -int(x["pos1"]), -int(x["pos2"]), i)
        if k not in best or rank > best[k][0]:
            best[k] = (rank, i)
    want = np.zeros(len(e), dtype=bool)
    want[[v[1] for v in best.values()]] = True
    assert np.array_equal(win, want)
    assert dups == len(e) - len(best)
    assert dups > 3000      # the case is dense in duplicates
