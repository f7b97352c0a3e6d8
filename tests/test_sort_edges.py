"""CPU suite: the restatement of OverlapGraph::sortEdges (oracle.sort_edges, src/OverlapGraph.cpp:722-764) against what the
unmodified reference's sortEdges() left in adj_out / adj_in (tests/golden/sorted_*.npz, oracle/make_golden_sorted.py)."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from util import GOLDEN, load_golden

NAMES = sorted(f[len("sorted_"):-4] for f in os.listdir(GOLDEN) if f.startswith("sorted_"))


@pytest.mark.parametrize("name", NAMES)
def test_sort_edges_restatement_is_pinned(name):
    g = load_golden(name)
    z = np.load(os.path.join(GOLDEN, "sorted_" + name + ".npz"))
    read_len = g.rs.descs["seq_len"].astype(np.int64).sum(axis=1)
    mine, (vs, off, src), ties = O.sort_edges(g.ref_graph, read_len)
    assert ties == 0
    assert mine.tobytes() == z["ref_sorted"].tobytes()
    assert np.array_equal(vs, z["in_vertices"]) and np.array_equal(off, z["in_off"]) and np.array_equal(src, z["in_src"])
    assert (mine["v2"] != g.ref_graph["v2"]).any()          # the sort does move edges in these fixtures
