"""CPU suite: the restatement of OverlapGraph::sortEdges (oracle.sort_edges, src/OverlapGraph.cpp:722-764) against what the
unmodified reference's sortEdges() left in adj_out / adj_in (tests/golden/sorted_*.npz, oracle/make_golden_sorted.py)."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from util import GOLDEN, load_golden

NAMES = sorted(f[len("sorted_"):-4] for f in os.listdir(GOLDEN) if f.startswith("sorted_"))


def _key_groups(rows, read_len):
    """Rows split into runs of equal (v1, non-overlap length, v2), each run as a sorted list of its rows' bytes."""
    ov = rows["len1"].astype(np.int64) + rows["len2"].astype(np.int64)
    nol = (read_len[rows["v1"].astype(np.int64)] + read_len[rows["v2"].astype(np.int64)] - 2 * ov) & 0xffffffff
    out, k = [], 0
    while k < len(rows):
        j = k
        while j < len(rows) and (rows["v1"][j], nol[j], rows["v2"][j]) == (rows["v1"][k], nol[k], rows["v2"][k]):
            j += 1
        out.append(((int(rows["v1"][k]), int(nol[k]), int(rows["v2"][k])), sorted(rows[i].tobytes() for i in range(k, j))))
        k = j
    return out


@pytest.mark.parametrize("name", NAMES)
def test_sort_edges_restatement_is_pinned(name):
    g = load_golden(name)
    z = np.load(os.path.join(GOLDEN, "sorted_" + name + ".npz"))
    read_len = g.rs.descs["seq_len"].astype(np.int64).sum(axis=1)
    mine, (vs, off, src), ties = O.sort_edges(g.ref_graph, read_len)
    if name.startswith("ties_"):
        # lists of more than 16 edges with equal keys: std::sort decides the order inside a run of equal keys (the restatement
        # keeps the list order there); everything else -- which key comes where, which edges carry it -- is pinned
        assert ties == g.rs.n_reads
        assert _key_groups(mine, read_len) == _key_groups(z["ref_sorted"], read_len)
        assert mine.tobytes() != z["ref_sorted"].tobytes()
        return
    assert ties == 0
    assert mine.tobytes() == z["ref_sorted"].tobytes()
    assert np.array_equal(vs, z["in_vertices"]) and np.array_equal(off, z["in_off"]) and np.array_equal(src, z["in_src"])
    assert (mine["v2"] != g.ref_graph["v2"]).any()          # the sort does move edges in these fixtures
