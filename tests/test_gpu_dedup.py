"""GPU suite: hc_dedup_edges (duplicate-edge resolution of the graph insert on the device) against the
reference's graphs (tests/golden) and against the sequential restatement on dense random ties."""
import numpy as np
import pytest

from haploconduct_b200 import capi
from oracle import oracle as O
from util import golden_names, load_golden, load_insert_golden, random_insert_edges

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", golden_names())
def test_dedup_reproduces_reference_graph(built_lib, name):
    g = load_golden(name)
    nv, ref_inc = load_insert_golden(name)
    e = O.normalise_ref_edges(g.ref_cands)
    win, inc, dups, incl = capi.dedup_edges(O.dedup_records(e), nv, ignore_inclusions=True)
    idx = np.nonzero(win)[0]
    adj = idx[np.argsort(e["v1"][idx], kind="stable")]
    assert e[adj].tobytes() == g.ref_graph.tobytes()            # the reference's adjacency lists, in order
    assert [int(win.sum()), dups, incl] == g.ref_counts.tolist()
    assert np.array_equal(np.nonzero(inc)[0], ref_inc)
    win2, inc2, _, _ = capi.dedup_edges(O.dedup_records(e), nv, ignore_inclusions=False)
    assert np.array_equal(win, win2) and not inc2.any()


@pytest.mark.parametrize("seed,n,nv", [(1, 5000, 30), (2, 200000, 500), (3, 300000, 60000), (4, 1, 2), (5, 64, 2)])
def test_dedup_dense_ties(built_lib, seed, n, nv):
    e = random_insert_edges(seed, n, nv)
    want_win, want_inc, want_dups, want_incl, _ = O.graph_insert(e, nv, ignore_inclusions=True)
    win, inc, dups, incl = capi.dedup_edges(O.dedup_records(e), nv, ignore_inclusions=True)
    assert np.array_equal(win, want_win)
    assert np.array_equal(inc, want_inc)
    assert (dups, incl) == (want_dups, want_incl)


def test_dedup_empty_and_errors(built_lib):
    win, inc, dups, incl = capi.dedup_edges(np.zeros(0, dtype=capi.F.DEDUP_EDGE), 5)
    assert len(win) == 0 and (dups, incl) == (0, 0)
    bad = np.zeros(1, dtype=capi.F.DEDUP_EDGE)
    bad["vertex2"] = 9
    with pytest.raises(capi.HcError):
        capi.dedup_edges(bad, 5)
