"""GPU suite: hc_build_adjacency (adjacency lists in insertion order, OverlapGraph::sortEdges, adj_in) against the
reference's own lists (tests/golden/sorted_*.npz) and, on random multigraphs with hubs and ties, against the pinned
restatement oracle.sort_edges."""
import os

import numpy as np
import pytest

from haploconduct_b200 import capi
from oracle import oracle as O
from util import GOLDEN, load_golden

pytestmark = pytest.mark.gpu
NAMES = sorted(f[len("sorted_"):-4] for f in os.listdir(GOLDEN) if f.startswith("sorted_"))


def _adj_edges(graph, read_len):
    e = np.zeros(len(graph), dtype=capi.ADJ_EDGE)
    e["vertex1"], e["vertex2"] = graph["v1"], graph["v2"]
    ov = graph["len1"].astype(np.int64) + graph["len2"].astype(np.int64)
    e["nonoverlap_len"] = (read_len[graph["v1"].astype(np.int64)] + read_len[graph["v2"].astype(np.int64)] - 2 * ov) & 0xffffffff
    return e


def _in_lists(in_off, in_src):
    deg = np.diff(in_off.astype(np.int64))
    vs = np.nonzero(deg)[0]
    return vs, np.append(in_off.astype(np.int64)[vs], int(in_off[-1])), in_src.astype(np.int64)


@pytest.mark.parametrize("name", NAMES)
def test_sorted_lists_equal_the_reference(built_lib, name):
    g = load_golden(name)
    z = np.load(os.path.join(GOLDEN, "sorted_" + name + ".npz"))
    read_len = g.rs.descs["seq_len"].astype(np.int64).sum(axis=1)
    # shuffle the edges and add dropped ones in between: the lists must come out in INPUT order (sort=False) resp. sortEdges order
    rng = np.random.default_rng(5)
    ref = g.ref_graph
    e = _adj_edges(ref, read_len)
    n = len(e)
    junk = e[rng.integers(0, n, n // 3)]
    total = n + len(junk)
    # keep the relative order of the kept edges (it IS the insertion order), interleave the dropped ones at random places
    pos = np.sort(rng.choice(total, n, replace=False))
    mixed = np.zeros(total, dtype=capi.ADJ_EDGE)
    mk = np.zeros(total, np.uint8)
    mixed[pos], mk[pos] = e, 1
    mixed[np.setdiff1d(np.arange(total), pos)] = junk
    src_index = np.full(total, -1, np.int64)
    src_index[pos] = np.arange(n)
    V = g.rs.n_reads
    out_off, perm, in_off, in_src, ties = capi.build_adjacency(mixed, V, keep=mk, sort=False)
    assert np.array_equal(src_index[perm], np.arange(n))                     # adjacency order == the reference's insertion order
    assert np.array_equal(np.diff(out_off.astype(np.int64)), np.bincount(ref["v1"].astype(np.int64), minlength=V))
    out_off, perm, in_off, in_src, ties = capi.build_adjacency(mixed, V, keep=mk, sort=True)
    if name.startswith("ties_"):
        # every list here has more than 16 edges and equal keys: the device reports them (the host mirror then runs std::sort
        # on them, tests/test_gpu_host_mirror.py) and orders equal keys by input index, like the pinned restatement
        assert ties.all()
        mine, _, n_ties = O.sort_edges(ref, read_len)
        assert n_ties == V and ref[src_index[perm]].tobytes() == mine.tobytes()
        return
    assert not ties.any()
    assert ref[src_index[perm]].tobytes() == z["ref_sorted"].tobytes()         # sortEdges, every field of every list
    vs, off, src = _in_lists(in_off, in_src)
    assert np.array_equal(vs, z["in_vertices"]) and np.array_equal(off, z["in_off"]) and np.array_equal(src, z["in_src"])


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_multigraph_with_hubs_and_ties(built_lib, seed):
    rng = np.random.default_rng(seed)
    V, n = 3000, 120000
    g = np.zeros(n, dtype=O.REF_EDGE)
    hub = rng.random(n) < 0.3
    g["v1"] = np.where(hub, rng.integers(0, 8, n), rng.integers(0, V, n))     # eight vertices with thousands of edges (heap sort path)
    g["v2"] = rng.integers(0, V, n)
    g["len1"] = rng.integers(10, 14, n)                                       # few distinct overlap lengths: many equal keys
    read_len = rng.integers(100, 104, V).astype(np.int64)
    order = np.argsort(g["v1"], kind="stable")                                # adjacency order for the restatement
    ref_sorted, (rv, ro, rs_), ties_long = O.sort_edges(g[order], read_len)
    e = _adj_edges(g, read_len)
    out_off, perm, in_off, in_src, ties = capi.build_adjacency(e, V, sort=True)
    assert g[perm].tobytes() == ref_sorted.tobytes()                          # ties broken by input order, like the restatement
    assert int(ties.sum()) == ties_long and ties_long > 0
    vs, off, src = _in_lists(in_off, in_src)
    assert np.array_equal(vs, rv) and np.array_equal(off, ro) and np.array_equal(src, rs_)
    out_off, perm, _, _, _ = capi.build_adjacency(e, V, sort=False)
    assert np.array_equal(perm, order)


def test_adjacency_empty_and_bad_input(built_lib):
    out_off, perm, in_off, in_src, ties = capi.build_adjacency(np.zeros(0, dtype=capi.ADJ_EDGE), 10, sort=True)
    assert len(perm) == 0 and not out_off.any() and not in_off.any()
    e = np.zeros(4, dtype=capi.ADJ_EDGE)
    e["vertex2"][2] = 99
    with pytest.raises(capi.HcError):
        capi.build_adjacency(e, 10)
