#!/usr/bin/env python
"""hc_build_adjacency on a random overlap-graph-shaped input (edges in input order, lists of ~3 edges, a few hubs):
adjacency lists in sortEdges order + adj_in through the host-buffer ABI, next to numpy's lexsort of the same keys on the
host cores (what the restatement oracle.sort_edges spends its time in).   python tools/bench_adjacency.py [--edges 10000000]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from haploconduct_b200 import capi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--edges", type=int, default=10_000_000)
    ap.add_argument("--vertices", type=int, default=3_000_000)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    rng = np.random.default_rng(3)
    e = np.zeros(a.edges, dtype=capi.ADJ_EDGE)
    e["vertex1"] = np.sort(rng.integers(0, a.vertices, a.edges))            # an overlaps file lists a read's overlaps together
    e["vertex2"] = rng.integers(0, a.vertices, a.edges)
    e["nonoverlap_len"] = rng.integers(0, 300, a.edges)
    keep = (rng.random(a.edges) < 0.9).astype(np.uint8)
    capi.build_adjacency(e, a.vertices, keep=keep, sort=True)
    t = []
    for _ in range(a.steps):
        t0 = time.perf_counter()
        out_off, perm, in_off, in_src, ties = capi.build_adjacency(e, a.vertices, keep=keep, sort=True)
        t.append(time.perf_counter() - t0)
    t0 = time.perf_counter()
    k = np.nonzero(keep)[0]
    order = k[np.lexsort((k, e["vertex2"][k], e["nonoverlap_len"][k], e["vertex1"][k]))]
    t_np = time.perf_counter() - t0
    assert np.array_equal(order, perm)
    print(json.dumps({"metric": "adjacency lists in sortEdges order + adj_in (host buffers)", "edges": a.edges, "kept": int(len(perm)), "vertices": a.vertices,
                      "ms": min(t) * 1e3, "edges_per_s": a.edges / min(t), "numpy_lexsort_ms": t_np * 1e3, "identical_order": True}))


if __name__ == "__main__":
    main()
