#!/usr/bin/env python
"""A fixture for the one place where OverlapGraph::sortEdges (src/OverlapGraph.cpp:722-764) leaves the order to std::sort:
adjacency lists of more than 16 edges that hold edges with equal (non-overlap length, vertex2).  Reads that are (AT)n repeats
equal their own reverse complement, so every pair of them overlaps perfectly in both orientation classes at every even
shift: a hub read gets dozens of edges, two per partner with the same key.  Writes tests/golden/ties_at_repeats.npz (the
usual fixture: candidates, the reference's results, its graph) and tests/golden/sorted_ties_at_repeats.npz (its lists and
adj_in after sortEdges).  Run in the build container."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from haploconduct_b200 import formats as F  # noqa: E402
import make_golden as G  # noqa: E402
import make_golden_sorted as GS  # noqa: E402


def main():
    rng = np.random.RandomState(5)
    n, L = 60, 150
    singles = []
    for i in range(n):
        # every read is a window of the infinite (AT)n repeat that starts with 'A': any two overlap exactly at every even shift
        q = "".join(chr(33 + int(x)) for x in rng.randint(30, 41, L))
        singles.append((i, "AT" * (L // 2), q))
    rs = F.ReadSet.from_lists(singles, [])
    rows = []
    for a in range(n):
        for b in range(n):
            if a == b:
                continue
            for shift in (int(s) for s in rng.choice(np.arange(0, 60, 2), 2, replace=False)):
                if shift == 0 and a > b:
                    continue                             # a pair at shift 0 is listed once
                for o2 in (1, 0):                        # both orientation classes: same position, same length, same key in sortEdges
                    ov = L - shift
                    rows.append((a, b, shift, 0, ov, 0, 100, 0, ord("-"), 1, o2, ord("s"), ord("s"), 0))
    cands = np.array(rows, dtype=F.CANDIDATE)
    cands = cands[rng.permutation(len(cands))[: 9000]]
    cands = cands[np.lexsort((cands["idx2"], cands["idx1"]))]
    ps = dict(edge_threshold=0.97, min_overlap_len=80)
    G.run_case("ties_at_repeats", rs, cands, ps)
    GS.main(names=("ties_at_repeats",))


if __name__ == "__main__":
    main()
