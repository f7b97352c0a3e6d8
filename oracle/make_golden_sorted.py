#!/usr/bin/env python
"""Golden vectors for OverlapGraph::sortEdges (src/OverlapGraph.cpp:722-764): the adjacency lists and adj_in of the
UNMODIFIED reference after construct_edges() + sortEdges() (oracle/_ref/ref_driver --run --dump-sorted), for the read
sets / candidate lists of existing fixtures.  Writes tests/golden/sorted_<name>.npz.  Run in the build container
(needs /root/reference compiled: make -C oracle)."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from haploconduct_b200 import formats as F  # noqa: E402
from oracle import oracle as O  # noqa: E402
from util import load_golden  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def main(names=("c1_savage_example_full", "c2_polyte_example_full", "synth_all_types", "synth_stage_c_contigs")):
    for name in names:
        try:
            g = load_golden(name)
        except Exception as ex:
            print("skip", name, ex)
            continue
        rs = g.rs
        d = tempfile.mkdtemp(prefix="hc_sorted_")
        F.write_fastq_set(rs, d + "/s.fastq", d + "/p1.fastq", d + "/p2.fastq")
        F.write_overlaps(d + "/ov.txt", g.cands, rs.ids)
        kw = dict(singles=d + "/s.fastq" if rs.n_single else None, paired1=d + "/p1.fastq" if rs.n_reads > rs.n_single else None,
                  paired2=d + "/p2.fastq" if rs.n_reads > rs.n_single else None)
        out = O.run_ref(d, d + "/ov.txt", run=True, dump_graph=True, dump_sorted=True, threads=1, **kw, **g.ps)
        assert out["graph"].tobytes() == g.ref_graph.tobytes()
        vs, off, src = out["adj_in"]
        np.savez_compressed(os.path.join(GOLDEN, "sorted_" + name + ".npz"), ref_sorted=out["sorted_graph"], in_vertices=vs, in_off=off, in_src=src)
        read_len = rs.descs["seq_len"].astype(np.int64).sum(axis=1)
        mine, (mv, mo, ms), ties = O.sort_edges(out["graph"], read_len)
        ok = mine.tobytes() == out["sorted_graph"].tobytes() and np.array_equal(mv, vs) and np.array_equal(mo, off) and np.array_equal(ms, src)
        print("%-28s edges=%d moved=%d ties_in_long_lists=%d restatement_equal=%s" %
              (name, len(out["graph"]), int((out["graph"]["v2"] != out["sorted_graph"]["v2"]).sum()), ties, ok))


if __name__ == "__main__":
    main()
