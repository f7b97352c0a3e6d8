#!/usr/bin/env python
"""Randomised parity sweep on the GPU box: many seeded read sets / candidate lists / parameter sets through
hc_score_batch (all record formats, both store layouts, fast and exact-score mode) against the pinned C oracle.
Not part of the test-suite (minutes of GPU time); prints one line per seed and a summary.  Uses oracle/ as the
checker only.

    python tools/fuzz_parity.py [--seeds 40]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from haploconduct_b200 import capi, formats as F, workloads as W  # noqa: E402
from oracle import oracle as O  # noqa: E402
from util import assert_results_match  # noqa: E402


def one(seed: int) -> str:
    rng = np.random.RandomState(seed)
    ns, npair = int(rng.randint(0, 150)), int(rng.randint(0, 150))
    if ns + npair < 20:
        ns += 40
    long_reads = rng.rand() < 0.3
    qmax = int(rng.choice([41, 41, 60, 93]))
    ss = W.synth_readset(ns, npair, genome_len=int(rng.randint(3000, 8000)) if long_reads else int(rng.randint(800, 6000)), read_len=(300, 2500) if long_reads else (60, 260),
                         qmax=qmax, q_lo=int(rng.choice([0, 2, 20])), seed=seed, n_rate=float(rng.choice([0.0, 0.002, 0.05])),
                         flip_fraction=float(rng.choice([0.0, 0.3])))
    cands = W.geometry_candidates(ss, int(rng.randint(2000, 12000)), seed=seed + 1, junk_fraction=float(rng.choice([0.0, 0.15, 0.4])),
                                  min_ov=int(rng.choice([10, 40, 100])))
    p = F.make_params(edge_threshold=float(rng.choice([0.9, 0.95, 0.97, 0.995, 1.0, 0.0])), ov_threshold=float(rng.choice([0.9, 0.5, 0.0])),
                      merge_contigs=float(rng.choice([0.0, 0.0, 0.01, 0.05])), mismatch=float(rng.choice([0.0, 0.0, 0.01, 0.2])),
                      min_read_len=int(rng.choice([0, 0, 100, 300])))
    ref, _ = O.score_batch(ss.rs, p, cands)
    for layout in ("auto", "planar"):
        os.environ["HC_STORE_LAYOUT"] = layout
        os.environ["HC_HOST_CHUNK"] = str(int(rng.choice([777, 4096, 10 ** 8])))
        os.environ["HC_HOST_WHOLE_MAX"] = str(int(rng.choice([0, 1 << 31])))       # two-slot / copy-ahead pipeline
        with capi.Store(ss.rs) as st:
            edges, nonedge, per, stats = st.score_batch(p, cands)
            assert_results_match(per, ref["score"], ref["mismatch_rate"], ref["pos3"], ref["pos4"], ref["cls"], what="seed %d %s" % (seed, layout))
            assert np.array_equal(per["mismatches"], ref["mismatches"]) and np.array_equal(per["compared"], ref["compared"])
            assert np.array_equal(per["status"], ref["status"])
            e2, n2, per2, _ = st.score_batch(p, cands, compact=True)
            assert per2.tobytes() == per.tobytes() and e2.tobytes() == edges.tobytes() and np.array_equal(n2, nonedge)
            fits = (cands["pos1"] < (1 << 14)) & (cands["pos2"] < (1 << 14))
            e3, n3, per3, _ = st.score_batch(p, cands[fits], compact="short")
            assert per3.tobytes() == per[fits].tobytes()
            if ss.rs.n_reads < (1 << 31):
                e4, n4, per4, _ = st.score_batch(p, cands[fits], compact="runs")
                assert per4.tobytes() == per3.tobytes() and e4.tobytes() == e3.tobytes() and np.array_equal(n4, n3)
            px = p.copy()
            px["flags"] = F.FLAG_EXACT_EDGE_SCORES
            ex, nx, perx, _ = st.score_batch(px, cands)
            assert np.array_equal(perx["cls"], ref["cls"]) and np.array_equal(nx, nonedge)
            ei = np.nonzero(ref["cls"] == F.CLASS_EDGE)[0]
            assert np.allclose(ex["score"], ref["score"][ei], rtol=1e-14, atol=0)          # exact sums, device exp
    cls = np.bincount(ref["cls"], minlength=3)
    return "seed %3d: %4d reads (%s), %5d candidates, Q<=%d, classes %s ok" % (seed, ss.rs.n_reads, "long" if long_reads else "short", len(cands), qmax, cls.tolist())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=40)
    ap.add_argument("--first", type=int, default=1000)
    a = ap.parse_args()
    for s in range(a.first, a.first + a.seeds):
        print(one(s), flush=True)
    print("all %d seeds match the oracle" % a.seeds)


if __name__ == "__main__":
    main()
