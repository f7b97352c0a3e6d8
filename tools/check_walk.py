#!/usr/bin/env python
"""Anchor walk of hc_score_kernel against the lane-chunk rounds and the oracle (GPU box).

The two device paths add the same integers, so every per-candidate field -- scores included -- must be bit-identical
whichever path scored a window.  Workloads: config-4 shaped runs (P-P), singles sorted by read (S-S runs), random lists
with the walk forced onto tiles of any size.    python tools/check_walk.py [--pairs 60000]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from haploconduct_b200 import capi, formats as F, workloads as W, workloads_torch as WT  # noqa: E402
from oracle import oracle as O  # noqa: E402


def run(st, p, cands, walk, walk_min=None, compact=False):
    os.environ.pop("HC_ANCHOR_WALK", None)
    os.environ.pop("HC_ANCHOR_WALK_MIN", None)
    if walk:
        os.environ["HC_ANCHOR_WALK"] = "1"
    if walk_min is not None:
        os.environ["HC_ANCHOR_WALK_MIN"] = str(walk_min)
    return st.score_batch(p, cands, compact=compact)


def compare(tag, a, b):
    ea, na, pa, _ = a
    eb, nb, pb, _ = b
    bad = 0
    for f in pa.dtype.names:
        x, y = pa[f], pb[f]
        neq = (x != y) & ~((x != x) & (y != y)) if x.dtype.kind == "f" else (x != y)
        if neq.ndim > 1:
            neq = neq.any(axis=1)
        if neq.any():
            i = int(np.nonzero(neq)[0][0])
            print("  %s: field %s differs for %d candidates, first %d: %r vs %r" % (tag, f, int(neq.sum()), i, pa[i], pb[i]))
            bad += 1
    if ea.tobytes() != eb.tobytes() or not np.array_equal(na, nb):
        print("  %s: edge / non-edge lists differ" % tag)
        bad += 1
    print("%s: %s (%d candidates, %d edges, %d non-edges)" % (tag, "IDENTICAL" if not bad else "MISMATCH", len(pa), len(ea), len(na)))
    return bad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=60000)
    a = ap.parse_args()
    bad = 0
    # ---- config-4 shaped: runs of ~100 P-P candidates per read pair
    pr = WT.make_paired_reads(a.pairs, read_len=150, genome_len=max(600, a.pairs // 100), seed=7, device="cpu")
    rs = pr.readset()
    cands = WT.candidates_as_numpy(WT.make_pp_candidates(pr, D=140, max_cands=3_000_000))
    p = F.make_params(edge_threshold=0.97, ov_threshold=0.9, merge_contigs=0.0, mismatch=0.0, min_read_len=0)
    with capi.Store(rs) as st:
        print("store: %d quality codes" % st.quality_alphabet)
        r_walk = run(st, p, cands, True)
        r_gen = run(st, p, cands, False)
        bad += compare("C4-shaped, hc_candidate", r_walk, r_gen)
        fits = (cands["pos1"] < (1 << 14)) & (cands["pos2"] < (1 << 14))
        bad += compare("C4-shaped, run-encoded", run(st, p, cands[fits], True, compact="runs"), run(st, p, cands[fits], False))
        pm = p.copy(); pm["mismatch"] = 0.2
        bad += compare("C4-shaped, mismatch=0.2 (void)", run(st, pm, cands[:400000], True), run(st, pm, cands[:400000], False))
        sub = cands[:40000]
        ref, _ = O.score_batch(rs, p, sub)
        per = run(st, p, sub, True)[2]
        ok = (np.array_equal(per["cls"], ref["cls"]) and np.array_equal(per["mismatches"], ref["mismatches"]) and
              np.array_equal(per["compared"], ref["compared"]) and np.allclose(per["score"], ref["score"], rtol=1e-6, atol=0))
        print("C4-shaped vs oracle on %d candidates: %s" % (len(sub), "ok" if ok else "MISMATCH"))
        bad += 0 if ok else 1
    # ---- random geometry lists (all read types / orientations / N), the walk forced onto every tile
    for seed in range(2000, 2012):
        rng = np.random.RandomState(seed)
        ss = W.synth_readset(int(rng.randint(0, 120)) + 20, int(rng.randint(0, 120)), genome_len=int(rng.randint(800, 5000)),
                             read_len=(60, 400), qmax=int(rng.choice([41, 41, 60])), q_lo=int(rng.choice([0, 2, 20])), seed=seed,
                             n_rate=float(rng.choice([0.0, 0.0, 0.002])), flip_fraction=float(rng.choice([0.0, 0.3])))
        c = W.geometry_candidates(ss, int(rng.randint(3000, 9000)), seed=seed + 1, junk_fraction=float(rng.choice([0.0, 0.15])),
                                  min_ov=int(rng.choice([10, 40])))
        if rng.rand() < 0.7:   # sorted by (min, max): runs, like an overlaps file
            c = c[np.lexsort((np.maximum(c["idx1"], c["idx2"]), np.minimum(c["idx1"], c["idx2"])))]
        pp = F.make_params(edge_threshold=float(rng.choice([0.9, 0.97, 0.995])), ov_threshold=0.9, merge_contigs=float(rng.choice([0.0, 0.02])),
                           mismatch=float(rng.choice([0.0, 0.0, 0.05])), min_read_len=int(rng.choice([0, 100])))
        with capi.Store(ss.rs) as st:
            g = run(st, pp, c, False)
            bad += compare("random seed %d, walk forced" % seed, run(st, pp, c, True, walk_min=1), g)
            bad += compare("random seed %d, default" % seed, run(st, pp, c, True), g)
    print("RESULT: %s" % ("all identical" if bad == 0 else "%d MISMATCHES" % bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
