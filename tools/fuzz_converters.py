#!/usr/bin/env python
"""Randomised comparison of lib/hc_sfo2overlaps and lib/hc_sam2overlaps with the reference's scripts (run from the temporary
Python-3 copies of oracle/make_golden_s*.py; build container only): output files, stdout and exit codes must be equal.
    python tools/fuzz_converters.py [--seeds 12]"""
import argparse
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import make_golden_sam as GM  # noqa: E402
import make_golden_sfo as GS  # noqa: E402
from haploconduct_b200 import build as B  # noqa: E402


def same(r1, r2, d):
    return r1.returncode == r2.returncode and (r1.returncode != 0 or (open(d + "/ref.txt", "rb").read() == open(d + "/mine.txt", "rb").read() and r1.stdout == r2.stdout))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=12)
    a = ap.parse_args()
    d = tempfile.mkdtemp(prefix="hc_fuzz_conv_")
    s_sfo, s_sam = GS.py3_copy(d), GM.py3_copy(d)
    bad = 0
    for seed in range(100, 100 + a.seeds):
        rng = np.random.RandomState(seed)
        ns, npairs = int(rng.randint(0, 60)), int(rng.randint(0, 60))
        if ns + npairs < 2:
            continue
        open(d + "/in.sfo", "w").write(GS.make_sfo(seed, ns, npairs, int(rng.randint(50, 3000))))
        args = ["--in", "in.sfo", "--num_singles", str(ns), "--num_pairs", str(npairs)]
        r1 = subprocess.run([sys.executable, s_sfo] + args + ["--out", "ref.txt"], cwd=d, env=dict(os.environ, LC_ALL="C"), stdout=subprocess.PIPE)
        r2 = subprocess.run([os.path.join(B.LIBDIR, "hc_sfo2overlaps")] + args + ["--out", "mine.txt"], cwd=d, stdout=subprocess.PIPE)
        if not same(r1, r2, d):
            bad += 1
            print("sfo2overlaps: seed %d differs" % seed)
    for seed in range(200, 200 + a.seeds):
        rng = np.random.RandomState(seed)
        ns, npairs = int(rng.randint(0, 300)), int(rng.randint(0, 300))
        if ns + npairs == 0:
            continue
        refs = [("r%d" % k, int(rng.randint(300, 1500))) for k in range(int(rng.randint(1, 4)))]
        fasta, sam_s, sam_p = GM.make_case(seed, ns, npairs, refs)
        for fn, t in (("ref.fasta", fasta), ("s.sam", sam_s), ("p.sam", sam_p)):
            open(d + "/" + fn, "w").write(t)
        args = ["--ref", "ref.fasta", "--min_overlap_len", str(int(rng.randint(0, 80)))] + (["--sam_s", "s.sam"] if ns else []) + \
               (["--sam_p", "p.sam"] if npairs else []) + (["--verbose"] if seed % 2 else [])
        r1 = subprocess.run([sys.executable, s_sam] + args + ["--out", "ref.txt"], cwd=d, stdout=subprocess.PIPE)
        r2 = subprocess.run([os.path.join(B.LIBDIR, "hc_sam2overlaps")] + args + ["--out", "mine.txt"], cwd=d, stdout=subprocess.PIPE)
        if not same(r1, r2, d):
            bad += 1
            print("sam2overlaps: seed %d differs" % seed)
    print("%d differences in %d + %d cases" % (bad, a.seeds, a.seeds))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
