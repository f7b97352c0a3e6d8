#!/usr/bin/env python
"""lib/hc_sam2overlaps next to the reference's scripts/sam2overlaps.py (run from the temporary Python-3 copy that
oracle/make_golden_sam.py makes; build container only) on seeded SAM files.   python tools/bench_sam2overlaps.py [--reads 20000]"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from haploconduct_b200 import build as B  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=20000)
    a = ap.parse_args()
    import make_golden_sam as G
    d = tempfile.mkdtemp(prefix="hc_sam_bench_")
    fasta, sam_s, sam_p = G.make_case(21, a.reads, a.reads, [("refA", 9000)])
    for fn, text in (("ref.fasta", fasta), ("s.sam", sam_s), ("p.sam", sam_p)):
        open(os.path.join(d, fn), "w").write(text)
    args = ["--ref", "ref.fasta", "--sam_s", "s.sam", "--sam_p", "p.sam", "--min_overlap_len", "50"]
    t0 = time.perf_counter()
    subprocess.run([os.path.join(B.LIBDIR, "hc_sam2overlaps")] + args + ["--out", "mine.txt"], cwd=d, check=True, stdout=subprocess.DEVNULL)
    t_mine = time.perf_counter() - t0
    res = {"metric": "sam2overlaps wall time", "alignments": 3 * a.reads, "threads": os.cpu_count(), "hc_sam2overlaps_s": t_mine,
           "overlap_lines": sum(1 for _ in open(os.path.join(d, "mine.txt")))}
    if os.path.exists(G.REF_SCRIPT):
        script = G.py3_copy(d)
        t0 = time.perf_counter()
        subprocess.run([sys.executable, script] + args + ["--out", "ref.txt"], cwd=d, check=True, stdout=subprocess.DEVNULL)
        res["reference_script_s"] = time.perf_counter() - t0
        res["speedup"] = res["reference_script_s"] / t_mine
        res["identical_output"] = open(os.path.join(d, "mine.txt"), "rb").read() == open(os.path.join(d, "ref.txt"), "rb").read()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
