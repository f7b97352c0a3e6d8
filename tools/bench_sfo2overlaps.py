#!/usr/bin/env python
"""lib/hc_sfo2overlaps next to the reference's scripts/sfo2overlaps.py (run from the temporary Python-3 copy that
oracle/make_golden_sfo.py makes; build container only) on a seeded SFO file.   python tools/bench_sfo2overlaps.py [--lines 400000]"""
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
    ap.add_argument("--lines", type=int, default=400000)
    a = ap.parse_args()
    import make_golden_sfo as G
    d = tempfile.mkdtemp(prefix="hc_sfo_bench_")
    ns, npairs = 20000, 60000
    text = G.make_sfo(11, ns, npairs, a.lines)
    open(os.path.join(d, "in.sfo"), "w").write(text)
    args = ["--in", "in.sfo", "--num_singles", str(ns), "--num_pairs", str(npairs)]
    t0 = time.perf_counter()
    subprocess.run([os.path.join(B.LIBDIR, "hc_sfo2overlaps")] + args + ["--out", "mine.txt"], cwd=d, check=True, stdout=subprocess.DEVNULL)
    t_mine = time.perf_counter() - t0
    res = {"metric": "sfo2overlaps wall time", "sfo_lines": text.count("\n"), "threads": os.cpu_count(), "hc_sfo2overlaps_s": t_mine}
    if os.path.exists(G.REF_SCRIPT):
        script = G.py3_copy(d)
        t0 = time.perf_counter()
        subprocess.run([sys.executable, script] + args + ["--out", "ref.txt"], cwd=d, env=dict(os.environ, LC_ALL="C"), check=True, stdout=subprocess.DEVNULL)
        res["reference_script_s"] = time.perf_counter() - t0
        res["speedup"] = res["reference_script_s"] / t_mine
        res["identical_output"] = open(os.path.join(d, "mine.txt"), "rb").read() == open(os.path.join(d, "ref.txt"), "rb").read()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
