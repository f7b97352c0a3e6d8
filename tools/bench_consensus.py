#!/usr/bin/env python
"""Throughput of hc_consensus on one B200 (host buffers in and out) next to the reference's own
SRBuilder::consensus (oracle/_ref/ref_driver --consensus, one host core) on a sample of the same pile-ups.

    python tools/bench_consensus.py [--problems 200000] [--steps 3]
Prints one JSON line.  Benchmark tool, not product code."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from haploconduct_b200 import capi, formats as F, workloads as W  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--problems", type=int, default=200_000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--cpu-problems", type=int, default=3000)
    a = ap.parse_args()
    # a few thousand distinct pile-ups, tiled: the device does not care, and generation stays in seconds
    rs, base = W.consensus_problems(seed=5, n_problems=2000)
    P0, S0 = F.consensus_arrays(base)
    reps = (a.problems + len(P0) - 1) // len(P0)
    P = np.tile(P0, reps)[: a.problems].copy()
    S = np.tile(S0, reps)
    for r in range(reps):
        sl = slice(r * len(P0), min((r + 1) * len(P0), a.problems))
        P["seq_begin"][sl] += r * len(S0)
        P["seq_end"][sl] += r * len(S0)
    P["out_offset"] = np.concatenate(([0], np.cumsum(P["total_len"][:-1].astype(np.int64)))).astype(np.uint64)
    total = int(P["total_len"].sum())
    lens = rs.descs["seq_len"]
    bases = int(sum(int(lens[e["read"], e["mate"]]) for e in S0)) * reps
    L = capi.lib()
    cs, cq = np.zeros(total, dtype=np.uint8), np.zeros(total, dtype=np.uint8)
    res = np.zeros(len(P), dtype=F.CONS_RESULT)
    with capi.Store(rs) as st:
        def step():
            rc = L.hc_consensus(st.handle, P.ctypes.data, len(P), S.ctypes.data, len(S), 3, 0.9, cs.ctypes.data, cq.ctypes.data, total, res.ctypes.data)
            assert rc == 0, capi.last_error()
        step()
        t = []
        for _ in range(a.steps):
            t0 = time.perf_counter()
            step()
            t.append(time.perf_counter() - t0)
        small = st.consensus(base[: a.cpu_problems], 3, 0.9)
    best = min(t)
    out = {"metric": "consensus columns per second (host buffers in and out)", "unit": "columns/s",
           "problems": len(P), "columns": total, "pileup_bases": bases, "ms": best * 1e3, "value": total / best, "bases_per_s": bases / best,
           "host_threads": os.cpu_count()}
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    if os.path.exists(ref):
        d = tempfile.mkdtemp(prefix="hc_cons_")
        sub = base[: a.cpu_problems]
        with open(d + "/in.txt", "w") as f:
            f.write(W.consensus_problem_text(rs, sub))
        open(d + "/ov.txt", "w").close()
        F.write_fastq_set(rs, d + "/s.fastq", d + "/p1.fastq", d + "/p2.fastq")
        o = subprocess.run([ref, "--overlaps", d + "/ov.txt", "--singles", d + "/s.fastq", "--paired1", d + "/p1.fastq", "--paired2",
                            d + "/p2.fastq", "--min_clique_size", "3", "--min_qual", "0.9", "--consensus", d + "/in.txt", d + "/out.txt"],
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        js = [json.loads(l) for l in o.stdout.split("\n") if l.startswith("{")]
        rows = [l.rstrip("\n").split("\t") for l in open(d + "/out.txt")]
        same = all((int(r[1]), "" if r[2] == "0" else r[3], "" if r[2] == "0" else r[4]) == g for r, g in zip(rows, small))
        cols = sum(p["total_len"] for p in sub)
        if js:
            out["reference"] = {"problems": len(sub), "columns": cols, "pileup_bases": js[-1]["consensus_bases"], "t_consensus_s": js[-1]["t_consensus_s"],
                                "columns_per_s": cols / js[-1]["t_consensus_s"], "bases_per_s": js[-1]["consensus_bases"] / js[-1]["t_consensus_s"],
                                "threads": 1, "identical_output": bool(same)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
