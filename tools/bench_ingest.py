#!/usr/bin/env python
"""Throughput of candidate ingestion (overlaps-file text -> candidates) on one B200, next to the reference's
single-threaded text loop on the host (oracle/_ref/ref_driver --time-scoring reports t_parse_s).

    python tools/bench_ingest.py [--lines 20000000] [--steps 5]

Prints one JSON line.  `device` = hc_ingest_overlaps_device on text resident in HBM (CUDA-event time of the
newline index + parse + compaction kernels), `e2e` = hc_ingest_overlaps on host buffers (copy in, kernels, copy
out), `reference` = the reference's loop on a bounded sample of the same text."""
import argparse
import ctypes
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from haploconduct_b200 import capi, formats as F, workloads as W  # noqa: E402


def make_text(n_lines: int, n_reads: int, seed: int = 1) -> bytes:
    """P-P lines as sfo2overlaps.py writes them (canonical spelling), ids uniform over the store."""
    rng = np.random.RandomState(seed)
    base = min(n_lines, 1_000_000)
    i1 = rng.randint(0, n_reads, base)
    i2 = (i1 + 1 + rng.randint(0, 100, base)) % n_reads
    pos1, pos2 = rng.randint(0, 100, base), rng.randint(0, 100, base)
    l1, l2 = 150 - pos1, 150 - pos2
    ori = rng.randint(0, 2, (base, 2))
    rows = ["%d\t%d\t%d\t%d\t%s\t%s\t%s\t%d\t%d\t%d\t%d\tp\tp" % (a, b, c, d, "12"[o & 1], "+-"[x], "+-"[y], int(e * 100 / 150), int(f * 100 / 150), e, f)
            for a, b, c, d, o, x, y, e, f in zip(i1, i2, pos1, pos2, rng.randint(0, 2, base), ori[:, 0], ori[:, 1], l1, l2)]
    chunk = ("\n".join(rows) + "\n").encode()
    reps = (n_lines + base - 1) // base
    return chunk * reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lines", type=int, default=20_000_000)
    ap.add_argument("--reads", type=int, default=10_000_000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--cpu-lines", type=int, default=2_000_000)
    a = ap.parse_args()
    L = capi.lib()
    text = make_text(a.lines, a.reads)
    n_lines = text.count(b"\n")
    ids = np.arange(a.reads, dtype=np.uint64)
    m = capi.IdMap(ids)
    p = F.make_ingest_params(min_overlap_len=100)
    dev = torch.device("cuda:0")
    h_text = torch.frombuffer(bytearray(text), dtype=torch.uint8).pin_memory()
    d_text = h_text.to(dev)
    d_cand = torch.empty(n_lines * 32, dtype=torch.uint8, device=dev)
    d_filt = torch.empty(n_lines * 48, dtype=torch.uint8, device=dev)
    st = np.zeros(1, dtype=F.INGEST_STATS)

    def device_step():
        rc = L.hc_ingest_overlaps_device(m.handle, None, d_text.data_ptr(), len(text), p.ctypes.data, d_cand.data_ptr(), None, n_lines,
                                         d_filt.data_ptr(), None, n_lines, st.ctypes.data)
        assert rc == 0, capi.last_error()
        return float(st[0]["device_ms"])

    for _ in range(3):
        device_step()
    dms = [device_step() for _ in range(a.steps)]
    torch.cuda.synchronize()
    wall0 = time.perf_counter()
    for _ in range(a.steps):
        device_step()
    torch.cuda.synchronize()
    wall_dev = (time.perf_counter() - wall0) / a.steps * 1e3
    # host buffers end to end
    h_cand = torch.empty(n_lines * 32, dtype=torch.uint8).pin_memory()       # pinned, like the text
    h_filt = torch.empty(n_lines * 48, dtype=torch.uint8).pin_memory()
    cand, filt = h_cand.numpy(), h_filt.numpy()
    buf = h_text.numpy()

    def host_step():
        rc = L.hc_ingest_overlaps(m.handle, buf.ctypes.data, len(text), p.ctypes.data, cand.ctypes.data, None, n_lines, filt.ctypes.data, None,
                                  n_lines, st.ctypes.data)
        assert rc == 0, capi.last_error()

    host_step()
    t0 = time.perf_counter()
    for _ in range(max(a.steps // 2, 1)):
        host_step()
    e2e_ms = (time.perf_counter() - t0) / max(a.steps // 2, 1) * 1e3
    n_scored = int(st[0]["n_scored"])
    out = {"metric": "overlap lines ingested per second", "unit": "lines/s", "lines": n_lines, "text_bytes": len(text),
           "scored": n_scored, "filtered": int(st[0]["n_filtered"]),
           "device": {"ms": float(np.mean(dms)), "wall_ms_incl_alloc": wall_dev, "lines_per_s": n_lines / (np.mean(dms) * 1e-3),
                      "text_GBps": len(text) / (np.mean(dms) * 1e-3) / 1e9,
                      "algorithmic_GBps": (len(text) + 32 * n_scored + 48 * int(st[0]["n_filtered"])) / (np.mean(dms) * 1e-3) / 1e9},
           "e2e": {"ms": e2e_ms, "lines_per_s": n_lines / (e2e_ms * 1e-3), "h2d_bytes": len(text), "d2h_bytes": 32 * n_scored + 48 * int(st[0]["n_filtered"])}}
    # the reference's text loop on a sample
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    if os.path.exists(ref):
        import subprocess
        d = tempfile.mkdtemp(prefix="hc_ing_")
        k = min(a.cpu_lines, n_lines)
        cut = 0
        for _ in range(k):
            cut = text.index(b"\n", cut) + 1
        with open(d + "/ov.txt", "wb") as f:
            f.write(text[:cut])
        # a token pair of reads is enough: the loop under test never looks at the store
        with open(d + "/p1.fastq", "w") as f1, open(d + "/p2.fastq", "w") as f2:
            for r in range(2):
                f1.write("@%d\n%s\n+\n%s\n" % (r, "A" * 150, "I" * 150))
                f2.write("@%d\n%s\n+\n%s\n" % (r, "C" * 150, "I" * 150))
        o = subprocess.run([ref, "--overlaps", d + "/ov.txt", "--paired1", d + "/p1.fastq", "--paired2", d + "/p2.fastq", "--threads", "1",
                            "--min_overlap_len", "1000000", "--time-scoring", "--reps", "1"], cwd=d, stdout=subprocess.PIPE,
                           stderr=subprocess.PIPE, text=True)
        js = [json.loads(l) for l in o.stdout.split("\n") if l.startswith("{")]
        if js and "t_parse_s" in js[-1]:
            out["reference"] = {"lines": k, "t_parse_s": js[-1]["t_parse_s"], "lines_per_s": k / js[-1]["t_parse_s"], "threads": 1,
                                "note": "all lines pre-filtered (min_overlap_len 1e6) so that only the text loop runs"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
