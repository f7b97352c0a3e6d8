#!/usr/bin/env python
"""Config-5-like workload: contigs of 1-10 kb with Phred up to 93 (64 quality values -> the planar
store layout, or packed when --qmax gives <= 63 values), S-S candidates with long windows (the warp-cooperative path).

    python tools/bench_contigs.py [--contigs 2000] [--cands 2000000] [--qmax 93]
Prints one JSON line (device-resident inputs, CUDA-event kernel time from hc_batch_stats).  Benchmark tool."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from haploconduct_b200 import capi, formats as F, workloads as W  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--contigs", type=int, default=2000)
    ap.add_argument("--cands", type=int, default=2_000_000)
    ap.add_argument("--qmax", type=int, default=93)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    ss = W.synth_readset(a.contigs, 0, genome_len=100 * a.contigs, read_len=(1000, 10000), qmax=a.qmax, q_lo=30, seed=20261019, n_rate=0.0)
    base = W.geometry_candidates(ss, 50000, seed=6, min_ov=100, junk_fraction=0.0)
    cands = np.tile(base, (a.cands + len(base) - 1) // len(base))[: a.cands]
    params = F.make_params(edge_threshold=0.995, min_read_len=100)          # stage c
    peak = 6539.5
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    dev = torch.device("cuda:0")
    n = len(cands)
    d_cand = torch.from_numpy(cands.view(np.uint8).reshape(n, 32)).to(dev)
    d_edges = torch.empty((n, 48), dtype=torch.uint8, device=dev)
    d_nonedge = torch.empty(n, dtype=torch.int64, device=dev)
    d_counts = torch.zeros(4, dtype=torch.int64, device=dev)
    with capi.Store(ss.rs) as st:
        ms, stats = [], None
        for k in range(3 + a.steps):
            stats = st.score_batch_device(0, 0, params, d_cand.data_ptr(), n, 0, d_edges.data_ptr(), n, d_nonedge.data_ptr(), n,
                                          d_counts.data_ptr(), True)
            if k >= 3:
                ms.append(float(stats["score_kernel_ms"]))
        layout = "packed" if st.quality_alphabet <= 63 else "planar"
        codes = st.quality_alphabet
    kms = float(np.mean(ms))
    alg = float(stats["algorithmic_bytes"])
    print(json.dumps({"workload": "C5-like: %d contigs 1-10 kb, Phred 30..%d (%d values, %s layout), %d S-S candidates, mean window %.0f"
                      % (a.contigs, a.qmax, codes, layout, n, float(stats["n_positions"]) / max(float(stats["n_windows"]), 1)),
                      "candidates_per_s": n / (kms * 1e-3), "positions_per_s": float(stats["n_positions"]) / (kms * 1e-3), "kernel_ms": kms,
                      "roofline": {"bound": "hbm", "achieved": alg / (kms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                   "frac": alg / (kms * 1e-3) / 1e9 / peak},
                      "edges": int(d_counts[0]), "nonedges": int(d_counts[1])}))


if __name__ == "__main__":
    main()
