#!/usr/bin/env python
"""Files in, overlap graph out: EdgeCalculator::construct_edges() of the UNMODIFIED reference (oracle/_ref/ref_driver
--run, all host threads) next to the host mirror on the B200 (haploconduct_b200/lib/hc_edgecalc), once with the host
parsers and once with FASTQ reading, overlaps-file parsing and duplicate resolution on the device.  The three graphs
must be identical (adjacency dump compared byte for byte).

    python tools/bench_pipeline.py [--pairs 100000] [--partners 20]
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

import torch  # noqa: E402

from haploconduct_b200 import build as B, formats as F, workloads_torch as WT  # noqa: E402


def write_inputs(d, pairs, read_len, partners, seed):
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    pr = WT.make_paired_reads(pairs, read_len=read_len, genome_len=max(20000, pairs // 10), seed=seed, device=dev)
    rec = WT.candidates_as_numpy(WT.make_pp_candidates(pr, D=partners))
    L = read_len
    b = pr.bases.numpy().reshape(pairs, 2, L)
    q = pr.quals.numpy().reshape(pairs, 2, L)
    # FASTQ records "@<id>\n<seq>\n+\n<qual>\n": ids with the same number of digits give fixed-length records, written as
    # one 2-D byte array per digit count (a Python loop over 3e6 pairs takes minutes)
    for m, name in ((0, "p1.fastq"), (1, "p2.fastq")):
        with open(os.path.join(d, name), "wb") as f:
            lo = 0
            while lo < pairs:
                nd = len(str(lo))
                hi = min(pairs, 10 ** nd)
                ids = np.arange(lo, hi)
                rec_len = 1 + nd + 1 + L + 3 + L + 1
                out = np.empty((hi - lo, rec_len), dtype=np.uint8)
                out[:, 0] = ord("@")
                for k in range(nd):
                    out[:, 1 + k] = (ids // 10 ** (nd - 1 - k)) % 10 + ord("0")
                out[:, 1 + nd] = 10
                out[:, 2 + nd:2 + nd + L] = b[lo:hi, m]
                out[:, 2 + nd + L] = 10; out[:, 3 + nd + L] = ord("+"); out[:, 4 + nd + L] = 10
                out[:, 5 + nd + L:5 + nd + 2 * L] = q[lo:hi, m]
                out[:, 5 + nd + 2 * L] = 10
                f.write(out.tobytes())
                lo = hi
    import pyarrow as pa
    import pyarrow.csv as pacsv
    n = len(rec)
    plus, pch = pa.array(["+"] * 1).take(pa.array(np.zeros(n, dtype=np.int32))), None
    ordc = pa.array(np.array(["-", "1", "2"])).take(pa.array(np.where(rec["ord"] == ord("1"), 1, np.where(rec["ord"] == ord("2"), 2, 0)).astype(np.int32)))
    pcol = pa.array(["p"]).take(pa.array(np.zeros(n, dtype=np.int32)))
    tab = pa.table({"a": rec["idx1"], "b": rec["idx2"], "c": rec["pos1"], "d": rec["pos2"], "e": ordc, "f": plus, "g": plus,
                    "h": rec["perc1"].astype(np.int32), "i": rec["perc2"].astype(np.int32), "j": rec["len1"], "k": rec["len2"], "l": pcol, "m": pcol})
    pacsv.write_csv(tab, os.path.join(d, "ov.txt"), pacsv.WriteOptions(include_header=False, delimiter="\t", quoting_style="none"))
    return len(rec)


def run(cmd, cwd):
    t0 = time.perf_counter()
    pr = subprocess.run(cmd, cwd=cwd, check=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    out = pr.stdout
    wall = time.perf_counter() - t0
    if os.environ.get("HC_MIRROR_TIMING"):
        sys.stderr.write("".join(l + "\n" for l in pr.stderr.split("\n") if l.startswith("[")))
    js = [json.loads(l) for l in out.split("\n") if l.startswith("{")]
    return wall, (js[-1] if js else {})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=100_000)
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--partners", type=int, default=20)
    ap.add_argument("--seed", type=int, default=7)
    ap.add_argument("--skip-host-parsers", action="store_true", help="only the device-ingest mirror (the host-parser mirror takes as long as the reference)")
    ap.add_argument("--one-thread-limit", type=int, default=6_000_000, help="also run the reference on 1 thread up to this many candidates")
    a = ap.parse_args()
    d = tempfile.mkdtemp(prefix="hc_pipe_")
    n_cand = write_inputs(d, a.pairs, a.read_len, a.partners, a.seed)
    common = ["--overlaps", "ov.txt", "--paired1", "p1.fastq", "--paired2", "p2.fastq", "--edge_threshold", "0.97", "--min_overlap_len", "150"]
    res = {"metric": "construct_edges wall time, files in -> overlap graph out", "pairs": a.pairs, "candidates": n_cand,
           "overlaps_file_bytes": os.path.getsize(d + "/ov.txt")}
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    threads = os.cpu_count() or 1
    if os.path.exists(ref):
        wall, js = run([ref] + common + ["--threads", str(threads), "--run", "--dump-graph", "ref.tsv"], d)
        res["reference"] = {"wall_s": wall, "t_fastq_s": js.get("t_fastq_s"), "t_construct_edges_s": js.get("t_construct_edges_s"),
                            "threads": threads, "graph_edges": js.get("graph_edges")}
        if n_cand <= a.one_thread_limit:       # the canonical order (1 thread) for the exact comparison
            wall1, js1 = run([ref] + common + ["--threads", "1", "--run", "--dump-graph", "ref1.tsv"], d)
            res["reference_1_thread"] = {"wall_s": wall1, "t_construct_edges_s": js1.get("t_construct_edges_s")}
    exe = os.path.join(B.LIBDIR, "hc_edgecalc")
    for name, flags in (("mirror_host_parsers", []), ("mirror_device_ingest", ["--gpu_fastq=true", "--gpu_parse=true", "--gpu_dedup=true"])):
        if a.skip_host_parsers and name == "mirror_host_parsers":
            continue
        wall, js = run([exe] + common + flags + ["--dump-graph", name + ".tsv"], d)
        res[name] = {"wall_s": wall, **{k: js.get(k) for k in ("t_fastq_s", "t_fastq_read_s", "t_cuda_init_s", "t_fastq_store_s", "t_fastq_index_s", "t_construct_edges_s", "graph_edges", "device_ms", "parse_device_ms", "t_ingest_s", "t_score_s", "t_edges_s", "t_write_s", "t_main_s", "t_graph_files_s") if k in js}}
        if os.path.exists(d + "/ref.tsv"):
            from oracle import oracle as O      # only its dump parser: every Edge field of every adjacency list, in order
            g_ref, g_own = O.parse_graph_dump(d + "/ref.tsv"), O.parse_graph_dump(d + "/" + name + ".tsv")
            # the reference ran with all host threads: its edge vector, hence the adjacency order, follows the thread
            # schedule; compare the edges as a set (all fields), and in order against a 1-thread run if there is one
            def canon(g):
                return np.sort(g, order=["v1", "v2", "ori1", "ori2", "pos1", "pos2"])
            res[name]["same_edges_as_reference"] = bool(len(g_ref) == len(g_own) and canon(g_ref).tobytes() == canon(g_own).tobytes())
            if os.path.exists(d + "/ref1.tsv"):
                g1 = O.parse_graph_dump(d + "/ref1.tsv")
                res[name]["identical_to_1_thread_reference"] = bool(len(g1) == len(g_own) and g1.tobytes() == g_own.tobytes())
    # where the mirror's time goes, next to what each part would take at the machine's rates (a roofline for the host side)
    m = res.get("mirror_device_ingest", {})
    if m and "reference" in res:
        ov, fq = res["overlaps_file_bytes"], sum(os.path.getsize(d + "/" + f) for f in ("p1.fastq", "p2.fastq"))
        res["breakdown"] = {
            "speedup_construct_edges_vs_reference": res["reference"]["t_construct_edges_s"] / max(m.get("t_construct_edges_s", 0), 1e-9),
            "speedup_wall_vs_reference": res["reference"]["wall_s"] / max(m["wall_s"], 1e-9),
            "speedup_main_vs_reference_fastq_plus_construct_edges": (res["reference"]["t_fastq_s"] + res["reference"]["t_construct_edges_s"]) / max(m.get("t_main_s") or 0, 1e-9),
            "speedup_wall_without_cuda_context_creation": res["reference"]["wall_s"] / max(m["wall_s"] - (m.get("t_cuda_init_s") or 0.0), 1e-9),
            "phases_s": {k: m.get(k) for k in ("t_fastq_s", "t_fastq_read_s", "t_cuda_init_s", "t_fastq_store_s", "t_fastq_index_s", "t_ingest_s", "t_score_s", "t_edges_s", "t_write_s")},
            "device_busy_ms": {"parse": m.get("parse_device_ms"), "score": m.get("device_ms")},
            "bytes": {"overlaps_file": ov, "fastq_files": fq},
            "floors_s": {"read both inputs once at 3 GB/s (page cache)": (ov + fq) / 3e9, "copy them to the device at 25 GB/s": (ov + fq) / 25e9},
            "note": "t_fastq_s includes the creation of the CUDA context (a few tenths of a second per process)",
        }
    print(json.dumps(res))


if __name__ == "__main__":
    main()
