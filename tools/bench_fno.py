#!/usr/bin/env python
"""Throughput of FindNextOverlaps 1 on one B200 (hc_fno1, host buffers in and out) next to the pinned C restatement
of the reference's sequential walk (oracle/fno_oracle.c, one host core) on the same input.

    python tools/bench_fno.py [--edges 10000000] [--vertices 1000000] [--steps 5]
Prints one JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from haploconduct_b200 import capi, formats as F  # noqa: E402


def make_input(V: int, n_sr: int, n_edges: int, seed: int = 3, paired_fraction: float = 0.4) -> F.FnoInput:
    """The dense random input of tests/util.random_fno_input, vectorised for 1e6+ vertices."""
    rng = np.random.RandomState(seed)
    visited = (rng.random_sample(V) < 0.6).astype(np.uint8)
    label = rng.randint(0, 2, size=V).astype(np.uint8)
    vr = np.zeros(V, dtype=F.FNO_READ)
    vr["id"] = rng.permutation(V) + n_sr
    vr["len1"] = rng.randint(60, 300, size=V)
    vr["len2"] = np.where(rng.random_sample(V) < paired_fraction, rng.randint(60, 300, size=V), 0)
    sr = np.zeros(n_sr, dtype=F.FNO_READ)
    sr["id"] = np.arange(n_sr)
    sr["len1"] = rng.randint(100, 900, size=n_sr)
    sr["len2"] = np.where(rng.random_sample(n_sr) < paired_fraction, rng.randint(100, 900, size=n_sr), 0)
    k = np.where(visited > 0, rng.randint(0, 5, size=V), 0)
    off = np.zeros(V + 1, dtype=np.uint64)
    off[1:] = np.cumsum(k)
    n_ent = int(off[-1])
    owner = np.repeat(np.arange(V), k)
    j = np.arange(n_ent) - off[:-1][owner].astype(np.int64)
    start = rng.randint(0, n_sr, size=V)[owner]
    idx = ((start + j * 7919) % n_sr).astype(np.uint32)           # distinct super-reads within one vertex
    sub = np.zeros(n_ent, dtype=F.FNO_SUBREAD)
    i1, i2 = rng.randint(0, 400, size=n_ent), rng.randint(0, 400, size=n_ent)
    names = sub.dtype.names
    sub[names[0]], sub[names[1]] = i1, i2
    sub[names[2]] = np.where((i1 > 0) & (rng.random_sample(n_ent) < 0.8), 0, rng.randint(0, 30, size=n_ent))
    sub[names[3]] = np.where((i2 > 0) & (rng.random_sample(n_ent) < 0.8), 0, rng.randint(0, 30, size=n_ent))
    e = np.zeros(n_edges, dtype=F.FNO_EDGE)
    e["u"] = rng.randint(0, V, size=n_edges)
    e["v"] = (e["u"] + rng.randint(1, V, size=n_edges)) % V
    e["pos1"], e["pos2"] = rng.randint(0, 250, size=n_edges), rng.randint(0, 250, size=n_edges)
    e["perc"] = rng.randint(20, 101, size=n_edges)
    e["len1"], e["len2"] = rng.randint(30, 250, size=n_edges), rng.randint(0, 250, size=n_edges)
    pu, pv = vr["len2"][e["u"]] > 0, vr["len2"][e["v"]] > 0
    e["ord"] = np.where(pu & pv, np.where(rng.random_sample(n_edges) < 0.5, ord("1"), ord("2")), ord("-"))
    e["ori1"], e["ori2"] = rng.randint(0, 2, size=n_edges), rng.randint(0, 2, size=n_edges)
    e["nonedge"] = rng.random_sample(n_edges) < 0.5
    return F.FnoInput(visited=visited, label=label, vertex_read=vr, sr_off=off, sr_idx=idx, sr_sub=sub, superread=sr,
                      resolve_orientations=1, no_inclusions=0, edges=e)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--edges", type=int, default=10_000_000)
    ap.add_argument("--vertices", type=int, default=1_000_000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--cpu-edges", type=int, default=2_000_000)
    a = ap.parse_args()
    fi = make_input(a.vertices, a.vertices // 3, a.edges)
    import ctypes
    out = capi.fno1(fi)                                            # warm-up (also sizes the output)
    L = capi.lib()
    keep = [np.ascontiguousarray(x) for x in (fi.visited, fi.label, fi.vertex_read, fi.sr_off, fi.sr_idx, fi.sr_sub, fi.superread)]
    st = capi._FnoInputC(len(fi.visited), *[k.ctypes.data for k in keep[:6]], len(fi.superread), keep[6].ctypes.data,
                         fi.resolve_orientations, fi.no_inclusions)
    edges = np.ascontiguousarray(fi.edges)
    buf = np.zeros(len(out) + 16, dtype=F.FNO_OVERLAP)              # touched once: no page faults inside the timed calls
    bufs = np.zeros(len(out) + 16, dtype=capi.FNO_OVERLAP_SMALL)
    n = ctypes.c_uint64(0)
    t, ts = [], []
    for _ in range(a.steps):
        t0 = time.perf_counter()
        rc = L.hc_fno1(ctypes.byref(st), edges.ctypes.data, len(edges), buf.ctypes.data, len(buf), ctypes.byref(n), 0)
        t.append(time.perf_counter() - t0)
        assert rc == 0 and n.value == len(out)
        t0 = time.perf_counter()
        rc = L.hc_fno1_small(ctypes.byref(st), edges.ctypes.data, len(edges), bufs.ctypes.data, len(bufs), ctypes.byref(n), 0)
        ts.append(time.perf_counter() - t0)
        assert rc == 0 and n.value == len(out)
    assert buf[: len(out)].tobytes() == out.tobytes()
    assert capi.fno_small_to_overlaps(bufs[: len(out)]).tobytes() == out.tobytes()
    best = min(ts)
    res = {"metric": "FNO1 edges processed per second (host buffers, incl. all copies and allocations)", "unit": "edges/s",
           "edges": a.edges, "vertices": a.vertices, "overlaps_out": int(len(out)), "ms": best * 1e3, "value": a.edges / best,
           "records": "hc_fno1_small: 24-byte result records (the call of the C++ binding hcb::SRBuilder)",
           "ms_48_byte_records": min(t) * 1e3,
           "bytes_in": int(fi.edges.nbytes + fi.sr_sub.nbytes + fi.sr_idx.nbytes + fi.vertex_read.nbytes + fi.superread.nbytes),
           "bytes_out": int(len(out) * 24), "bytes_out_48": int(out.nbytes)}
    try:
        from oracle import oracle as O
        import copy
        small = copy.copy(fi)
        small.edges = fi.edges[: a.cpu_edges]
        t0 = time.perf_counter()
        ref = O.fno1(small)
        dt = time.perf_counter() - t0
        mine = capi.fno1(small)
        res["cpu_baseline"] = {"kind": "port", "cores": 1, "edges": int(len(small.edges)), "value": len(small.edges) / dt,
                               "identical_output": bool(mine.tobytes() == ref.tobytes())}
    except Exception as ex:      # the oracle is test infrastructure; its absence only removes the baseline
        res["cpu_baseline"] = {"unavailable": str(ex)[:200]}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
