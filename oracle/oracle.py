"""ctypes binding of oracle/liboracle.so (the plain-C restatement, edgecalc_oracle.c) and a thin
runner for oracle/_ref/ref_driver (the UNMODIFIED reference C++).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import json
import os
import subprocess
import tempfile
from typing import Dict, Optional, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_DRIVER = os.path.join(HERE, "_ref", "ref_driver")
REFERENCE_ROOT = "/root/reference"

_lib = None


def build(force: bool = False, with_ref: bool = True) -> None:
    """make port (+ make ref when the reference tree is present; on the GPU box only the prebuilt
    oracle/_ref/ref_driver that travelled with the snapshot is used)."""
    targets = []
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(HERE, "edgecalc_oracle.c")):
        targets.append("port")
    if with_ref and os.path.isdir(os.path.join(REFERENCE_ROOT, "src")):
        targets.append("ref")
    if targets:
        subprocess.check_call(["make", "-s", "-C", HERE] + targets)


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build(with_ref=False)
        L = ctypes.CDLL(LIB_PATH)
        L.hco_score_batch.restype = ctypes.c_int
        L.hco_score_batch.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p]
        L.hco_overlap_score.restype = ctypes.c_double
        L.hco_overlap_score.argtypes = [ctypes.c_char_p, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_uint32, ctypes.c_char_p,
                                        ctypes.c_char_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.POINTER(ctypes.c_double)]
        L.hco_phred_to_prob.restype = ctypes.c_double
        L.hco_phred_to_prob.argtypes = [ctypes.c_int]
        L.hco_prefilter.restype = ctypes.c_int
        L.hco_prefilter.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int]
        _lib = L
    return _lib


def score_batch(rs, params: np.ndarray, cands: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """hco_score_batch: per-candidate RESULT records (+ per-window mean logs) in input order, 1 thread."""
    from haploconduct_b200.formats import RESULT

    n = len(cands)
    res = np.zeros(n, dtype=RESULT)
    means = np.zeros((n, 2), dtype=np.float64)
    cands = np.ascontiguousarray(cands)
    descs = np.ascontiguousarray(rs.descs)
    rc = lib().hco_score_batch(descs.ctypes.data, rs.n_reads, rs.n_single, rs.bases.ctypes.data, rs.quals.ctypes.data,
                               params.ctypes.data, cands.ctypes.data, n, res.ctypes.data, means.ctypes.data)
    if rc != 0:
        raise RuntimeError("oracle hco_score_batch failed with %d" % rc)
    return res, means


def overlap_score(seq1: str, seq2: str, q1: str, q2: str, pos: int, params: np.ndarray) -> Tuple[float, float]:
    mm = ctypes.c_double(0)
    s = lib().hco_overlap_score(seq1.encode(), len(seq1), seq2.encode(), len(seq2), q1.encode(), q2.encode(), pos,
                                params.ctypes.data, ctypes.byref(mm))
    return s, mm.value


def window_lengths(res: np.ndarray) -> np.ndarray:
    """sum of window lengths per candidate (the oracle parks it in the last word of its RESULT record -- the field the
    product reports as indel_count = 0)."""
    return res["indel_count"].astype(np.int64)


# ---- the compiled reference ------------------------------------------------------------------------
def have_ref() -> bool:
    return os.path.exists(REF_DRIVER) and os.access(REF_DRIVER, os.X_OK)


def run_ref(workdir: str, overlaps: str, singles: Optional[str] = None, paired1: Optional[str] = None,
            paired2: Optional[str] = None, threads: int = 1, dump_cands: bool = False, run: bool = False,
            dump_graph: bool = False, time_scoring: bool = False, reps: int = 1, dump_sorted: bool = False, **ps) -> Dict:
    """Runs oracle/_ref/ref_driver with cwd=workdir (the reference writes nonedge_overlaps.txt into
    its cwd, src/EdgeCalculator.cpp:566,549).  Returns the JSON summary plus parsed dumps."""
    cmd = [REF_DRIVER, "--overlaps", overlaps, "--threads", str(threads)]
    if singles:
        cmd += ["--singles", singles]
    if paired1:
        cmd += ["--paired1", paired1, "--paired2", paired2]
    for k, v in ps.items():
        cmd += ["--" + k, str(int(v) if isinstance(v, bool) else v)]
    if dump_cands:
        cmd += ["--dump-cands", os.path.join(workdir, "ref_cands.tsv")]
    if run:
        cmd += ["--run"]
    if dump_graph:
        cmd += ["--dump-graph", os.path.join(workdir, "ref_graph.tsv")]
    if dump_sorted:
        cmd += ["--dump-sorted", os.path.join(workdir, "ref_sorted.tsv")]
    if time_scoring:
        cmd += ["--time-scoring", "--reps", str(reps)]
    out = subprocess.run(cmd, cwd=workdir, check=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True).stdout
    jl = [json.loads(l) for l in out.split("\n") if l.startswith("{")]
    summary = jl[-1]
    for extra in jl[:-1]:
        summary.update(extra)
    if dump_cands:
        summary["cands"] = parse_cand_dump(os.path.join(workdir, "ref_cands.tsv"))
    if dump_graph:
        summary["graph"] = parse_graph_dump(os.path.join(workdir, "ref_graph.tsv"))
        summary["inclusions"] = parse_graph_inclusions(os.path.join(workdir, "ref_graph.tsv"))
    if dump_sorted:
        summary["sorted_graph"] = parse_graph_dump(os.path.join(workdir, "ref_sorted.tsv"))
        summary["adj_in"] = parse_adj_in(os.path.join(workdir, "ref_sorted.tsv"))
    if run and os.path.exists(os.path.join(workdir, "nonedge_overlaps.txt")):
        with open(os.path.join(workdir, "nonedge_overlaps.txt")) as f:
            summary["nonedge_lines"] = f.read().split("\n")[:-1]
    return summary


REF_CAND = np.dtype([("line", "<i8"), ("cls", "u1"), ("score", "<f8"), ("mismatch_rate", "<f8"), ("pos1", "<i4"),
                     ("pos2", "<i4"), ("pos3", "<i4"), ("pos4", "<i4"), ("v1", "<u8"), ("v2", "<u8"), ("ori1", "u1"),
                     ("ori2", "u1"), ("ord", "u1"), ("perc", "<i4"), ("len1", "<i4"), ("len2", "<i4")])
REF_EDGE = np.dtype([("v1", "<u8"), ("v2", "<u8"), ("score", "<f8"), ("mismatch_rate", "<f8"), ("pos1", "<i4"),
                     ("pos2", "<i4"), ("pos3", "<i4"), ("pos4", "<i4"), ("ori1", "u1"), ("ori2", "u1"), ("ord", "u1"),
                     ("perc", "<i4"), ("len1", "<i4"), ("len2", "<i4")])
_CLS = {"D": 0, "E": 1, "N": 2}


def parse_cand_dump(path: str) -> np.ndarray:
    rows = []
    with open(path) as f:
        for line in f:
            if line.startswith("#"):
                continue
            t = line.rstrip("\n").split("\t")
            rows.append((int(t[0]), _CLS[t[1]], float.fromhex(t[2]), float.fromhex(t[3]), int(t[4]), int(t[5]), int(t[6]),
                         int(t[7]), int(t[8]), int(t[9]), int(t[10]), int(t[11]), ord(t[12]) if t[12] else 0, int(t[13]),
                         int(t[14]), int(t[15])))
    return np.array(rows, dtype=REF_CAND)


def parse_graph_dump(path: str) -> np.ndarray:
    rows = []
    with open(path) as f:
        for line in f:
            if line.startswith("#"):
                continue
            t = line.rstrip("\n").split("\t")
            rows.append((int(t[0]), int(t[1]), float.fromhex(t[2]), float.fromhex(t[3]), int(t[4]), int(t[5]), int(t[6]),
                         int(t[7]), int(t[8]), int(t[9]), ord(t[10]) if t[10] else 0, int(t[11]), int(t[12]), int(t[13])))
    return np.array(rows, dtype=REF_EDGE)


def parse_adj_in(path: str):
    """The '#IN' lines of a --dump-sorted file as CSR-like arrays: (vertices with in-edges, offsets, sources)."""
    vs, off, src = [], [0], []
    with open(path) as f:
        for l in f:
            if l.startswith("#IN\t"):
                t = l.rstrip("\n").split("\t")
                vs.append(int(t[1]))
                src.extend(int(x) for x in t[2:])
                off.append(len(src))
    return np.array(vs, dtype=np.int64), np.array(off, dtype=np.int64), np.array(src, dtype=np.int64)


def sort_edges(graph: np.ndarray, read_len: np.ndarray):
    """OverlapGraph::sortEdges, src/OverlapGraph.cpp:722-764, restated on arrays: `graph` = REF_EDGE rows in adjacency order
    (vertex by vertex, list order), `read_len[v]` = Read::get_len() of vertex v (both mates of a pair, src/Read.h:203-212).
    Every list is ordered by (non-overlap length, vertex2) -- src/Edge.h:58-63: len1 + len2 - 2 * overlap_len in unsigned
    arithmetic.  std::sort is not stable; this restatement keeps the list order among equal keys, which is what std::sort
    does for lists of at most 16 edges (insertion sort) -- `ties_in_long_lists` counts the lists for which it may not hold.
    Returns (sorted rows, adj_in as (vertices, offsets, sources), ties_in_long_lists)."""
    ov = (graph["len1"].astype(np.int64) + graph["len2"].astype(np.int64))
    nol = ((read_len[graph["v1"].astype(np.int64)].astype(np.int64) + read_len[graph["v2"].astype(np.int64)].astype(np.int64) - 2 * ov)
           & 0xffffffff).astype(np.int64)
    order = np.lexsort((np.arange(len(graph)), graph["v2"], nol, graph["v1"]))
    out = graph[order]
    k = np.stack([out["v1"].astype(np.int64), nol[order], out["v2"].astype(np.int64)], axis=1)
    same = (k[1:] == k[:-1]).all(axis=1) if len(k) > 1 else np.zeros(0, bool)
    deg = np.bincount(out["v1"].astype(np.int64), minlength=int(len(read_len)))
    ties_long = int(np.unique(out["v1"][1:][same & (deg[out["v1"][1:].astype(np.int64)] > 16)]).size) if len(k) > 1 else 0
    # adj_in: for every vertex in order, for every edge of its sorted list: adj_in[v2].push_back(v1)  (:752-763)
    o2 = np.argsort(out["v2"].astype(np.int64), kind="stable")
    v2s = out["v2"].astype(np.int64)[o2]
    vs, first = np.unique(v2s, return_index=True)
    off = np.append(first, len(v2s)).astype(np.int64)
    return out, (vs.astype(np.int64), off, out["v1"].astype(np.int64)[o2]), ties_long


def calc_subread_info(trim_pos1, trim_pos2, pos1, vertices1, pos2, vertices2):
    """SRBuilder::calcSubreadInfo, src/SRBuilder.cpp:536-595, restated: {vertex: (index1, index2, startpos1, startpos2)}."""
    m = {}
    for p, v in zip(pos1, vertices1):
        if v in m:                                   # left index already there: a single-end super-read (:543-556)
            i1, _, s1, _ = m[v]
            m[v] = (i1, 0, s1, trim_pos1 - p) if trim_pos1 > p else (i1, p - trim_pos1, s1, 0)
        else:
            m[v] = (0, -1, trim_pos1 - p, -1) if trim_pos1 > p else (p - trim_pos1, -1, 0, -1)
    if trim_pos2 >= 0:                               # paired-end super-read: the /2 positions (:574-592)
        for p, v in zip(pos2, vertices2):
            i1, _, s1, _ = m[v]
            m[v] = (i1, 0, s1, trim_pos2 - p) if trim_pos2 > p else (i1, p - trim_pos2, s1, 0)
    return m


def parse_graph_inclusions(path: str) -> np.ndarray:
    """The '#I' lines of a --dump-graph file: vertices with OverlapGraph::inclusions set."""
    with open(path) as f:
        return np.array([int(l.split("\t")[1]) for l in f if l.startswith("#I\t")], dtype=np.int64)


# ---- FindNextOverlaps (FNO1) ---------------------------------------------------------------------------
def parse_fno_dump(path: str):
    """Reads the --merge-fno1 dump of ref_driver into a formats.FnoInput."""
    from haploconduct_b200 import formats as F

    verts, srs, edges = [], [], []
    ro = ni = 0
    with open(path) as f:
        for line in f:
            t = line.rstrip("\n").split("\t")
            if t[0] == "P":
                ro, ni = int(t[1]), int(t[2])
            elif t[0] == "S":
                srs.append((int(t[2]), int(t[3]), int(t[4])))
            elif t[0] == "V":
                subs = [tuple(int(x) for x in s.split(":")) for s in t[7:]]
                verts.append((int(t[2]), int(t[3]), int(t[4]), int(t[5]), int(t[6]), subs))
            elif t[0] == "E":
                edges.append((int(t[2]), int(t[3]), int(t[4]), int(t[5]), int(t[10]), int(t[11]), int(t[12]), ord(t[6]), int(t[7]),
                              int(t[8]), int(t[9])))
    V = len(verts)
    visited = np.array([v[0] for v in verts], dtype=np.uint8)
    label = np.array([v[2] for v in verts], dtype=np.uint8)
    vr = np.zeros(V, dtype=F.FNO_READ)
    vr["id"] = [max(v[1], 0) for v in verts]
    vr["len1"] = [v[3] for v in verts]
    vr["len2"] = [v[4] for v in verts]
    off = np.zeros(V + 1, dtype=np.uint64)
    idx, sub = [], []
    for i, v in enumerate(verts):
        for s in v[5]:
            idx.append(s[0])
            sub.append(s[1:])
        off[i + 1] = len(idx)
    sr = np.zeros(len(srs), dtype=F.FNO_READ)
    if srs:
        sr["id"], sr["len1"], sr["len2"] = zip(*srs)
    return F.FnoInput(visited=visited, label=label, vertex_read=vr, sr_off=off, sr_idx=np.array(idx, dtype=np.uint32),
                      sr_sub=np.array(sub, dtype=F.FNO_SUBREAD) if sub else np.zeros(0, dtype=F.FNO_SUBREAD), superread=sr,
                      resolve_orientations=ro, no_inclusions=ni, edges=np.array(edges, dtype=F.FNO_EDGE))


class _FnoInputC(ctypes.Structure):
    _fields_ = [("n_vertices", ctypes.c_uint64), ("visited", ctypes.c_void_p), ("label", ctypes.c_void_p),
                ("vertex_read", ctypes.c_void_p), ("sr_off", ctypes.c_void_p), ("sr_idx", ctypes.c_void_p),
                ("sr_sub", ctypes.c_void_p), ("n_superreads", ctypes.c_uint64), ("superread", ctypes.c_void_p),
                ("resolve_orientations", ctypes.c_uint8), ("no_inclusions", ctypes.c_uint8)]


def fno_input_struct(fi):
    """ctypes image of hc_fno_input; the numpy arrays of `fi` must stay alive while it is used."""
    keep = [np.ascontiguousarray(a) for a in (fi.visited, fi.label, fi.vertex_read, fi.sr_off, fi.sr_idx, fi.sr_sub, fi.superread)]
    st = _FnoInputC(len(fi.visited), keep[0].ctypes.data, keep[1].ctypes.data, keep[2].ctypes.data, keep[3].ctypes.data,
                    keep[4].ctypes.data, keep[5].ctypes.data, len(fi.superread), keep[6].ctypes.data, fi.resolve_orientations,
                    fi.no_inclusions)
    return st, keep


def fno1(fi) -> np.ndarray:
    """hco_fno1: the derived overlaps in processing order."""
    from haploconduct_b200 import formats as F

    L = lib()
    L.hco_fno1.restype = ctypes.c_int
    L.hco_fno1.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64,
                           ctypes.POINTER(ctypes.c_uint64)]
    st, keep = fno_input_struct(fi)
    edges = np.ascontiguousarray(fi.edges)
    cap = 1024
    while True:
        out = np.zeros(cap, dtype=F.FNO_OVERLAP)
        n = ctypes.c_uint64(0)
        rc = L.hco_fno1(ctypes.byref(st), edges.ctypes.data, len(edges), out.ctypes.data, cap, ctypes.byref(n))
        if rc == 0:
            return out[: n.value]
        if rc != -5:
            raise RuntimeError("hco_fno1 failed with %d" % rc)
        cap = int(n.value)


# ---- FindNextOverlaps3 ------------------------------------------------------------------------------------
def parse_fno3_dump(path: str):
    from haploconduct_b200 import formats as F

    srs, lists = [], []
    ni = 0
    with open(path) as f:
        for line in f:
            t = line.rstrip("\n").split("\t")
            if t[0] == "P":
                ni = int(t[1])
            elif t[0] == "S":
                srs.append((int(t[2]), int(t[3]), int(t[4])))
            elif t[0] == "O":
                lists.append([tuple(int(x) for x in s.split(":")) for s in t[2:]])
    off = np.zeros(len(lists) + 1, dtype=np.uint64)
    idx, pos = [], []
    for k, l in enumerate(lists):
        for s in l:
            idx.append(s[0])
            pos.append((s[1], s[2]))
        off[k + 1] = len(idx)
    reads = np.zeros(len(srs), dtype=F.FNO_READ)
    if srs:
        reads["id"], reads["len1"], reads["len2"] = zip(*srs)
    return F.Fno3Input(off=off, sr_idx=np.array(idx, dtype=np.uint32), sr_pos=np.array(pos, dtype=F.FNO3_POS), reads=reads,
                       no_inclusions=ni)


def fno3(fi) -> np.ndarray:
    from haploconduct_b200 import formats as F

    L = lib()
    L.hco_fno3.restype = ctypes.c_int
    L.hco_fno3.argtypes = [ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p,
                           ctypes.c_int, ctypes.c_void_p, ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint64)]
    off, idx, pos, reads = (np.ascontiguousarray(a) for a in (fi.off, fi.sr_idx, fi.sr_pos, fi.reads))
    cap = 1024
    while True:
        out = np.zeros(cap, dtype=F.FNO_OVERLAP)
        n = ctypes.c_uint64(0)
        rc = L.hco_fno3(len(off) - 1, off.ctypes.data, idx.ctypes.data, pos.ctypes.data, len(reads), reads.ctypes.data,
                        fi.no_inclusions, out.ctypes.data, cap, ctypes.byref(n))
        if rc == 0:
            return out[: n.value]
        if rc != -5:
            raise RuntimeError("hco_fno3 failed with %d" % rc)
        cap = int(n.value)


# ---- the serial graph insert (src/EdgeCalculator.cpp:429-545) ----------------------------------------------
def normalise_ref_edges(rc: np.ndarray) -> np.ndarray:
    """The accepted edges of a reference candidate dump (cls == 1) after the normalisation of :443-448 /
    Edge::swap_reads (src/Edge.h:74-88), as REF_EDGE rows in input order."""
    e = np.zeros(int((rc["cls"] == 1).sum()), dtype=REF_EDGE)
    src = rc[rc["cls"] == 1]
    for f in REF_EDGE.names:
        e[f] = src[f]
    sw = (e["pos1"] == 0) & (e["v1"] > e["v2"])
    v1, o1 = e["v1"].copy(), e["ori1"].copy()
    e["v1"][sw], e["v2"][sw] = e["v2"][sw], v1[sw]
    e["ori1"][sw], e["ori2"][sw] = e["ori2"][sw], o1[sw]
    od = e["ord"].copy()
    e["ord"][sw & (od == ord("1"))] = ord("2")
    e["ord"][sw & (od == ord("2"))] = ord("1")
    e["pos3"][sw] = -e["pos3"][sw]
    e["pos4"][sw] = -e["pos4"][sw]
    return e


def graph_insert(e: np.ndarray, n_vertices: int, ignore_inclusions: bool = False):
    """Sequential restatement of the serial section of EdgeCalculator::process_overlaps
    (src/EdgeCalculator.cpp:441-545) over normalised REF_EDGE rows: one pass, replace-if-not-worse.
    Returns (winner flags, inclusions, dup_count, inclusion_count, adjacency-ordered winner indices)."""
    have = {}                                  # (lo, hi, same_ori) -> index of the edge in the graph
    inclusions = np.zeros(n_vertices, dtype=np.uint8)
    dups = incl = 0
    for i in range(len(e)):
        x = e[i]
        v1, v2 = int(x["v1"]), int(x["v2"])
        if x["perc"] == 100:                                                   # :449-451
            incl += 1
        k = (min(v1, v2), max(v1, v2), bool(x["ori1"] == x["ori2"]))
        if k not in have:                                                      # :455-469
            have[k] = i
            if ignore_inclusions and x["perc"] == 100 and 0 <= x["mismatch_rate"] < 0.000001:
                if x["pos3"] < 0:
                    if x["pos1"] == 0:
                        inclusions[v1] = 1
                else:
                    inclusions[v2] = 1
            continue
        dups += 1                                                              # :472 / :537
        o = e[have[k]]
        if x["score"] < o["score"]:
            continue
        keep_old = False
        if x["score"] == o["score"]:                                           # :474-521, first difference decides
            ol, xl = int(o["len1"]) + int(o["len2"]), int(x["len1"]) + int(x["len2"])
            if ol != xl:
                keep_old = ol > xl
            elif o["mismatch_rate"] != x["mismatch_rate"]:
                keep_old = o["mismatch_rate"] < x["mismatch_rate"]
            elif o["v1"] != x["v1"]:
                keep_old = o["v1"] < x["v1"]
            elif o["ori1"] != x["ori1"]:
                keep_old = bool(o["ori1"])
            elif o["ori2"] != x["ori2"]:
                keep_old = bool(o["ori2"])
            elif o["pos1"] != x["pos1"]:
                keep_old = o["pos1"] < x["pos1"]
            elif o["pos2"] != x["pos2"]:
                keep_old = o["pos2"] < x["pos2"]
        if not keep_old:
            have[k] = i                                                        # :522-531 erase + append
    win = np.zeros(len(e), dtype=bool)
    win[list(have.values())] = True
    idx = np.nonzero(win)[0]
    adj = idx[np.argsort(e["v1"][idx], kind="stable")]       # adjacency dump order: by vertex1, then insertion time
    return win, inclusions, dups, incl, adj


def dedup_records(e: np.ndarray) -> np.ndarray:
    """REF_EDGE rows -> formats.DEDUP_EDGE (hc_dedup_edge) records."""
    from haploconduct_b200 import formats as F

    d = np.zeros(len(e), dtype=F.DEDUP_EDGE)
    d["vertex1"], d["vertex2"] = e["v1"], e["v2"]
    for f in ("score", "mismatch_rate", "pos1", "pos2", "pos3", "perc", "ori1", "ori2"):
        d[f] = e[f]
    d["overlap_len"] = e["len1"] + e["len2"]
    return d
