"""Shared helpers of the test-suite."""
import glob
import os

import numpy as np

from haploconduct_b200 import formats as F

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Golden:
    def __init__(self, path):
        z = np.load(path, allow_pickle=False)
        self.name = os.path.basename(path)[:-4]
        self.rs = F.ReadSet(ids=z["ids"], descs=z["descs"], bases=z["bases"], quals=z["quals"], n_single=int(z["n_single"]))
        self.cands = z["cands"]
        self.ps = {str(k): float(v) for k, v in zip(z["ps_keys"], z["ps_vals"])}
        self.ref_cands = z["ref_cands"]
        self.ref_graph = z["ref_graph"]
        self.ref_nonedge = [str(x) for x in z["ref_nonedge"]]
        self.ref_counts = z["ref_counts"]

    def params(self, flags=0):
        return F.make_params(edge_threshold=self.ps.get("edge_threshold", 0.99), ov_threshold=self.ps.get("ov_threshold", 0.9),
                             merge_contigs=self.ps.get("merge_contigs", 0.0), mismatch=self.ps.get("mismatch", 0.0),
                             min_read_len=int(self.ps.get("min_read_len", 0)), flags=flags)

    def prefilter(self):
        return F.prefilter(self.cands, int(self.ps.get("min_overlap_len", 150)), int(self.ps.get("min_overlap_perc", 0)),
                           bool(self.ps.get("relax_PE_edges", 0)))

    def scored(self):
        return self.cands[self.prefilter() == 1]


def golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")) if not os.path.basename(p).startswith(("fno", "insert_", "ingest_", "consensus_", "sorted_", "sfo_", "sam_", "subread_")))


def fno_golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "fno1_*.npz")))


def load_fno_golden(name):
    """(FnoInput, reference overlaps.txt lines) -- inputs and output of the reference's own findNextOverlaps()."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    fi = F.FnoInput(visited=z["visited"], label=z["label"], vertex_read=z["vertex_read"], sr_off=z["sr_off"], sr_idx=z["sr_idx"],
                    sr_sub=z["sr_sub"], superread=z["superread"], resolve_orientations=int(z["flags"][0]),
                    no_inclusions=int(z["flags"][1]), edges=z["edges"])
    return fi, [str(x) for x in z["ref_lines"]]


def fno3_golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "fno3_*.npz")))


def load_fno3_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    fi = F.Fno3Input(off=z["off"], sr_idx=z["sr_idx"], sr_pos=z["sr_pos"], reads=z["reads"], no_inclusions=int(z["flags"][0]))
    return fi, [str(x) for x in z["ref_lines"]]


def random_fno3_input(seed, n_originals=2000, n_reads=600, paired_fraction=0.4):
    rng = np.random.RandomState(seed)
    reads = np.zeros(n_reads, dtype=F.FNO_READ)
    reads["id"] = rng.permutation(n_reads)
    reads["len1"] = rng.randint(80, 600, size=n_reads)
    reads["len2"] = np.where(rng.random_sample(n_reads) < paired_fraction, rng.randint(80, 600, size=n_reads), 0)
    off = np.zeros(n_originals + 1, dtype=np.uint64)
    idx, pos = [], []
    for k in range(n_originals):
        c = int(rng.choice([1, 1, 2, 2, 3, 4, 6, 9]))
        for s in rng.choice(n_reads, size=c, replace=False):
            idx.append(int(s))
            pos.append((int(rng.randint(-20, 500)), int(rng.randint(-20, 500))))
        off[k + 1] = len(idx)
    return F.Fno3Input(off=off, sr_idx=np.array(idx, dtype=np.uint32), sr_pos=np.array(pos, dtype=F.FNO3_POS), reads=reads,
                       no_inclusions=int(seed % 2))


def random_fno_input(seed, n_vertices=300, n_sr=120, n_edges=4000, paired_fraction=0.4):
    """Random but self-consistent FNO1 input: many super-reads per vertex, so that the first-found-wins
    rule, failing derivations and all four read-type combinations are exercised far more densely
    than real merge iterations do."""
    rng = np.random.RandomState(seed)
    V = n_vertices
    visited = (rng.random_sample(V) < 0.6).astype(np.uint8)
    label = rng.randint(0, 2, size=V).astype(np.uint8)
    vr = np.zeros(V, dtype=F.FNO_READ)
    vr["id"] = rng.permutation(V) + n_sr          # new ids of unmerged reads follow the super-read ids
    vr["len1"] = rng.randint(60, 300, size=V)
    vr["len2"] = np.where(rng.random_sample(V) < paired_fraction, rng.randint(60, 300, size=V), 0)
    sr = np.zeros(n_sr, dtype=F.FNO_READ)
    sr["id"] = np.arange(n_sr)
    sr["len1"] = rng.randint(100, 900, size=n_sr)
    sr["len2"] = np.where(rng.random_sample(n_sr) < paired_fraction, rng.randint(100, 900, size=n_sr), 0)
    off = np.zeros(V + 1, dtype=np.uint64)
    idx, sub = [], []
    for v in range(V):
        if visited[v]:
            k = int(rng.randint(0, 5))
            for s in rng.choice(n_sr, size=k, replace=False):
                idx.append(int(s))
                i1, i2 = int(rng.randint(0, 400)), int(rng.randint(0, 400))
                s1 = 0 if i1 > 0 and rng.random_sample() < 0.8 else int(rng.randint(0, 30))
                s2 = 0 if i2 > 0 and rng.random_sample() < 0.8 else int(rng.randint(0, 30))
                sub.append((i1, i2, s1, s2))
        off[v + 1] = len(idx)
    e = np.zeros(n_edges, dtype=F.FNO_EDGE)
    e["u"] = rng.randint(0, V, size=n_edges)
    e["v"] = (e["u"] + rng.randint(1, V, size=n_edges)) % V
    e["pos1"] = rng.randint(0, 250, size=n_edges)
    e["pos2"] = rng.randint(0, 250, size=n_edges)
    e["perc"] = rng.randint(20, 101, size=n_edges)
    e["len1"] = rng.randint(30, 250, size=n_edges)
    e["len2"] = rng.randint(0, 250, size=n_edges)
    pu, pv = vr["len2"][e["u"]] > 0, vr["len2"][e["v"]] > 0
    e["ord"] = np.where(pu & pv, np.where(rng.random_sample(n_edges) < 0.5, ord("1"), ord("2")), ord("-"))
    e["ori1"] = rng.randint(0, 2, size=n_edges)
    e["ori2"] = rng.randint(0, 2, size=n_edges)
    e["nonedge"] = rng.random_sample(n_edges) < 0.5
    return F.FnoInput(visited=visited, label=label, vertex_read=vr, sr_off=off, sr_idx=np.array(idx, dtype=np.uint32),
                      sr_sub=np.array(sub, dtype=F.FNO_SUBREAD) if sub else np.zeros(0, dtype=F.FNO_SUBREAD), superread=sr,
                      resolve_orientations=1, no_inclusions=int(seed % 2), edges=e)


def load_golden(name):
    return Golden(os.path.join(GOLDEN, name + ".npz"))


def assert_results_match(res, ref_score, ref_mm, ref_pos3, ref_pos4, ref_cls, rel=1e-6, what=""):
    """The parity bar: integers and classes bit-exact, scores within 1e-6 relative."""
    assert np.array_equal(res["cls"], ref_cls), "%s: class mismatch at %s" % (what, np.nonzero(res["cls"] != ref_cls)[0][:10])
    assert np.array_equal(res["pos3"], ref_pos3), what + ": pos3"
    assert np.array_equal(res["pos4"], ref_pos4), what + ": pos4"
    assert np.array_equal(res["mismatch_rate"], ref_mm), "%s: mismatch_rate differs at %s" % (
        what, np.nonzero(res["mismatch_rate"] != ref_mm)[0][:10])
    err = np.abs(res["score"] - ref_score)
    tol = rel * np.abs(ref_score)
    bad = np.nonzero(err > tol)[0]
    assert len(bad) == 0, "%s: %d scores off by more than %g relative, first %s: %s vs %s" % (
        what, len(bad), rel, bad[:5], res["score"][bad[:5]], ref_score[bad[:5]])


def load_insert_golden(name):
    """(n_vertices, vertices the reference marked in OverlapGraph::inclusions under --ignore_inclusions)."""
    z = np.load(os.path.join(GOLDEN, "insert_" + name + ".npz"), allow_pickle=False)
    return int(z["n_vertices"]), z["ref_inclusions"]


def random_insert_edges(seed, n, n_vertices):
    """Dense duplicates with many exact ties on every level of the tie-break chain (oracle.REF_EDGE rows)."""
    from oracle import oracle as O

    rng = np.random.default_rng(seed)
    e = np.zeros(n, dtype=O.REF_EDGE)
    e["v1"] = rng.integers(0, n_vertices, n)
    e["v2"] = (e["v1"] + 1 + rng.integers(0, min(4, n_vertices - 1), n)) % n_vertices
    sw = rng.random(n) < 0.5
    e["v1"][sw], e["v2"][sw] = e["v2"][sw].copy(), e["v1"][sw].copy()
    e["score"] = rng.choice([0.97, 0.98, 0.99, 0.99 + 2.0 ** -40], n)
    e["mismatch_rate"] = rng.choice([0.0, 0.0, 1e-7, 0.01], n)
    e["pos1"] = rng.choice([0, 0, 5, 9], n)
    e["pos2"] = rng.choice([0, 3, 7], n)
    e["pos3"] = rng.choice([-4, 0, 6], n)
    e["pos4"] = rng.choice([-2, 0, 2], n)
    e["ori1"] = rng.integers(0, 2, n)
    e["ori2"] = rng.integers(0, 2, n)
    e["ord"] = rng.choice([ord("-"), ord("1"), ord("2")], n)
    e["perc"] = rng.choice([60, 100, 100], n)
    e["len1"] = rng.choice([100, 120], n)
    e["len2"] = rng.choice([0, 20], n)
    # the normalisation of :443-448 on the raw records
    rc = np.zeros(n, dtype=O.REF_CAND)
    for f in O.REF_EDGE.names:
        rc[f] = e[f]
    rc["cls"] = 1
    return O.normalise_ref_edges(rc)


class IngestGolden:
    """tests/golden/ingest_*.npz: an irregularly spelled overlaps file and what the reference's construct_edges()
    printed back for it (oracle/make_golden.py: run_ingest_case)."""

    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
        self.name = name
        self.ids = z["ids"]
        self.text = z["text"].tobytes()
        self.ps = {str(k): float(v) for k, v in zip(z["ps_keys"], z["ps_vals"])}
        self.allow_spaces = bool(z["allow_spaces"])
        self.ref_scored = [str(x) for x in z["ref_scored"]]
        self.ref_filtered = [str(x) for x in z["ref_filtered"]]
        self.ref_counts = z["ref_counts"]     # lines, scored, filtered, self overlaps, perc-dropped, skipped

    def kw(self):
        return dict(min_overlap_len=int(self.ps["min_overlap_len"]), min_overlap_perc=int(self.ps.get("min_overlap_perc", 0)),
                    relax_PE_edges=bool(self.ps.get("relax_PE_edges", 0)), allow_spaces=self.allow_spaces)


def ingest_golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "ingest_*.npz")))


class ConsensusGolden:
    """tests/golden/consensus_*.npz: pile-ups and what the reference's SRBuilder::consensus returned for them."""

    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
        self.rs = F.ReadSet(ids=z["ids"], descs=z["descs"], bases=z["bases"], quals=z["quals"], n_single=int(z["n_single"]))
        self.P, self.S = z["problems"], z["seqs"]
        self.min_clique_size, self.min_qual = int(z["params"][0]), float(z["params"][1])
        self.ref = [(int(r), str(s), str(q)) for r, s, q in zip(z["ref_ret"], z["ref_seq"], z["ref_qual"])]

    def problems(self):
        out = []
        for p in self.P:
            ent = [(int(e["read"]), int(e["mate"]), bool(e["rc"]), int(e["pos"])) for e in self.S[int(p["seq_begin"]):int(p["seq_end"])]]
            out.append(dict(total_len=int(p["total_len"]), subreads_needed=bool(p["subreads_needed"]),
                            error_correction=bool(p["error_correction"]), entries=ent))
        return out


def consensus_golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "consensus_*.npz")))


def consensus_oracle_results(rs, problems, min_clique_size, min_qual):
    from haploconduct_b200.workloads import revcomp
    from oracle import consensus_oracle as CO
    out = []
    for p in problems:
        pos, seqs, quals = [], [], []
        for read, mate, rc, ps in p["entries"]:
            s, q = rs.seq(read, mate), rs.qual(read, mate)
            if rc:
                s, q = revcomp(s), q[::-1]
            pos.append(ps); seqs.append(s); quals.append(q)
        out.append(CO.consensus(p["total_len"], pos, seqs, quals, p["subreads_needed"], p["error_correction"], min_clique_size, min_qual))
    return out
