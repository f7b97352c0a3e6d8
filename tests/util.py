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
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))


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
