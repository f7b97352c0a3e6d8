"""GPU suite, larger than what the oracle can check exhaustively: size-independent properties on a
config-4 shaped workload (2x150 bp pairs, P-P candidates, shuffled read ids) plus an oracle check on
a random sample of it."""
import numpy as np
import pytest
import torch

from haploconduct_b200 import capi, formats as F, workloads_torch as WT
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=[150, 250], ids=["c4_2x150", "c3_2x250"])
def c4_small(request):
    """config-4 shaped (2x150) and config-3 shaped (2x250) reads; candidates need half of min_overlap_len per mate"""
    L = request.param
    pr = WT.make_paired_reads(200_000 if L == 150 else 120_000, read_len=L, genome_len=2_000, seed=20261018 if L == 150 else 20261017,
                              insert=(450.0, 50.0) if L == 150 else (600.0, 60.0), device="cuda")
    rec = WT.make_pp_candidates(pr, D=40, shard=0, n_shards=1, max_cands=5_000_000)
    cands = WT.candidates_as_numpy(rec)
    return pr.readset(), cands


def test_shard_invariance_and_order(built_lib, c4_small):
    """Scoring the list in one call equals scoring 8 contiguous shards and concatenating in rank
    order (what N one-GPU ranks + the ordered gather produce): same edges, same order."""
    rs, cands = c4_small
    p = F.make_params(edge_threshold=0.97)
    n = len(cands)
    assert n >= 4_000_000
    with capi.Store(rs) as st:
        e_all, n_all, _, s_all = st.score_batch(p, cands, per_candidate=False)
        parts_e, parts_n = [], []
        for k in range(8):
            lo, hi = n * k // 8, n * (k + 1) // 8
            e, ne, _, _ = st.score_batch(p, cands[lo:hi], per_candidate=False, compact=(k % 2 == 1))
            e = e.copy()
            e["cand"] += np.uint64(lo)
            parts_e.append(e)
            parts_n.append(ne + np.uint64(lo))
    assert np.concatenate(parts_e).tobytes() == e_all.tobytes()
    assert np.array_equal(np.concatenate(parts_n), n_all)
    assert np.all(np.diff(e_all["cand"].astype(np.int64)) > 0) and np.all(np.diff(n_all.astype(np.int64)) > 0)   # input order
    assert 0 < len(e_all) < n and 0 < len(n_all) < n
    # the kernel's own bookkeeping of algorithmic bytes: 32 + sum_w(2*ceil(L/4) + 2*ceil(L/8) + 2L) + 16 per candidate
    RL = int(rs.descs["seq_len"][0, 0])       # every mate has the same length in this workload
    L1 = RL - cands["pos1"].astype(np.int64)
    L2 = RL - cands["pos2"].astype(np.int64)
    want = (48 * n + sum(int((2 * ((L + 3) // 4) + 2 * ((L + 7) // 8) + 2 * L).sum()) for L in (L1, L2)))
    assert int(s_all["algorithmic_bytes"]) == want and int(s_all["n_positions"]) == int((L1 + L2).sum())


def test_random_sample_against_oracle(built_lib, c4_small):
    rs, cands = c4_small
    p = F.make_params(edge_threshold=0.97)
    rng = np.random.RandomState(1)
    pick = np.sort(rng.choice(len(cands), size=40_000, replace=False))
    with capi.Store(rs) as st:
        e_all, n_all, _, _ = st.score_batch(p, cands, per_candidate=False)
        _, _, per, _ = st.score_batch(p, cands[pick])
    ref, _ = O.score_batch(rs, p, cands[pick])
    assert np.array_equal(per["cls"], ref["cls"])
    assert np.array_equal(per["mismatch_rate"], ref["mismatch_rate"]) and np.array_equal(per["mismatches"], ref["mismatches"])
    assert np.allclose(per["score"], ref["score"], rtol=1e-6, atol=0)
    # membership of the sample in the full run's lists agrees with the oracle's classes
    is_e = np.isin(pick.astype(np.uint64), e_all["cand"])
    is_n = np.isin(pick.astype(np.uint64), n_all)
    assert np.array_equal(is_e, ref["cls"] == 1) and np.array_equal(is_n, ref["cls"] == 2)


def test_permutation_equivariance(built_lib, c4_small):
    """Integer sums: a candidate's result does not depend on where it sits in the batch."""
    rs, cands = c4_small
    sub = cands[:300_000]
    perm = np.random.RandomState(2).permutation(len(sub))
    p = F.make_params(edge_threshold=0.97)
    with capi.Store(rs) as st:
        _, _, a, _ = st.score_batch(p, sub)
        _, _, b, _ = st.score_batch(p, sub[perm])
    assert a[perm].tobytes() == b.tobytes()


def test_store_replicated_on_two_devices(built_lib, monkeypatch):
    """n_devices = 2 in one process: the planes packed on the first device are copied to the second, the batch is cut into
    two contiguous shards and the lists come back in input order -- identical to the one-device result."""
    if capi.device_count() < 2:
        pytest.skip("needs two GPUs")
    from util import load_golden
    g = load_golden("synth_all_types")
    cands = np.tile(g.scored(), 5)
    with capi.Store(g.rs) as st1:
        e1, n1, p1, _ = st1.score_batch(g.params(), cands)
    with capi.Store(g.rs, first_device=0, n_devices=2) as st2:
        e2, n2, p2, _ = st2.score_batch(g.params(), cands)
        fits = (cands["pos1"] < (1 << 14)) & (cands["pos2"] < (1 << 14))
        er, nr, pr, _ = st2.score_batch(g.params(), cands[fits], compact="runs")     # run-encoded records, two shards
        monkeypatch.setenv("HC_HOST_CHUNK", "1000")
        er2, nr2, pr2, _ = st2.score_batch(g.params(), cands[fits], compact="runs")
        monkeypatch.setenv("HC_HOST_WHOLE_MAX", "0")                                # two-slot pipeline
        er3, nr3, pr3, _ = st2.score_batch(g.params(), cands[fits], compact="runs")
        monkeypatch.delenv("HC_HOST_CHUNK"); monkeypatch.delenv("HC_HOST_WHOLE_MAX")
    assert p1.tobytes() == p2.tobytes() and e1.tobytes() == e2.tobytes() and np.array_equal(n1, n2)
    assert pr.tobytes() == p1[fits].tobytes() and pr2.tobytes() == pr.tobytes() and pr3.tobytes() == pr.tobytes()
    assert er2.tobytes() == er.tobytes() and er3.tobytes() == er.tobytes() and np.array_equal(nr2, nr) and np.array_equal(nr3, nr)
    assert len(er) == int((p1["cls"][fits] == 1).sum())
    with capi.Store(g.rs, first_device=1, n_devices=1) as st3:           # a store that lives on the second device only
        e3, n3, p3, _ = st3.score_batch(g.params(), cands)
    assert p1.tobytes() == p3.tobytes() and e1.tobytes() == e3.tobytes()


def test_million_candidates_against_the_reference_binary(built_lib, c4_small, tmp_path):
    """One million candidates of the config-4 / config-3 shaped workload through the UNMODIFIED reference itself
    (oracle/_ref/ref_driver --dump-cands = EdgeCalculator::compute_overlap per candidate, not the C restatement): classes,
    mismatch rates, pos3 / pos4 bit for bit; scores within 1e-6 relative in the default mode and within 1e-14 (the device
    exp) with HC_FLAG_EXACT_EDGE_SCORES."""
    if not O.have_ref():
        pytest.skip("oracle/_ref/ref_driver has not travelled to this box")
    rs, cands = c4_small
    sub = cands[:1_000_000]
    used = np.unique(np.concatenate([sub["idx1"], sub["idx2"]]))
    small = rs.subset(used.tolist())
    c = sub.copy()
    c["idx1"] = np.searchsorted(used, sub["idx1"]).astype(np.uint32)
    c["idx2"] = np.searchsorted(used, sub["idx2"]).astype(np.uint32)
    d = str(tmp_path)
    F.write_fastq_set(small, d + "/s.fastq", d + "/p1.fastq", d + "/p2.fastq")
    F.write_overlaps(d + "/ov.txt", c, small.ids)
    out = O.run_ref(d, d + "/ov.txt", paired1=d + "/p1.fastq", paired2=d + "/p2.fastq", dump_cands=True, threads=1, edge_threshold=0.97,
                    min_overlap_len=150)        # the generator keeps candidates whose mates overlap by >= 75 = half of it each
    ref = out["cands"]
    assert len(ref) == len(c)                                    # every candidate passes the pre-filter (both mates overlap >= L/2)
    with capi.Store(small) as st:
        _, _, per, _ = st.score_batch(F.make_params(edge_threshold=0.97), c)
        ex, _, _, _ = st.score_batch(F.make_params(edge_threshold=0.97, flags=F.FLAG_EXACT_EDGE_SCORES), c, per_candidate=False)
    assert np.array_equal(per["cls"], ref["cls"])
    assert np.array_equal(per["mismatch_rate"], ref["mismatch_rate"])
    assert np.array_equal(per["pos3"], ref["pos3"]) and np.array_equal(per["pos4"], ref["pos4"])
    assert np.allclose(per["score"], ref["score"], rtol=1e-6, atol=0)
    ei = np.nonzero(ref["cls"] == F.CLASS_EDGE)[0]
    assert np.array_equal(ex["cand"], ei.astype(np.uint64)) and len(ei) > 10_000
    assert np.allclose(ex["score"], ref["score"][ei], rtol=1e-14, atol=0)
