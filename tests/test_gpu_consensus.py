"""GPU suite: hc_consensus (per-column scores on the device, pow/log10 step and column walk on the host) against the
reference's SRBuilder::consensus on the golden pile-ups and against the pinned restatement on fresh ones."""
import numpy as np
import pytest

from haploconduct_b200 import capi, workloads as W
from util import ConsensusGolden, consensus_golden_names, consensus_oracle_results

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["packed", "planar"])
def layout(request, monkeypatch):
    if request.param == "planar":
        monkeypatch.setenv("HC_STORE_LAYOUT", "planar")
    return request.param


@pytest.mark.parametrize("host_all", [False, True])
@pytest.mark.parametrize("name", consensus_golden_names())
def test_consensus_reproduces_reference(built_lib, monkeypatch, name, host_all):
    """Default: pow / log10 on the device, only columns within the error bound of a decision go back to the host libm.
    HC_CONS_HOST_ALL: every column through the host libm.  Both must give the reference's strings."""
    if host_all:
        monkeypatch.setenv("HC_CONS_HOST_ALL", "1")
    g = ConsensusGolden(name)
    with capi.Store(g.rs) as st:
        got = st.consensus(g.problems(), g.min_clique_size, g.min_qual)
    assert got == g.ref            # return value, consensus sequence and quality string of every pile-up


@pytest.mark.parametrize("seed,mcs,mq,kw", [(11, 2, 0.9, {}), (12, 4, 0.99, {}), (13, 1, 0.5, dict(qmax=93)), (14, 3, 0.9, dict(n_rate=0.2))])
def test_consensus_fresh_pileups_against_restatement(built_lib, seed, mcs, mq, kw):
    rs, probs = W.consensus_problems(seed=seed, n_problems=150, **kw)
    want = consensus_oracle_results(rs, probs, mcs, mq)
    with capi.Store(rs) as st:
        got = st.consensus(probs, mcs, mq)
        assert st.consensus([], mcs, mq) == []
    assert got == want


def test_consensus_rejects_bad_problems(built_lib):
    rs, probs = W.consensus_problems(seed=1, n_problems=5)
    with capi.Store(rs) as st:
        bad = [dict(probs[0], entries=[(rs.n_reads, 0, False, 0)])]
        with pytest.raises(capi.HcError):
            st.consensus(bad, 2, 0.9)
        e = probs[1]["entries"]
        if len(e) > 1:
            bad = [dict(probs[1], entries=[e[0], (e[1][0], e[1][1], e[1][2], -1)])]
            with pytest.raises(capi.HcError):
                st.consensus(bad, 2, 0.9)


def test_device_decisions_agree_with_host_libm_at_scale(built_lib, monkeypatch, layout):
    """A few million columns: the characters decided on the device (unmarked columns) must be the ones the host libm
    gives for the same scores -- i.e. the marking covers every column whose outcome could differ."""
    if layout == "planar":
        pytest.skip("one layout is enough for this one")
    rs, base = W.consensus_problems(seed=21, n_problems=1500, qmax=60)
    probs = base * 12
    with capi.Store(rs) as st:
        dev = st.consensus(probs, 3, 0.93)
        monkeypatch.setenv("HC_CONS_HOST_ALL", "1")
        host = st.consensus(probs, 3, 0.93)
    assert dev == host
    assert sum(len(r[1]) for r in dev) > 2_000_000
