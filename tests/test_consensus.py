"""CPU suite: the restatement of SRBuilder::consensus / consensus_pos (oracle/consensus_oracle.py) against what the
UNMODIFIED reference returned for the pile-ups of tests/golden/consensus_*.npz: return value and both strings."""
import pytest

from util import ConsensusGolden, consensus_golden_names, consensus_oracle_results


@pytest.mark.parametrize("name", consensus_golden_names())
def test_consensus_restatement_matches_reference(name):
    g = ConsensusGolden(name)
    got = consensus_oracle_results(g.rs, g.problems(), g.min_clique_size, g.min_qual)
    assert got == g.ref
    assert sum(1 for r in g.ref if r[1]) > 0.4 * len(g.ref) and any(r[0] == -1 for r in g.ref) and any("N" in r[1] for r in g.ref)
