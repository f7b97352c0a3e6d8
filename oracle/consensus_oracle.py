"""CPU restatement of SRBuilder::consensus / consensus_pos (src/SRBuilder.cpp:297-522) -- TEST INFRASTRUCTURE ONLY.

Pure Python on the same libm (math.log10, math.pow are glibc's log10 / pow), pinned bit for bit on what the unmodified
reference returned for the problems of tests/golden/consensus_*.npz (oracle/make_golden.py: run_consensus_case)."""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple


def phred_to_prob(q: int) -> float:                       # src/SRBuilder.cpp:289-293
    return math.pow(10, -q / 10.0)


def _c_round(x: float) -> int:                           # C round(): halves away from zero
    return int(math.copysign(math.floor(abs(x) + 0.5), x))


def consensus_pos(nucs: str, quals: str, min_qual: float):
    """:297-402.  Returns (ok, base, quality character); ok False <=> the reference returns 0 (NaN)."""
    sA = sC = sT = sG = 0.0
    for n, q in zip(nucs, quals):
        Q = ord(q) - 33
        p = phred_to_prob(Q)
        hit, miss = (math.log10(1 - p) if p < 1 else float("-inf")), (math.log10(p / 3.0) if p > 0 else float("-inf"))
        if n == "A":
            sA += hit; sC += miss; sT += miss; sG += miss
        elif n == "C":
            sC += hit; sA += miss; sT += miss; sG += miss
        elif n == "T":
            sT += hit; sC += miss; sA += miss; sG += miss
        elif n == "G":
            sG += hit; sC += miss; sT += miss; sA += miss
    mx = max(sA, sT, sC, sG)
    max_prob = math.pow(10.0, mx)
    total = math.pow(10.0, sA) + math.pow(10.0, sT) + math.pow(10.0, sC) + math.pow(10.0, sG)
    if mx == 0 or total == 0.0:
        return True, "N", "$"
    p_inc = 1 - (max_prob / total)
    if len(nucs) > 1 and (1 - p_inc) < min_qual:
        return True, "N", "$"
    if p_inc != p_inc:
        return False, "", ""
    if p_inc < math.pow(10.0, -9.3):
        phred = 93
    else:
        phred = _c_round(-10 * math.log10(p_inc)) if p_inc > 0 else 93
    phred = min(max(phred, 0), 93)
    base = "A" if mx == sA else ("T" if mx == sT else ("C" if mx == sC else "G"))
    return True, base, chr(phred + 33)


def consensus(total_len: int, pos: Sequence[int], seqs: Sequence[str], quals: Sequence[str], subreads_needed: bool,
              error_correction: bool, min_clique_size: int, min_qual: float) -> Tuple[int, str, str]:
    """:406-522.  Returns (return value, cons_seq, cons_qual)."""
    n = len(pos)
    min_support = 2 if subreads_needed else min_clique_size
    if error_correction:
        k, support = 0, 1
        while support < min_support and k < n:
            support += 1
            k += 1
        if k == n:
            return -1, "", ""
        trim = pos[k]
    else:
        trim = 0
    active_pos = [trim - p if p < trim else 0 for p in pos]
    active = [False] * n
    nxt = 0
    prefix_removed = False
    cs, cq = [], []
    for cur in range(total_len):
        while nxt < n and cur == pos[nxt]:
            active[nxt] = True
            nxt += 1
        if error_correction and sum(active) < min_support:
            if nxt == n:
                break
            if not prefix_removed:
                continue
        prefix_removed = True
        nucs, qs = [], []
        for j in range(n):
            if active[j]:
                p = active_pos[j]
                if p >= len(seqs[j]) or p >= len(quals[j]):
                    return 0, "", ""
                nucs.append(seqs[j][p])
                qs.append(quals[j][p])
                if p + 1 < len(seqs[j]):
                    active_pos[j] = p + 1
                else:
                    active[j] = False
        if not nucs:
            return 0, "", ""
        ok, b, q = consensus_pos("".join(nucs), "".join(qs), min_qual)
        if not ok:
            return trim, "", ""
        cs.append(b)
        cq.append(q)
    return trim, "".join(cs), "".join(cq)
