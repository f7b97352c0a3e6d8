"""Large seeded synthetic workloads, generated with torch tensor ops so
that the 10 M-pair read set of config 4 is produced in about a minute on the GPU (or, at test
sizes, on the CPU).  torch is used for plumbing only: this module produces INPUT data.

C4: genome 100 kb, 4 haplotypes at 0.5-2 % divergence, 2x150 bp pairs stored forward-forward,
insert ~ N(450, 50), Illumina-like position dependent qualities (Q in {2, 11..41}), substitution
errors at 10^(-Q/10), 0.05 % N with Q = 0.  Reads get their IDs in generation order (random genome
position, like a shuffled FASTQ).  Candidates: every pair of read-pairs whose left mates start
within `D` ranks of each other in position order and whose two mate overlaps are both >= 75 bp (the
P-P pre-filter 0.5*min_overlap_len of src/EdgeCalculator.cpp:618-620 at m = 150), written as P-P
records with the true POS1/POS2/ORD and LEN/PERC as scripts/sfo2overlaps.py:161,190-195 computes
them, sorted by (min id, max id) like sfo2overlaps.py:52 sorts its output.  With D = 140 this is
~100 partners per pair, ~1e9 candidates for 10 M pairs.  Shard k of G is the k-th contiguous range
of that sorted list, generated directly without materialising the other shards.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

from . import formats as F


@dataclass
class PairedReads:
    n_pairs: int
    read_len: int
    bases: torch.Tensor    # uint8 CPU [n_pairs * 2 * read_len], mate 0 then mate 1 per pair
    quals: torch.Tensor    # uint8 CPU
    s1: torch.Tensor       # int32 (device) genome start of the left mate
    s2: torch.Tensor       # int32 (device) genome start of the right mate

    def readset(self) -> F.ReadSet:
        n, L = self.n_pairs, self.read_len
        descs = np.zeros(n, dtype=F.READ_DESC)
        off = np.arange(n, dtype=np.uint64) * np.uint64(2 * L)
        descs["seq_off"][:, 0] = off
        descs["seq_off"][:, 1] = off + np.uint64(L)
        descs["seq_len"][:, :] = L
        return F.ReadSet(ids=np.arange(n, dtype=np.uint64), descs=descs, bases=self.bases.numpy(), quals=self.quals.numpy(),
                         n_single=0)


def make_paired_reads(n_pairs: int, read_len: int = 150, genome_len: int = 100_000, n_hap: int = 4,
                      divergence=(0.0, 0.005, 0.01, 0.02), insert=(450.0, 50.0), n_rate: float = 0.0005,
                      seed: int = 20261018, device: str = "cpu", chunk: int = 1 << 20,
                      position_sorted_ids: bool = False, binned_qualities: bool = False) -> PairedReads:
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    base = torch.randint(0, 4, (genome_len,), generator=g, device=dev, dtype=torch.uint8)
    haps = [base]
    for h in range(1, n_hap):
        m = torch.rand(genome_len, generator=g, device=dev) < divergence[min(h, len(divergence) - 1)]
        sub = torch.randint(1, 4, (genome_len,), generator=g, device=dev, dtype=torch.uint8)
        haps.append(torch.where(m, (base + sub) % 4, base))
    H = torch.stack(haps)                                                  # [n_hap, genome_len]
    L = read_len
    bases = torch.empty(n_pairs * 2 * L, dtype=torch.uint8, pin_memory=(dev.type == "cuda"))
    quals = torch.empty(n_pairs * 2 * L, dtype=torch.uint8, pin_memory=(dev.type == "cuda"))
    s1_all = torch.empty(n_pairs, dtype=torch.int32, device=dev)
    s2_all = torch.empty(n_pairs, dtype=torch.int32, device=dev)
    ar = torch.arange(L, device=dev)
    qmean = (36.0 - 8.0 * ar.float() / max(L - 1, 1))[None, :]
    acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    # diagnostic layout only: read IDs in genome-position order (partners become neighbours in memory)
    u_all = torch.sort(torch.rand(n_pairs, generator=g, device=dev))[0] if position_sorted_ids else None
    for lo in range(0, n_pairs, chunk):
        m = min(chunk, n_pairs - lo)
        hap = torch.randint(0, n_hap, (m,), generator=g, device=dev)
        ins = torch.clamp(torch.round(insert[0] + insert[1] * torch.randn(m, generator=g, device=dev)), L, genome_len).long()
        u = u_all[lo:lo + m] if position_sorted_ids else torch.rand(m, generator=g, device=dev)
        st = (u * (genome_len - ins + 1).float()).long()
        st = torch.minimum(st, genome_len - ins)
        s1 = st
        s2 = st + ins - L
        s1_all[lo:lo + m] = s1.int()
        s2_all[lo:lo + m] = s2.int()
        out_b = torch.empty((m, 2, L), dtype=torch.uint8, device=dev)
        out_q = torch.empty((m, 2, L), dtype=torch.uint8, device=dev)
        for mate, s in ((0, s1), (1, s2)):
            idx = s[:, None] + ar[None, :]
            seg = H[hap[:, None], idx]                                     # true bases
            q = torch.clamp(torch.round(qmean + 4.0 * torch.randn((m, L), generator=g, device=dev)), 11, 41)
            low = torch.rand((m, L), generator=g, device=dev) < 0.01
            q = torch.where(low, torch.full_like(q, 2.0), q)
            if binned_qualities:   # diagnostic: the four quality bins of current Illumina instruments (2, 12, 23, 37)
                q = torch.where(q < 7, torch.full_like(q, 2.0), torch.where(q < 18, torch.full_like(q, 12.0),
                                torch.where(q < 30, torch.full_like(q, 23.0), torch.full_like(q, 37.0))))
            err = torch.rand((m, L), generator=g, device=dev) < torch.pow(10.0, -q / 10.0)
            sub = torch.randint(1, 4, (m, L), generator=g, device=dev, dtype=torch.uint8)
            seg = torch.where(err, (seg + sub) % 4, seg)
            b = acgt[seg.long()]
            nn = torch.rand((m, L), generator=g, device=dev) < n_rate
            b = torch.where(nn, torch.full_like(b, ord("N")), b)
            q = torch.where(nn, torch.zeros_like(q), q)
            out_b[:, mate, :] = b
            out_q[:, mate, :] = (q + 33).to(torch.uint8)
        bases[lo * 2 * L:(lo + m) * 2 * L].copy_(out_b.reshape(-1))
        quals[lo * 2 * L:(lo + m) * 2 * L].copy_(out_q.reshape(-1))
    return PairedReads(n_pairs=n_pairs, read_len=L, bases=bases, quals=quals, s1=s1_all, s2=s2_all)


def shard_id_range(n: int, k: int, G: int) -> Tuple[int, int]:
    """Reads whose candidates (as the smaller id of the pair) make up the k-th of G equal parts of
    the (min id, max id)-sorted list: a read x is the smaller id in a share (1 - x/n) of its pairs,
    so the cumulative candidate fraction is 1 - (1 - x/n)^2."""
    a = int(round(n * (1.0 - math.sqrt(1.0 - k / G))))
    b = int(round(n * (1.0 - math.sqrt(1.0 - (k + 1) / G)))) if k + 1 < G else n
    return a, b


def make_pp_candidates(pr: PairedReads, D: int = 140, min_half: int = 75, shard: int = 0, n_shards: int = 1,
                       max_cands: Optional[int] = None) -> torch.Tensor:
    """Returns a uint8 tensor [M, 32] of hc_candidate records on pr.s1.device, sorted by (min id, max id)."""
    dev = pr.s1.device
    n, L = pr.n_pairs, pr.read_len
    s1, s2 = pr.s1.long(), pr.s2.long()
    order = torch.argsort(s1, stable=True)                # rank -> read id
    rank = torch.empty_like(order)
    rank[order] = torch.arange(n, device=dev)
    a, b = shard_id_range(n, shard, n_shards)
    I = torch.arange(a, b, device=dev)
    rI = rank[I]
    s1I, s2I = s1[I], s2[I]
    out_i, out_j, out_d1, out_d2 = [], [], [], []
    total = 0
    for d in [x for k in range(1, D + 1) for x in (k, -k)]:
        rr = rI + d
        ok = (rr >= 0) & (rr < n)
        J = order[rr.clamp(0, n - 1)]
        ok &= J > I
        d1 = s1[J] - s1I
        d2 = s2[J] - s2I
        # the left read is id1 (scripts/sfo2overlaps.py:164-185); its partner's mates are shifted by d1 / d2
        ok &= (L - d1.abs() >= min_half) & (L - d2.abs() >= min_half)
        if ok.any():
            out_i.append(I[ok].int()); out_j.append(J[ok].int()); out_d1.append(d1[ok].int()); out_d2.append(d2[ok].int())
            total += int(ok.sum())
    i = torch.cat(out_i).long(); j = torch.cat(out_j).long(); d1 = torch.cat(out_d1).long(); d2 = torch.cat(out_d2).long()
    del out_i, out_j, out_d1, out_d2
    perm = torch.argsort(i * n + j)
    if max_cands is not None and perm.numel() > max_cands:
        perm = perm[:max_cands]
    i, j, d1, d2 = i[perm], j[perm], d1[perm], d2[perm]
    swap = d1 < 0                                          # partner starts further left: it becomes ID1
    id1 = torch.where(swap, j, i)
    id2 = torch.where(swap, i, j)
    pos1 = d1.abs()
    d2o = torch.where(swap, -d2, d2)                       # right-mate shift of id2 relative to id1
    ord_ = torch.where(d2o >= 0, torch.full_like(d2o, ord("1")), torch.full_like(d2o, ord("2")))
    pos2 = d2o.abs()
    len1, len2 = L - pos1, L - pos2
    perc1 = torch.clamp(torch.round(100.0 * len1.double() / L), max=100).long()
    perc2 = torch.clamp(torch.round(100.0 * len2.double() / L), max=100).long()
    rec = torch.empty((i.numel(), 8), dtype=torch.int32, device=dev)
    rec[:, 0] = id1.int(); rec[:, 1] = id2.int(); rec[:, 2] = pos1.int(); rec[:, 3] = pos2.int()
    rec[:, 4] = len1.int(); rec[:, 5] = len2.int()
    rec[:, 6] = (perc1 | (perc2 << 8) | (ord_ << 16) | (1 << 24)).int()
    rec[:, 7] = (1 | (ord("p") << 8) | (ord("p") << 16)).__int__()
    return rec.view(torch.uint8).reshape(-1, 32)


def candidates_as_numpy(rec: torch.Tensor) -> np.ndarray:
    return rec.cpu().numpy().reshape(-1).view(F.CANDIDATE)
