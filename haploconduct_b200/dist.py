"""One-process-per-GPU sharding of a candidate batch and the ordered gather of the results.

Candidates are independent (src/EdgeCalculator.cpp:399-414), so the batch is cut into contiguous
index ranges, rank r scores range r on its own replica of the read store, and the only exchange is
the concatenation of the per-rank accepted-edge / non-edge lists IN RANK ORDER, which equals input
order.  torch.distributed is plumbing: NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import formats as F


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous range of rank `rank`; the same split hc_score_batch uses across one process' devices."""
    return n * rank // world, n * (rank + 1) // world


def _gather_bytes(local: np.ndarray, device: torch.device, group=None) -> np.ndarray:
    """all-gather of variable-length byte strings: counts first, then max-padded payloads."""
    world = dist.get_world_size(group)
    raw = np.ascontiguousarray(local).view(np.uint8).reshape(-1)
    cnt = torch.tensor([raw.size], dtype=torch.int64, device=device)
    counts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(counts, cnt, group=group)
    sizes = [int(c.item()) for c in counts]
    m = max(max(sizes), 1)
    buf = torch.zeros(m, dtype=torch.uint8, device=device)
    if raw.size:
        buf[: raw.size] = torch.from_numpy(raw.copy()).to(device)
    outs = [torch.zeros(m, dtype=torch.uint8, device=device) for _ in range(world)]
    dist.all_gather(outs, buf, group=group)
    return np.concatenate([o[:s].cpu().numpy() for o, s in zip(outs, sizes)]) if sum(sizes) else np.zeros(0, np.uint8)


class DeviceGather:
    """Ordered concatenation of per-rank record lists that stay in device memory (no host hop): an all-gather of the
    counts, then one all-gather of the lists padded to the longest one (NCCL over NVLink on GPUs; every rank ends up
    with all lists, the consumer slices them in rank order = input order).  Buffers are allocated once and reused, so
    a call is two collectives on the given stream."""

    def __init__(self, rec_bytes: int, max_records: int, device: torch.device, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rec = rec_bytes
        self.cap = max_records
        self.device = device
        self.counts = torch.zeros(self.world, dtype=torch.int64, device=device)
        self.padded = torch.zeros((self.world, max_records * rec_bytes), dtype=torch.uint8, device=device)

    def gather(self, records: torch.Tensor, count: torch.Tensor) -> None:
        """records: uint8 tensor holding this rank's list (at least cap * rec bytes are readable), count: int64[1] on the
        device.  Asynchronous on the current stream; results in self.counts / self.padded."""
        if records.numel() < self.cap * self.rec:
            raise ValueError("DeviceGather: the record buffer must hold cap = %d records (lists are padded to it)" % self.cap)
        dist.all_gather_into_tensor(self.counts, count.reshape(1), group=self.group)
        dist.all_gather_into_tensor(self.padded.reshape(-1), records.reshape(-1)[: self.cap * self.rec], group=self.group)

    def lists(self):
        """Per-rank views of the gathered records (device tensors), rank order; synchronises to read the counts."""
        c = self.counts.cpu().tolist()
        return [self.padded[r, : c[r] * self.rec] for r in range(self.world)]

    def concatenated(self) -> torch.Tensor:
        return torch.cat(self.lists())


class PeerGather:
    """The same ordered concatenation by ONE-SIDED PUTS over NVLink peer memory: every rank owns a buffer [world][cap * rec]
    that its peers can address (CUDA IPC through torch's symmetric-memory allocator -- plumbing), and a gather is, per rank,
    `world` device-to-device copies of its list into slot `rank` of every peer's buffer plus its count, bracketed by two
    device-side barriers (peers have finished reading the previous contents / all puts have landed).  The copies run on the
    copy engines: no SM is taken from the score kernel of the next step, which an NCCL all-gather's kernels have to wait for
    (they do not fit next to its two CTAs per SM).  Same results as DeviceGather: self.counts / self.padded.
    Raises at construction if peer memory cannot be set up (no P2P, a container without the rights to pass handles): the
    caller falls back to DeviceGather."""

    def __init__(self, rec_bytes: int, max_records: int, device: torch.device, group=None):
        import torch.distributed._symmetric_memory as symm_mem

        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.rec, self.cap, self.device = rec_bytes, max_records, device
        row = max_records * rec_bytes
        row += (-row) % 16
        self.row = row
        # one symmetric allocation: [world] rows of records, then [world] int64 counts
        self.bytes = self.world * row + 8 * self.world
        self.local = symm_mem.empty(self.bytes, dtype=torch.uint8, device=device)
        self.handle = symm_mem.rendezvous(self.local, self.group)
        self.peers = [self.handle.get_buffer(r, (self.bytes,), torch.uint8) for r in range(self.world)]
        self.local.zero_()
        self.padded = self.local[: self.world * row].view(self.world, row)
        self.counts = self.local[self.world * row:].view(torch.int64)
        torch.cuda.synchronize(device)
        dist.barrier(group=self.group)

    def gather(self, records: torch.Tensor, count: torch.Tensor) -> None:
        """Asynchronous on the current stream; results in self.counts / self.padded (valid once the stream has passed)."""
        if records.numel() < self.cap * self.rec:
            raise ValueError("PeerGather: the record buffer must hold cap = %d records (lists are padded to it)" % self.cap)
        src = records.reshape(-1)[: self.cap * self.rec]
        cnt = count.reshape(1).view(torch.uint8)
        self.handle.barrier(channel=0)                       # every peer has consumed what the previous gather put here
        for k in range(self.world):                          # start with the own slot, then the peers round robin from rank + 1
            r = (self.rank + k) % self.world
            dst = self.peers[r]
            dst[self.rank * self.row: self.rank * self.row + src.numel()].copy_(src, non_blocking=True)
            o = self.world * self.row + 8 * self.rank
            dst[o: o + 8].copy_(cnt, non_blocking=True)
        self.handle.barrier(channel=1)                       # all puts of all ranks have landed

    def lists(self):
        c = self.counts.cpu().tolist()
        return [self.padded[r, : c[r] * self.rec] for r in range(self.world)]

    def concatenated(self) -> torch.Tensor:
        return torch.cat(self.lists())


def make_device_gather(rec_bytes: int, max_records: int, device: torch.device, group=None, prefer_peer: bool = True):
    """PeerGather where peer memory works, DeviceGather (NCCL) otherwise; returns (gather object, its kind)."""
    ok = torch.zeros(1, dtype=torch.int32, device=device)
    g = None
    if prefer_peer and device.type == "cuda":
        try:
            g = PeerGather(rec_bytes, max_records, device, group)
            ok += 1
        except Exception:                                    # noqa: BLE001 -- any failure means "no peer memory here"
            g = None
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)   # all ranks or none
    if int(ok.item()) == 1:
        return g, "peer"
    return DeviceGather(rec_bytes, max_records, device, group), "nccl"


def gather_results(edges: np.ndarray, nonedge_idx: np.ndarray, shard_start: int, device: Optional[torch.device] = None,
                   group=None) -> Tuple[np.ndarray, np.ndarray]:
    """`edges` (formats.EDGE) / `nonedge_idx` (uint64) hold indices local to this rank's shard;
    returns the global lists, in input order, on every rank."""
    device = device or torch.device("cpu")
    e = edges.copy()
    e["cand"] += np.uint64(shard_start)
    ne = (nonedge_idx.astype(np.uint64) + np.uint64(shard_start)).astype(np.uint64)
    ge = _gather_bytes(e, device, group).view(F.EDGE)
    gn = _gather_bytes(ne, device, group).view(np.uint64)
    return ge, gn


def score_sharded(store, params: np.ndarray, cands: np.ndarray, device: Optional[torch.device] = None, group=None):
    """Score this rank's contiguous share of `cands` on `store` (a capi.Store replica on the rank's
    GPU) and return the globally gathered (edges, nonedge_idx)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_range(len(cands), rank, world)
    edges, nonedge, _, stats = store.score_batch(params, cands[lo:hi], per_candidate=False)
    ge, gn = gather_results(edges, nonedge, lo, device=device, group=group)
    return ge, gn, stats
