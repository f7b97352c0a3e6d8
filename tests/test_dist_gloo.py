"""CPU suite: the N > 1 host logic (contiguous sharding + ordered gather) on 2 gloo processes.
The per-shard scorer is the CPU oracle here (no GPU in this suite); the gather code is the
product's own (haploconduct_b200/dist.py) and is what bench.py / a multi-GPU caller uses."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from util import load_golden


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, outdir):
    import torch.distributed as dist
    from haploconduct_b200 import dist as D, formats as F
    from oracle import oracle as O
    from util import load_golden

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = load_golden(name)
    cands = g.scored()
    lo, hi = D.shard_range(len(cands), rank, world)
    res, _ = O.score_batch(g.rs, g.params(), cands[lo:hi])
    ei = np.nonzero(res["cls"] == F.CLASS_EDGE)[0]
    edges = np.zeros(len(ei), dtype=F.EDGE)
    edges["cand"] = ei
    edges["score"] = res["score"][ei]
    edges["mismatch_rate"] = res["mismatch_rate"][ei]
    edges["pos3"] = res["pos3"][ei]
    edges["pos4"] = res["pos4"][ei]
    nonedge = np.nonzero(res["cls"] == F.CLASS_NONEDGE)[0].astype(np.uint64)
    ge, gn = D.gather_results(edges, nonedge, lo)
    np.savez(os.path.join(outdir, "r%d.npz" % rank), edges=ge, nonedge=gn)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_gather_equals_single_process_order(tmp_path, world):
    name = "synth_all_types"
    mp.spawn(_worker, args=(world, _free_port(), name, str(tmp_path)), nprocs=world, join=True)
    g = load_golden(name)
    ref = g.ref_cands
    want_e = np.nonzero(ref["cls"] == 1)[0].astype(np.uint64)
    want_n = np.nonzero(ref["cls"] == 2)[0].astype(np.uint64)
    for r in range(world):
        z = np.load(str(tmp_path / ("r%d.npz" % r)))
        assert np.array_equal(z["edges"]["cand"], want_e)          # rank order == input order
        assert np.array_equal(z["nonedge"], want_n)
        assert np.array_equal(z["edges"]["score"], ref["score"][want_e.astype(np.int64)])


def test_shard_ranges_tile_the_batch():
    from haploconduct_b200 import dist as D

    for n in (0, 1, 7, 1000, 124717900):
        for w in (1, 2, 3, 8):
            r = [D.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))


def _gather_worker(rank, world, port, outdir):
    import torch
    import torch.distributed as dist
    from haploconduct_b200 import dist as D

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rec, cap = 24, 50
    g = D.DeviceGather(rec, cap, torch.device("cpu"))
    for rnd in range(2):                      # buffers are reused between calls
        n = (7 * rank + 3 + rnd) % cap
        mine = torch.zeros((cap, rec), dtype=torch.uint8)
        mine[:n] = torch.arange(n * rec, dtype=torch.int64).reshape(n, rec).to(torch.uint8) + rank
        mine[n:] = 255                        # beyond the count: must not show up
        g.gather(mine, torch.tensor([n], dtype=torch.int64))
        got = g.concatenated()
        want = []
        for r in range(world):
            k = (7 * r + 3 + rnd) % cap
            want.append((torch.arange(k * rec, dtype=torch.int64).reshape(k, rec).to(torch.uint8) + r).reshape(-1))
        assert torch.equal(got, torch.cat(want)), (rank, rnd)
    open(os.path.join(outdir, "ok%d" % rank), "w").close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_device_gather_concatenates_in_rank_order(tmp_path, world):
    """DeviceGather (the all-gather bench.py times at N > 1; NCCL there, gloo here): ragged lists, rank order."""
    mp.spawn(_gather_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / ("ok%d" % r)).exists() for r in range(world))


def _worker_padded(rank, world, port, outdir):
    import torch
    import torch.distributed as dist
    from haploconduct_b200 import dist as D

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = torch.device("cpu")
    rec, cap = 48, 40
    n = 7 + 11 * rank                                   # lists of different lengths, padded to one capacity
    mine = torch.zeros((cap, rec), dtype=torch.uint8)
    mine[:n] = torch.arange(n * rec, dtype=torch.int64).reshape(n, rec).remainder(251).to(torch.uint8) + rank
    g, kind = D.make_device_gather(rec, cap, dev)       # no peer memory on CPU tensors: the collective gather
    assert kind == "nccl" and isinstance(g, D.DeviceGather)
    for _ in range(2):
        g.gather(mine, torch.tensor([n], dtype=torch.int64))
        got = g.concatenated().numpy()
    np.save(os.path.join(outdir, "p%d.npy" % rank), got)
    dist.destroy_process_group()


def test_padded_list_gather_factory_on_cpu(tmp_path):
    """dist.make_device_gather hands back the collective gather where peer memory is not available (CPU tensors here; a box
    without P2P on GPUs): counts first, lists padded to the agreed capacity, concatenated in rank order."""
    world = 3
    mp.spawn(_worker_padded, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    want = []
    for r in range(world):
        n = 7 + 11 * r
        want.append((np.arange(n * 48, dtype=np.int64).reshape(n, 48) % 251).astype(np.uint8) + r)
    want = np.concatenate(want).reshape(-1)
    for r in range(world):
        assert np.array_equal(np.load(str(tmp_path / ("p%d.npy" % r))), want)
