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
