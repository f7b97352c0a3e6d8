"""GPU suite, two or more GPUs (one box): the two ways the path spreads over devices.
  * one process, a store replicated on two devices: hc_score_batch* cuts the batch into contiguous shards and gathers
    the lists in input order (small outputs included);
  * one process per GPU: every rank scores its contiguous range with hc_score_batch_device and the accepted-edge lists are
    concatenated by the NCCL all-gather of haploconduct_b200/dist.py (DeviceGather), device to device.
Both must reproduce the one-device result byte for byte.  On a box with a single GPU there is nothing to run (skip); with two
or more these tests run -- the log of such a run is kept under profiles/."""
import os
import socket

import numpy as np
import pytest

from haploconduct_b200 import capi, formats as F
from util import load_golden

pytestmark = pytest.mark.gpu


def _two():
    return capi.device_count() >= 2


@pytest.mark.skipif(not _two(), reason="one GPU on this box")
def test_small_outputs_on_two_devices(built_lib, monkeypatch):
    g = load_golden("synth_all_types")
    cands = np.tile(g.scored(), 7)
    cands = cands[(cands["pos1"] < (1 << 14)) & (cands["pos2"] < (1 << 14))]
    for exact in (False, True):
        p = g.params(flags=F.FLAG_EXACT_EDGE_SCORES if exact else 0)
        with capi.Store(g.rs) as st1:
            e1, f1, _ = st1.score_batch_small(p, cands)
        with capi.Store(g.rs, first_device=0, n_devices=2) as st2:
            e2, f2, _ = st2.score_batch_small(p, cands)
            monkeypatch.setenv("HC_HOST_CHUNK", "1000")
            e3, f3, _ = st2.score_batch_small(p, cands, runs=False)
            monkeypatch.delenv("HC_HOST_CHUNK")
        assert e1.tobytes() == e2.tobytes() and np.array_equal(f1, f2)
        assert e1.tobytes() == e3.tobytes() and np.array_equal(f1, f3)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank(rank, world, port, name, outdir):
    import torch
    import torch.distributed as dist
    from haploconduct_b200 import dist as D

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    g = load_golden(name)
    cands = np.tile(g.scored(), 9)
    lo, hi = D.shard_range(len(cands), rank, world)
    mine = np.ascontiguousarray(cands[lo:hi])
    n = len(mine)
    with capi.Store(g.rs, first_device=rank, n_devices=1) as st:
        d_c = torch.from_numpy(mine.view(np.uint8).reshape(-1)).to(dev)
        cap = len(cands) // world + 2                                   # the same on every rank: the lists are padded to it
        d_e = torch.zeros((cap, 48), dtype=torch.uint8, device=dev)
        d_n = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        d_cnt = torch.zeros(4, dtype=torch.int64, device=dev)
        stream = torch.cuda.current_stream(dev)
        st.score_batch_device(rank, stream.cuda_stream, g.params(), d_c.data_ptr(), n, 0, d_e.data_ptr(), n, d_n.data_ptr(), n, d_cnt.data_ptr(), False)
        d_e.view(torch.int64).reshape(cap, 6)[:, 0] += lo              # hc_edge.cand: local -> global index
        gather = D.DeviceGather(48, cap, dev)
        gather.gather(d_e, d_cnt[:1])
        allv = gather.concatenated().cpu().numpy().view(F.EDGE)
        # the same by one-sided puts over peer memory (copy engines), twice in a row: the second gather must wait for the
        # readers of the first (barrier protocol); where peer memory cannot be set up the factory hands back the NCCL gather
        pg, kind = D.make_device_gather(48, cap, dev)
        for rep in range(2):
            pg.gather(d_e, d_cnt[:1])
            torch.cuda.synchronize(dev)
            assert pg.concatenated().cpu().numpy().tobytes() == allv.tobytes(), "rank %d: %s gather, repetition %d" % (rank, kind, rep)
        with open(os.path.join(outdir, "kind%d.txt" % rank), "w") as f:
            f.write(kind)
    np.save(os.path.join(outdir, "r%d.npy" % rank), allv)
    dist.destroy_process_group()


@pytest.mark.skipif(not _two(), reason="one GPU on this box")
def test_nccl_gather_of_device_lists_equals_one_device(built_lib, tmp_path):
    import torch.multiprocessing as mp

    name = "synth_all_types"
    world = min(capi.device_count(), 4)
    mp.spawn(_rank, args=(world, _free_port(), name, str(tmp_path)), nprocs=world, join=True)
    g = load_golden(name)
    cands = np.tile(g.scored(), 9)
    with capi.Store(g.rs) as st:
        edges, _, _, _ = st.score_batch(g.params(), cands, per_candidate=False)
    for r in range(world):
        got = np.load(str(tmp_path / ("r%d.npy" % r)))
        assert got.tobytes() == edges.tobytes(), "rank %d" % r
    print("second gather implementation on this box:", (tmp_path / "kind0.txt").read_text())
