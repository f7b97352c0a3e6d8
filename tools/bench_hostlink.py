#!/usr/bin/env python
"""What the host side of an N-rank e2e step can move: every rank copies pinned host buffers to its GPU, back, and both at
once, all ranks at the same time (barrier before every leg, CUDA events on the copy streams, max over ranks) -- with and
without pinning the process to the GPU's NUMA node.  Run under torchrun like bench.py:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_hostlink.py [--numa-bind]

Rank 0 prints one JSON line: per-rank and aggregate GB/s of the three legs, the bytes of one e2e step of bench.py next to them."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--numa-bind", action="store_true")
    ap.add_argument("--write-combined", action="store_true", help="the host->device source in write-combined pinned memory (hc_host_alloc)")
    ap.add_argument("--mb-in", type=int, default=1024)
    ap.add_argument("--mb-out", type=int, default=256)
    ap.add_argument("--reps", type=int, default=8)
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    numa = None
    if a.numa_bind:
        import bench
        numa = bench.bind_to_gpu_numa_node(local)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if a.write_combined:
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from haploconduct_b200 import capi
        h_in = torch.from_numpy(capi.host_alloc(a.mb_in << 20, write_combined=True)); h_in.fill_(1)
    else:
        h_in = torch.empty(a.mb_in << 20, dtype=torch.uint8, pin_memory=True); h_in.fill_(1)
    d_in = torch.empty(a.mb_in << 20, dtype=torch.uint8, device=dev)
    h_out = torch.empty(a.mb_out << 20, dtype=torch.uint8, pin_memory=True); h_out.fill_(1)
    d_out = torch.empty(a.mb_out << 20, dtype=torch.uint8, device=dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def leg(do_in, do_out):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        barrier()
        if do_in:
            with torch.cuda.stream(s_in):
                ev[0].record(s_in)
                for _ in range(a.reps):
                    d_in.copy_(h_in, non_blocking=True)
                ev[1].record(s_in)
        if do_out:
            with torch.cuda.stream(s_out):
                ev[2].record(s_out)
                for _ in range(a.reps):
                    h_out.copy_(d_out, non_blocking=True)
                ev[3].record(s_out)
        barrier()
        gin = a.reps * h_in.numel() / ev[0].elapsed_time(ev[1]) / 1e6 if do_in else 0.0
        gout = a.reps * h_out.numel() / ev[2].elapsed_time(ev[3]) / 1e6 if do_out else 0.0
        t = torch.tensor([gin, gout], dtype=torch.float64, device=dev)
        if world > 1:
            lo = t.clone(); dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            sm = t.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            return {"h2d_min_rank": float(lo[0]), "h2d_sum": float(sm[0]), "d2h_min_rank": float(lo[1]), "d2h_sum": float(sm[1])}
        return {"h2d_min_rank": gin, "h2d_sum": gin, "d2h_min_rank": gout, "d2h_sum": gout}

    leg(True, True)   # warm-up
    res = {"h2d_only": leg(True, False), "d2h_only": leg(False, True), "both": leg(True, True)}
    if rank == 0:
        print(json.dumps({"ranks": world, "write_combined": bool(a.write_combined), "numa_bind": bool(a.numa_bind), "numa_rank0": numa, "unit": "GB/s", "mb_in": a.mb_in, "mb_out": a.mb_out,
                          "legs": res, "cpus": os.cpu_count(),
                          "e2e_step_bytes": {"h2d": 1002909812, "d2h_small": 186000000, "d2h_full": 611642544}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
