#!/usr/bin/env python
"""Sustained host<->device bandwidth with N processes (one per GPU) running at once.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 benchmarks/probe_pcie_multi.py

Every rank moves one bench-sized batch (4096 x 4 x 2500 fp32 = 164 MB) between pinned host memory
and its GPU for ``--seconds`` per leg and reports the MEAN rate (not the best repetition): H2D alone,
D2H alone, both directions at once on two streams, and the same three legs done by an SM kernel over
unified addressing (``pcgmix_copy_small``) instead of the copy engines.  All ranks start every leg
together (gloo barrier), so the per-rank numbers add up to what the box delivers.  This is the
ceiling ``bench.py``'s ``e2e`` leg is held against when N > 1.
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=0.6)
    ap.add_argument("--mb", type=float, default=163.84)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("gloo")
    from pcgmix_b200 import native
    native.load()

    n = int(args.mb * 1e6 / 4) // 4 * 4
    h_in = torch.empty(n, dtype=torch.float32).pin_memory()
    h_out = torch.empty(n, dtype=torch.float32).pin_memory()
    h_in.normal_()
    d_in = torch.empty(n, dtype=torch.float32, device="cuda")
    d_out = torch.empty(n, dtype=torch.float32, device="cuda").normal_()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    gb = n * 4 / 1e9

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def leg(up, down, sm=False):
        """mean GB/s per direction while every rank runs the same leg for args.seconds"""
        def issue():
            if up:
                with torch.cuda.stream(s1):
                    if sm:
                        native.copy_small(d_in, h_in, n * 4)
                    else:
                        d_in.copy_(h_in, non_blocking=True)
            if down:
                with torch.cuda.stream(s2):
                    if sm:
                        native.copy_small(h_out, d_out, n * 4)
                    else:
                        h_out.copy_(d_out, non_blocking=True)
        for _ in range(2):
            issue()
        barrier()
        t0 = time.perf_counter()
        reps = 0
        while time.perf_counter() - t0 < args.seconds:
            issue()
            s1.synchronize()
            s2.synchronize()
            reps += 1
        dt = time.perf_counter() - t0
        barrier()
        return reps * gb / dt

    res = {"rank": rank, "world": world, "mb": n * 4 / 1e6,
           "h2d_alone_GBps": leg(True, False), "d2h_alone_GBps": leg(False, True),
           "both_each_GBps": leg(True, True),
           "sm_h2d_alone_GBps": leg(True, False, True), "sm_d2h_alone_GBps": leg(False, True, True),
           "sm_both_each_GBps": leg(True, True, True),
           "cpus": len(os.sched_getaffinity(0))}
    line = json.dumps(res)
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, res)
        if rank == 0:
            total = {k: sum(g[k] for g in gathered) for k in res if k.endswith("GBps")}
            line = json.dumps({"per_rank": gathered, "box_total": total})
    if rank == 0:
        print(line)
        if args.out:
            with open(args.out, "w") as f:
                f.write(line + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
