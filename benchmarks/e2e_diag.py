#!/usr/bin/env python
"""Where the end-to-end step of ``bench.py`` goes when N ranks run at once: the same prefetching loop
around ``augmentations.augment`` with single pieces switched off.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 benchmarks/e2e_diag.py

Legs (ms per 4096-cycle step, max over ranks): full; without the result's D2H copy; without the
batch's H2D copy; with the host draws replaced by cached ones; copies only (no augment call at all).
Diagnostic only — nothing here is a benchmark number.
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

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("gloo")
    from pcgmix_b200 import augmentations, draws, synth

    B, C, L = 4096, 4, 2500
    batches = [bench.make_batch(synth.BENCH_SEED + 1000 * rank + i, B, C, L) for i in range(2)]
    host_in = [torch.from_numpy(b[0]).pin_memory() for b in batches]
    host_out = [torch.empty_like(host_in[0]).pin_memory() for _ in range(2)]
    frames_t = [torch.from_numpy(b[1]) for b in batches]
    ohe_t = [torch.nn.functional.one_hot(torch.from_numpy(b[2]), 2).to(dev) for b in batches]
    wav = ["a0001"] * B
    a = bench._Args(bench.METHOD, B)
    stream = torch.cuda.current_stream(dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    NIN = 3
    dev_in = [torch.from_numpy(batches[0][0]).to(dev) for _ in range(NIN)]
    in_ready = [torch.cuda.Event() for _ in range(NIN)]
    in_free = [torch.cuda.Event() for _ in range(NIN)]
    out_done = [torch.cuda.Event() for _ in range(2)]

    real = {k: getattr(draws, k) for k in ("pairing", "draw_knots", "draw_lambda")}
    cache = {}

    def cached(name):
        def fn(*p, **kw):
            if name not in cache:
                cache[name] = real[name](*p, **kw)
            return cache[name]
        return fn

    def run(n, seed0, h2d=True, d2h=True, call=True):
        for ev in in_free + out_done:
            ev.record(stream)

        def stage_in(i):
            with torch.cuda.stream(s_in):
                s_in.wait_event(in_free[i % NIN])
                if h2d:
                    dev_in[i % NIN].copy_(host_in[i % 2], non_blocking=True)
                in_ready[i % NIN].record(s_in)
        stage_in(0)
        stage_in(1)
        for i in range(n):
            j = i % 2
            if i + 2 < n:
                stage_in(i + 2)
            stream.wait_event(in_ready[i % NIN])
            if call:
                out = augmentations.augment(a, dev_in[i % NIN], ohe_t[j], frames_t[j], wav, bench._Step(seed0 + i), None, dev, None)[0]
            else:
                out = dev_in[i % NIN]
            in_free[i % NIN].record(stream)
            done = torch.cuda.Event()
            done.record(stream)
            with torch.cuda.stream(s_out):
                s_out.wait_event(done)
                s_out.wait_event(out_done[i % 2])
                if d2h:
                    host_out[i % 2].copy_(out, non_blocking=True)
                out.record_stream(s_out)
                out_done[i % 2].record(s_out)
        stream.wait_stream(s_out)
        stream.wait_stream(s_in)

    def timed(**kw):
        run(3, 100, **kw)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        run(args.steps, 1000, **kw)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / args.steps
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    res = {"world": world, "cpus": len(os.sched_getaffinity(0)), "torch_threads": torch.get_num_threads()}
    res["full"] = timed()
    res["no_d2h"] = timed(d2h=False)
    res["no_h2d"] = timed(h2d=False)
    res["copies_only"] = timed(call=False)
    for k in real:
        setattr(draws, k, cached(k))
    res["cached_draws"] = timed()
    res["cached_draws_no_d2h"] = timed(d2h=False)
    for k, v in real.items():
        setattr(draws, k, v)
    res["full_again"] = timed()
    if rank == 0:
        print(json.dumps(res))
        if args.out:
            with open(args.out, "w") as f:
                f.write(json.dumps(res) + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
