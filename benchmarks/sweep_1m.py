#!/usr/bin/env python
"""BASELINE config 4: 1 M synthetic cardiac cycles (245 batches of 4096 x 4 x 2500) through PCGmix+,
batch-sharded over the GPUs of one box.

    python benchmarks/sweep_1m.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29500 benchmarks/sweep_1m.py                  # N GPUs

Batch k is augmented with seed k on rank k mod N (`pcgmix_b200.sharding`), so the result does not
depend on N; there is no collective on the path (NCCL only carries the barrier and the MAX over
ranks).  All of a rank's batches are resident in its HBM (40 GB at N = 1); the signals are drawn on
the device (the values do not matter for the timing, the offsets and labels are drawn like in
bench.py), host draws are uploaded before the timed region, consecutive launches may overlap.
The host's share — pairing, processing order, lambda and knots of the rank's batches, drawn by the native replay
on a few worker threads (it releases the GIL), and their upload — is timed on the wall clock and reported beside
the device time (``host_draws_seconds``, ``seconds_including_host_draws``).  One JSON line per run.
"""
from __future__ import annotations

import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from pcgmix_b200 import augmentations, draws, native, sharding, staging, synth  # noqa: E402

N_BATCHES, B, C, L, KNOT, SIGMA = 245, 4096, 4, 2500, 4, 0.2


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    native.load()
    mine = list(sharding.batches_for_rank(N_BATCHES, rank, world))
    gen = torch.Generator(device=dev)
    data, metas, m_total = [], [], 0
    t = torch.arange(L, device=dev)[None, None, :]
    inputs = []
    for k in mine:
        rng = np.random.default_rng(synth.BENCH_SEED + k)
        frames = synth.cycle_frames(rng, B, limit=L)
        labels = rng.integers(0, 2, B)
        gen.manual_seed(k)
        x = torch.randn((B, C, L), device=dev, generator=gen)
        x *= (t < torch.from_numpy(frames[:, 4]).to(dev)[:, None, None])
        data.append(x)
        inputs.append((frames, labels, sharding.step_seed(k)))
    torch.cuda.synchronize()

    def host_draws(item):                                      # everything the host contributes to one batch
        frames, labels, seed = item
        mix = draws.same_label_pairing(labels, seed)
        lam, knots, _ = native.host_lambda_knots(seed, 1.0, SIGMA, (B, KNOT + 2, C), max_threads=1, want_state=False)
        return mix, draws.processing_order(mix), lam, knots

    workers = max(1, min(8, (os.cpu_count() or 1) // max(world, 1)))
    host_draws(inputs[0])                                       # (library load, thread start-up: not part of the sweep)
    staging.upload([np.zeros(16, np.int32)], dev)
    torch.cuda.synchronize()
    t_host = time.perf_counter()
    with ThreadPoolExecutor(max_workers=workers) as pool:
        drawn = list(pool.map(host_draws, inputs))
    for (frames, labels, seed), (mix, order, lam, knots) in zip(inputs, drawn):
        up = staging.upload([frames.astype(np.int32), mix.astype(np.int32), order, knots], dev)
        metas.append((up, draws.lambda_pair_fp32(lam)))
        m_total += synth.mixed_samples(frames, mix)
    torch.cuda.synchronize()
    host_seconds = time.perf_counter() - t_host
    outs = [torch.empty((B, C, L), device=dev) for _ in range(3)]

    prepared = [augmentations.prepare_on_device(data[i], up[0], up[1], lam[0], lam[1], outs[i % 3], up[3], KNOT,
                                                order_dev=up[2]) for i, (up, lam) in enumerate(metas)]
    handle = torch.cuda.current_stream(dev).cuda_stream

    def run(i):
        prepared[i].launch(handle)

    native.set_launch_overlap(True)
    for i in range(min(3, len(mine))):
        run(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(len(mine)):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    native.set_launch_overlap(False)
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    host_s = torch.tensor([host_seconds], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(host_s, op=dist.ReduceOp.MAX)
    bytes_local = torch.tensor([4.0 * C * (2.0 * L * B * len(mine) + m_total)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(bytes_local, op=dist.ReduceOp.SUM)
    if rank == 0:
        peak = 6544.7
        try:
            peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            pass
        sec = float(ms.item()) * 1e-3
        gbs = float(bytes_local.item()) / sec / 1e9
        print(json.dumps({"config": "cfg4: 1M cycles = 245 batches of 4096 x 4 x 2500, durmixmagwarp(0.2,4)",
                          "n_gpus": world, "cycles": N_BATCHES * B, "seconds": sec, "cycles_per_s": N_BATCHES * B / sec,
                          "aggregate_algorithmic_GBps": gbs, "frac_of_measured_peak_per_gpu": gbs / world / peak,
                          "host_draws_seconds": float(host_s.item()), "host_draw_threads_per_rank": workers,
                          "seconds_including_host_draws": sec + float(host_s.item()),
                          "cycles_per_s_including_host_draws": N_BATCHES * B / (sec + float(host_s.item())),
                          "sharding": "batch k -> rank k mod N, seed k, no collective"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
