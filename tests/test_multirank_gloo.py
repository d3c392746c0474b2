"""World-size-2 check (gloo, CPU) of the multi-GPU scheme: batches are sharded round-robin, every
rank replays the reference's draws for ITS batches with seed = global batch index, no collective
is needed on the augmentation path, and the union over ranks equals the single-process result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pcgmix_b200 import draws, sharding, synth

N_BATCHES = 7
BATCH = 64
METHOD = "durmixmagwarp(0.2,4)"


def _digest(batch_index: int):
    """Everything the device kernel consumes for one batch, reduced to a few numbers."""
    rng = np.random.default_rng(1000 + batch_index)
    frames = synth.cycle_frames(rng, BATCH, limit=2500)
    labels = rng.integers(0, 2, BATCH)
    seed = sharding.step_seed(batch_index)
    plan = draws.parse_method_1d(METHOD)
    mix = draws.pairing(METHOD, labels, None, seed)
    lam32, _ = draws.lambda_pair_fp32(draws.draw_lambda(plan.alpha, seed))
    knots = draws.draw_knots(BATCH, plan.knot, 4, plan.sigma)
    order = draws.processing_order(mix)
    return np.array([batch_index, float(lam32), float(knots.sum()), float((mix * np.arange(BATCH)).sum()),
                     float((order * np.arange(BATCH)).sum()), float(synth.mixed_samples(frames, mix))])


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = [_digest(k) for k in sharding.batches_for_rank(N_BATCHES, rank, world)]
    local = torch.zeros(N_BATCHES, 6, dtype=torch.float64)
    for row in mine:
        local[int(row[0])] = torch.from_numpy(row)
    # the only cross-rank traffic of the benchmark: a barrier and a MAX over per-rank times
    dist.barrier()
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == 10.0 + world - 1
    dist.all_reduce(local, op=dist.ReduceOp.SUM)      # test-only gather of the digests
    if rank == 0:
        np.save(os.path.join(out_dir, "union.npy"), local.numpy())
    dist.destroy_process_group()


def test_two_ranks_cover_all_batches_with_global_seeds(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    union = np.load(tmp_path / "union.npy")
    single = np.stack([_digest(k) for k in range(N_BATCHES)])
    assert np.array_equal(union, single)
