#!/usr/bin/env python
"""BASELINE config 5: on-device PCGmix+ feeding a 1D ResNet training step, data-parallel over the
GPUs of one box (one process per GPU, NCCL gradient all-reduce through DistributedDataParallel).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29500 examples/train_ddp_pcgmix.py --steps 50

What is shown: the reference's per-step call `augmentations.augment(args, data, target_ohe, frames,
wav, step_counter, model, device, EXPERIMENT_ARGS)` (train_model.py:507) served by the B200 kernels,
with every rank owning its mini-batch and its `step_counter` (all ranks use seed = step, as a
single-GPU run of the reference would), and `nn.DataParallel` (train_model.py:385) replaced by DDP.
The network has the reference ResNet9-1D's shape (models.py:520-589: 2 274 626 parameters, 9.1 MB of
fp32 gradients all-reduced per step); the optimiser step is the reference's (Adam + OneCycleLR +
gradient value clipping, train_model.py:405-409, 555-569).  Data is synthetic.  Rank 0 prints one
JSON line: step time, the share of the on-device PCGmix+ call (host wall clock and device time) and
whole-job cycles/s.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from pcgmix_b200 import augmentations, resident, synth  # noqa: E402


# (in, out, halve) per convolution stage; a residual pair follows stages 1 and 3.  Widths, the three
# halvings, the final pool of 4 and the 39 936-wide classifier input are those of the reference's
# ResNet9 for (4, 2500) inputs (models.py:468-473, 520-531, 589): 2 274 626 parameters.
STAGES = ((4, 64, False), (64, 128, True), (128, 256, True), (256, 512, True))
RESIDUAL_AFTER = (1, 3)
PARAMETERS = 2_274_626


def conv_unit(cin, cout, halve):
    unit = [nn.Conv1d(cin, cout, 3, padding=1), nn.BatchNorm1d(cout), nn.ReLU(inplace=True)]
    return nn.Sequential(*unit, nn.MaxPool1d(2)) if halve else nn.Sequential(*unit)


class CycleResNet9(nn.Module):
    """Consumer of the augmented cycles with the reference ResNet9-1D's shape and size."""

    def __init__(self, classes=2, length=2500):
        super().__init__()
        self.stages = nn.ModuleList(conv_unit(*spec) for spec in STAGES)
        self.residuals = nn.ModuleDict({str(i): nn.Sequential(conv_unit(STAGES[i][1], STAGES[i][1], False),
                                                              conv_unit(STAGES[i][1], STAGES[i][1], False))
                                        for i in RESIDUAL_AFTER})
        for _, _, halve in STAGES:
            length = length // 2 if halve else length
        self.head = nn.Sequential(nn.MaxPool1d(4), nn.Flatten(), nn.Linear(STAGES[-1][1] * (length // 4), classes))

    def forward(self, x):
        for i, stage in enumerate(self.stages):
            x = stage(x)
            if str(i) in self.residuals:
                x = self.residuals[str(i)](x) + x
        return self.head(x)


class StepCounter:
    def __init__(self):
        self.count = 0

    def add(self):
        self.count += 1


class Args:
    method = "durmixmagwarp(0.2,4)"
    batch_size = 64
    sample_rate = 1000
    num_classes = 2


def run_cfg5(steps=30, batch=64, use_resident=False, device_labels=False, init_process_group=True, sync_every_step=False):
    """Train for ``steps`` steps; returns the result dict on rank 0 (None elsewhere).  With
    ``init_process_group=False`` the caller (bench.py) already owns an NCCL process group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and init_process_group:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(4)
    model = CycleResNet9().to(dev)
    assert sum(p.numel() for p in model.parameters()) == PARAMETERS
    if world > 1:
        model = nn.parallel.DistributedDataParallel(model, device_ids=[local])
    optim = torch.optim.Adam(model.parameters(), lr=1e-3)
    sched = torch.optim.lr_scheduler.OneCycleLR(optim, max_lr=1e-3, total_steps=steps)
    args = Args()
    args.batch_size = batch
    counter = StepCounter()
    rng = np.random.default_rng(100 + rank)                    # every rank draws its own cycles
    wav = ["a0001"] * batch
    # a few host batches prepared up front (what the loader's workers would have ready)
    pool = []
    for _ in range(4):
        frames = synth.cycle_frames(rng, batch, limit=2500)
        pool.append((torch.from_numpy(synth.cycle_signals(rng, frames, (4,), 2500)).pin_memory(),
                     torch.from_numpy(frames), torch.from_numpy(rng.integers(0, 2, batch))))
    res = cycle_labels = None
    if use_resident:
        # this rank's recordings (64 x 4 bands x 40 s @ 1 kHz) with dense Springer states -> cycle table, once
        states = torch.from_numpy(synth.dense_states(rng, 64, 40000, 1000)).to(dev)
        signal = torch.from_numpy(rng.standard_normal((64, 4, 40000)).astype(np.float32)).to(dev)
        res = resident.from_dense_states(signal, states, 2500)
        cycle_labels = rng.integers(0, 2, res.n_cycles)
    # Timing: like a real loop, nothing waits for the GPU inside a step; the host runs at most two steps ahead (it waits
    # for step k-2 at the top of step k, as a loader / logging call would make it), the device is synchronised once
    # after the warm-up steps and once at the end, and the step time is the wall clock of the steps in between divided
    # by their number.  ``sync_every_step`` restores a synchronisation per step.
    warm = min(5, max(0, steps - 3))
    aug_events, aug_host_ms, step_done = [], [], []
    loss = None
    t_begin = None
    for step in range(steps):
        if step >= 2 and not sync_every_step:
            step_done[step - 2].synchronize()                  # the host stays at most two steps ahead of the device
        if step == warm:
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t_begin = time.perf_counter()
        host_data, frames, target = pool[step % len(pool)]
        if use_resident:
            ids = rng.integers(0, res.n_cycles, batch)          # what a sampler over the cycle table yields
            target = torch.from_numpy(cycle_labels[ids])
        if not use_resident:
            data = host_data.to(dev, non_blocking=True)            # train_model.py:499
        target_ohe = F.one_hot(target, args.num_classes).to(dev)
        if not device_labels:                                  # the loader's CPU target: pairing needs no device read-back
            augmentations.with_host_labels(target_ohe, target)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        h0 = time.perf_counter()
        if use_resident:
            data, target_ohe, _, _ = resident.augment(args, res, ids, target_ohe, wav, counter, model, dev, None)
        else:
            data, target_ohe, _, _ = augmentations.augment(args, data, target_ohe, frames, wav, counter, model, dev, None)
        if step >= warm:
            aug_host_ms.append((time.perf_counter() - h0) * 1e3)
        e1.record()
        aug_events.append((e0, e1))
        loss = F.cross_entropy(model(data), target_ohe.float().argmax(1))
        loss.backward()                                        # DDP all-reduces the gradients over NCCL here
        nn.utils.clip_grad_value_(model.parameters(), 0.1)
        optim.step()
        optim.zero_grad(set_to_none=True)
        sched.step()
        counter.add()
        done = torch.cuda.Event()
        done.record()
        step_done.append(done)
        if sync_every_step:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    step_ms_mean = (time.perf_counter() - t_begin) * 1e3 / max(1, steps - warm)
    aug_dev_ms = [a.elapsed_time(b) for a, b in aug_events[warm:]]
    t = torch.tensor([step_ms_mean], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    result = None
    if rank == 0:
        ms = float(t.item())
        result = {"config": "cfg5: on-device PCGmix+ -> ResNet9-1D training step, DDP" +
                  (", batches drawn from resident recordings" if use_resident else ""), "n_gpus": world,
                  "per_rank_batch": batch, "steps": steps, "loss": round(loss.item(), 4),
                  "step_ms_max_over_ranks": ms, "augment_call_host_ms_median": float(np.median(aug_host_ms)),
                  "augment_device_span_ms_median": float(np.median(aug_dev_ms)),
                  "timing": ("device synchronised every step" if sync_every_step else
                             "wall clock of steps %d..%d / their number, one synchronisation at each end" % (warm, steps - 1)),
                  "labels": "device one-hot read back every step (reference behaviour)" if device_labels else "loader's CPU target",
                  "cycles_per_s": world * batch / (ms * 1e-3)}
    if world > 1 and init_process_group:
        dist.destroy_process_group()
    return result


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--batch", type=int, default=64, help="per-rank batch (the reference trains with 64)")
    ap.add_argument("--resident", action="store_true",
                    help="keep the rank's recordings + cycle table on the GPU and draw batches as table rows "
                         "(pcgmix_b200.resident): no per-step upload, no padded array")
    ap.add_argument("--device-labels", action="store_true",
                    help="recover the class ids from the device one-hot tensor every step, like the reference "
                         "(augmentations.py:501), instead of taking them from the loader's CPU target")
    ap.add_argument("--sync-every-step", action="store_true", help="synchronise the device after every step (round-1 timing)")
    opt = ap.parse_args()
    result = run_cfg5(opt.steps, opt.batch, opt.resident, opt.device_labels, sync_every_step=opt.sync_every_step)
    if result is not None:
        import json
        print(json.dumps(result))


if __name__ == "__main__":
    main()
