#!/usr/bin/env python
"""BASELINE config 5: on-device PCGmix+ feeding a 1D ResNet training step, data-parallel over the
GPUs of one box (one process per GPU, NCCL gradient all-reduce through DistributedDataParallel).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29500 examples/train_ddp_pcgmix.py --steps 50

What is shown: the reference's per-step call `augmentations.augment(args, data, target_ohe, frames,
wav, step_counter, model, device, EXPERIMENT_ARGS)` (train_model.py:507) served by the B200 kernels,
with every rank owning its mini-batch and its `step_counter` (all ranks use seed = step, as a
single-GPU run of the reference would), and `nn.DataParallel` (train_model.py:385) replaced by DDP.
The network is a compact stand-in with the reference ResNet9-1D's input contract (B, 4, 2500) -> 2
logits; the reference's own `models.ResNet9` can be dropped in unchanged.  Data is synthetic.
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
from pcgmix_b200 import augmentations, synth  # noqa: E402


def block(cin, cout, pool):
    layers = [nn.Conv1d(cin, cout, 3, padding=1, bias=False), nn.BatchNorm1d(cout), nn.ReLU(inplace=True)]
    if pool:
        layers.append(nn.MaxPool1d(pool))
    return nn.Sequential(*layers)


class SmallResNet1D(nn.Module):
    def __init__(self, channels=4, classes=2):
        super().__init__()
        self.c1, self.c2 = block(channels, 64, 0), block(64, 128, 2)
        self.r1 = nn.Sequential(block(128, 128, 0), block(128, 128, 0))
        self.c3, self.c4 = block(128, 256, 2), block(256, 512, 2)
        self.r2 = nn.Sequential(block(512, 512, 0), block(512, 512, 0))
        self.head = nn.Linear(512, classes)

    def forward(self, x):
        x = self.c2(self.c1(x))
        x = x + self.r1(x)
        x = self.c4(self.c3(x))
        x = x + self.r2(x)
        return self.head(F.adaptive_max_pool1d(x, 1).flatten(1))


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--batch", type=int, default=64, help="per-rank batch (the reference trains with 64)")
    opt = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(4)
    model = SmallResNet1D().to(dev)
    if world > 1:
        model = nn.parallel.DistributedDataParallel(model, device_ids=[local])
    optim = torch.optim.Adam(model.parameters(), lr=1e-3)
    args = Args()
    args.batch_size = opt.batch
    counter = StepCounter()
    rng = np.random.default_rng(100 + rank)                    # every rank draws its own cycles
    wav = ["a0001"] * opt.batch
    aug_ms, step_ms = [], []
    for step in range(opt.steps):
        frames = synth.cycle_frames(rng, opt.batch, limit=2500)
        data = torch.from_numpy(synth.cycle_signals(rng, frames, (4,), 2500)).pin_memory()
        target = torch.from_numpy(rng.integers(0, 2, opt.batch))
        t0 = time.perf_counter()
        data = data.to(dev, non_blocking=True)
        target_ohe = F.one_hot(target, args.num_classes).to(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        data, target_ohe, _, _ = augmentations.augment(args, data, target_ohe, torch.from_numpy(frames), wav, counter,
                                                       model, dev, None)
        e1.record()
        loss = F.cross_entropy(model(data), target_ohe.float().argmax(1))
        optim.zero_grad(set_to_none=True)
        loss.backward()                                        # DDP all-reduces the gradients over NCCL here
        optim.step()
        counter.add()
        torch.cuda.synchronize()
        step_ms.append((time.perf_counter() - t0) * 1e3)
        aug_ms.append(e0.elapsed_time(e1))
    if rank == 0:
        print(f"ranks={world} per-rank batch={opt.batch} steps={opt.steps} loss={loss.item():.4f} "
              f"median step {np.median(step_ms[3:]):.2f} ms, of which on-device PCGmix+ {np.median(aug_ms[3:]):.3f} ms "
              f"(cycles/s over all ranks: {world * opt.batch / (np.median(step_ms[3:]) * 1e-3):.0f})")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
