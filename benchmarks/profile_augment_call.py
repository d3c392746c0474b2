#!/usr/bin/env python
"""Where the host time of one drop-in ``augment`` call goes at the reference's batch size (B = 64):
cProfile over a few hundred calls, plus wall time per call with and without a device sync."""
import cProfile
import io
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from pcgmix_b200 import augmentations, synth  # noqa: E402


class Args:
    method = "durmixmagwarp(0.2,4)"
    batch_size = 64
    sample_rate = 1000
    num_classes = 2


class Step:
    count = 0


def main():
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(0)
    b = 64
    frames = synth.cycle_frames(rng, b, limit=2500)
    data = torch.from_numpy(synth.cycle_signals(rng, frames, (4,), 2500)).to(dev)
    target = torch.from_numpy(rng.integers(0, 2, b))
    ohe = torch.nn.functional.one_hot(target, 2).to(dev)
    if "--host-labels" in sys.argv:                      # the loader's CPU target handed over: no device read-back
        augmentations.with_host_labels(ohe, target)
    ft = torch.from_numpy(frames)
    wav = ["a"] * b
    step = Step()

    def call():
        step.count += 1
        return augmentations.augment(Args, data, ohe, ft, wav, step, None, dev, None)

    for _ in range(20):
        call()
    torch.cuda.synchronize()
    n = 300
    t0 = time.perf_counter()
    for _ in range(n):
        call()
    t_async = (time.perf_counter() - t0) / n
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        call()
        torch.cuda.synchronize()
    t_sync = (time.perf_counter() - t0) / n
    print(f"augment() at B=64: {t_async * 1e6:.0f} us per call back to back, {t_sync * 1e6:.0f} us with a sync after each")
    prof = cProfile.Profile()
    prof.enable()
    for _ in range(n):
        call()
    prof.disable()
    torch.cuda.synchronize()
    out = io.StringIO()
    pstats.Stats(prof, stream=out).sort_stats("cumulative").print_stats(28)
    print(out.getvalue()[:6000])


if __name__ == "__main__":
    main()
