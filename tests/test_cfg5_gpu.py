"""BASELINE config 5 (the DDP training step fed by the on-device PCGmix+, examples/train_ddp_pcgmix.py) runs, at
one GPU in-process and — where the box has two — at two ranks under torchrun with NCCL."""
import json
import math
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
pytestmark = pytest.mark.gpu


def _example():
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    import train_ddp_pcgmix
    return train_ddp_pcgmix


@pytest.mark.parametrize("use_resident", [False, True])
@pytest.mark.parametrize("device_labels", [False, True])
def test_training_step_runs_on_one_gpu(use_resident, device_labels):
    res = _example().run_cfg5(steps=8, batch=64, use_resident=use_resident, device_labels=device_labels)
    assert res["n_gpus"] == 1 and res["per_rank_batch"] == 64 and math.isfinite(res["loss"])
    assert 0 < res["augment_call_host_ms_median"] < 5 and res["cycles_per_s"] > 0


def test_augment_inside_the_step_matches_the_oracle():
    """What the model is fed in that loop: the step counter is the seed, every step a fresh pairing / lambda / knots."""
    import numpy as np
    from oracle import pcgmix_oracle as orc
    from pcgmix_b200 import augmentations, synth
    ex = _example()
    rng = np.random.default_rng(3)
    frames = synth.cycle_frames(rng, 64, limit=2500)
    data = synth.cycle_signals(rng, frames, (4,), 2500)
    target = torch.from_numpy(rng.integers(0, 2, 64))
    dev = torch.device("cuda:0")
    counter = ex.StepCounter()
    args = ex.Args()
    for _ in range(3):
        ohe = augmentations.with_host_labels(torch.nn.functional.one_hot(target, 2).to(dev), target)
        out, _, mix, _ = augmentations.augment(args, torch.from_numpy(data).to(dev), ohe, torch.from_numpy(frames), ["a"] * 64,
                                               counter, None, dev, None)
        want, want_mix, _, _ = orc.augment_1d(args.method, data, target.numpy(), frames, counter.count)
        assert np.array_equal(mix, want_mix)
        rel = np.abs(out.cpu().numpy().astype(np.float64) - want) / np.maximum(np.abs(want), np.finfo(np.float32).tiny)
        assert rel.max() <= 1e-5
        counter.add()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("extra", [[], ["--resident"]])
def test_two_ranks_under_torchrun(extra):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29877", os.path.join(ROOT, "examples", "train_ddp_pcgmix.py"), "--steps", "8"] + extra
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert proc.returncode == 0, proc.stderr[-2000:]
    line = json.loads([ln for ln in proc.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["n_gpus"] == 2 and math.isfinite(line["loss"]) and line["cycles_per_s"] > 0
