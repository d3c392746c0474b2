"""Two calls of the first block (batch statistics, then running statistics) on BASELINE's batch — the command
`ncu --set full` is pointed at (scripts/gpu_first_block_ncu.sh).  Never a source of timings."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pcgmix_b200 import first_block, synth  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(synth.BENCH_SEED + 6)
    B, C, L, F = 4096, 4, 2500, 64
    frames = synth.cycle_frames(rng, 512, limit=L)
    data = torch.from_numpy(synth.cycle_signals(rng, frames, (C,), L)).to(dev).repeat(B // 512, 1, 1).contiguous()
    torch.manual_seed(6)
    block = torch.nn.Sequential(torch.nn.Conv1d(C, F, 3, padding=1), torch.nn.BatchNorm1d(F), torch.nn.ReLU(inplace=True)).to(dev)
    out = torch.empty((B, F, L), dtype=torch.float32, device=dev)
    for training in (True, False, True, False):
        block.train(training)
        first_block.first_conv_block(block, data, out=out)
    torch.cuda.synchronize()
    print("ok", float(out[0, 0, :8].sum()))


if __name__ == "__main__":
    main()
