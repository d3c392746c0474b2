"""Golden vectors for the first block of the reference's ResNet9-1D (models.py:468-473 ``conv_block``, ``conv1`` at
models.py:523), produced by running the UNMODIFIED reference: ``models.ResNet9(in_channels, num_classes, filters,
linear).conv1`` imported from /root/reference behind stubs for the absent ``tsai`` package (models.py:3-4 imports
it; the ResNet9 classes use nothing from it), in training mode (batch statistics, running statistics updated) and
then in evaluation mode, on zero-padded synthetic cycles.

Run in the build container:  python tests/golden/make_golden_first_block.py
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from pcgmix_b200 import synth  # noqa: E402

REF = os.environ.get("PCGMIX_REFERENCE_ROOT", "/root/reference")

# name, in_channels, filters, linear, batch, length, seed, what is special
CASES = [
    ("first_block_c4_f64", 4, [64, 128, 256, 512], 39936, 3, 500, 11, "the default ResNet9-1D on band-split cycles"),
    ("first_block_c4_f64_odd", 4, [64, 128, 256, 512], 39936, 5, 257, 12, "row length not a multiple of four"),
    ("first_block_c1_f16", 1, [16, 32, 64, 128], 9984, 4, 96, 13, "single channel, the 16-filter variant (train_model.py:348)"),
    ("first_block_c2_f7", 2, [7, 8, 8, 8], 8, 3, 64, 14, "odd filter count, two channels"),
]


def load_reference_models():
    saved = {}
    names = ("tsai", "tsai.models", "tsai.models.layers")
    for name in names:
        saved[name] = sys.modules.get(name)
        mod = types.ModuleType(name)
        if name.endswith("layers"):
            for attr in ("ConvBlock", "Add", "BN1d", "Squeeze", "ConvBN", "Conv1d", "Concat", "GAP1d"):
                setattr(mod, attr, object)
        sys.modules[name] = mod
    try:
        spec = importlib.util.spec_from_file_location("_pcgmix_ref_models", os.path.join(REF, "models.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for name in names:
            if saved[name] is None:
                sys.modules.pop(name, None)
            else:
                sys.modules[name] = saved[name]
    return mod


def main():
    models = load_reference_models()
    torch.set_num_threads(1)
    for name, cin, filters, linear, batch, length, seed, note in CASES:
        torch.manual_seed(seed)
        model = models.ResNet9(in_channels=cin, num_classes=2, filters=filters, linear=linear)
        block = model.conv1
        conv, bn = block[0], block[1]
        rng = np.random.default_rng(seed)
        with torch.no_grad():                       # a block in the middle of training: nothing at its initial value
            bn.weight.copy_(torch.from_numpy(rng.uniform(0.5, 1.5, bn.num_features).astype(np.float32)))
            bn.weight[1] = -0.75                    # a negative scale must survive the folding
            bn.bias.copy_(torch.from_numpy(rng.normal(0, 0.3, bn.num_features).astype(np.float32)))
            bn.running_mean.copy_(torch.from_numpy(rng.normal(0, 0.2, bn.num_features).astype(np.float32)))
            bn.running_var.copy_(torch.from_numpy(rng.uniform(0.5, 2.0, bn.num_features).astype(np.float32)))
        frames = synth.cycle_frames(rng, batch, limit=length)
        x = synth.cycle_signals(rng, frames, (cin,), length)
        x[0, :, 0] += 1.5                           # the row borders carry weight
        x[-1, :, length - 1] -= 2.0
        store = dict(entry=np.array(f"models.ResNet9(in_channels={cin}, num_classes=2, filters={filters}, linear={linear}).conv1 — {note}"),
                     x=x, weight=conv.weight.detach().numpy().copy(), bias=conv.bias.detach().numpy().copy(),
                     gamma=bn.weight.detach().numpy().copy(), beta=bn.bias.detach().numpy().copy(),
                     running_mean_before=bn.running_mean.numpy().copy(), running_var_before=bn.running_var.numpy().copy(),
                     eps=np.float64(bn.eps), momentum=np.float64(bn.momentum))
        with torch.no_grad():
            block.train()
            store["out_train"] = block(torch.from_numpy(x)).numpy().copy()
            store["running_mean_after"] = bn.running_mean.numpy().copy()
            store["running_var_after"] = bn.running_var.numpy().copy()
            store["num_batches_tracked_after"] = np.int64(bn.num_batches_tracked.item())
            block.eval()
            store["out_eval"] = block(torch.from_numpy(x)).numpy().copy()      # with the UPDATED running statistics
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **store)
        print(name, store["out_train"].shape, "zeros after ReLU:", float((store["out_train"] == 0).mean()))


if __name__ == "__main__":
    main()
