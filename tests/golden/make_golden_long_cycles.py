"""Golden fixtures for cycles whose offsets run past the padded row length, from the UNMODIFIED
reference (run in the build container: ``python tests/golden/make_golden_long_cycles.py``).

The data builder keeps such cycles (databuilder.ipynb cells 14/25 only print "segment ... longer than
... samples"; ``resize`` truncates the samples and the offsets stay), and the reference's slices are
clamped by Python, so a long cycle still blends with a partner whenever the clamped widths agree.
Stored here:

  long_pairs_1d   every ordered pair of eight offset lists (normal, long diastole, S2 across the row
                  end, everything beyond the row, one-sample clamps) through
                  ``augmentations.mixup_keepdur_multidim_tensors``: ``ok[i, j]`` says whether the
                  reference returned or raised, ``out[i, j]`` what it returned
  long_pairs_2d   the same through ``augmentations2d.mixup_keepdur_multidim_tensors``
  long_batch_1d   ``augmentations.augment`` (PCGmix and PCGmix+) on a batch that contains long cycles,
                  at a step whose pairing the reference can blend, plus a step at which it raises

Kept apart from make_golden.py so that the seeded stream behind the older fixtures is unchanged.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle.ref_import import load_reference  # noqa: E402
sys.path.insert(0, HERE)
from make_golden import _Args, _Step, cycle_frames, save  # noqa: E402


def main():
    ref1, ref2 = load_reference()
    rng = np.random.default_rng(20261018)

    length = 50
    frames = np.array([
        [0, 10, 20, 30, 40],      # inside the row
        [0, 10, 20, 30, 70],      # long diastole
        [0, 12, 24, 36, 50],      # ends exactly at the row end
        [0, 10, 20, 55, 60],      # S2 crosses the row end, diastole beyond it
        [0, 60, 70, 80, 90],      # S1 covers the whole row, the rest lies beyond
        [0, 10, 20, 49, 80],      # diastole clamps to one sample
        [0, 5, 10, 15, 120],      # very long diastole
        [2, 9, 21, 33, 52],       # f[0] != 0, ends just past the row
    ], dtype=np.int64)
    n = frames.shape[0]
    lam = 0.35
    for tag, ref, shape in (("1d", ref1, (3,)), ("2d", ref2, (1, 5))):
        x = rng.standard_normal((n,) + shape + (length,)).astype(np.float32)
        out = np.zeros((n, n) + shape + (length,), np.float32)
        ok = np.zeros((n, n), bool)
        lam_t = torch.from_numpy(np.full((1,) * (len(shape) + 0), lam, np.float32).reshape((1,) * len(shape)))
        for i in range(n):
            for j in range(n):
                try:
                    out[i, j] = ref.mixup_keepdur_multidim_tensors(
                        torch.from_numpy(x[i].copy()), torch.from_numpy(x[j].copy()), frames[i], frames[j], lam_t,
                        "durratiomixup", 0).numpy()
                    ok[i, j] = True
                except RuntimeError:
                    pass
        save(f"long_pairs_{tag}", entry=np.array(f"augmentations{'2d' if tag == '2d' else ''}.mixup_keepdur_multidim_tensors"),
             data=x, frames=frames, lam=np.float32(lam), out=out, ok=ok)
        print(tag, "pairs the reference blends:", int(ok.sum()), "of", n * n)

    # dispatcher level: 12 cycles, 4 of them longer than the row (samples fill the whole row)
    b, c, length = 12, 2, 600
    fr = cycle_frames(rng, b, limit=length)
    for k in (1, 4, 6, 9):
        fr[k, 4] = length + int(rng.integers(20, 400))         # diastole runs past the row end
    x = rng.standard_normal((b, c, length)).astype(np.float32)
    x *= (np.arange(length)[None, None, :] < fr[:, 4][:, None, None])
    lab = rng.integers(0, 2, b).astype(np.int64)

    def run(method, step):
        ohe = torch.nn.functional.one_hot(torch.from_numpy(lab), 2)
        out, _, mix, _ = ref1.augment(_Args(method, b), torch.from_numpy(x.copy()), ohe, torch.from_numpy(fr),
                                      ["a0001"] * b, _Step(step), None, "cpu", None)
        return out.numpy(), np.asarray(mix, np.int64)

    good = bad = None
    for step in range(400):
        try:
            run("durratiomixup", step)
            if good is None:
                mix = run("durratiomixup", step)[1]
                if any(fr[i, 4] > length or fr[mix[i], 4] > length for i in range(b) if mix[i] != i):
                    good = step
        except RuntimeError:
            if bad is None:
                bad = step
        if good is not None and bad is not None:
            break
    assert good is not None and bad is not None, (good, bad)
    out_mix, mix = run("durratiomixup", good)
    out_plus, mix_plus = run("durmixmagwarp(0.2,4)", good)
    assert np.array_equal(mix, mix_plus)
    save("long_batch_1d", entry=np.array("augmentations.augment"), step=np.int64(good), step_raises=np.int64(bad),
         data=x, labels=lab, frames=fr, mix=mix, out_durratiomixup=out_mix, out_durmixmagwarp=out_plus)
    print("dispatcher: blends at step", good, "raises at step", bad)


if __name__ == "__main__":
    main()
