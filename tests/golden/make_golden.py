"""Generate the golden fixtures in this directory from the UNMODIFIED reference code.

Run inside the build container (where /root/reference exists):

    python tests/golden/make_golden.py

It imports ``/root/reference/augmentations.py`` and ``augmentations2d.py`` through the stub
recipe in ``oracle/ref_import.py``, feeds them seeded inputs and stores inputs + outputs as
``.npz`` files.  The tests never need the reference tree: they read the fixtures.

Every fixture records which reference entry point produced it (``entry``), so a reader can tell
a dispatcher-level vector (``augment``) from a per-pair one (``mixup_keepdur_multidim_tensors``).
"""
from __future__ import annotations

import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle.ref_import import load_reference  # noqa: E402


class _Args:
    def __init__(self, method, batch_size):
        self.method = method
        self.batch_size = batch_size
        self.sample_rate = 1000
        self.num_classes = 2


class _Step:
    def __init__(self, count):
        self.count = count


def cycle_frames(rng, n, fs=1000, limit=None):
    """Four state durations per cycle in the physiological ranges of SURVEY.md 8d."""
    lo = np.array([90, 150, 70, 300])
    hi = np.array([160, 400, 130, 900])
    ms = rng.integers(lo, hi + 1, size=(n, 4))
    dur = (ms * fs // 1000).astype(np.int64)
    fr = np.concatenate([np.zeros((n, 1), np.int64), np.cumsum(dur, axis=1)], axis=1)
    if limit is not None:
        scale = np.minimum(1.0, (limit - 1) / fr[:, 4:5])
        fr = np.floor(fr * scale).astype(np.int64)
    return fr


def signals(rng, frames, channels, length):
    x = rng.standard_normal((frames.shape[0],) + tuple(channels) + (length,)).astype(np.float32)
    t = np.arange(length)[None, :]
    keep = (t < frames[:, 4:5]).reshape((frames.shape[0],) + (1,) * len(channels) + (length,))
    return (x * keep).astype(np.float32)


def run_1d(ref, method, step, data, labels, frames, wav=None):
    args = _Args(method, data.shape[0])
    ohe = torch.nn.functional.one_hot(torch.from_numpy(labels), 2)
    wav = wav if wav is not None else ["a0001"] * data.shape[0]
    d_in = torch.from_numpy(data.copy())
    out, tgt, mix, _ = ref.augment(args, d_in, ohe, torch.from_numpy(frames), wav, _Step(step), None, "cpu", None)
    same_obj = out is d_in
    return out.numpy(), np.asarray(tgt.numpy()), np.asarray(mix, dtype=np.int64), same_obj


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def main():
    ref1, ref2 = load_reference()
    rng = np.random.default_rng(20241018)

    # ---- known-answer scalars for the host draw chain -------------------------------------
    np.random.seed(7)
    beta7 = np.random.beta(1, 1)
    normals7 = np.random.normal(1, 0.2, (2, 6, 4))
    save("kat_draws",
         entry=np.array("augmentations.get_lambda / random.Random / np.random legacy stream"),
         lambda_alpha1_seed7=np.float64(ref1.get_lambda(1, 7)),
         lambda_alpha05_seed3=np.float64(ref1.get_lambda(0.5, 3)),
         lambda_alpha0=np.float64(ref1.get_lambda(0, 3)),
         beta7=np.float64(beta7), normals7=normals7,
         uniform7=np.float64(random.Random(7).uniform(0, 1)),
         sample7=np.array(random.Random(7).sample(list(range(10)), 10)),
         labels=np.array([0, 1, 1, 0, 1, 0, 0, 1, 1, 1, 0, 2]),
         same_label_mix_seed5=ref1.get_same_label_mix_indices(
             torch.nn.functional.one_hot(torch.tensor([0, 1, 1, 0, 1, 0, 0, 1, 1, 1, 0, 2]), 3), 5))

    # ---- 1D dispatcher-level vectors -------------------------------------------------------
    cases_1d = [
        ("pcgmix_c4_l2500", "durratiomixup", 3, 6, 4, 2500),
        ("pcgmixplus_c4_l2500", "durmixmagwarp(0.2,4)", 11, 6, 4, 2500),
        ("pcgmixplus_default_c2_l800", "durmixmagwarp", 2, 5, 2, 800),
        ("pcgmixplus_alpha_k2_oddlen", "(alpha=0.5)durmixmagwarp(0.3,2)+0.9", 4, 7, 1, 1001),
        ("pcgmixplus_k7_c3", "durmixmagwarp(0.1,7)", 9, 4, 3, 1203),
        ("pcgmix_alpha2_prob", "(alpha=2.0)durratiomixup+0.8", 6, 9, 3, 1502),
        ("pcgmix_mixall", "(mixAll)durratiomixup", 8, 8, 2, 1200),
        ("pcgmixplus_mixall", "(mixAll)durmixmagwarp(0.2,4)", 8, 8, 2, 1200),
    ]
    for name, method, step, b, c, length in cases_1d:
        fr = cycle_frames(rng, b, limit=length)
        x = signals(rng, fr, (c,), length)
        lab = rng.integers(0, 2, b).astype(np.int64)
        out, tgt, mix, same = run_1d(ref1, method, step, x, lab, fr)
        assert not same, (name, "gate unexpectedly failed; pick another step")
        save(name, entry=np.array("augmentations.augment"), method=np.array(method), step=np.int64(step),
             data=x, labels=lab, frames=fr, out=out, target=tgt, mix=mix)

    # pairing modifiers that use the recording names
    b, c, length = 10, 2, 900
    fr = cycle_frames(rng, b, limit=length)
    x = signals(rng, fr, (c,), length)
    lab = rng.integers(0, 2, b).astype(np.int64)
    wav = ["a0001", "a0001", "b0002", "b0002", "a0001", "c0003", "b0002", "a0007", "c0003", "a0007"]
    for tag in ("samePCG", "sameDataset"):
        method = f"({tag})durratiomixup"
        out, tgt, mix, same = run_1d(ref1, method, 13, x, lab, fr, wav)
        save(f"pcgmix_{tag.lower()}", entry=np.array("augmentations.augment"), method=np.array(method),
             step=np.int64(13), data=x, labels=lab, frames=fr, out=out, target=tgt, mix=mix,
             wav=np.array(wav))

    # gate failure: same objects back, empty pairing
    b, c, length = 4, 1, 64
    fr = cycle_frames(rng, b, limit=length)
    x = signals(rng, fr, (c,), length)
    lab = np.zeros(b, np.int64)
    step = next(s for s in range(100) if random.Random(s).uniform(0, 1) >= 0.1)
    out, tgt, mix, same = run_1d(ref1, "durratiomixup+0.1", step, x, lab, fr)
    assert same and len(mix) == 0
    save("pcgmix_gate_fail", entry=np.array("augmentations.augment"), method=np.array("durratiomixup+0.1"),
         step=np.int64(step), data=x, labels=lab, frames=fr, out=out, mix=mix, same_object=np.bool_(same))

    # ---- 1D edge cases at the per-pair level -----------------------------------------------
    length = 50
    edge_frames = np.array([
        [0, 10, 20, 30, 40],      # plain
        [0, 0, 25, 25, 50],       # zero-length S1 and S2, f[4] == L
        [0, 13, 14, 33, 47],      # odd offsets
        [0, 5, 9, 30, 31],        # short tail state
        [3, 11, 22, 35, 49],      # f[0] != 0
        [0, 12, 12, 12, 12],      # everything after S1 empty
    ], dtype=np.int64)
    n = edge_frames.shape[0]
    x = rng.standard_normal((n, 3, length)).astype(np.float32)
    x[1, :, 7] = np.float32(-0.0)
    pair_out = np.zeros((n, n, 3, length), np.float32)
    lam = torch.from_numpy(np.full((1, 1), 0.3, np.float32))
    for i in range(n):
        for j in range(n):
            pair_out[i, j] = ref1.mixup_keepdur_multidim_tensors(
                torch.from_numpy(x[i].copy()), torch.from_numpy(x[j].copy()),
                edge_frames[i], edge_frames[j], lam, "durratiomixup", 0).numpy()
    save("pairs_1d_edge", entry=np.array("augmentations.mixup_keepdur_multidim_tensors"),
         data=x, frames=edge_frames, lam=np.float32(0.3), out=pair_out)

    # ---- magnitude warp alone --------------------------------------------------------------
    xb = rng.standard_normal((3, 257, 2)).astype(np.float32)
    np.random.seed(21)
    state = np.random.get_state()
    warped = ref1.magnitude_warp(xb, 0.2, 4)
    np.random.set_state(state)
    knots = np.random.normal(1.0, 0.2, (3, 6, 2))
    save("magwarp_alone", entry=np.array("augmentations.magnitude_warp"), data_blc=xb, knots=knots,
         sigma=np.float64(0.2), knot=np.int64(4), out_blc=warped)

    # ---- 2D ---------------------------------------------------------------------------------
    class _StepC(_Step):
        pass

    def run_2d(method, step, data, labels, frames):
        args = _Args(method, data.shape[0])
        ohe = torch.nn.functional.one_hot(torch.from_numpy(labels), 2)
        out, _, mix, _ = ref2.augment(args, torch.from_numpy(data.copy()), ohe, torch.from_numpy(frames),
                                      ["a"] * data.shape[0], _StepC(step), None, "cpu", None)
        return out.numpy(), np.asarray(mix, dtype=np.int64)

    b, f_bins, t_cols = 6, 32, 32
    fr = cycle_frames(rng, b, limit=2200)
    fr = np.round(fr * (t_cols / 2200)).astype(np.int64)
    x = signals(rng, fr, (1, f_bins), t_cols)
    lab = rng.integers(0, 2, b).astype(np.int64)
    for name, method, step in (
        ("spec_pcgmix_square", "durratiomixup", 5),
        ("spec_timemask_square", "durmixtimemask(0.4)", 6),
        ("spec_timemask_default", "durmixtimemask", 12),
        ("spec_freqmask_square", "durmixfreqmask(0.5)", 7),
        ("spec_cutout_square", "durmixcutout(0.5,0.6)", 8),
    ):
        out, mix = run_2d(method, step, x, lab, fr)
        assert len(mix) == b
        save(name, entry=np.array("augmentations2d.augment"), method=np.array(method), step=np.int64(step),
             data=x, labels=lab, frames=fr, out=out, mix=mix)

    # non-square: the reference dispatcher cannot run (augmentations2d.py:409), so pin the
    # per-pair function in the dispatcher's own loop with the dispatcher's own draws.
    b, f_bins, t_cols = 5, 16, 250
    fr = cycle_frames(rng, b, limit=2200)
    fr = np.round(fr * (t_cols / 2200)).astype(np.int64)
    x = signals(rng, fr, (1, f_bins), t_cols)
    lab = rng.integers(0, 2, b).astype(np.int64)
    step = 14
    mix = ref2.get_same_label_mix_indices(torch.nn.functional.one_hot(torch.from_numpy(lab), 2), step)
    lam32 = np.array(np.ones(b) * ref2.get_lambda(1, step)).astype("float32")
    lam_t = torch.from_numpy(lam32)[:, None, None, None][0]
    xt = torch.from_numpy(x.copy())
    out = np.zeros_like(x)
    for i in range(b):
        out[i] = ref2.mixup_keepdur_multidim_tensors(xt[i], xt[mix][i], fr[i], fr[mix][i], lam_t,
                                                     "durratiomixup", step).numpy()
    save("spec_pcgmix_nonsquare", entry=np.array("augmentations2d.mixup_keepdur_multidim_tensors in the :419-426 loop"),
         method=np.array("durratiomixup"), step=np.int64(step), data=x, labels=lab, frames=fr, out=out,
         mix=np.asarray(mix, np.int64))

    # ---- '(rand)' displacement variant (added after the fixtures above; kept last so that the
    # seeded stream that produced them is unchanged) ---------------------------------------------
    for name, method, step, b, c, length in (
        ("pcgmix_rand", "(rand)durratiomixup", 15, 9, 2, 1600),
        ("pcgmixplus_rand", "(rand)durmixmagwarp(0.2,4)", 16, 6, 3, 2500),
    ):
        fr = cycle_frames(rng, b, limit=length)
        x = signals(rng, fr, (c,), length)
        lab = rng.integers(0, 2, b).astype(np.int64)
        out, tgt, mix, same = run_1d(ref1, method, step, x, lab, fr)
        assert not same
        save(name, entry=np.array("augmentations.augment"), method=np.array(method), step=np.int64(step),
             data=x, labels=lab, frames=fr, out=out, target=tgt, mix=mix)


if __name__ == "__main__":
    main()
