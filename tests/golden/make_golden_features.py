"""Golden vectors for the per-cycle classical features (amplitude and Hilbert-envelope blocks of
``classical.feature_vector_seg``, classical.py:284-360), produced by EXECUTING the reference's own
statements verbatim (``classical.py`` as a whole cannot be imported here: xgboost, lightgbm, pywt, antropy,
librosa ... are absent).  The statements from the top of the function body through the last mean-envelope
ratio are located by their text, compiled and run per cycle on float32 rows, exactly as
train_model.py:519-532 feeds them (one channel of an augmented batch, ``frames`` of the cycle).

Run in the build container:  python tests/golden/make_golden_features.py
"""
from __future__ import annotations

import os
import textwrap
import warnings

import numpy as np
from scipy.signal import hilbert

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PCGMIX_REFERENCE_ROOT", "/root/reference")

NAMES = (["max_amplitude_" + s for s in ("S1", "systole", "S2", "diastole")] +
         ["max_amplitude_ratio_" + s for s in ("S1_S2", "systole_diastole", "systole_S1", "systole_S2", "diastole_S1", "diastole_S2")] +
         ["envelope_integral_" + s for s in ("S1", "systole", "S2", "diastole", "RR")] +
         ["envelope_integral_ratio_" + s for s in ("S1_S2", "systole_diastole", "S1_RR", "systole_RR", "S2_RR", "diastole_RR",
                                                   "systole_S1", "diastole_S2")] +
         ["mean_envelope_" + s for s in ("S1", "systole", "S2", "diastole", "RR")] +
         ["mean_envelope_ratio_" + s for s in ("S1_RR", "systole_RR", "S2_RR", "diastole_RR", "systole_diastole", "systole_S1",
                                               "diastole_S2", "S1_S2")])


def reference_statements():
    src = open(os.path.join(REF, "classical.py")).read().split("\n")
    i0 = next(i for i, l in enumerate(src) if l.startswith("def feature_vector_seg("))
    i1 = next(i for i in range(i0, len(src)) if "mean_envelope_ratio_S1_S2 = mean_envelope_S1/mean_envelope_S2" in src[i])
    return compile(textwrap.dedent("\n".join(src[i0 + 1:i1 + 1])), "<classical.py feature_vector_seg>", "exec")


def main():
    code = reference_statements()
    rng = np.random.default_rng(20261019)
    n, length = 96, 2500
    lo = np.array([90, 150, 70, 300])
    hi = np.array([160, 400, 130, 900])
    dur = rng.integers(lo, hi + 1, size=(n, 4))
    frames = np.concatenate([np.zeros((n, 1), np.int64), np.cumsum(dur, axis=1)], axis=1)
    frames[0] = [0, 1, 2, 3, 4]                      # one-sample states
    frames[1] = [0, 2, 5, 7, 12]
    frames[2] = [3, 100, 301, 398, 1001]             # f0 != 0 (S1 still starts at column 0), odd lengths
    frames[3] = [0, 128, 384, 512, 1024]             # powers of two
    frames[4] = [0, 120, 400, 520, 2700]             # runs past the row: slices clamp
    frames[5] = [0, 97, 331, 433, 2500]              # ends exactly at the row end
    # heart-sound-like rows: band-limited noise bursts on S1 / S2, low-level noise elsewhere, zero padding
    t = np.arange(length)
    data = np.zeros((n, length), np.float32)
    for i in range(n):
        f = np.minimum(frames[i], length)
        x = 0.05 * rng.standard_normal(length)
        for a, b in ((0, f[1]), (f[2], f[3])):
            m = max(b - a, 1)
            x[a:b] += np.hanning(m)[: b - a] * np.sin(2 * np.pi * rng.uniform(0.02, 0.08) * t[: b - a] + rng.uniform(0, 6)) * rng.uniform(0.5, 2.0)
        x[f[4]:] = 0.0
        data[i] = x.astype(np.float32)
    data[7] *= np.float32(1e-3)                      # small amplitudes
    data[8] = -np.abs(data[8]) - np.float32(0.01)   # all-negative row: negative maxima and ratios
    data[8, min(frames[8, 4], length):] = 0.0
    feats = np.zeros((n, len(NAMES)), np.float32)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i in range(n):
            ns = dict(np=np, hilbert=hilbert, data=data[i], frames=frames[i])
            exec(code, ns)
            for k, name in enumerate(NAMES):
                assert isinstance(ns[name], np.float32), (name, type(ns[name]))
                feats[i, k] = ns[name]
    np.savez_compressed(os.path.join(HERE, "cycle_features.npz"),
                        entry=np.array("classical.py feature_vector_seg, amplitude + envelope blocks, executed verbatim"),
                        data=data, frames=frames, features=feats, names=np.array(NAMES))
    print("cycle features:", n, "cycles x", len(NAMES), "features;", os.path.getsize(os.path.join(HERE, "cycle_features.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
