"""Golden vectors for the power-spectral-density block of ``classical.feature_vector_seg`` (classical.py:358-643:
Welch PSD of the whole beat, the systole and the diastole, its Hilbert-envelope integral, mean PSD and mean
normalised PSD overall and in twelve frequency bands, two ratios), produced by EXECUTING the reference's own
statements verbatim — the segment slices at the top of the function body, then everything from the first
``signal.welch`` comment through ``mean_psd_ratio_diastole_RR`` — per cycle on float32 rows, as
train_model.py:519-532 feeds them.  (``classical.py`` as a whole cannot be imported here: librosa, pywt,
antropy ... are absent; SciPy, which is all this block needs, is present.)

Run in the build container:  python tests/golden/make_golden_psd_features.py
"""
from __future__ import annotations

import os
import textwrap
import warnings

import numpy as np
from scipy import signal
from scipy.signal import hilbert

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PCGMIX_REFERENCE_ROOT", "/root/reference")

BANDS = ((25, 40), (40, 60), (60, 80), (80, 100), (100, 120), (120, 140), (140, 160), (160, 180), (180, 200), (200, 250),
         (250, 300), (300, 400))
SEGMENTS = ("RR", "systole", "diastole")


def names():
    out = []
    for s in SEGMENTS:
        out += [f"mean_psd_{s}", f"mean_psd_{s}_normalized"]
        for lo, hi in BANDS:
            out += [f"mean_psd_{s}_{lo}_{hi}_hz", f"mean_psd_{s}_normalized_{lo}_{hi}_hz"]
    return out + ["mean_psd_ratio_systole_RR", "mean_psd_ratio_diastole_RR"]


NAMES = names()


def reference_statements():
    src = open(os.path.join(REF, "classical.py")).read().split("\n")
    i0 = next(i for i, l in enumerate(src) if l.startswith("def feature_vector_seg("))
    i_seg = next(i for i in range(i0, len(src)) if src[i].strip().startswith("diastole = data[frames[3]:frames[4]]"))
    j0 = next(i for i in range(i_seg, len(src)) if "freqs, psd_RR = signal.welch(RR, Fs)" in src[i])
    j1 = next(i for i in range(j0, len(src)) if src[i].strip().startswith("mean_psd_ratio_diastole_RR ="))
    body = src[i0 + 1:i_seg + 1] + src[j0:j1 + 1]
    return compile(textwrap.dedent("\n".join(body)), "<classical.py feature_vector_seg, PSD block>", "exec")


def cycles(rng, n, length):
    lo = np.array([90, 150, 70, 300])
    hi = np.array([160, 400, 130, 900])
    dur = rng.integers(lo, hi + 1, size=(n, 4))
    frames = np.concatenate([np.zeros((n, 1), np.int64), np.cumsum(dur, axis=1)], axis=1)
    frames[0] = [0, 100, 356, 456, 1224]             # systole of exactly 256 samples, diastole of 768 (5 windows)
    frames[1] = [0, 100, 355, 456, 968]              # 255 (one short window, odd) and 512
    frames[2] = [0, 90, 130, 200, 240]               # 40 samples each: 21 bins at multiples of 25 Hz, most bands empty
    frames[3] = [0, 5, 8, 10, 12]                    # three and two samples
    frames[4] = [0, 120, 400, 520, 2700]             # runs past the row: slices clamp
    frames[5] = [3, 100, 301, 398, 1001]             # odd lengths
    frames[6] = [0, 110, 366, 470, 2500]             # ends exactly at the row end
    t = np.arange(length)
    data = np.zeros((n, length), np.float32)
    for i in range(n):
        f = np.minimum(frames[i], length)
        # heart-sound-like: coloured noise (steep spectrum), bursts on S1 / S2, a murmur in some systoles
        x = np.cumsum(0.02 * rng.standard_normal(length)) * 0.1 + 0.02 * rng.standard_normal(length)
        for a, b in ((0, f[1]), (f[2], f[3])):
            m = max(b - a, 1)
            x[a:b] += np.hanning(m)[: b - a] * np.sin(2 * np.pi * rng.uniform(0.02, 0.08) * t[: b - a] + rng.uniform(0, 6)) * rng.uniform(0.5, 2.0)
        if i % 3 == 0:
            a, b = f[1], f[2]
            x[a:b] += 0.2 * rng.standard_normal(b - a) * np.hanning(max(b - a, 1))[: b - a]
        x[f[4]:] = 0.0
        data[i] = x.astype(np.float32)
    data[7] *= np.float32(1e-3)                      # small amplitudes
    data[8] = rng.standard_normal(length).astype(np.float32)     # white noise: flat spectrum
    data[8, min(frames[8, 4], length):] = 0.0
    return data, frames


def main():
    code = reference_statements()
    rng = np.random.default_rng(20261020)
    n, length = 64, 2500
    data, frames = cycles(rng, n, length)
    feats = np.zeros((n, len(NAMES)), np.float64)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i in range(n):
            ns = dict(np=np, hilbert=hilbert, signal=signal, data=data[i], frames=frames[i])
            exec(code, ns)
            for k, name in enumerate(NAMES):
                feats[i, k] = ns[name]
            assert ns["psd_RR"].dtype == np.float32
    np.savez_compressed(os.path.join(HERE, "cycle_psd_features.npz"),
                        entry=np.array("classical.py feature_vector_seg, PSD block (Welch + envelope integral + band means), executed verbatim"),
                        data=data, frames=frames, features=feats, names=np.array(NAMES))
    print("psd features:", n, "cycles x", len(NAMES), "features;", os.path.getsize(os.path.join(HERE, "cycle_psd_features.npz")) // 1024, "KiB;",
          "nan:", int(np.isnan(feats).sum()), "inf:", int(np.isinf(feats).sum()))


if __name__ == "__main__":
    main()
