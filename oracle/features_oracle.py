"""CPU oracle for the per-cycle classical features (amplitude and Hilbert-envelope blocks).

TEST INFRASTRUCTURE — NOT PRODUCT CODE (only ``tests/`` and ``bench.py``'s CPU legs may import it).

Restates, in this repo's own words, ``classical.feature_vector_seg`` of the reference from the slices at
classical.py:248-253 through the amplitude block (:284-303) and the envelope block (:305-360), for one
float32 row and its five offsets.  PARITY PIN: ``tests/golden/cycle_features.npz`` holds the outputs of the
reference's own statements executed verbatim (``tests/golden/make_golden_features.py``);
``tests/test_features.py`` holds this file to them bit-for-bit (it makes the same NumPy / SciPy calls on
the same float32 data, so nothing is left to tolerance here).
"""
from __future__ import annotations

import numpy as np
from scipy import signal
from scipy.signal import hilbert

N_FEATURES = 36
N_PSD_FEATURES = 80
PSD_BANDS = ((25, 40), (40, 60), (60, 80), (80, 100), (100, 120), (120, 140), (140, 160), (160, 180), (180, 200), (200, 250),
             (250, 300), (300, 400))


def _round4(v):
    return round(v, 4)                 # np.float32.__round__: rint(v * 1e4) / 1e4 in float32


def _trapz(y, dx):
    return (dx * (y[1:] + y[:-1]) / 2.0).sum()      # what np.trapz(y, dx=dx) evaluates (classical.py:307 ...)


def cycle_features(row: np.ndarray, frames) -> np.ndarray:
    """36 float32 features of one cycle (layout: ``pcgmix_cycle_features`` in include/pcgmix_b200.h)."""
    row = np.asarray(row, dtype=np.float32)
    f = [int(v) for v in frames[:5]]
    seg = [row[:f[1]], row[f[1]:f[2]], row[f[2]:f[3]], row[f[3]:f[4]], row[:f[4]]]    # S1, systole, S2, diastole, RR
    out = np.zeros(N_FEATURES, np.float32)
    mx = [np.max(s) for s in seg[:4]]
    out[0:4] = mx
    out[4] = _round4(mx[0] / mx[2])
    out[5] = _round4(mx[1] / mx[3])
    out[6] = _round4(mx[1] / mx[0])
    out[7] = _round4(mx[1] / mx[2])
    out[8] = _round4(mx[3] / mx[0])
    out[9] = _round4(mx[3] / mx[2])
    env = [np.abs(hilbert(s)) for s in seg]
    integ = [_trapz(e, 5) for e in env]
    out[10:15] = integ
    out[15] = _round4(integ[0] / integ[2])
    out[16] = _round4(integ[1] / integ[3])
    out[17] = _round4(integ[0] / integ[4])
    out[18] = _round4(integ[1] / integ[4])
    out[19] = _round4(integ[2] / integ[4])
    out[20] = _round4(integ[3] / integ[4])
    out[21] = _round4(integ[1] / integ[0])
    out[22] = _round4(integ[3] / integ[2])
    mean = [np.mean(e) for e in env]
    out[23:28] = mean
    out[28] = mean[0] / mean[4]
    out[29] = mean[1] / mean[4]
    out[30] = mean[2] / mean[4]
    out[31] = mean[3] / mean[4]
    out[32] = mean[1] / mean[3]
    out[33] = mean[1] / mean[0]
    out[34] = mean[3] / mean[2]
    out[35] = mean[0] / mean[2]
    return out


def cycle_psd_features(row: np.ndarray, frames, fs: int = 1000) -> np.ndarray:
    """The power-spectral-density block of ``feature_vector_seg`` (classical.py:358-643) for one cycle: for the whole
    beat, the systole and the diastole — Welch PSD (SciPy defaults: Hann window of min(256, n) samples, half overlap,
    mean removed per window, density scaling, one-sided, float32 arithmetic for float32 rows), the trapezoid integral
    (dx = 5) of the Hilbert envelope of the PSD, then the mean of the PSD and of PSD / integral over all bins and over
    the bins inside each of twelve closed frequency bands (an empty band gives NaN) — and two rounded ratios.
    80 values, layout: ``pcgmix_cycle_psd_features`` in include/pcgmix_b200.h.  PARITY PIN:
    tests/golden/cycle_psd_features.npz (the reference's statements executed verbatim), held bit-for-bit."""
    row = np.asarray(row, dtype=np.float32)
    f = [int(v) for v in frames[:5]]
    out = np.zeros(N_PSD_FEATURES, np.float64)
    norm_mean = []
    for k, seg in enumerate((row[:f[4]], row[f[1]:f[2]], row[f[3]:f[4]])):            # RR, systole, diastole
        freqs, psd = signal.welch(seg, fs)
        integral = _np_trapz(np.abs(hilbert(psd)), 5)
        normalized = psd / integral
        base = 26 * k
        out[base] = np.mean(psd)
        out[base + 1] = np.mean(normalized)
        norm_mean.append(np.mean(normalized))
        for j, (lo, hi) in enumerate(PSD_BANDS):
            inside = (lo <= freqs) & (freqs <= hi)
            out[base + 2 + 2 * j] = np.mean(psd[inside])
            out[base + 3 + 2 * j] = np.mean(normalized[inside])
    out[78] = round(norm_mean[1] / norm_mean[0], 4)
    out[79] = round(norm_mean[2] / norm_mean[0], 4)
    return out


def cycle_moment_features(row: np.ndarray, frames) -> np.ndarray:
    """The skewness / kurtosis block of ``feature_vector_seg`` (classical.py:893-905): ``scipy.stats.skew`` and
    ``scipy.stats.kurtosis`` (biased, Fisher) of RR, S1, systole, S2, diastole — 10 values, float32 arithmetic for float32
    rows.  PARITY PIN: tests/golden/cycle_moment_features.npz (the reference's statements executed verbatim), bit-for-bit."""
    from scipy import stats
    row = np.asarray(row, dtype=np.float32)
    f = [int(v) for v in frames[:5]]
    seg = [row[:f[4]], row[:f[1]], row[f[1]:f[2]], row[f[2]:f[3]], row[f[3]:f[4]]]
    return np.array([stats.skew(s) for s in seg] + [stats.kurtosis(s) for s in seg], dtype=np.float64)


def batch_moment_features(data: np.ndarray, frames: np.ndarray, channel: int) -> np.ndarray:
    return np.stack([cycle_moment_features(data[i, channel], frames[i]) for i in range(data.shape[0])])


def _np_trapz(y, dx):
    return (np.trapz if hasattr(np, "trapz") else np.trapezoid)(y, dx=dx)


def batch_psd_features(data: np.ndarray, frames: np.ndarray, channel: int, fs: int = 1000) -> np.ndarray:
    return np.stack([cycle_psd_features(data[i, channel], frames[i], fs) for i in range(data.shape[0])])


def batch_features(data: np.ndarray, frames: np.ndarray, channel: int) -> np.ndarray:
    """The loop of train_model.py:519-532: one call per cycle on ``data[i, channel]``."""
    return np.stack([cycle_features(data[i, channel], frames[i]) for i in range(data.shape[0])])
