"""Synthetic PhysioNet-shaped inputs: cardiac cycles with physiological
state durations, zero-padded signals, dense Springer state vectors.  Used by the benchmark and
by tests; there is no dataset in the reference repository and no network here."""
from __future__ import annotations

import numpy as np

# state durations in milliseconds: S1, systole, S2, diastole
DUR_LO_MS = np.array([90, 150, 70, 300])
DUR_HI_MS = np.array([160, 400, 130, 900])
BENCH_SEED = 20241018


def cycle_frames(rng: np.random.Generator, n: int, fs: int = 1000, limit: int | None = None) -> np.ndarray:
    """(n, 5) int64 cumulative offsets; durations uniform in the physiological ranges, in samples
    at ``fs``.  With ``limit`` the (rare) cycles longer than ``limit`` are shrunk to fit."""
    ms = rng.integers(DUR_LO_MS, DUR_HI_MS + 1, size=(n, 4))
    dur = (ms * fs // 1000).astype(np.int64)
    fr = np.concatenate([np.zeros((n, 1), np.int64), np.cumsum(dur, axis=1)], axis=1)
    if limit is not None:
        scale = np.minimum(1.0, limit / np.maximum(fr[:, 4:5], 1))
        fr = np.floor(fr * scale).astype(np.int64)
    return fr


def spectrogram_frames(rng: np.random.Generator, n: int, n_cols: int, seconds: float = 2.5) -> np.ndarray:
    """Cycle offsets in spectrogram-column units: durations at 1 kHz scaled to ``n_cols`` columns
    per ``seconds`` and rounded (half-even), as the reference's frame mapping does."""
    fr = cycle_frames(rng, n, 1000)
    cols = np.rint(fr * (n_cols / (seconds * 1000.0))).astype(np.int64)
    return np.minimum(cols, n_cols)


def cycle_signals(rng: np.random.Generator, frames: np.ndarray, row_shape: tuple, length: int) -> np.ndarray:
    """fp32 standard-normal samples inside each cycle, exact zeros after ``frames[:, 4]``;
    shape (n, *row_shape, length)."""
    n = frames.shape[0]
    x = rng.standard_normal((n,) + tuple(row_shape) + (length,), dtype=np.float32)
    t = np.arange(length)[None, :]
    keep = (t < frames[:, 4:5]).reshape((n,) + (1,) * len(row_shape) + (length,))
    x *= keep
    return x


def dense_states(rng: np.random.Generator, n_rec: int, n_samples: int, fs: int = 2000) -> np.ndarray:
    """(n_rec, n_samples) int8 Springer states in {1,2,3,4}; every recording starts at a random
    phase inside a random state."""
    out = np.empty((n_rec, n_samples), np.int8)
    for r in range(n_rec):
        state = int(rng.integers(0, 4))
        pos = 0
        first = True
        while pos < n_samples:
            ms = int(rng.integers(DUR_LO_MS[state], DUR_HI_MS[state] + 1))
            ln = max(1, ms * fs // 1000)
            if first:
                ln = max(1, int(rng.integers(1, ln + 1)))
                first = False
            out[r, pos:pos + ln] = state + 1
            pos += ln
            state = (state + 1) % 4
    return out


def mixed_samples(frames: np.ndarray, mix: np.ndarray) -> int:
    """Sum over cycles of M = sum_s min(len1_s, len2_s): the partner samples one pass must read."""
    d1 = np.diff(frames, axis=1)
    d2 = np.diff(frames[mix], axis=1)
    return int(np.minimum(d1, d2).sum())
