"""CPU oracle for the segmentation-index side of the hot path (annotations -> per-cycle frames,
cut + zero-pad, duration ratios).

TEST INFRASTRUCTURE — NOT PRODUCT CODE (same rule as ``pcgmix_oracle.py``).

PARITY UNPINNED by executable reference code: the cycle builder lives in notebook cells
(``databuilder.ipynb``) that read audio/annotation files absent from the reference repo, and
``classical.py`` cannot be imported here (it needs xgboost/lightgbm/pywt/antropy).  The
functions below restate those cells line by line instead; every rule cites the raw-JSON line
of ``/root/reference/databuilder.ipynb`` (what ``grep -n`` shows) or ``classical.py``:

  * dense per-sample state vector -> transitions ........ databuilder.ipynb:593-597 (cell 14)
  * (frame, state-name) table, ``//2`` downsample ........ databuilder.ipynb:928 (cell 25)
  * complete-cycle rule, noise skip, pattern check ....... databuilder.ipynb:599-606, 932-948
  * spectrogram frame mapping with ``round`` ............. databuilder.ipynb:370, 399 (cell 6)
  * cut + zero-pad (``ndarray.resize`` / ``np.pad``) ..... databuilder.ipynb:627-632, 973-978, 403-411
  * duration features ................................... classical.py:245-283
"""
from __future__ import annotations

import numpy as np

S1, SYSTOLE, S2, DIASTOLE, NOISE = 1, 2, 3, 4, 5
_NAME_TO_CODE = {"S1": S1, "systole": SYSTOLE, "S2": S2, "diastole": DIASTOLE}


class SegmentPatternError(Exception):
    """The reference raises ``Exception('Segment states are not correct!')``."""


def state_code(name) -> int:
    """Map a PhysioNet state label to the integer code the device kernels use.  Anything with an
    ``N`` in it is a noise marker (the reference tests ``'N' in ''.join(seg_states)``)."""
    if name in _NAME_TO_CODE:
        return _NAME_TO_CODE[name]
    if "N" in str(name):
        return NOISE
    return 0


def transitions_from_dense(states: np.ndarray):
    """Cell 14: indices where the per-sample state changes, and the state entered there."""
    states = np.asarray(states)
    pos = np.where(states[:-1] != states[1:])[0]
    pos = pos + 1
    entered = [int(states[p]) for p in pos]
    return [int(p) for p in pos], entered


def cycles_from_transitions(positions, codes, downsample: int = 1, noise_skip: bool = True):
    """Complete-cycle rule shared by cells 14 and 25.

    ``positions`` are absolute sample indices of the state changes, ``codes`` the state entered.
    Positions are floor-divided by ``downsample`` FIRST (on the absolute index), then a cycle
    starts at every transition into S1 that has another S1 somewhere later; its four states must
    read S1, systole, S2, diastole (else the reference raises), unless one of them is a noise
    marker, in which case the cycle is skipped (cell 25 only: ``noise_skip``).  Returns ``(rel_frames (n,5) int64,
    abs_start (n,) int64, abs_stop (n,) int64)`` in downsampled units.
    """
    pos = [int(p) // downsample for p in positions]
    codes = [int(c) for c in codes]
    rel, starts, stops = [], [], []
    for i, code in enumerate(codes):
        if code == S1 and S1 in codes[i + 1:]:
            four = codes[i:i + 4]
            if noise_skip and NOISE in four:
                continue
            if four != [S1, SYSTOLE, S2, DIASTOLE]:
                raise SegmentPatternError("Segment states are not correct!")
            window = np.asarray(pos[i:i + 5], dtype=np.int64)
            rel.append(window - window[0])
            starts.append(window[0])
            stops.append(window[4])
    if not rel:
        return (np.zeros((0, 5), np.int64), np.zeros((0,), np.int64), np.zeros((0,), np.int64))
    return np.stack(rel), np.asarray(starts, np.int64), np.asarray(stops, np.int64)


def cycles_from_dense(states: np.ndarray, downsample: int = 1):
    """Cell 14 end to end for one recording."""
    pos, entered = transitions_from_dense(states)
    return cycles_from_transitions(pos, entered, downsample, noise_skip=False)


def spectrogram_positions(positions, n_spec_cols: int, n_samples: int):
    """Cell 6 (:370): ``round(f * T_spec / len(y))`` — Python ``round`` is half-to-even."""
    return [round(int(f) * n_spec_cols / n_samples) for f in positions]


def cycles_from_transitions_spec(positions, codes, n_spec_cols: int, n_samples: int):
    """Cell 6: cycles are found on the raw positions (no downsample), the five offsets are then
    taken from the rounded spectrogram positions (:399)."""
    codes = [int(c) for c in codes]
    spec_pos = spectrogram_positions(positions, n_spec_cols, n_samples)
    rel, starts, stops = [], [], []
    for i, code in enumerate(codes):
        if code == S1 and S1 in codes[i + 1:]:
            four = codes[i:i + 4]
            if NOISE in four:
                continue
            if four != [S1, SYSTOLE, S2, DIASTOLE]:
                raise SegmentPatternError("Segment states are not correct!")
            window = np.asarray(spec_pos[i:i + 5], dtype=np.int64)
            rel.append(window - window[0])
            starts.append(window[0])
            stops.append(window[4])
    if not rel:
        return (np.zeros((0, 5), np.int64), np.zeros((0,), np.int64), np.zeros((0,), np.int64))
    return np.stack(rel), np.asarray(starts, np.int64), np.asarray(stops, np.int64)


def cut_and_pad(signal: np.ndarray, start: int, stop: int, length: int) -> np.ndarray:
    """``seg = y[start:stop]; seg.resize(length)`` (:627-632, :973-978): slice with Python
    clipping, then truncate or zero-fill to ``length``."""
    seg = np.array(signal[start:stop], copy=True)
    out = np.zeros(length, dtype=signal.dtype)
    n = min(length, seg.shape[0])
    out[:n] = seg[:n]
    return out


def cut_and_pad_spec(spec: np.ndarray, start: int, stop: int, n_cols: int) -> np.ndarray:
    """``np.pad(spec[:, start:stop], ((0,0),(0,max(0,n_cols-w))))`` (:403-411).  The reference
    does not truncate a cycle wider than ``n_cols``; neither does this."""
    seg = spec[:, start:stop]
    pad = max(0, n_cols - seg.shape[1])
    return np.pad(seg, ((0, 0), (0, pad)), mode="constant")


DURATION_FEATURE_NAMES = (
    "duration_RR", "BPM", "duration_S1", "duration_systole", "duration_S2", "duration_diastole",
    "ratio_S1_S2", "ratio_systole_diastole", "ratio_S1_RR", "ratio_systole_RR", "ratio_S2_RR",
    "ratio_diastole_RR", "ratio_systole_S1", "ratio_diastole_S2",
)


def duration_features(frames, fs: int = 1000) -> np.ndarray:
    """classical.py:245-283: state durations in ms (``int(len*1000/Fs)``), BPM and the eight
    duration ratios, each ``round(.., 4)``.  Returns 14 float64 values in the order of
    ``DURATION_FEATURE_NAMES``.  A zero denominator raises ZeroDivisionError, as in the
    reference."""
    f = [int(v) for v in frames]
    rr = int((f[4]) * 1000 / fs)            # RR = data[:frames[-1]]
    s1 = int((f[1]) * 1000 / fs)            # S1 = data[:frames[1]]
    sy = int((f[2] - f[1]) * 1000 / fs)
    s2 = int((f[3] - f[2]) * 1000 / fs)
    di = int((f[4] - f[3]) * 1000 / fs)
    return np.array([
        rr, round(60000 / rr, 4), s1, sy, s2, di,
        round(s1 / s2, 4), round(sy / di, 4), round(s1 / rr, 4), round(sy / rr, 4),
        round(s2 / rr, 4), round(di / rr, 4), round(sy / s1, 4), round(di / s2, 4),
    ], dtype=np.float64)
