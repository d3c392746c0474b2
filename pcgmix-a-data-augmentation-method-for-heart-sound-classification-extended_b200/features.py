"""Classical per-cycle features of augmented cycles, computed on the GPU where the batch already is.

With ``args.classical_space`` the reference takes every augmented batch back to the host and runs
``classical.feature_vector_seg(d[4], t, f, w, sq, 5, 'train')`` per cycle in a Python loop, concatenating
pandas columns (train_model.py:519-532).  This module provides the first three blocks of that function as
device kernels over the whole batch (one launch each, no D2H of the samples):

  duration block   classical.py:248-283  ->  ``segmentation.duration_features`` (14 float64 values)
  amplitude block  classical.py:284-303  ->  :func:`cycle_features`, columns 0..9   (exact float32)
  envelope block   classical.py:305-360  ->  :func:`cycle_features`, columns 10..35 (float32 rounding)

  PSD block        classical.py:358-643  ->  :func:`cycle_psd_features`, 80 values  (float32 rounding)
  moments block    classical.py:893-905  ->  :func:`cycle_moment_features`, 10 values (float32 rounding)

The blocks between and behind those in the reference (zero crossings, chroma, RMS, spectral statistics, MFCC, wavelets,
entropies: librosa, PyWavelets and antropy calls) are not provided.  Column names are the reference's variable names (:data:`FEATURE_NAMES`,
:data:`PSD_FEATURE_NAMES`), so a caller can build the same table.
"""
from __future__ import annotations

import numpy as np
import torch

from . import native, staging

__all__ = ["FEATURE_NAMES", "PSD_FEATURE_NAMES", "MOMENT_FEATURE_NAMES", "DURATION_NAMES", "cycle_features",
           "cycle_psd_features", "cycle_moment_features", "classical_space_features"]

_STATES = ("S1", "systole", "S2", "diastole")
FEATURE_NAMES = tuple(
    ["max_amplitude_" + s for s in _STATES] +
    ["max_amplitude_ratio_" + s for s in ("S1_S2", "systole_diastole", "systole_S1", "systole_S2", "diastole_S1", "diastole_S2")] +
    ["envelope_integral_" + s for s in _STATES + ("RR",)] +
    ["envelope_integral_ratio_" + s for s in ("S1_S2", "systole_diastole", "S1_RR", "systole_RR", "S2_RR", "diastole_RR",
                                              "systole_S1", "diastole_S2")] +
    ["mean_envelope_" + s for s in _STATES + ("RR",)] +
    ["mean_envelope_ratio_" + s for s in ("S1_RR", "systole_RR", "S2_RR", "diastole_RR", "systole_diastole", "systole_S1",
                                          "diastole_S2", "S1_S2")])
DURATION_NAMES = ("duration_RR", "BPM", "duration_S1", "duration_systole", "duration_S2", "duration_diastole",
                  "duration_ratio_S1_S2", "duration_ratio_systole_diastole", "duration_ratio_S1_RR",
                  "duration_ratio_systole_RR", "duration_ratio_S2_RR", "duration_ratio_diastole_RR",
                  "duration_ratio_systole_S1", "duration_ratio_diastole_S2")
assert len(FEATURE_NAMES) == native.CYCLE_FEATURES
PSD_BANDS = ((25, 40), (40, 60), (60, 80), (80, 100), (100, 120), (120, 140), (140, 160), (160, 180), (180, 200), (200, 250),
             (250, 300), (300, 400))
PSD_FEATURE_NAMES = tuple(
    [name for s in ("RR", "systole", "diastole")
     for name in [f"mean_psd_{s}", f"mean_psd_{s}_normalized"] +
     [n for lo, hi in PSD_BANDS for n in (f"mean_psd_{s}_{lo}_{hi}_hz", f"mean_psd_{s}_normalized_{lo}_{hi}_hz")]] +
    ["mean_psd_ratio_systole_RR", "mean_psd_ratio_diastole_RR"])
assert len(PSD_FEATURE_NAMES) == native.CYCLE_PSD_FEATURES
MOMENT_FEATURE_NAMES = tuple(["skew_" + s for s in ("RR",) + _STATES] + ["kurtosis_" + s for s in ("RR",) + _STATES])
assert len(MOMENT_FEATURE_NAMES) == native.CYCLE_MOMENT_FEATURES


def _frames_on_device(frames, batch: int, device) -> torch.Tensor:
    if isinstance(frames, torch.Tensor) and frames.is_cuda:
        if frames.dtype != torch.int32:
            raise TypeError("device frames must be int32")
        return frames
    f = frames.detach().cpu().numpy() if isinstance(frames, torch.Tensor) else np.asarray(frames)
    if not np.issubdtype(f.dtype, np.integer) or f.ndim != 2 or f.shape[0] != batch or f.shape[1] < 5:
        raise ValueError(f"frames must be an integer ({batch}, 5) array")
    if (f[:, :5] < 0).any() or (f[:, :5] > 2 ** 31 - 1).any():
        raise ValueError("frames must be non-negative int32 offsets")
    return staging.upload([np.ascontiguousarray(f[:, :5].astype(np.int32))], device)[0]


def cycle_features(data: torch.Tensor, frames, channel: int = 4, amplitude: bool = True, envelope: bool = True,
                   out: torch.Tensor = None, err_flag: torch.Tensor = None) -> torch.Tensor:
    """``(B, 36)`` float32 features of ``data[:, channel]`` (layout: :data:`FEATURE_NAMES`).

    ``data`` (B, C, L) float32 on a CUDA device (e.g. what ``augment`` returned), ``frames`` the cycles' five
    offsets (CPU integer tensor as the loader yields them, or int32 on the device).  Blocks that are switched
    off leave their columns untouched (NaN in a fresh result).  There is no CPU path."""
    if not isinstance(data, torch.Tensor) or not data.is_cuda:
        raise RuntimeError("cycle_features: data must be a CUDA tensor (there is no CPU fallback)")
    if data.dim() != 3 or data.dtype != torch.float32:
        raise ValueError("cycle_features: data must be (B, C, L) float32")
    if not 0 <= channel < data.shape[1]:
        raise ValueError(f"channel {channel} outside the {data.shape[1]} channels of the batch")
    what = (1 if amplitude else 0) | (2 if envelope else 0)
    if what == 0:
        raise ValueError("nothing to compute")
    data = data if data.is_contiguous() else data.contiguous()
    if out is None:
        out = torch.full((data.shape[0], native.CYCLE_FEATURES), float("nan"), dtype=torch.float32, device=data.device)
    native.cycle_features(data, _frames_on_device(frames, data.shape[0], data.device), channel, what, out, err_flag)
    return out


def cycle_psd_features(data: torch.Tensor, frames, channel: int = 4, fs: int = 1000, out: torch.Tensor = None,
                       err_flag: torch.Tensor = None) -> torch.Tensor:
    """``(B, 80)`` float32 Welch-PSD features of ``data[:, channel]`` (layout: :data:`PSD_FEATURE_NAMES`): mean PSD and
    mean normalised PSD of the whole beat, the systole and the diastole, overall and in twelve frequency bands (NaN for
    a band that holds no bin of a short segment, like the reference), and the two ratios of classical.py:640-643.
    Arguments as for :func:`cycle_features`; the reference calls Welch with ``Fs = 1000``.  There is no CPU path."""
    if not isinstance(data, torch.Tensor) or not data.is_cuda:
        raise RuntimeError("cycle_psd_features: data must be a CUDA tensor (there is no CPU fallback)")
    if data.dim() != 3 or data.dtype != torch.float32:
        raise ValueError("cycle_psd_features: data must be (B, C, L) float32")
    if not 0 <= channel < data.shape[1]:
        raise ValueError(f"channel {channel} outside the {data.shape[1]} channels of the batch")
    if fs <= 0:
        raise ValueError("fs must be positive")
    data = data if data.is_contiguous() else data.contiguous()
    if out is None:
        out = torch.empty((data.shape[0], native.CYCLE_PSD_FEATURES), dtype=torch.float32, device=data.device)
    native.cycle_psd_features(data, _frames_on_device(frames, data.shape[0], data.device), channel, fs, out, err_flag)
    return out


def cycle_moment_features(data: torch.Tensor, frames, channel: int = 4, out: torch.Tensor = None,
                          err_flag: torch.Tensor = None) -> torch.Tensor:
    """``(B, 10)`` float32: ``scipy.stats.skew`` and ``scipy.stats.kurtosis`` of the whole beat and the four states of
    ``data[:, channel]`` (layout: :data:`MOMENT_FEATURE_NAMES`; classical.py:893-905).  There is no CPU path."""
    if not isinstance(data, torch.Tensor) or not data.is_cuda:
        raise RuntimeError("cycle_moment_features: data must be a CUDA tensor (there is no CPU fallback)")
    if data.dim() != 3 or data.dtype != torch.float32:
        raise ValueError("cycle_moment_features: data must be (B, C, L) float32")
    if not 0 <= channel < data.shape[1]:
        raise ValueError(f"channel {channel} outside the {data.shape[1]} channels of the batch")
    data = data if data.is_contiguous() else data.contiguous()
    if out is None:
        out = torch.empty((data.shape[0], native.CYCLE_MOMENT_FEATURES), dtype=torch.float32, device=data.device)
    native.cycle_moment_features(data, _frames_on_device(frames, data.shape[0], data.device), channel, out, err_flag)
    return out


def classical_space_features(data: torch.Tensor, frames, channel: int = 4, fs: int = 1000, psd: bool = True):
    """Duration, amplitude, envelope and (``psd``) PSD and moments blocks for a whole batch: ``(names, values)`` with
    ``values`` a (B, 14 + 36 [+ 80 + 10]) float64 device tensor in the reference's order of computation (durations first)."""
    from . import segmentation
    frames_dev = _frames_on_device(frames, data.shape[0], data.device)
    dur = segmentation.duration_features(frames_dev, fs)
    feats = cycle_features(data, frames_dev, channel)
    names, cols = DURATION_NAMES + FEATURE_NAMES, [dur, feats.to(torch.float64)]
    if psd:
        names = names + PSD_FEATURE_NAMES + MOMENT_FEATURE_NAMES
        cols = cols + [cycle_psd_features(data, frames_dev, channel, 1000).to(torch.float64),
                       cycle_moment_features(data, frames_dev, channel).to(torch.float64)]
    return names, torch.cat(cols, dim=1)
