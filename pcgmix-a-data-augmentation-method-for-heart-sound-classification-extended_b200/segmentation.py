"""Annotations -> cardiac cycles on the device (the offline stage that produces the ``frames``
the augmentation consumes).

Mirrors, as batched CUDA kernels, what the reference does per recording in Python inside
``databuilder.ipynb`` (line numbers of the raw notebook JSON):

  ``cycles_from_dense_states``   cell 14 (:593-606): dense per-sample Springer states ->
                                 transitions -> complete S1,systole,S2,diastole cycles
  ``cycles_from_state_table``    cell 25 (:928-948): (position, state) table with ``//ds``
                                 downsampling and noise-cycle skip; with ``spec_cols`` the
                                 cell-6 spectrogram mapping ``round(f*T_spec/len(y))`` (:370)
  ``cut_cycles``                 ``y[start:stop]`` + zero-pad to L (:627-632, :973-978, :403-411)
  ``duration_features``          ``classical.py:245-283`` (durations, BPM, duration ratios)

A cycle table is an int32 tensor (n, 8): recording, abs_start, abs_stop, f0..f4; its column
slice ``[:, 3:]`` is directly usable as ``frames`` by the mix kernels (no copy).
"""
from __future__ import annotations

import dataclasses

import torch

from . import native

STATE_CODES = {"S1": 1, "systole": 2, "S2": 3, "diastole": 4}
NOISE_CODE = 5

FEATURE_NAMES = (
    "duration_RR", "BPM", "duration_S1", "duration_systole", "duration_S2", "duration_diastole",
    "ratio_S1_S2", "ratio_systole_diastole", "ratio_S1_RR", "ratio_systole_RR", "ratio_S2_RR",
    "ratio_diastole_RR", "ratio_systole_S1", "ratio_diastole_S2",
)


class SegmentationError(Exception):
    """Raised where the reference raises ``Exception('Segment states are not correct!')`` and for
    capacity overflows."""


def state_code(name) -> int:
    """PhysioNet state label -> code; labels containing ``N`` are noise markers (the reference
    tests ``'N' in ''.join(seg_states)``)."""
    if name in STATE_CODES:
        return STATE_CODES[name]
    return NOISE_CODE if "N" in str(name) else 0


@dataclasses.dataclass
class CycleTable:
    cycles: torch.Tensor        # (capacity, 8) int32; rows [0, total) are valid
    row_ptr: torch.Tensor       # (R+1,) int32 row pointers per recording
    err_flag: torch.Tensor      # (1,) int32 device flag word

    def total(self) -> int:
        """Number of cycles (device -> host read; synchronises)."""
        return int(self.row_ptr[-1].item())

    def check(self):
        """Raise if the kernels flagged a pattern error or an overflow (synchronises)."""
        flag = int(self.err_flag.item())
        if flag & native.ERR_BAD_PATTERN:
            raise SegmentationError("Segment states are not correct!")
        if flag & native.ERR_OVERFLOW:
            raise SegmentationError("more cycles or transitions than the table can hold")
        return self

    @property
    def frames(self) -> torch.Tensor:
        """(capacity, 5) strided int32 view: the relative offsets f0..f4."""
        return self.cycles[:, 3:]


def _need_cuda(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the segmentation kernels have no CPU path")


def cycles_from_dense_states(states: torch.Tensor, downsample: int = 1, max_cycles: int | None = None) -> CycleTable:
    """``states`` (R, T) int8 in {1,2,3,4} on the device."""
    _need_cuda(states, "states")
    if states.dtype != torch.int8 or states.dim() != 2:
        raise TypeError("states must be a (R, T) int8 tensor")
    states = states.contiguous()
    R, T = states.shape
    if max_cycles is None:
        max_cycles = R * (T // 4 + 1)
        max_cycles = min(max_cycles, max(1024, R * 8192 // 4))
    dev = states.device
    cycles = torch.empty((max(max_cycles, 1), 8), dtype=torch.int32, device=dev)
    row_ptr = torch.empty(R + 1, dtype=torch.int32, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    native.segment_dense(states, downsample, cycles, row_ptr, err)
    return CycleTable(cycles, row_ptr, err)


def cycles_from_state_table(positions: torch.Tensor, codes: torch.Tensor, rec_offsets: torch.Tensor,
                            downsample: int = 1, spec_cols: int = 0, rec_len: torch.Tensor | None = None,
                            max_cycles: int | None = None) -> CycleTable:
    """Concatenated transition tables of R recordings (``rec_offsets`` (R+1,) delimits them)."""
    for t, n in ((positions, "positions"), (codes, "codes"), (rec_offsets, "rec_offsets")):
        _need_cuda(t, n)
    n_trans = positions.shape[0]
    R = rec_offsets.shape[0] - 1
    if max_cycles is None:
        max_cycles = n_trans // 4 + 1
    dev = positions.device
    cycles = torch.empty((max(max_cycles, 1), 8), dtype=torch.int32, device=dev)
    row_ptr = torch.empty(R + 1, dtype=torch.int32, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    native.segment_table(positions.contiguous(), codes.contiguous(), rec_offsets.contiguous(), downsample, spec_cols,
                         rec_len, cycles, row_ptr, err)
    return CycleTable(cycles, row_ptr, err)


def cut_cycles(signal: torch.Tensor, table: CycleTable, length: int, n_cycles: int | None = None) -> torch.Tensor:
    """``signal`` (R, C, T) fp32 -> (n, C, length): every cycle cut out of its recording and
    zero-padded (or truncated) to ``length``.  ``n_cycles=None`` reads the count from the device
    (one host sync); pass it to stay asynchronous."""
    _need_cuda(signal, "signal")
    if signal.dtype != torch.float32 or signal.dim() != 3:
        raise TypeError("signal must be (R, C, T) float32")
    signal = signal.contiguous()
    if n_cycles is None:
        n_cycles = table.total()
    out = torch.empty((n_cycles, signal.shape[1], length), dtype=torch.float32, device=signal.device)
    native.cut_cycles(signal, table.cycles, n_cycles, out, n_cycles_dev=None)
    return out


def duration_features(frames: torch.Tensor, fs: int = 1000, err_flag: torch.Tensor | None = None) -> torch.Tensor:
    """(n, 14) float64 features in the order of ``FEATURE_NAMES``; a zero denominator yields NaN
    and raises ``ERR_ZERO_DIVISION`` in ``err_flag`` (the reference raises ZeroDivisionError)."""
    _need_cuda(frames, "frames")
    n = frames.shape[0]
    out = torch.empty((n, 14), dtype=torch.float64, device=frames.device)
    native.duration_features(frames, n, fs, out, err_flag=err_flag)
    return out
