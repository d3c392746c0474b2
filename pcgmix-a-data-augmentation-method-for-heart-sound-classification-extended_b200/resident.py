"""PCGmix / PCGmix+ on a data set that lives on the GPU as *recordings + cycle table*.

The reference builds, offline and on the host, one zero-padded row per cardiac cycle
(``databuilder.ipynb:627-632, :973-978``: ``seg_y = y_hat[start:stop]; seg_y.resize(L)``), stacks
the rows into an ``(n, C, L)`` array (``dataloader_physionet.py:43-48``), and every training step
uploads a batch of them (``train_model.py:499``) before ``augment`` mixes it.  With PhysioNet-shaped
cycles about 60 % of that array is padding.

Here the recordings stay on the device as they are, the cycle table written by the segmentation
kernels says where every cycle lies, and ONE kernel (``pcgmix_mix1d_resident``) cuts, pads, mixes
and warps a batch of table rows: the padded array is never materialised and a step moves
``4*C*(len1 + M + L)`` bytes instead of ``4*C*(len + L) + 4*C*(2L + M)``.

``augment`` below has the reference's parameters with ``data``/``frames`` replaced by the resident
set and the batch's table rows; its result is bit-identical to

    augmentations.augment(args, resident.padded(ids), target_ohe, resident.frames_of(ids), ...)
"""
from __future__ import annotations

import dataclasses

import numpy as np
import torch

from . import draws, native, segmentation, staging
from ._common import labels_from_one_hot
from .augmentations import _SMALL_DRAW, _device_tables, _plan_for

__all__ = ["ResidentCycles", "from_dense_states", "from_state_table", "augment", "mix_rows"]


@dataclasses.dataclass
class ResidentCycles:
    signal: torch.Tensor                 # (n_rec, C, T) float32, CUDA, 16-byte aligned
    table: segmentation.CycleTable       # cycles (capacity, 8) int32
    length: int                          # L: row length of an augmented cycle
    n_cycles: int                        # valid table rows
    err_flag: torch.Tensor               # (1,) int32 device flag word of the mix launches

    @property
    def channels(self) -> int:
        return self.signal.shape[1]

    def padded(self, ids=None) -> torch.Tensor:
        """The reference's layout, for comparison and for the gate-fail path: cycles ``ids`` (default:
        all) cut out and zero-padded to ``length`` -> (n, C, L)."""
        if ids is None:
            return segmentation.cut_cycles(self.signal, self.table, self.length, self.n_cycles)
        rows = self.table.cycles[_ids_on_device(ids, self.signal.device).long()].contiguous()
        sub = segmentation.CycleTable(rows, self.table.row_ptr, self.table.err_flag)
        return segmentation.cut_cycles(self.signal, sub, self.length, rows.shape[0])

    def frames_of(self, ids) -> torch.Tensor:
        """(B, 5) int64 CPU offsets of table rows ``ids`` — what the reference's loader would hand to
        ``augment`` as ``frames`` (device -> host read; synchronises)."""
        return self.table.cycles[_ids_on_device(ids, self.signal.device).long(), 3:].cpu().to(torch.int64)

    def check(self):
        """Raise if a mix launch met a table row, recording or partner out of range, or offsets the
        reference could not have blended (synchronises)."""
        flag = int(self.err_flag.item())
        if flag & native.ERR_BAD_PARTNER:
            raise IndexError("a cycle id, recording index or partner index was out of range")
        if flag & native.ERR_BAD_FRAMES:
            raise ValueError(f"a cycle's state offsets are decreasing, or its windows and its partner's clamp to unequal "
                             f"widths in a row of {self.length} samples (the reference raises a shape mismatch there)")
        return self


def _ids_on_device(ids, device) -> torch.Tensor:
    if isinstance(ids, torch.Tensor):
        if ids.dtype not in (torch.int32, torch.int64, torch.int16, torch.uint8, torch.int8):
            raise TypeError(f"cycle ids must be integers, got {ids.dtype}")
        return ids.to(device=device, dtype=torch.int32).contiguous()
    arr = np.asarray(ids)
    if not np.issubdtype(arr.dtype, np.integer):
        raise TypeError(f"cycle ids must be integers, got {arr.dtype}")
    return staging.upload([np.ascontiguousarray(arr.astype(np.int32))], device)[0]


def _wrap(signal, table, length) -> ResidentCycles:
    if signal.dtype != torch.float32 or signal.dim() != 3 or not signal.is_cuda:
        raise TypeError("signal must be a CUDA (n_rec, C, T) float32 tensor")
    signal = signal.contiguous()
    if signal.data_ptr() % 16:
        signal = signal.clone()
    table.check()
    err = torch.zeros(1, dtype=torch.int32, device=signal.device)
    return ResidentCycles(signal, table, int(length), table.total(), err)


def from_dense_states(signal: torch.Tensor, states: torch.Tensor, length: int, downsample: int = 1) -> ResidentCycles:
    """Recordings ``signal`` (n_rec, C, T) at the rate of the cycle offsets, dense Springer states
    ``states`` (n_rec, T*downsample) int8 (``databuilder.ipynb`` cell 14)."""
    return _wrap(signal, segmentation.cycles_from_dense_states(states, downsample), length)


def from_state_table(signal: torch.Tensor, positions, codes, rec_offsets, length: int, downsample: int = 1) -> ResidentCycles:
    """Recordings plus concatenated (position, state) transition tables (``databuilder.ipynb`` cell 25)."""
    return _wrap(signal, segmentation.cycles_from_state_table(positions, codes, rec_offsets, downsample), length)


def mix_rows(resident: ResidentCycles, sel_dev, mix_dev, lam32, one_minus_lam32, knots_dev=None, knot=0,
             order_dev=None, out=None, scratch=None) -> torch.Tensor:
    """Device-resident entry, everything already on the GPU (``sel_dev`` may be None for "rows 0..B-1").
    ``scratch``: (B, 8) int32 work space for the slot records of the pipelined kernel; allocated from
    PyTorch's caching allocator when not given."""
    batch = mix_dev.shape[0]
    if scratch is None:
        scratch = torch.empty((batch, 8), dtype=torch.int32, device=resident.signal.device)
    if out is None:
        out = torch.empty((batch, resident.channels, resident.length), dtype=torch.float32, device=resident.signal.device)
    pos_dev = mat_dev = None
    if knots_dev is not None:
        if knot > native.MAX_KNOT:
            raise ValueError(f"durmixmagwarp knot={knot} exceeds the supported maximum {native.MAX_KNOT}")
        pos_dev, mat_dev = _device_tables(resident.length, knot, resident.signal.device)
    native.mix1d_resident(resident.signal, resident.table.cycles, sel_dev, mix_dev, lam32, one_minus_lam32, out,
                          knots_dev, mat_dev, pos_dev, knot, order=order_dev, err_flag=resident.err_flag, scratch=scratch)
    return out


def augment(args, resident: ResidentCycles, cycle_ids, target_ohe, wav, step_counter, model, device, RESULTS_ARGS):
    """``augmentations.augment`` for a batch given as table rows ``cycle_ids`` of a resident set.

    Returns ``(data_new, target_ohe, mix_indices, None)`` with ``data_new`` (B, C, L) on the device.
    When the method is not a PCGmix one or the probability gate fails the batch is returned
    un-augmented (cut + padded), with ``mix_indices = []`` like the reference (``augmentations.py:938-939``)."""
    plan = _plan_for(args.method)
    step = step_counter.count
    if plan is None or (plan.probability < 1.0 and draws.gate(step) >= plan.probability):
        return resident.padded(cycle_ids), target_ohe, [], None
    if plan.rand_displacement:
        raise NotImplementedError("the (rand) displacement variant needs the offsets on the host; use "
                                  "augmentations.augment on resident.padded(ids)")
    sel_host = None if isinstance(cycle_ids, torch.Tensor) and cycle_ids.is_cuda else np.asarray(cycle_ids)
    batch = int(cycle_ids.shape[0]) if sel_host is None else int(sel_host.shape[0])
    labels = labels_from_one_hot(target_ohe)
    if labels.shape[0] != batch:
        raise ValueError(f"{batch} cycle ids but {labels.shape[0]} targets")
    mix_indices = draws.pairing(args.method, labels, wav, step)
    uploads = [mix_indices.astype(np.int32)]
    if plan.branch == "durmixmagwarp":
        if plan.knot > native.MAX_KNOT:
            raise ValueError(f"durmixmagwarp knot={plan.knot} exceeds the supported maximum {native.MAX_KNOT}")
        if batch * (plan.knot + 2) * resident.channels <= _SMALL_DRAW:
            # small draws: NumPy itself (a C call either way, and its global stream ends where the reference leaves it)
            lam = draws.draw_lambda(plan.alpha, step)
            knots = draws.draw_knots(batch, plan.knot, resident.channels, plan.sigma)
        else:
            lam, knots = draws.lambda_and_knots(plan.alpha, step, batch, plan.knot, resident.channels, plan.sigma)
            draws.prefetch_lambda_and_knots(plan.alpha, step + 1, batch, plan.knot, resident.channels, plan.sigma)
        uploads.append(knots)
    else:
        lam = draws.draw_lambda(plan.alpha, step)
    lam32, one_minus = draws.lambda_pair_fp32(lam)
    if sel_host is not None:
        if not np.issubdtype(sel_host.dtype, np.integer):
            raise TypeError(f"cycle ids must be integers, got {sel_host.dtype}")
        uploads.append(np.ascontiguousarray(sel_host.astype(np.int32)))
    on_dev = staging.upload(uploads, resident.signal.device)
    sel_dev = on_dev[-1] if sel_host is not None else _ids_on_device(cycle_ids, resident.signal.device)
    knots_dev = on_dev[1] if plan.branch == "durmixmagwarp" else None
    data_new = mix_rows(resident, sel_dev, on_dev[0], lam32, one_minus, knots_dev, plan.knot)

    if plan.mix_all:
        lams = torch.from_numpy(np.array(np.ones(batch) * lam).astype("float32")).to(resident.signal.device)
        lams_target = lams[:, None]
        target_ohe = target_ohe * lams_target + target_ohe[mix_indices] * (1 - lams_target)
    return data_new, target_ohe, mix_indices, None
