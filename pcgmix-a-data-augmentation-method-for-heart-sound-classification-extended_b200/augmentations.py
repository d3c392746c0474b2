"""Drop-in for the reference's ``augmentations.augment`` on the PCGmix / PCGmix+ branches.

Same nine positional parameters, same 4-tuple result as ``augmentations.py:698`` of the
reference; ``train_model.py:507`` can import this module instead of the original.  What runs
where:

  host (Python)  method-string parse, probability gate, pairing, lambda, knot draw — replayed
                 from ``step_counter.count`` exactly like the reference (see ``draws.py``)
  device (CUDA)  everything that touches the samples: one fused kernel launch per step
                 (``pcgmix_mix1d`` or ``pcgmix_mix1d_magwarp`` of ``include/pcgmix_b200.h``)

The reference loops over the batch in Python (``augmentations.py:969-977``) and, for PCGmix+,
takes the batch to the host for a SciPy spline per cycle and channel and back
(``:924-928``); here the batch never leaves the GPU and is read once and written once.
"""
from __future__ import annotations

import numpy as np
import torch

from . import draws, native, spline, staging
from ._common import check_pair_windows, host_frames, labels_from_one_hot, require_cuda_batch, with_host_labels

__all__ = ["augment", "pcgmix_on_device", "prepare_on_device", "with_host_labels"]

# Visit the cycles in pairing-chain order (draws.processing_order) so that a cycle read as
# "partner" is still in L2 when it is read as "itself".  Worth ~3 % of kernel time when batches are
# device-resident; it costs ~1 ms of host Python per 4096 cycles, so the per-step drop-in call, which
# is bound by the host and by PCIe, leaves it off unless asked.
use_processing_order = False

# PCGmix+: start drawing step k+1's lambda and knots on a worker thread while step k runs (the seed is the step
# count).  Pure overlap; results and NumPy's global stream are the same with it on or off.
prefetch_next_step = True

_table_cache = {}


def _device_tables(length: int, knot: int, device):
    """knot positions and coefficient matrix as device tensors, cached per (L, knot, device)."""
    key = (int(length), int(knot), str(device))
    hit = _table_cache.get(key)
    if hit is None:
        pos, mat = spline.magwarp_tables(length, knot)
        hit = (torch.from_numpy(np.array(pos)).to(device), torch.from_numpy(np.array(mat)).to(device))
        if len(_table_cache) > 32:
            _table_cache.clear()
        _table_cache[key] = hit
    return hit


def pcgmix_on_device(data, frames_dev, mix_dev, lam32, one_minus_lam32, knots_dev=None, knot=None,
                     order_dev=None, out=None, err_flag=None, windows_dev=None):
    """Device-resident entry: everything already on the GPU (int32 frames / pairing / order,
    float64 knots).  Launches exactly one kernel on the current stream and returns ``out``.
    ``windows_dev`` (B, 4, 3) replaces ``frames_dev`` for the ``(rand)`` displacement variant."""
    if out is None:
        out = torch.empty_like(data)
    if windows_dev is not None:
        if knots_dev is None:
            native.mix1d_windows(data, out, windows_dev, mix_dev, lam32, one_minus_lam32, order=order_dev,
                                 err_flag=err_flag)
        else:
            pos_dev, mat_dev = _device_tables(data.shape[2], knot, data.device)
            native.mix1d_windows(data, out, windows_dev, mix_dev, lam32, one_minus_lam32, knots_dev, mat_dev, pos_dev,
                                 knot, order=order_dev, err_flag=err_flag)
        return out
    if knots_dev is None:
        native.mix1d(data, out, frames_dev, mix_dev, lam32, one_minus_lam32, order=order_dev, err_flag=err_flag)
    else:
        if knot > native.MAX_KNOT:
            raise ValueError(f"durmixmagwarp knot={knot} exceeds the supported maximum {native.MAX_KNOT}")
        pos_dev, mat_dev = _device_tables(data.shape[2], knot, data.device)
        native.mix1d_magwarp(data, out, frames_dev, mix_dev, lam32, one_minus_lam32, knots_dev, mat_dev,
                             pos_dev, knot, order=order_dev, err_flag=err_flag)
    return out


def prepare_on_device(data, frames_dev, mix_dev, lam32, one_minus_lam32, out, knots_dev=None, knot=None,
                      order_dev=None, err_flag=None):
    """Same launch as :func:`pcgmix_on_device`, with all arguments resolved once: returns an object
    whose ``launch()`` costs one foreign call.  For sweeps over resident batches."""
    if knots_dev is None:
        return native.PreparedMix1D(data, out, frames_dev, mix_dev, lam32, one_minus_lam32, order=order_dev,
                                    err_flag=err_flag)
    if knot > native.MAX_KNOT:
        raise ValueError(f"durmixmagwarp knot={knot} exceeds the supported maximum {native.MAX_KNOT}")
    pos_dev, mat_dev = _device_tables(data.shape[2], knot, data.device)
    return native.PreparedMix1D(data, out, frames_dev, mix_dev, lam32, one_minus_lam32, knots_dev, mat_dev, pos_dev,
                                knot, order=order_dev, err_flag=err_flag)


def augment(args, data, target_ohe, frames, wav, step_counter, model, device, RESULTS_ARGS):
    """PCGmix (``durratiomixup``) / PCGmix+ (``durmixmagwarp(sigma,knot)``) on a (B, C, L) batch.

    Returns ``(data_new, target_ohe, mix_indices, None)``; when the method is not one the
    reference implements, or the probability gate fails, returns ``(data, target_ohe, [], None)``
    with the very same objects (``augmentations.py:731-732, 938-939``)."""
    plan = draws.parse_method_1d(args.method)
    if plan is None:
        return data, target_ohe, [], None
    step = step_counter.count
    if draws.gate(step) >= plan.probability:
        return data, target_ohe, [], None

    data = require_cuda_batch(data, 3, "augment")
    batch, channels, length = data.shape
    labels = labels_from_one_hot(target_ohe)
    mix_indices = draws.pairing(args.method, labels, wav, step)
    knots = None
    if plan.branch == "durmixmagwarp":
        if plan.knot > native.MAX_KNOT:
            raise ValueError(f"durmixmagwarp knot={plan.knot} exceeds the supported maximum {native.MAX_KNOT}")
        lam, knots = draws.lambda_and_knots(plan.alpha, step, batch, plan.knot, channels, plan.sigma)
        if prefetch_next_step:                       # step k+1's draws need nothing but the step count
            draws.prefetch_lambda_and_knots(plan.alpha, step + 1, batch, plan.knot, channels, plan.sigma)
    else:
        lam = draws.draw_lambda(plan.alpha, step)
    lam32, one_minus = draws.lambda_pair_fp32(lam)

    frames_i32 = host_frames(frames, batch, length)
    if not plan.rand_displacement:
        check_pair_windows(frames_i32, mix_indices, length)
    uploads = [draws.rand_windows(frames_i32, mix_indices, step, length) if plan.rand_displacement else frames_i32,
               mix_indices.astype(np.int32),
               draws.processing_order(mix_indices) if use_processing_order else np.zeros(0, np.int32)]
    if knots is not None:
        uploads.append(knots)
    on_dev = staging.upload(uploads, data.device)
    knots_dev = on_dev[3] if plan.branch == "durmixmagwarp" else None
    data_new = pcgmix_on_device(data, None if plan.rand_displacement else on_dev[0], on_dev[1], lam32, one_minus,
                                knots_dev, plan.knot, order_dev=on_dev[2] if use_processing_order else None,
                                windows_dev=on_dev[0] if plan.rand_displacement else None)

    if plan.mix_all:
        # soft labels, as augmentations.py:915-917 / :978-980
        lams = torch.from_numpy(np.array(np.ones(batch) * lam).astype("float32")).to(data.device)
        lams_target = lams[:, None]
        target_ohe = target_ohe * lams_target + target_ohe[mix_indices] * (1 - lams_target)
    return data_new, target_ohe, mix_indices, None
