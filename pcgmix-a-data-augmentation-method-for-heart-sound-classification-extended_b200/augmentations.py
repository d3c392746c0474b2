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

# Visit the cycles in pairing-chain order (draws.processing_order) so that a cycle read as "partner" is still in L2
# when it is read as "itself": ~3 % of kernel time on batches larger than L2.  The walk is native code (20 us per 4096
# cycles, 2 us at the reference's batch of 64), so the drop-in call has it on; any order gives the same output.
use_processing_order = True

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


_plans = {}


def _plan_for(method: str):
    """``draws.parse_method_1d`` with the result remembered per method string (a training run calls
    ``augment`` with the same string every step; ``None`` and refusals are not cached)."""
    plan = _plans.get(method)
    if plan is None:
        plan = draws.parse_method_1d(method)
        if plan is not None:
            if len(_plans) > 64:
                _plans.clear()
            _plans[method] = plan
    return plan


# below this many knot ordinates per step NumPy itself draws them (a C call either way, and NumPy then leaves
# its global stream where the reference leaves it without a 30 us set_state)
_SMALL_DRAW = 16384


def _plain_step(plan, data, target_ohe, frames, step):
    """The common case — default same-label pairing, no ``(rand)`` / ``(mixAll)`` / name-based modifiers — with
    the host's integer work done by ONE native call straight into the pinned staging buffer
    (``pcgmix_host_prepare_step``) and the tables handed to the kernel as offsets into one device buffer.
    Returns ``(data_new, mix_indices)`` or ``None`` if this path does not apply (the general path then runs)."""
    batch, channels, length = data.shape
    magwarp = plan.branch == "durmixmagwarp"
    if not isinstance(frames, torch.Tensor) or frames.is_cuda or frames.dim() != 2 or frames.shape[0] != batch \
            or frames.shape[1] < 5 or not isinstance(step, (int, np.integer)) or not 0 <= step < 2 ** 32 \
            or not plan.alpha > 0.0:
        return None
    if magwarp and plan.knot > native.MAX_KNOT:
        raise ValueError(f"durmixmagwarp knot={plan.knot} exceeds the supported maximum {native.MAX_KNOT}")
    if frames.dtype != torch.int64:
        if frames.dtype.is_floating_point or frames.dtype == torch.bool:
            raise TypeError(f"frames must hold integers, got {frames.dtype}")
        frames = frames.to(torch.int64)
    f_np = frames.numpy()
    if f_np.strides[1] != 8 or f_np.strides[0] % 8 or f_np.strides[0] < 40:
        f_np = np.ascontiguousarray(f_np)
    labels = np.ascontiguousarray(labels_from_one_hot(target_ohe), dtype=np.int64)
    knot = plan.knot if magwarp else -1
    n_knots = batch * (plan.knot + 2) * channels if magwarp else 0
    slot = staging.reserve(data.device, batch * 28 + n_knots * 8 + 64)
    rc, info, mix_indices = native.host_prepare_step(labels, f_np, length, step, knot, channels, use_processing_order, slot.buf)
    if rc == 2:
        bad = int(info[5])
        raise ValueError(f"frames[{bad}] = {f_np[bad, :5].tolist()} is not a non-negative, non-decreasing int32 offset list")
    if rc == 3:
        b, s_, wd, ws = (int(v) for v in info[5:9])
        raise RuntimeError(
            f"cycle {b} (offsets {f_np[b, :5].tolist()}) cannot be blended with its partner {int(mix_indices[b])} "
            f"(offsets {f_np[int(mix_indices[b]), :5].tolist()}) in a row of {length} samples: state {s_} clamps to "
            f"{wd} destination and {ws} source samples (the reference raises a shape mismatch here)")
    if rc != 0:
        return None
    # lambda (and knots) from NumPy's legacy global stream, as the reference draws them
    if magwarp:
        if n_knots <= _SMALL_DRAW:
            lam = draws.draw_lambda(plan.alpha, step)
            knots = draws.draw_knots(batch, plan.knot, channels, plan.sigma)
        else:
            lam, knots = draws.lambda_and_knots(plan.alpha, step, batch, plan.knot, channels, plan.sigma)
            if prefetch_next_step:                   # step k+1's draws need nothing but the step count
                draws.prefetch_lambda_and_knots(plan.alpha, step + 1, batch, plan.knot, channels, plan.sigma)
        off = int(info[3])
        slot.buf.numpy()[off:off + n_knots * 8] = knots.reshape(-1).view(np.uint8)
    else:
        lam = draws.draw_lambda(plan.alpha, step)
    lam32, one_minus = draws.lambda_pair_fp32(lam)
    tables = staging.commit(slot, int(info[4]), data.device)
    data_new = torch.empty_like(data)
    if magwarp:
        pos_dev, mat_dev = _device_tables(length, plan.knot, data.device)
        native.mix1d_packed(data, data_new, tables, info, lam32, one_minus, mat_dev, pos_dev, plan.knot)
    else:
        native.mix1d_packed(data, data_new, tables, info, lam32, one_minus)
    return data_new, mix_indices


def augment(args, data, target_ohe, frames, wav, step_counter, model, device, RESULTS_ARGS):
    """PCGmix (``durratiomixup``) / PCGmix+ (``durmixmagwarp(sigma,knot)``) on a (B, C, L) batch.

    Returns ``(data_new, target_ohe, mix_indices, None)``; when the method is not one the
    reference implements, or the probability gate fails, returns ``(data, target_ohe, [], None)``
    with the very same objects (``augmentations.py:731-732, 938-939``)."""
    plan = _plan_for(args.method)
    if plan is None:
        return data, target_ohe, [], None
    step = step_counter.count
    # (the gate variate lies in [0, 1): with probability 1 the draw cannot fail and has no other effect)
    if plan.probability < 1.0 and draws.gate(step) >= plan.probability:
        return data, target_ohe, [], None

    data = require_cuda_batch(data, 3, "augment")
    batch, channels, length = data.shape
    if not (plan.rand_displacement or plan.mix_all or "(samePCG)" in args.method or "(sameDataset)" in args.method):
        done = _plain_step(plan, data, target_ohe, frames, step)
        if done is not None:
            return done[0], target_ohe, done[1], None
    labels = labels_from_one_hot(target_ohe)
    mix_indices = draws.pairing(args.method, labels, wav, step)
    knots = None
    if plan.branch == "durmixmagwarp":
        if plan.knot > native.MAX_KNOT:
            raise ValueError(f"durmixmagwarp knot={plan.knot} exceeds the supported maximum {native.MAX_KNOT}")
        lam, knots = draws.lambda_and_knots(plan.alpha, step, batch, plan.knot, channels, plan.sigma)
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
