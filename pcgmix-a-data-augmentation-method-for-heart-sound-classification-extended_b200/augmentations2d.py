"""Drop-in for the reference's ``augmentations2d.augment`` on the PCGmix branches.

Covers ``durratiomixup`` (``augmentations2d.py:397-427``) and the three composites that zero a
seeded box after the mix: ``durmixtimemask``, ``durmixfreqmask``, ``durmixcutout``
(``:286-395``).  One kernel launch per step (``pcgmix_mix2d``), box included.

Deliberate difference from the reference: it allocates its output as
``(B, Ch, shape[2], shape[2])`` (``:409``), so it only works for square spectrograms; this
module uses the true time dimension, which gives the same result for square inputs and makes
non-square ones (e.g. 64 x 250) work.
"""
from __future__ import annotations

import numpy as np
import torch

from . import draws, native, staging
from ._common import check_pair_windows, host_frames, labels_from_one_hot, last_frame, require_cuda_batch

__all__ = ["augment"]


def augment(args, data, target_ohe, frames, wav, step_counter, model, device, RESULTS_ARGS):
    plan = draws.parse_method_2d(args.method)
    if plan is None:
        return data, target_ohe, [], None
    step = step_counter.count
    if draws.gate(step) >= plan.probability:
        return data, target_ohe, [], None

    data = require_cuda_batch(data, 4, "augment (2D)")
    batch, _, n_freq, n_time = data.shape
    labels = labels_from_one_hot(target_ohe)
    mix_indices = draws.same_label_pairing(labels, step)
    lam32, one_minus = draws.lambda_pair_fp32(draws.draw_lambda(1, step))

    frames_i32 = host_frames(frames, batch, n_time)
    check_pair_windows(frames_i32, mix_indices, n_time)
    uploads = [frames_i32, mix_indices.astype(np.int32)]
    h1 = h2 = 0
    has_tbox = False
    if plan.branch in ("durmixtimemask", "durmixcutout"):
        gap, frac1 = draws.mask_geometry(step, plan.time_region_max)
        beat = last_frame(frames)
        # int(frac * beat_len) per cycle: float64 product truncated toward zero (:358-360)
        tbox = np.stack([(frac1 * beat).astype(np.int64), ((frac1 + gap) * beat).astype(np.int64)], axis=1)
        uploads.append(np.ascontiguousarray(np.clip(tbox, 0, n_time).astype(np.int32)))
        has_tbox = True
        h1, h2 = 0, n_freq
    if plan.branch in ("durmixfreqmask", "durmixcutout"):
        gap, frac1 = draws.mask_geometry(step, plan.freq_region_max)
        h1 = int(n_freq * frac1)
        h2 = min(n_freq, h1 + int(gap * n_freq))
    on_dev = staging.upload(uploads, data.device)
    data_new = torch.empty_like(data)
    native.mix2d(data, data_new, on_dev[0], on_dev[1], lam32, one_minus,
                 tbox=on_dev[2] if has_tbox else None, h1=h1, h2=h2)
    return data_new, target_ohe, mix_indices, None
