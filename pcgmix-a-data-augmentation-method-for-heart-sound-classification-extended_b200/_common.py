"""Pieces shared by the 1D and 2D ``augment`` front ends."""
from __future__ import annotations

import numpy as np
import torch


def require_cuda_batch(data, ndim: int, what: str):
    if not isinstance(data, torch.Tensor):
        raise TypeError(f"{what}: data must be a torch.Tensor")
    if not data.is_cuda:
        raise RuntimeError(
            f"{what}: data lives on {data.device}; this implementation runs the PCGmix kernels on a CUDA "
            "device only (there is no CPU fallback) — move the batch to the GPU first, as train_model.py does")
    if data.dim() != ndim:
        raise ValueError(f"{what}: expected a {ndim}-D batch, got shape {tuple(data.shape)}")
    if data.dtype != torch.float32:
        raise TypeError(f"{what}: data must be float32 (the reference path is fp32), got {data.dtype}")
    return data if data.is_contiguous() else data.contiguous()


_label_slots = {}


def labels_from_one_hot(target_ohe) -> np.ndarray:
    """Class id per cycle, as the reference recovers it (augmentations.py:501): arg-max of the
    one-hot target, read back to the host.  On a CUDA tensor the read-back is a tiny kernel writing
    into pinned host memory followed by an event wait, not a ``.cpu()`` copy: a copy-engine transfer
    would wait behind any large device->host copy in flight (results of the previous step)."""
    idx = target_ohe.max(1, keepdim=True)[1].reshape(-1)
    if not idx.is_cuda:
        return idx.detach().numpy()
    from . import native
    idx = idx.to(torch.int64).contiguous()
    n = idx.shape[0]
    key = (idx.device.index, n)
    slot = _label_slots.get(key)
    if slot is None:
        if len(_label_slots) > 16:
            _label_slots.clear()
        slot = (torch.empty(n, dtype=torch.int64, pin_memory=True), torch.cuda.Event())
        _label_slots[key] = slot
    native.copy_small(slot[0], idx, n * 8)
    slot[1].record(torch.cuda.current_stream(idx.device))
    slot[1].synchronize()
    return slot[0].numpy().copy()


def host_frames(frames, batch: int, limit: int) -> np.ndarray:
    """Validate the CPU ``frames`` tensor and return it as int32 (B, 5).

    The reference slices with these offsets directly; offsets that are not monotone or that
    exceed the row length make it either raise a shape error or (through Python's negative-index
    wrap-around) blend unrelated samples.  Here they are rejected up front."""
    f = frames.detach().cpu().numpy() if isinstance(frames, torch.Tensor) else np.asarray(frames)
    if not np.issubdtype(f.dtype, np.integer):
        raise TypeError(f"frames must hold integers, got {f.dtype}")
    if f.ndim != 2 or f.shape[0] != batch or f.shape[1] < 5:
        raise ValueError(f"frames must be ({batch}, 5), got {f.shape}")
    f5 = f[:, :5].astype(np.int64)
    if (f5 < 0).any() or (f5 > limit).any() or (np.diff(f5, axis=1) < 0).any():
        bad = int(np.nonzero((f5 < 0).any(1) | (f5 > limit).any(1) | (np.diff(f5, axis=1) < 0).any(1))[0][0])
        raise ValueError(f"frames[{bad}] = {f5[bad].tolist()} is not a monotone offset list inside [0, {limit}]")
    return np.ascontiguousarray(f5.astype(np.int32))


def last_frame(frames) -> np.ndarray:
    """``f[-1]`` of every cycle (the beat length used by the 2D time masks)."""
    f = frames.detach().cpu().numpy() if isinstance(frames, torch.Tensor) else np.asarray(frames)
    return f[:, -1].astype(np.int64)
