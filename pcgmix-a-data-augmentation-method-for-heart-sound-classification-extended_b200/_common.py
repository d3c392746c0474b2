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


def with_host_labels(target_ohe, target):
    """Attach the loader's CPU ``target`` (class id per cycle, what ``F.one_hot`` was built from:
    train_model.py:500-501) to the device one-hot tensor and return that tensor.

    The reference recovers the class ids from the DEVICE tensor every step (``target_ohe.max(1)[1].cpu()``,
    augmentations.py:501), which makes the host wait for everything queued on the GPU.  A caller that
    still has the CPU ``target`` can hand it over this way; ``augment`` then pairs from it without any
    device synchronisation.  The ids must be what the arg-max of ``target_ohe`` would give."""
    ids = target.detach().cpu().numpy() if isinstance(target, torch.Tensor) else np.asarray(target)
    if ids.ndim != 1 or ids.shape[0] != target_ohe.shape[0] or not np.issubdtype(ids.dtype, np.integer):
        raise ValueError("target must be a 1-D integer array with one class id per cycle")
    target_ohe.pcgmix_host_labels = ids.astype(np.int64)
    return target_ohe


def labels_from_one_hot(target_ohe) -> np.ndarray:
    """Class id per cycle, as the reference recovers it (augmentations.py:501): arg-max of the
    one-hot target, read back to the host.  On a CUDA tensor the read-back is a tiny kernel writing
    into pinned host memory followed by an event wait, not a ``.cpu()`` copy: a copy-engine transfer
    would wait behind any large device->host copy in flight (results of the previous step).  No
    device work at all when the caller attached the CPU ids (:func:`with_host_labels`)."""
    attached = getattr(target_ohe, "pcgmix_host_labels", None)
    if attached is not None and attached.shape[0] == target_ohe.shape[0]:
        return attached
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

    Offsets must be non-negative and non-decreasing (a negative or decreasing offset makes the
    reference blend unrelated samples through Python's negative-index wrap-around; refused here).
    Offsets BEYOND the row length are accepted, because the reference accepts them: the data builder
    keeps cycles longer than the padded length (databuilder.ipynb cells 14/25 print a warning,
    ``resize`` truncates the samples, the offsets stay) and the reference's slices are clamped to the
    row.  Whether such a cycle can be blended depends on its partner: see :func:`check_pair_windows`."""
    f = frames.detach().cpu().numpy() if isinstance(frames, torch.Tensor) else np.asarray(frames)
    if not np.issubdtype(f.dtype, np.integer):
        raise TypeError(f"frames must hold integers, got {f.dtype}")
    if f.ndim != 2 or f.shape[0] != batch or f.shape[1] < 5:
        raise ValueError(f"frames must be ({batch}, 5), got {f.shape}")
    f5 = f[:, :5].astype(np.int64)
    if (f5 < 0).any() or (f5 > 2 ** 31 - 1).any() or (np.diff(f5, axis=1) < 0).any():
        bad = int(np.nonzero((f5 < 0).any(1) | (f5 > 2 ** 31 - 1).any(1) | (np.diff(f5, axis=1) < 0).any(1))[0][0])
        raise ValueError(f"frames[{bad}] = {f5[bad].tolist()} is not a non-negative, non-decreasing int32 offset list")
    return np.ascontiguousarray(f5.astype(np.int32))


def clamped_windows(f1: np.ndarray, f2: np.ndarray, limit: int):
    """Per (cycle, state): the reference's blend slices after Python's slice clamping to the row
    length (augmentations.py:289-304).  Returns ``(dst_start, src_start, dst_width, src_width)``."""
    f1 = np.asarray(f1, dtype=np.int64)
    f2 = np.asarray(f2, dtype=np.int64)
    n0 = np.minimum(np.diff(f1, axis=1), np.diff(f2, axis=1))
    a1, a2 = np.minimum(f1[:, :4], limit), np.minimum(f2[:, :4], limit)
    return a1, a2, np.minimum(f1[:, :4] + n0, limit) - a1, np.minimum(f2[:, :4] + n0, limit) - a2


def check_pair_windows(frames_i32: np.ndarray, mix: np.ndarray, limit: int) -> None:
    """Raise where the reference raises: a cycle whose offsets run past the row length can only be
    blended with a partner if, state by state, the clamped slices of both have the same width (or an
    empty destination meets a one-sample source, which broadcasts to nothing).  Any other combination
    is a shape mismatch in the reference's tensor expression (``RuntimeError``) — except a one-sample
    source against a wider destination, which torch broadcasts over the whole window; that is refused
    here too.  Batches whose offsets all lie inside the row (the normal case) return immediately."""
    if frames_i32.size == 0 or int(frames_i32.max()) <= limit:
        return
    f1 = frames_i32[:, :5]
    _, _, wd, ws = clamped_windows(f1, f1[np.asarray(mix, dtype=np.int64)], limit)
    bad = (wd != ws) & ~((wd == 0) & (ws == 1))
    if bad.any():
        b, s = (int(v[0]) for v in np.nonzero(bad))
        raise RuntimeError(
            f"cycle {b} (offsets {f1[b].tolist()}) cannot be blended with its partner {int(mix[b])} "
            f"(offsets {f1[int(mix[b])].tolist()}) in a row of {limit} samples: state {s} clamps to "
            f"{int(wd[b, s])} destination and {int(ws[b, s])} source samples (the reference raises a shape mismatch here)")


def last_frame(frames) -> np.ndarray:
    """``f[-1]`` of every cycle (the beat length used by the 2D time masks)."""
    f = frames.detach().cpu().numpy() if isinstance(frames, torch.Tensor) else np.asarray(frames)
    return f[:, -1].astype(np.int64)
