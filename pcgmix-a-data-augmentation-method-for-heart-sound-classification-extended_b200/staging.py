"""Host -> device upload of the small per-step arrays (frames, pairing, order, knots).

All arrays of one step are packed into ONE pinned host buffer and moved with ONE small kernel on
the current stream (it reads the pinned buffer through unified addressing); the device side is one allocation sliced into typed views.  No
``cudaMalloc`` per call (PyTorch's caching allocator), no host synchronisation except when a
pinned slot is about to be reused while its previous copy is still in flight (a ring of slots
makes that rare).
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch

from . import native

_ALIGN = 16
_RING = 4


class _Slot:
    def __init__(self):
        self.buf = None
        self.event = None


class Uploader:
    def __init__(self):
        self._slots = {}
        self._next = {}

    def _slot(self, device: torch.device, nbytes: int) -> _Slot:
        key = (device.type, device.index)
        ring = self._slots.setdefault(key, [_Slot() for _ in range(_RING)])
        i = self._next.get(key, 0)
        self._next[key] = (i + 1) % _RING
        slot = ring[i]
        if slot.event is not None:
            slot.event.synchronize()          # previous copy out of this slot must have finished
        if slot.buf is None or slot.buf.numel() < nbytes:
            slot.buf = torch.empty(max(nbytes, 1 << 16), dtype=torch.uint8, pin_memory=True)
        return slot

    def reserve(self, device: torch.device, nbytes: int) -> _Slot:
        """A pinned staging slot of at least ``nbytes`` for the caller to fill (``slot.buf``); hand it to
        :meth:`commit` afterwards."""
        return self._slot(device, nbytes)

    def commit(self, slot: _Slot, nbytes: int, device: torch.device) -> torch.Tensor:
        """Upload the first ``nbytes`` of a filled slot; returns the device copy (uint8)."""
        nbytes = (max(int(nbytes), _ALIGN) + _ALIGN - 1) // _ALIGN * _ALIGN
        dev_buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        native.copy_small_raw(dev_buf.data_ptr(), slot.buf.data_ptr(), nbytes, device)
        if slot.event is None:
            slot.event = torch.cuda.Event()
        if device.index is None or device.index == torch.cuda.current_device():
            slot.event.record()                            # current stream of the current device: no Stream object built
        else:
            slot.event.record(torch.cuda.current_stream(device))
        return dev_buf

    def upload(self, arrays: Sequence[np.ndarray], device: torch.device) -> List[torch.Tensor]:
        """Copy ``arrays`` (C-contiguous ndarrays) to ``device``; returns tensors of the same
        shapes/dtypes, all views into one device allocation."""
        offsets, total = [], 0
        for a in arrays:
            total = (total + _ALIGN - 1) // _ALIGN * _ALIGN
            offsets.append(total)
            total += a.nbytes
        total = (max(total, _ALIGN) + _ALIGN - 1) // _ALIGN * _ALIGN
        slot = self._slot(device, total)
        host = slot.buf.numpy()
        for a, off in zip(arrays, offsets):
            host[off:off + a.nbytes] = np.ascontiguousarray(a).reshape(-1).view(np.uint8)
        dev_buf = torch.empty(total, dtype=torch.uint8, device=device)
        # by a kernel reading the pinned buffer, not by the copy engine: a prefetching loader keeps
        # large batch copies in flight on that engine and this upload must not wait behind them
        native.copy_small(dev_buf, slot.buf, total)
        if slot.event is None:
            slot.event = torch.cuda.Event()
        slot.event.record(torch.cuda.current_stream(device))
        out = []
        for a, off in zip(arrays, offsets):
            t = dev_buf[off:off + a.nbytes].view(torch.from_numpy(np.empty(0, a.dtype)).dtype)
            out.append(t.view(a.shape))
        return out


_default = Uploader()


def upload(arrays: Sequence[np.ndarray], device) -> List[torch.Tensor]:
    return _default.upload(arrays, torch.device(device))


def reserve(device, nbytes: int) -> _Slot:
    return _default.reserve(torch.device(device), nbytes)


def commit(slot: _Slot, nbytes: int, device) -> torch.Tensor:
    return _default.commit(slot, nbytes, torch.device(device))
