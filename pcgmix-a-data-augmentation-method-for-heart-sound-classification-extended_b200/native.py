"""ctypes binding of ``csrc/libpcgmix_b200.so`` — the C ABI declared in ``include/pcgmix_b200.h``.

This is the only place where Python meets the CUDA kernels.  PyTorch is used for device
memory and streams only: every call passes raw device pointers (``tensor.data_ptr()``) and the
current CUDA stream handle, and the library neither allocates nor synchronises.

There is no fallback: if the library is missing or a tensor is not a contiguous CUDA tensor of
the expected dtype, the call raises.
"""
from __future__ import annotations

import ctypes
import os
import threading

import torch

from . import build_native

ERR_BAD_PARTNER = 1
ERR_BAD_FRAMES = 2
ERR_BAD_PATTERN = 4
ERR_OVERFLOW = 8
ERR_ZERO_DIVISION = 16
ERR_EMPTY_STATE = 32
CYCLE_FEATURES = 36
CYCLE_PSD_FEATURES = 80
CYCLE_MOMENT_FEATURES = 10
MAX_KNOT = 30
MAX_FIRST_BLOCK_FILTERS = 512

_c_i32 = ctypes.c_int32
_c_f32 = ctypes.c_float
_ptr = ctypes.c_void_p

# name -> argtypes; restype is int for all but pcgmix_last_error.  Mirrors include/pcgmix_b200.h
# (tests/test_host_logic.py checks the two against each other).
SIGNATURES = {
    "pcgmix_version": [],
    "pcgmix_last_error": [],
    "pcgmix_device_info": [_ptr, _ptr, _ptr],
    "pcgmix_set_launch_overlap": [_c_i32],
    "pcgmix_overlap_launches": [],
    "pcgmix_set_spline_precision": [_c_i32],
    "pcgmix_get_spline_precision": [],
    "pcgmix_set_tuning": [_c_i32, _c_i32, _c_i32, _c_i32, _c_i32, _c_i32, _c_i32],
    "pcgmix_mix1d": [_ptr, _ptr, _ptr, _c_i32, _ptr, _ptr, _c_f32, _c_f32, _c_i32, _c_i32, _c_i32, _ptr, _ptr],
    "pcgmix_mix1d_magwarp": [_ptr, _ptr, _ptr, _c_i32, _ptr, _ptr, _c_f32, _c_f32, _ptr, _ptr, _ptr,
                             _c_i32, _c_i32, _c_i32, _c_i32, _ptr, _ptr],
    "pcgmix_mix1d_windows": [_ptr, _ptr, _ptr, _ptr, _ptr, _c_f32, _c_f32, _ptr, _ptr, _ptr,
                             _c_i32, _c_i32, _c_i32, _c_i32, _ptr, _ptr],
    "pcgmix_mix2d": [_ptr, _ptr, _ptr, _c_i32, _ptr, _ptr, _c_f32, _c_f32, _c_i32, _c_i32, _c_i32, _c_i32,
                     _ptr, _c_i32, _c_i32, _ptr, _ptr],
    "pcgmix_segment_dense": [_ptr, _c_i32, _c_i32, _c_i32, _ptr, _c_i32, _ptr, _ptr, _ptr],
    "pcgmix_segment_table": [_ptr, _ptr, _ptr, _c_i32, _c_i32, _c_i32, _ptr, _ptr, _c_i32, _ptr, _ptr, _ptr],
    "pcgmix_cut_cycles": [_ptr, _c_i32, _c_i32, _c_i32, _ptr, _c_i32, _ptr, _ptr, _c_i32, _ptr],
    "pcgmix_duration_features": [_ptr, _c_i32, _c_i32, _c_i32, _ptr, _ptr, _ptr],
    "pcgmix_cycle_features": [_ptr, _ptr, _c_i32, _c_i32, _c_i32, _c_i32, _c_i32, _c_i32, _ptr, _ptr, _ptr],
    "pcgmix_cycle_psd_features": [_ptr, _ptr, _c_i32, _c_i32, _c_i32, _c_i32, _c_i32, _c_i32, _ptr, _ptr, _ptr],
    "pcgmix_cycle_moment_features": [_ptr, _ptr, _c_i32, _c_i32, _c_i32, _c_i32, _c_i32, _ptr, _ptr, _ptr],
    "pcgmix_first_conv_block_workspace": [_c_i32, _c_i32],
    "pcgmix_first_conv_block": [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _c_i32, _c_i32, _c_i32, _c_i32, _c_i32,
                                ctypes.c_double, ctypes.c_double, _ptr, _ptr, _ptr],
    "pcgmix_mix1d_resident": [_ptr, _c_i32, _c_i32, _c_i32, _ptr, _c_i32, _ptr, _ptr, _ptr, _c_f32, _c_f32,
                              _ptr, _ptr, _ptr, _c_i32, _ptr, _c_i32, _c_i32, _ptr, _ptr, _ptr],
    "pcgmix_copy_small": [_ptr, _ptr, ctypes.c_int64, _ptr],
    "pcgmix_host_group_permutation": [_ptr, ctypes.c_int64, ctypes.c_int64, ctypes.c_uint64, _ptr],
    "pcgmix_host_lambda_knots": [ctypes.c_uint64, ctypes.c_double, ctypes.c_double, ctypes.c_int64, _c_i32,
                                 _ptr, _ptr, _ptr, _ptr, _ptr, _ptr],
    "pcgmix_host_processing_order": [_ptr, ctypes.c_int64, _ptr],
    "pcgmix_host_prepare_step": [_ptr, ctypes.c_int64, _ptr, ctypes.c_int64, ctypes.c_int64, ctypes.c_uint64, _c_i32,
                                 _c_i32, _c_i32, _ptr, ctypes.c_int64, _ptr, _ptr],
}

_lib = None
_lock = threading.Lock()
launch_count = 0     # kernels launched through this binding (bench.py reports it)


class NativeLibraryError(RuntimeError):
    pass


def library_path() -> str:
    # PCGMIX_LIB=<path>: a differently built library of the same ABI (kernel experiments: benchmarks/build_variant.py)
    override = os.environ.get("PCGMIX_LIB")
    if override:
        if not os.path.exists(override):
            raise NativeLibraryError(f"PCGMIX_LIB points to {override}, which does not exist")
        return override
    # PCGMIX_PROFILING_LIB=1: the build with the skip switches compiled in (profiling scripts only)
    if os.environ.get("PCGMIX_PROFILING_LIB") == "1" and os.path.exists(build_native.PROFILING_LIB_PATH):
        return build_native.PROFILING_LIB_PATH
    return build_native.LIB_PATH


def load(build_if_missing: bool = True):
    """Load the shared library (once).  If it has not been built yet it is compiled in-tree with
    nvcc first; if that is impossible, raises ``NativeLibraryError`` — there is no other path."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = library_path()
        if not os.path.exists(path):
            try:
                if not build_if_missing:
                    raise RuntimeError("automatic build disabled")
                build_native.build()
            except Exception as exc:
                raise NativeLibraryError(
                    f"{path} is missing and could not be built ({exc}): build it with "
                    "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc).  "
                    "There is no CPU fallback for the PCGmix kernels.") from exc
        lib = ctypes.CDLL(path)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = (ctypes.c_char_p if name == "pcgmix_last_error" else
                          ctypes.c_longlong if name in ("pcgmix_overlap_launches", "pcgmix_first_conv_block_workspace")
                          else ctypes.c_int)
        _lib = lib
    return _lib


def _check(rc: int, what: str):
    if rc != 0:
        msg = load().pcgmix_last_error()
        raise RuntimeError(f"{what} failed (status {rc}): {msg.decode() if msg else 'unknown error'}")


def _dev_ptr(t, dtype, name, allow_none=False):
    if t is None:
        if allow_none:
            return None
        raise ValueError(f"{name} is required")
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the PCGmix kernels have no CPU path")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t.data_ptr()


def _frames_ptr(frames):
    """Pointer and row stride (in int32) of a frames table.  Accepts a packed (B, 5) tensor or a
    strided view such as ``cycles[:, 3:]`` of the (n, 8) cycle table written by the segmentation
    kernels — the kernels take the stride, no copy is made."""
    if not isinstance(frames, torch.Tensor) or not frames.is_cuda:
        raise RuntimeError("frames must be a CUDA tensor: the PCGmix kernels have no CPU path")
    if frames.dtype != torch.int32:
        raise TypeError(f"frames must be int32 on the device, got {frames.dtype}")
    if frames.dim() != 2 or frames.shape[1] < 5:
        raise ValueError("frames must be (B, >=5)")
    if frames.shape[0] > 1 and (frames.stride(1) != 1 or frames.stride(0) < 5):
        raise ValueError("frames rows must be unit-stride with a row stride >= 5")
    stride = frames.stride(0) if frames.shape[0] > 1 else max(5, frames.stride(0))
    return frames.data_ptr(), int(stride)


def _check_rows(batch: int, **tables):
    """Per-cycle tables must cover the batch: the kernels index them with cycle ids up to ``batch - 1``
    (the entries themselves are range-checked on the device)."""
    for name, t in tables.items():
        if t is not None and (t.dim() < 1 or t.shape[0] < batch):
            raise ValueError(f"{name} has {t.shape[0] if t.dim() else 0} rows, the batch has {batch} cycles")


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream_handle(device) -> int:
    """Raw handle of the device's current stream (the C-level getter when this PyTorch has it: building a
    ``torch.cuda.Stream`` object per launch costs several microseconds)."""
    if _raw_stream is not None:
        index = device.index if device.index is not None else torch.cuda.current_device()
        return _raw_stream(index)
    return torch.cuda.current_stream(device).cuda_stream


class _on_device:
    """``with _on_device(dev):`` makes ``dev`` the current CUDA device for the launch.  Switching
    the device costs several microseconds, which shows on the small latency-bound batches, so it
    is skipped when ``dev`` already is the current device (the normal one-process-per-GPU case)."""

    __slots__ = ("ctx",)

    def __init__(self, device):
        self.ctx = None
        if device.index is not None and device.index != torch.cuda.current_device():
            self.ctx = torch.cuda.device(device)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


def copy_small(dst, src, nbytes: int):
    """Small copy between device memory and pinned host memory (either direction) done by a kernel
    instead of a copy engine (see ``pcgmix_copy_small``); enqueued on the current stream of the
    CUDA tensor's device."""
    global launch_count
    cuda_t, host_t = (dst, src) if dst.is_cuda else (src, dst)
    if not cuda_t.is_cuda or host_t.is_cuda or not host_t.is_pinned():
        raise RuntimeError("copy_small needs one CUDA tensor and one pinned host tensor")
    dev = cuda_t.device
    with _on_device(dev):
        rc = load().pcgmix_copy_small(dst.data_ptr(), src.data_ptr(), int(nbytes), _stream_handle(dev))
    _check(rc, "pcgmix_copy_small")
    launch_count += 1 if nbytes > 0 else 0


def copy_small_raw(dst_ptr: int, src_ptr: int, nbytes: int, device):
    """``copy_small`` for callers that own both buffers (the staging ring): no tensor checks, raw addresses."""
    global launch_count
    with _on_device(device):
        rc = load().pcgmix_copy_small(dst_ptr, src_ptr, int(nbytes), _stream_handle(device))
    if rc != 0:
        _check(rc, "pcgmix_copy_small")
    launch_count += 1 if nbytes > 0 else 0


def host_group_permutation(group_ids, n_groups: int, seed: int):
    """``mix[idx_g] = random.Random(seed).sample(idx_g, len(idx_g))`` for every group, computed by
    the C++ replay of CPython's algorithm (host code, no GPU involved)."""
    import numpy as np
    g = np.ascontiguousarray(group_ids, dtype=np.int64)
    mix = np.arange(g.shape[0], dtype=np.int64)
    rc = load().pcgmix_host_group_permutation(g.ctypes.data, g.shape[0], int(n_groups), int(seed), mix.ctypes.data)
    if rc != 0:
        raise RuntimeError("pcgmix_host_group_permutation: bad arguments")
    return mix


def host_lambda_knots(seed: int, alpha: float, sigma: float, shape, max_threads: int = 4, want_state: bool = True):
    """``np.random.seed(seed); lam = beta(alpha, alpha); knots = normal(1, sigma, shape)`` replayed in C++
    (bit-equal to NumPy's legacy stream).  Returns ``(lam, knots, state)``; ``state`` is the tuple
    ``np.random.set_state`` accepts (the stream position NumPy would be left at) or None.  Raises
    ``ValueError`` where the replay does not apply (seed outside [0, 2^32), alpha <= 0)."""
    import numpy as np
    n = 1
    for d in shape:
        n *= int(d)
    knots = np.empty(shape, dtype=np.float64)
    lam = ctypes.c_double()
    key = np.empty(624, dtype=np.uint32) if want_state else None
    pos, has_gauss, gauss = _c_i32(), _c_i32(), ctypes.c_double()
    if not (0 <= int(seed) < 2 ** 32):
        raise ValueError("seed outside [0, 2^32)")
    rc = load().pcgmix_host_lambda_knots(int(seed), float(alpha), float(sigma), n, int(max_threads), ctypes.byref(lam),
                                         knots.ctypes.data, key.ctypes.data if want_state else None, ctypes.byref(pos),
                                         ctypes.byref(has_gauss), ctypes.byref(gauss))
    if rc != 0:
        raise ValueError("pcgmix_host_lambda_knots: the replay does not apply to these arguments")
    state = ("MT19937", key, pos.value, has_gauss.value, gauss.value) if want_state else None
    return lam.value, knots, state


def host_prepare_step(labels, frames, length: int, seed: int, knot: int, channels: int, want_order: bool, packed):
    """One foreign call for the integer half of a plain step (see ``pcgmix_host_prepare_step``): ``labels``
    int64 (B,), ``frames`` int64 (B, >=5) CPU arrays, ``packed`` a pinned uint8 tensor.  Returns
    ``(status, info, mix)`` — ``info`` the int64[16] layout / diagnostics block, ``mix`` the int64 pairing."""
    import numpy as np
    B = labels.shape[0]
    info = np.zeros(16, dtype=np.int64)
    mix = np.empty(B, dtype=np.int64)
    rc = load().pcgmix_host_prepare_step(labels.ctypes.data, B, frames.ctypes.data, frames.strides[0] // 8, int(length),
                                         int(seed), int(knot), int(channels), int(bool(want_order)), packed.data_ptr(),
                                         packed.numel(), info.ctypes.data, mix.ctypes.data)
    return rc, info, mix


def mix1d_packed(x, out, tables, info, lam32, one_minus_lam32, coefmat=None, knot_pos=None, knot=-1, err_flag=None):
    """``pcgmix_mix1d`` / ``pcgmix_mix1d_magwarp`` with the per-step tables taken from ONE device buffer at the
    byte offsets ``info[0..3]`` written by ``host_prepare_step`` (no views, no re-validation: the layout was
    produced by the library itself)."""
    global launch_count
    B, C, L = x.shape
    base = tables.data_ptr()
    order = base + int(info[2]) if info[2] >= 0 else None
    dev = x.device
    lib = load()
    with _on_device(dev):
        if knot < 0:
            rc = lib.pcgmix_mix1d(x.data_ptr(), out.data_ptr(), base + int(info[0]), 5, base + int(info[1]), order,
                                  float(lam32), float(one_minus_lam32), B, C, L,
                                  _dev_ptr(err_flag, torch.int32, "err_flag", True), _stream_handle(dev))
        else:
            rc = lib.pcgmix_mix1d_magwarp(x.data_ptr(), out.data_ptr(), base + int(info[0]), 5, base + int(info[1]), order,
                                          float(lam32), float(one_minus_lam32), base + int(info[3]), coefmat.data_ptr(),
                                          knot_pos.data_ptr(), int(knot), B, C, L,
                                          _dev_ptr(err_flag, torch.int32, "err_flag", True), _stream_handle(dev))
    _check(rc, "pcgmix_mix1d (packed tables)")
    launch_count += 1 if B > 0 else 0


def host_processing_order(mix):
    import numpy as np
    m = np.ascontiguousarray(mix, dtype=np.int64)
    order = np.empty(m.shape[0], dtype=np.int32)
    if load().pcgmix_host_processing_order(m.ctypes.data, m.shape[0], order.ctypes.data) != 0:
        raise RuntimeError("pcgmix_host_processing_order: bad arguments")
    return order


def version() -> int:
    return load().pcgmix_version()


def device_info():
    sm, major, minor = _c_i32(), _c_i32(), _c_i32()
    _check(load().pcgmix_device_info(ctypes.byref(sm), ctypes.byref(major), ctypes.byref(minor)), "pcgmix_device_info")
    return sm.value, major.value, minor.value


def set_tuning(use_pipeline: bool = True, stages: int = 0, max_slice: int = 0, ctas_per_sm: int = 0,
               pbuf_pct: int = 0, consumer_threads: int = 0, debug: int = 0):
    """Kernel selection knobs (see ``pcgmix_set_tuning`` in the header); results never depend on them."""
    _check(load().pcgmix_set_tuning(int(bool(use_pipeline)), int(stages), int(max_slice), int(ctas_per_sm),
                                    int(pbuf_pct), int(consumer_threads), int(debug)),
           "pcgmix_set_tuning")


def set_launch_overlap(enable: bool):
    """Let consecutive, buffer-disjoint PCGmix launches on one stream overlap (programmatic dependent
    launch); see ``pcgmix_set_launch_overlap`` in the header for what the caller asserts."""
    _check(load().pcgmix_set_launch_overlap(int(bool(enable))), "pcgmix_set_launch_overlap")


def set_spline_precision(precision: str):
    """``"float32"`` (default: PCGmix+ warp factor evaluated in fp32, <= 1e-5 relative to the reference, as
    fast as plain PCGmix) or ``"float64"`` (bit-faithful evaluation, ~15 % slower); pipelined kernel only."""
    if precision not in ("float32", "float64"):
        raise ValueError("precision must be 'float32' or 'float64'")
    _check(load().pcgmix_set_spline_precision(1 if precision == "float32" else 0), "pcgmix_set_spline_precision")


def spline_precision() -> str:
    return "float32" if load().pcgmix_get_spline_precision() else "float64"


def overlap_launches() -> int:
    return int(load().pcgmix_overlap_launches())


def _same_device(*tensors):
    dev = None
    for t in tensors:
        if t is None:
            continue
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError("all tensors must live on the same CUDA device")
    return dev


def mix1d(x, out, frames, mix, lam32, one_minus_lam32, order=None, err_flag=None):
    """``out[B,C,L] = PCGmix(x)``; see ``pcgmix_mix1d`` in the header."""
    global launch_count
    if x.dim() != 3 or out.shape != x.shape:
        raise ValueError("x and out must be (B, C, L) of equal shape")
    B, C, L = x.shape
    _check_rows(B, frames=frames, mix=mix, order=order)
    dev = _same_device(x, out, frames, mix, order, err_flag)
    fptr, fstride = _frames_ptr(frames)
    with _on_device(dev):
        rc = load().pcgmix_mix1d(
            _dev_ptr(x, torch.float32, "x"), _dev_ptr(out, torch.float32, "out"),
            fptr, fstride, _dev_ptr(mix, torch.int32, "mix"),
            _dev_ptr(order, torch.int32, "order", True), float(lam32), float(one_minus_lam32), B, C, L,
            _dev_ptr(err_flag, torch.int32, "err_flag", True), _stream_handle(dev))
    _check(rc, "pcgmix_mix1d")
    launch_count += 1 if B > 0 else 0


def mix1d_magwarp(x, out, frames, mix, lam32, one_minus_lam32, knots, coefmat, knot_pos, knot,
                  order=None, err_flag=None):
    """Fused PCGmix+; see ``pcgmix_mix1d_magwarp`` in the header."""
    global launch_count
    if x.dim() != 3 or out.shape != x.shape:
        raise ValueError("x and out must be (B, C, L) of equal shape")
    B, C, L = x.shape
    if tuple(knots.shape) != (B, knot + 2, C):
        raise ValueError(f"knots must be (B, knot+2, C) = {(B, knot + 2, C)}, got {tuple(knots.shape)}")
    if tuple(coefmat.shape) != ((knot + 1) * 4, knot + 2) or tuple(knot_pos.shape) != (knot + 3,):
        raise ValueError("coefmat / knot_pos do not match knot")
    _check_rows(B, frames=frames, mix=mix, order=order)
    dev = _same_device(x, out, frames, mix, order, err_flag, knots, coefmat, knot_pos)
    fptr, fstride = _frames_ptr(frames)
    with _on_device(dev):
        rc = load().pcgmix_mix1d_magwarp(
            _dev_ptr(x, torch.float32, "x"), _dev_ptr(out, torch.float32, "out"),
            fptr, fstride, _dev_ptr(mix, torch.int32, "mix"),
            _dev_ptr(order, torch.int32, "order", True), float(lam32), float(one_minus_lam32),
            _dev_ptr(knots, torch.float64, "knots"), _dev_ptr(coefmat, torch.float64, "coefmat"),
            _dev_ptr(knot_pos, torch.float64, "knot_pos"), int(knot), B, C, L,
            _dev_ptr(err_flag, torch.int32, "err_flag", True), _stream_handle(dev))
    _check(rc, "pcgmix_mix1d_magwarp")
    launch_count += 1 if B > 0 else 0


def mix1d_windows(x, out, windows, mix, lam32, one_minus_lam32, knots=None, coefmat=None, knot_pos=None, knot=0,
                  order=None, err_flag=None):
    """PCGmix / PCGmix+ with explicit windows (B, 4, 3) int32; see ``pcgmix_mix1d_windows``."""
    global launch_count
    if x.dim() != 3 or out.shape != x.shape:
        raise ValueError("x and out must be (B, C, L) of equal shape")
    B, C, L = x.shape
    if tuple(windows.shape) != (B, 4, 3):
        raise ValueError(f"windows must be (B, 4, 3), got {tuple(windows.shape)}")
    if knots is not None and tuple(knots.shape) != (B, knot + 2, C):
        raise ValueError("knots must be (B, knot+2, C)")
    _check_rows(B, mix=mix, order=order)
    dev = _same_device(x, out, windows, mix, order, err_flag, knots, coefmat, knot_pos)
    with _on_device(dev):
        rc = load().pcgmix_mix1d_windows(
            _dev_ptr(x, torch.float32, "x"), _dev_ptr(out, torch.float32, "out"),
            _dev_ptr(windows, torch.int32, "windows"), _dev_ptr(mix, torch.int32, "mix"),
            _dev_ptr(order, torch.int32, "order", True), float(lam32), float(one_minus_lam32),
            _dev_ptr(knots, torch.float64, "knots", True), _dev_ptr(coefmat, torch.float64, "coefmat", True),
            _dev_ptr(knot_pos, torch.float64, "knot_pos", True), int(knot), B, C, L,
            _dev_ptr(err_flag, torch.int32, "err_flag", True), _stream_handle(dev))
    _check(rc, "pcgmix_mix1d_windows")
    launch_count += 1 if B > 0 else 0


class PreparedMix1D:
    """A PCGmix / PCGmix+ launch with every argument resolved once.

    ``mix1d`` / ``mix1d_magwarp`` validate and unpack their tensors on every call (~20 us of Python),
    which is comparable to the kernel itself when resident batches are swept back to back.  This
    object does that work at construction and ``launch()`` is a single foreign call on the current
    stream.  It keeps references to all tensors, so their memory stays valid."""

    def __init__(self, x, out, frames, mix, lam32, one_minus_lam32, knots=None, coefmat=None, knot_pos=None,
                 knot=0, order=None, err_flag=None):
        if x.dim() != 3 or out.shape != x.shape:
            raise ValueError("x and out must be (B, C, L) of equal shape")
        B, C, L = x.shape
        _check_rows(B, frames=frames, mix=mix, order=order)
        self._keep = (x, out, frames, mix, knots, coefmat, knot_pos, order, err_flag)
        self._device = _same_device(x, out, frames, mix, order, err_flag, knots, coefmat, knot_pos)
        fptr, fstride = _frames_ptr(frames)
        lib = load()
        common = (_dev_ptr(x, torch.float32, "x"), _dev_ptr(out, torch.float32, "out"), fptr, fstride,
                  _dev_ptr(mix, torch.int32, "mix"), _dev_ptr(order, torch.int32, "order", True),
                  float(lam32), float(one_minus_lam32))
        tail = (B, C, L, _dev_ptr(err_flag, torch.int32, "err_flag", True))
        if knots is None:
            self._fn, self._args, self._name = lib.pcgmix_mix1d, common + tail, "pcgmix_mix1d"
        else:
            if tuple(knots.shape) != (B, knot + 2, C):
                raise ValueError("knots must be (B, knot+2, C)")
            spline = (_dev_ptr(knots, torch.float64, "knots"), _dev_ptr(coefmat, torch.float64, "coefmat"),
                      _dev_ptr(knot_pos, torch.float64, "knot_pos"), int(knot))
            self._fn, self._args, self._name = lib.pcgmix_mix1d_magwarp, common + spline + tail, "pcgmix_mix1d_magwarp"
        self._count = 1 if B > 0 else 0

    def launch(self, stream_handle=None):
        global launch_count
        if stream_handle is None:
            stream_handle = _stream_handle(self._device)
        with _on_device(self._device):                     # no-op when the tensors' device already is current
            rc = self._fn(*self._args, stream_handle)
        if rc != 0:
            _check(rc, self._name)
        launch_count += self._count


def mix2d(x, out, frames, mix, lam32, one_minus_lam32, tbox=None, h1=0, h2=0, order=None, err_flag=None):
    """PCGmix on (B, Ch, F, T) with the optional zero box; see ``pcgmix_mix2d`` in the header."""
    global launch_count
    if x.dim() != 4 or out.shape != x.shape:
        raise ValueError("x and out must be (B, Ch, F, T) of equal shape")
    B, Ch, F, T = x.shape
    _check_rows(B, frames=frames, mix=mix, order=order, tbox=tbox)
    dev = _same_device(x, out, frames, mix, order, err_flag, tbox)
    fptr, fstride = _frames_ptr(frames)
    with _on_device(dev):
        rc = load().pcgmix_mix2d(
            _dev_ptr(x, torch.float32, "x"), _dev_ptr(out, torch.float32, "out"),
            fptr, fstride, _dev_ptr(mix, torch.int32, "mix"),
            _dev_ptr(order, torch.int32, "order", True), float(lam32), float(one_minus_lam32), B, Ch, F, T,
            _dev_ptr(tbox, torch.int32, "tbox", True), int(h1), int(h2),
            _dev_ptr(err_flag, torch.int32, "err_flag", True), _stream_handle(dev))
    _check(rc, "pcgmix_mix2d")
    launch_count += 1 if B > 0 else 0


def segment_dense(states, downsample, cycles, cycle_count, err_flag=None):
    global launch_count
    R, T = states.shape
    dev = _same_device(states, cycles, cycle_count, err_flag)
    with _on_device(dev):
        rc = load().pcgmix_segment_dense(
            _dev_ptr(states, torch.int8, "states"), R, T, int(downsample),
            _dev_ptr(cycles, torch.int32, "cycles"), cycles.shape[0], _dev_ptr(cycle_count, torch.int32, "cycle_count"),
            _dev_ptr(err_flag, torch.int32, "err_flag", True), _stream_handle(dev))
    _check(rc, "pcgmix_segment_dense")
    launch_count += 3 if R > 0 else 0


def segment_table(positions, codes, rec_offsets, downsample, spec_cols, rec_len, cycles, cycle_count, err_flag=None):
    global launch_count
    R = rec_offsets.shape[0] - 1
    dev = _same_device(positions, codes, rec_offsets, rec_len, cycles, cycle_count, err_flag)
    with _on_device(dev):
        rc = load().pcgmix_segment_table(
            _dev_ptr(positions, torch.int32, "positions"), _dev_ptr(codes, torch.int8, "codes"),
            _dev_ptr(rec_offsets, torch.int32, "rec_offsets"), R, int(downsample), int(spec_cols),
            _dev_ptr(rec_len, torch.int32, "rec_len", True), _dev_ptr(cycles, torch.int32, "cycles"),
            cycles.shape[0], _dev_ptr(cycle_count, torch.int32, "cycle_count"),
            _dev_ptr(err_flag, torch.int32, "err_flag", True), _stream_handle(dev))
    _check(rc, "pcgmix_segment_table")
    launch_count += 3 if R > 0 else 0


def cut_cycles(signal, cycles, n_cycles, out, n_cycles_dev=None):
    global launch_count
    R, C, T = signal.shape
    L = out.shape[-1]
    dev = _same_device(signal, cycles, out, n_cycles_dev)
    with _on_device(dev):
        rc = load().pcgmix_cut_cycles(
            _dev_ptr(signal, torch.float32, "signal"), R, C, T, _dev_ptr(cycles, torch.int32, "cycles"),
            int(n_cycles), _dev_ptr(n_cycles_dev, torch.int32, "n_cycles_dev", True),
            _dev_ptr(out, torch.float32, "out"), L, _stream_handle(dev))
    _check(rc, "pcgmix_cut_cycles")
    launch_count += 1 if n_cycles > 0 else 0


def mix1d_resident(signal, cycles, sel, mix, lam32, one_minus_lam32, out, knots=None, coefmat=None, knot_pos=None,
                   knot=0, order=None, err_flag=None, scratch=None):
    """Cut + zero-pad + PCGmix(+) in one launch: ``signal`` (n_rec, C, T) fp32, ``cycles`` (n_table, 8) int32
    cycle table, ``sel`` (B,) int32 table rows of the batch or None, ``out`` (B, C, L); ``scratch`` (B, 8) int32
    lets the library use the pipelined kernel (slot records are written there first)."""
    global launch_count
    n_rec, C, T = signal.shape
    B, C_out, L = out.shape
    if C_out != C:
        raise ValueError(f"out has {C_out} channels, the recordings have {C}")
    if cycles.dim() != 2 or cycles.shape[1] != 8 or not cycles.is_contiguous():
        raise ValueError("cycles must be a contiguous (n, 8) int32 cycle table")
    _check_rows(B, sel=sel, mix=mix, order=order, knots=knots)
    if knots is not None and tuple(knots.shape) != (B, knot + 2, C):
        raise ValueError(f"knots must be (B, knot+2, C) = {(B, knot + 2, C)}, got {tuple(knots.shape)}")
    dev = _same_device(signal, cycles, sel, mix, out, knots, coefmat, knot_pos, order, err_flag, scratch)
    if scratch is not None and (scratch.numel() < 8 * B or not scratch.is_contiguous()):
        raise ValueError("scratch must be a contiguous int32 tensor of at least B*8 elements")
    with _on_device(dev):
        rc = load().pcgmix_mix1d_resident(
            _dev_ptr(signal, torch.float32, "signal"), n_rec, C, T, _dev_ptr(cycles, torch.int32, "cycles"),
            cycles.shape[0], _dev_ptr(sel, torch.int32, "sel", True), _dev_ptr(mix, torch.int32, "mix"),
            _dev_ptr(order, torch.int32, "order", True), float(lam32), float(one_minus_lam32),
            _dev_ptr(knots, torch.float64, "knots", True), _dev_ptr(coefmat, torch.float64, "coefmat", True),
            _dev_ptr(knot_pos, torch.float64, "knot_pos", True), int(knot),
            _dev_ptr(out, torch.float32, "out"), B, L, _dev_ptr(scratch, torch.int32, "scratch", True),
            _dev_ptr(err_flag, torch.int32, "err_flag", True), _stream_handle(dev))
    _check(rc, "pcgmix_mix1d_resident")
    launch_count += (2 if scratch is not None else 1) if B > 0 else 0


def duration_features(frames, n, fs, features, err_flag=None):
    global launch_count
    dev = _same_device(frames, features, err_flag)
    fptr, fstride = _frames_ptr(frames)
    with _on_device(dev):
        rc = load().pcgmix_duration_features(
            fptr, fstride, int(n), int(fs),
            _dev_ptr(features, torch.float64, "features"), _dev_ptr(err_flag, torch.int32, "err_flag", True),
            _stream_handle(dev))
    _check(rc, "pcgmix_duration_features")
    launch_count += 1 if n > 0 else 0


def cycle_moment_features(x, frames, channel: int, features, err_flag=None):
    """Skewness / kurtosis block of ``feature_vector_seg`` for ``x[:, channel]``; see ``pcgmix_cycle_moment_features``."""
    global launch_count
    if x.dim() != 3:
        raise ValueError("x must be (B, C, L)")
    B, C, L = x.shape
    if tuple(features.shape) != (B, CYCLE_MOMENT_FEATURES):
        raise ValueError(f"features must be (B, {CYCLE_MOMENT_FEATURES})")
    _check_rows(B, frames=frames)
    dev = _same_device(x, frames, features, err_flag)
    fptr, fstride = _frames_ptr(frames)
    with _on_device(dev):
        rc = load().pcgmix_cycle_moment_features(_dev_ptr(x, torch.float32, "x"), fptr, fstride, B, C, L, int(channel),
                                                 _dev_ptr(features, torch.float32, "features"),
                                                 _dev_ptr(err_flag, torch.int32, "err_flag", True), _stream_handle(dev))
    _check(rc, "pcgmix_cycle_moment_features")
    launch_count += 1 if B > 0 else 0


def first_conv_block_workspace(C: int, F: int) -> int:
    """Bytes of device workspace ``first_conv_block`` needs for ``C`` input channels and ``F`` filters."""
    n = int(load().pcgmix_first_conv_block_workspace(int(C), int(F)))
    if n < 0:
        raise ValueError(f"first_conv_block: unsupported size (C={C} must be 1..4, F={F} at most {MAX_FIRST_BLOCK_FILTERS})")
    return n


def first_conv_block(x, weight, bias, gamma, beta, running_mean, running_var, out, workspace, batch_stats: bool,
                     eps: float, momentum: float, save_mean=None, save_invstd=None):
    """Conv1d(C, F, 3, padding=1) + BatchNorm1d + ReLU forward; see ``pcgmix_first_conv_block`` in the header."""
    global launch_count
    if x.dim() != 3 or weight.dim() != 3:
        raise ValueError("x must be (B, C, L) and weight (F, C, 3)")
    B, C, L = x.shape
    F = weight.shape[0]
    if tuple(weight.shape) != (F, C, 3):
        raise ValueError(f"weight must be (F, {C}, 3), got {tuple(weight.shape)}")
    if tuple(out.shape) != (B, F, L):
        raise ValueError(f"out must be ({B}, {F}, {L}), got {tuple(out.shape)}")
    for name, t in (("bias", bias), ("gamma", gamma), ("beta", beta), ("running_mean", running_mean),
                    ("running_var", running_var), ("save_mean", save_mean), ("save_invstd", save_invstd)):
        if t is not None and tuple(t.shape) != (F,):
            raise ValueError(f"{name} must have {F} entries, got {tuple(t.shape)}")
    if workspace.dtype != torch.uint8 or workspace.numel() < first_conv_block_workspace(C, F):
        raise ValueError("workspace must be a uint8 tensor of first_conv_block_workspace(C, F) bytes")
    dev = _same_device(x, weight, bias, gamma, beta, running_mean, running_var, out, workspace, save_mean, save_invstd)
    f32 = torch.float32
    with _on_device(dev):
        rc = load().pcgmix_first_conv_block(
            _dev_ptr(x, f32, "x"), _dev_ptr(weight, f32, "weight"), _dev_ptr(bias, f32, "bias", True),
            _dev_ptr(gamma, f32, "gamma", True), _dev_ptr(beta, f32, "beta", True),
            _dev_ptr(running_mean, f32, "running_mean", True), _dev_ptr(running_var, f32, "running_var", True),
            _dev_ptr(out, f32, "out"), _dev_ptr(workspace, torch.uint8, "workspace"), B, C, L, F, 1 if batch_stats else 0,
            float(eps), float(momentum), _dev_ptr(save_mean, f32, "save_mean", True),
            _dev_ptr(save_invstd, f32, "save_invstd", True), _stream_handle(dev))
    _check(rc, "pcgmix_first_conv_block")
    launch_count += (3 if batch_stats else 2) if B > 0 else 0


def cycle_psd_features(x, frames, channel: int, fs: int, features, err_flag=None):
    """Welch-PSD block of ``feature_vector_seg`` for ``x[:, channel]``; see ``pcgmix_cycle_psd_features`` in the header."""
    global launch_count
    if x.dim() != 3:
        raise ValueError("x must be (B, C, L)")
    B, C, L = x.shape
    if tuple(features.shape) != (B, CYCLE_PSD_FEATURES):
        raise ValueError(f"features must be (B, {CYCLE_PSD_FEATURES})")
    _check_rows(B, frames=frames)
    dev = _same_device(x, frames, features, err_flag)
    fptr, fstride = _frames_ptr(frames)
    with _on_device(dev):
        rc = load().pcgmix_cycle_psd_features(_dev_ptr(x, torch.float32, "x"), fptr, fstride, B, C, L, int(channel), int(fs),
                                              _dev_ptr(features, torch.float32, "features"),
                                              _dev_ptr(err_flag, torch.int32, "err_flag", True), _stream_handle(dev))
    _check(rc, "pcgmix_cycle_psd_features")
    launch_count += 1 if B > 0 else 0


def cycle_features(x, frames, channel: int, what: int, features, err_flag=None):
    """Amplitude (what & 1) and Hilbert-envelope (what & 2) features of ``x[:, channel]``; see
    ``pcgmix_cycle_features`` in the header."""
    global launch_count
    if x.dim() != 3:
        raise ValueError("x must be (B, C, L)")
    B, C, L = x.shape
    if tuple(features.shape) != (B, CYCLE_FEATURES):
        raise ValueError(f"features must be (B, {CYCLE_FEATURES})")
    _check_rows(B, frames=frames)
    dev = _same_device(x, frames, features, err_flag)
    fptr, fstride = _frames_ptr(frames)
    with _on_device(dev):
        rc = load().pcgmix_cycle_features(_dev_ptr(x, torch.float32, "x"), fptr, fstride, B, C, L, int(channel), int(what),
                                          _dev_ptr(features, torch.float32, "features"),
                                          _dev_ptr(err_flag, torch.int32, "err_flag", True), _stream_handle(dev))
    _check(rc, "pcgmix_cycle_features")
    launch_count += 1 if B > 0 else 0
