"""Host-side replay of the reference's seeded draws and of its method-string mini-language.

Everything random on the PCGmix path is drawn on the host from the integer training step
(``step_counter.count``) and handed to the device kernels, so this implementation and the
reference consume identical randomness.  The calls below are made on the *same* generators, in
the *same* order, as the reference makes them:

  gate ............ ``random.Random(step).uniform(0, 1)``            augmentations.py:936-939
  pairing ......... per class ``random.Random(step).sample(idx, n)`` augmentations.py:500-514
  lambda .......... ``np.random.seed(step); np.random.beta(a, a)``   augmentations.py:659-666
  knots ........... ``np.random.normal(1, sigma, (B, knot+2, C))``   augmentations.py:677
                    from the global legacy stream right after the beta draw

The global NumPy stream is re-seeded on purpose: the reference does it every step, and code
that runs after ``augment`` in the same process sees that state.
"""
from __future__ import annotations

import dataclasses
import random
from typing import Optional, Sequence

import numpy as np

# Branches of the reference dispatcher that are tested BEFORE the PCGmix branches
# (augmentations.py:734, 777, 807).  A method string that also contains one of these would never
# reach PCGmix in the reference, so it is refused here instead of being silently reinterpreted.
_EARLIER_1D = ("durmixrespscale", "respiratoryscale", "timemask")

# Every branch keyword of the reference dispatchers (augmentations.py:700-729,
# augmentations2d.py:269-281): a method containing none of them is returned untouched.
METHODS_1D = ("durratiocutmix", "lengthcutmix", "datasetcutmix", "wav-durratiocutmix", "wavcutmix",
              "lc-nointrusion", "labelcutmix", "swapsysdia", "s1s2mask", "cont-cutmix", "saliency-cutmix",
              "latentmixup", "manifold-cutmix(ch)", "manifold-cutmix", "manifold-cutout(ch)",
              "manifold-cutout", "cutmix(ch)", "cutmix", "cutout(ch)", "cutout", "gaussiannoise",
              "magnitudewarp", "timewarp", "mixup", "timemask", "durratiomixup", "durmixmagwarp",
              "respiratoryscale", "durmixrespscale")
METHODS_2D = ("durratiocutmix", "cutmix", "mixup", "latentmixup", "freqmask", "timemask", "cutout",
              "durratiomixup", "durmixfreqmask", "durmixtimemask", "durmixcutout")

# Pairing / displacement modifiers this implementation does not provide (they need files or
# trained models that are not part of the reference repository, or are ablations).
_UNSUPPORTED_MODIFIERS = ("(sameCVD)", "(closestbins=", "(closestknn=", "(salopt")


@dataclasses.dataclass
class Plan1D:
    branch: str                 # 'durratiomixup' | 'durmixmagwarp'
    probability: float
    alpha: float
    sigma: float = 0.2
    knot: int = 4
    mix_all: bool = False
    rand_displacement: bool = False


@dataclasses.dataclass
class Plan2D:
    branch: str                 # 'durratiomixup' | 'durmixtimemask' | 'durmixfreqmask' | 'durmixcutout'
    probability: float
    time_region_max: float = 0.2
    freq_region_max: float = 0.2


def _probability(method: str) -> float:
    parts = method.split("+")
    return float(parts[-1]) if len(parts) > 1 else 1.0


def _alpha(method: str, branch: str) -> float:
    if len(method.split("(alpha=")) > 1:
        return float(method.split("(alpha=")[1].split(")" + branch)[0])
    return 1.0


def _clamp01(v: float) -> float:
    return min(max(v, 0), 1)


def parse_method_1d(method: str) -> Optional[Plan1D]:
    """Return the plan for a PCGmix(+) method string, ``None`` if the reference would return the
    batch untouched, and raise for strings that the reference routes to another augmentation."""
    if not any(k in method for k in METHODS_1D):
        return None
    for earlier in _EARLIER_1D:
        if earlier in method:
            raise NotImplementedError(
                f"method {method!r} is routed to the {earlier!r} branch by the reference dispatcher; "
                "only durratiomixup / durmixmagwarp are provided here")
    if "durmixmagwarp" in method:
        branch = "durmixmagwarp"
    elif "durratiomixup" in method:
        branch = "durratiomixup"
    else:
        raise NotImplementedError(f"method {method!r} is not on the PCGmix hot path")
    for mod in _UNSUPPORTED_MODIFIERS:
        if mod in method:
            raise NotImplementedError(f"modifier {mod} of {method!r} is not provided by this implementation")
    plan = Plan1D(branch=branch, probability=_probability(method), alpha=_alpha(method, branch),
                  mix_all="(mixAll)" in method, rand_displacement="(rand)" in method)
    if branch == "durmixmagwarp" and len(method.split("durmixmagwarp(")) > 1:
        plan.sigma = float(method.split("durmixmagwarp(")[1].split(",")[0])
        plan.knot = int(method.split(",")[1].split(")")[0])
    return plan


def parse_method_2d(method: str) -> Optional[Plan2D]:
    """Same for the spectrogram dispatcher (augmentations2d.py:283-427); branch order as there."""
    if not any(k in method for k in METHODS_2D):
        return None
    if "durmixcutout" in method:
        plan = Plan2D("durmixcutout", _probability(method))
        if len(method.split("cutout(")) > 1:
            plan.time_region_max = _clamp01(float(method.split("cutout(")[1].split(",")[0]))
            plan.freq_region_max = _clamp01(float(method.split(",")[1].split(")")[0]))
        return plan
    if "durmixtimemask" in method:
        plan = Plan2D("durmixtimemask", _probability(method))
        if len(method.split("timemask(")) > 1:
            plan.time_region_max = _clamp01(float(method.split("timemask(")[1].split(")")[0]))
        return plan
    if "durmixfreqmask" in method:
        plan = Plan2D("durmixfreqmask", _probability(method))
        if len(method.split("freqmask(")) > 1:
            plan.freq_region_max = _clamp01(float(method.split("freqmask(")[1].split(")")[0]))
        return plan
    if "durratiomixup" in method:
        if "(salopt" in method:
            raise NotImplementedError("the (salopt...) variants are not provided by this implementation")
        return Plan2D("durratiomixup", _probability(method))
    raise NotImplementedError(f"method {method!r} is not on the PCGmix hot path")


def gate(step: int) -> float:
    return random.Random(step).uniform(0, 1)


def _grouped_permutation(keys: Sequence, step: int) -> np.ndarray:
    """Seeded permutation inside every group of equal keys.  Uses the C++ replay of CPython's
    ``random.Random(step).sample`` from the native library (20x faster than the Python loop, held
    bit-equal to it by tests); falls back to CPython itself if the library cannot be loaded or the
    seed is negative (this is host logic — the GPU kernels have no such fallback)."""
    if isinstance(step, (int, np.integer)) and step >= 0:
        try:
            from . import native
        except (ImportError, OSError):
            return _grouped_permutation_python(keys, step)
        try:
            _, inverse = np.unique(keys if isinstance(keys, np.ndarray) else np.asarray(keys), return_inverse=True)
            return native.host_group_permutation(inverse.reshape(-1), int(inverse.max()) + 1 if inverse.size else 0, int(step))
        except (native.NativeLibraryError, OSError):
            pass
    return _grouped_permutation_python(keys, step)


def _grouped_permutation_python(keys: Sequence, step: int) -> np.ndarray:
    groups = {}
    for i, k in enumerate(keys):
        groups.setdefault(k, []).append(i)
    mix = np.arange(0, len(keys), 1)
    for members in groups.values():
        mix[members] = random.Random(step).sample(list(mix[members]), len(members))
    return mix


def same_label_pairing(labels: np.ndarray, step: int) -> np.ndarray:
    """Default pairing: a seeded permutation inside every class, classes visited in order of first
    appearance, each from a fresh ``Random(step)`` (augmentations.py:500-514)."""
    labels = np.asarray(labels).reshape(-1)
    if labels.dtype.kind in "iub":
        return _grouped_permutation(labels, step)              # class ids: no detour through Python objects
    return _grouped_permutation(labels.tolist(), step)


def pairing(method: str, labels: np.ndarray, wav, step: int) -> np.ndarray:
    """Pairing with the modifiers applied in the reference's order (augmentations.py:943-952)."""
    labels = np.asarray(labels).reshape(-1)
    mix = same_label_pairing(labels, step)
    if "(samePCG)" in method:
        mix = _grouped_permutation(list(wav), step)
    if "(sameDataset)" in method:
        mix = _grouped_permutation([f"{w[0]}_{t}" for w, t in zip(wav, labels.tolist())], step)
    if "(mixAll)" in method:
        mix = np.array(random.Random(step).sample(list(np.arange(0, len(labels), 1)), len(labels)))
    return mix


def draw_lambda(alpha: float, step: int) -> float:
    if alpha > 0.0:
        np.random.seed(step)
        return float(np.random.beta(alpha, alpha))
    return 1.0


# The reference leaves NumPy's GLOBAL legacy stream re-seeded and advanced past the lambda (and knot) draws
# after every step; code that runs later in the same process sees that state.  The C++ replay does not touch
# NumPy, so the state it ends in is written back (np.random.set_state, ~30 us).  Callers that own every use
# of np.random in their process may switch the mirroring off.
mirror_numpy_global_state = True


class _Prefetch:
    """Lambda + knots of one future step, computed on a worker thread (the C++ replay releases the GIL)."""
    __slots__ = ("key", "future")


_prefetch_pool = None
_prefetched = {}
_PREFETCH_DEPTH = 8


def _replay_key(alpha, sigma, step, shape):
    return (float(alpha), float(sigma), int(step), tuple(int(d) for d in shape))


def _replay(alpha, sigma, step, shape):
    from . import native
    return native.host_lambda_knots(step, alpha, sigma, shape, max_threads=1, want_state=mirror_numpy_global_state)


def prefetch_lambda_and_knots(alpha: float, step: int, batch: int, knot: int, channels: int, sigma: float) -> None:
    """Start computing ``lambda_and_knots(...)`` for a future ``step`` in the background.  The seed is the
    step count (train_model.py:105-109, :577), so step k+1's label-independent draws are known while step k
    runs; ``lambda_and_knots`` picks the result up (or computes it itself if nobody asked)."""
    global _prefetch_pool
    if not (alpha > 0.0 and isinstance(step, (int, np.integer)) and 0 <= step < 2 ** 32):
        return
    key = _replay_key(alpha, sigma, step, (batch, knot + 2, channels))
    if key in _prefetched:
        return
    if _prefetch_pool is None:
        from concurrent.futures import ThreadPoolExecutor
        _prefetch_pool = ThreadPoolExecutor(max_workers=4, thread_name_prefix="pcgmix-draws")
    while len(_prefetched) >= _PREFETCH_DEPTH:                 # forget the oldest request nobody collected
        _prefetched.pop(next(iter(_prefetched)))
    _prefetched[key] = _prefetch_pool.submit(_replay, alpha, sigma, int(step), key[3])


def lambda_and_knots(alpha: float, step: int, batch: int, knot: int, channels: int, sigma: float):
    """``(lam, knots)`` exactly as the reference draws them (augmentations.py:659-666 then :677):
    ``np.random.seed(step); lam = np.random.beta(alpha, alpha)`` followed by
    ``knots = np.random.normal(1.0, sigma, (B, knot+2, C))`` from the same global stream.

    Computed by the C++ replay of NumPy's legacy stream (bit-equal, ~3x faster, no GIL, prefetchable);
    NumPy itself is used when the replay does not apply: ``alpha <= 0`` (the reference then neither
    re-seeds nor draws lambda, the knots continue the global stream wherever it stands), seeds NumPy
    refuses, or a missing library (host logic only: the GPU kernels have no such fallback)."""
    shape = (int(batch), int(knot) + 2, int(channels))
    if alpha > 0.0 and isinstance(step, (int, np.integer)) and 0 <= step < 2 ** 32:
        future = _prefetched.pop(_replay_key(alpha, sigma, step, shape), None)
        try:
            lam, knots, state = future.result() if future is not None else _replay(alpha, sigma, int(step), shape)
        except (ImportError, OSError, RuntimeError):
            state = None
        else:
            if mirror_numpy_global_state:
                if state is None:                           # prefetched while mirroring was off
                    np.random.seed(step)
                    np.random.beta(alpha, alpha)
                    np.random.normal(loc=1.0, scale=sigma, size=shape)
                else:
                    np.random.set_state(state)
            return lam, knots
    lam = draw_lambda(alpha, step)
    return lam, draw_knots(batch, knot, channels, sigma)


def lambda_pair_fp32(lam: float):
    """``(lam32, 1 - lam32)`` rounded the way the reference's float32 tensor expression rounds: lambda is stored in a
    float32 array (``np.array(np.ones(B) * lam).astype('float32')``, augmentations.py:962-963 — one rounding of the
    float64 value to nearest) and ``1 - lams`` is evaluated in float32."""
    lam32 = np.float32(lam)
    return lam32, np.float32(1) - lam32


def draw_knots(batch: int, knot: int, channels: int, sigma: float) -> np.ndarray:
    return np.random.normal(loc=1.0, scale=sigma, size=(batch, knot + 2, channels))


def mask_geometry(step: int, region_max: float):
    """``(gap, start_fraction)`` of the seeded zero box (augmentations2d.py:317-319, 354-356)."""
    gap = random.Random(step + 131071).uniform(0, region_max)
    frac1 = random.Random(step + 13119).uniform(0, 1 - gap)
    return gap, frac1


def rand_windows(frames: np.ndarray, mix: np.ndarray, step: int, limit: Optional[int] = None) -> np.ndarray:
    """Blended windows of the reference's ``(rand)`` variant (augmentations.py:305-337), as
    ``(B, 4, 3)`` int32 ``{start in the cycle, blended length, shift to the partner's sample}``.

    Per state the shorter of the two durations is blended, placed at a seeded offset inside the
    longer one: ``disp = random.Random(step).randint(0, |len2 - len1|)`` from a FRESH generator, so
    it depends only on the gap; if the partner's state is longer the offset moves the read window
    in the partner, otherwise it moves the write window in the cycle.  With ``limit`` (the row
    length) the windows are clamped like the reference's slices; unequal clamped widths raise like the
    reference's tensor expression does."""
    f1 = np.asarray(frames, dtype=np.int64)[:, :5]
    f2 = f1[np.asarray(mix, dtype=np.int64)]
    len1, len2 = np.diff(f1, axis=1), np.diff(f2, axis=1)
    gap = len2 - len1
    table = {}
    for g in np.unique(np.abs(gap)).tolist():
        table[g] = random.Random(step).randint(0, g)
    disp = np.vectorize(table.__getitem__, otypes=[np.int64])(np.abs(gap))
    dst = f1[:, :4] + np.where(gap < 0, disp, 0)
    src = f2[:, :4] + np.where(gap >= 0, disp, 0)
    n = np.minimum(len1, len2)
    if limit is not None and f1.size and int(f1.max()) > limit:
        d0, s0 = np.minimum(dst, limit), np.minimum(src, limit)
        wd, ws = np.minimum(dst + n, limit) - d0, np.minimum(src + n, limit) - s0
        bad = (wd != ws) & ~((wd == 0) & (ws == 1))
        if bad.any():
            b, s = (int(v[0]) for v in np.nonzero(bad))
            raise RuntimeError(f"cycle {b}, state {s}: the displaced windows clamp to {int(wd[b, s])} destination and "
                               f"{int(ws[b, s])} source samples in a row of {limit} (the reference raises a shape mismatch here)")
        dst, src, n = d0, s0, np.where(wd == ws, wd, 0)
    out = np.stack([dst, n, src - dst], axis=2)
    return np.ascontiguousarray(out.astype(np.int32))


def processing_order(mix: np.ndarray) -> np.ndarray:
    """Order in which the device visits the cycles: follow the pairing permutation's chains
    (b, mix[b], mix[mix[b]], ...), so that a cycle is read as "partner" and as "itself" by CTAs
    that are in flight together and the second read is served by L2 instead of HBM.  Any
    permutation gives the same output; this one only changes locality.  Computed by the native
    library (20 us per 4096 cycles); the Python loop below is the same walk."""
    try:
        from . import native
        return native.host_processing_order(mix)
    except (ImportError, OSError, RuntimeError):
        return _processing_order_python(mix)


def _processing_order_python(mix: np.ndarray) -> np.ndarray:
    mix = np.asarray(mix, dtype=np.int64)
    n = mix.shape[0]
    seen = np.zeros(n, dtype=bool)
    order = np.empty(n, dtype=np.int32)
    k = 0
    for start in range(n):
        b = start
        while not seen[b]:
            seen[b] = True
            order[k] = b
            k += 1
            b = int(mix[b])
            if not (0 <= b < n):
                break
    return order
