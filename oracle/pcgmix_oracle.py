"""CPU oracle for the PCGmix / PCGmix+ augmentation hot path.

TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / ``--impl reference`` legs may import this module; the product
package never does (it has no CPU fallback and fails loudly without its CUDA library).

This is a restatement, in this repo's own words, of what the reference computes on the path
named in BASELINE.json (all citations are into /root/reference):

  * gate / pairing / lambda / knot draws ......... augmentations.py:500-514, 659-666, 677,
                                                    866-871, 932-939 (2D: augmentations2d.py:19-26,
                                                    251-265, 398-405)
  * per-pair state mixer (1D and 2D) ............. augmentations.py:289-304,
                                                    augmentations2d.py:206-221
  * magnitude warp ............................... augmentations.py:674-683
  * dispatcher branches .......................... augmentations.py:864-929 (PCGmix+),
                                                    931-981 (PCGmix), augmentations2d.py:397-427,
                                                    286-395 (durmix{cutout,timemask,freqmask})

PARITY PIN: the reference ships no tests or golden vectors for this path,
so the pin is the reference's own code executed in the build container through
``oracle/ref_import.py``; ``tests/golden/make_golden.py`` stores its outputs as fixtures and
``tests/test_oracle_golden.py`` holds this file to them bit-for-bit.  Third-party arithmetic
the reference leans on (scipy ``CubicSpline`` default not-a-knot boundary, NumPy legacy
``RandomState`` stream, CPython ``random.Random``) is called here exactly as the reference
calls it, so the oracle inherits the same results.

The loops deliberately keep the reference's shape (one Python iteration per cardiac cycle,
one SciPy spline per cycle and channel): ``bench.py`` times this module as the host-CPU
baseline, and a restructured/vectorised port would not be representative of the reference.
"""
from __future__ import annotations

import random

import numpy as np

STATE_NAMES = ("S1", "systole", "S2", "diastole")


# --------------------------------------------------------------------------------------
# host draws
# --------------------------------------------------------------------------------------
def gate_draw(step: int) -> float:
    """Probability-gate variate: a fresh ``random.Random(step)`` uniform (augmentations.py:936-937)."""
    return random.Random(step).uniform(0, 1)


def parse_probability(method: str) -> float:
    """Text after the last ``+`` is the apply-probability, default 1 (augmentations.py:932-935)."""
    parts = method.split("+")
    return float(parts[-1]) if len(parts) > 1 else 1.0


def same_label_mix_indices(labels, step: int) -> np.ndarray:
    """Within-class seeded permutation (augmentations.py:500-514).

    ``labels`` is the per-cycle class id (the reference recovers it as the arg-max of the
    one-hot target, :501).  Classes are visited in order of first appearance; every class
    starts from a *fresh* ``random.Random(step)`` and is shuffled with ``sample(k=n)``.
    """
    labels = np.asarray(labels).reshape(-1)
    groups = {}
    for i, lab in enumerate(labels.tolist()):
        groups.setdefault(lab, []).append(i)
    mix = np.arange(0, labels.shape[0], 1)
    for members in groups.values():
        mix[members] = random.Random(step).sample(list(mix[members]), len(members))
    return mix


def mix_all_indices(size: int, step: int) -> np.ndarray:
    """``(mixAll)`` pairing: one seeded permutation of the whole batch (augmentations.py:883-884)."""
    return np.array(random.Random(step).sample(list(np.arange(0, size, 1)), size))


def same_wav_mix_indices(wav, step: int) -> np.ndarray:
    """``(samePCG)`` pairing: permutation within each recording name (augmentations.py:528-540)."""
    groups = {}
    for i, w in enumerate(wav):
        groups.setdefault(w, []).append(i)
    mix = np.arange(0, len(wav), 1)
    for members in groups.values():
        mix[members] = random.Random(step).sample(list(mix[members]), len(members))
    return mix


def same_dataset_mix_indices(labels, wav, step: int) -> np.ndarray:
    """``(sameDataset)`` pairing: groups keyed by first letter of the recording name and the
    class (augmentations.py:542-556)."""
    labels = np.asarray(labels).reshape(-1)
    groups = {}
    for i, (w, lab) in enumerate(zip(wav, labels.tolist())):
        groups.setdefault(f"{w[0]}_{lab}", []).append(i)
    mix = np.arange(0, len(wav), 1)
    for members in groups.values():
        mix[members] = random.Random(step).sample(list(mix[members]), len(members))
    return mix


def draw_lambda(alpha: float, step: int) -> float:
    """Mixing weight (augmentations.py:659-666).  Re-seeds NumPy's *global* legacy stream, which
    is the stream the magnitude-warp knots are drawn from right afterwards."""
    if alpha > 0.0:
        np.random.seed(step)
        return float(np.random.beta(alpha, alpha))
    return 1.0


def lambda_as_float32(lam: float) -> np.float32:
    """The reference stores lambda in a float32 array before use (augmentations.py:962-963)."""
    return np.array(np.ones(1) * lam).astype("float32")[0]


def draw_knots(batch: int, knot: int, channels: int, sigma: float) -> np.ndarray:
    """Knot ordinates N(1, sigma) of shape (B, knot+2, C) from the global stream
    (augmentations.py:677) — call right after :func:`draw_lambda`."""
    return np.random.normal(loc=1.0, scale=sigma, size=(batch, knot + 2, channels))


# --------------------------------------------------------------------------------------
# per-pair state mixer (augmentations.py:289-304 / augmentations2d.py:206-221)
# --------------------------------------------------------------------------------------
def mix_pair(d1, d2, f1, f2, lam):
    """Blend the four heart states of cycle ``d1`` with its partner ``d2``.

    Works on torch tensors or NumPy arrays of shape (..., T); ``f1``/``f2`` are the five
    cumulative state offsets of each cycle and ``lam`` a float32 scalar/1-element tensor.
    The result keeps ``d1``'s durations: in state ``s`` only the first
    ``min(len1_s, len2_s)`` samples are blended (aligned at the state start), the rest stays
    a copy of ``d1``.  Slices use Python semantics, so the expression is evaluated in fp32
    as ``(a*lam) + (b*(1-lam))`` with every operation rounded separately.
    """
    out = d1.clone() if hasattr(d1, "clone") else d1.copy()
    for s in range(4):
        n = min(f1[s + 1] - f1[s], f2[s + 1] - f2[s])
        a, b = f1[s], f2[s]
        out[..., a:a + n] = out[..., a:a + n] * lam + d2[..., b:b + n] * (1 - lam)
    return out


def mix_pair_rand(d1, d2, f1, f2, lam, step):
    """The ``(rand)`` variant of the per-pair mixer (augmentations.py:305-337): the shorter state is
    blended at a seeded offset inside the longer one.  ``random.Random(step).randint(0, |gap|)`` is
    drawn from a fresh generator for every state; a longer partner state shifts the READ window,
    a longer own state shifts the WRITE window."""
    out = d1.clone() if hasattr(d1, "clone") else d1.copy()
    for s in range(4):
        n = min(f1[s + 1] - f1[s], f2[s + 1] - f2[s])
        gap = (f2[s + 1] - f2[s]) - (f1[s + 1] - f1[s])
        disp = random.Random(step).randint(0, np.abs(gap))
        if gap >= 0:
            a, b = f1[s], f2[s] + disp
        else:
            a, b = f1[s] + disp, f2[s]
        out[..., a:a + n] = out[..., a:a + n] * lam + d2[..., b:b + n] * (1 - lam)
    return out


def mix_batch(data, frames, mix_indices, lam32, rand_step=None):
    """The reference's per-cycle loop (augmentations.py:969-977, augmentations2d.py:419-426)."""
    is_torch = type(data).__module__.split(".")[0] == "torch"
    if is_torch:
        import torch
    frames_np = frames.numpy() if hasattr(frames, "numpy") else np.asarray(frames)
    if is_torch:
        out = torch.zeros(tuple(data.shape), dtype=data.dtype)
        lam = torch.from_numpy(np.full((1,) * (data.dim() - 1), lam32, dtype=np.float32))
    else:
        out = np.zeros(data.shape, dtype=data.dtype)
        lam = np.float32(lam32)
    partners = data[mix_indices]
    partner_frames = frames_np[mix_indices]
    for i in range(data.shape[0]):
        if rand_step is None:
            out[i] = mix_pair(data[i], partners[i], frames_np[i], partner_frames[i], lam)
        else:
            out[i] = mix_pair_rand(data[i], partners[i], frames_np[i], partner_frames[i], lam, rand_step)
    return out


# --------------------------------------------------------------------------------------
# magnitude warp (augmentations.py:674-683)
# --------------------------------------------------------------------------------------
def magnitude_warp(x_blc: np.ndarray, knots: np.ndarray) -> np.ndarray:
    """``x_blc`` is (B, L, C) float32, ``knots`` (B, knot+2, C) float64.

    One not-a-knot cubic spline (SciPy default boundary) per cycle and channel through the
    knots placed at ``linspace(0, L-1, knot+2)``, evaluated at every integer sample; the
    product is formed in float64 and stored to float32.
    """
    from scipy.interpolate import CubicSpline

    n_samples = x_blc.shape[1]
    sample_pos = np.arange(n_samples)
    knot_pos = (np.ones((x_blc.shape[2], 1)) * np.linspace(0, n_samples - 1.0, num=knots.shape[1])).T
    out = np.zeros_like(x_blc)
    for i, cyc in enumerate(x_blc):
        curve = np.array([
            CubicSpline(knot_pos[:, c], knots[i, :, c])(sample_pos) for c in range(x_blc.shape[2])
        ]).T
        out[i] = cyc * curve
    return out


def warp_curves(n_samples: int, knots: np.ndarray) -> np.ndarray:
    """The float64 warping curves alone, shape (B, C, L) — used by tests to look at the spline
    stage in isolation."""
    from scipy.interpolate import CubicSpline

    pos = np.linspace(0, n_samples - 1.0, num=knots.shape[1])
    t = np.arange(n_samples)
    out = np.empty((knots.shape[0], knots.shape[2], n_samples))
    for b in range(knots.shape[0]):
        for c in range(knots.shape[2]):
            out[b, c] = CubicSpline(pos, knots[b, :, c])(t)
    return out


# --------------------------------------------------------------------------------------
# method-string parsing shared by the dispatchers
# --------------------------------------------------------------------------------------
def parse_alpha(method: str, branch: str) -> float:
    """``(alpha=a)<branch>`` prefix, default 1 (augmentations.py:895-897, 958-960)."""
    if len(method.split("(alpha=")) > 1:
        return float(method.split("(alpha=")[1].split(")" + branch)[0])
    return 1.0


def parse_magwarp(method: str):
    """``durmixmagwarp(sigma,knot)``, defaults 0.2 and 4 (augmentations.py:919-923)."""
    sigma, knot = 0.2, 4
    if len(method.split("durmixmagwarp(")) > 1:
        sigma = float(method.split("durmixmagwarp(")[1].split(",")[0])
        knot = int(method.split(",")[1].split(")")[0])
    return sigma, knot


def pick_pairing(method: str, labels, wav, step: int) -> np.ndarray:
    """Pairing selection in the reference's order (augmentations.py:875-884, 943-952).  The
    modifiers that need files or models absent from the reference repo are not restated."""
    for unsupported in ("(sameCVD)", "(closestbins=", "(closestknn=", "(salopt"):
        if unsupported in method:
            raise NotImplementedError(f"oracle does not restate the {unsupported} modifier")
    mix = same_label_mix_indices(labels, step)
    if "(samePCG)" in method:
        mix = same_wav_mix_indices(wav, step)
    if "(sameDataset)" in method:
        mix = same_dataset_mix_indices(labels, wav, step)
    if "(mixAll)" in method:
        mix = mix_all_indices(len(labels), step)
    return mix


# --------------------------------------------------------------------------------------
# dispatchers
# --------------------------------------------------------------------------------------
def augment_1d(method: str, data, labels, frames, step: int, wav=None):
    """PCGmix / PCGmix+ on time series, ``data`` (B, C, L) float32 torch-CPU tensor or ndarray.

    Returns ``(data_out, mix_indices, lam32, knots)``; on a failed probability gate returns
    ``(data, [], None, None)`` with ``data`` the same object (augmentations.py:938-939).
    Follows augmentations.py:864-929 (``durmixmagwarp``) and :931-981 (``durratiomixup``);
    like the reference, ``durmixmagwarp`` is tested first.
    """
    is_torch = type(data).__module__.split(".")[0] == "torch"
    if "durmixmagwarp" in method:
        branch = "durmixmagwarp"
    elif "durratiomixup" in method:
        branch = "durratiomixup"
    else:
        raise ValueError(f"not a PCGmix method: {method!r}")
    if gate_draw(step) >= parse_probability(method):
        return data, [], None, None
    labels = np.asarray(labels).reshape(-1)
    mix = pick_pairing(method, labels, wav, step)
    lam = draw_lambda(parse_alpha(method, branch), step)
    lam32 = lambda_as_float32(lam)
    out = mix_batch(data, frames, mix, lam32, rand_step=step if "(rand)" in method else None)
    knots = None
    if branch == "durmixmagwarp":
        sigma, knot = parse_magwarp(method)
        as_np = out.detach().cpu().numpy() if is_torch else out
        as_np = np.transpose(as_np, (0, 2, 1))
        knots = draw_knots(as_np.shape[0], knot, as_np.shape[2], sigma)
        as_np = np.transpose(magnitude_warp(as_np, knots), (0, 2, 1))
        if is_torch:
            import torch
            out = torch.from_numpy(np.ascontiguousarray(as_np))
        else:
            out = as_np
    return out, mix, lam32, knots


def soft_targets(target_ohe, mix, lam32):
    """``(mixAll)`` label blend (augmentations.py:915-917, 978-980); float32 result."""
    t = np.asarray(target_ohe)
    lam = np.float32(lam32)
    return t * lam + t[mix] * (1 - lam)


def mask_draws(step: int, region_max: float):
    """Seeded mask geometry shared by the 2D composites (augmentations2d.py:317-319, 354-356,
    390-391): ``gap ~ U(0, region_max)`` from ``Random(step+131071)``, start fraction
    ``~ U(0, 1-gap)`` from ``Random(step+13119)``."""
    gap = random.Random(step + 131071).uniform(0, region_max)
    frac1 = random.Random(step + 13119).uniform(0, 1 - gap)
    return gap, frac1


def _clamp01(v: float) -> float:
    return min(max(v, 0), 1)


def augment_2d(method: str, data, labels, frames, step: int):
    """PCGmix on spectrograms, ``data`` (B, Ch, F, T) (augmentations2d.py:397-427) and the three
    composites that zero a seeded box afterwards (:286-395).

    Departure from the reference, on purpose: the reference allocates its output with
    ``spec_dim2 = data.shape[2]`` (:409), which only works for square spectrograms; the oracle
    uses the true T so that BASELINE config 3 (64 x 250) is defined.  For square inputs the
    two agree (checked against the live reference in the golden fixtures).

    Returns ``(data_out, mix_indices, lam32)`` or ``(data, [], None)`` on a failed gate.
    """
    branches = ("durmixcutout", "durmixtimemask", "durmixfreqmask", "durratiomixup")
    branch = next((b for b in branches if b in method), None)
    if branch is None:
        raise ValueError(f"not a 2D PCGmix method: {method!r}")
    if gate_draw(step) >= parse_probability(method):
        return data, [], None
    labels = np.asarray(labels).reshape(-1)
    mix = same_label_mix_indices(labels, step)
    lam32 = lambda_as_float32(draw_lambda(1, step))
    out = mix_batch(data, frames, mix, lam32)
    frames_np = frames.numpy() if hasattr(frames, "numpy") else np.asarray(frames)
    n_freq = data.shape[2]
    if branch == "durmixcutout":
        t_max, f_max = 0.2, 0.2
        if len(method.split("cutout(")) > 1:
            t_max = _clamp01(float(method.split("cutout(")[1].split(",")[0]))
            f_max = _clamp01(float(method.split(",")[1].split(")")[0]))
        t_gap, t_frac1 = mask_draws(step, t_max)
        f_gap, f_frac1 = mask_draws(step, f_max)
        h1 = int(n_freq * f_frac1)
        h2 = min(n_freq, h1 + int(f_gap * n_freq))
        for i in range(out.shape[0]):
            beat = frames_np[i][-1]
            out[i][:, h1:h2, int(t_frac1 * beat):int((t_frac1 + t_gap) * beat)] = 0
    elif branch == "durmixtimemask":
        r_max = 0.2
        if len(method.split("timemask(")) > 1:
            r_max = _clamp01(float(method.split("timemask(")[1].split(")")[0]))
        gap, frac1 = mask_draws(step, r_max)
        for i in range(out.shape[0]):
            beat = frames_np[i][-1]
            out[i][:, :, int(frac1 * beat):int((frac1 + gap) * beat)] = 0
    elif branch == "durmixfreqmask":
        r_max = 0.2
        if len(method.split("freqmask(")) > 1:
            r_max = _clamp01(float(method.split("freqmask(")[1].split(")")[0]))
        gap, frac1 = mask_draws(step, r_max)
        h1 = int(n_freq * frac1)
        h2 = min(n_freq, h1 + int(gap * n_freq))
        out[:, :, h1:h2, :] = 0
    return out, mix, lam32


# --------------------------------------------------------------------------------------
# vectorised cross-check (not the reference's structure; used only to test big batches fast)
# --------------------------------------------------------------------------------------
def mix_batch_vectorised(data: np.ndarray, frames: np.ndarray, mix: np.ndarray, lam32) -> np.ndarray:
    """Same result as :func:`mix_batch` on ndarrays, computed with masks instead of a Python
    loop.  ``tests/test_oracle_golden.py`` holds it bit-equal to :func:`mix_batch`."""
    data = np.asarray(data)
    frames = np.asarray(frames).astype(np.int64)
    n_t = data.shape[-1]
    lam = np.float32(lam32)
    one_minus = np.float32(1) - lam
    f2 = frames[mix]
    t = np.arange(n_t, dtype=np.int64)[None, :]
    out = data.copy()
    lead = (slice(None),) + (None,) * (data.ndim - 2)
    for s in range(4):
        n = np.minimum(frames[:, s + 1] - frames[:, s], f2[:, s + 1] - f2[:, s])
        sel = (t >= frames[:, s:s + 1]) & (t < (frames[:, s] + n)[:, None])          # (B, T)
        src = np.clip(t + (f2[:, s] - frames[:, s])[:, None], 0, n_t - 1)            # (B, T)
        idx = np.broadcast_to(src[lead + (slice(None),)], data.shape)
        partner = np.take_along_axis(data[mix], idx, axis=-1)
        blended = data * lam + partner * one_minus
        out = np.where(np.broadcast_to(sel[lead + (slice(None),)], data.shape), blended, out)
    return out
