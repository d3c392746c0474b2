"""Not-a-knot cubic spline as a linear map (host side, float64).

The magnitude warp of PCGmix+ (reference ``augmentations.py:674-683``) interpolates ``knot+2``
random ordinates placed at ``np.linspace(0, L-1, knot+2)`` with SciPy's default ``CubicSpline``
(not-a-knot at both ends).  The abscissae are the same for every cycle and channel, and the
spline coefficients are linear in the ordinates, so the whole construction collapses into one
constant matrix per ``(L, knot)``:

    coef[k*4 + i] = sum_j  M[k*4 + i, j] * y[j]          k = piece, i = 0..3

with piece ``k`` evaluated as ``((c0*dt + c1)*dt + c2)*dt + c3``, ``dt = t - x[k]``.  The device
kernel applies ``M`` per (cycle, channel) in its prologue and evaluates the cubic in float64.

The matrix is obtained by solving the standard slope system (continuity of the second
derivative at interior knots, continuity of the third derivative at the second and the
second-to-last knot) for unit ordinate vectors.  Degenerate sizes follow SciPy: two points give
the chord, three points give the parabola through them.
"""
from __future__ import annotations

import functools

import numpy as np


def knot_positions(n_samples: int, knot: int) -> np.ndarray:
    """Abscissae exactly as the reference forms them (``augmentations.py:678``)."""
    return np.linspace(0, n_samples - 1.0, num=knot + 2)


def _slopes_operator(x: np.ndarray) -> np.ndarray:
    """Matrix S (n x n) with ``s = S @ y`` the spline's first derivative at the knots."""
    n = x.shape[0]
    dx = np.diff(x)
    # secant slopes m = D @ y
    D = np.zeros((n - 1, n))
    for i in range(n - 1):
        D[i, i] = -1.0 / dx[i]
        D[i, i + 1] = 1.0 / dx[i]
    if n == 2:
        return np.vstack([D[0], D[0]])
    if n == 3:
        # parabola through three points: slopes of the quadratic at the knots
        A = np.zeros((3, 3))
        B = np.zeros((3, n))
        A[0, 0], A[0, 1] = 1.0, 1.0
        B[0] = 2.0 * D[0]
        A[1, 0], A[1, 1], A[1, 2] = dx[1], 2.0 * (dx[0] + dx[1]), dx[0]
        B[1] = 3.0 * (dx[1] * D[0] + dx[0] * D[1])
        A[2, 1], A[2, 2] = 1.0, 1.0
        B[2] = 2.0 * D[1]
        return np.linalg.solve(A, B)
    A = np.zeros((n, n))
    B = np.zeros((n, n))
    for i in range(1, n - 1):
        A[i, i - 1] = dx[i]
        A[i, i] = 2.0 * (dx[i - 1] + dx[i])
        A[i, i + 1] = dx[i - 1]
        B[i] = 3.0 * (dx[i] * D[i - 1] + dx[i - 1] * D[i])
    # not-a-knot: third derivative continuous across x[1] ...
    d = x[2] - x[0]
    A[0, 0] = dx[1]
    A[0, 1] = d
    B[0] = ((dx[0] + 2.0 * d) * dx[1] * D[0] + dx[0] ** 2 * D[1]) / d
    # ... and across x[n-2]
    d = x[-1] - x[-3]
    A[-1, -1] = dx[-2]
    A[-1, -2] = d
    B[-1] = (dx[-1] ** 2 * D[-2] + (2.0 * d + dx[-1]) * dx[-2] * D[-1]) / d
    return np.linalg.solve(A, B)


def coefficient_matrix_for(x: np.ndarray) -> np.ndarray:
    """``M`` of shape ``((n-1)*4, n)`` for knots at ``x`` (strictly increasing, n >= 2)."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    if n < 2 or not np.all(np.diff(x) > 0):
        raise ValueError("spline abscissae must be strictly increasing (needs L >= 2)")
    dx = np.diff(x)
    S = _slopes_operator(x)
    eye = np.eye(n)
    M = np.zeros(((n - 1) * 4, n))
    for k in range(n - 1):
        secant = (eye[k + 1] - eye[k]) / dx[k]
        t = (S[k] + S[k + 1] - 2.0 * secant) / dx[k]
        M[k * 4 + 0] = t / dx[k]
        M[k * 4 + 1] = (secant - S[k]) / dx[k] - t
        M[k * 4 + 2] = S[k]
        M[k * 4 + 3] = eye[k]
    return M


def _curves(coef_rows: np.ndarray, x: np.ndarray, n_samples: int) -> np.ndarray:
    t = np.arange(n_samples, dtype=np.float64)
    piece = np.clip(np.searchsorted(x, t, side="right") - 1, 0, x.shape[0] - 2)
    dt = t - x[piece]
    c = coef_rows.reshape(coef_rows.shape[:-1] + (x.shape[0] - 1, 4))[..., piece, :]
    return ((c[..., 0] * dt + c[..., 1]) * dt + c[..., 2]) * dt + c[..., 3]


def safe_deviation(x: np.ndarray, M: np.ndarray, n_samples: int) -> float:
    """Largest ``max_j |y_j - 1|`` for which the warp curve is certainly positive at every sample.

    The curve is linear in the ordinates and reproduces constants, so at sample ``t`` it equals
    ``1 + sum_j l_j(t) (y_j - 1)`` with ``l_j`` the curve through the j-th unit vector; hence
    ``|w(t) - 1| <= Lambda * max_j |y_j - 1|`` with ``Lambda = max_t sum_j |l_j(t)|`` taken over
    the samples actually evaluated (1.94 for knot = 4).  Below ``0.999 / Lambda`` every factor is
    ``> 1e-3``: exact zeros (the padding after a cycle) times the factor are ``+0.0`` and a kernel
    that knows which samples are padding may leave them alone.  With sigma = 0.2 about 94 % of all
    (cycle, channel) rows qualify."""
    basis = _curves(M.T.copy(), x, n_samples)                  # (n, L): row j = l_j at every sample
    lebesgue = float(np.abs(basis).sum(axis=0).max())
    return 0.999 / lebesgue if np.isfinite(lebesgue) and lebesgue > 0 else 0.0


@functools.lru_cache(maxsize=64)
def _cached(n_samples: int, knot: int):
    x = knot_positions(n_samples, knot)
    M = coefficient_matrix_for(x)
    pos = np.concatenate([x, [safe_deviation(x, M, n_samples)]])
    pos.setflags(write=False)
    M.setflags(write=False)
    return pos, M


def magwarp_tables(n_samples: int, knot: int):
    """``(knot_pos (knot+3,), coefmat ((knot+1)*4, knot+2))`` float64, cached per ``(L, knot)``.
    ``knot_pos[:knot+2]`` are the abscissae, ``knot_pos[knot+2]`` is :func:`safe_deviation` — the
    layout ``include/pcgmix_b200.h`` asks for."""
    if knot < 0:
        raise ValueError("knot must be >= 0")
    return _cached(int(n_samples), int(knot))


def evaluate(knots: np.ndarray, n_samples: int) -> np.ndarray:
    """Host evaluation of the warp curves through ``M`` — used by tests to compare the linear-map
    formulation with SciPy; ``knots`` (..., knot+2) -> (..., L)."""
    knots = np.asarray(knots, dtype=np.float64)
    pos, M = magwarp_tables(n_samples, knots.shape[-1] - 2)
    return _curves(knots @ M.T, pos[:-1], n_samples)           # coefficients (..., (n-1)*4) -> curves
