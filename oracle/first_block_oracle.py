"""CPU restatement of the first block of the reference's ResNet9-1D — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this; the product
path (``pcgmix_b200/first_block.py`` -> ``pcgmix_first_conv_block``) never does.

What is restated: ``conv_block(in_channels, out_channels)`` (models.py:468-473) as instantiated for ``conv1``
(models.py:523) and applied at models.py:538: ``nn.Conv1d(C, F, kernel_size=3, padding=1)`` ->
``nn.BatchNorm1d(F)`` -> ``nn.ReLU``.  The arithmetic lives in torch (third-party, reference pins torch==1.13.1,
container 2.11): cross-correlation with zero padding; in training mode normalisation with the batch mean and the
BIASED batch variance over (B, L), running statistics updated with ``momentum`` and the UNBIASED variance; in
evaluation mode normalisation with the running statistics.

Pinned: ``tests/golden/first_block_*.npz`` hold what the reference's own ``models.ResNet9(...).conv1`` produced in
this container (``tests/golden/make_golden_first_block.py``); ``tests/test_first_block.py`` holds this restatement
to them (it computes in float64, torch in float32: 2e-6 relative + 2e-6 absolute).
"""
from __future__ import annotations

import numpy as np


def conv_block_forward(x, weight, bias, gamma, beta, running_mean, running_var, training, eps=1e-5, momentum=0.1):
    """``(out float32 (B, F, L), new_running_mean, new_running_var, mean, invstd)`` of the block for ``x`` (B, C, L)."""
    x = np.asarray(x, np.float64)
    w = np.asarray(weight, np.float64)
    B, C, L = x.shape
    F = w.shape[0]
    assert w.shape == (F, C, 3)
    xp = np.zeros((B, C, L + 2))
    xp[:, :, 1:L + 1] = x                                         # padding=1, zeros
    z = np.zeros((B, F, L))
    for k in range(3):                                            # cross-correlation, like nn.Conv1d
        z += np.einsum("fc,bcl->bfl", w[:, :, k], xp[:, :, k:k + L])
    if bias is not None:
        z += np.asarray(bias, np.float64)[None, :, None]
    rm = None if running_mean is None else np.asarray(running_mean, np.float64).copy()
    rv = None if running_var is None else np.asarray(running_var, np.float64).copy()
    if training or rm is None:
        mean = z.mean(axis=(0, 2))
        var = z.var(axis=(0, 2))                                  # biased: what normalises
        n = B * L
        if training and rm is not None:
            rm = (1.0 - momentum) * rm + momentum * mean
            rv = (1.0 - momentum) * rv + momentum * var * (n / (n - 1.0))
    else:
        mean, var = rm, rv
    invstd = 1.0 / np.sqrt(var + eps)
    y = (z - mean[None, :, None]) * invstd[None, :, None]
    if gamma is not None:
        y = y * np.asarray(gamma, np.float64)[None, :, None]
    if beta is not None:
        y = y + np.asarray(beta, np.float64)[None, :, None]
    out = np.maximum(y, 0.0)
    out = np.where(np.isnan(y), np.nan, out)                      # torch's ReLU keeps NaN
    return (out.astype(np.float32), None if rm is None else rm.astype(np.float32),
            None if rv is None else rv.astype(np.float32), mean.astype(np.float32), invstd.astype(np.float32))
