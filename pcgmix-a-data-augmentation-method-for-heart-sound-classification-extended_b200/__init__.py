"""PCGmix hot path on B200: segment-aware heart-sound augmentation as hand-written sm_100a
CUDA kernels behind the reference's own ``augment(...)`` call signature.

Modules
-------
augmentations, augmentations2d   drop-in ``augment`` for time series / spectrograms
draws                            host-side replay of the reference's seeded draws
spline                           not-a-knot cubic coefficient map (host, float64)
segmentation                     annotations -> cycles, cut + pad, duration features (device)
native                           ctypes binding of ``csrc/libpcgmix_b200.so`` (the C ABI in
                                 ``include/pcgmix_b200.h``)
sharding                         how a batch stream is split across one-process-per-GPU ranks
synth                            synthetic PhysioNet-shaped cycles and annotations

There is no CPU fallback: every compute entry point raises if the CUDA library is missing or
the tensors are not on a CUDA device.

The directory name contains hyphens (it mirrors the upstream repository name); import it as
``pcgmix_b200`` (alias package at the repository root).
"""

__version__ = "0.1.0"
