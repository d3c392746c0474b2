"""How the augmentation workload is split across GPUs (one process per GPU).

The path has no exchange step: the reference pairs cycles *inside* one batch
(``augmentations.py:943``), so whole batches are the unit of distribution and every rank runs the
same kernel on its own batches with no collective.  Two layouts are used:

  * a stream of batches (BASELINE config 4: 1 M cycles = 245 batches of 4096): batch ``k`` goes
    to rank ``k mod world`` and is augmented with seed ``k`` (the reference's seed is the global
    step counter), so the result does not depend on the number of GPUs;
  * data-parallel training (config 5): every rank draws its own mini-batch and owns a
    ``step_counter``; at step ``k`` all ranks use seed ``k`` (same gate decision and lambda,
    different pairings because the labels differ) — exactly what the reference would do if it
    were launched once per GPU.
"""
from __future__ import annotations


def batches_for_rank(n_batches: int, rank: int, world: int) -> range:
    """Indices of the batches rank ``rank`` augments (round-robin)."""
    if not 0 <= rank < world:
        raise ValueError("rank outside [0, world)")
    return range(rank, n_batches, world)


def step_seed(batch_index: int) -> int:
    """Seed (= the reference's ``step_counter.count``) used for batch ``batch_index``."""
    return int(batch_index)


def rows_for_rank(n_rows: int, rank: int, world: int):
    """Contiguous block ``[lo, hi)`` of ``n_rows`` recordings for ``rank`` (segmentation stage)."""
    base, extra = divmod(n_rows, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)
