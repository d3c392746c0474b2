"""Import the UNMODIFIED reference modules behind module stubs.

TEST / BASELINE INFRASTRUCTURE ONLY.  Two users: ``tests/golden/make_golden*.py`` inside the build container
(source tree at /root/reference), to pin the oracle in ``oracle/pcgmix_oracle.py`` against the reference's
own code; and ``bench.py --impl reference`` / its ``cpu_baseline`` leg, which time the reference's ``augment``
from the byte-compiled copies under ``oracle/_ref/`` (``oracle/build_ref.py``) — the only form in which the
reference exists on the GPU box.  Nothing on the product path or in the ``-m gpu`` tests may call this.

Why stubs: ``augmentations.py:1-23`` imports tkinter, matplotlib,
tsp_solver, audiomentations, python_tsp and the sibling modules latent_space / saliency /
train_model / utils (none importable here), and reads a CSV from a hard-coded absent path at
import time (``augmentations.py:26-28``).  None of that is touched by the hot path
(``augment`` -> ``get_same_label_mix_indices`` / ``get_lambda`` /
``mixup_keepdur_multidim_tensors`` / ``magnitude_warp``), which only needs torch, numpy,
``random`` and scipy.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PCGMIX_REFERENCE_ROOT", "/root/reference")

_STUBS = {
    "tkinter": {"wantobjects": 0},
    "matplotlib": {},
    "matplotlib.pyplot": {},
    "tsp_solver": {},
    "tsp_solver.greedy": {"solve_tsp": None},
    "tsp_solver.util": {"path_cost": None},
    "audiomentations": {"AddGaussianSNR": None},
    "python_tsp": {},
    "python_tsp.heuristics": {"solve_tsp_local_search": None},
    "latent_space": {},
    "saliency": {},
    "train_model": {},
    "utils": {},
    "torchvision": {},
}


COMPILED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "augmentations.py"))


def compiled_reference_available() -> bool:
    return all(os.path.isfile(os.path.join(COMPILED_ROOT, m + ".bytecode")) for m in ("augmentations", "augmentations2d"))


def load_reference(compiled: bool = False):
    """Return ``(augmentations, augmentations2d)`` — the reference's own modules, from the source tree or
    (``compiled=True``) from the sourceless modules under ``oracle/_ref/``."""
    if compiled:
        if not compiled_reference_available():
            raise RuntimeError(f"no compiled reference under {COMPILED_ROOT} (run oracle/build_ref.py in the build container)")
    elif not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if "_pcgmix_ref_augmentations" in sys.modules:
        return sys.modules["_pcgmix_ref_augmentations"], sys.modules["_pcgmix_ref_augmentations2d"]

    import importlib.util

    import pandas as pd

    saved = {}
    for name, attrs in _STUBS.items():
        saved[name] = sys.modules.get(name)
        try:
            if name == "torchvision":
                # use the real one when it imports cleanly, otherwise a stub
                __import__(name)
                continue
        except Exception:
            pass
        mod = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(mod, k, v)
        sys.modules[name] = mod

    real_read_csv = pd.read_csv
    pd.read_csv = lambda *a, **k: pd.DataFrame({"wav": [], "diagnosis": []})
    try:
        mods = []
        for fname, alias in (("augmentations.py", "_pcgmix_ref_augmentations"),
                             ("augmentations2d.py", "_pcgmix_ref_augmentations2d")):
            if compiled:
                import importlib.machinery
                path = os.path.join(COMPILED_ROOT, fname[:-3] + ".bytecode")
                spec = importlib.util.spec_from_loader(alias, importlib.machinery.SourcelessFileLoader(alias, path))
            else:
                spec = importlib.util.spec_from_file_location(alias, os.path.join(REFERENCE_ROOT, fname))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[alias] = mod
            spec.loader.exec_module(mod)
            mods.append(mod)
    finally:
        pd.read_csv = real_read_csv
        # remove the stubs again so they cannot leak into anything else in the process
        for name, old in saved.items():
            if old is None:
                sys.modules.pop(name, None)
            else:
                sys.modules[name] = old
    return tuple(mods)
