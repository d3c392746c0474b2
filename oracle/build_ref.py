"""Recipe for ``oracle/_ref/``: the reference's two hot-path modules, COMPILED where they lie.

    python oracle/build_ref.py            # needs /root/reference (the build container); writes oracle/_ref/*.bytecode

TEST / BASELINE INFRASTRUCTURE.  The reference is 100 % Python, so "building" it means byte-compiling
``/root/reference/augmentations.py`` and ``augmentations2d.py`` with this interpreter into sourceless
modules under ``oracle/_ref/`` — a build output like a ``.so``: git-ignored (no reference source enters the
repository or its history) but not gpurun-ignored, so it travels to the GPU box, where ``/root/reference``
does not exist.  ``bench.py --impl reference`` then times the UNMODIFIED reference ``augment`` on the box's
host cores (``cpu_baseline.kind = "reference"``); without ``oracle/_ref`` it falls back to the oracle port.
``__graft_entry__.build()`` runs this recipe whenever the reference tree is present.
"""
from __future__ import annotations

import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
MODULES = ("augmentations", "augmentations2d")
COMPILED_SUFFIX = ".bytecode"      # (CPython bytecode; not named *.pyc: snapshot tools drop those as interpreter caches)


def build(reference_root: str = None) -> bool:
    root = reference_root or os.environ.get("PCGMIX_REFERENCE_ROOT", "/root/reference")
    if not all(os.path.isfile(os.path.join(root, m + ".py")) for m in MODULES):
        return False
    os.makedirs(OUT, exist_ok=True)
    for m in MODULES:
        py_compile.compile(os.path.join(root, m + ".py"), cfile=os.path.join(OUT, m + COMPILED_SUFFIX), doraise=True, optimize=0)
    with open(os.path.join(OUT, "BUILT_WITH"), "w") as f:
        f.write(f"python {sys.version.split()[0]} magic {py_compile.importlib.util.MAGIC_NUMBER.hex()}\n")
    return True


if __name__ == "__main__":
    print("oracle/_ref built" if build() else "reference tree not found: nothing built")
