"""Importable alias of ``pcgmix-a-data-augmentation-method-for-heart-sound-classification-extended_b200``.

The real package directory carries the upstream repository's hyphenated name, which the
``import`` statement cannot spell.  This alias makes ``import pcgmix_b200.augmentations`` work by
pointing the package search path at that directory.
"""
import os as _os

_REAL = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "pcgmix-a-data-augmentation-method-for-heart-sound-classification-extended_b200")
if not _os.path.isdir(_REAL):  # pragma: no cover
    raise ImportError(f"package directory not found: {_REAL}")
__path__ = [_REAL]
with open(_os.path.join(_REAL, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_REAL, "__init__.py"), "exec"))
del _f
