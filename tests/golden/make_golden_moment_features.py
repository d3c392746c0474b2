"""Golden vectors for the skewness / kurtosis block of ``classical.feature_vector_seg`` (classical.py:893-905:
``scipy.stats.skew`` and ``scipy.stats.kurtosis`` of the whole beat and of the four states), produced by EXECUTING the
reference's own statements verbatim on the cycles of ``cycle_psd_features.npz`` (same generator, same seed: only the
outputs are stored here).

Run in the build container:  python tests/golden/make_golden_moment_features.py
"""
from __future__ import annotations

import os
import sys
import textwrap
import warnings

import numpy as np
from scipy import stats

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_psd_features as psd  # noqa: E402

REF = os.environ.get("PCGMIX_REFERENCE_ROOT", "/root/reference")
SEGMENTS = ("RR", "S1", "systole", "S2", "diastole")
NAMES = ["skew_" + s for s in SEGMENTS] + ["kurtosis_" + s for s in SEGMENTS]


def reference_statements():
    src = open(os.path.join(REF, "classical.py")).read().split("\n")
    i0 = next(i for i, l in enumerate(src) if l.startswith("def feature_vector_seg("))
    i_seg = next(i for i in range(i0, len(src)) if src[i].strip().startswith("diastole = data[frames[3]:frames[4]]"))
    j0 = next(i for i in range(i_seg, len(src)) if src[i].strip().startswith("skew_RR = stats.skew(RR)"))
    j1 = next(i for i in range(j0, len(src)) if src[i].strip().startswith("kurtosis_diastole = stats.kurtosis(diastole)"))
    body = src[i0 + 1:i_seg + 1] + src[j0:j1 + 1]
    return compile(textwrap.dedent("\n".join(body)), "<classical.py feature_vector_seg, moments block>", "exec")


def main():
    code = reference_statements()
    data, frames = psd.cycles(np.random.default_rng(20261020), 64, 2500)
    stored = np.load(os.path.join(HERE, "cycle_psd_features.npz"))
    assert np.array_equal(stored["data"], data) and np.array_equal(stored["frames"], frames)
    feats = np.zeros((len(data), len(NAMES)), np.float64)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i in range(len(data)):
            ns = dict(np=np, stats=stats, data=data[i], frames=frames[i])
            exec(code, ns)
            for k, name in enumerate(NAMES):
                assert ns[name].dtype == np.float32
                feats[i, k] = ns[name]
    np.savez_compressed(os.path.join(HERE, "cycle_moment_features.npz"),
                        entry=np.array("classical.py feature_vector_seg, skewness / kurtosis block, executed verbatim on the cycles of cycle_psd_features.npz"),
                        features=feats, names=np.array(NAMES))
    print("moment features:", feats.shape, "nan:", int(np.isnan(feats).sum()))


if __name__ == "__main__":
    main()
