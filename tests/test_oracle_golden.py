"""Pin the CPU oracle to the fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  Everything here is bit-exact: the oracle performs the same
fp32 / fp64 operations in the same order as the reference."""
import random

import numpy as np
import pytest
import torch

from oracle import pcgmix_oracle as orc

CASES_1D = [
    "pcgmix_c4_l2500", "pcgmixplus_c4_l2500", "pcgmixplus_default_c2_l800",
    "pcgmixplus_alpha_k2_oddlen", "pcgmixplus_k7_c3", "pcgmix_alpha2_prob",
    "pcgmix_mixall", "pcgmixplus_mixall", "pcgmix_rand", "pcgmixplus_rand",
]
CASES_2D = ["spec_pcgmix_square", "spec_timemask_square", "spec_timemask_default",
            "spec_freqmask_square", "spec_cutout_square", "spec_pcgmix_nonsquare"]


def test_known_answers(golden):
    g = golden("kat_draws")
    assert orc.draw_lambda(1, 7) == g["lambda_alpha1_seed7"] == 0.08912155549916569
    assert orc.draw_lambda(0.5, 3) == g["lambda_alpha05_seed3"]
    assert orc.draw_lambda(0, 3) == 1.0 == g["lambda_alpha0"]
    assert orc.gate_draw(7) == g["uniform7"] == 0.32383276483316237
    assert random.Random(7).sample(list(range(10)), 10) == g["sample7"].tolist() == [5, 2, 6, 9, 0, 7, 4, 1, 3, 8]
    orc.draw_lambda(1, 7)
    knots = orc.draw_knots(2, 4, 4, 0.2)
    assert np.array_equal(knots, g["normals7"])
    np.testing.assert_allclose(knots[0, :, 0], [1.33810514, 0.84221539, 1.2035316, 1.10105987, 1.11091606, 1.33013994], rtol=1e-8)
    assert np.array_equal(orc.same_label_mix_indices(g["labels"], 5), g["same_label_mix_seed5"])


@pytest.mark.parametrize("name", CASES_1D)
@pytest.mark.parametrize("as_torch", [True, False])
def test_augment_1d_matches_reference(golden, name, as_torch):
    g = golden(name)
    data = torch.from_numpy(g["data"].copy()) if as_torch else g["data"].copy()
    out, mix, lam32, _ = orc.augment_1d(str(g["method"]), data, g["labels"], g["frames"], int(g["step"]))
    out = out.numpy() if as_torch else out
    assert np.array_equal(mix, g["mix"])
    assert out.dtype == np.float32
    assert np.array_equal(out.view(np.uint32), g["out"].view(np.uint32)), "oracle differs from the reference bitwise"
    if "(mixAll)" in str(g["method"]):
        ohe = np.eye(2, dtype=np.int64)[g["labels"]]
        assert np.array_equal(orc.soft_targets(ohe, mix, lam32).astype(np.float32), g["target"].astype(np.float32))


@pytest.mark.parametrize("tag", ["samepcg", "samedataset"])
def test_pairing_modifiers(golden, tag):
    g = golden(f"pcgmix_{tag}")
    out, mix, _, _ = orc.augment_1d(str(g["method"]), g["data"].copy(), g["labels"], g["frames"], int(g["step"]),
                                    wav=[str(w) for w in g["wav"]])
    assert np.array_equal(mix, g["mix"])
    assert np.array_equal(out.view(np.uint32), g["out"].view(np.uint32))


def test_gate_failure_returns_same_object(golden):
    g = golden("pcgmix_gate_fail")
    data = g["data"].copy()
    out, mix, lam32, knots = orc.augment_1d(str(g["method"]), data, g["labels"], g["frames"], int(g["step"]))
    assert out is data and mix == [] and lam32 is None and knots is None
    assert bool(g["same_object"]) and g["mix"].size == 0


def test_pairs_edge_cases(golden):
    g = golden("pairs_1d_edge")
    x, fr = g["data"], g["frames"]
    for i in range(x.shape[0]):
        for j in range(x.shape[0]):
            got = orc.mix_pair(x[i], x[j], fr[i], fr[j], np.float32(g["lam"]))
            assert np.array_equal(got.view(np.uint32), g["out"][i, j].view(np.uint32)), (i, j)


def test_magnitude_warp_alone(golden):
    g = golden("magwarp_alone")
    got = orc.magnitude_warp(g["data_blc"], g["knots"])
    assert np.array_equal(got.view(np.uint32), g["out_blc"].view(np.uint32))


@pytest.mark.parametrize("name", CASES_2D)
def test_augment_2d_matches_reference(golden, name):
    g = golden(name)
    out, mix, _ = orc.augment_2d(str(g["method"]), torch.from_numpy(g["data"].copy()), g["labels"],
                                 torch.from_numpy(g["frames"]), int(g["step"]))
    assert np.array_equal(mix, g["mix"])
    assert np.array_equal(out.numpy().view(np.uint32), g["out"].view(np.uint32))


@pytest.mark.parametrize("name", ["pcgmix_c4_l2500", "spec_pcgmix_nonsquare", "pcgmix_alpha2_prob"])
def test_vectorised_cross_check_is_bit_equal(golden, name):
    g = golden(name)
    step = int(g["step"])
    lam32 = orc.lambda_as_float32(orc.draw_lambda(orc.parse_alpha(str(g["method"]), "durratiomixup"), step))
    got = orc.mix_batch_vectorised(g["data"], g["frames"], g["mix"], lam32)
    assert np.array_equal(got.view(np.uint32), g["out"].view(np.uint32))


def test_mix_lambda_one_is_identity_and_no_fma():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((4, 2, 300)).astype(np.float32)
    fr = np.array([[0, 30, 90, 120, 280]] * 4)
    mix = np.array([1, 0, 3, 2])
    assert np.array_equal(orc.mix_batch(x, fr, mix, np.float32(1.0)), x)
    lam = np.float32(0.3)
    got = orc.mix_batch(x, fr, mix, lam)
    expect = x * lam + x[mix] * (np.float32(1) - lam)
    assert np.array_equal(got[..., :280], expect[..., :280])


# ---- cycles whose offsets run past the row (tests/golden/make_golden_long_cycles.py) ------------------
@pytest.mark.parametrize("tag", ["1d", "2d"])
def test_long_cycle_pairs_blend_or_raise_like_the_reference(golden, tag):
    g = golden(f"long_pairs_{tag}")
    x, fr, ok = g["data"], g["frames"], g["ok"]
    assert ok.sum() == 28 and not ok.all()
    for i in range(x.shape[0]):
        for j in range(x.shape[0]):
            if ok[i, j]:
                got = orc.mix_pair(x[i], x[j], fr[i], fr[j], np.float32(g["lam"]))
                assert np.array_equal(got.view(np.uint32), g["out"][i, j].view(np.uint32)), (i, j)
            else:
                with pytest.raises(ValueError):
                    orc.mix_pair(x[i], x[j], fr[i], fr[j], np.float32(g["lam"]))


def test_long_cycle_batch_matches_reference(golden):
    g = golden("long_batch_1d")
    assert (g["frames"][:, 4] > g["data"].shape[-1]).sum() == 4
    for method in ("durratiomixup", "durmixmagwarp(0.2,4)"):
        out, mix, _, _ = orc.augment_1d(method, g["data"].copy(), g["labels"], g["frames"], int(g["step"]))
        assert np.array_equal(mix, g["mix"])
        assert np.array_equal(out.view(np.uint32), g["out_" + method.split("(")[0]].view(np.uint32))
    with pytest.raises(ValueError):
        orc.augment_1d("durratiomixup", g["data"].copy(), g["labels"], g["frames"], int(g["step_raises"]))
