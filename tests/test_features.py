"""Per-cycle classical features (the consumer at train_model.py:519-532).

CPU: the oracle restatement against fixtures produced by executing the reference's own statements.
GPU: the CUDA kernel against the same fixtures.  Tolerances: the amplitude block (columns 0..9) is exact —
maxima are selections and NumPy's float32 ``round(x, 4)`` is reproduced operation for operation; the envelope
block is two different float32 evaluations of the same analytic signal (SciPy: single-precision FFT; here:
circular convolution with the discrete Hilbert kernel), so integrals and means must agree to 2e-5 relative
and the 4-decimal rounded ratios to one unit of the last decimal."""
import numpy as np
import pytest
import torch

from oracle import features_oracle as forc

REL = 2e-5


def test_oracle_matches_reference_statements_bitwise(golden):
    g = golden("cycle_features")
    assert g["features"].shape == (96, 36) and len(g["names"]) == 36
    for i in range(g["data"].shape[0]):
        got = forc.cycle_features(g["data"][i], g["frames"][i])
        assert np.array_equal(got.view(np.uint32), g["features"][i].view(np.uint32)), i


def test_feature_names_are_the_reference_variable_names(golden):
    from pcgmix_b200 import features
    assert [str(n) for n in golden("cycle_features")["names"]] == list(features.FEATURE_NAMES)


def test_cpu_tensor_is_refused():
    from pcgmix_b200 import features
    with pytest.raises(RuntimeError):
        features.cycle_features(torch.zeros(2, 5, 100), torch.zeros(2, 5, dtype=torch.int64))


def _check_block(got, want, what):
    amp, integ, iratio, mean, mratio = slice(0, 10), slice(10, 15), slice(15, 23), slice(23, 28), slice(28, 36)
    if what & 1:
        assert np.array_equal(got[:, amp].view(np.uint32), want[:, amp].view(np.uint32)), "amplitude block must be exact"
    if what & 2:
        # one-sample segments have a zero integral: the reference's ratios are then 0/0 = NaN or x/0 = inf, and so are ours
        g, w = got[:, 10:].astype(np.float64), want[:, 10:].astype(np.float64)
        fin = np.isfinite(w)
        assert np.array_equal(np.isnan(g), np.isnan(w)) and np.array_equal(g[~fin & ~np.isnan(w)], w[~fin & ~np.isnan(w)])
        got, want = np.where(np.isfinite(got), got, 0.0).astype(np.float32), np.where(np.isfinite(want), want, 0.0).astype(np.float32)
        for sl in (integ, mean, mratio):
            err = np.abs(got[:, sl].astype(np.float64) - want[:, sl]) / np.maximum(np.abs(want[:, sl]), 1e-30)
            err[want[:, sl] == 0] = np.abs(got[:, sl].astype(np.float64))[want[:, sl] == 0]
            assert err.max() <= REL, (sl, float(err.max()))
        step = np.abs(got[:, iratio].astype(np.float64) - want[:, iratio])
        assert (step <= 1.0001e-4 + REL * np.abs(want[:, iratio])).all(), float(step.max())
        assert np.mean(got[:, iratio] == want[:, iratio]) > 0.97


@pytest.mark.gpu
@pytest.mark.parametrize("what", [1, 2, 3])
def test_kernel_vs_reference_fixture(golden, what):
    from pcgmix_b200 import features
    g = golden("cycle_features")
    n, length = g["data"].shape
    batch = np.zeros((n, 5, length), np.float32)
    batch[:, 4] = g["data"]                                   # the reference extracts the fifth channel (d[4])
    batch[:, 1] = 9.0                                         # other channels must not matter
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    out = features.cycle_features(torch.from_numpy(batch).cuda(), torch.from_numpy(g["frames"]), channel=4,
                                  amplitude=bool(what & 1), envelope=bool(what & 2), err_flag=err)
    got = out.cpu().numpy()
    assert int(err.item()) == 0
    _check_block(got, g["features"], what)
    if what == 1:
        assert np.isnan(got[:, 10:]).all()                    # blocks not requested stay untouched
    if what == 2:
        assert np.isnan(got[:, :10]).all()


@pytest.mark.gpu
def test_features_of_an_augmented_batch_match_the_oracle():
    """The real consumer: PCGmix+ output of a 5-channel batch -> features of channel 4 with the cycle's own frames."""
    from pcgmix_b200 import augmentations, features, synth
    rng = np.random.default_rng(4)
    b, c, length = 48, 5, 2500
    frames = synth.cycle_frames(rng, b, limit=length)
    x = synth.cycle_signals(rng, frames, (c,), length)
    labels = rng.integers(0, 2, b)

    class Args:
        method, batch_size, sample_rate, num_classes = "durmixmagwarp(0.2,4)", b, 1000, 2

    class Step:
        count = 7

    dev = torch.device("cuda:0")
    ohe = torch.nn.functional.one_hot(torch.from_numpy(labels), 2).to(dev)
    out, _, _, _ = augmentations.augment(Args, torch.from_numpy(x).to(dev), ohe, torch.from_numpy(frames), ["a"] * b, Step, None, dev, None)
    names, table = features.classical_space_features(out, torch.from_numpy(frames), channel=4)
    assert len(names) == 50 and table.shape == (b, 50) and table.dtype == torch.float64
    want = forc.batch_features(out.cpu().numpy(), frames, 4)
    _check_block(table[:, 14:].cpu().numpy().astype(np.float32), want, 3)
    assert table[0, 0].item() == int(frames[0, 4] * 1000 / 1000)          # duration_RR in ms


@pytest.mark.gpu
def test_empty_state_gives_nan_and_a_flag():
    from pcgmix_b200 import features, native
    x = torch.randn(3, 1, 64, device="cuda")
    frames = torch.tensor([[0, 10, 20, 30, 40], [0, 10, 10, 30, 40], [0, 70, 80, 90, 100]])     # empty systole; everything past the row
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    out = features.cycle_features(x, frames, channel=0, err_flag=err).cpu().numpy()
    assert int(err.item()) == native.ERR_EMPTY_STATE
    assert not np.isnan(out[0]).any() and np.isnan(out[1]).all() and np.isnan(out[2]).all()


@pytest.mark.gpu
def test_nan_propagates_like_np_max():
    from pcgmix_b200 import features
    x = torch.randn(1, 1, 200, device="cuda")
    x[0, 0, 15] = float("nan")                                # inside systole
    frames = torch.tensor([[0, 10, 60, 80, 150]])
    out = features.cycle_features(x, frames, channel=0, envelope=False).cpu().numpy()[0]
    want = forc.cycle_features(x[0, 0].cpu().numpy(), frames[0].numpy())
    assert np.isnan(out[1]) and np.isnan(want[1]) and out[0] == want[0] and out[2] == want[2]
    assert np.array_equal(np.isnan(out[:10]), np.isnan(want[:10]))
