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
    assert len(names) == 140 and table.shape == (b, 140) and table.dtype == torch.float64
    want = forc.batch_features(out.cpu().numpy(), frames, 4)
    _check_block(table[:, 14:50].cpu().numpy().astype(np.float32), want, 3)
    _check_psd(table[:, 50:130].cpu().numpy(), forc.batch_psd_features(out.cpu().numpy(), frames, 4))
    _check_moments(table[:, 130:].cpu().numpy(), forc.batch_moment_features(out.cpu().numpy(), frames, 4))
    names50, table50 = features.classical_space_features(out, torch.from_numpy(frames), channel=4, psd=False)
    assert len(names50) == 50 and torch.equal(table50, table[:, :50])
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


# ---------------------------------------------------------------------------------------------------------------
# PSD block (classical.py:358-643): Welch PSD of the beat, the systole and the diastole, envelope integral of the
# spectrum, overall and band means, two ratios.  SciPy: single-precision FFTs; the kernel: direct float64 sums over
# float32 windows.  2e-5 relative on the means, one unit of the last decimal on the two rounded ratios, identical NaNs.

def _check_psd(got, want):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    assert got.shape == want.shape and got.shape[1] == 80
    assert np.array_equal(np.isnan(got), np.isnan(want)), "NaN pattern (empty bands, one-sample segments) must be the reference's"
    ok = ~np.isnan(want)
    means = np.ones(80, bool)
    means[78:] = False
    sel = ok & means[None, :]
    err = np.abs(got[sel] - want[sel]) / np.maximum(np.abs(want[sel]), 1e-30)
    err[want[sel] == 0] = np.abs(got[sel])[want[sel] == 0]
    assert err.max() <= REL, float(err.max())
    sel = ok & ~means[None, :]
    step = np.abs(got[sel] - want[sel])
    assert (step <= 1.0001e-4 + REL * np.abs(want[sel])).all(), float(step.max())


def test_psd_oracle_matches_reference_statements_bitwise(golden):
    import warnings
    g = golden("cycle_psd_features")
    assert g["features"].shape == (64, 80) and len(g["names"]) == 80
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                       # SciPy warns about windows longer than a short segment, NumPy about empty bands
        for i in range(g["data"].shape[0]):
            got = forc.cycle_psd_features(g["data"][i], g["frames"][i])
            want = g["features"][i]
            assert np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(got[~np.isnan(got)], want[~np.isnan(want)]), i
    assert np.isnan(g["features"]).any(), "the fixture must hold short segments with empty bands"


def test_psd_feature_names_are_the_reference_variable_names(golden):
    from pcgmix_b200 import features
    assert [str(n) for n in golden("cycle_psd_features")["names"]] == list(features.PSD_FEATURE_NAMES)


def test_psd_cpu_tensor_is_refused():
    from pcgmix_b200 import features
    with pytest.raises(RuntimeError):
        features.cycle_psd_features(torch.zeros(2, 5, 100), torch.zeros(2, 5, dtype=torch.int64))


@pytest.mark.gpu
def test_psd_kernel_vs_reference_fixture(golden):
    from pcgmix_b200 import features
    g = golden("cycle_psd_features")
    data = torch.from_numpy(np.ascontiguousarray(g["data"][:, None, :])).cuda()          # (B, 1, L)
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    out = features.cycle_psd_features(data, torch.from_numpy(g["frames"]), channel=0, err_flag=err)
    assert out.shape == (64, 80) and out.dtype == torch.float32 and int(err.item()) == 0
    _check_psd(out.cpu().numpy(), g["features"])


@pytest.mark.gpu
def test_psd_kernel_on_a_strided_channel_and_device_frames(golden):
    from pcgmix_b200 import features
    g = golden("cycle_psd_features")
    rng = np.random.default_rng(3)
    batch = rng.standard_normal((64, 5, 2500)).astype(np.float32)
    batch[:, 4] = g["data"]
    frames_dev = torch.from_numpy(g["frames"].astype(np.int32)).cuda()
    out = features.cycle_psd_features(torch.from_numpy(batch).cuda(), frames_dev, channel=4)
    _check_psd(out.cpu().numpy(), g["features"])


@pytest.mark.gpu
def test_psd_empty_segment_gives_nan_and_a_flag():
    from pcgmix_b200 import features, native
    x = torch.randn(3, 1, 600, device="cuda")
    frames = torch.tensor([[0, 100, 300, 380, 590], [0, 100, 100, 180, 590], [0, 100, 300, 380, 380]])     # empty systole; empty diastole
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    out = features.cycle_psd_features(x, frames, channel=0, err_flag=err).cpu().numpy()
    assert int(err.item()) == native.ERR_EMPTY_STATE
    assert np.isfinite(out[0, :2]).all() and np.isnan(out[1]).all() and np.isnan(out[2]).all()
    want = forc.cycle_psd_features(x[0, 0].cpu().numpy(), frames[0].numpy())
    _check_psd(out[:1], want[None])


@pytest.mark.gpu
def test_psd_other_sampling_rate_moves_the_bands():
    """``fs`` is the reference's ``Fs`` argument of ``signal.welch``: at 2 kHz the bins are twice as far apart."""
    import warnings
    from pcgmix_b200 import features
    rng = np.random.default_rng(11)
    x = rng.standard_normal((4, 1, 3000)).astype(np.float32)
    frames = np.array([[0, 200, 700, 900, 2400], [0, 250, 760, 1000, 2999], [0, 180, 436, 600, 1500], [0, 300, 555, 800, 2000]])
    out = features.cycle_psd_features(torch.from_numpy(x).cuda(), torch.from_numpy(frames), channel=0, fs=2000).cpu().numpy()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = np.stack([forc.cycle_psd_features(x[i, 0], frames[i], 2000) for i in range(4)])
    _check_psd(out, want)


# ---------------------------------------------------------------------------------------------------------------
# Moments block (classical.py:893-905): scipy.stats.skew / kurtosis of the beat and the four states, float32 like
# SciPy's for float32 rows.  2e-5 relative + 2e-6 absolute: a skewness near zero is a difference of large terms in
# both implementations.

def _check_moments(got, want):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    assert got.shape == want.shape and got.shape[1] == 10
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    assert (np.abs(got[ok] - want[ok]) <= REL * np.abs(want[ok]) + 2e-6).all(), float(np.abs(got[ok] - want[ok]).max())


def test_moment_oracle_matches_reference_statements_bitwise(golden):
    import warnings
    g, m = golden("cycle_psd_features"), golden("cycle_moment_features")
    assert m["features"].shape == (64, 10)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = forc.batch_moment_features(g["data"][:, None, :], g["frames"], 0)
    assert np.array_equal(got, m["features"])


def test_moment_feature_names_are_the_reference_variable_names(golden):
    from pcgmix_b200 import features
    assert [str(n) for n in golden("cycle_moment_features")["names"]] == list(features.MOMENT_FEATURE_NAMES)


@pytest.mark.gpu
def test_moment_kernel_vs_reference_fixture(golden):
    from pcgmix_b200 import features
    g, m = golden("cycle_psd_features"), golden("cycle_moment_features")
    rng = np.random.default_rng(4)
    batch = rng.standard_normal((64, 5, 2500)).astype(np.float32)
    batch[:, 4] = g["data"]
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    out = features.cycle_moment_features(torch.from_numpy(batch).cuda(), torch.from_numpy(g["frames"]), channel=4, err_flag=err)
    assert out.shape == (64, 10) and out.dtype == torch.float32 and int(err.item()) == 0
    _check_moments(out.cpu().numpy(), m["features"])


@pytest.mark.gpu
def test_moment_kernel_flat_and_empty_segments():
    import warnings
    from pcgmix_b200 import features, native
    x = torch.randn(3, 1, 400, device="cuda")
    x[1, 0, 100:200] = 0.25                                   # a constant systole: no variance -> NaN like SciPy
    frames = torch.tensor([[0, 100, 200, 300, 390], [0, 100, 200, 300, 390], [0, 100, 100, 300, 390]])     # third: empty systole
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    out = features.cycle_moment_features(x, frames, channel=0, err_flag=err).cpu().numpy()
    assert int(err.item()) == native.ERR_EMPTY_STATE
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = np.stack([forc.cycle_moment_features(x[i, 0].cpu().numpy(), frames[i].numpy()) for i in range(2)])
    _check_moments(out[:2], want)
    assert np.isnan(out[1, 2]) and np.isnan(out[1, 7])        # skew / kurtosis of the constant systole
    assert np.isnan(out[2, 2]) and np.isnan(out[2, 7]) and np.isfinite(out[2, [0, 1, 3, 4]]).all()
