"""First block of the reference's ResNet9-1D on augmented cycles (models.py:468-473, :523).

CPU: the oracle restatement against fixtures the UNMODIFIED reference module produced
(tests/golden/make_golden_first_block.py), and the identity the kernels rest on (batch statistics of the
convolution's output from second moments of the input patches).
GPU: ``pcgmix_first_conv_block`` through the C ABI and through the host mirror against the same fixtures.
Tolerance: float32 FMAs in a different order from torch's CPU kernels, normalised values of order one:
2e-5 relative + 2e-5 absolute on the output, 1e-5 relative + 1e-6 absolute on the statistics."""
import numpy as np
import pytest
import torch

from oracle import first_block_oracle as fborc

@pytest.fixture(autouse=True)
def _float32_convolutions():
    """torch's own GPU convolution must be a float32 one to be a yardstick (TF32 is its default)."""
    saved = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved


CASES = ["first_block_c4_f64", "first_block_c4_f64_odd", "first_block_c1_f16", "first_block_c2_f7"]
REL, ABS = 2e-5, 2e-5


def _close(got, want, rel=REL, abs_=ABS):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return float(np.max(np.abs(got - want) - rel * np.abs(want))) <= abs_


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_module(golden, name):
    g = golden(name)
    out, rm, rv, _, _ = fborc.conv_block_forward(g["x"], g["weight"], g["bias"], g["gamma"], g["beta"],
                                                 g["running_mean_before"], g["running_var_before"], True,
                                                 float(g["eps"]), float(g["momentum"]))
    assert _close(out, g["out_train"], 2e-6, 2e-6)
    assert _close(rm, g["running_mean_after"], 1e-6, 1e-7) and _close(rv, g["running_var_after"], 1e-6, 1e-7)
    out, rm2, rv2, _, _ = fborc.conv_block_forward(g["x"], g["weight"], g["bias"], g["gamma"], g["beta"],
                                                   g["running_mean_after"], g["running_var_after"], False,
                                                   float(g["eps"]), float(g["momentum"]))
    assert _close(out, g["out_eval"], 2e-6, 2e-6)
    assert np.array_equal(rm2, g["running_mean_after"]) and np.array_equal(rv2, g["running_var_after"])
    assert int(g["num_batches_tracked_after"]) == 1


def test_fixture_exercises_relu_and_negative_scale(golden):
    g = golden("first_block_c4_f64")
    assert 0.2 < float((g["out_train"] == 0).mean()) < 0.8 and g["gamma"][1] < 0


@pytest.mark.parametrize("name", CASES)
def test_batch_statistics_follow_from_patch_moments(golden, name):
    """mean_f = bias_f + w_f . E[v], var_f = w_f^T Cov[v] w_f with v the zero-padded 3C-sample patch: the identity that
    lets the device normalise without ever holding the un-normalised output."""
    g = golden(name)
    x = g["x"].astype(np.float64)
    B, C, L = x.shape
    xp = np.zeros((B, C, L + 2))
    xp[:, :, 1:L + 1] = x
    v = np.stack([xp[:, c, k:k + L] for c in range(C) for k in range(3)], axis=-1).reshape(B * L, 3 * C)
    mu, second = v.mean(0), v.T @ v / (B * L)
    w = g["weight"].astype(np.float64).reshape(-1, 3 * C)
    mean = g["bias"] + w @ mu
    var = np.einsum("fi,ij,fj->f", w, second - np.outer(mu, mu), w)
    _, _, _, want_mean, want_invstd = fborc.conv_block_forward(g["x"], g["weight"], g["bias"], g["gamma"], g["beta"],
                                                               None, None, True, float(g["eps"]))
    assert _close(mean, want_mean, 1e-6, 1e-7)
    assert _close(1.0 / np.sqrt(var + float(g["eps"])), want_invstd, 1e-6, 1e-7)


def _reference_block(g, device):
    """An equivalent torch module (same layer types as models.py:468-473) holding the fixture's parameters."""
    F, C, _ = g["weight"].shape
    block = torch.nn.Sequential(torch.nn.Conv1d(C, F, kernel_size=3, padding=1), torch.nn.BatchNorm1d(F),
                                torch.nn.ReLU(inplace=True))
    with torch.no_grad():
        block[0].weight.copy_(torch.from_numpy(g["weight"]))
        block[0].bias.copy_(torch.from_numpy(g["bias"]))
        block[1].weight.copy_(torch.from_numpy(g["gamma"]))
        block[1].bias.copy_(torch.from_numpy(g["beta"]))
        block[1].running_mean.copy_(torch.from_numpy(g["running_mean_before"]))
        block[1].running_var.copy_(torch.from_numpy(g["running_var_before"]))
    return block.to(device)


def test_cpu_tensor_is_refused(golden):
    from pcgmix_b200 import first_block
    g = golden("first_block_c2_f7")
    with pytest.raises(RuntimeError):
        first_block.first_conv_block(_reference_block(g, "cpu"), torch.from_numpy(g["x"]))


def test_workspace_size_is_a_host_side_query():
    """No device needed: moments (3C + 3C(3C+1)/2 doubles, padded to 16 bytes) + one interleaved record per filter pair."""
    from pcgmix_b200 import native
    assert native.first_conv_block_workspace(4, 64) == 90 * 8 + 32 * 28 * 4
    assert native.first_conv_block_workspace(1, 16) == 16 * ((9 * 8 + 15) // 16) + 8 * 8 * 4
    assert native.first_conv_block_workspace(2, 7) == 16 * ((27 * 8 + 15) // 16) + 4 * 16 * 4        # odd filter count: 4 pairs
    for bad in ((0, 8), (5, 8), (4, 0), (4, native.MAX_FIRST_BLOCK_FILTERS + 1)):
        with pytest.raises(ValueError):
            native.first_conv_block_workspace(*bad)


def test_other_blocks_are_refused():
    from pcgmix_b200 import first_block
    pooled = torch.nn.Sequential(torch.nn.Conv1d(4, 8, 3, padding=1), torch.nn.BatchNorm1d(8), torch.nn.ReLU(), torch.nn.MaxPool1d(2))
    wide = torch.nn.Sequential(torch.nn.Conv1d(4, 8, 5, padding=1), torch.nn.BatchNorm1d(8), torch.nn.ReLU())
    for block in (pooled, wide, torch.nn.ReLU()):
        with pytest.raises(ValueError):
            first_block.check_block(block)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_kernel_vs_reference_fixture(golden, name):
    from pcgmix_b200 import first_block
    g = golden(name)
    dev = torch.device("cuda:0")
    block = _reference_block(g, dev)
    x = torch.from_numpy(g["x"]).to(dev)
    block.train()
    out, mean, invstd = first_block.first_conv_block(block, x, return_statistics=True)
    assert out.shape == g["out_train"].shape and out.dtype == torch.float32 and not out.requires_grad
    assert _close(out.cpu().numpy(), g["out_train"])
    assert _close(block[1].running_mean.cpu().numpy(), g["running_mean_after"], 1e-5, 1e-6)
    assert _close(block[1].running_var.cpu().numpy(), g["running_var_after"], 1e-5, 1e-6)
    assert int(block[1].num_batches_tracked.item()) == int(g["num_batches_tracked_after"])
    _, _, _, want_mean, want_invstd = fborc.conv_block_forward(g["x"], g["weight"], g["bias"], g["gamma"], g["beta"],
                                                               None, None, True, float(g["eps"]))
    assert _close(mean.cpu().numpy(), want_mean, 1e-5, 1e-6) and _close(invstd.cpu().numpy(), want_invstd, 1e-5, 1e-6)
    block.eval()
    before = block[1].running_mean.clone(), block[1].running_var.clone()
    out = first_block.first_conv_block(block, x)
    assert _close(out.cpu().numpy(), g["out_eval"])
    assert torch.equal(before[0], block[1].running_mean) and torch.equal(before[1], block[1].running_var)
    assert int(block[1].num_batches_tracked.item()) == 1


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(64, 4, 2500, 64), (7, 4, 2499, 64), (33, 1, 2500, 32), (9, 3, 1000, 96), (1, 4, 4, 2), (2, 2, 1, 3)])
@pytest.mark.parametrize("training", [True, False])
def test_kernel_vs_torch_modules_on_the_device(shape, training):
    """The same modules torch runs on the GPU (cuDNN) — an independent float32 evaluation — on synthetic cycles at the
    reference's batch (64 x 4 x 2500) and on ragged shapes; the host mirror must also leave the module's buffers as
    torch's forward leaves them."""
    from pcgmix_b200 import first_block, synth
    B, C, L, F = shape
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(B * 1000 + L)
    if L >= 700:
        frames = synth.cycle_frames(rng, B, limit=L)
        x = torch.from_numpy(synth.cycle_signals(rng, frames, (C,), L)).to(dev)
    else:
        x = torch.from_numpy(rng.standard_normal((B, C, L)).astype(np.float32)).to(dev)
    torch.manual_seed(B + F)
    ours = torch.nn.Sequential(torch.nn.Conv1d(C, F, 3, padding=1), torch.nn.BatchNorm1d(F), torch.nn.ReLU(inplace=True)).to(dev)
    with torch.no_grad():
        ours[1].weight.uniform_(0.5, 1.5)
        ours[1].bias.normal_(0, 0.3)
        ours[1].running_mean.normal_(0, 0.2)
        ours[1].running_var.uniform_(0.5, 2.0)
    import copy
    theirs = copy.deepcopy(ours)
    ours.train(training)
    theirs.train(training)
    with torch.no_grad():
        want = theirs(x.clone())
    got = first_block.first_conv_block(ours, x)
    assert _close(got.cpu().numpy(), want.cpu().numpy(), 5e-5, 5e-5)
    assert _close(ours[1].running_mean.cpu().numpy(), theirs[1].running_mean.cpu().numpy(), 1e-5, 1e-6)
    assert _close(ours[1].running_var.cpu().numpy(), theirs[1].running_var.cpu().numpy(), 1e-5, 1e-6)
    assert int(ours[1].num_batches_tracked.item()) == int(theirs[1].num_batches_tracked.item())


@pytest.mark.gpu
def test_cumulative_average_and_modules_without_statistics(golden):
    """``momentum=None`` (cumulative moving average) and ``track_running_stats=False`` follow torch's rules."""
    from pcgmix_b200 import first_block
    import copy
    dev = torch.device("cuda:0")
    g = golden("first_block_c1_f16")
    x = torch.from_numpy(g["x"]).to(dev)
    for kwargs in (dict(momentum=None), dict(track_running_stats=False), dict(affine=False)):
        torch.manual_seed(3)
        ours = torch.nn.Sequential(torch.nn.Conv1d(1, 16, 3, padding=1), torch.nn.BatchNorm1d(16, **kwargs), torch.nn.ReLU()).to(dev)
        theirs = copy.deepcopy(ours)
        for _ in range(3):                                  # three training steps, then evaluation
            with torch.no_grad():
                want = theirs(x)
            got = first_block.first_conv_block(ours, x)
            assert _close(got.cpu().numpy(), want.cpu().numpy(), 5e-5, 5e-5), kwargs
        if ours[1].running_mean is not None:
            assert _close(ours[1].running_mean.cpu().numpy(), theirs[1].running_mean.cpu().numpy(), 1e-5, 1e-6), kwargs
            assert _close(ours[1].running_var.cpu().numpy(), theirs[1].running_var.cpu().numpy(), 1e-5, 1e-6), kwargs
        ours.eval(), theirs.eval()
        with torch.no_grad():
            want = theirs(x)
        assert _close(first_block.first_conv_block(ours, x).cpu().numpy(), want.cpu().numpy(), 5e-5, 5e-5), kwargs


@pytest.mark.gpu
def test_nan_propagates_like_torch(golden):
    from pcgmix_b200 import first_block
    dev = torch.device("cuda:0")
    g = golden("first_block_c2_f7")
    block = _reference_block(g, dev).eval()
    x = torch.from_numpy(g["x"]).to(dev)
    x[1, 0, 10] = float("nan")
    with torch.no_grad():
        want = block(x.clone()).cpu().numpy()
    got = first_block.first_conv_block(block, x).cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(want)) and np.isnan(got).sum() == 3 * 7
    assert _close(np.nan_to_num(got), np.nan_to_num(want), 5e-5, 5e-5)
    block.train()                                           # batch statistics of a batch with a NaN are NaN everywhere
    assert np.isnan(first_block.first_conv_block(block, x).cpu().numpy()).all()


@pytest.mark.gpu
def test_c_abi_argument_errors(golden):
    from pcgmix_b200 import native
    dev = torch.device("cuda:0")
    x = torch.zeros(2, 4, 64, device=dev)
    w = torch.zeros(8, 4, 3, device=dev)
    out = torch.empty(2, 8, 64, device=dev)
    ws = torch.empty(native.first_conv_block_workspace(4, 8), dtype=torch.uint8, device=dev)
    with pytest.raises(RuntimeError, match="running statistics"):       # evaluation mode needs them
        native.first_conv_block(x, w, None, None, None, None, None, out, ws, False, 1e-5, 0.1)
    with pytest.raises(ValueError):
        native.first_conv_block_workspace(5, 8)
    with pytest.raises(ValueError):
        native.first_conv_block(x, w, None, None, None, None, None, torch.empty(2, 8, 63, device=dev), ws, True, 1e-5, 0.1)
    flat = torch.zeros(2 * 8 * 64 + 2 * 4 * 64, device=dev)
    with pytest.raises(RuntimeError, match="overlap"):
        native.first_conv_block(flat[:512].view(2, 4, 64), w, None, None, None, None, None, flat[256:256 + 1024].view(2, 8, 64),
                                ws, True, 1e-5, 0.1)
    native.first_conv_block(x[:0], w, None, None, None, None, None, out[:0], ws, True, 1e-5, 0.1)     # empty batch: nothing to do


@pytest.mark.gpu
def test_full_size_batch_properties():
    """BASELINE's batch (4096 x 4 x 2500 -> 4096 x 64 x 2500, 2.6 GB): per-filter statistics of the training-mode output
    (mean = beta, variance = gamma^2 where the ReLU is undone by construction: beta large) and agreement of 64 sampled
    cycles with torch's modules."""
    from pcgmix_b200 import first_block, synth
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(99)
    B, C, L, F = 4096, 4, 2500, 64
    frames = synth.cycle_frames(rng, 256, limit=L)
    base = torch.from_numpy(synth.cycle_signals(rng, frames, (C,), L)).to(dev)
    x = base.repeat(B // 256, 1, 1) * torch.linspace(0.5, 1.5, B, device=dev)[:, None, None]
    torch.manual_seed(5)
    block = torch.nn.Sequential(torch.nn.Conv1d(C, F, 3, padding=1), torch.nn.BatchNorm1d(F), torch.nn.ReLU(inplace=True)).to(dev).train()
    with torch.no_grad():
        block[1].weight.uniform_(0.5, 1.5)
        block[1].bias.fill_(40.0)                           # far above zero: the ReLU passes everything
    out, mean, invstd = first_block.first_conv_block(block, x, return_statistics=True)
    m = out.mean(dim=(0, 2), dtype=torch.float64)
    v = torch.stack([((out[:, f].double() - m[f]) ** 2).mean() for f in range(F)])
    assert torch.allclose(m, torch.full_like(m, 40.0), atol=2e-4)
    assert torch.allclose(v.sqrt(), block[1].weight.double().abs(), rtol=2e-4)
    sel = torch.from_numpy(rng.choice(B, 64, replace=False)).to(dev)
    with torch.no_grad():
        z = torch.nn.functional.conv1d(x[sel], block[0].weight, block[0].bias, padding=1)
        want = torch.relu((z - mean[None, :, None]) * invstd[None, :, None] * block[1].weight[None, :, None] + 40.0)
    assert _close(out[sel].cpu().numpy(), want.cpu().numpy(), 2e-5, 2e-4)
