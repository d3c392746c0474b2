"""Parity of the CUDA path (through the public ``augment`` and the C ABI) with the CPU oracle and
with the fixtures produced by the unmodified reference.

Tolerances (BASELINE.json north_star): pairing / indices bit-exact; PCGmix (mix only) bit-exact
— the kernel performs the same three separately rounded fp32 operations as the reference;
PCGmix+ within 1e-5 relative of the reference's float64 spline (measured: ~1e-7, >99.99 %
of samples bit-equal)."""
import numpy as np
import pytest
import torch

from oracle import pcgmix_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["pipeline", "pipeline-f32", "direct"])
def kernel_choice(request):
    """Every parity test runs three times: through the persistent TMA-pipelined kernel (used for rows of
    >= 1024 samples) with the PCGmix+ factor evaluated in float64 (bit-faithful to the reference's float64
    spline) and in float32 (the library default: <= 1e-5 relative), and with the pipelined kernel switched off,
    i.e. through the direct-load kernel only (always float64)."""
    from pcgmix_b200 import native
    native.set_tuning(use_pipeline=request.param != "direct")
    native.set_spline_precision("float32" if request.param.endswith("f32") else "float64")
    yield request.param
    native.set_tuning(use_pipeline=True)
    native.set_spline_precision("float32")


def _bit_faithful_warp():
    """True when PCGmix+ outputs are expected to equal the float64 reference computation sample for sample."""
    from pcgmix_b200 import native
    return native.spline_precision() == "float64"

REL_TOL = 1e-5


class _Args:
    def __init__(self, method, batch):
        self.method, self.batch_size, self.sample_rate, self.num_classes = method, batch, 1000, 2


class _Step:
    def __init__(self, count):
        self.count = count


def _rel_err(got, want):
    denom = np.maximum(np.abs(want), np.finfo(np.float32).tiny)
    return float(np.max(np.abs(got.astype(np.float64) - want.astype(np.float64)) / denom))


def _run_1d(method, step, data, labels, frames, wav=None):
    from pcgmix_b200 import augmentations
    dev = torch.device("cuda:0")
    ohe = torch.nn.functional.one_hot(torch.from_numpy(np.asarray(labels)), int(max(2, labels.max() + 1))).to(dev)
    d = torch.from_numpy(data).to(dev)
    out, tgt, mix, cut = augmentations.augment(_Args(method, data.shape[0]), d, ohe, torch.from_numpy(frames),
                                               wav or ["a0001"] * data.shape[0], _Step(step), None, dev, None)
    torch.cuda.synchronize()
    assert cut is None
    return out, tgt, mix, d


GOLDEN_MIX = ["pcgmix_c4_l2500", "pcgmix_alpha2_prob", "pcgmix_mixall", "pcgmix_rand"]
GOLDEN_WARP = ["pcgmixplus_c4_l2500", "pcgmixplus_default_c2_l800", "pcgmixplus_alpha_k2_oddlen",
               "pcgmixplus_k7_c3", "pcgmixplus_mixall", "pcgmixplus_rand"]


@pytest.mark.parametrize("name", GOLDEN_MIX)
def test_pcgmix_bit_exact_vs_reference_fixture(golden, name):
    g = golden(name)
    out, tgt, mix, d_in = _run_1d(str(g["method"]), int(g["step"]), g["data"], g["labels"], g["frames"])
    assert np.array_equal(mix, g["mix"]) and mix.dtype == np.int64
    got = out.cpu().numpy()
    assert out.data_ptr() != d_in.data_ptr() and out.dtype == torch.float32 and out.is_contiguous()
    assert np.array_equal(got.view(np.uint32), g["out"].view(np.uint32))
    assert np.array_equal(d_in.cpu().numpy(), g["data"]), "input batch was modified"
    assert np.array_equal(tgt.cpu().numpy().astype(np.float32), g["target"].astype(np.float32))


@pytest.mark.parametrize("name", GOLDEN_WARP)
def test_pcgmix_plus_vs_reference_fixture(golden, name):
    g = golden(name)
    out, tgt, mix, _ = _run_1d(str(g["method"]), int(g["step"]), g["data"], g["labels"], g["frames"])
    assert np.array_equal(mix, g["mix"])
    got = out.cpu().numpy()
    assert _rel_err(got, g["out"]) <= REL_TOL
    if _bit_faithful_warp():
        assert np.mean(got == g["out"]) > 0.999, "fp64 spline should reproduce almost every sample bit-for-bit"
    assert np.array_equal(tgt.cpu().numpy().astype(np.float32), g["target"].astype(np.float32))


@pytest.mark.parametrize("tag", ["samepcg", "samedataset"])
def test_pairing_modifiers(golden, tag):
    g = golden(f"pcgmix_{tag}")
    out, _, mix, _ = _run_1d(str(g["method"]), int(g["step"]), g["data"], g["labels"], g["frames"],
                             wav=[str(w) for w in g["wav"]])
    assert np.array_equal(mix, g["mix"])
    assert np.array_equal(out.cpu().numpy().view(np.uint32), g["out"].view(np.uint32))


def test_gate_failure_returns_same_objects(golden):
    from pcgmix_b200 import augmentations
    g = golden("pcgmix_gate_fail")
    dev = torch.device("cuda:0")
    d = torch.from_numpy(g["data"]).to(dev)
    ohe = torch.nn.functional.one_hot(torch.from_numpy(g["labels"]), 2).to(dev)
    out, tgt, mix, cut = augmentations.augment(_Args(str(g["method"]), 4), d, ohe, torch.from_numpy(g["frames"]),
                                               ["a"] * 4, _Step(int(g["step"])), None, dev, None)
    assert out is d and tgt is ohe and mix == [] and cut is None


def test_unknown_method_passthrough():
    from pcgmix_b200 import augmentations
    d = torch.zeros(2, 1, 8, device="cuda:0")
    out, tgt, mix, cut = augmentations.augment(_Args("nothing", 2), d, None, None, None, _Step(0), None, "cuda:0", None)
    assert out is d and mix == [] and cut is None


def test_pairs_edge_cases_through_c_abi(golden):
    """Every (i, j) pair of the edge-case cycles, one kernel launch: cycle k = (i, j) mixes row i
    with partner row n*n + j."""
    from pcgmix_b200 import native
    g = golden("pairs_1d_edge")
    x, fr = g["data"], g["frames"]
    n, c, length = x.shape
    data = np.concatenate([np.repeat(x, n, axis=0), x], axis=0)                       # (n*n + n, c, L)
    frames = np.concatenate([np.repeat(fr, n, axis=0), fr], axis=0).astype(np.int32)
    mix = np.concatenate([n * n + np.tile(np.arange(n), n), np.arange(n * n, n * n + n)]).astype(np.int32)
    dev = torch.device("cuda:0")
    d = torch.from_numpy(data).to(dev)
    out = torch.empty_like(d)
    lam = np.float32(g["lam"])
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    native.mix1d(d, out, torch.from_numpy(frames).to(dev), torch.from_numpy(mix).to(dev), lam,
                 np.float32(1) - lam, err_flag=err)
    got = out.cpu().numpy()[: n * n].reshape(n, n, c, length)
    assert int(err.item()) == 0
    assert np.array_equal(got.view(np.uint32), g["out"].view(np.uint32))


@pytest.mark.parametrize("shape", [(1, 1, 7), (3, 2, 33), (5, 4, 2500), (9, 3, 1001), (4, 1, 4400), (6, 5, 250)])
@pytest.mark.parametrize("method", ["durratiomixup", "durmixmagwarp(0.2,4)", "durmixmagwarp(0.1,12)", "durmixmagwarp(0.3,0)",
                                    "(rand)durratiomixup", "(rand)durmixmagwarp(0.2,4)"])
def test_random_shapes_vs_oracle(shape, method):
    from pcgmix_b200 import synth
    b, c, length = shape
    if "magwarp" in method and length < 16 and "12" in method:
        pytest.skip("more knots than samples: SciPy itself rejects duplicate abscissae")
    rng = np.random.default_rng(b * 1000 + length)
    frames = synth.cycle_frames(rng, b, limit=length)
    data = synth.cycle_signals(rng, frames, (c,), length)
    labels = rng.integers(0, 2, b)
    step = 17
    out, _, mix, _ = _run_1d(method, step, data, labels, frames)
    want, want_mix, _, _ = orc.augment_1d(method, data.copy(), labels, frames, step)
    assert np.array_equal(mix, want_mix)
    got = out.cpu().numpy()
    if "magwarp" in method:
        assert _rel_err(got, want) <= REL_TOL
    else:
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("method", ["durratiomixup", "durmixmagwarp(0.2,4)"])
def test_windows_larger_than_the_staging_buffer(method):
    """Cycles that blend (almost) the whole row: the pipelined kernel cannot stage the partner
    windows in shared memory and must fall back to reading them from global memory."""
    rng = np.random.default_rng(77)
    b, c, length = 12, 2, 2400
    frames = np.tile(np.array([[0, 600, 1200, 1800, 2400]]), (b, 1))
    frames[1] = [0, 599, 1201, 1795, 2399]
    frames[5] = [0, 10, 20, 30, 2400]
    data = rng.standard_normal((b, c, length)).astype(np.float32)
    labels = np.zeros(b, dtype=np.int64)
    out, _, mix, _ = _run_1d(method, 21, data, labels, frames)
    want, want_mix, _, _ = orc.augment_1d(method, data.copy(), labels, frames, 21)
    assert np.array_equal(mix, want_mix)
    got = out.cpu().numpy()
    if "magwarp" in method:
        assert _rel_err(got, want) <= REL_TOL
    else:
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("length", [4096, 6000, 10000])
def test_long_rows_are_sliced(length):
    """Rows longer than one pipeline slice (e.g. BASELINE config 1's 5 s @ 2 kHz = 10 000 samples)."""
    from pcgmix_b200 import synth
    rng = np.random.default_rng(length)
    b, c = 5, 2
    frames = synth.cycle_frames(rng, b, fs=2000 if length > 4096 else 1000, limit=length)
    data = synth.cycle_signals(rng, frames, (c,), length)
    labels = rng.integers(0, 2, b)
    for method in ("durratiomixup", "durmixmagwarp(0.2,4)"):
        out, _, mix, _ = _run_1d(method, 3, data, labels, frames)
        want, _, _, _ = orc.augment_1d(method, data.copy(), labels, frames, 3)
        got = out.cpu().numpy()
        if "magwarp" in method:
            assert _rel_err(got, want) <= REL_TOL
        else:
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_misaligned_base_pointer_uses_scalar_path():
    """A batch whose storage does not start on a 16-byte boundary must still be exact."""
    from pcgmix_b200 import native, synth
    rng = np.random.default_rng(3)
    b, c, length = 6, 2, 404
    frames = synth.cycle_frames(rng, b, limit=length)
    data = synth.cycle_signals(rng, frames, (c,), length)
    mix = rng.permutation(b).astype(np.int32)
    dev = torch.device("cuda:0")
    backing = torch.zeros(b * c * length + 1, device=dev)
    d = backing[1:].view(b, c, length)
    d.copy_(torch.from_numpy(data))
    assert d.data_ptr() % 16 != 0
    out = torch.empty(b, c, length, device=dev)
    lam = np.float32(0.37)
    native.mix1d(d, out, torch.from_numpy(frames.astype(np.int32)).to(dev), torch.from_numpy(mix).to(dev), lam,
                 np.float32(1) - lam)
    want = orc.mix_batch(data, frames, mix, lam)
    assert np.array_equal(out.cpu().numpy().view(np.uint32), want.view(np.uint32))


def test_lambda_one_and_special_values():
    """lambda = 1: samples equal x1 wherever the partner is finite; NaN/Inf/-0 propagate like IEEE."""
    from pcgmix_b200 import native
    dev = torch.device("cuda:0")
    x = np.zeros((2, 1, 16), np.float32)
    x[0, 0, :] = np.arange(16)
    x[0, 0, 3] = -0.0
    x[1, 0, :] = 5
    x[1, 0, 2] = np.inf
    x[1, 0, 5] = np.nan
    frames = np.array([[0, 4, 8, 12, 16], [0, 4, 8, 12, 16]], np.int32)
    mix = np.array([1, 0], np.int32)
    for lam in (np.float32(1.0), np.float32(0.25)):
        out = torch.empty(2, 1, 16, device=dev)
        native.mix1d(torch.from_numpy(x).to(dev), out, torch.from_numpy(frames).to(dev), torch.from_numpy(mix).to(dev),
                     lam, np.float32(1) - lam)
        want = orc.mix_batch(x, frames, mix, lam)
        got = out.cpu().numpy()
        # NaN sign/payload is not specified by IEEE 754 and differs between x86 and the GPU:
        # NaNs must appear in the same places, everything else must be bit-equal (incl. -0.0)
        assert np.array_equal(np.isnan(got), np.isnan(want))
        ok = ~np.isnan(want)
        assert np.array_equal(got.view(np.uint32)[ok], want.view(np.uint32)[ok])


def test_invalid_frames_are_rejected_on_host_and_flagged_on_device():
    from pcgmix_b200 import augmentations, draws, native
    dev = torch.device("cuda:0")
    d = torch.randn(2, 1, 32, device=dev)
    ohe = torch.nn.functional.one_hot(torch.tensor([0, 0]), 2).to(dev)
    bad = torch.tensor([[0, 10, 5, 20, 30], [0, 5, 10, 20, 30]])
    with pytest.raises(ValueError):
        augmentations.augment(_Args("durratiomixup", 2), d, ohe, bad, ["a"] * 2, _Step(1), None, dev, None)
    # offsets past the row end are legal when the clamped windows agree (the reference's slices clamp) ...
    too_long = torch.tensor([[0, 5, 10, 20, 40], [0, 5, 10, 20, 30]])
    out, _, mix, _ = augmentations.augment(_Args("durratiomixup", 2), d, ohe, too_long, ["a"] * 2, _Step(1), None, dev, None)
    want = orc.mix_batch(d.cpu().numpy(), too_long.numpy(), mix, draws.lambda_pair_fp32(draws.draw_lambda(1, 1))[0])
    assert np.array_equal(out.cpu().numpy().view(np.uint32), want.view(np.uint32))
    # ... and a shape mismatch, as in the reference, when they do not
    clash = torch.tensor([[0, 5, 10, 20, 40], [0, 5, 10, 25, 45]])
    with pytest.raises(RuntimeError):
        augmentations.augment(_Args("durratiomixup", 2), d, ohe, clash, ["a"] * 2, _Step(0), None, dev, None)
    # device-resident frames cannot be validated on the host: the kernel copies the cycle and flags it
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    out = torch.empty_like(d)
    native.mix1d(d, out, bad.to(torch.int32).to(dev), torch.tensor([1, 7], dtype=torch.int32, device=dev),
                 0.5, 0.5, err_flag=err)
    torch.cuda.synchronize()
    assert int(err.item()) == (native.ERR_BAD_FRAMES | native.ERR_BAD_PARTNER)
    assert torch.equal(out, d)


def test_cpu_tensor_is_refused():
    from pcgmix_b200 import augmentations
    d = torch.randn(2, 1, 32)
    ohe = torch.nn.functional.one_hot(torch.tensor([0, 0]), 2)
    fr = torch.tensor([[0, 5, 10, 20, 30]] * 2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        augmentations.augment(_Args("durratiomixup", 2), d, ohe, fr, ["a"] * 2, _Step(1), None, "cpu", None)


def test_order_does_not_change_result():
    from pcgmix_b200 import draws, native, synth
    rng = np.random.default_rng(11)
    b, c, length = 64, 4, 2500
    frames = synth.cycle_frames(rng, b, limit=length)
    data = synth.cycle_signals(rng, frames, (c,), length)
    mix = rng.permutation(b)
    dev = torch.device("cuda:0")
    d = torch.from_numpy(data).to(dev)
    f = torch.from_numpy(frames.astype(np.int32)).to(dev)
    m = torch.from_numpy(mix.astype(np.int32)).to(dev)
    o1, o2 = torch.empty_like(d), torch.empty_like(d)
    native.mix1d(d, o1, f, m, 0.3, 0.7)
    order = draws.processing_order(mix)
    assert sorted(order.tolist()) == list(range(b))
    native.mix1d(d, o2, f, m, 0.3, 0.7, order=torch.from_numpy(order).to(dev))
    assert torch.equal(o1, o2)


def test_large_batch_properties_full_size():
    """BASELINE config 2 size (B=4096, C=4, L=2500): checked through size-independent properties
    and a vectorised oracle on a slice."""
    from pcgmix_b200 import augmentations, synth
    rng = np.random.default_rng(synth.BENCH_SEED)
    b, c, length = 4096, 4, 2500
    frames = synth.cycle_frames(rng, b, limit=length)
    data = synth.cycle_signals(rng, frames, (c,), length)
    labels = rng.integers(0, 2, b)
    out, _, mix, d_in = _run_1d("durratiomixup", 5, data, labels, frames)
    got = out.cpu().numpy()
    # pairing is a within-class permutation
    assert sorted(mix.tolist()) == list(range(b)) and np.array_equal(labels[mix], labels)
    # samples outside every blended window are copies of the input; padding stays zero
    lam32 = orc.lambda_as_float32(orc.draw_lambda(1, 5))
    want = orc.mix_batch_vectorised(data, frames, mix, lam32)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    t = np.arange(length)[None, None, :]
    assert not got[np.broadcast_to(t >= frames[:, 4][:, None, None], got.shape)].any()
    # linearity in lambda: out(lam) - x1 == (1-lam) * (x2 - x1) only on blended samples -> checksum of
    # unblended region equals the input's
    changed = got != data
    assert changed.sum() <= c * synth.mixed_samples(frames, mix)


def test_pcgmix_plus_large_batch_sampled():
    from pcgmix_b200 import synth
    rng = np.random.default_rng(synth.BENCH_SEED + 1)
    b, c, length = 2048, 4, 2500
    frames = synth.cycle_frames(rng, b, limit=length)
    data = synth.cycle_signals(rng, frames, (c,), length)
    labels = rng.integers(0, 2, b)
    step = 9
    out, _, mix, _ = _run_1d("durmixmagwarp(0.2,4)", step, data, labels, frames)
    got = out.cpu().numpy()
    lam32 = orc.lambda_as_float32(orc.draw_lambda(1, step))
    knots = orc.draw_knots(b, 4, c, 0.2)
    mixed = orc.mix_batch_vectorised(data, frames, mix, lam32)
    sample = rng.choice(b, 64, replace=False)
    curves = orc.warp_curves(length, knots[sample])
    want = (mixed[sample].astype(np.float64) * curves).astype(np.float32)
    assert _rel_err(got[sample], want) <= REL_TOL


def test_pcgmix_plus_benchmark_size_sampled():
    """bench.py's exact workload (BASELINE config 2: 4096 cycles x 4 x 2500, pairing-chain order, through the
    prepared-launch path the benchmark uses) against the oracle on 96 sampled cycles."""
    from pcgmix_b200 import augmentations, draws, staging, synth
    rng = np.random.default_rng(synth.BENCH_SEED)
    b, c, length, step = 4096, 4, 2500, 23
    frames = synth.cycle_frames(rng, b, fs=1000, limit=length)
    data = synth.cycle_signals(rng, frames, (c,), length)
    labels = rng.integers(0, 2, b)
    dev = torch.device("cuda:0")
    mix = draws.same_label_pairing(labels, step)
    lam, knots = draws.lambda_and_knots(1.0, step, b, 4, c, 0.2)
    lam32, oml = draws.lambda_pair_fp32(lam)
    up = staging.upload([frames.astype(np.int32), mix.astype(np.int32), draws.processing_order(mix), knots], dev)
    x = torch.from_numpy(data).to(dev)
    out = torch.empty_like(x)
    augmentations.prepare_on_device(x, up[0], up[1], lam32, oml, out, up[3], 4, order_dev=up[2]).launch()
    torch.cuda.synchronize()
    assert np.array_equal(mix, orc.same_label_mix_indices(labels, step)) and lam == orc.draw_lambda(1, step)
    assert np.array_equal(knots, orc.draw_knots(b, 4, c, 0.2))                   # (global stream: right after draw_lambda)
    sample = np.sort(rng.choice(b, 96, replace=False))
    mixed = np.stack([orc.mix_pair(data[i], data[mix[i]], frames[i], frames[mix[i]], lam32) for i in sample])
    want = (mixed.astype(np.float64) * orc.warp_curves(length, knots[sample])).astype(np.float32)
    got = out[torch.from_numpy(sample).to(dev)].cpu().numpy()
    assert _rel_err(got, want) <= REL_TOL
    if _bit_faithful_warp():
        assert np.mean(got == want) > 0.999


@pytest.mark.parametrize("shape", [(7, 4, 2500), (3, 2, 10000), (5, 3, 1001), (4, 1, 64 * 250)])
@pytest.mark.parametrize("magwarp", [False, True])
def test_output_guard_bands_stay_intact(shape, magwarp):
    """compute-sanitizer is closed on this GPU pool, so out-of-bounds WRITES are looked for by hand:
    the output lives inside a larger buffer filled with a sentinel that must survive the launch."""
    from pcgmix_b200 import native, spline, synth
    b, c, length = shape
    rng = np.random.default_rng(sum(shape))
    frames = synth.cycle_frames(rng, b, limit=length)
    data = synth.cycle_signals(rng, frames, (c,), length)
    mix = rng.permutation(b).astype(np.int32)
    dev = torch.device("cuda:0")
    n = b * c * length
    guard = 4096
    backing = torch.full((n + 2 * guard,), 12345.0, device=dev)
    out = backing[guard:guard + n].view(b, c, length)
    d = torch.from_numpy(data).to(dev)
    f = torch.from_numpy(frames.astype(np.int32)).to(dev)
    m = torch.from_numpy(mix).to(dev)
    lam = np.float32(0.6)
    if magwarp:
        knots = rng.normal(1, 0.2, (b, 6, c))
        pos, mat = spline.magwarp_tables(length, 4)
        native.mix1d_magwarp(d, out, f, m, lam, np.float32(1) - lam, torch.from_numpy(knots).to(dev),
                             torch.from_numpy(np.array(mat)).to(dev), torch.from_numpy(np.array(pos)).to(dev), 4)
    else:
        native.mix1d(d, out, f, m, lam, np.float32(1) - lam)
    torch.cuda.synchronize()
    assert bool((backing[:guard] == 12345.0).all()) and bool((backing[guard + n:] == 12345.0).all())
    assert not bool((out == 12345.0).any())
    if not magwarp:
        assert np.array_equal(out.cpu().numpy().view(np.uint32), orc.mix_batch(data, frames, mix, lam).view(np.uint32))


@pytest.mark.parametrize("shape", [(1, 1, 1024), (2, 1, 1028), (13, 3, 3584), (5, 2, 3588), (3, 1, 7168), (37, 2, 1536), (2, 70, 1024)])
@pytest.mark.parametrize("method", ["durratiomixup", "durmixmagwarp(0.2,4)", "durmixmagwarp(0.05,30)", "durmixmagwarp(0.2,0)"])
def test_pipeline_kernel_shape_corners(shape, method):
    """Smallest / largest single slice, first multi-slice length, one cycle, one channel, many rows,
    and the extreme knot counts (0 and PCGMIX_MAX_KNOT: 124 coefficients per row, 32 KB matrix)."""
    from pcgmix_b200 import synth
    b, c, length = shape
    rng = np.random.default_rng(b + 7 * c + length)
    fs = 2000 if length > 3000 else 1000
    frames = synth.cycle_frames(rng, b, fs=fs, limit=length)
    data = synth.cycle_signals(rng, frames, (c,), length)
    labels = rng.integers(0, 2, b)
    out, _, mix, _ = _run_1d(method, 29, data, labels, frames)
    want, want_mix, _, _ = orc.augment_1d(method, data.copy(), labels, frames, 29)
    assert np.array_equal(mix, want_mix)
    got = out.cpu().numpy()
    if "magwarp" in method:
        assert _rel_err(got, want) <= REL_TOL
    else:
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_launch_overlap_keeps_results_and_respects_dependencies():
    """With launch overlap enabled, independent back-to-back launches may overlap, but a launch that
    consumes the previous launch's output (or overwrites its input) must still see ordinary stream
    order — the library checks the buffers.  Every launch is compared with the oracle."""
    from pcgmix_b200 import native, synth
    rng = np.random.default_rng(55)
    b, c, length = 512, 2, 2500
    dev = torch.device("cuda:0")
    frames = synth.cycle_frames(rng, b, limit=length)
    f = torch.from_numpy(frames.astype(np.int32)).to(dev)
    xs = [synth.cycle_signals(rng, frames, (c,), length) for _ in range(3)]
    mixes = [rng.permutation(b).astype(np.int32) for _ in range(8)]
    lam = np.float32(0.35)
    native.set_launch_overlap(True)
    try:
        # (a) independent launches, rotating buffers
        d = [torch.from_numpy(x).to(dev) for x in xs]
        outs = [torch.empty_like(d[0]) for _ in range(3)]
        for k in range(6):
            native.mix1d(d[k % 3], outs[k % 3], f, torch.from_numpy(mixes[k]).to(dev), lam, np.float32(1) - lam)
            if k >= 3:
                continue
        torch.cuda.synchronize()
        for k in range(3, 6):
            want = orc.mix_batch(xs[k % 3], frames, mixes[k], lam)
            assert np.array_equal(outs[k % 3].cpu().numpy().view(np.uint32), want.view(np.uint32))
        # (b) a dependent chain: out of launch k is the input of launch k+1
        cur = torch.from_numpy(xs[0]).to(dev)
        ref = xs[0]
        bufs = [torch.empty_like(cur) for _ in range(2)]
        for k in range(5):
            native.mix1d(cur, bufs[k % 2], f, torch.from_numpy(mixes[k]).to(dev), lam, np.float32(1) - lam)
            cur = bufs[k % 2]
            ref = orc.mix_batch(ref, frames, mixes[k], lam)
        torch.cuda.synchronize()
        assert np.array_equal(cur.cpu().numpy().view(np.uint32), ref.view(np.uint32))
    finally:
        native.set_launch_overlap(False)


def test_launch_overlap_bookkeeping_is_reset_by_other_launches_of_the_library(kernel_choice):
    """A mix launch may only be made a programmatic dependent of another PCGmix launch whose buffers are disjoint —
    never of the kernel that uploaded its per-step tables (or cut its cycles): every other entry point that launches
    on the stream resets the bookkeeping.  Counted through `pcgmix_overlap_launches`."""
    from pcgmix_b200 import native, staging, synth
    if kernel_choice == "direct":
        pytest.skip("only pipelined launches overlap")
    rng = np.random.default_rng(56)
    b, c, length = 1024, 4, 2500                               # enough items to fill the GPU (overlap needs a full grid)
    dev = torch.device("cuda:0")
    frames = synth.cycle_frames(rng, b, limit=length)
    f = torch.from_numpy(frames.astype(np.int32)).to(dev)
    x = [torch.from_numpy(synth.cycle_signals(rng, frames, (c,), length)).to(dev) for _ in range(3)]
    outs = [torch.empty_like(x[0]) for _ in range(3)]
    mix = torch.from_numpy(rng.permutation(b).astype(np.int32)).to(dev)
    lam = np.float32(0.4)
    native.set_launch_overlap(True)
    try:
        native.mix1d(x[0], outs[0], f, mix, lam, np.float32(1) - lam)
        before = native.overlap_launches()
        native.mix1d(x[1], outs[1], f, mix, lam, np.float32(1) - lam)        # disjoint from the previous launch: overlapped
        assert native.overlap_launches() == before + 1
        up = staging.upload([frames.astype(np.int32)], dev)                  # the library's own table upload on the same stream
        native.mix1d(x[2], outs[2], up[0], mix, lam, np.float32(1) - lam)    # reads what that upload wrote: ordinary stream order
        assert native.overlap_launches() == before + 1
        native.mix1d(x[0], outs[0], f, mix, lam, np.float32(1) - lam)        # and the launch after it may overlap again
        assert native.overlap_launches() == before + 2
        torch.cuda.synchronize()
        want = orc.mix_batch(x[2].cpu().numpy(), frames, mix.cpu().numpy(), lam)
        assert np.array_equal(outs[2].cpu().numpy().view(np.uint32), want.view(np.uint32))
    finally:
        native.set_launch_overlap(False)


def test_prepared_launch_matches_plain_call():
    from pcgmix_b200 import augmentations, draws, native, synth
    rng = np.random.default_rng(91)
    b, c, length = 96, 4, 2500
    dev = torch.device("cuda:0")
    frames = synth.cycle_frames(rng, b, limit=length)
    data = synth.cycle_signals(rng, frames, (c,), length)
    mix = rng.permutation(b).astype(np.int32)
    knots = rng.normal(1, 0.2, (b, 6, c))
    d = torch.from_numpy(data).to(dev)
    f = torch.from_numpy(frames.astype(np.int32)).to(dev)
    m = torch.from_numpy(mix).to(dev)
    kn = torch.from_numpy(knots).to(dev)
    lam = np.float32(0.4)
    for knots_dev in (None, kn):
        o1, o2 = torch.empty_like(d), torch.empty_like(d)
        augmentations.pcgmix_on_device(d, f, m, lam, np.float32(1) - lam, knots_dev, 4, out=o1)
        prep = augmentations.prepare_on_device(d, f, m, lam, np.float32(1) - lam, o2, knots_dev, 4)
        before = native.launch_count
        prep.launch()
        prep.launch()
        torch.cuda.synchronize()
        assert native.launch_count == before + 2
        assert torch.equal(o1, o2)


@pytest.mark.parametrize("length", [1024, 2500, 3584, 7168])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_arbitrary_monotone_offsets_stress(length, seed):
    """Offsets drawn uniformly (sorted) from [0, L]: zero-length states, f[0] > 0, f[4] == L, windows
    cut by slice boundaries of multi-slice rows, windows too large for the staging buffer — against
    the vectorised oracle, for PCGmix (bit-exact) and PCGmix+ (1e-5)."""
    from pcgmix_b200 import native, spline
    rng = np.random.default_rng(1000 * seed + length)
    b, c = 48, 3
    frames = np.sort(rng.integers(0, length + 1, size=(b, 5)), axis=1)
    frames[0] = [0, 0, 0, 0, 0]
    frames[1] = [0, 0, 0, 0, length]
    frames[2] = [length, length, length, length, length]
    frames[3] = [0, length // 4, length // 2, 3 * length // 4, length]
    frames[4] = [3, 3, length // 3 + 1, length // 3 + 1, length - 1]
    data = rng.standard_normal((b, c, length)).astype(np.float32)
    mix = rng.integers(0, b, b).astype(np.int32)            # any mapping, not only permutations
    dev = torch.device("cuda:0")
    d = torch.from_numpy(data).to(dev)
    f = torch.from_numpy(frames.astype(np.int32)).to(dev)
    m = torch.from_numpy(mix).to(dev)
    lam = np.float32(rng.uniform(0.05, 0.95))
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    out = torch.empty_like(d)
    native.mix1d(d, out, f, m, lam, np.float32(1) - lam, err_flag=err)
    want = orc.mix_batch_vectorised(data, frames, mix, lam)
    assert int(err.item()) == 0
    assert np.array_equal(out.cpu().numpy().view(np.uint32), want.view(np.uint32))
    knots = rng.normal(1, 0.2, (b, 6, c))
    pos, mat = spline.magwarp_tables(length, 4)
    native.mix1d_magwarp(d, out, f, m, lam, np.float32(1) - lam, torch.from_numpy(knots).to(dev),
                         torch.from_numpy(np.array(mat)).to(dev), torch.from_numpy(np.array(pos)).to(dev), 4, err_flag=err)
    curves = orc.warp_curves(length, knots)
    want_w = (want.astype(np.float64) * curves).astype(np.float32)
    assert int(err.item()) == 0
    assert _rel_err(out.cpu().numpy(), want_w) <= REL_TOL


@pytest.mark.parametrize("sigma", [0.2, 0.5, 1.5])
def test_zero_samples_keep_the_sign_the_reference_gives_them(sigma):
    """Zero padding (and zero samples inside a cycle) times the warp factor: +0.0 where the factor is
    positive, -0.0 where it is negative, exactly like the reference's float64 product.  With
    sigma = 0.2 nearly every row takes the kernels' "factor certainly positive" shortcut, with
    sigma = 1.5 many factors are negative and the shortcut must not fire."""
    from pcgmix_b200 import synth
    rng = np.random.default_rng(int(sigma * 10))
    b, c, length = 96, 4, 2500
    frames = synth.cycle_frames(rng, b, limit=length)
    data = synth.cycle_signals(rng, frames, (c,), length)
    data[::3, :, 40:300] = 0.0                                # digital silence inside blended windows
    data[1::3, 1, 500:508] = -0.0                             # negative zeros stay negative times a positive factor
    labels = rng.integers(0, 2, b)
    method = f"durmixmagwarp({sigma},4)"
    out, _, mix, _ = _run_1d(method, 13, data, labels, frames)
    want, want_mix, _, _ = orc.augment_1d(method, data, labels, frames, 13)
    got = out.cpu().numpy()
    assert np.array_equal(mix, want_mix)
    assert _rel_err(got, want) <= REL_TOL
    zeros = want == 0
    assert zeros.mean() > 0.4
    assert np.array_equal(np.signbit(got[zeros]), np.signbit(want[zeros]))
    assert np.array_equal(got == 0, zeros)
    if sigma >= 1.5:
        assert np.signbit(want[zeros]).mean() > 0.02          # the case is exercised: negative factors exist


@pytest.mark.parametrize("knot", [0, 4, 12, 30])
@pytest.mark.parametrize("with_order", [False, True])
def test_coefficient_warp_against_the_direct_kernel(kernel_choice, knot, with_order):
    """In the pipelined kernel a dedicated warp turns every item's knots into the cubic coefficients its
    consumers use (up to 124 of them at knot = 30, four per lane).  A launch with many items per CTA, with the
    float64 evaluation, must equal the direct-load kernel's result bit for bit (same fma order); the float32
    evaluation must stay within the tolerance."""
    from pcgmix_b200 import draws, native, spline, synth
    if kernel_choice == "direct":
        pytest.skip("compares the pipelined kernel against the direct-load one itself")
    rng = np.random.default_rng(100 + knot)
    b, c, length = 3000, 4, 2500                      # 12 000 items over 444 CTAs: 27-28 per CTA
    dev = torch.device("cuda:0")
    frames = synth.cycle_frames(rng, b, limit=length)
    data = torch.from_numpy(synth.cycle_signals(rng, frames, (c,), length)).to(dev)
    f = torch.from_numpy(frames.astype(np.int32)).to(dev)
    mix_host = draws.same_label_pairing(rng.integers(0, 2, b), knot)
    mix = torch.from_numpy(mix_host.astype(np.int32)).to(dev)
    order = torch.from_numpy(draws.processing_order(mix_host)).to(dev) if with_order else None
    knots = torch.from_numpy(rng.normal(1.0, 0.2, (b, knot + 2, c))).to(dev)
    pos, mat = spline.magwarp_tables(length, knot)
    pos_d, mat_d = torch.from_numpy(np.array(pos)).to(dev), torch.from_numpy(np.array(mat)).to(dev)
    lam = draws.lambda_pair_fp32(0.61)
    outs = []
    for use_pipeline in (True, False):
        native.set_tuning(use_pipeline=use_pipeline)
        out = torch.empty_like(data)
        native.mix1d_magwarp(data, out, f, mix, lam[0], lam[1], knots, mat_d, pos_d, knot, order=order)
        outs.append(out)
    native.set_tuning(use_pipeline=True)
    torch.cuda.synchronize()
    if _bit_faithful_warp():
        assert torch.equal(outs[0].view(torch.int32), outs[1].view(torch.int32))
    else:
        assert _rel_err(outs[0].cpu().numpy(), outs[1].cpu().numpy()) <= REL_TOL


# ---- cycles whose offsets run past the row end (fixtures from the unmodified reference) ----------------
def _all_pairs_launch(x, fr, lam, two_d=False):
    """cycle k = (i, j) mixes row i with partner row n*n + j, one launch; returns (out[n, n, ...], err)."""
    from pcgmix_b200 import native
    n = x.shape[0]
    data = np.concatenate([np.repeat(x, n, axis=0), x], axis=0)
    frames = np.concatenate([np.repeat(fr, n, axis=0), fr], axis=0).astype(np.int32)
    mix = np.concatenate([n * n + np.tile(np.arange(n), n), np.arange(n * n, n * n + n)]).astype(np.int32)
    dev = torch.device("cuda:0")
    d = torch.from_numpy(data).to(dev)
    out = torch.empty_like(d)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    fn = native.mix2d if two_d else native.mix1d
    fn(d, out, torch.from_numpy(frames).to(dev), torch.from_numpy(mix).to(dev), lam, np.float32(1) - lam, err_flag=err)
    torch.cuda.synchronize()
    return out.cpu().numpy()[: n * n].reshape((n, n) + x.shape[1:]), int(err.item())


@pytest.mark.parametrize("tag", ["1d", "2d"])
def test_long_cycle_pairs_vs_reference_fixture(golden, tag):
    """Where the reference blends, the kernel blends bit-exactly; where it raises a shape mismatch (or
    broadcasts one partner sample over a window), the kernel copies the cycle and raises BAD_FRAMES."""
    from pcgmix_b200 import _common, native
    g = golden(f"long_pairs_{tag}")
    x, fr, ok = g["data"], g["frames"], g["ok"]
    n, length = x.shape[0], x.shape[-1]
    got, err = _all_pairs_launch(x, fr, np.float32(g["lam"]), two_d=tag == "2d")
    assert err == native.ERR_BAD_FRAMES
    blended = 0
    for i in range(n):
        for j in range(n):
            _, _, wd, ws = _common.clamped_windows(fr[i:i + 1], fr[j:j + 1], length)
            if ((wd == ws) | ((wd == 0) & (ws == 1))).all():
                assert ok[i, j]
                assert np.array_equal(got[i, j].view(np.uint32), g["out"][i, j].view(np.uint32)), (i, j)
                blended += 1
            else:
                assert np.array_equal(got[i, j].view(np.uint32), x[i].view(np.uint32)), (i, j)
    assert blended == 23


def test_long_cycle_pairs_in_long_rows():
    """The same offset lists scaled to rows of 1500 samples, so that the pipelined kernel takes them
    (its producer derives the windows itself); expected values from the oracle's slice semantics."""
    from pcgmix_b200 import native
    rng = np.random.default_rng(77)
    base = np.array([[0, 10, 20, 30, 40], [0, 10, 20, 30, 70], [0, 12, 24, 36, 50], [0, 10, 20, 55, 60],
                     [0, 60, 70, 80, 90], [0, 5, 10, 15, 120], [2, 9, 21, 33, 52]], dtype=np.int64)
    fr = base * 30 + np.array([0, 1, 2, 3, 1])              # odd shifts between the pairs
    length, lam = 1500, np.float32(0.41)
    x = rng.standard_normal((fr.shape[0], 2, length)).astype(np.float32)
    got, err = _all_pairs_launch(x, fr, lam)
    n_ok = 0
    for i in range(len(fr)):
        for j in range(len(fr)):
            try:
                want = orc.mix_pair(x[i], x[j], fr[i], fr[j], lam)
                n_ok += 1
            except ValueError:
                want = x[i]
            assert np.array_equal(got[i, j].view(np.uint32), want.view(np.uint32)), (i, j)
    assert 0 < n_ok < len(fr) ** 2 and err == native.ERR_BAD_FRAMES


def test_long_cycle_batch_vs_reference_fixture(golden):
    g = golden("long_batch_1d")
    out, _, mix, _ = _run_1d("durratiomixup", int(g["step"]), g["data"], g["labels"], g["frames"])
    assert np.array_equal(mix, g["mix"])
    assert np.array_equal(out.cpu().numpy().view(np.uint32), g["out_durratiomixup"].view(np.uint32))
    out, _, mix, _ = _run_1d("durmixmagwarp(0.2,4)", int(g["step"]), g["data"], g["labels"], g["frames"])
    assert _rel_err(out.cpu().numpy(), g["out_durmixmagwarp"]) <= REL_TOL
    with pytest.raises(RuntimeError):
        _run_1d("durratiomixup", int(g["step_raises"]), g["data"], g["labels"], g["frames"])


def test_order_entries_out_of_range_are_flagged_not_followed():
    from pcgmix_b200 import native, synth
    rng = np.random.default_rng(8)
    b, c, length = 6, 2, 1200
    frames = synth.cycle_frames(rng, b, limit=length)
    x = synth.cycle_signals(rng, frames, (c,), length)
    dev = torch.device("cuda:0")
    d = torch.from_numpy(x).to(dev)
    guard = torch.full((b + 2, c, length), 3.0, device=dev)
    out = guard[1:b + 1]
    mix = np.array([1, 0, 3, 2, 5, 4], np.int32)
    order = torch.tensor([0, 1, 99, 3, -4, 5], dtype=torch.int32, device=dev)      # slots 2 and 4 name no cycle
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    lam = np.float32(0.3)
    native.mix1d(d, out, torch.from_numpy(frames.astype(np.int32)).to(dev), torch.from_numpy(mix).to(dev), lam,
                 np.float32(1) - lam, order=order, err_flag=err)
    torch.cuda.synchronize()
    assert int(err.item()) & native.ERR_BAD_PARTNER
    want = orc.mix_batch(x, frames, mix, lam)
    assert np.array_equal(out.cpu().numpy().view(np.uint32), want.view(np.uint32))   # the slot's own cycle was processed
    assert torch.all(guard[0] == 3.0) and torch.all(guard[-1] == 3.0)


def test_tables_shorter_than_the_batch_are_refused():
    from pcgmix_b200 import native
    dev = torch.device("cuda:0")
    d = torch.zeros(4, 1, 64, device=dev)
    out = torch.empty_like(d)
    fr = torch.zeros(4, 5, dtype=torch.int32, device=dev)
    with pytest.raises(ValueError):
        native.mix1d(d, out, fr, torch.zeros(3, dtype=torch.int32, device=dev), 0.5, 0.5)
    with pytest.raises(ValueError):
        native.mix1d(d, out, fr[:2], torch.zeros(4, dtype=torch.int32, device=dev), 0.5, 0.5)
    with pytest.raises(ValueError):
        native.mix1d(d, out, fr, torch.zeros(4, dtype=torch.int32, device=dev), 0.5, 0.5,
                     order=torch.zeros(2, dtype=torch.int32, device=dev))
    with pytest.raises(RuntimeError):                                            # overlapping views of one buffer
        buf = torch.zeros(5, 1, 64, device=dev)
        native.mix1d(buf[:4], buf[1:], fr, torch.zeros(4, dtype=torch.int32, device=dev), 0.5, 0.5)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_two_devices_in_one_process():
    """The reference trains single-process multi-GPU (nn.DataParallel, train_model.py:385): the library keeps
    its per-device facts (shared-memory opt-in, SM count, launch history) apart."""
    from pcgmix_b200 import synth
    rng = np.random.default_rng(12)
    b, c, length = 64, 4, 2500
    frames = synth.cycle_frames(rng, b, limit=length)
    x = synth.cycle_signals(rng, frames, (c,), length)
    labels = rng.integers(0, 2, b)
    from pcgmix_b200 import augmentations
    for rep in range(2):
        for index in (0, 1):
            dev = torch.device("cuda", index)
            ohe = torch.nn.functional.one_hot(torch.from_numpy(labels), 2).to(dev)
            for method in ("durratiomixup", "durmixmagwarp(0.2,4)"):
                out, _, mix, _ = augmentations.augment(_Args(method, b), torch.from_numpy(x).to(dev), ohe, torch.from_numpy(frames),
                                                       ["a"] * b, _Step(3 + rep), None, dev, None)
                torch.cuda.synchronize(dev)
                want, want_mix, _, _ = orc.augment_1d(method, x, labels, frames, 3 + rep)
                assert out.device == dev and np.array_equal(mix, want_mix)
                assert _rel_err(out.cpu().numpy(), want) <= REL_TOL
