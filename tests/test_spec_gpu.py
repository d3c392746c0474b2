"""Parity of the spectrogram path (``augmentations2d.augment`` -> ``pcgmix_mix2d``) with the
fixtures produced by the unmodified reference and with the CPU oracle.  The blend is the same
three separately rounded fp32 operations as the reference, so everything here is bit-exact."""
import numpy as np
import pytest
import torch

from oracle import pcgmix_oracle as orc

pytestmark = pytest.mark.gpu


class _Args:
    def __init__(self, method, batch):
        self.method, self.batch_size, self.sample_rate, self.num_classes = method, batch, 1000, 2


class _Step:
    def __init__(self, count):
        self.count = count


def _run_2d(method, step, data, labels, frames):
    from pcgmix_b200 import augmentations2d
    dev = torch.device("cuda:0")
    ohe = torch.nn.functional.one_hot(torch.from_numpy(np.asarray(labels)), 2).to(dev)
    d = torch.from_numpy(data).to(dev)
    out, tgt, mix, cut = augmentations2d.augment(_Args(method, data.shape[0]), d, ohe, torch.from_numpy(frames),
                                                 ["a"] * data.shape[0], _Step(step), None, dev, None)
    torch.cuda.synchronize()
    assert cut is None and tgt is ohe
    return out, mix, d


SPEC_CASES = ["spec_pcgmix_square", "spec_timemask_square", "spec_timemask_default", "spec_freqmask_square",
              "spec_cutout_square", "spec_pcgmix_nonsquare"]


@pytest.mark.parametrize("name", SPEC_CASES)
def test_spectrogram_fixtures_bit_exact(golden, name):
    g = golden(name)
    out, mix, d_in = _run_2d(str(g["method"]), int(g["step"]), g["data"], g["labels"], g["frames"])
    assert np.array_equal(mix, g["mix"])
    assert out.shape == d_in.shape and out.data_ptr() != d_in.data_ptr()
    assert np.array_equal(out.cpu().numpy().view(np.uint32), g["out"].view(np.uint32))
    assert np.array_equal(d_in.cpu().numpy(), g["data"])


@pytest.mark.parametrize("shape", [(3, 1, 8, 12), (5, 1, 128, 128), (6, 2, 16, 250), (4, 1, 64, 250), (2, 3, 7, 33),
                                   (3, 1, 4, 1024)])
@pytest.mark.parametrize("method", ["durratiomixup", "durmixtimemask(0.3)", "durmixfreqmask(0.4)", "durmixcutout(0.5,0.5)"])
def test_random_spectrogram_shapes_vs_oracle(shape, method):
    from pcgmix_b200 import synth
    b, ch, f, t = shape
    rng = np.random.default_rng(b * 100 + t)
    frames = synth.spectrogram_frames(rng, b, t)
    data = synth.cycle_signals(rng, frames, (ch, f), t)
    labels = rng.integers(0, 2, b)
    step = 23
    out, mix, _ = _run_2d(method, step, data, labels, frames)
    want, want_mix, _ = orc.augment_2d(method, torch.from_numpy(data.copy()), labels, torch.from_numpy(frames), step)
    assert np.array_equal(mix, want_mix)
    assert np.array_equal(out.cpu().numpy().view(np.uint32), want.numpy().view(np.uint32))


def test_config3_full_size_vectorised_oracle():
    """BASELINE config 3: 1024 x (1 x 64 x 250) — the shape the reference dispatcher cannot run."""
    from pcgmix_b200 import synth
    rng = np.random.default_rng(synth.BENCH_SEED + 3)
    b, f, t = 1024, 64, 250
    frames = synth.spectrogram_frames(rng, b, t)
    data = synth.cycle_signals(rng, frames, (1, f), t)
    labels = rng.integers(0, 2, b)
    out, mix, _ = _run_2d("durratiomixup", 4, data, labels, frames)
    lam32 = orc.lambda_as_float32(orc.draw_lambda(1, 4))
    want = orc.mix_batch_vectorised(data, frames, mix, lam32)
    assert sorted(mix.tolist()) == list(range(b)) and np.array_equal(labels[mix], labels)
    assert np.array_equal(out.cpu().numpy().view(np.uint32), want.view(np.uint32))


def test_2d_gate_and_passthrough():
    from pcgmix_b200 import augmentations2d
    d = torch.zeros(2, 1, 4, 8, device="cuda:0")
    out, tgt, mix, cut = augmentations2d.augment(_Args("unknown", 2), d, None, None, None, _Step(0), None, "cuda:0", None)
    assert out is d and mix == [] and cut is None
