"""Size-independent properties of the oracle (hypothesis): they hold for any offsets and pairing, so
they also describe what the CUDA path must do at sizes where no fixture exists."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import pcgmix_oracle as orc


@st.composite
def batches(draw):
    b = draw(st.integers(1, 6))
    c = draw(st.integers(1, 3))
    length = draw(st.integers(8, 120))
    cuts = draw(st.lists(st.lists(st.integers(0, length), min_size=5, max_size=5), min_size=b, max_size=b))
    frames = np.sort(np.array(cuts, dtype=np.int64), axis=1)
    seed = draw(st.integers(0, 2 ** 16))
    rng = np.random.default_rng(seed)
    data = rng.standard_normal((b, c, length)).astype(np.float32)
    mix = rng.integers(0, b, b)
    lam = np.float32(draw(st.floats(0, 1, width=32)))
    return data, frames, mix, lam


@settings(max_examples=60, deadline=None)
@given(batches())
def test_loop_and_vectorised_oracle_agree_and_only_windows_change(case):
    data, frames, mix, lam = case
    got = orc.mix_batch(data, frames, mix, lam)
    assert np.array_equal(got.view(np.uint32), orc.mix_batch_vectorised(data, frames, mix, lam).view(np.uint32))
    # samples outside the four blended windows are bit copies of the input
    t = np.arange(data.shape[-1])[None, :]
    inside = np.zeros((data.shape[0], data.shape[-1]), bool)
    f2 = frames[mix]
    for s in range(4):
        n = np.minimum(frames[:, s + 1] - frames[:, s], f2[:, s + 1] - f2[:, s])
        inside |= (t >= frames[:, s:s + 1]) & (t < (frames[:, s] + n)[:, None])
    keep = ~np.broadcast_to(inside[:, None, :], data.shape)
    assert np.array_equal(got[keep].view(np.uint32), data[keep].view(np.uint32))
    # lambda = 1 is the identity, lambda = 0 copies the partner's window
    assert np.array_equal(orc.mix_batch(data, frames, mix, np.float32(1)), data)
    zero = orc.mix_batch(data, frames, mix, np.float32(0))
    for i in range(data.shape[0]):
        for s in range(4):
            n = min(frames[i, s + 1] - frames[i, s], f2[i, s + 1] - f2[i, s])
            a, b = frames[i, s], f2[i, s]
            assert np.array_equal(zero[i, :, a:a + n], data[mix[i], :, b:b + n] + np.float32(0) * data[i, :, a:a + n])


@settings(max_examples=30, deadline=None)
@given(st.integers(2, 300), st.integers(0, 6), st.integers(0, 2 ** 16))
def test_warp_curve_interpolates_the_knots(length, knot, seed):
    if knot + 2 > length:
        return
    rng = np.random.default_rng(seed)
    knots = rng.normal(1, 0.2, (1, knot + 2, 1))
    curve = orc.warp_curves(length, knots)[0, 0]
    pos = np.linspace(0, length - 1.0, knot + 2)
    on_grid = np.isclose(pos, np.round(pos))
    np.testing.assert_allclose(curve[np.round(pos[on_grid]).astype(int)], knots[0, on_grid, 0], rtol=1e-9, atol=1e-12)
