"""Fused cut + pad + PCGmix(+) over resident recordings (``pcgmix_mix1d_resident``) against
(a) the two-step device path it replaces (``pcgmix_cut_cycles`` then ``pcgmix_mix1d[_magwarp]``,
bit for bit), and (b) the CPU oracle (cut + pad of the notebook cells, then the reference's
``augment``): bit-exact for PCGmix, 1e-5 relative for PCGmix+ (north_star's tolerance)."""
import numpy as np
import pytest
import torch

from oracle import pcgmix_oracle as orc
from oracle import segmentation_oracle as seg_orc

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5       # north_star: float outputs within 1e-5 relative of the NumPy reference


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


class _Args:
    def __init__(self, method, batch):
        self.method, self.batch_size, self.sample_rate, self.num_classes = method, batch, 1000, 2


class _Step:
    def __init__(self, count):
        self.count = count


def _resident(seed, n_rec, channels, t_len, length, fs=1000):
    from pcgmix_b200 import resident, synth
    rng = np.random.default_rng(seed)
    states = synth.dense_states(rng, n_rec, t_len, fs)
    signal = rng.standard_normal((n_rec, channels, t_len)).astype(np.float32)
    res = resident.from_dense_states(torch.from_numpy(signal).cuda(), torch.from_numpy(states).cuda(), length)
    return res, signal, states, rng


def _bits(a):
    return a.view(np.uint32)


# sigma = 1.5: many warp factors are negative, padding becomes -0.0 and the "factor certainly positive"
# shortcut of the pipelined kernel must not fire (the comparison is bitwise, so the sign of zero counts)
@pytest.mark.parametrize("method", ["durratiomixup", "durmixmagwarp(0.2,4)", "(alpha=0.4)durmixmagwarp(0.3,7)",
                                    "durmixmagwarp(1.5,4)", "durmixmagwarp(0.2,12)"])
@pytest.mark.parametrize("n_rec,channels,t_len,length", [
    (12, 4, 9001, 2500),     # odd recording length: every row starts at a different 16-byte phase
    (9, 1, 12002, 2500),
    (6, 3, 8003, 1598),      # L not a multiple of 4: scalar variant
    (5, 2, 30000, 4400),     # two slices per row
])
def test_resident_equals_cut_then_mix(method, n_rec, channels, t_len, length):
    from pcgmix_b200 import augmentations, resident
    res, _, _, rng = _resident(n_rec * 7 + channels, n_rec, channels, t_len, length)
    assert res.n_cycles > 8
    batch = 2 * res.n_cycles + 3                                  # rows repeat inside the batch
    ids = rng.integers(0, res.n_cycles, batch)
    labels = rng.integers(0, 2, batch)
    ohe = torch.nn.functional.one_hot(torch.from_numpy(labels), 2).cuda()
    wav = ["a"] * batch
    for step, ids_arg in ((3, ids), (11, torch.from_numpy(ids).cuda())):
        got, _, mix, _ = resident.augment(_Args(method, batch), res, ids_arg, ohe, wav, _Step(step), None, "cuda", None)
        want, _, mix2, _ = augmentations.augment(_Args(method, batch), res.padded(ids), ohe, res.frames_of(ids), wav,
                                                 _Step(step), None, "cuda", None)
        res.check()
        assert np.array_equal(mix, mix2)
        assert got.shape == (batch, channels, length)
        assert np.array_equal(_bits(got.cpu().numpy()), _bits(want.cpu().numpy()))


@pytest.mark.parametrize("method", ["durratiomixup", "durmixmagwarp(0.2,4)"])
def test_resident_against_cpu_oracle(method):
    from pcgmix_b200 import resident
    res, signal, states, rng = _resident(5, 8, 4, 10007, 2500)
    # CPU: the notebook's cycle rule + cut + pad, then the reference's augment
    cyc, frames = [], []
    for r in range(states.shape[0]):
        rel, a0, a1 = seg_orc.cycles_from_dense(states[r])
        for i in range(len(a0)):
            cyc.append(np.stack([seg_orc.cut_and_pad(signal[r, c], a0[i], a1[i], 2500) for c in range(4)]))
            frames.append(rel[i])
    cyc, frames = np.stack(cyc), np.stack(frames)
    assert cyc.shape[0] == res.n_cycles
    ids = rng.permutation(res.n_cycles)[: res.n_cycles - 1]
    labels = rng.integers(0, 2, len(ids))
    want, want_mix, _, _ = orc.augment_1d(method, cyc[ids], labels, frames[ids], 9)
    ohe = torch.nn.functional.one_hot(torch.from_numpy(labels), 2).cuda()
    got, _, mix, _ = resident.augment(_Args(method, len(ids)), res, ids, ohe, ["a"] * len(ids), _Step(9), None, "cuda", None)
    res.check()
    got = got.cpu().numpy()
    assert np.array_equal(mix, want_mix)
    if method == "durratiomixup":
        assert np.array_equal(_bits(got), _bits(want))
    else:
        denom = np.maximum(np.abs(want), np.finfo(np.float32).tiny)
        assert float(np.max(np.abs(got - want) / denom)) <= REL_TOL
        if _bit_faithful_warp():
            assert np.mean(got == want) > 0.999


def test_gate_fail_and_unknown_method_return_the_plain_batch():
    from pcgmix_b200 import resident
    res, _, _, rng = _resident(1, 4, 2, 9000, 2500)
    ids = rng.integers(0, res.n_cycles, 10)
    ohe = torch.nn.functional.one_hot(torch.from_numpy(rng.integers(0, 2, 10)), 2).cuda()
    for method, step in (("durratiomixup+0.0", 1), ("base", 1)):
        out, t, mix, cut = resident.augment(_Args(method, 10), res, ids, ohe, ["a"] * 10, _Step(step), None, "cuda", None)
        assert mix == [] and cut is None and t is ohe
        assert torch.equal(out, res.padded(ids))


def _hand_made(signal, rows, length):
    """A resident set around a hand-written cycle table."""
    from pcgmix_b200 import resident, segmentation
    dev = torch.device("cuda")
    table = segmentation.CycleTable(torch.tensor(rows, dtype=torch.int32, device=dev),
                                    torch.tensor([0, len(rows)], dtype=torch.int32, device=dev),
                                    torch.zeros(1, dtype=torch.int32, device=dev))
    return resident.ResidentCycles(torch.from_numpy(signal).cuda(), table, length, len(rows),
                                   torch.zeros(1, dtype=torch.int32, device=dev))


def _two_step(res, mix, lam, knots=None, knot=0):
    """pcgmix_cut_cycles followed by pcgmix_mix1d(_magwarp) through the direct-load and the pipelined kernel."""
    from pcgmix_b200 import native
    from pcgmix_b200.augmentations import pcgmix_on_device
    padded = res.padded()
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    out = pcgmix_on_device(padded, res.table.frames[: res.n_cycles], mix, lam[0], lam[1], knots, knot, err_flag=err)
    return out, int(err.item())


@pytest.mark.parametrize("warp", [False, True])
def test_cycles_touching_the_end_of_the_tensor_and_too_long_cycles(warp):
    """Rows that end exactly at the last element of a tensor whose size is not a multiple of 4 (the
    aligned loads must not run past it), a cycle clipped by the end of its recording, and cycles longer
    than L: blended where the clamped windows agree (like the reference's slices), copied unmixed and
    flagged where they do not — like the two-step path."""
    from pcgmix_b200 import draws, resident
    rng = np.random.default_rng(3)
    t_len, length = 3001, 1200
    signal = rng.standard_normal((2, 1, t_len)).astype(np.float32)              # 6002 floats
    rows = [
        [1, 2301, 3001, 0, 100, 300, 400, 700],      # ends at the very last element of the tensor
        [0, 2303, 3001, 0, 120, 280, 410, 698],      # ends at the end of recording 0
        [1, 5, 905, 0, 90, 350, 460, 900],
        [0, 2, 1502, 0, 200, 600, 800, 1500],        # longer than L, blends with a short partner (and as partner of row 4)
        [1, 2500, 3400, 0, 100, 400, 500, 900],      # table says 900 samples, the recording has 501: rest is padding
        [0, 1000, 1000, 0, 0, 0, 0, 0],              # empty cycle
        [1, 0, 1600, 0, 200, 600, 900, 1600],        # longer than L against long row 3: clamped widths differ -> flagged, copied
    ]
    res = _hand_made(signal, rows, length)
    mix = torch.tensor([1, 2, 4, 0, 3, 4, 3], dtype=torch.int32, device="cuda")
    lam = draws.lambda_pair_fp32(0.37)
    knots = torch.from_numpy(rng.normal(1.0, 0.2, (7, 6, 1))).cuda() if warp else None
    got = resident.mix_rows(res, None, mix, lam[0], lam[1], knots, 4)
    want, flag = _two_step(res, mix, lam, knots, 4)
    assert np.array_equal(_bits(got.cpu().numpy()), _bits(want.cpu().numpy()))
    assert int(res.err_flag.item()) == flag and flag != 0
    with pytest.raises(ValueError):
        res.check()
    if not warp:                                                                  # the long row really was blended
        padded = res.padded().cpu().numpy()
        want3 = orc.mix_pair(padded[3], padded[0], np.array(rows[3][3:]), np.array(rows[0][3:]), lam[0])
        assert np.array_equal(_bits(got[3].cpu().numpy()), _bits(want3))
        assert np.array_equal(_bits(got[6].cpu().numpy()), _bits(padded[6]))


def test_out_of_range_rows_and_partners_are_flagged_not_followed():
    from pcgmix_b200 import draws, native, resident
    rng = np.random.default_rng(4)
    signal = rng.standard_normal((1, 2, 4000)).astype(np.float32)
    rows = [[0, 0, 900, 0, 100, 400, 500, 900], [0, 1000, 1800, 0, 90, 300, 420, 800], [7, 0, 900, 0, 100, 400, 500, 900]]
    res = _hand_made(signal, rows, 1024)
    lam = draws.lambda_pair_fp32(0.5)
    guard = torch.full((6, 2, 1024), 7.0, device="cuda")
    out = guard[1:5]
    sel = torch.tensor([0, 5, 2, 1], dtype=torch.int32, device="cuda")          # row 5 does not exist; row 2 names recording 7
    mix = torch.tensor([3, 0, 0, 9], dtype=torch.int32, device="cuda")          # partner 9 does not exist
    resident.mix_rows(res, sel, mix, lam[0], lam[1], out=out)
    torch.cuda.synchronize()
    assert int(res.err_flag.item()) & native.ERR_BAD_PARTNER
    with pytest.raises(IndexError):
        res.check()
    assert torch.all(guard[0] == 7.0) and torch.all(guard[5] == 7.0)             # nothing written outside out
    padded = res.padded()
    assert torch.equal(out[1], torch.zeros_like(out[1])) and torch.equal(out[2], torch.zeros_like(out[2]))
    assert torch.equal(out[3], padded[1])                                       # bad partner: own cycle, unmixed
    want0 = orc.mix_pair(padded[0].cpu().numpy(), padded[1].cpu().numpy(), np.array(rows[0][3:]), np.array(rows[1][3:]),
                         np.float32(0.5))
    assert np.array_equal(_bits(out[0].cpu().numpy()), _bits(np.asarray(want0, dtype=np.float32)))


def test_full_size_batch_property():
    """BASELINE-sized batch (4096 cycles x 4 ch x 2500) drawn from resident recordings: outside the
    blended windows the output is the cut cycle (times the warp curve), padding stays zero for PCGmix."""
    from pcgmix_b200 import resident
    res, _, _, rng = _resident(21, 96, 4, 60000, 2500)
    batch = 4096
    ids = rng.integers(0, res.n_cycles, batch)
    labels = rng.integers(0, 2, batch)
    ohe = torch.nn.functional.one_hot(torch.from_numpy(labels), 2).cuda()
    got, _, mix, _ = resident.augment(_Args("durratiomixup", batch), res, ids, ohe, ["a"] * batch, _Step(2), None, "cuda", None)
    res.check()
    padded = res.padded(ids)
    frames = res.frames_of(ids).numpy()
    cols = torch.arange(2500, device="cuda")[None, :]
    f = torch.from_numpy(frames).cuda()
    fp = f[torch.from_numpy(mix).cuda()]
    n = torch.minimum(f[:, 1:] - f[:, :-1], fp[:, 1:] - fp[:, :-1])
    blended = torch.zeros((batch, 2500), dtype=torch.bool, device="cuda")
    for s in range(4):
        blended |= (cols >= f[:, s:s + 1]) & (cols < f[:, s:s + 1] + n[:, s:s + 1])
    keep = ~blended[:, None, :].expand(-1, 4, -1)
    assert torch.equal(got[keep], padded[keep])
    assert torch.all(got[(cols >= f[:, 4:5])[:, None, :].expand(-1, 4, -1)] == 0)
    # a slice of it against the vectorised oracle
    sl = slice(0, 256)
    sub_ids = np.concatenate([ids[sl], ids[mix[sl]]])
    x = res.padded(sub_ids).cpu().numpy()
    fr = res.frames_of(sub_ids).numpy()
    sub_mix = np.arange(256, 512)
    want = orc.mix_batch_vectorised(x, fr, np.concatenate([sub_mix, np.arange(256)]), orc.lambda_as_float32(orc.draw_lambda(1.0, 2)))[:256]
    assert np.array_equal(_bits(got[sl].cpu().numpy()), _bits(want))


def test_processing_order_changes_nothing():
    """`order` only decides which slots are in flight together; the result must not depend on it."""
    from pcgmix_b200 import draws, resident
    res, _, _, rng = _resident(17, 10, 4, 20000, 2500)
    batch = 300
    ids = torch.from_numpy(rng.integers(0, res.n_cycles, batch).astype(np.int32)).cuda()
    mix_host = draws.same_label_pairing(rng.integers(0, 2, batch), 5)
    mix = torch.from_numpy(mix_host.astype(np.int32)).cuda()
    lam = draws.lambda_pair_fp32(0.42)
    knots = torch.from_numpy(rng.normal(1.0, 0.2, (batch, 6, 4))).cuda()
    plain = resident.mix_rows(res, ids, mix, lam[0], lam[1], knots, 4)
    for order in (draws.processing_order(mix_host), rng.permutation(batch).astype(np.int32)):
        got = resident.mix_rows(res, ids, mix, lam[0], lam[1], knots, 4, order_dev=torch.from_numpy(np.asarray(order, np.int32)).cuda())
        assert torch.equal(got, plain)
    res.check()


def test_recordings_beyond_two_to_the_31_elements():
    """Element indices into the recordings are 64-bit: cycles cut from the far end of a 2.4 G-element
    tensor (9.7 GB) must come out exactly like cycles cut from a small copy of that region."""
    from pcgmix_b200 import draws, resident
    free, _ = torch.cuda.mem_get_info()
    if free < 24 * 2 ** 30:
        pytest.skip("needs ~10 GB of free device memory")
    rng = np.random.default_rng(8)
    t_len, length = 805_306_371, 2500                                   # 3 recordings x 1 channel: 2 415 919 113 elements
    big = torch.zeros((3, 1, t_len), dtype=torch.float32, device="cuda")
    tail = 40_000
    region = torch.from_numpy(rng.standard_normal(tail).astype(np.float32)).cuda()
    big[2, 0, t_len - tail:] = region                                   # the last 40 000 samples of the last recording
    small = torch.zeros((3, 1, tail), dtype=torch.float32, device="cuda")
    small[2, 0] = region
    rows_small, pos = [], 3
    while pos + 1600 < tail:
        d = rng.integers([90, 150, 70, 300], [160, 400, 130, 900])
        f = np.concatenate([[0], np.cumsum(d)])
        rows_small.append([2, pos, pos + int(f[4]), *f.tolist()])
        pos += int(f[4]) + int(rng.integers(0, 7))
    rows_small[-1][2] = tail                                            # the last cycle ends with the tensor
    rows_small[-1][7] = tail - rows_small[-1][1]
    rows_small[-1][3:7] = [0, 1, 2, 3]
    rows_big = [[r[0], r[1] + t_len - tail, r[2] + t_len - tail, *r[3:]] for r in rows_small]
    n = len(rows_small)
    mix = torch.from_numpy(rng.permutation(n).astype(np.int32)).cuda()
    lam = draws.lambda_pair_fp32(0.3)
    knots = torch.from_numpy(rng.normal(1.0, 0.2, (n, 6, 1))).cuda()
    want = resident.mix_rows(_hand_made_dev(small, rows_small, length), None, mix, lam[0], lam[1], knots, 4)
    res_big = _hand_made_dev(big, rows_big, length)
    got = resident.mix_rows(res_big, None, mix, lam[0], lam[1], knots, 4)
    assert int(res_big.err_flag.item()) == 0
    assert torch.equal(got, want)
    assert torch.equal(res_big.padded(), _hand_made_dev(small, rows_small, length).padded())


def _hand_made_dev(signal_dev, rows, length):
    from pcgmix_b200 import resident, segmentation
    dev = signal_dev.device
    table = segmentation.CycleTable(torch.tensor(rows, dtype=torch.int32, device=dev),
                                    torch.tensor([0, len(rows)], dtype=torch.int32, device=dev),
                                    torch.zeros(1, dtype=torch.int32, device=dev))
    return resident.ResidentCycles(signal_dev, table, length, len(rows), torch.zeros(1, dtype=torch.int32, device=dev))
