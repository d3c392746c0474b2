"""Segmentation-index kernels (annotations -> cycles, cut + pad, duration features) against the
CPU restatement of the reference's notebook logic.  Integer work: everything is bit-exact."""
import numpy as np
import pytest
import torch

from oracle import segmentation_oracle as seg_orc

pytestmark = pytest.mark.gpu


def _oracle_dense(states, downsample):
    rel, starts, stops, recs = [], [], [], []
    for r in range(states.shape[0]):
        f, a, b = seg_orc.cycles_from_dense(states[r], downsample)
        rel.append(f); starts.append(a); stops.append(b); recs.append(np.full(len(a), r))
    return np.concatenate(rel), np.concatenate(starts), np.concatenate(stops), np.concatenate(recs)


@pytest.mark.parametrize("shape,fs,ds", [((32, 10000), 2000, 1), ((32, 10000), 2000, 2), ((7, 40001), 4000, 4),
                                         ((3, 17), 2000, 1), ((1, 0), 2000, 1), ((5, 120000), 4000, 4)])
def test_dense_states_to_cycles(shape, fs, ds):
    from pcgmix_b200 import segmentation, synth
    rng = np.random.default_rng(shape[0] * 31 + shape[1])
    states = synth.dense_states(rng, shape[0], shape[1], fs) if shape[1] > 0 else np.zeros(shape, np.int8)
    table = segmentation.cycles_from_dense_states(torch.from_numpy(states).cuda(), ds).check()
    n = table.total()
    rel, starts, stops, recs = _oracle_dense(states, ds)
    assert n == len(starts)
    cyc = table.cycles[:n].cpu().numpy()
    assert np.array_equal(cyc[:, 0], recs)
    assert np.array_equal(cyc[:, 1], starts) and np.array_equal(cyc[:, 2], stops)
    assert np.array_equal(cyc[:, 3:], rel)
    ptr = table.row_ptr.cpu().numpy()
    assert np.array_equal(np.diff(ptr), np.bincount(recs.astype(np.int64), minlength=shape[0]))


def test_dense_bad_pattern_is_flagged():
    from pcgmix_b200 import segmentation
    states = np.array([[4, 4, 1, 1, 2, 2, 4, 4, 1, 1, 2, 3, 4, 1, 1]], np.int8)   # S1, systole, then diastole: S2 missing
    with pytest.raises(seg_orc.SegmentPatternError):
        seg_orc.cycles_from_dense(states[0])
    table = segmentation.cycles_from_dense_states(torch.from_numpy(states).cuda())
    with pytest.raises(segmentation.SegmentationError):
        table.check()


def _random_tables(rng, n_rec, noisy):
    names = ["S1", "systole", "S2", "diastole"]
    pos_all, code_all, offs, lens = [], [], [0], []
    per_rec = []
    for r in range(n_rec):
        n_tr = int(rng.integers(0, 60))
        state = int(rng.integers(0, 4))
        t = int(rng.integers(1, 500))
        pos, codes = [], []
        for _ in range(n_tr):
            name = names[state]
            if noisy and rng.random() < 0.08:
                name = "(N" if rng.random() < 0.5 else "N)"
            pos.append(t)
            codes.append(seg_orc.state_code(name))
            t += int(rng.integers(40, 900))
            state = (state + 1) % 4
        per_rec.append((pos, codes, t + 100))
        pos_all += pos; code_all += codes; offs.append(len(pos_all)); lens.append(t + 100)
    return per_rec, np.array(pos_all, np.int32), np.array(code_all, np.int8), np.array(offs, np.int32), np.array(lens, np.int32)


@pytest.mark.parametrize("mode", ["physionet_ds2", "plain", "spectrogram"])
def test_state_table_to_cycles(mode):
    from pcgmix_b200 import segmentation
    rng = np.random.default_rng(hash(mode) % 1000)
    per_rec, pos, codes, offs, lens = _random_tables(rng, 40, noisy=True)
    ds = 2 if mode == "physionet_ds2" else 1
    spec_cols = 517 if mode == "spectrogram" else 0
    table = segmentation.cycles_from_state_table(
        torch.from_numpy(pos).cuda(), torch.from_numpy(codes).cuda(), torch.from_numpy(offs).cuda(), ds, spec_cols,
        torch.from_numpy(lens).cuda() if spec_cols else None).check()
    n = table.total()
    want = []
    for r, (p, c, ln) in enumerate(per_rec):
        if spec_cols:
            f, a, b = seg_orc.cycles_from_transitions_spec(p, c, spec_cols, ln)
        else:
            f, a, b = seg_orc.cycles_from_transitions(p, c, ds)
        for i in range(len(a)):
            want.append([r, a[i], b[i]] + f[i].tolist())
    want = np.array(want, np.int64).reshape(-1, 8)
    assert n == want.shape[0]
    assert np.array_equal(table.cycles[:n].cpu().numpy(), want)


def test_round_half_even_in_spectrogram_mapping():
    from pcgmix_b200 import segmentation
    # positions chosen so that f*T_spec/len(y) is exactly k + 0.5
    pos = np.array([5, 15, 25, 35, 45, 55], np.int32)          # *1/10 -> 0.5, 1.5, 2.5 ...
    codes = np.array([1, 2, 3, 4, 1, 2], np.int8)
    table = segmentation.cycles_from_state_table(torch.from_numpy(pos).cuda(), torch.from_numpy(codes).cuda(),
                                                 torch.tensor([0, 6], dtype=torch.int32).cuda(), 1, 10,
                                                 torch.tensor([100], dtype=torch.int32).cuda()).check()
    f, a, b = seg_orc.cycles_from_transitions_spec(pos.tolist(), codes.tolist(), 10, 100)
    assert table.total() == 1
    assert table.cycles[0].cpu().tolist() == [0, int(a[0]), int(b[0])] + f[0].tolist() == [0, 0, 4, 0, 2, 2, 4, 4]


@pytest.mark.parametrize("length", [4400, 2500, 1001])
def test_cut_and_pad_cycles(length):
    from pcgmix_b200 import segmentation, synth
    rng = np.random.default_rng(length)
    n_rec, n_samples, bands = 6, 10000, 4
    states = synth.dense_states(rng, n_rec, n_samples, 2000)
    signal = rng.standard_normal((n_rec, bands, n_samples)).astype(np.float32)
    table = segmentation.cycles_from_dense_states(torch.from_numpy(states).cuda()).check()
    out = segmentation.cut_cycles(torch.from_numpy(signal).cuda(), table, length).cpu().numpy()
    cyc = table.cycles[: table.total()].cpu().numpy()
    assert out.shape == (len(cyc), bands, length)
    for i, (rec, start, stop) in enumerate(cyc[:, :3]):
        for c in range(bands):
            assert np.array_equal(out[i, c], seg_orc.cut_and_pad(signal[rec, c], start, stop, length))


def test_cycle_table_feeds_the_mixer_without_a_copy():
    """frames = cycles[:, 3:] (row stride 8) goes straight into the mix kernel."""
    from oracle import pcgmix_oracle as orc
    from pcgmix_b200 import native, segmentation, synth
    rng = np.random.default_rng(8)
    states = synth.dense_states(rng, 32, 10000, 2000)
    signal = rng.standard_normal((32, 2, 10000)).astype(np.float32)
    dev = torch.device("cuda:0")
    table = segmentation.cycles_from_dense_states(torch.from_numpy(states).to(dev)).check()
    n = table.total()
    assert 90 <= n <= 160                                  # BASELINE config 1: ~110-140 complete cycles
    x = segmentation.cut_cycles(torch.from_numpy(signal).to(dev), table, 4400, n)
    mix = rng.permutation(n).astype(np.int32)
    out = torch.empty_like(x)
    lam = np.float32(0.42)
    native.mix1d(x, out, table.frames[:n], torch.from_numpy(mix).to(dev), lam, np.float32(1) - lam)
    want = orc.mix_batch(x.cpu().numpy(), table.frames[:n].cpu().numpy(), mix, lam)
    assert np.array_equal(out.cpu().numpy().view(np.uint32), want.view(np.uint32))


def test_duration_features_match_python_rounding():
    from pcgmix_b200 import native, segmentation
    rng = np.random.default_rng(2)
    n = 5000
    dur = rng.integers(1, 1200, size=(n, 4))
    frames = np.concatenate([np.zeros((n, 1), np.int64), np.cumsum(dur, axis=1)], axis=1).astype(np.int32)
    # include exact decimal ties: ratios like 1/32 = 0.03125 -> round-half-even on the exact value
    frames[0] = [0, 1, 100, 132, 500]
    frames[1] = [0, 5, 165, 325, 800]
    for fs in (1000, 2000, 4000):
        err = torch.zeros(1, dtype=torch.int32, device="cuda:0")
        got = segmentation.duration_features(torch.from_numpy(frames).cuda(), fs, err).cpu().numpy()
        bad = 0
        for i in range(n):
            try:
                want = seg_orc.duration_features(frames[i], fs)
            except ZeroDivisionError:
                bad += 1
                assert np.isnan(got[i]).any()
                continue
            assert np.array_equal(got[i].view(np.uint64), want.view(np.uint64)), (i, fs, got[i], want)
        assert (int(err.item()) & native.ERR_ZERO_DIVISION != 0) == (bad > 0)
