"""Segmentation-index oracle and CUDA kernels against fixtures produced by executing the
reference's OWN statements (tests/golden/make_golden_segmentation.py cuts them out of
databuilder.ipynb cells 14 / 25 / 6 and classical.py:feature_vector_seg and runs them verbatim on
synthetic annotations).  Integer results must be equal; the cut signals are ramps, so the stored
(sum, non-zero count, first value) of every cut pins the slice it came from."""
import numpy as np
import pytest
import torch

from oracle import segmentation_oracle as seg_orc


def _split(flat, offs):
    return [flat[offs[i]:offs[i + 1]] for i in range(len(offs) - 1)]


def _summary(a):
    return [float(np.sum(a, dtype=np.float64)), float(np.count_nonzero(a)), float(np.ravel(a)[0])]


def _signals(offs):
    return [(np.arange(offs[r + 1] - offs[r]) + 1000 * r).astype(np.float32) for r in range(len(offs) - 1)]


# ------------------------------------------------------------------------------------------ oracle
def test_oracle_cell14_dense_states(golden):
    g = golden("seg_cell14")
    states = _split(g["states"], g["states_offsets"])
    signals = _signals(g["signal_offsets"])
    frames, recs, cuts = [], [], []
    for r, st in enumerate(states):
        rel, a0, a1 = seg_orc.cycles_from_dense(st, int(g["downsample"]))
        for i in range(len(a0)):
            frames.append(rel[i]); recs.append(r)
            cuts.append(seg_orc.cut_and_pad(signals[r], a0[i], a1[i], int(g["cut_length"])))
    assert np.array_equal(np.stack(frames), g["frames"]) and np.array_equal(recs, g["recording"])
    assert np.array_equal(np.stack(cuts[:len(g["cut_head"])]), g["cut_head"])
    assert np.array_equal(np.array([_summary(c) for c in cuts]), g["cut_summary"])


def test_oracle_cell25_state_table(golden):
    g = golden("seg_cell25")
    pos = _split(g["positions"], g["offsets"])
    names = _split(g["names"], g["offsets"])
    signals = _signals(g["signal_offsets"])
    frames, recs, cuts = [], [], []
    for r in range(len(pos)):
        codes = [seg_orc.state_code(str(n)) for n in names[r]]
        rel, a0, a1 = seg_orc.cycles_from_transitions(pos[r], codes, int(g["downsample"]))
        for i in range(len(a0)):
            frames.append(rel[i]); recs.append(r)
            cuts.append(seg_orc.cut_and_pad(signals[r], a0[i], a1[i], int(g["cut_length"])))
    assert np.array_equal(np.stack(frames), g["frames"]) and np.array_equal(recs, g["recording"])
    assert np.array_equal(np.stack(cuts[:len(g["cut_head"])]), g["cut_head"])
    assert np.array_equal(np.array([_summary(c) for c in cuts]), g["cut_summary"])


def _mel(r, cols):
    return (np.arange(8)[:, None] * 10000 + np.arange(cols)[None, :] + 7 * r).astype(np.float32)


def test_oracle_cell6_spectrogram_frames(golden):
    g = golden("seg_cell6")
    pos = _split(g["positions"], g["offsets"])
    names = _split(g["names"], g["offsets"])
    frames, recs, specs = [], [], []
    for r in range(len(pos)):
        codes = [seg_orc.state_code(str(n)) for n in names[r]]
        rel, a0, a1 = seg_orc.cycles_from_transitions_spec(pos[r], codes, int(g["spec_cols"][r]), int(g["rec_len"][r]))
        for i in range(len(a0)):
            if rel[i][4] <= int(g["spec_frames"]):
                frames.append(rel[i]); recs.append(r)
                specs.append(seg_orc.cut_and_pad_spec(_mel(r, int(g["spec_cols"][r])), a0[i], a1[i], int(g["spec_frames"])))
    assert np.array_equal(np.stack(frames), g["frames"]) and np.array_equal(recs, g["recording"])
    assert np.array_equal(np.stack(specs[:len(g["spec_head"])]), g["spec_head"])
    assert np.array_equal(np.array([_summary(s) for s in specs]), g["spec_summary"])


def test_oracle_duration_features(golden):
    g = golden("duration_features")
    for i in range(g["frames"].shape[0]):
        got = seg_orc.duration_features(g["frames"][i], int(g["fs"]))
        assert np.array_equal(got.view(np.uint64), g["features"][i].view(np.uint64)), i


# -------------------------------------------------------------------------------------------- CUDA
@pytest.mark.gpu
def test_cuda_cell14_dense_states(golden):
    from pcgmix_b200 import segmentation
    g = golden("seg_cell14")
    states = _split(g["states"], g["states_offsets"])
    signals = _signals(g["signal_offsets"])
    t_max = max(len(s) for s in states)
    padded = np.stack([np.concatenate([s, np.full(t_max - len(s), s[-1], np.int8)]) for s in states])   # hold the last state
    table = segmentation.cycles_from_dense_states(torch.from_numpy(padded).cuda(), int(g["downsample"])).check()
    n = table.total()
    cyc = table.cycles[:n].cpu().numpy()
    assert np.array_equal(cyc[:, 3:], g["frames"]) and np.array_equal(cyc[:, 0], g["recording"])
    s_max = max(len(s) for s in signals)
    sig = np.zeros((len(signals), 1, s_max), np.float32)
    for r, s in enumerate(signals):
        sig[r, 0, :len(s)] = s
    cut = segmentation.cut_cycles(torch.from_numpy(sig).cuda(), table, int(g["cut_length"]), n).cpu().numpy()[:, 0]
    assert np.array_equal(cut[:len(g["cut_head"])], g["cut_head"])
    assert np.array_equal(np.array([_summary(c) for c in cut]), g["cut_summary"])


@pytest.mark.gpu
def test_cuda_cell25_state_table(golden):
    from pcgmix_b200 import segmentation
    g = golden("seg_cell25")
    codes = np.array([segmentation.state_code(str(n)) for n in g["names"]], np.int8)
    table = segmentation.cycles_from_state_table(torch.from_numpy(g["positions"].astype(np.int32)).cuda(),
                                                 torch.from_numpy(codes).cuda(),
                                                 torch.from_numpy(g["offsets"].astype(np.int32)).cuda(), int(g["downsample"])).check()
    n = table.total()
    cyc = table.cycles[:n].cpu().numpy()
    assert np.array_equal(cyc[:, 3:], g["frames"]) and np.array_equal(cyc[:, 0], g["recording"])
    signals = _signals(g["signal_offsets"])
    s_max = max(len(s) for s in signals)
    sig = np.zeros((len(signals), 1, s_max), np.float32)
    for r, s in enumerate(signals):
        sig[r, 0, :len(s)] = s
    cut = segmentation.cut_cycles(torch.from_numpy(sig).cuda(), table, int(g["cut_length"]), n).cpu().numpy()[:, 0]
    assert np.array_equal(cut[:len(g["cut_head"])], g["cut_head"])
    assert np.array_equal(np.array([_summary(c) for c in cut]), g["cut_summary"])


@pytest.mark.gpu
def test_cuda_cell6_spectrogram_frames(golden):
    from pcgmix_b200 import segmentation
    g = golden("seg_cell6")
    codes = np.array([segmentation.state_code(str(n)) for n in g["names"]], np.int8)
    frames_all, recs_all = [], []
    # spec_cols differs per recording in the fixture; the kernel takes one value per call
    for r in range(len(g["rec_len"])):
        lo, hi = int(g["offsets"][r]), int(g["offsets"][r + 1])
        table = segmentation.cycles_from_state_table(
            torch.from_numpy(g["positions"][lo:hi].astype(np.int32)).cuda(), torch.from_numpy(codes[lo:hi]).cuda(),
            torch.tensor([0, hi - lo], dtype=torch.int32).cuda(), 1, int(g["spec_cols"][r]),
            torch.tensor([int(g["rec_len"][r])], dtype=torch.int32).cuda()).check()
        cyc = table.cycles[:table.total()].cpu().numpy()
        for row in cyc:
            if row[7] <= int(g["spec_frames"]):
                frames_all.append(row[3:]); recs_all.append(r)
    assert np.array_equal(np.stack(frames_all), g["frames"]) and np.array_equal(recs_all, g["recording"])


@pytest.mark.gpu
def test_cuda_duration_features(golden):
    from pcgmix_b200 import segmentation
    g = golden("duration_features")
    got = segmentation.duration_features(torch.from_numpy(g["frames"].astype(np.int32)).cuda(), int(g["fs"])).cpu().numpy()
    assert np.array_equal(got.view(np.uint64), g["features"].view(np.uint64))
