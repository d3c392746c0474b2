"""Pin the segmentation-index logic against the reference's OWN statements.

The cycle builder of the reference lives in notebook cells (``databuilder.ipynb``) that cannot be
run as a whole: they read audio and annotation files that are not in the repository.  The index
logic itself is a handful of plain-Python statements in the middle of those cells, so this script
cuts exactly those statements out of the notebook (located by their text, checked verbatim),
executes them unmodified on synthetic annotations, and stores inputs + outputs as fixtures:

  seg_cell14.npz   dense per-sample states -> transitions -> cycles, //4, cut + resize(2000)
                   (databuilder.ipynb cell 14: the ``np.where`` block and the per-segment cut)
  seg_cell25.npz   (position, state-name) table, //2, noise skip, cut + resize(2500)   (cell 25)
  seg_cell6.npz    spectrogram positions via round(), offsets, column cut + np.pad       (cell 6)
  duration_features.npz   classical.py:feature_vector_seg, the duration block (lines 248-283)

Run in the build container:  python tests/golden/make_golden_segmentation.py
"""
from __future__ import annotations

import ast
import copy
import io
import json
import os
import sys
import textwrap
from contextlib import redirect_stdout

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PCGMIX_REFERENCE_ROOT", "/root/reference")


def cell_source(index):
    nb = json.load(open(os.path.join(REF, "databuilder.ipynb")))
    return "".join(nb["cells"][index]["source"]).split("\n")


def block(lines, first_text, last_text, start_at=0):
    """Lines from the first one containing ``first_text`` through the first later one containing
    ``last_text``, dedented.  Both anchors must exist."""
    i0 = next(i for i in range(start_at, len(lines)) if first_text in lines[i])
    i1 = next(i for i in range(i0, len(lines)) if last_text in lines[i])
    return textwrap.dedent("\n".join(lines[i0:i1 + 1])), i1 + 1


def run(code, ns):
    with redirect_stdout(io.StringIO()):
        exec(compile(code, "<reference statements>", "exec"), ns)
    return ns


def summarise(arrays, keep=8):
    """Full copies of the first ``keep`` outputs plus (float64 sum, non-zero count, first value) of all:
    the inputs are ramps, so these three numbers pin the slice a cycle was cut from."""
    full = np.stack(arrays[:keep]) if arrays else np.zeros((0,))
    summ = np.array([[float(np.sum(a, dtype=np.float64)), float(np.count_nonzero(a)), float(np.ravel(a)[0])] for a in arrays])
    return full, summ


def ragged(list_of_arrays, dtype):
    offs = np.cumsum([0] + [len(a) for a in list_of_arrays]).astype(np.int64)
    flat = np.concatenate([np.asarray(a, dtype=dtype).reshape(-1) for a in list_of_arrays]) if list_of_arrays else np.zeros(0, dtype)
    return flat, offs


def synth_states(rng, n, fs):
    lo, hi = np.array([90, 150, 70, 300]), np.array([160, 400, 130, 900])
    out = np.empty(n)
    state, pos, first = int(rng.integers(0, 4)), 0, True
    while pos < n:
        ln = max(1, int(rng.integers(lo[state], hi[state] + 1)) * fs // 1000)
        if first:
            ln, first = max(1, int(rng.integers(1, ln + 1))), False
        out[pos:pos + ln] = state + 1
        pos += ln
        state = (state + 1) % 4
    return out


def main():
    rng = np.random.default_rng(20241019)

    # ------------------------------------------------------------------ cell 14 (dense states)
    c14 = cell_source(14)
    find_code, nxt = block(c14, "frames = np.where(states[:-1] != states[1:])[0]", "all_data['excluded'].append(excluded)")
    cut_code, _ = block(c14, "seg_frames = frames[start:start+5] - frames[start]", "seg_y.resize(2000)", nxt)
    assert "frames = [f//4 for f in frames]" in find_code and "raise Exception('Segment states are not correct!')" in find_code
    rec_states, rec_signal, out_frames, out_rec, out_cut = [], [], [], [], []
    for r in range(12):
        n = int(rng.integers(30000, 70000))
        states = synth_states(rng, n, 4000)                      # np.loadtxt gives float64
        # the 1 kHz signal the cycles are cut from: a ramp (exact in fp32, compresses to nothing, and a
        # wrong offset shows up as a wrong value)
        y_hat = (np.arange(n // 4 + 8) + 1000 * r).astype(np.float32)
        all_data = {k: [] for k in ("label", "frames", "wav", "id", "sig_qual", "excluded")}
        ns = run(find_code, dict(np=np, states=states.copy(), label=0, rec=f"r{r}", idx="ID_0", sig_qual=1, excluded=1,
                                 all_data=all_data))
        frames, seg_starts = ns["frames"], ns["seg_starts"]
        for start in seg_starts:
            ns2 = run(cut_code, dict(np=np, copy=copy, frames=frames, start=start, y_hat=y_hat, rec=f"r{r}"))
            out_cut.append(np.array(ns2["seg_y"]))
        out_frames += [np.asarray(f, dtype=np.int64) for f in all_data["frames"]]
        out_rec += [r] * len(all_data["frames"])
        rec_states.append(states.astype(np.int8))
        rec_signal.append(y_hat)
    st_flat, st_off = ragged(rec_states, np.int8)
    sg_flat, sg_off = ragged(rec_signal, np.float32)
    np.savez_compressed(os.path.join(HERE, "seg_cell14.npz"), entry=np.array("databuilder.ipynb cell 14 statements, executed verbatim"),
                        states=st_flat, states_offsets=st_off, signal_offsets=sg_off, downsample=np.int64(4),
                        signal_rule=np.array("signal of recording r = arange(len) + 1000*r, float32"),
                        frames=np.stack(out_frames), recording=np.array(out_rec, np.int64), cut_head=summarise(out_cut)[0],
                        cut_summary=summarise(out_cut)[1], cut_length=np.int64(2000))
    print("cell 14:", len(out_frames), "cycles")

    # ------------------------------------------------------------------ cell 25 (PhysioNet table)
    c25 = cell_source(25)
    ds_code, nxt = block(c25, "frames = [x//2 for x in frames]", "frames = [x//2 for x in frames]")
    find_code, nxt = block(c25, "seg_starts = []", "train_data['sig_qual'].append(sig_qual)", nxt)
    cut_code, _ = block(c25, "seg_frames = frames[start:start+5] - frames[start]", "seg_y.resize(2500)", nxt)
    assert "if 'N' in ''.join(seg_states):" in find_code and "continue" in find_code
    names = ["S1", "systole", "S2", "diastole"]
    tab_pos, tab_names, tab_sig, out_frames, out_rec, out_cut = [], [], [], [], [], []
    for r in range(25):
        n_tr = int(rng.integers(0, 70))
        state, t = int(rng.integers(0, 4)), int(rng.integers(1, 900))
        pos, st = [], []
        for _ in range(n_tr):
            name = names[state]
            if rng.random() < 0.06:
                name = "(N" if rng.random() < 0.5 else "N)"
            pos.append(np.int64(t))
            st.append(name)
            t += int(rng.integers(60, 1800))
            state = (state + 1) % 4
        y_hat = (np.arange(t // 2 + 8) + 1000 * r).astype(np.float32)
        data = {"frames": [], "label": [], "wav": [], "sig_qual": []}
        ns = dict(np=np, frames=list(pos), fr=list(pos), states=list(st), wav=f"a{r:04d}", label=0, sig_qual=1,
                  test_wavs=[], train_data=data, test_data=copy.deepcopy(data))
        run(ds_code, ns)
        run(find_code, ns)
        for start in ns["seg_starts"]:
            ns2 = run(cut_code, dict(np=np, copy=copy, frames=ns["frames"], states=ns["states"], start=start, y_hat=y_hat, wav=f"a{r:04d}"))
            out_cut.append(np.array(ns2["seg_y"]))
        out_frames += [np.asarray(f, dtype=np.int64) for f in data["frames"]]
        out_rec += [r] * len(data["frames"])
        tab_pos.append(np.asarray(pos, np.int64))
        tab_names.append(st)
        tab_sig.append(y_hat)
    p_flat, p_off = ragged(tab_pos, np.int64)
    sg_flat, sg_off = ragged(tab_sig, np.float32)
    np.savez_compressed(os.path.join(HERE, "seg_cell25.npz"), entry=np.array("databuilder.ipynb cell 25 statements, executed verbatim"),
                        positions=p_flat, offsets=p_off, names=np.array([n for rec in tab_names for n in rec]),
                        signal_offsets=sg_off, downsample=np.int64(2),
                        signal_rule=np.array("signal of recording r = arange(len) + 1000*r, float32"),
                        frames=np.stack(out_frames) if out_frames else np.zeros((0, 5), np.int64),
                        recording=np.array(out_rec, np.int64), cut_head=summarise(out_cut)[0], cut_summary=summarise(out_cut)[1],
                        cut_length=np.int64(2500))
    print("cell 25:", len(out_frames), "cycles")

    # ------------------------------------------------------------------ cell 6 (spectrogram frames)
    c6 = cell_source(6)
    find_code, nxt = block(c6, "seg_starts = []", "seg_frames = frames[i:i+5] - frames[i]")
    map_code, nxt = block(c6, "frames_spec = [round(f*mel_spectrogram_db.shape[1]/len(y)) for f in frames]",
                          "frames_spec = [round(f*mel_spectrogram_db.shape[1]/len(y)) for f in frames]", nxt)
    rel_code, nxt = block(c6, "seg_frames_spec = np.array(frames_spec[start:start+5]) - frames_spec[start]",
                          "seg_frames_spec = np.array(frames_spec[start:start+5]) - frames_spec[start]", nxt)
    spec_code, nxt = block(c6, "spec = mel_spectrogram_db[:, frames_spec[start]:frames_spec[start+4]]",
                           "spec = mel_spectrogram_db[:, frames_spec[start]:frames_spec[start+4]]", nxt)
    pad_code, _ = block(c6, "pad_size = ((0, 0), (0, max(0, spec_frames - spec.shape[1])))", "spec = np.pad(spec, pad_size, mode='constant')", nxt)
    tab_pos, tab_names, tab_len, tab_cols, mels, out_frames, out_rec, out_spec = [], [], [], [], [], [], [], []
    for r in range(20):
        n_tr = int(rng.integers(0, 50))
        state, t = int(rng.integers(0, 4)), int(rng.integers(1, 900))
        pos, st = [], []
        for _ in range(n_tr):
            pos.append(np.int64(t))
            st.append(names[state])
            t += int(rng.integers(120, 1800))
            state = (state + 1) % 4
        n_samples = t + 50 if r % 3 else 2 * max(t, 1)            # some lengths that make f*T/len land on .5
        n_cols = int(rng.integers(100, 900)) if r % 3 else max(n_samples // 4, 1)
        mel = (np.arange(8)[:, None] * 10000 + np.arange(n_cols)[None, :] + 7 * r).astype(np.float32)
        y = np.zeros(n_samples, np.float32)
        ns = dict(np=np, frames=list(pos), fr=list(pos), states=list(st), wav=f"a{r:04d}", label=0, sig_qual=1)
        run(find_code, ns)
        ns.update(mel_spectrogram_db=mel, y=y)
        run(map_code, ns)
        for start in ns["seg_starts"]:
            ns2 = dict(np=np, frames_spec=ns["frames_spec"], start=start, mel_spectrogram_db=mel, spec_frames=128)
            run(rel_code, ns2)
            run(spec_code, ns2)
            if ns2["spec"].shape[1] <= 128:                      # the reference does not truncate wider cycles
                run(pad_code, ns2)
                out_spec.append(np.array(ns2["spec"]))
                out_frames.append(np.asarray(ns2["seg_frames_spec"], np.int64))
                out_rec.append(r)
        tab_pos.append(np.asarray(pos, np.int64)); tab_names.append(st); tab_len.append(n_samples); tab_cols.append(n_cols); mels.append(mel)
    p_flat, p_off = ragged(tab_pos, np.int64)
    m_flat, m_off = ragged([m.reshape(-1) for m in mels], np.float32)
    np.savez_compressed(os.path.join(HERE, "seg_cell6.npz"), entry=np.array("databuilder.ipynb cell 6 statements, executed verbatim"),
                        positions=p_flat, offsets=p_off, names=np.array([n for rec in tab_names for n in rec]),
                        rec_len=np.array(tab_len, np.int64), spec_cols=np.array(tab_cols, np.int64),
                        mel_rule=np.array("mel of recording r = arange(8)[:,None]*10000 + arange(cols)[None,:] + 7*r, float32"),
                        mel_rows=np.int64(8), frames=np.stack(out_frames), recording=np.array(out_rec, np.int64),
                        spec_head=summarise(out_spec, 4)[0], spec_summary=summarise(out_spec)[1], spec_frames=np.int64(128))
    print("cell 6:", len(out_frames), "cycles (<= 128 columns)")

    # ------------------------------------------------------------------ classical.py duration block
    src = open(os.path.join(REF, "classical.py")).read().split("\n")
    i0 = next(i for i, l in enumerate(src) if l.startswith("def feature_vector_seg("))
    i1 = next(i for i in range(i0, len(src)) if "duration_ratio_diastole_S2 = round(duration_diastole/duration_S2, 4)" in src[i])
    body = textwrap.dedent("\n".join(src[i0 + 1:i1 + 1]))
    ast.parse(body)
    names_out = ["duration_RR", "BPM", "duration_S1", "duration_systole", "duration_S2", "duration_diastole",
                 "duration_ratio_S1_S2", "duration_ratio_systole_diastole", "duration_ratio_S1_RR", "duration_ratio_systole_RR",
                 "duration_ratio_S2_RR", "duration_ratio_diastole_RR", "duration_ratio_systole_S1", "duration_ratio_diastole_S2"]
    n = 3000
    dur = rng.integers(1, 1300, size=(n, 4))
    frames = np.concatenate([np.zeros((n, 1), np.int64), np.cumsum(dur, axis=1)], axis=1)
    frames[0] = [0, 1, 100, 132, 500]          # 1/32 = 0.03125: an exact decimal tie
    frames[1] = [0, 5, 165, 325, 800]
    frames[2] = [0, 3, 99, 195, 1000]
    feats = np.zeros((n, 14))
    for i in range(n):
        ns = run(body, dict(np=np, data=np.zeros(int(frames[i, 4]) + 5, np.float32), frames=frames[i]))
        feats[i] = [ns[k] for k in names_out]
    np.savez_compressed(os.path.join(HERE, "duration_features.npz"), entry=np.array("classical.py feature_vector_seg, duration block, executed verbatim"),
                        frames=frames, features=feats, fs=np.int64(1000))
    print("duration features:", n, "cycles")


if __name__ == "__main__":
    main()
