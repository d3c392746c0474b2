"""CPU-only checks of the host side: method-string mini-language, seeded draws, spline map,
processing order, sharding, ABI of the shared library (symbols only — no compute without a GPU)."""
import ctypes
import os
import random
import re

import numpy as np
import pytest
import torch

from oracle import pcgmix_oracle as orc
from pcgmix_b200 import build_native, draws, native, sharding, spline, synth

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_parse_1d_methods():
    p = draws.parse_method_1d("durratiomixup")
    assert (p.branch, p.probability, p.alpha, p.mix_all) == ("durratiomixup", 1.0, 1.0, False)
    p = draws.parse_method_1d("(alpha=0.5)durmixmagwarp(0.3,2)+0.9")
    assert (p.branch, p.probability, p.alpha, p.sigma, p.knot) == ("durmixmagwarp", 0.9, 0.5, 0.3, 2)
    p = draws.parse_method_1d("durmixmagwarp")
    assert (p.sigma, p.knot) == (0.2, 4)
    assert draws.parse_method_1d("(mixAll)durratiomixup").mix_all
    assert draws.parse_method_1d("no-such-method") is None
    # reference dispatcher order: these never reach the PCGmix branches -> refuse loudly
    for method in ("respiratoryscale(12,20)durratiomixup", "timemask+durratiomixup", "cutmix", "mixup(same)"):
        with pytest.raises(NotImplementedError):
            draws.parse_method_1d(method)
    assert draws.parse_method_1d("(rand)durratiomixup").rand_displacement
    for method in ("(sameCVD)durratiomixup", "(saloptenv)durratiomixup", "(closestknn=3)durmixmagwarp(0.2,4)"):
        with pytest.raises(NotImplementedError):
            draws.parse_method_1d(method)


def test_parse_2d_methods():
    assert draws.parse_method_2d("durratiomixup+0.5").probability == 0.5
    p = draws.parse_method_2d("durmixcutout(0.3,0.9)")
    assert (p.branch, p.time_region_max, p.freq_region_max) == ("durmixcutout", 0.3, 0.9)
    assert draws.parse_method_2d("durmixtimemask(7)").time_region_max == 1
    assert draws.parse_method_2d("durmixfreqmask").freq_region_max == 0.2
    assert draws.parse_method_2d("durmixmagwarp(0.2,4)") is None       # not a 2D method: passthrough like the reference
    with pytest.raises(NotImplementedError):
        draws.parse_method_2d("freqmask")


def test_draws_match_oracle_and_known_answers(golden):
    g = golden("kat_draws")
    assert draws.gate(7) == g["uniform7"]
    assert draws.draw_lambda(1, 7) == g["lambda_alpha1_seed7"]
    assert np.array_equal(draws.draw_knots(2, 4, 4, 0.2), g["normals7"])
    assert np.array_equal(draws.same_label_pairing(g["labels"], 5), g["same_label_mix_seed5"])
    rng = np.random.default_rng(0)
    labels = rng.integers(0, 3, 200)
    wav = [f"{'abc'[i % 3]}{i % 7:04d}" for i in range(200)]
    for method in ("durratiomixup", "(samePCG)durratiomixup", "(sameDataset)durratiomixup", "(mixAll)durratiomixup"):
        assert np.array_equal(draws.pairing(method, labels, wav, 11), orc.pick_pairing(method, labels, wav, 11))
    lam32, oml = draws.lambda_pair_fp32(0.3)
    assert lam32.dtype == np.float32 and oml == np.float32(1) - np.float32(0.3)
    assert draws.mask_geometry(9, 0.4) == orc.mask_draws(9, 0.4)


def test_global_numpy_stream_is_left_like_the_reference_leaves_it():
    draws.draw_lambda(1, 42)
    a = np.random.normal(size=3)
    orc.draw_lambda(1, 42)
    assert np.array_equal(a, np.random.normal(size=3))


@pytest.mark.parametrize("length,knot", [(2500, 4), (4400, 4), (1001, 2), (250, 1), (128, 0), (1203, 7), (2500, 30), (17, 4)])
def test_spline_map_matches_scipy(length, knot):
    from scipy.interpolate import CubicSpline
    rng = np.random.default_rng(length + knot)
    y = rng.normal(1, 0.2, (4, knot + 2))
    x = np.linspace(0, length - 1.0, knot + 2)
    ref = np.stack([CubicSpline(x, row)(np.arange(length)) for row in y])
    got = spline.evaluate(y, length)
    assert np.max(np.abs(got - ref) / np.abs(ref)) < 1e-13
    pos, mat = spline.magwarp_tables(length, knot)
    assert np.array_equal(pos[:-1], x) and mat.shape == ((knot + 1) * 4, knot + 2)
    # the safe-deviation bound: every curve whose knots stay inside it is positive at every sample
    safe = pos[-1]
    assert 0.0 < safe < 1.0
    worst = 1.0 + safe * np.sign(rng.standard_normal((64, knot + 2)))
    assert spline.evaluate(worst, length).min() > 5e-4


def test_processing_order_is_a_permutation_following_chains():
    rng = np.random.default_rng(3)
    for n in (1, 2, 17, 4096):
        mix = rng.permutation(n)
        order = draws.processing_order(mix)
        assert sorted(order.tolist()) == list(range(n))
        nxt = {int(order[i]): int(order[i + 1]) for i in range(n - 1)}
        follows = sum(1 for b, c in nxt.items() if mix[b] == c)
        n_cycles = len({frozenset(_orbit(mix, s)) for s in range(n)}) if n <= 17 else None
        if n_cycles is not None:
            assert follows == n - n_cycles          # every step follows the pairing except at chain ends


def _orbit(mix, s):
    seen, b = [], s
    while b not in seen:
        seen.append(b)
        b = int(mix[b])
    return seen


def test_sharding_partitions_batches():
    n_batches = 245
    for world in (1, 2, 4, 8):
        seen = []
        for rank in range(world):
            mine = sharding.batches_for_rank(n_batches, rank, world)
            seen += list(mine)
            assert all(sharding.step_seed(k) == k for k in mine)
        assert sorted(seen) == list(range(n_batches))
    lo, hi = sharding.rows_for_rank(10, 3, 4)
    assert (lo, hi) == (8, 10) and sharding.rows_for_rank(10, 0, 4) == (0, 3)


def test_synthetic_generator_shapes():
    rng = np.random.default_rng(0)
    fr = synth.cycle_frames(rng, 50, 1000, 2500)
    assert fr.shape == (50, 5) and (np.diff(fr, axis=1) >= 0).all() and fr[:, 4].max() <= 2500
    x = synth.cycle_signals(rng, fr, (4,), 2500)
    assert x.dtype == np.float32 and not x[np.arange(2500)[None, None, :] >= fr[:, 4][:, None, None] + 0 * x.astype(int)].any()
    st = synth.dense_states(rng, 3, 10000, 2000)
    assert st.dtype == np.int8 and set(np.unique(st)) <= {1, 2, 3, 4}
    sf = synth.spectrogram_frames(rng, 20, 250)
    assert sf.max() <= 250 and (np.diff(sf, axis=1) >= 0).all()


def test_library_exports_every_symbol_in_the_header():
    """The C ABI in include/pcgmix_b200.h and the built library / ctypes table must agree."""
    header = open(os.path.join(ROOT, "include", "pcgmix_b200.h")).read()
    declared = set(re.findall(r"^(?:int|long long|const char\*)\s+(pcgmix_\w+)\s*\(", header, flags=re.M))
    assert declared == set(native.SIGNATURES), declared ^ set(native.SIGNATURES)
    if not os.path.exists(build_native.LIB_PATH):
        pytest.skip("library not built yet (run __graft_entry__.build())")
    lib = ctypes.CDLL(build_native.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.pcgmix_version() >= 100
    # argument counts of the ctypes table follow the header
    for name, args in native.SIGNATURES.items():
        m = re.search(r"^(?:int|long long|const char\*)\s+" + name + r"\s*\(([^;]*?)\)\s*;", header, flags=re.S | re.M)
        assert m, name
        params = [p for p in m.group(1).replace("\n", " ").split(",") if p.strip() and p.strip() != "void"]
        assert len(params) == len(args), (name, len(params), len(args))


def test_no_cpu_fallback_and_loud_failure_without_library(monkeypatch):
    import torch
    from pcgmix_b200 import augmentations

    class A:
        method = "durratiomixup"
        batch_size = 2

    class S:
        count = 0

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        augmentations.augment(A(), torch.zeros(2, 1, 8), torch.eye(2, dtype=torch.int64), torch.zeros(2, 5, dtype=torch.int64),
                              ["a", "b"], S(), None, "cpu", None)
    monkeypatch.setattr(native, "_lib", None)
    monkeypatch.setattr(build_native, "LIB_PATH", "/nonexistent/libpcgmix_b200.so")
    with pytest.raises(native.NativeLibraryError, match="no CPU fallback"):
        native.load(build_if_missing=False)


def test_product_code_does_not_import_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may touch oracle/."""
    roots = ["pcgmix-a-data-augmentation-method-for-heart-sound-classification-extended_b200", "pcgmix_b200",
             "benchmarks", "examples", "include"]
    for root in roots:
        for dirpath, _, files in os.walk(os.path.join(ROOT, root)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                    text = open(os.path.join(dirpath, f)).read()
                    assert "import oracle" not in text and "from oracle" not in text, f


def test_native_pairing_replays_cpython_random_sample():
    """The C++ host helper must give exactly CPython's random.Random(seed).sample per group."""
    if not os.path.exists(build_native.LIB_PATH):
        pytest.skip("library not built yet (run __graft_entry__.build())")
    rng = np.random.default_rng(12)
    for trial in range(400):
        n = int(rng.integers(1, 700))
        keys = rng.integers(0, int(rng.integers(1, 6)), n)
        seed = int(rng.integers(0, 2 ** 45)) if trial % 4 == 0 else trial
        got = draws._grouped_permutation(keys.tolist(), seed)
        want = draws._grouped_permutation_python(keys.tolist(), seed)
        assert np.array_equal(got, want), (trial, n, seed)
    # n = 1 groups, string keys, and the oracle's own pairing
    assert np.array_equal(draws._grouped_permutation(["x"], 3), [0])
    labels = rng.integers(0, 2, 4096)
    assert np.array_equal(draws.same_label_pairing(labels, 77), orc.same_label_mix_indices(labels, 77))


def test_header_is_plain_c(tmp_path):
    """include/pcgmix_b200.h must be usable from C (the boundary is a C ABI): compile a C99 translation
    unit that includes it and takes the address of every declared entry point."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    header = open(os.path.join(ROOT, "include", "pcgmix_b200.h")).read()
    names = sorted(set(re.findall(r"^(?:int|long long|const char\*)\s+(pcgmix_\w+)\s*\(", header, flags=re.M)))
    src = tmp_path / "abi.c"
    src.write_text('#include "pcgmix_b200.h"\nconst void* const entry_points[] = {\n'
                   + "".join(f"    (const void*)&{n},\n" for n in names) + "};\n"
                   "int version_macro = PCGMIX_B200_VERSION;\n")
    out = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic-errors", "-Wno-pedantic", "-I",
                          os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / "abi.o")],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stderr


def test_resident_path_has_no_cpu_fallback_either():
    """Recordings, annotations and cycle ids on the CPU must be refused loudly, not processed some other way."""
    import torch
    from pcgmix_b200 import resident, segmentation
    signal = torch.zeros(2, 1, 100)
    states = torch.ones(2, 100, dtype=torch.int8)
    with pytest.raises(RuntimeError, match="CUDA"):
        resident.from_dense_states(signal, states, 64)
    with pytest.raises(RuntimeError, match="CUDA"):
        segmentation.cut_cycles(signal, None, 64, 0)
    with pytest.raises(TypeError):
        resident._ids_on_device(np.array([0.5, 1.5]), torch.device("cpu"))


def test_clamped_window_rule_predicts_the_reference(golden):
    """`_common.clamped_windows` / `check_pair_windows` (and `pair_window` in csrc/common.cuh, which
    states the same rule) against what the unmodified reference did with every ordered pair of the
    long-cycle offset lists: it blends iff every state's clamped widths agree, an empty destination
    meets one source sample, or one source sample is broadcast over a wider destination (refused here)."""
    from pcgmix_b200 import _common
    g = golden("long_pairs_1d")
    fr, ok, L = g["frames"], g["ok"], g["data"].shape[-1]
    n_broadcast = 0
    for i in range(len(fr)):
        for j in range(len(fr)):
            _, _, wd, ws = _common.clamped_windows(fr[i:i + 1], fr[j:j + 1], L)
            fine = (wd == ws) | ((wd == 0) & (ws == 1))
            broadcast = (ws == 1) & (wd > 1)
            assert bool((fine | broadcast).all()) == bool(ok[i, j]), (i, j)
            pair = np.stack([fr[i], fr[j]]).astype(np.int32)
            if fine.all():
                _common.check_pair_windows(pair, np.array([1, 1]), L)
            else:
                n_broadcast += bool(ok[i, j])
                with pytest.raises(RuntimeError):
                    _common.check_pair_windows(pair, np.array([1, 1]), L)
    assert n_broadcast == 5


def test_host_frames_accepts_offsets_beyond_the_row_and_rejects_disorder():
    from pcgmix_b200 import _common
    ok = _common.host_frames(torch.tensor([[0, 10, 20, 30, 70], [0, 5, 9, 12, 40]]), 2, 50)
    assert ok.dtype == np.int32 and ok[0, 4] == 70
    for bad in ([[0, 10, 9, 30, 40]], [[-1, 10, 20, 30, 40]], [[0, 10, 20, 30, 2 ** 31]]):
        with pytest.raises(ValueError):
            _common.host_frames(torch.tensor(bad), 1, 50)
    with pytest.raises(TypeError):
        _common.host_frames(torch.tensor([[0.0, 1, 2, 3, 4]]), 1, 50)


@pytest.mark.parametrize("alpha", [1.0, 0.5, 2.0, 0.3, 7.5])
def test_lambda_and_knots_replay_is_numpy_bit_for_bit(alpha):
    """The C++ replay of NumPy's legacy stream (seed -> beta -> normal) against NumPy itself: values, and the
    state the GLOBAL stream is left in (the reference re-seeds it every step; later code sees that)."""
    for step in list(range(120)) + [2 ** 32 - 1, 3_000_000_019 % 2 ** 32]:
        for shape in ((5, 6, 4), (1, 2, 1), (33, 9, 3)):
            np.random.seed(step)
            lam = np.random.beta(alpha, alpha)
            knots = np.random.normal(loc=1.0, scale=0.2, size=shape)
            after = np.random.random_sample(3)
            np.random.seed(12345)
            got_lam, got_knots = draws.lambda_and_knots(alpha, step, shape[0], shape[1] - 2, shape[2], 0.2)
            assert got_lam == lam and np.array_equal(got_knots, knots), (alpha, step, shape)
            assert np.array_equal(np.random.random_sample(3), after), "global NumPy stream not left where the reference leaves it"


def test_lambda_and_knots_prefetch_and_fallbacks():
    np.random.seed(77)
    lam = np.random.beta(1, 1)
    knots = np.random.normal(1.0, 0.2, (64, 6, 4))
    draws.prefetch_lambda_and_knots(1, 77, 64, 4, 4, 0.2)
    draws.prefetch_lambda_and_knots(1, 77, 64, 4, 4, 0.2)            # asking twice is harmless
    got = draws.lambda_and_knots(1, 77, 64, 4, 4, 0.2)
    assert got[0] == lam and np.array_equal(got[1], knots)
    # more requests than the pool keeps: the oldest are dropped, results stay right
    for k in range(20):
        draws.prefetch_lambda_and_knots(1, 1000 + k, 8, 4, 2, 0.2)
    np.random.seed(1003)
    lam = np.random.beta(1, 1)
    knots = np.random.normal(1.0, 0.2, (8, 6, 2))
    got = draws.lambda_and_knots(1, 1003, 8, 4, 2, 0.2)
    assert got[0] == lam and np.array_equal(got[1], knots)
    # alpha <= 0: no re-seed, lambda 1, knots continue the global stream wherever it stands
    np.random.seed(5)
    np.random.random_sample(7)
    state = np.random.get_state()
    want = np.random.normal(1.0, 0.3, (4, 6, 2))
    np.random.set_state(state)
    lam0, knots0 = draws.lambda_and_knots(0.0, 9, 4, 4, 2, 0.3)
    assert lam0 == 1.0 and np.array_equal(knots0, want)
    # mirroring switched off: values unchanged, global stream untouched
    draws.mirror_numpy_global_state = False
    try:
        np.random.seed(31)
        probe = np.random.get_state()[1].copy()
        draws.lambda_and_knots(1, 8, 4, 4, 2, 0.2)
        assert np.array_equal(np.random.get_state()[1], probe)
    finally:
        draws.mirror_numpy_global_state = True


def test_processing_order_native_equals_python_walk():
    rng = np.random.default_rng(3)
    for n in (0, 1, 2, 17, 4096):
        mix = rng.permutation(n)
        assert np.array_equal(draws.processing_order(mix), draws._processing_order_python(mix))
    mix = np.array([1, 0, 7, 2])                                     # an entry outside the batch ends its chain
    assert np.array_equal(draws.processing_order(mix), draws._processing_order_python(mix))


def test_host_labels_side_channel():
    from pcgmix_b200 import _common
    target = torch.tensor([0, 1, 1, 0, 1])
    ohe = torch.nn.functional.one_hot(target, 2)
    assert np.array_equal(_common.labels_from_one_hot(ohe), [0, 1, 1, 0, 1])
    tagged = _common.with_host_labels(ohe, target)
    assert tagged is ohe and np.array_equal(_common.labels_from_one_hot(ohe), [0, 1, 1, 0, 1])
    with pytest.raises(ValueError):
        _common.with_host_labels(ohe, torch.tensor([0, 1]))


def test_prepare_step_packs_what_the_python_pieces_compute():
    """`pcgmix_host_prepare_step` (the single-call host path of a plain step) against the separate Python/C++
    pieces it replaces: pairing, int32 offsets, processing order, layout, and its refusals."""
    rng = np.random.default_rng(11)
    for batch, classes, want_order, knot, channels in ((64, 2, False, 4, 4), (257, 3, True, -1, 1), (1, 1, True, 2, 3), (4096, 2, True, 4, 4)):
        frames = synth.cycle_frames(rng, batch, limit=2500).astype(np.int64)
        wide = np.concatenate([frames, np.full((batch, 2), -7, np.int64)], axis=1)[:, :5]      # a strided (B, 5) view of (B, 7)
        labels = rng.integers(0, classes, batch).astype(np.int64)
        packed = torch.zeros(batch * 28 + max(0, batch * (knot + 2) * channels * 8) + 64, dtype=torch.uint8)
        for step in (0, 5, 2 ** 32 - 1, 2 ** 40 + 3):
            rc, info, mix = native.host_prepare_step(labels, wide, 2500, step, knot, channels, want_order, packed)
            assert rc == 0
            assert np.array_equal(mix, draws.same_label_pairing(labels, step))
            raw = packed.numpy()
            got_frames = raw[info[0]:info[0] + batch * 20].view(np.int32).reshape(batch, 5)
            assert np.array_equal(got_frames, frames)
            assert np.array_equal(raw[info[1]:info[1] + batch * 4].view(np.int32), mix)
            if want_order:
                assert np.array_equal(raw[info[2]:info[2] + batch * 4].view(np.int32), draws.processing_order(mix))
            else:
                assert info[2] == -1
            assert (info[3] >= 0) == (knot >= 0) and all(o % 16 == 0 for o in info[:5] if o >= 0)
            if knot >= 0:
                assert info[4] >= info[3] + batch * (knot + 2) * channels * 8
    # refusals
    frames = np.array([[0, 10, 20, 30, 40], [0, 10, 5, 30, 40]], np.int64)
    packed = torch.zeros(4096, dtype=torch.uint8)
    rc, info, _ = native.host_prepare_step(np.zeros(2, np.int64), frames, 50, 1, -1, 1, False, packed)
    assert rc == 2 and info[5] == 1
    frames = np.array([[0, 10, 20, 30, 70], [0, 10, 20, 35, 75]], np.int64)            # both run past a row of 50, unequal clamps
    rc, info, mix = native.host_prepare_step(np.zeros(2, np.int64), frames, 50, 0, -1, 1, False, packed)
    if mix[0] == 1:
        assert rc == 3 and info[6] == 3
    rc, info, _ = native.host_prepare_step(np.zeros(2, np.int64), frames, 50, 0, 4, 4, True, torch.zeros(64, dtype=torch.uint8))
    assert rc == 1 and info[4] > 64
