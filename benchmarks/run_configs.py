#!/usr/bin/env python
"""Device-side timing of every BASELINE.json configuration (bench.py covers only the headline
config 2 in the driver's contract).  One process, one GPU; prints one JSON line per config.

    python benchmarks/run_configs.py [--reps 200]

cfg1  32 recordings x 5 s @ 2 kHz, dense Springer states -> segmentation kernels -> cut + pad to
      4400 -> durratiomixup (tiny: reported as latency per call, not as a roofline fraction)
cfg2  durmixmagwarp(0.2,4) on 4096 x 4 x 2500 (the bench.py workload), plus PCGmix on the same batch
cfg3  2D durratiomixup on 1024 x 1 x 64 x 250
cfg4  1 M cycles = 245 batches of 4096, seed = batch index (single GPU here; shards by batch)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)

import bench  # noqa: E402  (the CPU-baseline legs live in bench.py: the only non-test code that may run oracle/)
from pcgmix_b200 import augmentations, draws, native, resident, segmentation, staging, synth  # noqa: E402

PEAK = 6544.7
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, reps, warm=10):
    for _ in range(warm):
        fn(0)
    torch.cuda.synchronize()
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    marks[0].record()
    for i in range(reps):
        fn(i)
        marks[i + 1].record()
    torch.cuda.synchronize()
    per = [marks[i].elapsed_time(marks[i + 1]) for i in range(reps)]
    return statistics.fmean(per), min(per)


RESULTS = []          # every line emitted in this process (bench.py collects them for its "configs" array)
QUIET = False


def emit(name, cycles, ms_mean, ms_min, bytes_per_call=None, **extra):
    line = {"config": name, "cycles_per_call": cycles, "ms_mean": ms_mean, "ms_min": ms_min,
            "cycles_per_s": cycles / (ms_mean * 1e-3)}
    if bytes_per_call is not None:
        gbs = bytes_per_call / (ms_mean * 1e-3) / 1e9
        line.update({"algorithmic_bytes": bytes_per_call, "achieved_GBps": gbs, "frac_of_measured_peak": gbs / PEAK})
    line.update(extra)
    RESULTS.append(line)
    if not QUIET:
        print(json.dumps(line), flush=True)


def prepared_steps(frames, labels, batch, channels, n_steps, magwarp, dev, nb):
    steps = []
    for s in range(n_steps):
        f, lab = frames[s % nb], labels[s % nb]
        mix = draws.same_label_pairing(lab, s)
        arrays = [f.astype(np.int32), mix.astype(np.int32), draws.processing_order(mix)]
        if magwarp:
            lam_f, knots = draws.lambda_and_knots(1.0, s, batch, 4, channels, 0.2)
            arrays.append(knots)
        else:
            lam_f = draws.draw_lambda(1, s)
        lam = draws.lambda_pair_fp32(lam_f)
        steps.append((staging.upload(arrays, dev), lam, synth.mixed_samples(f, mix)))
    return steps


def resident_section(args, dev):
    C, L = 4, 2500
    n_rec, Tr, Br = 512, 40000, 4096                            # 328 MB of recordings (>> L2), ~18 k cycles
    rr = np.random.default_rng(synth.BENCH_SEED + 77)
    st = torch.from_numpy(synth.dense_states(rr, n_rec, Tr, 1000)).to(dev)
    sig = torch.from_numpy(rr.standard_normal((n_rec, C, Tr)).astype(np.float32)).to(dev)
    res = resident.from_dense_states(sig, st, L)
    frames_all = res.table.cycles[: res.n_cycles, 3:].cpu().numpy().astype(np.int64)
    out_r = torch.empty((Br, C, L), dtype=torch.float32, device=dev)
    n_sets = 8
    sets = []
    for k in range(n_sets):
        ids = rr.integers(0, res.n_cycles, Br)
        mix = draws.same_label_pairing(rr.integers(0, 2, Br), k)
        lam = draws.lambda_pair_fp32(draws.draw_lambda(1, k))
        up = staging.upload([ids.astype(np.int32), mix.astype(np.int32), draws.draw_knots(Br, 4, C, 0.2),
                             draws.processing_order(mix)], dev)
        f = frames_all[ids]
        sets.append((up, lam, int(f[:, 4].clip(max=L).sum()), synth.mixed_samples(f, mix)))
    for magwarp, name in ((True, "durmixmagwarp(0.2,4)"), (False, "durratiomixup")):
        def fused(i, magwarp=magwarp):
            up, lam, _, _ = sets[i % n_sets]
            resident.mix_rows(res, up[0], up[1], lam[0], lam[1], up[2] if magwarp else None, 4, out=out_r,
                              order_dev=up[3] if getattr(args, "resident_order", False) else None)

        def two_step(i, magwarp=magwarp):
            up, lam, _, _ = sets[i % n_sets]
            rows = res.table.cycles[up[0].long()]
            sub = segmentation.CycleTable(rows, res.table.row_ptr, res.table.err_flag)
            padded = segmentation.cut_cycles(sig, sub, L, Br)
            augmentations.pcgmix_on_device(padded, rows[:, 3:], up[1], lam[0], lam[1], up[2] if magwarp else None, 4, out=out_r)
        own = statistics.fmean(s_[2] for s_ in sets)
        mm = statistics.fmean(s_[3] for s_ in sets)
        ms, mn = timed(fused, args.reps)
        emit("resident/%s fused cut+pad+mix from %d recordings x %d x %d, batch %d x %d" % (name, n_rec, C, Tr, Br, L),
             Br, ms, mn, 4.0 * C * (own + mm + L * Br),
             note="bytes = 4*C*(sum len1 + sum M + B*L); mean cycle %.0f samples of L=%d" % (own / Br, L))
        if magwarp and getattr(args, "resident_e2e", True):
            # end to end from the host's point of view: per step only the table rows of the batch (16 KB) and the
            # per-step draws go up; the augmented batch comes back into pinned host memory (bench.py's e2e loop
            # moves the 164 MB padded batch up as well)
            class _A:
                method, batch_size, sample_rate, num_classes = "durmixmagwarp(0.2,4)", Br, 1000, 2

            class _S:
                count = 0
            host_out = [torch.empty((Br, C, L), dtype=torch.float32).pin_memory() for _ in range(2)]
            ids_host = [rr.integers(0, res.n_cycles, Br) for _ in range(4)]
            ohe = [torch.nn.functional.one_hot(torch.from_numpy(rr.integers(0, 2, Br)), 2).to(dev) for _ in range(4)]
            s_out = torch.cuda.Stream(dev)
            done_out = [torch.cuda.Event() for _ in range(2)]
            cur = torch.cuda.current_stream(dev)

            def e2e_steps(n, download=True):
                for ev in done_out:
                    ev.record(cur)
                for i in range(n):
                    _S.count = 1000 + i
                    out, _, _, _ = resident.augment(_A, res, ids_host[i % 4], ohe[i % 4], None, _S, None, dev, None)
                    if not download:
                        continue
                    ready = torch.cuda.Event()
                    ready.record(cur)
                    with torch.cuda.stream(s_out):
                        s_out.wait_event(ready)
                        s_out.wait_event(done_out[i % 2])
                        host_out[i % 2].copy_(out, non_blocking=True)
                        out.record_stream(s_out)
                        done_out[i % 2].record(s_out)
                cur.wait_stream(s_out)
            import time
            e2e_steps(3)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e2e_steps(24)
            torch.cuda.synchronize()
            ms_e2e = (time.perf_counter() - t0) / 24 * 1e3
            t0 = time.perf_counter()
            e2e_steps(24, download=False)
            torch.cuda.synchronize()
            ms_host = (time.perf_counter() - t0) / 24 * 1e3
            emit("resident/durmixmagwarp(0.2,4) end to end: table rows + draws up, augmented batch down to pinned host memory",
                 Br, ms_e2e, ms_e2e, note="h2d per step: %d B of cycle ids + draws; d2h: %d B; the same loop without the download "
                 "(host draws + launches only): %.2f ms per step" % (Br * 4, Br * C * L * 4, ms_host))
        if args.only == "resident" and (args.stages or args.ctas_per_sm or args.consumer_threads or args.max_slice or args.pbuf_pct):
            continue
        ms2, mn2 = timed(two_step, 50)
        emit("resident/%s two-step: row gather + cut_cycles + mix kernel (what the fused kernel replaces)" % name,
             Br, ms2, mn2, note="fused is %.2fx faster" % (ms2 / ms))



def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=200)
    ap.add_argument("--only", default="", choices=["", "resident", "first_block"])
    ap.add_argument("--stages", type=int, default=0)
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    ap.add_argument("--consumer-threads", type=int, default=0)
    ap.add_argument("--max-slice", type=int, default=0)
    ap.add_argument("--pbuf-pct", type=int, default=0)
    ap.add_argument("--spline", default="float32", choices=["float32", "float64"])
    ap.add_argument("--resident-order", action="store_true",
                    help="resident section: process the slots in pairing-chain order (measured: no gain for PCGmix+, -9 %% for PCGmix; "
                         "the partner's samples come out of another recording either way)")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    native.load()
    native.set_spline_precision(args.spline)
    native.set_tuning(True, args.stages, args.max_slice, args.ctas_per_sm, args.pbuf_pct, args.consumer_threads, 0)
    run(args, dev)


def collect(dev, reps=60, cpu_legs=False, resident_e2e=False):
    """All sections, quietly, for bench.py's "configs" array: returns the list of result lines."""
    global QUIET
    ns = argparse.Namespace(reps=reps, only="", stages=0, ctas_per_sm=0, consumer_threads=0, max_slice=0, pbuf_pct=0,
                            cpu_legs=cpu_legs, resident_e2e=resident_e2e)
    QUIET, before = True, len(RESULTS)
    try:
        run(ns, dev)
    finally:
        QUIET = False
    return RESULTS[before:]


def run(args, dev):
    cpu_legs = getattr(args, "cpu_legs", True)
    if args.only == "resident":
        return resident_section(args, dev)
    if args.only == "first_block":
        return first_block_section(args, dev, cpu_legs)
    rng = np.random.default_rng(synth.BENCH_SEED)

    # ---------------- cfg1 ----------------
    n_rec, n_samp, bands, L1 = 32, 10000, 4, 4400
    states = torch.from_numpy(synth.dense_states(rng, n_rec, n_samp, 2000)).to(dev)
    signal = torch.from_numpy(rng.standard_normal((n_rec, bands, n_samp)).astype(np.float32)).to(dev)
    table = segmentation.cycles_from_dense_states(states).check()
    n_cyc = table.total()
    cycles = segmentation.cut_cycles(signal, table, L1, n_cyc)
    labels1 = rng.integers(0, 2, n_cyc)
    mix1 = torch.from_numpy(draws.same_label_pairing(labels1, 0).astype(np.int32)).to(dev)
    out1 = torch.empty_like(cycles)
    ms, mn = timed(lambda i: segmentation.cycles_from_dense_states(states), 50)
    emit("cfg1/segment_dense (32 x 10000 int8 -> cycle table, 3 launches)", n_cyc, ms, mn)
    ms, mn = timed(lambda i: segmentation.cut_cycles(signal, table, L1, n_cyc), 50)
    emit("cfg1/cut_cycles (-> %d x 4 x 4400)" % n_cyc, n_cyc, ms, mn)
    ms, mn = timed(lambda i: native.mix1d(cycles, out1, table.frames[:n_cyc], mix1, 0.3, 0.7), args.reps)
    emit("cfg1/durratiomixup (%d cycles x 4 x 4400, frames = cycle table view)" % n_cyc, n_cyc, ms, mn,
         note="latency-bound: %.1f MB per call" % (2 * cycles.numel() * 4 / 1e6))
    # CPU port of the same pipeline (reference structure: Python loops), single process
    if cpu_legs:
        t_seg, t_mix = bench.cpu_baseline_cfg1(states.cpu().numpy(), signal.cpu().numpy(), labels1, L1)
        emit("cfg1/CPU port: segmentation + cut (once) and durratiomixup per call", n_cyc, t_mix * 1e3, t_mix * 1e3,
             note="segmentation+cut %.1f ms; mix %.2f ms per call on the host (torch-CPU tensors like the reference)" % (t_seg * 1e3, t_mix * 1e3))
    # the same batch without the intermediate padded array: recordings + cycle table -> fused cut + pad + mix
    res1 = resident.ResidentCycles(signal, table, L1, n_cyc, torch.zeros(1, dtype=torch.int32, device=dev))
    scratch1 = torch.empty((n_cyc, 8), dtype=torch.int32, device=dev)
    ms, mn = timed(lambda i: resident.mix_rows(res1, None, mix1, 0.3, 0.7, out=out1, scratch=scratch1), args.reps)
    emit("cfg1/durratiomixup fused with cut + pad (pcgmix_mix1d_resident, 2 launches)", n_cyc, ms, mn,
         note="replaces cut_cycles + durratiomixup above")
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        native.mix1d(cycles, out1, table.frames[:n_cyc], mix1, 0.3, 0.7)
        with torch.cuda.graph(graph, stream=side):
            native.mix1d(cycles, out1, table.frames[:n_cyc], mix1, 0.3, 0.7)
    torch.cuda.synchronize()
    ms, mn = timed(lambda i: graph.replay(), args.reps)
    emit("cfg1/durratiomixup replayed from a CUDA graph", n_cyc, ms, mn)

    # ---------------- cfg2 / cfg4 ----------------
    B, C, L, NB = 4096, 4, 2500, 4
    frames, labels, data = [], [], []
    for i in range(NB):
        r = np.random.default_rng(synth.BENCH_SEED + i)
        f = synth.cycle_frames(r, B, limit=L)
        frames.append(f)
        data.append(torch.from_numpy(synth.cycle_signals(r, f, (C,), L)).to(dev))
        labels.append(r.integers(0, 2, B))
    outs = [torch.empty_like(data[0]) for _ in range(2)]
    for magwarp, name in ((True, "cfg2/durmixmagwarp(0.2,4) 4096 x 4 x 2500"), (False, "cfg2b/durratiomixup 4096 x 4 x 2500")):
        n_steps = args.reps + 10
        steps = prepared_steps(frames, labels, B, C, n_steps, magwarp, dev, NB)

        def run(i, steps=steps, magwarp=magwarp):
            up, lam, _ = steps[i % len(steps)]
            augmentations.pcgmix_on_device(data[i % NB], up[0], up[1], lam[0], lam[1], up[3] if magwarp else None, 4,
                                           order_dev=up[2], out=outs[i % 2])
        ms, mn = timed(run, args.reps)
        m_mean = statistics.fmean(s[2] for s in steps)
        emit(name, B, ms, mn, 4.0 * C * (2.0 * L * B + m_mean))
    n_b = 245
    import time
    t_host = time.perf_counter()
    steps = prepared_steps(frames, labels, B, C, n_b, True, dev, NB)
    torch.cuda.synchronize()
    host_ms = (time.perf_counter() - t_host) * 1e3             # pairing, order, lambda, knots and their upload for 245 batches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(n_b):
        up, lam, _ = steps[k]
        augmentations.pcgmix_on_device(data[k % NB], up[0], up[1], lam[0], lam[1], up[3], 4, order_dev=up[2], out=outs[k % 2])
    e1.record()
    torch.cuda.synchronize()
    tot = e0.elapsed_time(e1)
    emit("cfg4/1M cycles = 245 batches of 4096, PCGmix+, 1 GPU", n_b * B, tot, tot,
         4.0 * C * (2.0 * L * B * n_b + sum(s[2] for s in steps)),
         host_draws_ms_total=host_ms, wall_ms_including_host_draws=host_ms + tot,
         cycles_per_s_including_host_draws=n_b * B / ((host_ms + tot) * 1e-3),
         note="device time is the 245 launches between two events; the host's draws for the 245 batches (single thread, "
              "before the timed region) are reported beside it")

    # ---------------- cfg3 ----------------
    B3, F3, T3 = 1024, 64, 250
    frames3 = synth.spectrogram_frames(rng, B3, T3)
    data3 = torch.from_numpy(synth.cycle_signals(rng, frames3, (1, F3), T3)).to(dev)
    out3 = torch.empty_like(data3)
    labels3 = rng.integers(0, 2, B3)
    mix3 = draws.same_label_pairing(labels3, 0)
    up3 = staging.upload([frames3.astype(np.int32), mix3.astype(np.int32), draws.processing_order(mix3)], dev)
    m3 = synth.mixed_samples(frames3, mix3)
    ms, mn = timed(lambda i: native.mix2d(data3, out3, up3[0], up3[1], 0.3, 0.7, order=up3[2]), args.reps)
    emit("cfg3/2D durratiomixup 1024 x 1 x 64 x 250 (65.5 MB in)", B3, ms, mn, 4.0 * F3 * (2.0 * T3 * B3 + m3),
         note="working set 131 MB ~ L2 size: partly L2-resident across repetitions")
    n3 = 128                                                   # bounded CPU sample of the same workload
    if cpu_legs:
        t3 = bench.cpu_baseline_cfg3(data3[:n3].cpu(), labels3[:n3], frames3[:n3])
        emit("cfg3/CPU port on a %d-item sample (throughput of the per-item loop is flat in B)" % n3, n3, t3 * 1e3, t3 * 1e3)
    B3b = 8192
    frames3b = synth.spectrogram_frames(rng, B3b, T3)
    data3b = torch.from_numpy(synth.cycle_signals(rng, frames3b, (1, F3), T3)).to(dev)
    out3b = torch.empty_like(data3b)
    mix3b = draws.same_label_pairing(rng.integers(0, 2, B3b), 0)
    up3b = staging.upload([frames3b.astype(np.int32), mix3b.astype(np.int32), draws.processing_order(mix3b)], dev)
    ms, mn = timed(lambda i: native.mix2d(data3b, out3b, up3b[0], up3b[1], 0.3, 0.7, order=up3b[2]), 50)
    emit("cfg3x8/2D durratiomixup 8192 x 1 x 64 x 250 (524 MB in, >> L2)", B3b, ms, mn,
         4.0 * F3 * (2.0 * T3 * B3b + synth.mixed_samples(frames3b, mix3b)))
    # reference's real spectrogram shape
    B5, F5, T5 = 4096, 128, 128
    frames5 = synth.spectrogram_frames(rng, B5, T5, seconds=2.2)
    data5 = torch.from_numpy(synth.cycle_signals(rng, frames5, (1, F5), T5)).to(dev)
    out5 = torch.empty_like(data5)
    mix5 = draws.same_label_pairing(rng.integers(0, 2, B5), 0)
    up5 = staging.upload([frames5.astype(np.int32), mix5.astype(np.int32), draws.processing_order(mix5)], dev)
    ms, mn = timed(lambda i: native.mix2d(data5, out5, up5[0], up5[1], 0.3, 0.7, order=up5[2]), 50)
    emit("spec128/2D durratiomixup 4096 x 1 x 128 x 128 (268 MB in)", B5, ms, mn,
         4.0 * F5 * (2.0 * T5 * B5 + synth.mixed_samples(frames5, mix5)))
    resident_section(args, dev)
    features_section(args, dev, cpu_legs)
    first_block_section(args, dev, cpu_legs)


def features_section(args, dev, cpu_legs):
    """The consumer side (train_model.py:519-532): classical features of an augmented batch, computed on the device."""
    from pcgmix_b200 import features
    rng = np.random.default_rng(synth.BENCH_SEED + 5)
    B, C, L = 4096, 5, 2500                                    # the reference reads channel 4 of a five-row cycle
    frames = synth.cycle_frames(rng, B, limit=L)
    data = torch.from_numpy(synth.cycle_signals(rng, frames, (C,), L)).to(dev)
    frames_dev = staging.upload([frames.astype(np.int32)], dev)[0]
    out36 = torch.empty((B, 36), dtype=torch.float32, device=dev)
    out80 = torch.empty((B, 80), dtype=torch.float32, device=dev)
    legs = (("amplitude + Hilbert-envelope block (36 values)", lambda i: features.cycle_features(data, frames_dev, 4, out=out36)),
            ("Welch-PSD block (80 values)", lambda i: features.cycle_psd_features(data, frames_dev, 4, out=out80)),
            ("duration block (14 values)", lambda i: segmentation.duration_features(frames_dev, 1000)))
    for name, fn in legs:
        ms, mn = timed(fn, 20, warm=3)
        emit("features/%s of 4096 cycles x 2500, one channel" % name, B, ms, mn)
    if cpu_legs:
        t_env, t_psd = bench.cpu_baseline_features(data[:64].cpu().numpy(), frames[:64], 4)
        emit("features/CPU port (NumPy / SciPy calls of the reference, one core) on a 64-cycle sample: amplitude + envelope", 64,
             t_env * 1e3, t_env * 1e3)
        emit("features/CPU port on a 64-cycle sample: Welch-PSD block", 64, t_psd * 1e3, t_psd * 1e3)


def first_block_section(args, dev, cpu_legs):
    """The model side (train_model.py:536 -> models.py:538): conv1 = Conv1d(4, 64, 3, padding=1) + BatchNorm1d + ReLU of the
    reference's ResNet9-1D on an augmented batch.  Bound by writing the 16x larger output; bytes = input read (twice in
    training mode: patch moments, then the pass that writes) + output written."""
    from pcgmix_b200 import first_block
    rng = np.random.default_rng(synth.BENCH_SEED + 6)
    C, L, F = 4, 2500, 64
    torch.manual_seed(6)
    block = torch.nn.Sequential(torch.nn.Conv1d(C, F, 3, padding=1), torch.nn.BatchNorm1d(F), torch.nn.ReLU(inplace=True)).to(dev)
    frames = synth.cycle_frames(rng, 512, limit=L)
    base = torch.from_numpy(synth.cycle_signals(rng, frames, (C,), L)).to(dev)
    saved_tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False                   # the library leg beside ours must be a float32 one too
    for B in (4096, 64):
        data = base.repeat(B // 512, 1, 1).contiguous() if B >= 512 else base[:B].contiguous()
        out = torch.empty((B, F, L), dtype=torch.float32, device=dev)
        for training in (True, False):
            block.train(training)
            ms, mn = timed(lambda i: first_block.first_conv_block(block, data, out=out), 20, warm=3)
            emit("first block/conv1 + BatchNorm + ReLU (%s statistics) of %d cycles x 4 x 2500 -> 64 x 2500" %
                 ("batch" if training else "running", B), B, ms, mn, 4.0 * B * L * (C * (2 if training else 1) + F))
            with torch.no_grad():
                ms, mn = timed(lambda i: block(data), 10, warm=3)
            emit("first block/the same modules run by torch (cuDNN, float32; library, beside ours), %s statistics, %d cycles" %
                 ("batch" if training else "running", B), B, ms, mn, 4.0 * B * L * (C * (2 if training else 1) + F))
        del out
    torch.backends.cudnn.allow_tf32 = saved_tf32
    if cpu_legs:
        import time
        cpu_block = torch.nn.Sequential(torch.nn.Conv1d(C, F, 3, padding=1), torch.nn.BatchNorm1d(F), torch.nn.ReLU(inplace=True)).train()
        sample = base[:64].cpu()
        with torch.no_grad():
            cpu_block(sample)
            t0 = time.perf_counter()
            for _ in range(5):
                cpu_block(sample)
            dt = (time.perf_counter() - t0) / 5
        emit("first block/CPU: the reference's modules (torch, %d threads) on a 64-cycle batch, batch statistics" % torch.get_num_threads(),
             64, dt * 1e3, dt * 1e3)


if __name__ == "__main__":
    main()
