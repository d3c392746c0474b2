#!/usr/bin/env python
"""Benchmark of the PCGmix+ hot path (BASELINE.json: augmented cardiac cycles/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A *step* is one pass of the hot path over one batch of synthetic cycles.  Workload at every N:
BASELINE config 2 — ``durmixmagwarp(0.2,4)`` (PCGmix+, fused mix + magnitude warp) on a batch of
4096 cycles x 4 channels x 2500 samples per GPU (weak scaling: every rank owns its own batches,
pairing is drawn inside each batch exactly like the reference does, no collective on the
augmentation path).

Numbers on the JSON line
  value ........ whole-job cycles/s with the batches already resident in HBM and the per-step
                 draws (frames, pairing, order, knots) already uploaded: K kernel launches
                 between two CUDA events on the launching stream, max over ranks
  roofline ..... algorithmic bytes per launch 4*C*(2*L*B + sum_b M_b) over the mean launch time,
                 against the measured HBM copy peak in MEASURED_PEAKS.json
  e2e .......... the same metric through the public ``augmentations.augment`` call with HOST
                 buffers: pinned host batch -> device, host draws, kernel, result -> pinned host,
                 all inside the timed region
  cpu_baseline . the CPU oracle (a port of the reference's Python loop + SciPy splines) timed on
                 this box's host cores on a bounded sample of the same workload

``--impl reference`` times only that CPU port (the reference itself is Python and is not
available on the GPU box), spread over all host cores, and prints the same JSON shape.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "augmented_cardiac_cycles_per_sec"
UNIT = "cycles/s"
METHOD = "durmixmagwarp(0.2,4)"
WORKLOAD = "cfg2: durmixmagwarp(0.2,4) PCGmix+ on 4096 cycles x 4 ch x 2500 samples (fp32) per GPU"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--channels", type=int, default=4)
    ap.add_argument("--length", type=int, default=2500)
    ap.add_argument("--method", default=METHOD)
    ap.add_argument("--resident-batches", type=int, default=4, help="distinct input batches kept in HBM per rank")
    ap.add_argument("--e2e-steps", type=int, default=24)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-order", action="store_true", help="visit cycles in index order instead of pairing-chain order")
    ap.add_argument("--kernel", default="pipeline", choices=["pipeline", "direct"],
                    help="pipeline: persistent TMA-pipelined kernel (default); direct: direct-load kernel")
    ap.add_argument("--stages", type=int, default=0)
    ap.add_argument("--max-slice", type=int, default=0)
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    ap.add_argument("--pbuf-pct", type=int, default=0)
    ap.add_argument("--consumer-threads", type=int, default=0)
    ap.add_argument("--no-launch-overlap", action="store_true",
                    help="serialise consecutive launches (default: consecutive independent steps may overlap their "
                         "pipeline fill/drain through programmatic dependent launch)")
    ap.add_argument("--no-graph", action="store_true",
                    help="issue the K timed launches from Python instead of replaying them from one CUDA graph")
    ap.add_argument("--debug-skip", type=int, default=0,
                    help="profiling build only (PCGMIX_PROFILING_LIB=1): 1 no stores, 2 no arithmetic, 4 no partner staging; "
                         "outputs are wrong, so this implies --no-verify and is recorded in config")
    ap.add_argument("--no-verify", action="store_true", help="do not check the timed outputs against the oracle")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------
# synthetic workload
# ----------------------------------------------------------------------------------------------
def make_batch(seed: int, batch: int, channels: int, length: int):
    from pcgmix_b200 import synth
    rng = np.random.default_rng(seed)
    frames = synth.cycle_frames(rng, batch, fs=1000, limit=length)
    data = synth.cycle_signals(rng, frames, (channels,), length)
    labels = rng.integers(0, 2, batch)
    return data, frames, labels


class _Args:
    def __init__(self, method, batch):
        self.method, self.batch_size, self.sample_rate, self.num_classes = method, batch, 1000, 2


class _Step:
    def __init__(self, count):
        self.count = count


# ----------------------------------------------------------------------------------------------
# CPU baseline (oracle port), optionally spread over processes
# ----------------------------------------------------------------------------------------------
def _cpu_worker(payload):
    """Run the oracle's per-cycle loop (mix + SciPy magnitude warp) on a slice of a batch."""
    x1, x2, f1, f2, lam32, knots = payload
    from oracle import pcgmix_oracle as orc
    out = np.zeros_like(x1)
    lam = np.float32(lam32)
    for i in range(x1.shape[0]):
        out[i] = orc.mix_pair(x1[i], x2[i], f1[i], f2[i], lam)
    if knots is not None:
        out = np.transpose(orc.magnitude_warp(np.transpose(out, (0, 2, 1)), knots), (0, 2, 1))
    return out


def cpu_reference_step(method, data, frames, labels, step, pool, workers):
    """One step of the CPU port on host arrays; draws exactly as the reference; the per-cycle
    work is split over ``workers`` processes when a pool is given."""
    from oracle import pcgmix_oracle as orc
    if pool is None:
        out, _, _, _ = orc.augment_1d(method, data, labels, frames, step)
        return out
    mix = orc.same_label_mix_indices(labels, step)
    lam32 = orc.lambda_as_float32(orc.draw_lambda(orc.parse_alpha(method, "durmixmagwarp"), step))
    knots = None
    if "durmixmagwarp" in method:
        sigma, knot = orc.parse_magwarp(method)
        knots = orc.draw_knots(data.shape[0], knot, data.shape[1], sigma)
    partners = data[mix]
    pframes = frames[mix]
    bounds = np.linspace(0, data.shape[0], workers + 1).astype(int)
    jobs = [(data[a:b], partners[a:b], frames[a:b], pframes[a:b], lam32, None if knots is None else knots[a:b])
            for a, b in zip(bounds[:-1], bounds[1:]) if b > a]
    return np.concatenate(list(pool.map(_cpu_worker, jobs)), axis=0)


def time_cpu_baseline(method, batch, channels, length, seconds, workers, seed=7):
    """Cycles/s of the CPU port on a bounded sample: batches of ``batch`` cycles until ``seconds``."""
    from concurrent.futures import ProcessPoolExecutor
    import multiprocessing as mp
    data, frames, labels = make_batch(seed, batch, channels, length)
    pool = ProcessPoolExecutor(max_workers=workers, mp_context=mp.get_context("spawn")) if workers > 1 else None
    try:
        if pool is not None:
            list(pool.map(_cpu_worker, [(data[:2], data[:2], frames[:2], frames[:2], 0.5, None)] * workers))  # start workers
        cpu_reference_step(method, data[: max(8, workers)], frames[: max(8, workers)], labels[: max(8, workers)], 0, pool, workers)
        done, t0, per_step = 0, time.perf_counter(), []
        step = 1
        while True:
            t1 = time.perf_counter()
            cpu_reference_step(method, data, frames, labels, step, pool, workers)
            per_step.append(time.perf_counter() - t1)
            done += batch
            step += 1
            if time.perf_counter() - t0 >= seconds:
                break
        elapsed = time.perf_counter() - t0
    finally:
        if pool is not None:
            pool.shutdown()
    return done / elapsed, done, elapsed, per_step


def cpu_baseline_cfg1(states, signal, labels, length, reps=5):
    """CPU-baseline leg of BASELINE config 1 (used by benchmarks/run_configs.py): the oracle's
    segmentation + cut over dense states, then ``durratiomixup`` per call.  Returns seconds
    (segmentation + cut once, mix per call)."""
    import torch
    from oracle import pcgmix_oracle as orc
    from oracle import segmentation_oracle as seg_orc
    t0 = time.perf_counter()
    cycles, frames = [], []
    for r in range(states.shape[0]):
        rel, a0, a1 = seg_orc.cycles_from_dense(states[r])
        for i in range(len(a0)):
            cycles.append(np.stack([seg_orc.cut_and_pad(signal[r, c], a0[i], a1[i], length) for c in range(signal.shape[1])]))
            frames.append(rel[i])
    t_seg = time.perf_counter() - t0
    cycles, frames = torch.from_numpy(np.stack(cycles)), torch.from_numpy(np.stack(frames))
    t0 = time.perf_counter()
    for rep in range(reps):
        orc.augment_1d("durratiomixup", cycles, labels, frames, rep)
    return t_seg, (time.perf_counter() - t0) / reps


def cpu_baseline_cfg3(data, labels, frames):
    """CPU-baseline leg of BASELINE config 3 on a bounded sample: the oracle's 2D per-item loop."""
    import torch
    from oracle import pcgmix_oracle as orc
    t0 = time.perf_counter()
    orc.augment_2d("durratiomixup", data, labels, torch.from_numpy(np.asarray(frames)), 0)
    return time.perf_counter() - t0


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t_begin=None, t_end=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for ts, line in self.lines:
            if t_begin is not None and not (t_begin - 0.05 <= ts <= t_end + 0.15):
                continue
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons), "samples": len(sm)}


def recorded_traffic(default_workload: bool):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/),
    only for the workload that capture was taken on; otherwise None."""
    if not default_workload:
        return None, None
    best = None
    pdir = os.path.join(ROOT, "profiles")
    try:
        for name in sorted(os.listdir(pdir)):
            if name.endswith("_traffic.json"):
                with open(os.path.join(pdir, name)) as f:
                    best = (json.load(f), name)
    except Exception:
        return None, None
    if best is None:
        return None, None
    return float(best[0]["dram_traffic_bytes_per_launch"]), f"profiles/{best[1]}"


def bind_to_gpu_numa_node(gpu_index: int):
    """Multi-rank runs: keep this rank's host threads (and therefore its first-touched pinned buffers)
    on the CPU cores NVML reports as local to its GPU.  Best effort; silently skipped if NVML or
    sched_setaffinity is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------------------------
# reference arm
# ----------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workers = os.cpu_count() or 1
    sample = min(args.batch, 1024)
    data, frames, labels = make_batch(7, sample, args.channels, args.length)
    from concurrent.futures import ProcessPoolExecutor
    import multiprocessing as mp
    pool = ProcessPoolExecutor(max_workers=workers, mp_context=mp.get_context("spawn")) if workers > 1 else None
    try:
        if pool is not None:
            list(pool.map(_cpu_worker, [(data[:2], data[:2], frames[:2], frames[:2], 0.5, None)] * workers))
        for w in range(args.warmup):
            cpu_reference_step(args.method, data, frames, labels, w, pool, workers)
        t0 = time.perf_counter()
        for k in range(args.steps):
            cpu_reference_step(args.method, data, frames, labels, args.warmup + k, pool, workers)
        elapsed = time.perf_counter() - t0
    finally:
        if pool is not None:
            pool.shutdown()
    value = sample * args.steps / elapsed
    sample_txt = (f"{args.steps} steps x {sample} cycles x {args.channels} ch x {args.length} samples of the same "
                  f"workload (the full step is {args.batch} cycles; throughput of the per-cycle loop is flat in B)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 mix, f64 spline",
        "data": "synthetic", "config": {"workload": WORKLOAD, "method": args.method, "sample": sample_txt},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample_txt},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    from pcgmix_b200 import augmentations, draws, native, spline, staging, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU port")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    local_cpus = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    # Anything a library prints to stdout (NCCL announces its version there) goes to stderr: stdout
    # carries exactly one JSON line.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    native.load()
    if os.environ.get("PCGMIX_SPLINE"):
        native.set_spline_precision(os.environ["PCGMIX_SPLINE"])
    native.set_tuning(args.kernel == "pipeline", args.stages, args.max_slice, args.ctas_per_sm, args.pbuf_pct,
                      args.consumer_threads, args.debug_skip)

    B, C, L = args.batch, args.channels, args.length
    K, W = args.steps, args.warmup
    plan = draws.parse_method_1d(args.method)
    magwarp = plan.branch == "durmixmagwarp"

    # ---- resident inputs: NB distinct batches per rank (>> L2), per-step draws pre-uploaded -----
    NB = max(1, args.resident_batches)
    batches = []
    for i in range(NB):
        data, frames, labels = make_batch(synth.BENCH_SEED + 1000 * rank + i, B, C, L)
        batches.append((data, frames, labels))
    dev_data = [torch.from_numpy(b[0]).to(dev) for b in batches]
    NOUT = 3                                                   # output buffers in rotation
    outs = [torch.empty_like(dev_data[0]) for _ in range(NOUT)]
    steps_meta = []
    for s in range(W + K):
        data, frames, labels = batches[s % NB]
        seed = rank * (W + K) + s                       # the "training step" of this batch
        mix = draws.same_label_pairing(labels, seed)
        lam32, oml = draws.lambda_pair_fp32(draws.draw_lambda(plan.alpha, seed))
        arrays = [frames.astype(np.int32), mix.astype(np.int32),
                  np.arange(B, dtype=np.int32) if args.no_order else draws.processing_order(mix)]
        if magwarp:
            arrays.append(draws.draw_knots(B, plan.knot, C, plan.sigma))
        on_dev = staging.upload(arrays, dev)
        steps_meta.append({"dev": on_dev, "lam": (lam32, oml), "M": synth.mixed_samples(frames, mix)})
    torch.cuda.synchronize()

    # every step's launch is resolved once (pointers, sizes), so that issuing it is one foreign call
    prepared = []
    for s in range(W + K):
        m = steps_meta[s]
        prepared.append(augmentations.prepare_on_device(
            dev_data[s % NB], m["dev"][0], m["dev"][1], m["lam"][0], m["lam"][1], outs[s % NOUT],
            m["dev"][3] if magwarp else None, plan.knot, order_dev=m["dev"][2]))
    stream_handle = torch.cuda.current_stream(dev).cuda_stream

    def launch(s):
        prepared[s].launch(stream_handle)

    # consecutive steps are independent (distinct input batches, two alternating output buffers): let
    # launch k+1 fill its pipeline while launch k drains; the library re-checks buffer disjointness
    native.set_launch_overlap(not args.no_launch_overlap)
    for s in range(W):
        launch(s)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    stream = torch.cuda.current_stream(dev)
    # With launch overlap, nothing may sit between two launches on the stream (an event record would
    # re-serialise them), so only the two bracketing events are recorded and the per-launch time is
    # total / K.  With --no-launch-overlap every launch is bracketed by its own pair of events.
    per_launch_events = args.no_launch_overlap
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1 if per_launch_events else 2)]
    # The K launches are captured once into a CUDA graph (K kernel nodes, programmatic edges between
    # them) and replayed: the host issues nothing inside the timed region, so ranks do not drift apart
    # with host scheduling noise.  --no-graph / --no-launch-overlap launch from Python instead.
    graph = None
    if not per_launch_events and not args.no_graph:
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                capture_handle = torch.cuda.current_stream(dev).cuda_stream
                for k in range(K):
                    prepared[W + k].launch(capture_handle)
            graph.replay()                                      # untimed: instantiate + upload
            torch.cuda.synchronize()
        except Exception as exc:                                # capture unsupported: fall back to direct launches
            print(f"bench: CUDA graph capture failed ({exc!r}); launching directly", file=sys.stderr)
            graph = None
            torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches_before = native.launch_count
    overlap_before = native.overlap_launches()
    t_begin = time.perf_counter()
    marks[0].record(stream)
    if graph is not None:
        graph.replay()
    else:
        for k in range(K):
            launch(W + k)
            if per_launch_events:
                marks[k + 1].record(stream)
    if not per_launch_events:
        marks[1].record(stream)
    torch.cuda.synchronize()
    t_end = time.perf_counter()
    gpu_launches = K if graph is not None else native.launch_count - launches_before
    overlapped = (K - 1 if graph is not None else native.overlap_launches() - overlap_before)
    native.set_launch_overlap(False)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    total_ms = marks[0].elapsed_time(marks[-1])
    per_launch_ms = ([marks[k].elapsed_time(marks[k + 1]) for k in range(K)] if per_launch_events else [total_ms / K])
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None

    # the same K steps once more with ordinary stream serialisation and an event pair around every
    # launch: per-launch statistics, reported next to the headline for comparison
    serial_ms = per_launch_ms
    if not per_launch_events:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
        ev[0].record(stream)
        for k in range(K):
            launch(W + k)
            ev[k + 1].record(stream)
        torch.cuda.synchronize()
        serial_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(K)]

    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    max_ms = float(t.item())
    value = world * K * B / (max_ms * 1e-3)

    # ---- roofline of the fused kernel (rank 0's launches) ---------------------------------------
    bytes_per_launch = [4.0 * C * (2.0 * L * B + steps_meta[W + k]["M"]) for k in range(K)]
    mean_launch_s = statistics.fmean(per_launch_ms) * 1e-3
    achieved = statistics.fmean(bytes_per_launch) / mean_launch_s / 1e9
    peak, peak_src = measured_peak()
    is_default = (B, C, L, args.method) == (4096, 4, 2500, METHOD) and args.kernel == "pipeline"
    traffic, traffic_src = recorded_traffic(is_default)

    # ---- end to end through the public augment() with host buffers ------------------------------
    E = max(3, min(args.e2e_steps, K))
    host_in = [torch.from_numpy(b[0]).pin_memory() for b in batches[: min(NB, 2)]]
    host_out = [torch.empty_like(host_in[0]).pin_memory() for _ in range(2)]
    frames_t = [torch.from_numpy(b[1]) for b in batches[: min(NB, 2)]]
    ohe_t = [torch.nn.functional.one_hot(torch.from_numpy(b[2]), 2).to(dev) for b in batches[: min(NB, 2)]]
    wav = ["a0001"] * B
    a = _Args(args.method, B)

    # The loop around augment() is what a prefetching data loader does: the next batch is copied
    # host->device on a side stream while the current one is augmented, and results leave on a third
    # stream, so the two PCIe directions overlap.  Every step still moves its own input and output.
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    NIN = 3                                                    # input buffers: copies run two steps ahead
    dev_in = [torch.empty_like(dev_data[0]) for _ in range(NIN)]
    in_ready = [torch.cuda.Event() for _ in range(NIN)]
    in_free = [torch.cuda.Event() for _ in range(NIN)]
    out_done = [torch.cuda.Event() for _ in range(2)]

    def stage_in(i):
        with torch.cuda.stream(s_in):
            s_in.wait_event(in_free[i % NIN])                  # the augment that read this buffer is done
            dev_in[i % NIN].copy_(host_in[i % len(host_in)], non_blocking=True)
            in_ready[i % NIN].record(s_in)

    host_s = [0.0]                                             # wall time spent inside augment() (draws + launch)

    def run_e2e(n, seed0):
        for ev in in_free + out_done:
            ev.record(stream)
        stage_in(0)
        if n > 1:
            stage_in(1)
        for i in range(n):
            j = i % len(host_in)
            if i + 2 < n:
                stage_in(i + 2)
            stream.wait_event(in_ready[i % NIN])
            h0 = time.perf_counter()
            out, _, _, _ = augmentations.augment(a, dev_in[i % NIN], ohe_t[j], frames_t[j], wav, _Step(seed0 + i), None, dev, None)
            host_s[0] += time.perf_counter() - h0
            in_free[i % NIN].record(stream)
            done = torch.cuda.Event()
            done.record(stream)
            with torch.cuda.stream(s_out):
                s_out.wait_event(done)
                s_out.wait_event(out_done[i % 2])              # host_out slot is free again
                host_out[i % 2].copy_(out, non_blocking=True)
                out.record_stream(s_out)
                out_done[i % 2].record(s_out)
        stream.wait_stream(s_out)
        stream.wait_stream(s_in)

    run_e2e(3, 10_000)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches_e2e0 = native.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.perf_counter()
    e0.record(stream)
    host_s[0] = 0.0
    run_e2e(E, 20_000 + rank * E)
    e1.record(stream)
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - wall0) * 1e3
    e2e_ms = max(e0.elapsed_time(e1), wall_ms)          # host work is part of the step: take the longer clock
    te = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * E * B / (float(te.item()) * 1e-3)
    e2e_launches = native.launch_count - launches_e2e0
    # the host share of a step on its own (no GPU involved): the draws augment() makes for one batch
    t_h = time.perf_counter()
    for rep in range(5):
        m_ = draws.pairing(args.method, batches[0][2], wav, 30_000 + rep)
        draws.lambda_pair_fp32(draws.draw_lambda(plan.alpha, 30_000 + rep))
        if magwarp:
            draws.draw_knots(B, plan.knot, C, plan.sigma)
        m_.astype(np.int32)
    host_draws_ms = (time.perf_counter() - t_h) / 5 * 1e3
    in_bytes = B * C * L * 4
    small_bytes = B * 5 * 4 + B * 4 * 2 + (B * (plan.knot + 2) * C * 8 if magwarp else 0)

    # ---- CPU baseline on rank 0 at N=1 ----------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        workers = os.cpu_count() or 1
        sample_b = 1024
        v, done, elapsed, _ = time_cpu_baseline(args.method, sample_b, C, L, args.cpu_seconds, workers)
        cpu = {"value": v, "unit": UNIT, "cores": workers, "kind": "port",
               "sample": f"{done} cycles ({done // sample_b} steps of {sample_b} x {C} ch x {L}) in {elapsed:.1f} s; "
                         f"oracle port of the reference loop, per-cycle work spread over {workers} processes"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": max_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 mix, f64 spline", "data": "synthetic",
            "config": {"workload": WORKLOAD if (B, C, L, args.method) == (4096, 4, 2500, METHOD) else
                       f"{args.method} on {B} cycles x {C} ch x {L} samples per GPU",
                       "method": args.method, "cycles_per_step_per_gpu": B, "channels": C, "samples": L,
                       "resident_input_batches": NB,
                       "l2_policy": f"inputs larger than L2: {NB} x {in_bytes / 1e6:.0f} MB input batches + 3 output "
                                    "buffers rotate, every step reads/writes ~330 MB",
                       "cycle_order": "index" if args.no_order else "pairing-chain",
                       "launch_overlap": "off" if args.no_launch_overlap else
                       "programmatic dependent launch between consecutive independent steps (buffers checked disjoint)",
                       "launch_mode": "one CUDA graph of K kernel nodes, replayed" if graph is not None else "K launches from Python",
                       "kernel": args.kernel, "stages": args.stages, "max_slice": args.max_slice, "ctas_per_sm": args.ctas_per_sm,
                       "sharding": "batches per rank, pairing inside each batch, no collective",
                       "host_affinity": f"rank bound to the {local_cpus} CPU cores local to its GPU" if local_cpus else "default"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": statistics.fmean(bytes_per_launch),
                         "kernel_ms_mean": statistics.fmean(per_launch_ms),
                         "serialized_launches": {
                             "kernel_ms_mean": statistics.fmean(serial_ms), "kernel_ms_median": statistics.median(serial_ms),
                             "kernel_ms_min": min(serial_ms),
                             "achieved": statistics.fmean(bytes_per_launch) / (statistics.fmean(serial_ms) * 1e-3) / 1e9,
                             "frac": statistics.fmean(bytes_per_launch) / (statistics.fmean(serial_ms) * 1e-3) / 1e9 / peak,
                             "cycles_per_s_per_gpu": B / (statistics.fmean(serial_ms) * 1e-3)}},
            "e2e": {"value": e2e_value, "unit": UNIT, "steps": E, "ms_per_step": float(te.item()) / E,
                    "h2d_bytes_per_step": in_bytes + small_bytes, "d2h_bytes_per_step": in_bytes + B * 8,
                    "host_ms_per_step_inside_augment": 1e3 * host_s[0] / E,      # includes waiting for the batch's H2D copy
                    "host_draws_ms_per_step": host_draws_ms,                     # pairing + lambda + knots alone, rank 0
                    "api": "pcgmix_b200.augmentations.augment (host draws + 1 kernel) inside a prefetching loop: pinned host in/out, "
                           "H2D of steps k+1, k+2 and D2H of step k-1 on side streams"},
            "gpu_launches": gpu_launches, "gpu_launches_overlapped": overlapped, "gpu_launches_e2e": e2e_launches,
            "clocks": clocks,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
