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
                 against the measured HBM copy peak in MEASURED_PEAKS.json; ``serialized_launches``
                 is the same kernel with ordinary stream order and an event pair per launch;
                 ``dram`` relates the ncu-measured DRAM traffic of a launch to the same times
  verified ..... sampled cycles of the LAST timed outputs (device leg and e2e leg) recomputed by the
                 CPU oracle; the run exits non-zero if they differ by more than 1e-5 relative
  e2e .......... the same metric through the public ``augmentations.augment`` call with HOST
                 buffers: pinned host batch -> device, host draws, kernel, result -> pinned host,
                 all inside the timed region.  ``e2e_variants`` holds the same loop with fewer bytes
                 on the PCIe link (what a training loop really moves): the result consumed on the
                 device, and batches drawn from recordings that stay on the device
  configs ...... device-timed legs of the other BASELINE configurations (cfg1, cfg3, cfg4, the resident
                 path) and of the consumers of the augmented batch (classical features, the model's first
                 block); cfg5 ..... the DDP training step fed by the on-device augmentation
  cpu_baseline . the reference's own ``augment`` (byte-compiled under oracle/_ref) or, without it, the
                 CPU oracle port, timed on this box's host cores on a bounded sample of the workload

``--impl reference`` times only that CPU implementation, one process per host core, each running
whole steps of the unmodified reference on its own batches, and prints the same JSON shape.
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
    ap.add_argument("--no-configs", action="store_true", help="skip the device-timed legs of the other BASELINE configs")
    ap.add_argument("--no-variants", action="store_true", help="skip the e2e variants (device-side consumer, resident recordings)")
    ap.add_argument("--no-cfg5", action="store_true", help="skip the DDP training-step leg (BASELINE config 5)")
    ap.add_argument("--cfg5-steps", type=int, default=40)
    ap.add_argument("--spline", default=os.environ.get("PCGMIX_SPLINE", "float32"), choices=["float32", "float64"],
                    help="how the pipelined kernel evaluates the warp factor (float32: library default, <= 1e-5 relative; "
                         "float64: bit-faithful)")
    ap.add_argument("--reference-sample", type=int, default=256, help="cycles per step of a reference-arm process")
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


def cpu_baseline_features(data, frames, channel):
    """Seconds the feature oracle (the reference's NumPy / SciPy calls, one core) takes for the amplitude + envelope
    blocks and for the PSD block of ``data[:, channel]``."""
    import warnings
    from oracle import features_oracle as forc
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        t0 = time.perf_counter()
        forc.batch_features(data, frames, channel)
        t1 = time.perf_counter()
        forc.batch_psd_features(data, frames, channel)
        t2 = time.perf_counter()
    return t1 - t0, t2 - t1


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
def _reference_worker(conn, method, sample, channels, length, seed, steps, warmup, threads):
    """One process of the reference arm: the UNMODIFIED reference ``augment`` (oracle/_ref, sourceless) on this
    process's own batch, CPU tensors, ``threads`` torch threads; reports the seconds its ``steps`` timed steps took."""
    import torch
    torch.set_num_threads(threads)
    from oracle.ref_import import load_reference
    ref1, _ = load_reference(compiled=True)
    data, frames, labels = make_batch(seed, sample, channels, length)
    d = torch.from_numpy(data)
    ohe = torch.nn.functional.one_hot(torch.from_numpy(labels), 2)
    f = torch.from_numpy(frames)
    wav = ["a0001"] * sample
    a = _Args(method, sample)
    for w in range(warmup):
        ref1.augment(a, d, ohe, f, wav, _Step(w), None, "cpu", None)
    conn.send("ready")
    conn.recv()                                                # all processes start their timed steps together
    t0 = time.perf_counter()
    for k in range(steps):
        ref1.augment(a, d, ohe, f, wav, _Step(warmup + k), None, "cpu", None)
    conn.send(time.perf_counter() - t0)
    conn.close()


def time_reference(method, sample, channels, length, steps, warmup, workers, threads=1):
    """Cycles/s of the unmodified reference on this box: ``workers`` processes with ``threads`` torch threads each,
    every process running whole steps of ``sample`` cycles on its own batch.  ``workers = 1, threads = all cores`` is
    the reference as it is (one Python process; SURVEY section 8d's CPU timing); ``workers = cores, threads = 1`` shards
    independent batches over the cores the way this repo's ranks shard them over GPUs (a stronger baseline than the
    reference itself offers).  Returns (cycles/s, seconds of the slowest process)."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    procs, pipes = [], []
    for w in range(workers):
        parent, child = ctx.Pipe()
        p = ctx.Process(target=_reference_worker, args=(child, method, sample, channels, length, 1000 + w, steps, warmup, threads))
        p.start()
        procs.append(p)
        pipes.append(parent)
    for c in pipes:
        c.recv()
    for c in pipes:
        c.send("go")
    secs = [c.recv() for c in pipes]
    for p in procs:
        p.join()
    return workers * steps * sample / max(secs), max(secs)


def reference_kind():
    from oracle.ref_import import compiled_reference_available
    return "reference" if compiled_reference_available() else "port"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workers = os.cpu_count() or 1
    kind = reference_kind()
    extra = {}
    if kind == "reference":
        sample = max(8, min(args.batch, args.reference_sample))
        value, elapsed = time_reference(args.method, sample, args.channels, args.length, args.steps, args.warmup, 1, workers)
        sample_txt = (f"{args.steps} steps x {sample} cycles x {args.channels} ch x {args.length} samples: the unmodified reference "
                      f"augment() (oracle/_ref, byte-compiled from the reference tree) as it is — one Python process, CPU tensors, "
                      f"torch.set_num_threads({workers}) (SURVEY 8d) — on a bounded sample of the step (the full step is {args.batch} "
                      f"cycles; the reference's per-cycle loop is flat in B)")
        ms_per_step = 1e3 * elapsed / args.steps
        # beside it: independent batches sharded over the cores, one single-threaded reference process per core
        v_all, _ = time_reference(args.method, sample, args.channels, args.length, max(1, min(args.steps, 3)), 1, workers, 1)
        extra = {"one_process_per_core": {"value": v_all, "unit": UNIT, "processes": workers,
                                          "what": "the same unmodified augment(), independent batches sharded over the host cores "
                                                  "(one single-threaded process per core): a baseline the reference does not offer itself"}}
    else:
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
        sample_txt = (f"{args.steps} steps x {sample} cycles x {args.channels} ch x {args.length} samples of the same workload through the "
                      f"oracle PORT of the reference loop (oracle/_ref not built), per-cycle work spread over {workers} processes")
        ms_per_step = 1e3 * elapsed / args.steps
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 mix, f64 spline",
        "data": "synthetic", "config": {"workload": WORKLOAD, "method": args.method, "sample": sample_txt},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": kind, "sample": sample_txt, **extra},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# verification of timed outputs against the CPU oracle
# ----------------------------------------------------------------------------------------------
def oracle_cycles(method, data, frames, mix, lam32, knots, sample):
    """What the reference computes for cycles ``sample`` of a batch: per-pair blend (oracle.mix_pair) and, for
    PCGmix+, the SciPy not-a-knot spline of the cycle's knots times the blend in float64, stored as float32."""
    from oracle import pcgmix_oracle as orc
    mixed = np.stack([orc.mix_pair(data[i], data[mix[i]], frames[i], frames[mix[i]], np.float32(lam32)) for i in sample])
    if knots is None:
        return mixed
    curves = orc.warp_curves(data.shape[-1], knots[sample])
    return (mixed.astype(np.float64) * curves).astype(np.float32)


class Verifier:
    """Accumulates comparisons of sampled output cycles with the oracle."""

    def __init__(self, tolerance=1e-5):
        self.tolerance, self.cycles, self.max_rel, self.equal, self.total = tolerance, 0, 0.0, 0, 0

    def check(self, got, want):
        denom = np.maximum(np.abs(want.astype(np.float64)), np.finfo(np.float32).tiny)
        rel = np.abs(got.astype(np.float64) - want.astype(np.float64)) / denom
        self.max_rel = max(self.max_rel, float(rel.max()))
        self.equal += int((got.view(np.uint32) == want.view(np.uint32)).sum())
        self.total += got.size
        self.cycles += got.shape[0]

    def report(self):
        return {"cycles": self.cycles, "max_rel": self.max_rel, "bit_equal": self.equal / max(self.total, 1),
                "tolerance": self.tolerance, "ok": self.cycles > 0 and self.max_rel <= self.tolerance,
                "against": "oracle/pcgmix_oracle.py (per-pair blend + SciPy not-a-knot spline, float64 product)"}


# ----------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    from pcgmix_b200 import augmentations, draws, native, resident, staging, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
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
    native.set_spline_precision(args.spline)
    native.set_tuning(args.kernel == "pipeline", args.stages, args.max_slice, args.ctas_per_sm, args.pbuf_pct,
                      args.consumer_threads, args.debug_skip)
    verify = not args.no_verify and args.debug_skip == 0

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
        knots = None
        if magwarp:
            lam, knots = draws.lambda_and_knots(plan.alpha, seed, B, plan.knot, C, plan.sigma)
        else:
            lam = draws.draw_lambda(plan.alpha, seed)
        lam32, oml = draws.lambda_pair_fp32(lam)
        arrays = [frames.astype(np.int32), mix.astype(np.int32),
                  np.arange(B, dtype=np.int32) if args.no_order else draws.processing_order(mix)]
        if magwarp:
            arrays.append(knots)
        on_dev = staging.upload(arrays, dev)
        keep = s >= W + K - NOUT                                 # host copies of the last steps' draws, for the verification
        steps_meta.append({"dev": on_dev, "lam": (lam32, oml), "M": synth.mixed_samples(frames, mix),
                           "mix": mix if keep else None, "knots": knots if keep else None})
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

    # consecutive steps are independent (distinct input batches, three rotating output buffers): let
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
    overlapped_in_graph = 0
    if not per_launch_events and not args.no_graph:
        try:
            graph = torch.cuda.CUDAGraph()
            before = native.overlap_launches()
            with torch.cuda.graph(graph):
                capture_handle = torch.cuda.current_stream(dev).cuda_stream
                for k in range(K):
                    prepared[W + k].launch(capture_handle)
            overlapped_in_graph = native.overlap_launches() - before      # launches the library issued with the overlap attribute
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
    overlapped = overlapped_in_graph if graph is not None else native.overlap_launches() - overlap_before
    native.set_launch_overlap(False)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    total_ms = marks[0].elapsed_time(marks[-1])
    per_launch_ms = ([marks[k].elapsed_time(marks[k + 1]) for k in range(K)] if per_launch_events else [total_ms / K])
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None

    # ---- the timed region's own outputs against the oracle (the last NOUT steps still sit in the output buffers) ----
    verifier = Verifier()
    if verify:
        vr = np.random.default_rng(12345 + rank)
        for s in range(max(W, W + K - NOUT), W + K):
            data, frames, labels = batches[s % NB]
            sample = np.sort(vr.choice(B, min(32, B), replace=False))
            got = outs[s % NOUT][torch.from_numpy(sample).to(dev)].cpu().numpy()
            want = oracle_cycles(args.method, data, frames, steps_meta[s]["mix"], steps_meta[s]["lam"][0],
                                 steps_meta[s]["knots"], sample)
            verifier.check(got, want)

    # the same K steps once more with ordinary stream serialisation and an event pair around every
    # launch: per-launch statistics, reported next to the headline for comparison
    serial_ms = per_launch_ms
    if not per_launch_events:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
        for k in range(min(3, K)):                          # untimed: the host gets ahead of the device, so that no interval
            launch(W + k)                                   # below contains the CPU side of a launch on an idle stream
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
    checksum_host = [torch.empty(B, dtype=torch.float32).pin_memory() for _ in range(2)]

    def stage_in(i):
        with torch.cuda.stream(s_in):
            s_in.wait_event(in_free[i % NIN])                  # the augment that read this buffer is done
            dev_in[i % NIN].copy_(host_in[i % len(host_in)], non_blocking=True)
            in_ready[i % NIN].record(s_in)

    host_s = [0.0]                                             # wall time spent inside augment() (draws + launch)
    last = {}

    def run_e2e(n, seed0, result="host"):
        """``result``: "host" = the augmented batch goes back to pinned host memory (164 MB per step);
        "device" = it is consumed on the device (what a training step does) and only a per-cycle checksum
        (16 KB) is read back."""
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
            out, _, mix_i, _ = augmentations.augment(a, dev_in[i % NIN], ohe_t[j], frames_t[j], wav, _Step(seed0 + i), None, dev, None)
            host_s[0] += time.perf_counter() - h0
            in_free[i % NIN].record(stream)
            if result == "device":
                sums = out.sum(dim=(1, 2))                     # the device-side consumer's stand-in
            done = torch.cuda.Event()
            done.record(stream)
            out_done[i % 2].synchronize()                      # the host has the result of step i-2 (and would consume it now)
            with torch.cuda.stream(s_out):
                s_out.wait_event(done)
                if result == "device":
                    checksum_host[i % 2].copy_(sums, non_blocking=True)
                    sums.record_stream(s_out)
                else:
                    host_out[i % 2].copy_(out, non_blocking=True)
                out.record_stream(s_out)
                out_done[i % 2].record(s_out)
            last.update(step=seed0 + i, slot=i % 2, batch=j, mix=mix_i)
        stream.wait_stream(s_out)
        stream.wait_stream(s_in)

    def time_loop(fn, n, seed0, **kw):
        fn(3, seed0 - 1000, **kw)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        wall0 = time.perf_counter()
        e0.record(stream)
        fn(n, seed0, **kw)
        e1.record(stream)
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - wall0) * 1e3
        ms = max(e0.elapsed_time(e1), wall_ms)                 # host work is part of the step: take the longer clock
        tt = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    launches_e2e0 = native.launch_count
    host_s[0] = 0.0
    e2e_ms_total = time_loop(run_e2e, E, 20_000 + rank * E)
    host_inside_ms = 1e3 * host_s[0] / (E + 3)
    e2e_value = world * E * B / (e2e_ms_total * 1e-3)
    e2e_launches = native.launch_count - launches_e2e0
    if verify:                                                 # the last e2e result, as it arrived in pinned host memory
        data, frames, labels = batches[last["batch"]]
        np.random.seed(last["step"])
        lam_e = np.random.beta(plan.alpha, plan.alpha)
        knots_e = np.random.normal(1.0, plan.sigma, (B, plan.knot + 2, C)) if magwarp else None
        sample = np.sort(np.random.default_rng(777 + rank).choice(B, min(32, B), replace=False))
        want = oracle_cycles(args.method, data, frames, last["mix"], draws.lambda_pair_fp32(lam_e)[0], knots_e, sample)
        verifier.check(host_out[last["slot"]].numpy()[sample], want)
    # the host share of a step on its own (no GPU involved): the draws augment() makes for one batch
    t_h = time.perf_counter()
    for rep in range(5):
        m_ = draws.pairing(args.method, batches[0][2], wav, 30_000 + rep)
        if magwarp:
            draws.lambda_and_knots(plan.alpha, 30_000 + rep, B, plan.knot, C, plan.sigma)
        else:
            draws.draw_lambda(plan.alpha, 30_000 + rep)
        if augmentations.use_processing_order:
            draws.processing_order(m_)
    host_draws_ms = (time.perf_counter() - t_h) / 5 * 1e3
    in_bytes = B * C * L * 4
    small_bytes = B * 5 * 4 + B * 4 * 2 + (B * (plan.knot + 2) * C * 8 if magwarp else 0)

    # ---- the same loop with fewer bytes on the link ---------------------------------------------
    variants = {}
    if not args.no_variants:
        ms_dev = time_loop(run_e2e, E, 40_000 + rank * E, result="device")
        variants["padded_batch_up_result_consumed_on_device"] = {
            "value": world * E * B / (ms_dev * 1e-3), "unit": UNIT, "ms_per_step": ms_dev / E,
            "h2d_bytes_per_step": in_bytes + small_bytes, "d2h_bytes_per_step": B * 4 + B * 8,
            "what": "train_model.py:499-507 as it is: data.to(device) of the padded batch, augment(); the result stays on the "
                    "device (its consumer is the model) and a per-cycle checksum is read back"}
        # batches drawn from recordings that stay on the device: per step only table rows + draws go up
        rr = np.random.default_rng(synth.BENCH_SEED + 77 + rank)
        n_rec, t_rec = 256, 40000
        st = torch.from_numpy(synth.dense_states(rr, n_rec, t_rec, 1000)).to(dev)
        sig = torch.from_numpy(rr.standard_normal((n_rec, C, t_rec)).astype(np.float32)).to(dev)
        res = resident.from_dense_states(sig, st, L)
        ids_host = [rr.integers(0, res.n_cycles, B) for _ in range(4)]
        tgt_host = [torch.from_numpy(rr.integers(0, 2, B)) for _ in range(4)]
        ohe_r = [augmentations.with_host_labels(torch.nn.functional.one_hot(t_, 2).to(dev), t_) for t_ in tgt_host]

        def run_resident(n, seed0, result="host"):
            for ev in out_done:
                ev.record(stream)
            for i in range(n):
                out, _, _, _ = resident.augment(a, res, ids_host[i % 4], ohe_r[i % 4], None, _Step(seed0 + i), None, dev, None)
                if result == "device":
                    sums = out.sum(dim=(1, 2))
                done = torch.cuda.Event()
                done.record(stream)
                out_done[i % 2].synchronize()                  # the host has the result of step i-2
                with torch.cuda.stream(s_out):
                    s_out.wait_event(done)
                    if result == "device":
                        checksum_host[i % 2].copy_(sums, non_blocking=True)
                        sums.record_stream(s_out)
                    else:
                        host_out[i % 2].copy_(out, non_blocking=True)
                    out.record_stream(s_out)
                    out_done[i % 2].record(s_out)
            stream.wait_stream(s_out)

        ms_res = time_loop(run_resident, E, 50_000 + rank * E)
        ms_res_dev = time_loop(run_resident, E, 60_000 + rank * E, result="device")
        res.check()
        up_res = B * 4 + B * 4 + (B * (plan.knot + 2) * C * 8 if magwarp else 0)
        variants["resident_recordings_result_to_host"] = {
            "value": world * E * B / (ms_res * 1e-3), "unit": UNIT, "ms_per_step": ms_res / E,
            "h2d_bytes_per_step": up_res, "d2h_bytes_per_step": in_bytes,
            "what": "pcgmix_b200.resident.augment: recordings + cycle table stay on the device (uploaded once), per step the "
                    "batch's table rows, pairing and knots go up and the augmented batch comes back to pinned host memory"}
        variants["resident_recordings_result_consumed_on_device"] = {
            "value": world * E * B / (ms_res_dev * 1e-3), "unit": UNIT, "ms_per_step": ms_res_dev / E,
            "h2d_bytes_per_step": up_res, "d2h_bytes_per_step": B * 4,
            "what": "the same with the result consumed on the device (per-cycle checksum read back): bounded by the host's "
                    "draws, not by the link"}
        del st, sig, res

    # ---- device-timed legs of the other BASELINE configurations (rank 0's GPU; every rank runs them so that ranks stay in step)
    configs = None
    if not args.no_configs:
        sys.path.insert(0, os.path.join(ROOT, "benchmarks"))
        import run_configs
        run_configs.PEAK = peak
        lines = run_configs.collect(dev, reps=max(20, min(K, 60)))
        configs = [{k: v for k, v in ln.items() if k in ("config", "cycles_per_call", "ms_mean", "ms_min", "cycles_per_s",
                                                          "achieved_GBps", "frac_of_measured_peak", "host_draws_ms_total",
                                                          "wall_ms_including_host_draws", "cycles_per_s_including_host_draws")}
                   for ln in lines]
    native.set_spline_precision(args.spline)

    # ---- BASELINE config 5: the DDP training step fed by the on-device augmentation ----------------
    cfg5 = None
    if not args.no_cfg5:
        sys.path.insert(0, os.path.join(ROOT, "examples"))
        import train_ddp_pcgmix
        cfg5 = {"padded_batches": train_ddp_pcgmix.run_cfg5(args.cfg5_steps, 64, False, False, init_process_group=False),
                "resident_recordings": train_ddp_pcgmix.run_cfg5(args.cfg5_steps, 64, True, False, init_process_group=False)}

    # ---- CPU baseline on rank 0 at N=1 ----------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        workers = os.cpu_count() or 1
        if reference_kind() == "reference":
            sample_b = max(8, min(B, args.reference_sample))
            t0 = time.perf_counter()
            v, _ = time_reference(args.method, sample_b, C, L, 6, 1, 1, workers)
            v_all, _ = time_reference(args.method, sample_b, C, L, 2, 1, workers, 1)
            cpu = {"value": v, "unit": UNIT, "cores": workers, "kind": "reference",
                   "sample": f"6 steps x {sample_b} cycles x {C} ch x {L}: the unmodified reference augment() (oracle/_ref) as it is, one "
                             f"Python process with torch.set_num_threads({workers}); {time.perf_counter() - t0:.1f} s incl. the leg below",
                   "one_process_per_core": {"value": v_all, "unit": UNIT, "processes": workers,
                                            "what": "independent batches sharded over the host cores, one single-threaded reference "
                                                    "process per core (2 steps each)"}}
        else:
            sample_b = 1024
            v, done, elapsed, _ = time_cpu_baseline(args.method, sample_b, C, L, args.cpu_seconds, workers)
            cpu = {"value": v, "unit": UNIT, "cores": workers, "kind": "port",
                   "sample": f"{done} cycles ({done // sample_b} steps of {sample_b} x {C} ch x {L}) in {elapsed:.1f} s; "
                             f"oracle port of the reference loop, per-cycle work spread over {workers} processes"}

    verified = verifier.report() if verify else {"ok": None, "skipped": "--no-verify or --debug-skip"}
    if rank == 0:
        serial_mean = statistics.fmean(serial_ms)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": max_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 mix, " + ("f32" if args.spline == "float32" else "f64") + " spline", "data": "synthetic",
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
                       "kernel": args.kernel, "spline_evaluation": args.spline, "stages": args.stages, "max_slice": args.max_slice,
                       "ctas_per_sm": args.ctas_per_sm, "pbuf_pct": args.pbuf_pct, "consumer_threads": args.consumer_threads,
                       "debug_skip": args.debug_skip, "library": os.path.basename(native.library_path()),
                       "timed_region": "K = %d launches, %.2f ms; robust at the driver's K = 20 (SCALE and BENCH of round 1 "
                                       "agreed to 0.03 %%)" % (K, max_ms),
                       "sharding": "batches per rank, pairing inside each batch, no collective",
                       "host_affinity": f"rank bound to the {local_cpus} CPU cores local to its GPU" if local_cpus else "default"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": statistics.fmean(bytes_per_launch),
                         "kernel_ms_mean": statistics.fmean(per_launch_ms),
                         "note": "achieved = ALGORITHMIC bytes / time (effective bandwidth: partner reads that hit L2 count); "
                                 "`dram` relates the DRAM bytes ncu measured for one launch to the same times",
                         "dram": None if traffic is None else {
                             "overlapped_GBps": traffic / mean_launch_s / 1e9, "overlapped_frac": traffic / mean_launch_s / 1e9 / peak,
                             "serialized_GBps": traffic / (serial_mean * 1e-3) / 1e9,
                             "serialized_frac": traffic / (serial_mean * 1e-3) / 1e9 / peak},
                         "serialized_launches": {
                             "kernel_ms_mean": serial_mean, "kernel_ms_median": statistics.median(serial_ms),
                             "kernel_ms_min": min(serial_ms),
                             "achieved": statistics.fmean(bytes_per_launch) / (serial_mean * 1e-3) / 1e9,
                             "frac": statistics.fmean(bytes_per_launch) / (serial_mean * 1e-3) / 1e9 / peak,
                             "cycles_per_s_per_gpu": B / (serial_mean * 1e-3)}},
            "verified": verified,
            "e2e": {"value": e2e_value, "unit": UNIT, "steps": E, "ms_per_step": e2e_ms_total / E,
                    "h2d_bytes_per_step": in_bytes + small_bytes, "d2h_bytes_per_step": in_bytes + B * 8,
                    "host_ms_per_step_inside_augment": host_inside_ms,          # includes waiting for the batch's H2D copy
                    "host_draws_ms_per_step": host_draws_ms,                     # pairing + lambda + knots alone, rank 0
                    "link_ceiling": "profiles/r2_pcie_ranks.json: with N ranks moving 164 MB each way at once this box class delivers "
                                    "99 / 147 / 105 / 144 GB/s in total at N = 1 / 2 / 4 / 8 (8.1 GB/s of D2H for the slowest rank at "
                                    "N = 8), and the copies alone take 3.4 / 4.7 / - / 19.9 ms per step",
                    "api": "pcgmix_b200.augmentations.augment (host draws + 1 kernel) inside a prefetching loop: pinned host in/out, "
                           "H2D of steps k+1, k+2 and D2H of step k-1 on side streams"},
            "e2e_variants": variants,
            "gpu_launches": gpu_launches, "gpu_launches_overlapped": overlapped, "gpu_launches_e2e": e2e_launches,
            "clocks": clocks,
        }
        if configs is not None:
            line["configs"] = configs
        if cfg5 is not None:
            line["cfg5"] = cfg5
        if cpu is not None:
            line["cpu_baseline"] = cpu
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()
    if verify and not verified["ok"]:
        print(f"bench: VERIFICATION FAILED: {verified}", file=sys.stderr)
        sys.exit(3)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
