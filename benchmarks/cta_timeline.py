#!/usr/bin/env python
"""When do the CTAs of ONE pipelined launch start, consume their first item and finish?  (profiling build)

    PCGMIX_PROFILING_LIB=1 python benchmarks/cta_timeline.py [--method durratiomixup]

Prints the spread of the three time stamps over the grid: how long the pipeline takes to fill and how far
apart the CTAs finish (the tail a dynamic work distribution could recover)."""
import argparse
import ctypes
import json
import os
import sys

import numpy as np
import torch

os.environ.setdefault("PCGMIX_PROFILING_LIB", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pcgmix_b200 import augmentations, build_native, draws, native, staging, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--method", default="durmixmagwarp(0.2,4)")
    args = ap.parse_args()
    if not os.path.exists(build_native.PROFILING_LIB_PATH):
        build_native.build_profiling()
    lib = native.load()
    lib.pcgmix_debug_timeline.argtypes = [ctypes.c_void_p, ctypes.c_int32]
    dev = torch.device("cuda:0")
    B, C, L = 4096, 4, 2500
    plan = draws.parse_method_1d(args.method)
    magwarp = plan.branch == "durmixmagwarp"
    batches = [bench.make_batch(synth.BENCH_SEED + i, B, C, L) for i in range(3)]
    data = [torch.from_numpy(b[0]).to(dev) for b in batches]
    out = torch.empty_like(data[0])
    ups = []
    for s, (_, frames, labels) in enumerate(batches):
        mix = draws.same_label_pairing(labels, s)
        lam, knots = draws.lambda_and_knots(1.0, s, B, 4, C, 0.2)
        arrays = [frames.astype(np.int32), mix.astype(np.int32), draws.processing_order(mix)] + ([knots] if magwarp else [])
        ups.append((staging.upload(arrays, dev), draws.lambda_pair_fp32(lam)))
    grid = 148 * 3
    rows = []
    per_sm, by_arrival, by_wave = {}, {}, {}
    for rep in range(6):
        up, lam = ups[rep % 3]
        augmentations.pcgmix_on_device(data[rep % 3], up[0], up[1], lam[0], lam[1], up[3] if magwarp else None, 4, order_dev=up[2], out=out)
        torch.cuda.synchronize()
        t = np.zeros(4 * grid, dtype=np.uint64)
        assert lib.pcgmix_debug_timeline(t.ctypes.data, grid) == 0
        t = t.reshape(grid, 4).astype(np.int64)
        sm = t[:, 3]
        t0 = t[:, 0].min()
        rel = (t[:, :3] - t0) / 1e3
        per_sm[rep] = {int(k): float(rel[sm == k, 2].mean()) for k in np.unique(sm)}
        within = float(np.mean([np.ptp(rel[sm == k, 2]) for k in np.unique(sm)]))
        # the CTAs of an SM in the order of their block index (= the order in which the SM received them)
        ranked = np.array([rel[np.flatnonzero(sm == k), 2] for k in np.unique(sm) if (sm == k).sum() == 3])
        by_arrival[rep] = [round(float(v), 2) for v in ranked.mean(axis=0)] if len(ranked) else []
        by_wave[rep] = [round(float(rel[w * 148:(w + 1) * 148, 2].mean()), 2) for w in range(grid // 148)]
        rows.append({"start_spread_us": float(rel[:, 0].max()), "first_item_us_median": float(np.median(rel[:, 1])),
                     "first_item_us_max": float(rel[:, 1].max()), "end_us_min": float(rel[:, 2].min()),
                     "end_us_p10": float(np.percentile(rel[:, 2], 10)), "end_us_median": float(np.median(rel[:, 2])),
                     "end_us_p90": float(np.percentile(rel[:, 2], 90)), "end_us_max": float(rel[:, 2].max()),
                     "end_spread_within_an_sm_us_mean": within})
    for r in rows[2:]:
        print(json.dumps({"method": args.method, **{k: round(v, 2) for k, v in r.items()}}))
    # is a CTA's finish time a property of the SM it ran on?  correlation of the per-SM mean finish time between launches
    # (different batches), and the SMs ranked by it
    sms = sorted(set(per_sm[3]) & set(per_sm[4]) & set(per_sm[5]))
    a, b, c = (np.array([per_sm[r][k] for k in sms]) for r in (3, 4, 5))
    order = np.argsort(a + b + c)
    print(json.dumps({"method": args.method, "sms": len(sms),
                      "per_sm_mean_end_us_min_max": [round(float((a + b + c).min() / 3), 2), round(float((a + b + c).max() / 3), 2)],
                      "corr_between_launches": [round(float(np.corrcoef(a, b)[0, 1]), 3), round(float(np.corrcoef(a, c)[0, 1]), 3)],
                      "mean_end_us_of_an_sms_ctas_in_block_index_order": [by_arrival[r] for r in (3, 4, 5)],
                      "mean_end_us_by_block_index_third": [by_wave[r] for r in (3, 4, 5)],
                      "fastest_sms": [int(sms[i]) for i in order[:12]], "slowest_sms": [int(sms[i]) for i in order[-12:]]}))


if __name__ == "__main__":
    main()
