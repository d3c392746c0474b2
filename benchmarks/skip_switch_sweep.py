#!/usr/bin/env python
"""What each part of the pipelined kernel costs: bench.py's device-timed leg with single parts switched off
through the PROFILING build of the library (outputs are wrong by construction; timing only).

    python benchmarks/skip_switch_sweep.py            # builds csrc/libpcgmix_b200_prof.so if needed
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [("float32", 0, "nothing skipped"), ("float32", 128, "one elected consumer warp polls the full barrier, the others park on bar.sync"), ("float32", 64, "no row safety check / knot loads in the producer"),
         ("float32", 2, "no consumer arithmetic"), ("float32", 1, "no stores"), ("float32", 4, "no partner staging"),
         ("float32", 2 + 64, "no arithmetic, no safety check"), ("float64", 0, "float64: nothing skipped"),
         ("float64", 2, "float64: no consumer arithmetic")]


def main():
    from pcgmix_b200 import build_native
    if not os.path.exists(build_native.PROFILING_LIB_PATH) or build_native.needs_build():
        build_native.build_profiling()
    env = dict(os.environ, PCGMIX_PROFILING_LIB="1")
    for method in ("durmixmagwarp(0.2,4)", "durratiomixup"):
        for prec, skip, what in CASES:
            if method == "durratiomixup" and (prec == "float64" or skip & 64 or skip == 4):
                continue
            e = dict(env, PCGMIX_SPLINE=prec)
            out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "40", "--warmup", "5", "--no-cpu-baseline",
                                  "--e2e-steps", "3", "--method", method, "--debug-skip", str(skip), "--no-verify"],
                                 env=e, capture_output=True, text=True)
            try:
                d = json.loads(out.stdout.strip().splitlines()[-1])
                r = d["roofline"]
                print(json.dumps({"method": method, "spline": prec, "skip": skip, "what": what,
                                  "overlapped_ms": round(r["kernel_ms_mean"], 4),
                                  "serialized_ms": round(r["serialized_launches"]["kernel_ms_mean"], 4)}), flush=True)
            except Exception:
                print(json.dumps({"method": method, "skip": skip, "error": out.stderr[-400:]}), flush=True)


if __name__ == "__main__":
    main()
