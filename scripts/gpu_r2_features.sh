#!/bin/bash
# feature kernels: tests, then device-timed legs (with the CPU port beside them)
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_features.py -m gpu -q -x 2>&1 | tail -5
timeout -s KILL 300 python - <<'PY'
import sys, json, argparse, torch
sys.path.insert(0, "."); sys.path.insert(0, "benchmarks")
import run_configs
from pcgmix_b200 import native
native.load()
run_configs.features_section(argparse.Namespace(), torch.device("cuda:0"), True)
PY
