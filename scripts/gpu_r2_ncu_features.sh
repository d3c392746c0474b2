#!/bin/bash
# ncu --set full of the two feature kernels (4096 cycles x 2500, channel 4 of 5)
mkdir -p gpurun_out
cat > /tmp/feat_once.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
from pcgmix_b200 import features, staging, synth
rng = np.random.default_rng(synth.BENCH_SEED + 5)
B, C, L = 4096, 5, 2500
frames = synth.cycle_frames(rng, B, limit=L)
dev = torch.device("cuda:0")
data = torch.from_numpy(synth.cycle_signals(rng, frames, (C,), L)).to(dev)
fd = staging.upload([frames.astype(np.int32)], dev)[0]
for _ in range(3):
    features.cycle_features(data, fd, 4)
    features.cycle_psd_features(data, fd, 4)
torch.cuda.synchronize()
PY
python /tmp/feat_once.py || exit 1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:cycle_ --launch-skip 4 -c 2 -o gpurun_out/r2_prof_features -f python /tmp/feat_once.py > gpurun_out/ncu_feat.log 2>&1
tail -2 gpurun_out/ncu_feat.log
