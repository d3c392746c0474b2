"""Where the host time of the drop-in ``augment`` call goes INSIDE the config-5 training loop (cProfile enabled only
around the call; 100 calls after 10 warm-up steps)."""
import cProfile, pstats, io, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'examples'))
import train_ddp_pcgmix as ex
from pcgmix_b200 import augmentations
prof = cProfile.Profile()
real = augmentations.augment
calls = [0]
def wrapped(*a, **k):
    calls[0] += 1
    if calls[0] <= 10:
        return real(*a, **k)
    prof.enable()
    try:
        return real(*a, **k)
    finally:
        prof.disable()
augmentations.augment = wrapped
ex.augmentations.augment = wrapped
print(ex.run_cfg5(steps=110, batch=64))
out = io.StringIO()
pstats.Stats(prof, stream=out).sort_stats("cumulative").print_stats(30)
print(out.getvalue()[:5000])
