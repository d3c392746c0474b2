"""The reference arm of bench.py runs on the CPU, so its JSON contract can be checked here."""
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                           "--warmup", "0", "--batch", "64"], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [l for l in proc.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "augmented_cardiac_cycles_per_sec" and d["unit"] == "cycles/s"
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    # "reference" when oracle/_ref holds the byte-compiled reference (oracle/build_ref.py), otherwise the oracle port
    built = all(os.path.isfile(os.path.join(ROOT, "oracle", "_ref", m + ".bytecode")) for m in ("augmentations", "augmentations2d"))
    assert d["cpu_baseline"]["kind"] == ("reference" if built else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"] and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "cycles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None and d["gpu_launches"] == 0 and "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                           "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert proc.returncode == 0 and proc.stdout.strip() == ""


def test_reference_arm_falls_back_to_the_port_without_oracle_ref(tmp_path):
    """The arm must work on a tree where oracle/_ref was never built (kind "port")."""
    import shutil
    work = tmp_path / "repo"
    shutil.copytree(ROOT, work, ignore=shutil.ignore_patterns(".git", "gpurun_out", "_ref", "*.so", "build", "profiles", "__pycache__"))
    proc = subprocess.run([sys.executable, str(work / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--batch", "32"],
                          capture_output=True, text=True, timeout=300, env=dict(os.environ, OMP_NUM_THREADS="1"), cwd=str(work))
    assert proc.returncode == 0, proc.stderr[-2000:]
    d = json.loads([l for l in proc.stdout.splitlines() if l.startswith("{")][0])
    assert d["cpu_baseline"]["kind"] == "port" and d["value"] > 0
