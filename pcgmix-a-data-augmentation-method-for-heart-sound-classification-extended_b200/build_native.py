"""Compile ``csrc/*.cu`` into ``csrc/libpcgmix_b200.so`` for sm_100a with nvcc (in-tree, so the
built library travels with the repository snapshot)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB_PATH = os.path.join(CSRC, "libpcgmix_b200.so")
OBJ_DIR = os.path.join(CSRC, "build")                      # intermediate objects (git- and gpurun-ignored)
SOURCES = ("mix_kernels.cu", "mix_pipeline.cu", "mix_resident.cu", "segment_kernels.cu", "feature_kernels.cu", "psd_kernels.cu", "first_block_kernels.cu", "capi.cu", "host_draws.cpp")
NVCC_FLAGS = ("-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--shared", "-Xcompiler", "-fPIC",
              # host code replays NumPy's / CPython's generators bit for bit: no FMA contraction there either
              "-Xcompiler", "-ffp-contract=off")


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libpcgmix_b200.so")


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.join(CSRC, "common.cuh"),
                                                       os.path.join(INCLUDE, "pcgmix_b200.h")]
    return any(os.path.getmtime(d) > built for d in deps)


PROFILING_LIB_PATH = os.path.join(CSRC, "libpcgmix_b200_prof.so")


def build_profiling(verbose: bool = False) -> str:
    """A second library with the pipelined kernel's skip switches compiled IN (-DPCGMIX_PROFILING): parts of the
    kernel can be switched off through ``pcgmix_set_tuning(debug=...)`` to see what they cost — the output is then
    WRONG.  Never loaded unless ``PCGMIX_PROFILING_LIB=1`` is set (benchmarks/skip_switch_sweep.py does)."""
    from concurrent.futures import ThreadPoolExecutor
    odir = os.path.join(OBJ_DIR, "prof")
    os.makedirs(odir, exist_ok=True)
    jobs = [(os.path.join(CSRC, n), os.path.join(odir, os.path.splitext(n)[0] + ".o")) for n in SOURCES]
    with ThreadPoolExecutor(max_workers=len(jobs)) as pool:
        list(pool.map(lambda j: _compile_one(j[0], j[1], verbose, ("-DPCGMIX_PROFILING",)), jobs))
    proc = subprocess.run([_nvcc(), "--shared", "-o", PROFILING_LIB_PATH, *[j[1] for j in jobs]], capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + proc.stdout + proc.stderr)
    return PROFILING_LIB_PATH


def _compile_one(src: str, obj: str, verbose: bool, extra=()):
    cmd = [_nvcc(), *[f for f in NVCC_FLAGS if f != "--shared"], *extra, "-I", INCLUDE, "-c", "-o", obj, src]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed on {os.path.basename(src)}:\n" + proc.stdout + proc.stderr)
    return proc.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    # several ranks may arrive here together (torchrun): one compiles, the others wait and reuse
    import fcntl
    from concurrent.futures import ThreadPoolExecutor
    with open(os.path.join(CSRC, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():
                return LIB_PATH
            # one nvcc per translation unit, side by side; objects are reused while their source is older
            headers = [os.path.join(CSRC, "common.cuh"), os.path.join(INCLUDE, "pcgmix_b200.h")]
            newest_header = max(os.path.getmtime(h) for h in headers)
            jobs, objects = [], []
            os.makedirs(OBJ_DIR, exist_ok=True)
            for name in SOURCES:
                src = os.path.join(CSRC, name)
                obj = os.path.join(OBJ_DIR, os.path.splitext(name)[0] + ".o")
                objects.append(obj)
                stale = (force or not os.path.exists(obj) or os.path.getmtime(obj) < os.path.getmtime(src)
                         or os.path.getmtime(obj) < newest_header)
                if stale:
                    jobs.append((src, obj))
            with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as pool:
                logs = list(pool.map(lambda j: _compile_one(j[0], j[1], verbose), jobs))
            tmp = LIB_PATH + f".tmp{os.getpid()}"
            proc = subprocess.run([_nvcc(), "--shared", "-o", tmp, *objects], capture_output=True, text=True)
            if proc.returncode != 0:
                raise RuntimeError("nvcc link failed:\n" + proc.stdout + proc.stderr)
            os.replace(tmp, LIB_PATH)                      # atomic: readers never see a half-written library
            if verbose:
                print("".join(logs))
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
