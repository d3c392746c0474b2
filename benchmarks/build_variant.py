#!/usr/bin/env python
"""Build a variant of the native library for an A/B measurement: the pipelined kernel's source with textual
substitutions applied, linked with the other (unchanged) objects into ``csrc/variants/lib<name>.so``.  Load it with
``PCGMIX_LIB=<that path>``.  Developer tool; nothing in the product uses it.

    python benchmarks/build_variant.py NAME 'old text' 'new text' ['old2' 'new2' ...]
"""
import importlib
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
bn = importlib.import_module("pcgmix-a-data-augmentation-method-for-heart-sound-classification-extended_b200.build_native")


def main():
    name, pairs = sys.argv[1], sys.argv[2:]
    bn.build()
    src = open(os.path.join(bn.CSRC, "mix_pipeline.cu")).read()
    for old, new in zip(pairs[0::2], pairs[1::2]):
        if src.count(old) != 1:
            raise SystemExit(f"{src.count(old)} occurrences of {old[:60]!r}")
        src = src.replace(old, new)
    vdir = os.path.join(bn.CSRC, "variants")
    os.makedirs(vdir, exist_ok=True)
    vsrc = os.path.join(bn.CSRC, f"_variant_{name}.cu")
    obj = os.path.join(bn.OBJ_DIR, f"_variant_{name}.o")
    open(vsrc, "w").write(src)
    try:
        bn._compile_one(vsrc, obj, False)
    finally:
        os.remove(vsrc)
    objects = [os.path.join(bn.OBJ_DIR, os.path.splitext(n)[0] + ".o") for n in bn.SOURCES if n != "mix_pipeline.cu"] + [obj]
    out = os.path.join(vdir, f"lib{name}.so")
    subprocess.run([bn._nvcc(), "--shared", "-o", out, *objects], check=True)
    print(out)


if __name__ == "__main__":
    main()
