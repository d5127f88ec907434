"""Builds A/B variants of the library under build/variants/ (each a set of -D macros); tools/ab_run.sh times them.
usage: python tools/build_variants.py name=MACRO1,MACRO2 ..."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iseg_b200 import build as b  # noqa: E402

out = os.path.join(os.path.dirname(b.PKG), "build", "variants")
os.makedirs(out, exist_ok=True)
for f in os.listdir(out):
    os.remove(os.path.join(out, f))
procs = []
for spec in sys.argv[1:]:
    name, _, macros = spec.partition("=")
    defs = [f"-D{m}" for m in macros.split(",") if m]
    cmd = ["nvcc", *b.NVCC_FLAGS, *defs, "-o", os.path.join(out, f"{name}.so"), *b.sources()]
    procs.append((name, subprocess.Popen(cmd)))
for name, p in procs:
    assert p.wait() == 0, name
    print("built", name)
