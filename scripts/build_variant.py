#!/usr/bin/env python3
"""Build an experimental variant of libsoftray_cuda.so into variants/<name>.so (git-ignored; travels to the GPU box).
usage: build_variant.py <name> [-DMACRO=value ...]      then:  SOFTRAY_SO=variants/<name>.so python bench.py ..."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from softray_b200 import lib  # noqa: E402

name, flags = sys.argv[1], sys.argv[2:]
os.makedirs(os.path.join(ROOT, "variants"), exist_ok=True)
out = os.path.join(ROOT, "variants", name + ".so")
cmd = ["nvcc"] + lib.NVCC_FLAGS + flags + ["-o", out] + [os.path.join(lib.CSRC, s) for s in lib.SOURCES]
res = subprocess.run(cmd, capture_output=True, text=True)
if res.returncode != 0:
    raise SystemExit(res.stdout + res.stderr)
print(out)
