#!/usr/bin/env python3
"""Render exactly N frames of a bench workload on cuda:0 and nothing else (no stats frames, no peaks, no CPU arm):
the subject of `ncu` captures.  usage: one_frame.py <workload> [n_frames]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from softray_b200 import lib  # noqa: E402

name, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1
meshes, spheres, frame, _ = bench.workload(name)
ctx = lib.Context(0)
scene = lib.Scene(ctx, meshes, spheres)
fb = torch.zeros((frame.height, frame.width), dtype=torch.int32, device="cuda")
for _ in range(n):
    scene.render_device(frame, fb.data_ptr())
torch.cuda.synchronize()
print(name, n, "frames")
