#!/usr/bin/env python3
"""softray_create_multi: one process, k GPUs behind one softray_render into a page-locked host surface.
usage: bench_group.py [workload] [steps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from softray_b200 import lib  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "config5"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
meshes, spheres, frame, _ = bench.workload(name)
ref = None
for k in range(1, torch.cuda.device_count() + 1):
    if k not in (1, 2, 4, 8):
        continue
    ctx = lib.Context(n_devices=k)
    t = time.perf_counter()
    scene = lib.Scene(ctx, meshes, spheres)
    t_scene = time.perf_counter() - t
    px = torch.empty((frame.height, frame.width), dtype=torch.int32).pin_memory()
    hp = px.numpy().view(np.uint32)
    for _ in range(3):
        st = scene.render(frame, pixels=hp)["stats"]
    t = time.perf_counter()
    for _ in range(steps):
        scene.render(frame, pixels=hp, want_stats=False)
    ms = (time.perf_counter() - t) * 1e3 / steps
    if ref is None:
        ref = hp.copy()
    print(f"{name}: group of {k} device(s): {ms:.3f} ms per frame end to end ({st.rays / ms / 1e3:.0f} Mrays/s), scene created + replicated in "
          f"{t_scene:.2f} s, frame identical to 1 device: {bool(np.array_equal(ref, hp))}", flush=True)
    scene.close()
    ctx.close()
