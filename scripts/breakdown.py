"""Where a frame's time goes: the same workload with stages switched off (kernel ms from softray_stats)."""
import sys

import numpy as np

from softray_b200 import abi, lib, synth

name = sys.argv[1] if len(sys.argv) > 1 else "config2"
ctx = lib.Context(0)
meshes, spheres, p = getattr(synth, name)()
sc = lib.Scene(ctx, meshes, spheres)
px = np.zeros((p.height, p.width), dtype=np.uint32)


def run(label, **kw):
    old = {k: getattr(p, k) for k in kw}
    for k, v in kw.items():
        setattr(p, k, v)
    best = 1e9
    for _ in range(4):
        st = sc.render(p, pixels=px)["stats"]
        best = min(best, st.ms_kernel)
    print(f"{label:42s} {best:8.3f} ms  rays {st.rays:>11d} nodes {st.node_visits:>12d} prim {st.prim_tests:>9d} "
          f"filt {st.filter_tests:>11d} unsure {st.filter_unsure:>8d} bundled {st.rays_bundled}")
    for k, v in old.items():
        setattr(p, k, v)


run("full")
run("no shadows", shadows=False)
run("no shadows, Lambert", shadows=False, specular_lighting=False)
run("no shadows, no shading", shadows=False, shading=False)
run("shadows, no shading", shading=False)
run("shadows 16 samples", shadow_samples=16)
run("shadows 1 sample", shadow_samples=1)
if p.reflection_depth:
    run("no reflection", reflection_depth=0)
if p.texture3d_id:
    run("no texture", texture3d_id=0)
run("exact only (filter off)", filter_mode=abi.FILTER_OFF)
