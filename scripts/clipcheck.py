"""SR_CLIP_CHECK build: render every config (reduced) in AUTO mode; filter_mismatch counts fast-clip vs reference_clip differences."""
import numpy as np
from softray_b200 import lib, synth, abi
import tests.util as U
import oracle
ctx = lib.Context(0)
fx = np.load("tests/golden/reference_fixtures.npz")
jobs = []
for name in ("model/obj.3ds", "model/obj2.3ds"):
    mesh = oracle.load_3ds(fx[name].tobytes())
    for kw in (dict(), dict(shadows=True, shadow_samples=8), dict(sub_pixel_res=2), dict(yaw_deg=20.0, pitch_deg=35.0, object_depth=0.8),
               dict(yaw_deg=0.0, pitch_deg=0.0), dict(yaw_deg=90.0, pitch_deg=0.0), dict(yaw_deg=45.0, pitch_deg=-89.0)):
        jobs.append((name + str(kw), [mesh], None, U.scenario(resolution=160, **kw)))
m, s, p = synth.config2(width=480, height=270, shadow_samples=16); jobs.append(("config2", m, s, p))
m, s, p = synth.config3(width=480, height=270, nx=301, nz=151, shadow_samples=4); jobs.append(("config3", m, s, p))
m, s, p = synth.config4(width=320, height=180, n_lon=60, n_lat=40, n_side=5, sub_pixel_res=2); jobs.append(("config4", m, s, p))
m, s, p = synth.config5(width=480, height=270, n_lon=60, n_lat=40, n_side=5, shadow_samples=2); jobs.append(("config5", m, s, p))
bad = 0
for name, meshes, sph, p in jobs:
    sc = lib.Scene(ctx, meshes, sph)
    st = sc.render(p)["stats"]
    print(f"{name:60s} rays {st.rays:>10d} exact tests {st.prim_tests:>9d} clip mismatches {st.filter_mismatch}")
    bad += st.filter_mismatch
    sc.close()
ctx.close()
print("TOTAL MISMATCH", bad)
