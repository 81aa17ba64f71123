import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from softray_b200 import lib, abi, synth
from tests.util import scenario
fx = np.load("tests/golden/reference_fixtures.npz")
mesh = oracle.load_3ds(fx["model/obj.3ds"].tobytes())
ctx = lib.Context(0)
sc = lib.Scene(ctx, [mesh])
for mode in (0, 1, 2):
    p = scenario(resolution=96, shadows=True); p.filter_mode = mode
    st = sc.render(p)["stats"]
    print(mode, st.as_dict())
meshes, spheres, p = synth.config2(width=160, height=90)
sc = lib.Scene(ctx, meshes, spheres)
for mode in (0, 1, 2):
    p.filter_mode = mode
    st = sc.render(p)["stats"]
    print(mode, st.as_dict())
