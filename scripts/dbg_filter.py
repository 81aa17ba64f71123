import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from softray_b200 import lib, abi, synth
from tests.util import scenario
fx = np.load("tests/golden/reference_fixtures.npz")
mesh = oracle.load_3ds(fx["model/obj2.3ds"].tobytes())
ctx = lib.Context(0)
sc = lib.Scene(ctx, [mesh])
for shadows in (False, True):
    p = scenario(resolution=96, shadows=shadows); p.filter_mode = 2
    st = sc.render(p)["stats"]
    print(shadows, st.as_dict())
