import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from softray_b200 import lib, abi, synth
ctx = lib.Context(0)
meshes, spheres, p = synth.config2()
sc = lib.Scene(ctx, meshes, spheres)
for name, lp in (("far light (default)", None), ("light inside the room", (0.1, 0.2, 1.3))):
    p.light_pos_view = lp
    for mode in (0, 1):
        p.filter_mode = mode
        sc.render(p)
        st = sc.render(p)["stats"]
        print(name, "mode", mode, "ms_kernel %.3f" % st.ms_kernel, "sphere_tests", st.sphere_tests, "filter_tests", st.filter_tests, "unsure", st.filter_unsure, "bundled", st.rays_bundled)
