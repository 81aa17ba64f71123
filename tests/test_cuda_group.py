"""softray_create_multi: one process, every GPU of the box behind one softray_render (the reference fans the rows of
one Render() out to tasks that share surface.Pixels, Renderer.cs:1655-1680).  A group of ONE device runs the whole
group code path on a single-GPU box; the two-device cases need two visible devices."""
import ctypes as C

import numpy as np
import pytest

from softray_b200 import abi, synth
from tests.util import scenario

pytestmark = pytest.mark.gpu

EXACT_COUNTERS = ("rays_primary", "rays_shadow", "rays_secondary", "hits_primary", "shaded_hits")


@pytest.fixture(scope="module")
def lib():
    from softray_b200 import lib as L

    return L


def n_visible():
    import torch

    return torch.cuda.device_count()


def compare_group_with_single(lib, n_devices, obj_mesh):
    single = lib.Context(0)
    group = lib.Context(n_devices=n_devices)
    try:
        assert group.n_devices == n_devices and single.n_devices == 1
        cases = [([obj_mesh], None, scenario(width=203, height=167, shadows=True, shadow_samples=6, sub_pixel_res=2))]
        m3, _, p3 = synth.config3(width=240, height=135, nx=301, nz=151, shadow_samples=20)       # stage kernels
        cases.append((m3, None, p3))
        m2, s2, p2 = synth.config2(width=160, height=90, shadow_samples=8, n_spheres=200)          # spheres: fused kernel
        cases.append((m2, s2, p2))
        m4, _, p4 = synth.config4(width=128, height=72, n_lon=30, n_lat=20, n_side=3, sub_pixel_res=2)
        cases.append((m4, None, p4))
        for meshes, spheres, p in cases:
            a = lib.Scene(single, meshes, spheres).render(p, want_ids=True)
            sg = lib.Scene(group, meshes, spheres)
            px = np.full((p.height, p.width), 0xDEADBEEF, dtype=np.uint32)
            b = sg.render(p, want_ids=True, pixels=px)
            assert (a["pixels"] == b["pixels"]).all() and (a["ids"] == b["ids"]).all()
            for k in EXACT_COUNTERS:
                assert getattr(a["stats"], k) == getattr(b["stats"], k), k
            assert sg.fingerprint() == lib.Scene(single, meshes, spheres).fingerprint()
            # a row range: only those rows are written, whatever device rendered them
            p.start_row, p.end_row = 10, p.height - 20
            px[:] = 0xDEADBEEF
            sg.render(p, pixels=px)
            assert (px[:10] == 0xDEADBEEF).all() and (px[p.height - 19:] == 0xDEADBEEF).all()
            assert (px[10:p.height - 19] == a["pixels"][10:p.height - 19]).all()
            p.start_row, p.end_row = None, None
            # page-locked surface: every device stores its bands straight into it
            group.host_register(px)
            try:
                px[:] = 0
                sg.render(p, pixels=px)
                assert (px == a["pixels"]).all()
            finally:
                group.host_unregister(px)
            # device framebuffer every member can store into (peer-mapped memory of device 0)
            d = group.device_alloc(p.width * p.height * 4)
            try:
                st = sg.render_device(p, d, want_stats=True)
                import torch

                got = _device_to_numpy(torch, d, p.height, p.width)
                assert (got == a["pixels"]).all()
                assert st.rays_primary == a["stats"].rays_primary
            finally:
                group.device_free(d)
        # the group partitions the rows itself
        p = scenario(resolution=32, band_height=4, band_count=2, band_index=0)
        with pytest.raises(lib.SoftRayError) as e:
            lib.Scene(group, [obj_mesh]).render(p)
        assert e.value.code == abi.E_INVALID_ARG
    finally:
        group.close()
        single.close()


def _device_to_numpy(torch, ptr, H, W):
    class _Holder:
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (H, W), "typestr": "<i4", "data": (int(ptr), False), "version": 3}
    return torch.as_tensor(h, device="cuda:0").cpu().numpy().view(np.uint32)


def test_group_of_one_device_equals_plain_context(lib, obj_mesh):
    compare_group_with_single(lib, 1, obj_mesh)


def test_group_of_two_devices_equals_one_device(lib, obj_mesh):
    if n_visible() < 2:
        pytest.skip("needs two visible CUDA devices")
    compare_group_with_single(lib, 2, obj_mesh)


def test_group_of_all_devices(lib, obj_mesh):
    if n_visible() < 3:
        pytest.skip("needs more than two visible CUDA devices")
    compare_group_with_single(lib, n_visible(), obj_mesh)


def test_create_multi_rejects_more_devices_than_visible(lib):
    L = lib.load()
    h = C.c_void_p()
    assert L.softray_create_multi(n_visible() + 1, C.byref(h)) == abi.E_INVALID_ARG
    assert L.softray_create_multi(-1, C.byref(h)) == abi.E_INVALID_ARG
    assert L.softray_device_count(None) == 0
