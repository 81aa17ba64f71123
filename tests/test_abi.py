"""CPU-side checks of the drop-in boundary: the shared library builds, loads and exports every
symbol include/softray_cuda.h declares, its PODs match the ctypes mirror, its pure-host helpers
agree with the oracle's twins, and without a CUDA device it fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from softray_b200 import abi, lib, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    lib.build()
    return lib.load()


def test_header_symbols_are_all_exported(L):
    hdr = open(os.path.join(ROOT, "include", "softray_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(softray_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(lib.EXPORTS)
    for name in declared:
        assert hasattr(L, name), name


def test_struct_layouts_match(L):
    assert L.softray_abi_version() == abi.ABI_VERSION
    for which, cls in enumerate([abi.Mesh, abi.Sphere, abi.SceneDesc, abi.Instance, abi.Frame, abi.Stats]):
        assert L.softray_abi_sizeof(which) == C.sizeof(cls)
    assert L.softray_abi_sizeof(99) == -1


def test_instance_init_and_frame_defaults_match_oracle(L):
    """Instance.InitRender matrices (Instance.cs:134-135) and the Renderer() defaults
    (Renderer.cs:207-230): product helper vs oracle twin, bit for bit."""
    import oracle

    O = oracle.lib()
    rng = np.random.default_rng(7)
    for _ in range(50):
        pos = (C.c_double * 3)(*rng.uniform(-3, 3, 3))
        yaw, pitch, roll = (float(v) for v in rng.uniform(-7, 7, 3))
        a, b = abi.Instance(), abi.Instance()
        L.softray_instance_init(C.byref(a), pos, yaw, pitch, roll, 2)
        O.orc_instance_init(C.byref(b), pos, yaw, pitch, roll, 2)
        assert bytes(a) == bytes(b)
    fa, fb = abi.Frame(), abi.Frame()
    L.softray_frame_defaults(C.byref(fa), 640, 480)
    O.orc_frame_defaults(C.byref(fb), 640, 480)
    assert bytes(fa) == bytes(fb)
    assert fa.shadow_samples == 100 and fa.random_seed == 1234567890 and fa.end_row == 479


def test_no_device_means_loud_failure(L):
    """There is no CPU fallback: without a GPU softray_create returns SOFTRAY_E_NO_DEVICE."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    assert L.softray_create(0, C.byref(h)) == abi.E_NO_DEVICE
    assert not h.value
    assert b"no CUDA device" in L.softray_last_error(None)
    with pytest.raises(lib.SoftRayError) as e:
        lib.Context(0)
    assert e.value.code == abi.E_NO_DEVICE
    # ... and so does the group context (softray_create_multi): no device, no context, no fallback
    assert L.softray_create_multi(0, C.byref(h)) == abi.E_NO_DEVICE and not h.value
    with pytest.raises(lib.SoftRayError) as e:
        lib.Context(n_devices=0)
    assert e.value.code == abi.E_NO_DEVICE
    assert L.softray_device_count(None) == 0


def test_null_arguments_are_rejected(L):
    assert L.softray_create(0, None) == abi.E_INVALID_ARG
    assert L.softray_scene_create(None, None, None) == abi.E_INVALID_ARG
    assert L.softray_render(None, None, None, None, None, None) == abi.E_INVALID_ARG
    assert L.softray_render_device(None, None, None, None, None, None, None) == abi.E_INVALID_ARG
    assert L.softray_scene_fingerprint(None, None) == abi.E_INVALID_ARG
    L.softray_destroy(None)
    L.softray_scene_destroy(None)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under softray_b200/ may import, load or link it."""
    pkg = os.path.join(ROOT, "softray_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cpp", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert "import oracle" not in text and "from oracle" not in text, fn
                assert "libsoftray_oracle" not in text and "softray_oracle.h" not in text, fn


def test_synthetic_scenes_are_deterministic_and_valid():
    m1, s1, _ = synth.config2(n_spheres=64)
    m2, s2, _ = synth.config2(n_spheres=64)
    assert (s1.cxyzr == s2.cxyzr).all() and (s1.argb == s2.argb).all()
    assert (np.abs(s1.cxyzr[:, :3]) <= 0.45).all() and (s1.cxyzr[:, 3] >= 0.01).all() and (s1.cxyzr[:, 3] <= 0.04).all()
    box = m1[0]
    assert box.n_tris == 12 and (box.bbox_min == -0.5).all() and (box.bbox_max == 0.5).all()
    # all 12 normals point into the room
    v = box.verts
    for a, b, c in box.tris:
        n = np.cross(v[b] - v[a], v[c] - v[a])
        assert np.dot(n, -v[a]) > 0
    hf = synth.height_field()
    assert hf.n_tris == 1_000_000
    assert (hf.bbox_min >= -0.5).all() and (hf.bbox_max <= 0.5).all()
    us = synth.uv_sphere()
    assert us.n_tris == 100_000 and np.isclose(np.abs(us.verts).max(), 0.5)
    assert len(synth.instance_grid()) == 100


def test_host_barrier_single_rank_and_timeout(L, monkeypatch):
    """softray_host_barrier: with one rank it returns at once (any number of generations); a rank that waits for
    one that never arrives gives up with SOFTRAY_E_TIMEOUT instead of spinning for ever."""
    import time

    words = np.zeros(2, dtype=np.uint32)
    for _ in range(5):
        assert L.softray_host_barrier(words.ctypes.data, 1) == abi.OK
    assert words[0] == 0 and words[1] == 5
    monkeypatch.setenv("SOFTRAY_BARRIER_TIMEOUT_S", "1")
    t = time.perf_counter()
    assert L.softray_host_barrier(words.ctypes.data, 2) == abi.E_TIMEOUT
    assert 0.9 < time.perf_counter() - t < 10.0
    assert L.softray_host_barrier(None, 2) == abi.E_INVALID_ARG
