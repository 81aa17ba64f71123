"""meshes flattened and their trees built on the device (SOFTRAY_ACCEL_LBVH, sr_lbvh.cu).

The tree only decides WHICH primitives a ray is tested against; the winner, its rayFrac and the pixel come
from the exact layer, so a frame must not depend on the builder: every test renders the same frame through
the host-built SAH tree and the device-built tree (and, where the oracle is cheap, the oracle) and requires
identical pixels, hit ids and ray counters.  The device layout itself must be bit-identical across builds."""
import time

import numpy as np
import pytest

from softray_b200 import MeshData, SphereData, abi, synth
from tests.util import scenario

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from softray_b200 import lib as L

    return L


@pytest.fixture(scope="module")
def ctx(lib):
    c = lib.Context(0)
    yield c
    c.close()


def same_frame(lib, ctx, meshes, spheres, p, filter_modes=(abi.FILTER_AUTO,)):
    host = lib.Scene(ctx, meshes, spheres)
    dev = lib.Scene(ctx, meshes, spheres, accel=abi.ACCEL_LBVH)
    out = None
    for mode in filter_modes:
        p.filter_mode = mode
        a = host.render(p, want_ids=True)
        b = dev.render(p, want_ids=True)
        assert np.array_equal(a["pixels"], b["pixels"]), (mode, int((a["pixels"] != b["pixels"]).sum()))
        assert np.array_equal(a["ids"], b["ids"]), mode
        for k in ("rays_primary", "rays_shadow", "rays_secondary", "hits_primary", "shaded_hits"):
            assert getattr(a["stats"], k) == getattr(b["stats"], k), (mode, k)
        if mode == abi.FILTER_VERIFY:
            assert b["stats"].filter_mismatch == 0
        out = b
    p.filter_mode = abi.FILTER_AUTO
    host.close()
    dev.close()
    return out


@pytest.mark.parametrize("kw", [dict(), dict(shadows=True), dict(shadows=True, sub_pixel_res=2), dict(subdivision=False, shadows=True),
                                dict(shadows=True, point_lighting=False), dict(sub_pixel_res=3, focal_blur=True)],
                         ids=lambda v: str(v))
def test_obj3ds_device_tree_equals_host_tree(lib, ctx, obj_mesh, kw):
    same_frame(lib, ctx, [obj_mesh], None, scenario(resolution=96, **kw),
               (abi.FILTER_AUTO, abi.FILTER_OFF, abi.FILTER_VERIFY))


def test_obj2_3ds_boundary_faces(lib, ctx, obj2_mesh):
    same_frame(lib, ctx, [obj2_mesh], None, scenario(resolution=96, shadows=True), (abi.FILTER_AUTO, abi.FILTER_VERIFY))


def test_reference_goldens_through_the_device_tree(lib, ctx, fixtures, obj_mesh):
    """Every on-path golden image of the reference (RendererTests.cs) reproduced with the device-built tree."""
    from tests.test_cuda_parity import ON_PATH
    from tests.util import count_diff, golden_name

    sc = lib.Scene(ctx, [obj_mesh], accel=abi.ACCEL_LBVH)
    for res, kw in ON_PATH:
        want = fixtures[f"golden/{res}x{res}/{golden_name(**kw)}"]
        got = sc.render(scenario(resolution=res, **kw))
        assert count_diff(got["pixels"], want) == 0, (res, kw)
    sc.close()


def test_configs_small(lib, ctx):
    meshes, spheres, p = synth.config2(width=240, height=135, shadow_samples=12, n_spheres=200)
    same_frame(lib, ctx, meshes, spheres, p)
    meshes, _, p = synth.config3(width=240, height=135, nx=151, nz=81, shadow_samples=8)
    same_frame(lib, ctx, meshes, None, p, (abi.FILTER_AUTO, abi.FILTER_VERIFY))
    meshes, _, p = synth.config4(width=200, height=112, n_lon=40, n_lat=30, n_side=4, sub_pixel_res=2)
    same_frame(lib, ctx, meshes, None, p)
    meshes, _, p = synth.config5(width=200, height=112, n_lon=50, n_lat=30, n_side=4, shadow_samples=6)
    same_frame(lib, ctx, meshes, None, p, (abi.FILTER_AUTO, abi.FILTER_VERIFY))


def test_config3_against_oracle(lib, ctx):
    import oracle

    meshes, _, p = synth.config3(width=96, height=54, nx=61, nz=41, shadow_samples=5)
    got = lib.Scene(ctx, meshes, accel=abi.ACCEL_LBVH).render(p, want_ids=True)
    want = oracle.Scene(meshes).render(p, want_ids=True)
    assert np.array_equal(got["pixels"] & 0xFFFFFF, want["pixels"] & 0xFFFFFF)
    assert np.array_equal(got["ids"], want["ids"])


def soup(rng, n, size):
    c = rng.uniform(3 * n).reshape(n, 3) - 0.5
    e = (rng.uniform(6 * n).reshape(n, 2, 3) - 0.5) * size
    v = np.concatenate([c, c + e[:, 0], c + e[:, 1]]).clip(-0.5, 0.5)
    t = np.stack([np.arange(n), np.arange(n) + n, np.arange(n) + 2 * n], axis=1).astype(np.int32)
    return MeshData(v, t, synth.PALETTE[np.arange(n) % 8], v.min(axis=0), v.max(axis=0))


def test_random_soups(lib, ctx):
    rng = synth.SplitMix64(321)
    for n, size in ((1, 0.8), (2, 0.8), (3, 0.5), (17, 0.4), (300, 0.2), (6000, 0.05), (50000, 0.01)):
        mesh = soup(rng, n, size)
        u = rng.uniform(3)
        p = scenario(resolution=80, shadows=True, shadow_samples=9, yaw_deg=360.0 * u[0], pitch_deg=60.0 * u[1] - 30.0,
                     object_depth=1.0 + u[2])
        same_frame(lib, ctx, [mesh], None, p, (abi.FILTER_AUTO, abi.FILTER_VERIFY))


def test_coincident_and_degenerate_triangles(lib, ctx):
    """Equal Morton keys (many copies of the same triangle, ties broken by the sorted position), zero-area
    triangles and triangles in the faces of the root box: the lowest triangle index must still win ties
    (GeometryCollection.cs:53, SpatialSubdivision.cs:644)."""
    v = np.array([[-0.4, -0.4, 0.0], [0.4, -0.4, 0.0], [0.0, 0.4, 0.0], [0.0, 0.0, 0.0], [0.1, 0.0, 0.0], [0.2, 0.0, 0.0],
                  [-0.5, -0.5, -0.5], [0.5, -0.5, -0.5], [0.5, -0.5, 0.5], [-0.5, 0.5, 0.5]], dtype=np.float64)
    t = [[0, 1, 2]] * 70 + [[0, 2, 1]] * 70 + [[3, 4, 5], [3, 3, 3]] * 5 + [[6, 8, 7]]
    t = np.array(t, dtype=np.int32)
    argb = (0xFF000000 | (np.arange(len(t)) * 2654435761 & 0xFFFFFF)).astype(np.uint32)
    mesh = MeshData(v, t, argb, v.min(axis=0), v.max(axis=0))
    for yaw in (0.0, 135.0, 180.0):
        out = same_frame(lib, ctx, [mesh], None, scenario(resolution=64, shadows=True, shadow_samples=5, yaw_deg=yaw, object_depth=1.4),
                         (abi.FILTER_AUTO, abi.FILTER_OFF))
        hit = out["ids"][out["ids"] >= 0]
        assert hit.size == 0 or set(np.unique(hit)) <= {0, 70, 150}, np.unique(hit)


def test_empty_and_mixed_meshes(lib, ctx, obj_mesh):
    empty = MeshData(np.zeros((0, 3)), np.zeros((0, 3), np.int32), np.zeros(0, np.uint32), [-0.5] * 3, [0.5] * 3)
    out = lib.Scene(ctx, [empty], accel=abi.ACCEL_LBVH).render(scenario(resolution=40, shadows=True), want_ids=True)
    assert ((out["pixels"] & 0xFFFFFF) == 0xFF00FF).all() and (out["ids"] == -1).all()
    c = np.array([[0.2, 0.1, 0.0, 0.15], [-0.2, 0.0, 0.1, 0.1]])
    sph = SphereData(c, np.array([0xFF20C040, 0xFFC04020], dtype=np.uint32))
    same_frame(lib, ctx, [obj_mesh], sph, scenario(resolution=96, shadows=True, shadow_samples=10))


def test_device_layout_is_bit_identical_across_builds(lib, ctx):
    big = synth.height_field(201, 101)
    rng = synth.SplitMix64(5)
    s = soup(rng, 20000, 0.02)
    fps = {lib.Scene(ctx, [big, s], accel=abi.ACCEL_LBVH).fingerprint() for _ in range(4)}
    assert len(fps) == 1
    assert lib.Scene(ctx, [big, s]).fingerprint() not in fps
    assert lib.Scene(ctx, [big], accel=abi.ACCEL_LBVH).fingerprint() not in fps


def test_error_contracts(lib, ctx, obj_mesh):
    bad = MeshData(obj_mesh.verts, obj_mesh.tris, obj_mesh.argb, obj_mesh.bbox_min * 0.5, obj_mesh.bbox_max * 0.5)
    with pytest.raises(lib.SoftRayError) as e:
        lib.Scene(ctx, [bad], accel=abi.ACCEL_LBVH)
    assert e.value.code == abi.E_VERTEX_OUTSIDE_BBOX
    tris = obj_mesh.tris.copy()
    tris[5, 1] = len(obj_mesh.verts)
    bad = MeshData(obj_mesh.verts, tris, obj_mesh.argb, obj_mesh.bbox_min, obj_mesh.bbox_max)
    with pytest.raises(lib.SoftRayError) as e:
        lib.Scene(ctx, [bad], accel=abi.ACCEL_LBVH)
    assert e.value.code == abi.E_INVALID_ARG
    with pytest.raises(lib.SoftRayError) as e:
        lib.Scene(ctx, [obj_mesh], accel=7)
    assert e.value.code == abi.E_INVALID_ARG
    # the context is still usable afterwards
    lib.Scene(ctx, [obj_mesh], accel=abi.ACCEL_LBVH).render(scenario(resolution=16))


def test_one_million_triangles_build_and_trace(lib, ctx):
    """configs[2]'s mesh (1 M triangles): built on the device in milliseconds, traced identically."""
    meshes, _, p = synth.config3(width=480, height=270, shadow_samples=4)
    t0 = time.perf_counter()
    dev = lib.Scene(ctx, meshes, accel=abi.ACCEL_LBVH)
    t1 = time.perf_counter()
    host = lib.Scene(ctx, meshes)
    t2 = time.perf_counter()
    print(f"scene_create, {meshes[0].n_tris} triangles: device {1e3 * (t1 - t0):.1f} ms, host {1e3 * (t2 - t1):.1f} ms")
    a = host.render(p, want_ids=True)
    b = dev.render(p, want_ids=True)
    assert np.array_equal(a["pixels"], b["pixels"]) and np.array_equal(a["ids"], b["ids"])
    # (timings are printed, not asserted: a cold box has taken hundreds of ms for a first launch)
