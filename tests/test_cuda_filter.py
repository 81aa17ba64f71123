"""The FP32 filtered predicates (DESIGN.md "Filtered predicates") must never change a result: every
frame is rendered three ways through the C ABI -- filter + exact fallback (the default), exact FP64
reference arithmetic only, and VERIFY (both on every ray) -- and the pixels / hit ids must be
identical while the VERIFY run reports zero contradictions."""
import math

import numpy as np
import pytest

from softray_b200 import MeshData, SphereData, abi, synth
from tests.util import path_trace_spheres, scenario

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


def three_ways(scene, p, max_unsure_frac=0.05):
    out = {}
    for mode in (abi.FILTER_AUTO, abi.FILTER_OFF, abi.FILTER_VERIFY):
        p.filter_mode = mode
        out[mode] = scene.render(p, want_ids=True)
    p.filter_mode = abi.FILTER_AUTO
    auto, off, ver = out[abi.FILTER_AUTO], out[abi.FILTER_OFF], out[abi.FILTER_VERIFY]
    assert np.array_equal(auto["pixels"], off["pixels"]), int((auto["pixels"] != off["pixels"]).sum())
    assert np.array_equal(ver["pixels"], off["pixels"])
    assert np.array_equal(auto["ids"], off["ids"]) and np.array_equal(ver["ids"], off["ids"])
    assert ver["stats"].filter_mismatch == 0, ver["stats"].filter_mismatch
    for k in ("rays_primary", "rays_shadow", "rays_secondary", "hits_primary", "shaded_hits"):
        assert getattr(auto["stats"], k) == getattr(off["stats"], k), k
    assert off["stats"].filter_tests == 0 and off["stats"].filter_unsure == 0
    if auto["stats"].rays_shadow:
        assert auto["stats"].filter_tests > 0 or auto["stats"].filter_unsure == 0   # (all decided at the root box)
        frac = auto["stats"].filter_unsure / auto["stats"].rays_shadow
        assert frac <= max_unsure_frac, f"{frac:.4f} of the shadow rays fell back to the exact path"
    return auto


def test_obj3ds_shadows(lib, ctx, obj_mesh):
    sc = lib.Scene(ctx, [obj_mesh])
    for kw in (dict(shadows=True), dict(shadows=True, subdivision=False), dict(shadows=True, sub_pixel_res=2),
               dict(shadows=True, point_lighting=False), dict(shadows=True, yaw_deg=20.0, pitch_deg=35.0, object_depth=0.8)):
        three_ways(sc, scenario(resolution=96, **kw))


def test_obj2_3ds_shadows(lib, ctx, obj2_mesh):
    """obj2.3DS has flat faces lying in the faces of its own bounding box: a shadow ray that enters the
    root box through such a face is clipped onto the triangle's plane (SpatialSubdivision.cs:389-401) and
    whether it 'hits' it is decided by the last bit of the FP64 clip -- the filter must leave exactly
    those rays to the reference arithmetic (about 5 % here)."""
    three_ways(lib.Scene(ctx, [obj2_mesh]), scenario(resolution=96, shadows=True), max_unsure_frac=0.15)


def test_light_inside_the_bounding_box(lib, ctx, obj_mesh):
    """Shadow rays that start inside the root box are not clipped (SpatialSubdivision.cs:389-398)."""
    p = scenario(resolution=80, shadows=True, shadow_samples=40)
    p.light_pos_view = (0.05, 0.1, 1.0)       # view space; the object sits at depth 1.0
    three_ways(lib.Scene(ctx, [obj_mesh]), p, max_unsure_frac=0.2)


def test_config2_spheres_room(lib, ctx):
    meshes, spheres, p = synth.config2(width=160, height=90, shadow_samples=100, n_spheres=1000)
    auto = three_ways(lib.Scene(ctx, meshes, spheres), p, max_unsure_frac=0.001)
    assert auto["stats"].sphere_tests < auto["stats"].rays_shadow     # spheres cannot shadow here: light too far


def test_spheres_that_can_shadow(lib, ctx, obj_mesh):
    """A light within 1.0 of the spheres: the sphere part of every shadow ray is exact."""
    sph = path_trace_spheres()
    p = scenario(resolution=80, object_depth=3.0, shadows=True, shadow_samples=10)
    three_ways(lib.Scene(ctx, [obj_mesh], sph), p, max_unsure_frac=1.0)
    meshes, spheres, p = synth.config2(width=96, height=54, shadow_samples=20, n_spheres=300)
    p.light_pos_view = (0.1, 0.2, 1.3)
    auto = three_ways(lib.Scene(ctx, meshes, spheres), p, max_unsure_frac=0.2)
    assert auto["stats"].sphere_tests > 0


def test_config3_heightfield_small_triangles(lib, ctx):
    meshes, _, p = synth.config3(width=160, height=90, nx=301, nz=201, shadow_samples=16)
    three_ways(lib.Scene(ctx, meshes), p)


def test_config5_flattened_grid(lib, ctx):
    meshes, _, p = synth.config5(width=160, height=90, n_lon=60, n_lat=40, n_side=4, shadow_samples=8)
    three_ways(lib.Scene(ctx, meshes), p)


def test_degenerate_and_boundary_triangles(lib, ctx):
    """Zero-area triangles (never hit, Triangle.cs:42-43), triangles lying in the faces of the root
    box (hits at the clipped start) and an axis-parallel light direction."""
    v = np.array([[-0.5, -0.5, -0.5], [0.5, -0.5, -0.5], [0.5, -0.5, 0.5], [-0.5, -0.5, 0.5],      # floor y = -0.5
                  [-0.2, 0.5, -0.2], [0.2, 0.5, -0.2], [0.2, 0.5, 0.2], [-0.2, 0.5, 0.2],          # lid y = +0.5
                  [0.0, 0.0, 0.0], [0.1, 0.0, 0.0], [0.2, 0.0, 0.0]], dtype=np.float64)            # collinear
    t = np.array([[0, 2, 1], [0, 3, 2], [4, 6, 5], [4, 7, 6], [4, 5, 6], [4, 6, 7], [8, 9, 10], [8, 8, 8]], dtype=np.int32)
    mesh = MeshData(v, t, np.full(len(t), 0xFFC8C8C8, dtype=np.uint32), v.min(axis=0), v.max(axis=0))
    sc = lib.Scene(ctx, [mesh])
    for yaw, pitch in ((135.0, -22.0), (0.0, -89.0), (45.0, 60.0)):
        p = scenario(resolution=72, shadows=True, shadow_samples=24, yaw_deg=yaw, pitch_deg=pitch, object_depth=1.6)
        three_ways(sc, p, max_unsure_frac=1.0)
    p = scenario(resolution=72, shadows=True, shadow_samples=24, object_depth=1.6, point_lighting=False)
    p.light_dir_view = (0.0, -1.0, 0.0)
    three_ways(sc, p, max_unsure_frac=1.0)


def test_random_soups(lib, ctx):
    """Random triangle soups of very different triangle sizes, random cameras."""
    rng = synth.SplitMix64(99)
    for n, size in ((200, 0.3), (5000, 0.05), (40000, 0.01)):
        c = rng.uniform(3 * n).reshape(n, 3) - 0.5
        e = (rng.uniform(6 * n).reshape(n, 2, 3) - 0.5) * size
        v = np.concatenate([c, c + e[:, 0], c + e[:, 1]]).clip(-0.5, 0.5)
        t = np.stack([np.arange(n), np.arange(n) + n, np.arange(n) + 2 * n], axis=1).astype(np.int32)
        mesh = MeshData(v, t, synth.PALETTE[np.arange(n) % 8], v.min(axis=0), v.max(axis=0))
        sc = lib.Scene(ctx, [mesh])
        u = rng.uniform(3)
        p = scenario(resolution=96, shadows=True, shadow_samples=16, yaw_deg=360.0 * u[0], pitch_deg=60.0 * u[1] - 30.0,
                     object_depth=1.0 + u[2])
        three_ways(sc, p, max_unsure_frac=0.1)


def test_bundles_resolve_unoccluded_shading_points(lib, ctx):
    """config2: the light is outside the room, no triangle can occlude -> every shading point's 100 rays
    are answered by one cone test; with an occluder in the way only part of them are."""
    meshes, spheres, p = synth.config2(width=160, height=90, shadow_samples=100, n_spheres=1000)
    auto = three_ways(lib.Scene(ctx, meshes, spheres), p, max_unsure_frac=0.001)
    assert auto["stats"].rays_bundled == auto["stats"].rays_shadow
    mesh = synth.height_field(61, 41)
    p = scenario(resolution=96, shadows=True, shadow_samples=32, object_depth=1.3)
    auto = three_ways(lib.Scene(ctx, [mesh]), p)
    assert 0 < auto["stats"].rays_bundled <= auto["stats"].rays_shadow


def test_closest_hit_filter_primary_and_reflection(lib, ctx, obj_mesh, obj2_mesh):
    """Camera and mirror rays: the filter names the winning triangle, the reference arithmetic evaluates
    only that one (or everything, when the filter cannot separate the candidates)."""
    for mesh in (obj_mesh, obj2_mesh):
        sc = lib.Scene(ctx, [mesh])
        for kw in (dict(), dict(subdivision=False), dict(sub_pixel_res=3, focal_blur=False), dict(sub_pixel_res=2, focal_blur=True),
                   dict(reflection_depth=2, texture3d_id=1), dict(object_depth=0.3), dict(yaw_deg=0.0, pitch_deg=0.0)):
            auto = three_ways(sc, scenario(resolution=128, **kw))
            assert auto["stats"].filter_unsure <= 0.25 * auto["stats"].rays_primary + 50   # (obj2: faces in the root box faces)
    meshes, _, p = synth.config3(width=200, height=112, nx=301, nz=201, shadow_samples=4)
    auto = three_ways(lib.Scene(ctx, meshes), p)
    assert auto["stats"].rays_secondary > 0
    meshes, _, p = synth.config4(width=160, height=90, n_lon=60, n_lat=40, n_side=4, sub_pixel_res=2)
    three_ways(lib.Scene(ctx, meshes), p)


def test_camera_inside_the_bounding_box(lib, ctx):
    """A camera inside the room: rays start inside the root box (no clip, SpatialSubdivision.cs:389-398)."""
    meshes, spheres, p = synth.config2(width=128, height=72, shadow_samples=8, n_spheres=200)
    p.instances[0].position = (0.0, 0.0, 0.2)
    three_ways(lib.Scene(ctx, meshes, spheres), p, max_unsure_frac=1.0)
