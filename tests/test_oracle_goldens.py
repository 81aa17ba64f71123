"""Pins the CPU oracle against the reference's own golden images:
every raytrace golden BMP that exercises only on-path features must be reproduced with ZERO
differing pixels -- the same assertion as RendererTests.RenderAndTest (RendererTests.cs:540-543).
"""
import numpy as np
import pytest

import oracle
from tests.util import count_diff, golden_name, path_trace_spheres, scenario

# (resolution, kwargs) of RaytraceScenario calls in the reference's tests
ON_PATH = [
    (100, dict(shading=False)),
    (100, dict()),
    (100, dict(sub_pixel_res=2)),                                   # RaytraceAntialised
    (100, dict(sub_pixel_res=4)),
    (100, dict(sub_pixel_res=8)),
    (100, dict(shadows=True)),                                      # RaytraceDynamicShadow
    (100, dict(shading=False, shadows=True)),
    (100, dict(shading=False, sub_pixel_res=4)),
    (100, dict(shadows=True, sub_pixel_res=4)),
    (100, dict(shading=False, shadows=True, sub_pixel_res=4)),
    (100, dict(focal_blur=True, sub_pixel_res=2)),
    (100, dict(focal_blur=True, sub_pixel_res=4)),
    (100, dict(shading=False, focal_blur=True, sub_pixel_res=2)),
    (100, dict(shading=False, focal_blur=True, sub_pixel_res=4)),
    (100, dict(shadows=True, focal_blur=True, sub_pixel_res=2)),
    (100, dict(shadows=True, focal_blur=True, sub_pixel_res=4)),
    (100, dict(shading=False, shadows=True, focal_blur=True, sub_pixel_res=2)),
    (100, dict(shading=False, shadows=True, focal_blur=True, sub_pixel_res=4)),
    (50, dict(shadows=True, sub_pixel_res=4)),                      # RaytraceShadowAndAntiAlias
    (50, dict(shadows=True, focal_blur=True, sub_pixel_res=4)),     # RaytraceShadowAndFocalBlur
]


@pytest.fixture(scope="module")
def obj_scene(obj_mesh):
    return oracle.Scene([obj_mesh])


@pytest.mark.parametrize("res,kw", ON_PATH, ids=lambda v: str(v))
def test_obj3ds_goldens_bit_exact(fixtures, obj_scene, res, kw):
    name = golden_name(**kw)
    golden = fixtures[f"golden/{res}x{res}/{name}"]
    out = obj_scene.render(scenario(resolution=res, **kw))
    assert count_diff(out["pixels"], golden) == 0, name
    assert (out["pixels"] >> 24 == 0xFF).all()


def test_brute_force_path_matches_tree_path_on_goldens(fixtures, obj_scene):
    """rayTraceSubdivision=false (GeometryCollection) renders the same images."""
    for kw in (dict(shading=False), dict(), dict(shadows=True)):
        golden = fixtures["golden/100x100/" + golden_name(**kw)]
        out = obj_scene.render(scenario(subdivision=False, **kw))
        assert count_diff(out["pixels"], golden) == 0


def test_known_answer_counters(obj_scene):
    """10 000 primary rays, 6 716 hits; 671 600 shadow rays with dynamic shadows."""
    out = obj_scene.render(scenario(shadows=True), want_ids=True)
    st = out["stats"]
    assert st.rays_primary == 10000
    assert st.hits_primary == 6716
    assert st.rays_shadow == 671600
    assert int((out["ids"] >= 0).sum()) == 6716
    assert int((out["ids"] == -1).sum()) == 3284


def test_reference_tree_shape(obj_scene, obj2_mesh):
    """SURVEY App. C tree known answers for the two test models with the default limits."""
    assert obj_scene.tree_stats() == dict(depth=3, nodes=7, leaves=4, internal=3, refs=232, largest_leaf=76)
    s2 = oracle.Scene([obj2_mesh]).tree_stats()
    assert s2 == dict(depth=6, nodes=23, leaves=12, internal=11, refs=270, largest_leaf=30)


# ---- path tracing: pins Sphere.IntersectRay, the mixed rayFrac units, GeometryCollection order and
# ---- the per-block System.Random against the reference (oracle-only decorator)
def test_path_tracing_spheres_golden(fixtures, obj_mesh):
    """PathTracePrimitivesTest, first scenario (RendererTests.cs:248-261)."""
    sc = oracle.Scene([obj_mesh], spheres=path_trace_spheres())
    out = sc.render(scenario(shading=False, object_depth=3.0), options=oracle.default_options(path_tracing=1))
    golden = fixtures["golden/100x100/pathTracing_noShading_6_geometry"]
    assert count_diff(out["pixels"], golden) == 0
    assert int((golden == 0xFF00FF).sum()) == 400


@pytest.mark.parametrize("kw", [dict(), dict(sub_pixel_res=2), dict(sub_pixel_res=4), dict(sub_pixel_res=8),
                                dict(focal_blur=True, sub_pixel_res=2), dict(focal_blur=True, sub_pixel_res=4),
                                dict(focal_blur=True, sub_pixel_res=8)], ids=lambda v: str(v))
def test_path_tracing_obj2_goldens(fixtures, obj2_mesh, kw):
    """PathTraceTrianglesTest (RendererTests.cs:268-281): obj2.3DS, depthOfFocus = objectDepth."""
    sc = oracle.Scene([obj2_mesh])
    p = scenario(shading=False, focal_depth=1.0, **kw)
    out = sc.render(p, options=oracle.default_options(path_tracing=1))
    golden = fixtures["golden/100x100/" + golden_name(shading=False, path_tracing=True, **kw)]
    assert count_diff(out["pixels"], golden) == 0


@pytest.mark.parametrize("n,name", [(4, "pathTracing_noShading_focalBlurx4_7_geometry"),
                                    (8, "pathTracing_noShading_focalBlurx8_8_geometry")])
def test_path_tracing_spheres_focal_blur_goldens(fixtures, obj_mesh, n, name):
    """PathTracePrimitivesTest scenarios 2 and 3 (RendererTests.cs:262-265).  The shared
    GeometryCollection accumulates one more copy of the same tree per Renderer ("the triangle
    geometry ends up added twice/3x"); a duplicate can never win the strict `<`
    (GeometryCollection.cs:53), so one tree gives the same image."""
    sc = oracle.Scene([obj_mesh], spheres=path_trace_spheres())
    p = scenario(shading=False, object_depth=3.0, focal_blur=True, focal_depth=2.5, sub_pixel_res=n)
    out = sc.render(p, options=oracle.default_options(path_tracing=1))
    assert count_diff(out["pixels"], fixtures["golden/100x100/" + name]) == 0
