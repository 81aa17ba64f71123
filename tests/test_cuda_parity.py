"""Parity of the CUDA path (through the C ABI, softray_b200.lib) against the reference's golden
images and against the CPU oracle.  Bars (BASELINE.json north_star): primary-hit ids bit-exact on
non-grazing rays (|cos theta| > 1e-4), 8-bit colours within +-1 LSB on >= 99.9 % of pixels.  The
kernel uses the reference's FP64 operations in the reference's order, so most cases are held to
the stronger bar of ZERO differing pixels, like RendererTests.RenderAndTest (RendererTests.cs:540-543).
"""
import math

import numpy as np
import pytest

from softray_b200 import FrameParams, InstanceData, MeshData, SphereData, abi, multi_gpu, synth
from tests.util import channel_absdiff, count_diff, golden_name, path_trace_spheres, scenario

pytestmark = pytest.mark.gpu

COS_GRAZING = 1e-4


@pytest.fixture(scope="module")
def lib():
    from softray_b200 import lib as L

    return L


@pytest.fixture(scope="module")
def ctx(lib):
    c = lib.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def obj_scene(lib, ctx, obj_mesh):
    return lib.Scene(ctx, [obj_mesh])


@pytest.fixture(scope="module")
def obj_oracle(obj_mesh):
    import oracle

    return oracle.Scene([obj_mesh])


def assert_parity(got, want, exact=True, what=""):
    """got: lib.Scene.render dict, want: oracle.Scene.render dict (with ids + aux)."""
    gp, wp = got["pixels"], want["pixels"]
    assert ((gp >> 24) == 0xFF).all(), what
    n = gp.size
    if got.get("ids") is not None and want.get("ids") is not None:
        bad = got["ids"] != want["ids"]
        if want.get("cos_theta") is not None:
            cos = np.nan_to_num(np.abs(want["cos_theta"]), nan=1.0)
            bad &= cos > COS_GRAZING
        assert int(bad.sum()) == 0, f"{what}: {int(bad.sum())} non-grazing hit ids differ"
    d = channel_absdiff(gp, wp)
    n_diff = int((d > 0).sum())
    if exact:
        assert n_diff == 0, f"{what}: {n_diff} of {n} pixels differ (max {int(d.max())})"
    else:
        assert int((d > 1).sum()) <= 0.001 * n, f"{what}: {(d > 1).sum()} pixels differ by more than 1 LSB"


# ------------------------------------------------------------------------------------------------
# the reference's own golden images
# ------------------------------------------------------------------------------------------------
ON_PATH = [
    (100, dict(shading=False)),
    (100, dict()),
    (100, dict(sub_pixel_res=2)),
    (100, dict(sub_pixel_res=4)),
    (100, dict(sub_pixel_res=8)),
    (100, dict(shadows=True)),
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
    (50, dict(shadows=True, sub_pixel_res=4)),
    (50, dict(shadows=True, focal_blur=True, sub_pixel_res=4)),
]


@pytest.mark.parametrize("res,kw", ON_PATH, ids=lambda v: str(v))
def test_reference_goldens(fixtures, obj_scene, res, kw):
    """Every on-path golden BMP of the reference straight from the CUDA kernel."""
    name = golden_name(**kw)
    golden = fixtures[f"golden/{res}x{res}/{name}"]
    out = obj_scene.render(scenario(resolution=res, **kw))
    assert count_diff(out["pixels"], golden) == 0, name
    assert (out["pixels"] >> 24 == 0xFF).all()


def test_known_answer_counters(obj_scene):
    """10 000 primary rays, 6 716 hits, 671 600 shadow rays."""
    out = obj_scene.render(scenario(shadows=True), want_ids=True)
    st = out["stats"]
    assert (st.rays_primary, st.hits_primary, st.rays_shadow) == (10000, 6716, 671600)
    assert int((out["ids"] >= 0).sum()) == 6716 and int((out["ids"] == -1).sum()) == 3284
    assert st.launches == 1 and st.ms_kernel > 0


@pytest.mark.parametrize("kw", [dict(), dict(shadows=True), dict(subdivision=False), dict(subdivision=False, shadows=True)],
                         ids=lambda v: str(v))
def test_brute_accel_equals_bvh(lib, ctx, obj_mesh, obj_scene, kw):
    """SOFTRAY_ACCEL_BRUTE (GeometryCollection order) and the BVH give identical frames and ids."""
    brute = lib.Scene(ctx, [obj_mesh], accel=abi.ACCEL_BRUTE)
    p = scenario(resolution=96, **kw)
    a = obj_scene.render(p, want_ids=True)
    b = brute.render(p, want_ids=True)
    assert count_diff(a["pixels"], b["pixels"]) == 0
    assert (a["ids"] == b["ids"]).all()
    brute.close()


# ------------------------------------------------------------------------------------------------
# oracle comparisons beyond the goldens
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kw", [
    dict(resolution=512, specular_lighting=False),                                  # configs[0]
    dict(resolution=256, shadows=True, point_lighting=False),                       # directional light
    dict(resolution=200, shadows=True, shadow_samples=7),
    dict(width=333, height=77),                                                     # ragged tiles
    dict(resolution=160, pitch_deg=35.0, yaw_deg=20.0, roll_deg=50.0, object_depth=0.3),   # camera inside the box
    dict(resolution=128, subdivision=False, shadows=True),
    dict(resolution=64, sub_pixel_res=3, focal_blur=True, shadows=True, focal_strength=25.0),
], ids=lambda v: str(v))
def test_obj3ds_against_oracle(obj_scene, obj_oracle, kw):
    p = scenario(**kw)
    got = obj_scene.render(p, want_ids=True)
    want = obj_oracle.render(p, want_ids=True, want_aux=True)
    assert_parity(got, want, what=str(kw))
    for k in ("rays_primary", "rays_shadow", "hits_primary"):
        assert getattr(got["stats"], k) == getattr(want["stats"], k), k


def test_spheres_and_mesh_mixed_units(lib, ctx, obj_mesh):
    """PathTracePrimitivesTest's scene (RendererTests.cs:251-257) without the path tracer: sphere
    rayFrac is a distance, triangle rayFrac a multiple of |dir|, list order."""
    import oracle

    sph = path_trace_spheres()
    for kw in (dict(shading=False), dict(), dict(shadows=True, shadow_samples=10)):
        p = scenario(resolution=120, object_depth=3.0, **kw)
        got = lib.Scene(ctx, [obj_mesh], sph).render(p, want_ids=True)
        want = oracle.Scene([obj_mesh], sph).render(p, want_ids=True, want_aux=True)
        assert_parity(got, want, what=str(kw))
        assert int((want["ids"] <= -2).sum()) > 1000


def test_config2_small_against_oracle(lib, ctx):
    import oracle

    meshes, spheres, p = synth.config2(width=240, height=135, shadow_samples=100, n_spheres=1000)
    got = lib.Scene(ctx, meshes, spheres).render(p, want_ids=True)
    want = oracle.Scene(meshes, spheres).render(p, want_ids=True, want_aux=True)
    assert_parity(got, want, what="config2")
    assert got["stats"].rays_shadow == want["stats"].rays_shadow == 100 * want["stats"].hits_primary


def test_config2_brute_equals_bvh(lib, ctx):
    meshes, spheres, p = synth.config2(width=96, height=54, shadow_samples=16, n_spheres=300)
    a = lib.Scene(ctx, meshes, spheres).render(p, want_ids=True)
    b = lib.Scene(ctx, meshes, spheres, accel=abi.ACCEL_BRUTE).render(p, want_ids=True)
    assert count_diff(a["pixels"], b["pixels"]) == 0 and (a["ids"] == b["ids"]).all()


def test_sphere_set_staged_in_shared_memory(lib, ctx):
    """A small sphere set is walked from a shared-memory copy of its tree and filter records (north_star: "shared-memory
    staging of small sphere sets"); the frame is the oracle's, and the one without staging."""
    import os

    import oracle

    meshes, spheres, p = synth.config2(width=200, height=112, shadow_samples=10, n_spheres=300)
    sc = lib.Scene(ctx, meshes, spheres)
    staged = sc.render(p, want_ids=True)
    os.environ["SOFTRAY_STAGE_SPHERES_MAX"] = "0"
    try:
        plain = sc.render(p, want_ids=True)
    finally:
        os.environ.pop("SOFTRAY_STAGE_SPHERES_MAX")
    assert (staged["pixels"] == plain["pixels"]).all() and (staged["ids"] == plain["ids"]).all()
    assert staged["stats"].node_visits == plain["stats"].node_visits
    want = oracle.Scene(meshes, spheres).render(p, want_ids=True, want_aux=True)
    assert_parity(staged, want, what="staged spheres")
    assert int((want["ids"] <= -2).sum()) > 2000


def test_config3_small_against_oracle(lib, ctx):
    """Shadows + 2-bounce mirror reflection + Texture3D (oracle-defined extensions)."""
    import oracle

    meshes, _, p = synth.config3(width=200, height=112, nx=121, nz=81, shadow_samples=12)
    got = lib.Scene(ctx, meshes).render(p, want_ids=True)
    opt = oracle.default_options(tree_max_depth=15, tree_max_per_node=25)
    want = oracle.Scene(meshes, options=opt).render(p, want_ids=True, want_aux=True)
    assert_parity(got, want, what="config3")
    assert got["stats"].rays_secondary == want["stats"].rays_secondary > 0


def test_config4_small_against_oracle(lib, ctx):
    """Composite instances (nearest hit across instances) with 2x2 supersampling."""
    import oracle

    meshes, _, p = synth.config4(width=160, height=90, n_lon=40, n_lat=30, n_side=4, sub_pixel_res=2)
    got = lib.Scene(ctx, meshes).render(p, want_ids=True)
    want = oracle.Scene(meshes).render(p, want_ids=True, want_aux=True)
    assert_parity(got, want, what="config4")


def test_config5_small_against_oracle(lib, ctx):
    import oracle

    meshes, _, p = synth.config5(width=192, height=108, n_lon=40, n_lat=30, n_side=4, shadow_samples=3)
    got = lib.Scene(ctx, meshes).render(p, want_ids=True)
    want = oracle.Scene(meshes).render(p, want_ids=True, want_aux=True)
    assert_parity(got, want, what="config5")


# ------------------------------------------------------------------------------------------------
# row ranges, row bands, determinism, errors
# ------------------------------------------------------------------------------------------------
def test_only_requested_rows_are_written(obj_scene, obj_oracle):
    """rayTraceStartRow/EndRow (Renderer.cs:134-136): exactly those rows."""
    p = scenario(resolution=100, start_row=37, end_row=58)
    px = np.full((100, 100), 0xDEADBEEF, dtype=np.uint32)
    ids = np.full((100, 100), 12345, dtype=np.int32)
    obj_scene.render(p, pixels=px, ids=ids)
    full = obj_oracle.render(scenario(resolution=100))["pixels"]
    assert (px[:37] == 0xDEADBEEF).all() and (px[59:] == 0xDEADBEEF).all()
    assert (ids[:37] == 12345).all() and (ids[59:] == 12345).all()
    assert count_diff(px[37:59], full[37:59]) == 0


@pytest.mark.parametrize("bands,band_height", [(2, 8), (3, 5), (8, 16), (4, 1)])
def test_row_bands_tile_the_frame(obj_scene, bands, band_height):
    """The union of all ranks' bands is the full frame and no band touches another's rows."""
    W, H = 96, 83
    full = obj_scene.render(scenario(width=W, height=H, shadows=True, shadow_samples=5), want_ids=True)
    px = np.full((H, W), 0xDEADBEEF, dtype=np.uint32)
    total_primary = 0
    for r in range(bands):
        mine = np.full((H, W), 0xDEADBEEF, dtype=np.uint32)
        out = obj_scene.render(scenario(width=W, height=H, shadows=True, shadow_samples=5, band_height=band_height,
                                        band_count=bands, band_index=r), pixels=mine)
        rows = [y for y in range(H) if (y // band_height) % bands == r]
        other = [y for y in range(H) if (y // band_height) % bands != r]
        assert (mine[other] == 0xDEADBEEF).all()
        px[rows] = mine[rows]
        total_primary += out["stats"].rays_primary
    assert count_diff(px, full["pixels"]) == 0 and (px >> 24 == 0xFF).all()
    assert total_primary == W * H


def test_scene_layout_is_bit_identical_across_builds(lib, ctx):
    meshes, spheres, _ = synth.config2(n_spheres=500)
    big = synth.height_field(101, 51)
    fps = {lib.Scene(ctx, meshes + [big], spheres).fingerprint() for _ in range(3)}
    assert len(fps) == 1
    assert lib.Scene(ctx, meshes, spheres).fingerprint() not in fps


def test_render_is_deterministic(lib, ctx):
    meshes, spheres, p = synth.config2(width=320, height=180, shadow_samples=10)
    sc = lib.Scene(ctx, meshes, spheres)
    a = sc.render(p, want_ids=True)
    b = sc.render(p, want_ids=True)
    assert (a["pixels"] == b["pixels"]).all() and (a["ids"] == b["ids"]).all()
    assert a["stats"].rays == b["stats"].rays


def test_error_contracts(lib, ctx, obj_mesh):
    """The .NET exceptions of the reference map to SOFTRAY_E_* codes (include/softray_cuda.h)."""
    # SpatialSubdivision ctor: "A triangle vertex is outside the bounding box" (SpatialSubdivision.cs:285-295)
    bad = MeshData(obj_mesh.verts, obj_mesh.tris, obj_mesh.argb, obj_mesh.bbox_min * 0.5, obj_mesh.bbox_max * 0.5)
    with pytest.raises(lib.SoftRayError) as e:
        lib.Scene(ctx, [bad])
    assert e.value.code == abi.E_VERTEX_OUTSIDE_BBOX
    sc = lib.Scene(ctx, [obj_mesh])
    for kw, code in ((dict(sub_pixel_res=0), abi.E_INVALID_ARG), (dict(reflection_depth=9), abi.E_INVALID_ARG),
                     (dict(shadows=True, shadow_samples=0), abi.E_INVALID_ARG),
                     (dict(texture3d_id=7), abi.E_INVALID_ARG),
                     (dict(band_count=4, band_height=4, band_index=4), abi.E_INVALID_ARG)):
        with pytest.raises(lib.SoftRayError) as e:
            sc.render(scenario(resolution=16, **kw))
        assert e.value.code == code, kw
    p = scenario(resolution=16, shadows=True)
    p.instances = p.instances * 2
    with pytest.raises(lib.SoftRayError) as e:
        sc.render(p)
    assert e.value.code == abi.E_UNSUPPORTED
    p = scenario(resolution=16)
    p.instances[0].mesh_id = 3
    with pytest.raises(lib.SoftRayError) as e:
        sc.render(p)
    assert e.value.code == abi.E_INVALID_ARG


def test_empty_row_range_renders_nothing(lib, ctx, obj_scene):
    """start_row > end_row after the clamp (Renderer.cs:1652-1653): the reference's row loop runs zero times.
    Nothing is written, every counter is zero, nothing divides by the empty band (both pipelines, both entry points)."""
    import os

    for pl in ("fused", "wave"):
        os.environ["SOFTRAY_PIPELINE"] = pl
        try:
            for kw in (dict(start_row=5, end_row=4), dict(start_row=90, end_row=3, band_height=4, band_count=2, band_index=1)):
                px = np.full((100, 100), 0xDEADBEEF, dtype=np.uint32)
                out = obj_scene.render(scenario(resolution=100, shadows=True, **kw), pixels=px, want_ids=True)
                assert (px == 0xDEADBEEF).all() and (out["ids"] == -1).all()
                assert out["stats"].rays == 0 and out["stats"].launches == 0
        finally:
            os.environ.pop("SOFTRAY_PIPELINE", None)


def test_empty_scene_and_empty_mesh(lib, ctx):
    """A Model with no triangles renders the background everywhere."""
    empty = MeshData(np.zeros((0, 3)), np.zeros((0, 3), np.int32), np.zeros(0, np.uint32), [-0.5] * 3, [0.5] * 3)
    out = lib.Scene(ctx, [empty]).render(scenario(resolution=40, shadows=True), want_ids=True)
    assert ((out["pixels"] & 0xFFFFFF) == 0xFF00FF).all() and (out["ids"] == -1).all()
    assert out["stats"].hits_primary == 0 and out["stats"].rays_shadow == 0


# ------------------------------------------------------------------------------------------------
# full-size properties (BASELINE.json sizes; the oracle only checks a sampled sub-range)
# ------------------------------------------------------------------------------------------------
def test_config2_full_size_rows_against_oracle(lib, ctx):
    """configs[1] at 1920x1080 with 100 shadow rays per hit: the oracle renders three row ranges
    through its own start_row/end_row (the reference's mechanism, Renderer.cs:134-136)."""
    import oracle

    meshes, spheres, p = synth.config2()
    sc = lib.Scene(ctx, meshes, spheres)
    got = sc.render(p, want_ids=True)
    st = got["stats"]
    assert st.rays_primary == 1920 * 1080 and st.rays_shadow == 100 * st.hits_primary
    orc = oracle.Scene(meshes, spheres)
    for top in (0, 539, 1078):
        p.start_row, p.end_row = top, top + 1
        want = orc.render(p, want_ids=True, want_aux=True)
        sl = slice(top, top + 2)
        sub = dict(pixels=got["pixels"][sl], ids=got["ids"][sl])
        wsub = dict(pixels=want["pixels"][sl], ids=want["ids"][sl], cos_theta=want["cos_theta"][sl])
        assert_parity(sub, wsub, what=f"rows {top}..{top + 1}")


def _window_parity(got, orc, p, windows, what):
    """Windows (row0, row1, col0, col1) of a FULL-SIZE frame against the oracle: the oracle traces exactly those
    pixels of the same frame (its own start_row / end_row, Renderer.cs:134-136, plus its test-only column window:
    the reference tree degenerates to brute force at these sizes, a whole row takes minutes)."""
    import oracle

    for r0, r1, c0, c1 in windows:
        p.start_row, p.end_row = r0, r1
        want = orc.render(p, options=oracle.default_options(col_start=c0, col_end=c1), want_ids=True, want_aux=True)
        sl = (slice(r0, r1 + 1), slice(c0, c1 + 1))
        sub = dict(pixels=got["pixels"][sl], ids=got["ids"][sl])
        wsub = dict(pixels=want["pixels"][sl], ids=want["ids"][sl], cos_theta=want["cos_theta"][sl])
        assert_parity(sub, wsub, what=f"{what} rows {r0}..{r1} cols {c0}..{c1}")
        assert int((want["ids"][sl] >= 0).sum()) > 0.3 * wsub["ids"].size, "the window should look at geometry"
    p.start_row, p.end_row = None, None


def test_config3_full_size_windows_against_oracle(lib, ctx):
    """configs[2] exactly as named -- 1M triangles, 3840x2160, 100 soft-shadow rays per hit, 2 mirror bounces,
    Texture3D -- rendered once at full size; three windows of it against the oracle."""
    import oracle

    meshes, _, p = synth.config3()
    got = lib.Scene(ctx, meshes).render(p, want_ids=True)
    st = got["stats"]
    assert st.rays_primary == 3840 * 2160 and st.rays_shadow == 100 * st.shaded_hits and st.rays_secondary > 0
    assert st.launches > 1, "the 4K configuration should run the stage kernels"
    orc = oracle.Scene(meshes)
    _window_parity(got, orc, p, [(700, 701, 1850, 1977), (1400, 1401, 1000, 1127), (1079, 1080, 2600, 2727), (1700, 1701, 2000, 2063)], "config3")


def test_config5_full_size_windows_against_oracle(lib, ctx):
    """configs[4] exactly as named -- 10M triangles, 7680x4320, Phong + 1 shadow ray per hit: windows of the full
    frame against the oracle, and of the same frame rendered as two interleaved row bands."""
    import oracle

    meshes, _, p = synth.config5()
    sc = lib.Scene(ctx, meshes)
    got = sc.render(p, want_ids=True)
    st = got["stats"]
    assert st.rays_primary == 7680 * 4320 and st.rays_shadow == st.hits_primary
    banded = dict(pixels=np.zeros_like(got["pixels"]), ids=np.full_like(got["ids"], -7))
    for r in range(2):
        p.band_height, p.band_count, p.band_index = 36, 2, r
        sc.render(p, pixels=banded["pixels"], ids=banded["ids"])
    p.band_height, p.band_count, p.band_index = 0, 1, 0
    assert (banded["pixels"] == got["pixels"]).all() and (banded["ids"] == got["ids"]).all()
    orc = oracle.Scene(meshes)
    _window_parity(got, orc, p, [(2160, 2163, 3800, 4055), (1200, 1201, 4900, 5155), (3300, 3303, 1900, 2027)], "config5")


def test_config3_full_size_band_invariance(lib, ctx):
    """1M triangles at 3840x2160 (1 shadow sample to bound the time): two interleaved bands
    reproduce the unbanded frame exactly."""
    meshes, _, p = synth.config3(shadow_samples=1)
    sc = lib.Scene(ctx, meshes)
    full = sc.render(p)
    px = np.zeros_like(full["pixels"])
    for r in range(2):
        p.band_height, p.band_count, p.band_index = 32, 2, r
        sc.render(p, pixels=px)
    assert (px == full["pixels"]).all()
    assert full["stats"].rays_secondary > 0 and full["stats"].hits_primary > 1_000_000


# ------------------------------------------------------------------------------------------------
# the step after the path: PostProcessImage + AntiAliasImage on the device
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("aa", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("style", [0, 1, 2])
def test_resolve_matches_oracle(ctx, aa, style):
    import oracle

    rng = synth.SplitMix64(1000 + 10 * aa + style)
    for h, w in ((37, 53), (41, 132)):                       # ragged scalar path; 128-bit vector path (w % 4 == 0)
        src = (rng.next_u64(h * aa * w * aa) & np.uint64(0xFFFFFFFF)).astype(np.uint32).reshape(h * aa, w * aa)
        src[::3, ::5] = 0x00FF00FF                           # the 24-bit background colour itself
        src[1::4, 2::7] |= 0xFF000000
        got = ctx.resolve(src, aa_res=aa, style=style, background=0xFF00FF)
        want = oracle.resolve(src, aa_res=aa, style=style, background=0xFF00FF)
        assert np.array_equal(got, want), (h, w)


def test_resolve_of_a_supersampled_render_equals_antialiasimage(lib, ctx, obj_scene, obj_oracle):
    """AntiAliasResolution = 2: the frame is traced at 2x the size, then box-filtered (Renderer.cs:366-410,937-978)."""
    import oracle

    p = scenario(resolution=128)
    hi = obj_scene.render(p)["pixels"]
    got = ctx.resolve(hi, aa_res=2)
    want = oracle.resolve(obj_oracle.render(p)["pixels"], aa_res=2)
    assert np.array_equal(got, want) and got.shape == (64, 64) and ((got >> 24) == 0xFF).all()


def test_resolve_error_contracts(lib, ctx):
    src = np.zeros((8, 8), dtype=np.uint32)
    for kw in (dict(style=3), dict(style=5)):
        with pytest.raises(lib.SoftRayError) as e:
            ctx.resolve(src, **kw)
        assert e.value.code == abi.E_UNSUPPORTED
    with pytest.raises(lib.SoftRayError) as e:
        ctx.resolve(src, aa_res=0)
    assert e.value.code == abi.E_INVALID_ARG


def test_pinned_host_framebuffer_is_written_by_the_kernel_itself(obj_scene):
    """softray_render stores straight into a page-locked caller buffer (no device framebuffer, no D2H copy);
    a pageable buffer goes through the staging copy.  Same pixels, same ids, and only the requested rows."""
    import torch

    p = scenario(resolution=160, shadows=True, shadow_samples=8, start_row=11, end_row=140)
    want = obj_scene.render(p, want_ids=True, pixels=np.full((160, 160), 0xABCDEF01, dtype=np.uint32),
                            ids=np.full((160, 160), 777, dtype=np.int32))
    px = torch.full((160, 160), 0xABCDEF01 - (1 << 32), dtype=torch.int32).pin_memory()
    ids = torch.full((160, 160), 777, dtype=torch.int32).pin_memory()
    got = obj_scene.render(p, want_ids=True, pixels=px.numpy().view(np.uint32), ids=ids.numpy())
    assert np.array_equal(got["pixels"], want["pixels"]) and np.array_equal(got["ids"], want["ids"])
    assert (got["pixels"][:11] == 0xABCDEF01).all() and (got["pixels"][141:] == 0xABCDEF01).all()


def test_host_register_makes_a_caller_buffer_zero_copy(lib, ctx, obj_scene):
    """softray_host_register page-locks a caller-owned surface (the pinned int[] of the C# host, or the shared
    section of the multi-GPU host path): same frame as the staged copy, banded frames write only their bands, and
    the buffer can be unregistered and used again."""
    p = scenario(resolution=128, shadows=True, shadow_samples=4)
    want = obj_scene.render(p)["pixels"]
    buf = np.full((128, 128), 0x01020304, dtype=np.uint32)
    ctx.host_register(buf)
    try:
        got = obj_scene.render(p, pixels=buf)["pixels"]
        assert np.array_equal(got, want)
        buf[:] = 0x01020304
        multi_gpu.apply_partition(p, 1, 2, 8)              # rank 1 of 2, bands of 8 rows
        obj_scene.render(p, pixels=buf)
        rows = multi_gpu.rows_of_rank(128, 2, 8, 1)
        mask = np.zeros(128, dtype=bool)
        mask[rows] = True
        assert np.array_equal(buf[mask], want[mask]) and (buf[~mask] == 0x01020304).all()
    finally:
        ctx.host_unregister(buf)
    multi_gpu.apply_partition(p, 0, 1, 0)
    assert np.array_equal(obj_scene.render(p, pixels=buf)["pixels"], want)      # staged copy again
    with pytest.raises(lib.SoftRayError):
        ctx.host_unregister(buf)                            # not registered any more: CUDA error code, no crash


def test_config1_full_size_against_oracle(lib, ctx, fixtures, obj_oracle):
    """configs[0] as named: obj.3DS through the NATIVE loader, 512x512, 1 spp, primary rays + Lambert
    (specularLighting = false) -- every pixel and hit id against the oracle."""
    mesh = lib.load_3ds(fixtures["model/obj.3ds"].tobytes())
    p = scenario(resolution=512, specular_lighting=False)
    got = lib.Scene(ctx, [mesh]).render(p, want_ids=True)
    want = obj_oracle.render(p, want_ids=True, want_aux=True)
    assert_parity(got, want, what="config1 512x512")
    assert got["stats"].rays_primary == 512 * 512 and got["stats"].rays_shadow == 0


@pytest.mark.parametrize("seed", [11, 12, 13, 14])
def test_random_soups_and_spheres_against_oracle(lib, ctx, seed):
    """Fuzz: random intersecting triangle soups + random spheres, random camera / light / flags, every
    pixel and hit id against the oracle (whose mesh path is the reference's kd-style tree)."""
    import oracle

    rng = synth.SplitMix64(seed)
    n = [60, 400, 2500, 9000][seed % 4]
    size = [0.4, 0.15, 0.06, 0.03][seed % 4]
    c = rng.uniform(3 * n).reshape(n, 3) - 0.5
    e = (rng.uniform(6 * n).reshape(n, 2, 3) - 0.5) * size
    v = np.concatenate([c, c + e[:, 0], c + e[:, 1]]).clip(-0.5, 0.5)
    t = np.stack([np.arange(n), np.arange(n) + n, np.arange(n) + 2 * n], axis=1).astype(np.int32)
    mesh = MeshData(v, t, synth.PALETTE[np.arange(n) % 8], v.min(axis=0), v.max(axis=0))
    ns = 12
    u = rng.uniform(4 * ns).reshape(ns, 4)
    sph = SphereData(np.concatenate([(u[:, :3] - 0.5) * 0.8, 0.02 + 0.08 * u[:, 3:]], axis=1), synth.PALETTE[np.arange(ns) % 8])
    u = rng.uniform(8)
    p = scenario(resolution=72, shadows=True, shadow_samples=6, sub_pixel_res=1 + int(u[0] * 2.99), focal_blur=u[1] < 0.3,
                 yaw_deg=360.0 * u[2], pitch_deg=80.0 * u[3] - 40.0, object_depth=0.9 + 1.2 * u[4], point_lighting=u[5] < 0.8,
                 specular_lighting=u[6] < 0.5, subdivision=u[7] < 0.8)
    got = lib.Scene(ctx, [mesh], sph).render(p, want_ids=True)
    want = oracle.Scene([mesh], sph).render(p, want_ids=True, want_aux=True)
    assert_parity(got, want, exact=True, what=f"soup seed {seed}")
    assert got["stats"].hits_primary == want["stats"].hits_primary > 100


def test_config4_full_size_rows_against_oracle(lib, ctx):
    """configs[3] as named (100 k-triangle mesh x 100 instances, 3840x2160, 16 spp): two full-width rows
    against the oracle, which traces every instance for every ray."""
    import oracle

    meshes, _, p = synth.config4()
    sc = lib.Scene(ctx, meshes)
    orc = oracle.Scene(meshes)
    for top in (700, 1403):
        p.start_row, p.end_row = top, top
        got = sc.render(p, want_ids=True)
        want = orc.render(p, want_ids=True, want_aux=True)
        sl = slice(top, top + 1)
        sub = dict(pixels=got["pixels"][sl], ids=got["ids"][sl])
        wsub = dict(pixels=want["pixels"][sl], ids=want["ids"][sl], cos_theta=want["cos_theta"][sl])
        assert_parity(sub, wsub, what=f"config4 row {top}")
        assert got["stats"].rays_primary == 3840 * 16 and got["stats"].hits_primary > 3840 * 8


def test_config5_full_size_filter_modes_and_bands(lib, ctx):
    """configs[4] as named (10 M triangles flattened, 7680x4320, 1 shadow ray per hit): a 96-row slice of
    the 8K frame rendered with the FP32 filters, with the FP64 reference arithmetic only, and as two
    interleaved bands (the multi-GPU partition) gives the same pixels and ids."""
    meshes, _, p = synth.config5()
    sc = lib.Scene(ctx, meshes)
    p.start_row, p.end_row = 2100, 2195
    a = sc.render(p, want_ids=True)
    p.filter_mode = abi.FILTER_OFF
    b = sc.render(p, want_ids=True)
    assert np.array_equal(a["pixels"], b["pixels"]) and np.array_equal(a["ids"], b["ids"])
    assert a["stats"].hits_primary == b["stats"].hits_primary > 100_000 and a["stats"].rays_shadow == a["stats"].hits_primary
    p.filter_mode = abi.FILTER_AUTO
    px = np.zeros_like(a["pixels"])
    for r in range(2):
        p.band_height, p.band_count, p.band_index = 8, 2, r
        sc.render(p, pixels=px)
    assert np.array_equal(px[2100:2196], a["pixels"][2100:2196])


def test_handles_may_be_released_in_any_order(lib, obj_mesh):
    """Finalisers run in any order (Renderer.Dispose vs the finaliser thread, CPython at exit): a context
    releases the scenes it still owns, and destroying a dead handle is a no-op (include/softray_cuda.h)."""
    import ctypes as C

    L = lib.load()
    c = lib.Context(0)
    s1 = lib.Scene(c, [obj_mesh])
    s2 = lib.Scene(c, [obj_mesh], accel=abi.ACCEL_LBVH)
    h_ctx, h1, h2 = C.c_void_p(c._h.value), C.c_void_p(s1._h.value), C.c_void_p(s2._h.value)
    s1.render(scenario(resolution=16))
    L.softray_scene_destroy(h1)
    L.softray_scene_destroy(h1)          # twice
    L.softray_destroy(h_ctx)             # takes s2 with it
    L.softray_scene_destroy(h2)          # after its context
    L.softray_destroy(h_ctx)             # twice
    for o in (s1, s2, c):
        o._h = C.c_void_p()
    # and the library still works
    c2 = lib.Context(0)
    out = lib.Scene(c2, [obj_mesh]).render(scenario(resolution=16))
    assert (out["pixels"] >> 24 == 0xFF).all()
    c2.close()


def test_phase_sync_changes_nothing(lib, ctx, obj_mesh, monkeypatch):
    """Stage barriers (sr_render.cu "Phase synchronisation") only change WHEN warps run what: frames, hit ids and
    the ray counters must be identical with them forced on and off -- including frames whose tiles do not fill the
    last block (warps without a tile), widths / heights that are not multiples of the 8x4 tile (lanes without a
    pixel), supersampling, focal blur, shadows, mirror bounces and spheres."""
    meshes2, spheres2, p2 = synth.config2(width=203, height=117, shadow_samples=16, n_spheres=300)
    meshes3, _, p3 = synth.config3(width=150, height=85, nx=101, nz=51, shadow_samples=3)
    jobs = [([obj_mesh], None, scenario(resolution=100, shadows=True)),
            ([obj_mesh], None, scenario(width=77, height=53, sub_pixel_res=2)),
            ([obj_mesh], None, scenario(width=90, height=61, sub_pixel_res=2, focal_blur=True, start_row=7, end_row=49)),
            ([obj_mesh], path_trace_spheres(), scenario(resolution=64, shadows=True, shadow_samples=5)),
            (meshes2, spheres2, p2), (meshes3, None, p3)]
    for meshes, sph, p in jobs:
        sc = lib.Scene(ctx, meshes, sph)
        out = {}
        for mode in ("0", "1"):
            monkeypatch.setenv("SOFTRAY_PHASE_SYNC", {"0": "0", "1": "31"}[mode])
            out[mode] = sc.render(p, want_ids=True)
        a, b = out["0"], out["1"]
        assert np.array_equal(a["pixels"], b["pixels"]) and np.array_equal(a["ids"], b["ids"])
        # (which shading points try a cone walk depends on their block's running score, i.e. on timing: the search
        # counters may differ between any two runs, the rays and what they hit may not)
        for k in ("rays_primary", "rays_shadow", "rays_secondary", "hits_primary", "shaded_hits"):
            assert getattr(a["stats"], k) == getattr(b["stats"], k), k
        sc.close()
