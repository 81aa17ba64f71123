"""The stage-kernel pipeline (softray_b200/csrc/sr_wave.cu) against the fused kernel, the reference's golden images
and the oracle.  Both pipelines run the same per-ray arithmetic, so every comparison is held to ZERO differing
pixels, hit ids and ray counters.  SOFTRAY_PIPELINE=wave|fused picks the pipeline for frames both can render."""
import os

import numpy as np
import pytest

from softray_b200 import MeshData, synth
from tests.test_cuda_parity import ON_PATH, assert_parity
from tests.util import count_diff, golden_name, scenario

pytestmark = pytest.mark.gpu

EXACT_COUNTERS = ("rays_primary", "rays_shadow", "rays_secondary", "hits_primary", "shaded_hits")


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


class pipeline:
    def __init__(self, name, chunk=None):
        self.env = {"SOFTRAY_PIPELINE": name}
        if chunk is not None:
            self.env["SOFTRAY_WAVE_CHUNK"] = str(chunk)

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.env}
        os.environ.update(self.env)

    def __exit__(self, *exc):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def both(scene, params, chunk=None, **kw):
    with pipeline("fused"):
        a = scene.render(params, want_ids=True, **kw)
    with pipeline("wave", chunk):
        b = scene.render(params, want_ids=True, **kw)
    assert a["stats"].launches == 1 and b["stats"].launches > 1, "the two pipelines were not both exercised"
    return a, b


def assert_same(a, b, what=""):
    assert count_diff(a["pixels"], b["pixels"]) == 0, f"{what}: {count_diff(a['pixels'], b['pixels'])} pixels differ"
    assert (a["pixels"] == b["pixels"]).all(), what
    assert (a["ids"] == b["ids"]).all(), f"{what}: {(a['ids'] != b['ids']).sum()} hit ids differ"
    for k in EXACT_COUNTERS:
        assert getattr(a["stats"], k) == getattr(b["stats"], k), f"{what}: {k}"


@pytest.mark.parametrize("res,kw", ON_PATH, ids=lambda v: str(v))
def test_wave_reference_goldens(fixtures, obj_scene, res, kw):
    """Every on-path golden BMP of the reference through the stage kernels."""
    golden = fixtures[f"golden/{res}x{res}/{golden_name(**kw)}"]
    with pipeline("wave"):
        out = obj_scene.render(scenario(resolution=res, **kw))
    assert out["stats"].launches > 1
    assert count_diff(out["pixels"], golden) == 0
    assert (out["pixels"] >> 24 == 0xFF).all()


@pytest.mark.parametrize("kw", [
    dict(resolution=256, shadows=True, point_lighting=False),
    dict(resolution=200, shadows=True, shadow_samples=7),
    dict(width=333, height=77),
    dict(resolution=160, pitch_deg=35.0, yaw_deg=20.0, roll_deg=50.0, object_depth=0.3),
    dict(resolution=128, subdivision=False, shadows=True),
    dict(resolution=64, sub_pixel_res=3, focal_blur=True, shadows=True, focal_strength=25.0),
    dict(resolution=100, start_row=37, end_row=58, shadows=True, shadow_samples=5),
    dict(width=96, height=83, shadows=True, shadow_samples=5, band_height=5, band_count=3, band_index=1),
    dict(resolution=90, reflection_depth=3, shadows=True, shadow_samples=4, texture3d_id=1),
], ids=lambda v: str(v))
def test_wave_equals_fused_on_obj3ds(obj_scene, kw):
    p = scenario(**kw)
    px0 = np.full((p.height, p.width), 0xDEADBEEF, dtype=np.uint32)
    px1 = px0.copy()
    with pipeline("fused"):
        a = obj_scene.render(p, pixels=px0, want_ids=True)
    with pipeline("wave", chunk=2048):
        b = obj_scene.render(p, pixels=px1, want_ids=True)
    assert a["stats"].launches == 1 and b["stats"].launches > 1
    assert (px0 == px1).all(), f"{(px0 != px1).sum()} pixels differ"      # incl. the rows / bands nobody may touch
    for k in EXACT_COUNTERS:
        assert getattr(a["stats"], k) == getattr(b["stats"], k), k


def test_wave_config3_small(lib, ctx):
    """Shadows + 2 mirror bounces + Texture3D: wave == fused == oracle."""
    import oracle

    meshes, _, p = synth.config3(width=200, height=112, nx=121, nz=81, shadow_samples=12)
    sc = lib.Scene(ctx, meshes)
    a, b = both(sc, p, chunk=4096)
    assert_same(a, b, "config3")
    opt = oracle.default_options(tree_max_depth=15, tree_max_per_node=25)
    want = oracle.Scene(meshes, options=opt).render(p, want_ids=True, want_aux=True)
    assert_parity(b, want, what="config3 wave vs oracle")
    assert b["stats"].rays_secondary == want["stats"].rays_secondary > 0


def test_wave_config3_100_samples_bundles(lib, ctx):
    """100 soft-shadow rays per hit (cone tests on): wave == fused, and single-chunk == many chunks."""
    meshes, _, p = synth.config3(width=160, height=90, nx=201, nz=101, shadow_samples=100)
    sc = lib.Scene(ctx, meshes)
    a, b = both(sc, p)
    assert_same(a, b, "config3 x100")
    with pipeline("wave", chunk=1024):
        c = sc.render(p, want_ids=True)
    assert_same(b, c, "chunked")


def test_wave_config4_small(lib, ctx):
    """Composite frame (instance hierarchy) with 2x2 supersampling: wave == fused == oracle."""
    import oracle

    meshes, _, p = synth.config4(width=160, height=90, n_lon=40, n_lat=30, n_side=4, sub_pixel_res=2)
    sc = lib.Scene(ctx, meshes)
    a, b = both(sc, p, chunk=8192)
    assert_same(a, b, "config4")
    want = oracle.Scene(meshes).render(p, want_ids=True, want_aux=True)
    assert_parity(b, want, what="config4 wave vs oracle")


def test_wave_config4_overlapping_instances(lib, ctx):
    """Many instances behind one another (the candidate list of a sample spans instances)."""
    meshes, _, p = synth.config4(width=200, height=112, n_lon=60, n_lat=40, n_side=6, sub_pixel_res=1)
    for i, inst in enumerate(p.instances):          # pull the grid together so that instances overlap on screen
        x, y, z = inst.position
        inst.position = (x * 0.35, y * 0.35, z + 0.6 * (i % 5))
    sc = lib.Scene(ctx, meshes)
    a, b = both(sc, p)
    assert_same(a, b, "overlapping instances")


def test_wave_config5_small(lib, ctx):
    import oracle

    meshes, _, p = synth.config5(width=192, height=108, n_lon=40, n_lat=30, n_side=4, shadow_samples=3)
    sc = lib.Scene(ctx, meshes)
    a, b = both(sc, p, chunk=4096)
    assert_same(a, b, "config5")
    want = oracle.Scene(meshes).render(p, want_ids=True, want_aux=True)
    assert_parity(b, want, what="config5 wave vs oracle")


def test_wave_triangle_soup_and_degenerate_triangles(lib, ctx):
    """Random overlapping triangles, zero-area triangles and faces lying in the bounding box itself (the clipped
    start decides them: SpatialSubdivision.cs:389-401): the rays the search cannot bracket take the fallback kernels."""
    rng = np.random.default_rng(5)
    n = 4000
    c = rng.uniform(-0.45, 0.45, (n, 1, 3))
    v = (c + rng.normal(0, 0.03, (n, 3, 3))).clip(-0.5, 0.5).reshape(-1, 3)
    v[:30] = v[0]                                                # ten zero-area triangles
    tris = np.arange(3 * n, dtype=np.int32).reshape(n, 3)
    box = synth.room_box()
    verts = np.concatenate([v, box.verts])
    tris = np.concatenate([tris, box.tris + 3 * n])
    argb = np.concatenate([synth.PALETTE[rng.integers(0, 8, n)], box.argb])
    mesh = MeshData(verts, tris, argb, verts.min(axis=0), verts.max(axis=0))
    sc = lib.Scene(ctx, [mesh])
    for kw in (dict(resolution=150, shadows=True, shadow_samples=9), dict(resolution=120, object_depth=0.2, yaw_deg=10.0, pitch_deg=5.0),
               dict(resolution=96, sub_pixel_res=2, reflection_depth=2)):
        a, b = both(sc, scenario(**kw), chunk=4096)
        assert_same(a, b, str(kw))


def test_wave_is_the_default_for_big_scenes_and_composites(lib, ctx, obj_scene):
    os.environ.pop("SOFTRAY_PIPELINE", None)
    assert obj_scene.render(scenario(resolution=32))["stats"].launches == 1            # small scene: fused kernel
    meshes, _, p = synth.config3(width=64, height=36, nx=301, nz=151, shadow_samples=2)   # 90 000 triangles
    assert lib.Scene(ctx, meshes).render(p)["stats"].launches > 1
    meshes, _, p = synth.config4(width=64, height=36, n_lon=20, n_lat=10, n_side=2, sub_pixel_res=1)
    assert lib.Scene(ctx, meshes).render(p)["stats"].launches > 1


def test_wave_big_frame_leaves_chunk_by_chunk_by_dma(lib, ctx):
    """A big stage-kernel frame into a page-locked surface is copied out chunk by chunk while the next chunk traces:
    same pixels and ids as through a pageable surface, only the requested rows / bands touched."""
    import torch

    meshes, _, p = synth.config3(width=2048, height=1100, nx=301, nz=151, shadow_samples=2)
    sc = lib.Scene(ctx, meshes)
    want = sc.render(p, want_ids=True)
    assert want["stats"].launches > 1
    px = torch.full((1100, 2048), 0x01020304, dtype=torch.int32).pin_memory()
    ids = torch.full((1100, 2048), 777, dtype=torch.int32).pin_memory()
    got = sc.render(p, want_ids=True, pixels=px.numpy().view(np.uint32), ids=ids.numpy())
    assert np.array_equal(got["pixels"], want["pixels"]) and np.array_equal(got["ids"], want["ids"])
    for kw in (dict(start_row=101, end_row=1003), dict(band_height=36, band_count=3, band_index=2), dict(start_row=7, end_row=1090, band_height=20, band_count=2, band_index=0)):
        for k, v in kw.items():
            setattr(p, k, v)
        px.fill_(0x01020304)
        ids.fill_(777)
        sc.render(p, pixels=px.numpy().view(np.uint32), ids=ids.numpy())
        s0, e0 = (p.start_row or 0), (p.end_row if p.end_row is not None else 1099)
        rows = np.arange(1100)
        mine = (rows >= s0) & (rows <= e0)
        if p.band_count > 1:
            mine &= ((rows - s0) // p.band_height) % p.band_count == p.band_index
        gp, gi = px.numpy().view(np.uint32), ids.numpy()
        assert np.array_equal(gp[mine], want["pixels"][mine]) and np.array_equal(gi[mine], want["ids"][mine]), kw
        assert (gp[~mine] == 0x01020304).all() and (gi[~mine] == 777).all(), kw
        p.start_row, p.end_row, p.band_height, p.band_count, p.band_index = None, None, 0, 1, 0
