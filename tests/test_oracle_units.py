"""The reference's unit facts and known-answer tests, run against the CPU oracle:
TriangleTests.cs:35-75, SphereTests.cs:20-30, SpatialSubdivisionTests.cs:24-137,284-392."""
import numpy as np
import pytest

import oracle
from softray_b200 import abi

ORIGIN, RIGHT, UP = (0, 0, 0), (1, 0, 0), (0, 1, 0)
FORWARD, BACKWARD = (0, 0, -1), (0, 0, 1)

TRIANGLE_SPACE = 100   # SpatialSubdivisionTests.cs:394
TRIANGLE_EXTENT = 10   # SpatialSubdivisionTests.cs:395


def make_random_triangles(rng, n):
    """MakeRandomTriangles (SpatialSubdivisionTests.cs:397-411): NextDouble order x,y,z per vector,
    then random.Next() for the colour."""
    tris = np.zeros((n, 3, 3))
    colors = np.zeros(n, dtype=np.uint32)

    def vec(size):
        return np.array([rng.next_double() * size, rng.next_double() * size, rng.next_double() * size])

    for i in range(n):
        v1 = vec(TRIANGLE_SPACE)
        tris[i, 0] = v1
        tris[i, 1] = v1 + vec(TRIANGLE_EXTENT)
        tris[i, 2] = v1 + vec(TRIANGLE_EXTENT)
        colors[i] = rng.next() & 0xFFFFFFFF
    return tris, colors


def build_random_tree(n, depth, per_node, seed):
    rng = oracle.SystemRandom(seed)
    tris, colors = make_random_triangles(rng, n)
    size = TRIANGLE_SPACE + TRIANGLE_EXTENT
    return oracle.Tree(tris, (0, 0, 0), (size, size, size), depth, per_node, colors), rng


# ------------------------------------------------------------------ System.Random
def test_system_random_first_values():
    """Well-known .NET Framework outputs: new Random(0).Next() and new Random(1).Next()."""
    assert oracle.SystemRandom(0).next() == 1559595546
    assert oracle.SystemRandom(1).next() == 534011718
    r = oracle.SystemRandom(12345)
    vals = [r.next_double() for _ in range(1000)]
    assert all(0.0 <= v < 1.0 for v in vals)


def test_area_light_offsets_are_radius_point2():
    off = oracle.area_light_offsets(1234567890, 100)
    assert np.allclose(np.linalg.norm(off, axis=1), 0.2, atol=1e-15)


# ------------------------------------------------------------------ TriangleTests.cs
def test_ray_hits_triangle():
    h = oracle.triangle_intersect(ORIGIN, RIGHT, UP, BACKWARD, FORWARD)
    assert h is not None
    assert h["pos"] == (0.0, 0.0, 0.0)
    assert h["normal"] == (0.0, 0.0, 1.0)
    assert h["ray_frac"] == 1.0
    assert h["color"] == 0xFFFFFFFF


def test_ray_from_triangle_vertex_hits_triangle():
    h = oracle.triangle_intersect(ORIGIN, RIGHT, UP, RIGHT, FORWARD)
    assert h is not None
    assert h["pos"] == (1.0, 0.0, 0.0)
    assert h["normal"] == (0.0, 0.0, 1.0)
    assert h["ray_frac"] == 0.0


def test_triangle_is_one_sided():
    assert oracle.triangle_intersect(ORIGIN, RIGHT, UP, FORWARD, BACKWARD) is None


def test_zero_size_triangle_never_hit():
    """CreateZeroSizeTriangle + SURVEY App. A #2: normal forced to (1,0,0), divisions give NaN."""
    assert oracle.triangle_intersect(ORIGIN, ORIGIN, ORIGIN, (1, 0, 0), (-1, 0, 0)) is None


def test_triangle_edges_inclusive():
    # ray through the midpoint of the hypotenuse (s+u == 1) and through an edge (u == 0)
    assert oracle.triangle_intersect(ORIGIN, RIGHT, UP, (0.5, 0.5, 1), FORWARD) is not None
    assert oracle.triangle_intersect(ORIGIN, RIGHT, UP, (0.5, 0.0, 1), FORWARD) is not None
    assert oracle.triangle_intersect(ORIGIN, RIGHT, UP, (0.5, -1e-9, 1), FORWARD) is None


# ------------------------------------------------------------------ SphereTests.cs
def test_sphere_contains_point():
    z = (0, 0, 0)
    assert oracle.sphere_contains_point(z, 1.0, z)
    assert oracle.sphere_contains_point(z, 1.0, (0.999, 0, 0))
    assert not oracle.sphere_contains_point(z, 1.0, (1, 0, 0))
    assert not oracle.sphere_contains_point(z, 1.0, (1.0001, 0, 0))
    assert not oracle.sphere_contains_point(z, 1.0, (2, 0, 0))


def test_sphere_ray_frac_is_a_distance():
    """the sphere normalises dir, so rayFrac is Euclidean distance."""
    h = oracle.sphere_intersect((0, 0, 0), 1.0, (0, 0, 5), (0, 0, -10))
    assert h is not None and h["ray_frac"] == 4.0 and h["normal"] == (0.0, 0.0, 1.0)
    # from inside: the far root
    h = oracle.sphere_intersect((0, 0, 0), 1.0, (0, 0, 0), (0, 0, -3))
    assert h is not None and h["ray_frac"] == 1.0
    # grazing line: term < 1e-10 => miss (Sphere.cs:179)
    assert oracle.sphere_intersect((0, 0, 0), 1.0, (1, 0, 5), (0, 0, -1)) is None
    # sphere entirely behind the start
    assert oracle.sphere_intersect((0, 0, 0), 1.0, (0, 0, -5), (0, 0, -1)) is None


# ------------------------------------------------------------------ AxisAlignedBox
def test_box_contains_epsilon():
    assert oracle.box_contains_point((0, 0, 0), (1, 1, 1), (1 + 5e-11, 0.5, 0.5))
    assert not oracle.box_contains_point((0, 0, 0), (1, 1, 1), (1 + 2e-10, 0.5, 0.5))


def test_box_clip_line_segment():
    got = oracle.box_clip_line_segment((0, 0, 0), (1, 1, 1), (-1, 0.5, 0.5), (2, 0.5, 0.5))
    assert got == ((0.0, 0.5, 0.5), (1.0, 0.5, 0.5))
    assert oracle.box_clip_line_segment((0, 0, 0), (1, 1, 1), (-1, 2, 0.5), (2, 2, 0.5)) is None
    inside = oracle.box_clip_line_segment((0, 0, 0), (1, 1, 1), (0.25, 0.5, 0.5), (0.75, 0.5, 0.5))
    assert inside == ((0.25, 0.5, 0.5), (0.75, 0.5, 0.5))


# ------------------------------------------------------------------ SpatialSubdivisionTests.cs KATs
@pytest.mark.parametrize("n,depth,per_node,expect", [
    (10, 5, 3, (4, 9, 5, 4)),          # ConstructArbitraryTree
    (5, 3, 1, (2, 3, 2, 1)),           # ConstructMaxDepthTree
    (8, 3, 1, (3, 5, 3, 2)),           # ConstructBalancedTree
    (4, 100, 1, (2, 3, 2, 1)),         # ConstructUnbalancedTree
    (1000, 10, 5, (10, 885, 443, 442)),  # ConstructBigTree
])
def test_tree_build_known_answers(n, depth, per_node, expect):
    tree, _ = build_random_tree(n, depth, per_node, 12345)
    s = tree.stats()
    assert (s["depth"], s["nodes"], s["leaves"], s["internal"]) == expect


def test_tree_ctor_error_contracts():
    """A vertex outside the bounding box -> ArgumentOutOfRangeException (SpatialSubdivision.cs:285-295);
    empty input and degenerate triangles are fine (SpatialSubdivisionTests.cs:31-57)."""
    tri = np.array([[[0.5, 0.5, 0.5], [2.0, 0.5, 0.5], [0.5, 0.9, 0.5]]])
    with pytest.raises(oracle.OracleError) as e:
        oracle.Tree(tri, (0, 0, 0), (1, 1, 1))
    assert e.value.code == abi.E_VERTEX_OUTSIDE_BBOX
    assert oracle.Tree(np.zeros((0, 3, 3)), (0, 0, 0), (0, 0, 0)).stats()["nodes"] == 1
    assert oracle.Tree(np.zeros((1, 3, 3)), (0, 0, 0), (1, 1, 1)).stats()["leaves"] == 1


def _check_tree_vs_brute(tree, rng, n_rays, outside_in):
    for i in range(n_rays):
        if outside_in:
            start = [rng.next_double() * TRIANGLE_SPACE * 10 for _ in range(3)]
            end = [rng.next_double() * TRIANGLE_SPACE for _ in range(3)]
            d = [e - s for s, e in zip(start, end)]
        else:
            start = [rng.next_double() * TRIANGLE_SPACE for _ in range(3)]
            d = [2 * rng.next_double() - 1 for _ in range(3)]
        a = tree.intersect(start, d)
        b = tree.brute_intersect(start, d)
        assert (a is None) == (b is None), f"ray {i}"
        if a is not None:
            assert a["tri_index"] == b["tri_index"], f"ray {i}"
            if outside_in:
                assert abs(a["ray_frac"] - b["ray_frac"]) <= 1e-10
            else:
                assert a["ray_frac"] == b["ray_frac"]
            # Vector == is (a-b).LengthSqr < 1e-10 (Vector.cs:43-47)
            assert sum((p - q) ** 2 for p, q in zip(a["pos"], b["pos"])) < 1e-10
            assert a["normal"] == b["normal"] and a["color"] == b["color"]


@pytest.mark.parametrize("n,depth,per_node,seed,rays,outside_in", [
    (100, 10, 5, 12345, 4000, False),     # TreeCorrectness1
    (20, 10, 1, 123456, 4000, False),     # TreeCorrectness2
    (20, 10, 1, 123456, 4000, True),      # TreeCorrectness_OutsideIn
    (100, 10, 5, 1234567, 4000, False),   # ..._EnsureIntersectionCheckedAgainstTreeNodeBoundingBox
    (10000, 10, 5, 1234567, 34, False),   # ..._BoundingBox2
])
def test_tree_equals_brute_force(n, depth, per_node, seed, rays, outside_in):
    """TestTree_InsideOut / TestTree_OutsideIn (SpatialSubdivisionTests.cs:341-392) with the RNG
    stream continuing after the triangles, as the reference does (a bounded number of rays)."""
    tree, rng = build_random_tree(n, depth, per_node, seed)
    _check_tree_vs_brute(tree, rng, rays, outside_in)


# ------------------------------------------------------------------ Model / loader
def test_obj3ds_model_facts(obj_mesh, obj2_mesh):
    """obj.3ds = 269 verts / 152 tris, one material diffuse 0.5882353 ->
    0xFF969696; normalised so the longest axis is exactly [-0.5, 0.5]; obj2 = 112 / 107."""
    assert (obj_mesh.n_verts, obj_mesh.n_tris) == (269, 152)
    assert set(obj_mesh.argb.tolist()) == {0xFF969696}
    assert obj_mesh.bbox_min[0] == -0.5 and obj_mesh.bbox_max[0] == 0.5
    assert (obj2_mesh.n_verts, obj2_mesh.n_tris) == (112, 107)
    assert len(set(obj2_mesh.argb.tolist())) == 2
    assert np.all(obj_mesh.verts.min(axis=0) == obj_mesh.bbox_min) and np.all(obj_mesh.verts.max(axis=0) == obj_mesh.bbox_max)


def test_loader_rejects_garbage():
    with pytest.raises(oracle.OracleError) as e:
        oracle.load_3ds(b"not a 3ds file at all")
    assert e.value.code == abi.E_FORMAT
    with pytest.raises(oracle.OracleError):
        oracle.load_3ds(b"MM\x10\x00\x00\x00" + b"\x00" * 10)   # primary chunk with no entities


def test_model_from_arrays_normalises():
    v = np.array([[0, 0, 0], [4, 0, 0], [0, 2, 0], [0, 0, 1]], dtype=float)
    t = np.array([[0, 1, 2], [0, 2, 3]], dtype=np.int32)
    m = oracle.model_from_arrays(v, t)
    assert m.bbox_min.tolist() == [-0.5, -0.25, -0.125] and m.bbox_max.tolist() == [0.5, 0.25, 0.125]
    raw = oracle.model_from_arrays(v, t, normalise=False)
    assert raw.bbox_max.tolist() == [4, 2, 1]


def test_oracle_column_window_is_the_same_pixels(obj_mesh):
    """The test-only column window (orc_options.col_start/col_end) traces exactly the pixels of the window: same
    colours, ids and ray counts as the same pixels of the whole rows; nothing outside the window is written."""
    import numpy as np

    import oracle
    from tests.util import scenario

    orc = oracle.Scene([obj_mesh])
    p = scenario(resolution=64, shadows=True, shadow_samples=5, sub_pixel_res=2, start_row=20, end_row=41)
    full = orc.render(p, want_ids=True)
    px = np.full((64, 64), 0xDEADBEEF, dtype=np.uint32)
    win = orc.render(p, options=oracle.default_options(col_start=17, col_end=40), want_ids=True, pixels=px)
    assert (win["pixels"][20:42, 17:41] == full["pixels"][20:42, 17:41]).all()
    assert (win["ids"][20:42, 17:41] == full["ids"][20:42, 17:41]).all()
    outside = np.ones((64, 64), dtype=bool)
    outside[20:42, 17:41] = False
    assert (px[outside] == 0xDEADBEEF).all()
    assert win["stats"].rays_primary == 22 * 24 * 4 and full["stats"].rays_primary == 22 * 64 * 4
    # an empty window produces nothing
    none = orc.render(p, options=oracle.default_options(col_start=50, col_end=40))
    assert none["stats"].rays == 0


def test_multiply_high_division_constants():
    """The stage kernels divide ray indices by per-frame constants with one multiply-high and a shift (FastDiv,
    sr_types.h; constants from make_fastdiv, sr_api.cu: p = 31 + ceil(log2 d), mul = ceil(2^p / d), shift = p - 32).
    The same construction, checked exhaustively near every boundary for the divisors frames use."""
    def make(d):
        if d == 1:
            return None
        lg = (d - 1).bit_length()
        p = 31 + lg
        return ((1 << p) + d - 1) // d, p - 32

    for d in [1, 2, 3, 4, 5, 7, 9, 16, 25, 32, 36, 49, 64, 81, 128, 288, 960, 1000, 2048, 32 * 64, 7680, 65535, 1 << 20, (1 << 20) + 7]:
        c = make(d)
        ns = {0, 1, d - 1, d, d + 1, 2 * d - 1, 2 * d, (1 << 31) - 1, (1 << 31) - d, 123456789, 987654321}
        ns |= {k * d + off for k in (3, 1000, (1 << 31) // d - 1) for off in (-1, 0, 1)}
        for n in ns:
            if not 0 <= n < (1 << 31):
                continue
            if c is not None:
                assert c[0] < (1 << 32)
            q = n if c is None else ((n * c[0]) >> 32) >> c[1]
            assert q == n // d, (n, d)
