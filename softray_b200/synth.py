"""Deterministic synthetic scenes of the shapes BASELINE.json's configs name.

Everything is generated the way the reference's procedural models are (the Cloth.cs pattern,
Cloth.cs:8-45: fill the vertex / triangle lists of a Model) and already sits inside the unit cube
that Model.PostProcessGeometry (Model.cs:750-831) would normalise it to; bbox = exact min/max of
the vertices (Model.CalcExtent).  Randomness comes from SplitMix64 (documented below), never from
System.Random, so the scenes are identical on every platform.
"""
import math

import numpy as np

from .scene import FrameParams, InstanceData, MeshData, SphereData

PALETTE = np.array([0xFFE6194B, 0xFF3CB44B, 0xFFFFE119, 0xFF4363D8, 0xFFF58231, 0xFF911EB4, 0xFF42D4F4, 0xFFF0F0F0],
                   dtype=np.uint32)


class SplitMix64:
    """Vigna's SplitMix64: z = (x += 0x9E3779B97F4A7C15); z = (z ^ z>>30) * 0xBF58476D1CE4E5B9;
    z = (z ^ z>>27) * 0x94D049BB133111EB; return z ^ z>>31.  uniform() = top 53 bits / 2^53."""

    def __init__(self, seed):
        self.x = np.uint64(seed)

    def next_u64(self, n):
        with np.errstate(over="ignore"):
            idx = np.arange(1, n + 1, dtype=np.uint64)
            z = self.x + idx * np.uint64(0x9E3779B97F4A7C15)
            self.x = z[-1] if n else self.x
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            return z ^ (z >> np.uint64(31))

    def uniform(self, n):
        return (self.next_u64(n) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def mesh_from_arrays(verts, tris, argb):
    verts = np.ascontiguousarray(verts, dtype=np.float64).reshape(-1, 3)
    return MeshData(verts, tris, argb, verts.min(axis=0), verts.max(axis=0))


def room_box(argb=0xFFB4B4B4):
    """The unit cube [-0.5,0.5]^3 as 12 triangles whose normals (edge1 x edge2, Triangle.cs:41) point
    INTO the cube: rays enter through the one-sided faces (Plane.cs:75) and hit the far walls, so
    the cube is a room around whatever ExtraGeometry is placed in it."""
    v = np.array([[x, y, z] for x in (-0.5, 0.5) for y in (-0.5, 0.5) for z in (-0.5, 0.5)], dtype=np.float64)
    # vertex index = 4*ix + 2*iy + iz
    quads = [
        (0, 1, 3, 2),  # x = -0.5, normal +x
        (4, 6, 7, 5),  # x = +0.5, normal -x
        (0, 4, 5, 1),  # y = -0.5, normal +y
        (2, 3, 7, 6),  # y = +0.5, normal -y
        (0, 2, 6, 4),  # z = -0.5, normal +z
        (1, 5, 7, 3),  # z = +0.5, normal -z
    ]
    tris = []
    for a, b, c, d in quads:
        tris += [(a, b, c), (a, c, d)]
    tris = np.array(tris, dtype=np.int32)
    # make every normal point at the origin (the cube centre), whatever the winding above
    for i, (a, b, c) in enumerate(tris):
        n = np.cross(v[b] - v[a], v[c] - v[a])
        if np.dot(n, -v[a]) < 0:
            tris[i] = (a, c, b)
    return mesh_from_arrays(v, tris, np.full(12, argb, dtype=np.uint32))


def sphere_cloud(n=1000, seed=1, extent=0.45, r_min=0.01, r_max=0.04):
    """n spheres: centres uniform in [-extent,extent]^3, radii uniform in [r_min,r_max], colours
    cycling through an 8-entry palette by a random draw."""
    rng = SplitMix64(seed)
    u = rng.uniform(5 * n).reshape(n, 5)
    c = (u[:, 0:3] * 2.0 - 1.0) * extent
    r = r_min + u[:, 3] * (r_max - r_min)
    col = PALETTE[(u[:, 4] * 8).astype(np.int64) % 8]
    return SphereData(np.concatenate([c, r[:, None]], axis=1), col)


def height_field(nx=1001, nz=501, amplitude=0.15, seed=2):
    """2*(nx-1)*(nz-1) triangles (default exactly 1 000 000): y = amplitude*sin(9x)*cos(7z) over
    x,z in [-0.5,0.5], normals up (+y), per-quad palette colours."""
    x = np.linspace(-0.5, 0.5, nx)
    z = np.linspace(-0.5, 0.5, nz)
    X, Z = np.meshgrid(x, z, indexing="ij")
    Y = amplitude * np.sin(9.0 * X) * np.cos(7.0 * Z)
    verts = np.stack([X, Y, Z], axis=-1).reshape(-1, 3)
    i, j = np.meshgrid(np.arange(nx - 1), np.arange(nz - 1), indexing="ij")
    a = (i * nz + j).ravel()
    b = a + 1          # (i, j+1)
    c = a + nz         # (i+1, j)
    d = c + 1          # (i+1, j+1)
    # (a, b, c): edge1 = +z, edge2 = +x  ->  z cross x = +y
    tris = np.empty((2 * a.size, 3), dtype=np.int32)
    tris[0::2] = np.stack([a, b, c], axis=1)
    tris[1::2] = np.stack([b, d, c], axis=1)
    rng = SplitMix64(seed)
    quad_col = PALETTE[(rng.uniform(a.size) * 8).astype(np.int64) % 8]
    argb = np.repeat(quad_col, 2)
    return mesh_from_arrays(verts, tris, argb)


def uv_sphere(n_lon=250, n_lat=200, radius=0.5, seed=3, bumps=0.04):
    """2*n_lon*n_lat triangles (default 100 000) on a bumpy sphere, outward normals.  The two
    polar rows contain zero-area triangles: the reference's Triangle ctor accepts them and they
    can never be hit (Triangle.cs:42-43, TriangleTests.cs:35-44)."""
    th = np.linspace(0.0, math.pi, n_lat + 1)
    ph = np.linspace(0.0, 2.0 * math.pi, n_lon + 1)[:-1]
    T, P = np.meshgrid(th, ph, indexing="ij")
    R = radius * (1.0 - bumps + bumps * np.sin(8.0 * T) * np.cos(6.0 * P))
    verts = np.stack([R * np.sin(T) * np.cos(P), R * np.cos(T), R * np.sin(T) * np.sin(P)], axis=-1).reshape(-1, 3)
    i, j = np.meshgrid(np.arange(n_lat), np.arange(n_lon), indexing="ij")
    a = (i * n_lon + j).ravel()
    b = (i * n_lon + (j + 1) % n_lon).ravel()
    c = a + n_lon
    d = b + n_lon
    tris = np.empty((2 * a.size, 3), dtype=np.int32)
    tris[0::2] = np.stack([a, b, c], axis=1)
    tris[1::2] = np.stack([b, d, c], axis=1)
    # orient outward
    v = verts
    n = np.cross(v[tris[:, 1]] - v[tris[:, 0]], v[tris[:, 2]] - v[tris[:, 0]])
    cen = (v[tris[:, 0]] + v[tris[:, 1]] + v[tris[:, 2]]) / 3.0
    flip = (n * cen).sum(axis=1) < 0
    tris[flip] = tris[flip][:, [0, 2, 1]]
    rng = SplitMix64(seed)
    argb = np.repeat(PALETTE[(rng.uniform(a.size) * 8).astype(np.int64) % 8], 2)
    verts = verts / (2.0 * np.abs(verts).max())      # longest axis spans [-0.5, 0.5] (Model.cs:762-790)
    return mesh_from_arrays(verts, tris, argb)


def instance_grid(n_side=10, depth=7.0, spread=5.0, seed=4):
    """n_side^2 instances of mesh 0 on a grid in view space, each with its own yaw/pitch."""
    rng = SplitMix64(seed)
    u = rng.uniform(2 * n_side * n_side).reshape(-1, 2)
    out = []
    for k in range(n_side * n_side):
        ix, iy = k % n_side, k // n_side
        x = ((ix + 0.5) / n_side - 0.5) * spread
        y = ((iy + 0.5) / n_side - 0.5) * spread * 9.0 / 16.0
        out.append(InstanceData(position=(x, y, depth), yaw=float(u[k, 0]) * 2.0 * math.pi,
                                pitch=(float(u[k, 1]) - 0.5) * math.pi * 0.5, roll=0.0, mesh_id=0))
    return out


def flattened_grid(base: MeshData, n_side=10):
    """n_side^2 copies of `base` laid out on a plane and flattened into ONE mesh inside the unit
    cube (config 5: 10 M triangles from a 100 k base)."""
    cells = n_side
    s = 1.0 / cells
    vs, ts, cs = [], [], []
    nv = base.n_verts
    for k in range(cells * cells):
        ix, iz = k % cells, k // cells
        off = np.array([(ix + 0.5) * s - 0.5, 0.0, (iz + 0.5) * s - 0.5])
        vs.append(base.verts * (s * 0.95) + off)
        ts.append(base.tris + k * nv)
        cs.append(base.argb)
    return mesh_from_arrays(np.concatenate(vs), np.concatenate(ts).astype(np.int32), np.concatenate(cs))


BACKGROUND = 0xFF00FF   # RendererTests.cs:68


def camera(depth=1.5, yaw_deg=135.0, pitch_deg=-22.0):
    """The reference tests' default camera (RendererTests.cs:47-49) at a chosen object depth."""
    return InstanceData(position=(0.0, 0.0, depth), yaw=yaw_deg / 180.0 * math.pi, pitch=pitch_deg / 180.0 * math.pi,
                        roll=0.0)


def config2(width=1920, height=1080, shadow_samples=100, n_spheres=1000):
    """configs[1]: procedural 1000-sphere scene, 1920x1080, Phong + shadow rays."""
    meshes = [room_box()]
    spheres = sphere_cloud(n_spheres, seed=1)
    frame = FrameParams(width=width, height=height, instances=[camera(1.5)], background=BACKGROUND, shading=True,
                        specular_lighting=True, shadows=True, shadow_samples=shadow_samples)
    return meshes, spheres, frame


def config3(width=3840, height=2160, nx=1001, nz=501, shadow_samples=100):
    """configs[2]: synthetic 1M-triangle mesh, shadows + 2-bounce reflection + Texture3D."""
    meshes = [height_field(nx, nz)]
    frame = FrameParams(width=width, height=height, instances=[camera(1.5)], background=BACKGROUND, shading=True,
                        shadows=True, shadow_samples=shadow_samples, reflection_depth=2, texture3d_id=1)
    return meshes, None, frame


def config4(width=3840, height=2160, n_lon=250, n_lat=200, n_side=10, sub_pixel_res=4):
    """configs[3]: 10M-triangle instanced scene (100 k base x 100 instances), 16 spp."""
    meshes = [uv_sphere(n_lon, n_lat)]
    frame = FrameParams(width=width, height=height, instances=instance_grid(n_side), background=BACKGROUND,
                        shading=True, shadows=False, sub_pixel_res=sub_pixel_res, focal_blur=False)
    return meshes, None, frame


def config5(width=7680, height=4320, n_lon=250, n_lat=200, n_side=10, shadow_samples=1):
    """configs[4]: 8K frame, 10M triangles flattened, row bands across GPUs."""
    meshes = [flattened_grid(uv_sphere(n_lon, n_lat), n_side)]
    frame = FrameParams(width=width, height=height, instances=[camera(1.5)], background=BACKGROUND, shading=True,
                        shadows=True, shadow_samples=shadow_samples)
    return meshes, None, frame
