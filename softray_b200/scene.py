"""Plain host-side containers for what crosses the C ABI: meshes, spheres, instances, frames.

These hold numpy arrays / Python floats only and know how to lay themselves out as the POD
structs of include/softray_cuda.h.  No computation happens here.
"""
import ctypes as C
import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import abi


def _as(a, dtype, shape_tail):
    a = np.ascontiguousarray(a, dtype=dtype)
    if a.size == 0:
        a = a.reshape((0,) + shape_tail)
    assert a.shape[1:] == shape_tail, (a.shape, shape_tail)
    return a


@dataclass
class MeshData:
    """One Model after PostProcessGeometry (Model.cs:750-831): unit-cube vertices, triangles in
    Model.Triangles order, packed diffuse colours, Model.Min/Max."""

    verts: np.ndarray  # (n_verts, 3) float64
    tris: np.ndarray  # (n_tris, 3) int32
    argb: np.ndarray  # (n_tris,) uint32, alpha 0xFF
    bbox_min: np.ndarray  # (3,) float64
    bbox_max: np.ndarray  # (3,) float64

    def __post_init__(self):
        self.verts = _as(self.verts, np.float64, (3,))
        self.tris = _as(self.tris, np.int32, (3,))
        self.argb = np.ascontiguousarray(self.argb, dtype=np.uint32).reshape(-1)
        self.bbox_min = np.ascontiguousarray(self.bbox_min, dtype=np.float64).reshape(3)
        self.bbox_max = np.ascontiguousarray(self.bbox_max, dtype=np.float64).reshape(3)
        assert self.argb.shape[0] == self.tris.shape[0]

    @property
    def n_tris(self):
        return int(self.tris.shape[0])

    @property
    def n_verts(self):
        return int(self.verts.shape[0])

    def to_c(self) -> abi.Mesh:
        m = abi.Mesh()
        m.verts_xyz = self.verts.ctypes.data_as(abi.c_double_p)
        m.tri_vidx = self.tris.ctypes.data_as(abi.c_int32_p)
        m.tri_argb = self.argb.ctypes.data_as(abi.c_uint32_p)
        m.n_verts = self.n_verts
        m.n_tris = self.n_tris
        m.bbox_min[:] = self.bbox_min.tolist()
        m.bbox_max[:] = self.bbox_max.tolist()
        return m


@dataclass
class SphereData:
    """ExtraGeometryToRaytrace spheres: (n,4) cx,cy,cz,r and (n,) ARGB."""

    cxyzr: np.ndarray
    argb: np.ndarray

    def __post_init__(self):
        self.cxyzr = _as(self.cxyzr, np.float64, (4,))
        self.argb = np.ascontiguousarray(self.argb, dtype=np.uint32).reshape(-1)
        assert self.argb.shape[0] == self.cxyzr.shape[0]

    def __len__(self):
        return int(self.cxyzr.shape[0])

    def to_c(self):
        n = len(self)
        arr = (abi.Sphere * max(n, 1))()
        for i in range(n):
            arr[i].cx, arr[i].cy, arr[i].cz, arr[i].r = (float(v) for v in self.cxyzr[i])
            arr[i].argb = int(self.argb[i])
        return arr


class SceneDescHolder:
    """Keeps the numpy arrays and ctypes arrays referenced by a softray_scene_desc alive."""

    def __init__(self, meshes: Sequence[MeshData], spheres: Optional[SphereData] = None, accel: int = abi.ACCEL_BVH):
        self.meshes = list(meshes)
        self.spheres = spheres
        self._c_meshes = (abi.Mesh * max(len(self.meshes), 1))()
        for i, m in enumerate(self.meshes):
            self._c_meshes[i] = m.to_c()
        self._c_spheres = spheres.to_c() if spheres is not None and len(spheres) else None
        self.desc = abi.SceneDesc()
        self.desc.meshes = C.cast(self._c_meshes, C.POINTER(abi.Mesh))
        self.desc.n_meshes = len(self.meshes)
        self.desc.accel = accel
        self.desc.spheres = C.cast(self._c_spheres, C.POINTER(abi.Sphere)) if self._c_spheres is not None else None
        self.desc.n_spheres = len(spheres) if spheres is not None else 0


@dataclass
class InstanceData:
    """Instance (Instance.cs): position in view space, Euler angles in radians."""

    position: Sequence[float] = (0.0, 0.0, 1.5)  # Instance.cs:31
    yaw: float = 0.0
    pitch: float = 0.0
    roll: float = 0.0
    mesh_id: int = 0


@dataclass
class FrameParams:
    """The Renderer fields RaytraceGeometry reads (Renderer.cs:35-85,134-136,207-230)."""

    width: int = 1
    height: int = 1
    instances: List[InstanceData] = field(default_factory=list)
    ambient: float = 0.1
    shininess: float = 100.0
    light_dir_view: Optional[Sequence[float]] = None  # default normalise(-1,-1,1)
    light_pos_view: Optional[Sequence[float]] = None  # default (0,0,1.5) - 2*dir
    fov_depth: float = 0.5 / math.tan((45.0 / 180.0 * math.pi) / 2)
    focal_depth: float = 1.5
    focal_strength: float = 10.0
    start_row: Optional[int] = None
    end_row: Optional[int] = None
    sub_pixel_res: int = 1
    focal_blur: bool = True
    subdivision: bool = True
    shading: bool = True
    shadows: bool = False
    shadow_samples: int = 100
    point_lighting: bool = True
    specular_lighting: bool = True
    random_seed: int = 1234567890
    background: int = 0
    reflection_depth: int = 0
    texture3d_id: int = 0
    band_height: int = 0  # row-band partition (one band set per GPU); 0 / band_count<=1 = all rows
    band_count: int = 1
    band_index: int = 0
    filter_mode: int = 0  # abi.FILTER_AUTO / FILTER_OFF / FILTER_VERIFY (results identical in every mode)
    profile_stages: bool = False  # time every stage kernel (softray_stats.ms_stage); a measuring aid

    def default_light(self):
        inv = 1.0 / math.sqrt((-1.0) * (-1.0) + (-1.0) * (-1.0) + 1.0 * 1.0)  # Vector.Normalise
        d = (-1.0 * inv, -1.0 * inv, 1.0 * inv)
        p = (0.0 - d[0] * 2, 0.0 - d[1] * 2, 1.5 - d[2] * 2)
        return d, p

    def to_c(self, instance_init):
        """instance_init(inst_struct, pos[3], yaw, pitch, roll, mesh_id) fills the matrices (the
        product's softray_instance_init, or the oracle's twin)."""
        n = len(self.instances)
        inst = (abi.Instance * max(n, 1))()
        for i, it in enumerate(self.instances):
            pos = (C.c_double * 3)(*[float(v) for v in it.position])
            instance_init(C.byref(inst[i]), pos, float(it.yaw), float(it.pitch), float(it.roll), int(it.mesh_id))
        f = abi.Frame()
        d, p = self.default_light()
        f.ambient = self.ambient
        f.shininess = self.shininess
        f.light_dir_view[:] = list(self.light_dir_view if self.light_dir_view is not None else d)
        f.light_pos_view[:] = list(self.light_pos_view if self.light_pos_view is not None else p)
        f.fov_depth = self.fov_depth
        f.focal_depth = self.focal_depth
        f.focal_strength = self.focal_strength
        f.instances = C.cast(inst, C.POINTER(abi.Instance))
        f.n_instances = n
        f.width = self.width
        f.height = self.height
        f.start_row = 0 if self.start_row is None else self.start_row
        f.end_row = self.height - 1 if self.end_row is None else self.end_row
        f.sub_pixel_res = self.sub_pixel_res
        f.focal_blur = int(bool(self.focal_blur))
        f.subdivision = int(bool(self.subdivision))
        f.shading = int(bool(self.shading))
        f.shadows = int(bool(self.shadows))
        f.shadow_samples = self.shadow_samples
        f.point_lighting = int(bool(self.point_lighting))
        f.specular_lighting = int(bool(self.specular_lighting))
        f.random_seed = self.random_seed
        f.background_argb = self.background & 0x00FFFFFF  # BackgroundColor setter (Renderer.cs:318)
        f.reflection_depth = self.reflection_depth
        f.texture3d_id = self.texture3d_id
        f.band_height = self.band_height
        f.band_count = self.band_count
        f.band_index = self.band_index
        f.filter_mode = self.filter_mode
        f.profile_stages = int(bool(self.profile_stages))
        f._keepalive = inst
        return f
