"""ctypes mirror of include/softray_cuda.h (the C ABI of libsoftray_cuda.so).

Field order and types must match the header exactly; tests/test_abi.py checks the struct sizes
against the values the C side reports through softray_abi_sizeof().
"""
import ctypes as C

ABI_VERSION = 4

OK = 0
E_INVALID_ARG = -1
E_VERTEX_OUTSIDE_BBOX = -2
E_NO_DEVICE = -3
E_CUDA = -4
E_OOM = -5
E_UNSUPPORTED = -6
E_FORMAT = -7
E_TIMEOUT = -8

ACCEL_BVH = 0
ACCEL_BRUTE = 1
ACCEL_LBVH = 2

FILTER_AUTO = 0
FILTER_OFF = 1
FILTER_VERIFY = 2

STAGE_NAMES = ["search", "hit", "fallback", "search_ref", "hit_ref", "fallback_ref", "shadow", "shadow_fb", "compose"]

ERROR_NAMES = {
    OK: "SOFTRAY_OK",
    E_INVALID_ARG: "SOFTRAY_E_INVALID_ARG",
    E_VERTEX_OUTSIDE_BBOX: "SOFTRAY_E_VERTEX_OUTSIDE_BBOX",
    E_NO_DEVICE: "SOFTRAY_E_NO_DEVICE",
    E_CUDA: "SOFTRAY_E_CUDA",
    E_OOM: "SOFTRAY_E_OOM",
    E_UNSUPPORTED: "SOFTRAY_E_UNSUPPORTED",
    E_FORMAT: "SOFTRAY_E_FORMAT",
    E_TIMEOUT: "SOFTRAY_E_TIMEOUT",
}

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_uint32_p = C.POINTER(C.c_uint32)


class Mesh(C.Structure):
    _fields_ = [
        ("verts_xyz", c_double_p),
        ("tri_vidx", c_int32_p),
        ("tri_argb", c_uint32_p),
        ("n_verts", C.c_int32),
        ("n_tris", C.c_int32),
        ("bbox_min", C.c_double * 3),
        ("bbox_max", C.c_double * 3),
    ]


class Sphere(C.Structure):
    _fields_ = [
        ("cx", C.c_double),
        ("cy", C.c_double),
        ("cz", C.c_double),
        ("r", C.c_double),
        ("argb", C.c_uint32),
        ("_pad", C.c_uint32),
    ]


class SceneDesc(C.Structure):
    _fields_ = [
        ("meshes", C.POINTER(Mesh)),
        ("n_meshes", C.c_int32),
        ("accel", C.c_int32),
        ("spheres", C.POINTER(Sphere)),
        ("n_spheres", C.c_int32),
        ("_pad", C.c_int32),
    ]


class Instance(C.Structure):
    _fields_ = [
        ("M", C.c_double * 16),
        ("Minv", C.c_double * 16),
        ("pos", C.c_double * 3),
        ("mesh_id", C.c_int32),
        ("_pad", C.c_int32),
    ]


class Frame(C.Structure):
    _fields_ = [
        ("ambient", C.c_double),
        ("shininess", C.c_double),
        ("light_dir_view", C.c_double * 3),
        ("light_pos_view", C.c_double * 3),
        ("fov_depth", C.c_double),
        ("focal_depth", C.c_double),
        ("focal_strength", C.c_double),
        ("instances", C.POINTER(Instance)),
        ("n_instances", C.c_int32),
        ("width", C.c_int32),
        ("height", C.c_int32),
        ("start_row", C.c_int32),
        ("end_row", C.c_int32),
        ("sub_pixel_res", C.c_int32),
        ("focal_blur", C.c_int32),
        ("subdivision", C.c_int32),
        ("shading", C.c_int32),
        ("shadows", C.c_int32),
        ("shadow_samples", C.c_int32),
        ("point_lighting", C.c_int32),
        ("specular_lighting", C.c_int32),
        ("random_seed", C.c_int32),
        ("background_argb", C.c_uint32),
        ("reflection_depth", C.c_int32),
        ("texture3d_id", C.c_int32),
        ("band_height", C.c_int32),
        ("band_count", C.c_int32),
        ("band_index", C.c_int32),
        ("filter_mode", C.c_int32),
        ("profile_stages", C.c_int32),
        ("_reserved", C.c_int32 * 2),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("rays_primary", C.c_uint64),
        ("rays_shadow", C.c_uint64),
        ("rays_secondary", C.c_uint64),
        ("node_visits", C.c_uint64),
        ("prim_tests", C.c_uint64),
        ("sphere_tests", C.c_uint64),
        ("hits_primary", C.c_uint64),
        ("shaded_hits", C.c_uint64),
        ("launches", C.c_uint64),
        ("filter_tests", C.c_uint64),
        ("filter_unsure", C.c_uint64),
        ("filter_mismatch", C.c_uint64),
        ("rays_bundled", C.c_uint64),
        ("rays_fallback", C.c_uint64),
        ("rays_short_listed", C.c_uint64),
        ("ms_kernel", C.c_double),
        ("ms_h2d", C.c_double),
        ("ms_d2h", C.c_double),
        ("ms_total", C.c_double),
        ("ms_stage", C.c_double * 10),
    ]

    def as_dict(self):
        return {name: (list(getattr(self, name)) if name == "ms_stage" else getattr(self, name)) for name, _ in self._fields_}

    @property
    def stages(self):
        """ms per stage of a frame rendered with profile_stages (empty for the fused kernel)."""
        return {n: self.ms_stage[i] for i, n in enumerate(STAGE_NAMES) if self.ms_stage[i] > 0.0}

    @property
    def rays(self):
        return self.rays_primary + self.rays_shadow + self.rays_secondary


EXPECTED_SIZES = {"mesh": 80, "sphere": 40, "scene_desc": 32, "instance": 288, "frame": 192, "stats": 232}

assert C.sizeof(Mesh) == EXPECTED_SIZES["mesh"]
assert C.sizeof(Sphere) == EXPECTED_SIZES["sphere"]
assert C.sizeof(SceneDesc) == EXPECTED_SIZES["scene_desc"]
assert C.sizeof(Instance) == EXPECTED_SIZES["instance"]
assert C.sizeof(Frame) == EXPECTED_SIZES["frame"]
assert C.sizeof(Stats) == EXPECTED_SIZES["stats"]
