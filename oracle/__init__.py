"""ctypes binding of the CPU oracle (oracle/softray_oracle.c).  TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs -- never from
softray_b200/.  The shared object is built on demand with oracle/Makefile (plain gcc, seconds).
"""
import ctypes as C
import os
import subprocess

import numpy as np

from softray_b200 import abi
from softray_b200.scene import FrameParams, MeshData, SceneDescHolder, SphereData

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsoftray_oracle.so")
_lib = None


class OracleError(RuntimeError):
    def __init__(self, code, what):
        super().__init__(f"{what}: {abi.ERROR_NAMES.get(code, code)}")
        self.code = code


class Hit(C.Structure):
    _fields_ = [
        ("ray_frac", C.c_double),
        ("pos", C.c_double * 3),
        ("normal", C.c_double * 3),
        ("color", C.c_uint32),
        ("tri_index", C.c_int32),
        ("prim_id", C.c_int32),
        ("_pad", C.c_int32),
    ]


class Random(C.Structure):
    _fields_ = [("a", C.c_int32 * 56), ("inext", C.c_int32), ("inextp", C.c_int32)]


class Options(C.Structure):
    _fields_ = [
        ("concurrency", C.c_int32),
        ("n_threads", C.c_int32),
        ("tree_max_depth", C.c_int32),
        ("tree_max_per_node", C.c_int32),
        ("path_tracing", C.c_int32),
        ("_pad", C.c_int32),
        ("col_start", C.c_int32),
        ("col_end", C.c_int32),
    ]


class Aux(C.Structure):
    _fields_ = [("ray_frac", abi.c_double_p), ("cos_theta", abi.c_double_p)]


class CModel(C.Structure):
    _fields_ = [
        ("verts_xyz", abi.c_double_p),
        ("n_verts", C.c_int32),
        ("n_tris", C.c_int32),
        ("tri_vidx", abi.c_int32_p),
        ("tri_argb", abi.c_uint32_p),
        ("bbox_min", C.c_double * 3),
        ("bbox_max", C.c_double * 3),
    ]


def build(force=False):
    src = os.path.join(_HERE, "softray_oracle.c")
    hdrs = [os.path.join(_HERE, "softray_oracle.h"), os.path.join(_HERE, "..", "include", "softray_cuda.h")]
    stale = (not os.path.exists(_SO)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_SO) for p in [src] + hdrs
    )
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libsoftray_oracle.so"], check=True, capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_SO)
    dp = abi.c_double_p
    L.orc_random_init.argtypes = [C.POINTER(Random), C.c_int32]
    L.orc_random_next.argtypes = [C.POINTER(Random)]
    L.orc_random_next.restype = C.c_int32
    L.orc_random_next_double.argtypes = [C.POINTER(Random)]
    L.orc_random_next_double.restype = C.c_double
    L.orc_triangle_intersect.argtypes = [dp, dp, dp, C.c_uint32, dp, dp, C.POINTER(Hit)]
    L.orc_sphere_intersect.argtypes = [dp, C.c_double, C.c_uint32, dp, dp, C.POINTER(Hit)]
    L.orc_sphere_contains_point.argtypes = [dp, C.c_double, dp]
    L.orc_box_contains_point.argtypes = [dp, dp, dp]
    L.orc_box_clip_line_segment.argtypes = [dp, dp, dp, dp]
    L.orc_tree_build.argtypes = [dp, abi.c_uint32_p, C.c_int32, dp, dp, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
    L.orc_tree_free.argtypes = [C.c_void_p]
    L.orc_tree_free.restype = None
    L.orc_tree_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int32 * 6)]
    L.orc_tree_stats.restype = None
    L.orc_tree_intersect.argtypes = [C.c_void_p, dp, dp, C.POINTER(Hit)]
    L.orc_tree_brute_intersect.argtypes = [C.c_void_p, dp, dp, C.POINTER(Hit)]
    L.orc_model_load_3ds.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.POINTER(CModel))]
    L.orc_model_from_arrays.argtypes = [dp, C.c_int32, abi.c_int32_p, abi.c_uint32_p, C.c_int32, C.c_int32,
                                        C.POINTER(C.POINTER(CModel))]
    L.orc_model_free.argtypes = [C.POINTER(CModel)]
    L.orc_model_free.restype = None
    L.orc_options_defaults.argtypes = [C.POINTER(Options)]
    L.orc_options_defaults.restype = None
    L.orc_scene_create.argtypes = [C.POINTER(abi.SceneDesc), C.POINTER(Options), C.POINTER(C.c_void_p)]
    L.orc_scene_free.argtypes = [C.c_void_p]
    L.orc_scene_free.restype = None
    L.orc_scene_tree_stats.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int32 * 6)]
    L.orc_scene_tree_stats.restype = None
    L.orc_render.argtypes = [C.c_void_p, C.POINTER(abi.Frame), C.POINTER(Options), abi.c_uint32_p, abi.c_int32_p,
                             C.POINTER(Aux), C.POINTER(abi.Stats)]
    L.orc_area_light_offsets.argtypes = [C.c_int32, C.c_int32, dp]
    L.orc_area_light_offsets.restype = None
    L.orc_texture3d_sample.argtypes = [C.c_int32, dp]
    L.orc_texture3d_sample.restype = C.c_uint8
    L.orc_instance_init.argtypes = [C.POINTER(abi.Instance), dp, C.c_double, C.c_double, C.c_double, C.c_int32]
    L.orc_instance_init.restype = None
    L.orc_frame_defaults.argtypes = [C.POINTER(abi.Frame), C.c_int32, C.c_int32]
    L.orc_frame_defaults.restype = None
    _lib = L
    return L


def _d3(v):
    return (C.c_double * 3)(*[float(x) for x in v])


# ---------------------------------------------------------------- System.Random
class SystemRandom:
    """System.Random(seed) (.NET Framework 4.x), SURVEY.md Appendix B."""

    def __init__(self, seed):
        self._r = Random()
        lib().orc_random_init(C.byref(self._r), int(seed))

    def next(self):
        return lib().orc_random_next(C.byref(self._r))

    def next_double(self):
        return lib().orc_random_next_double(C.byref(self._r))


# ---------------------------------------------------------------- primitives
def _hit_to_dict(h):
    return dict(ray_frac=h.ray_frac, pos=tuple(h.pos), normal=tuple(h.normal), color=h.color, tri_index=h.tri_index,
                prim_id=h.prim_id)


def triangle_intersect(v1, v2, v3, start, direction, color=0xFFFFFFFF):
    h = Hit()
    ok = lib().orc_triangle_intersect(_d3(v1), _d3(v2), _d3(v3), color, _d3(start), _d3(direction), C.byref(h))
    return _hit_to_dict(h) if ok else None


def sphere_intersect(center, radius, start, direction, color=0xFFFFFFFF):
    h = Hit()
    ok = lib().orc_sphere_intersect(_d3(center), float(radius), color, _d3(start), _d3(direction), C.byref(h))
    return _hit_to_dict(h) if ok else None


def sphere_contains_point(center, radius, pt):
    return bool(lib().orc_sphere_contains_point(_d3(center), float(radius), _d3(pt)))


def box_contains_point(bmin, bmax, pt):
    return bool(lib().orc_box_contains_point(_d3(bmin), _d3(bmax), _d3(pt)))


def box_clip_line_segment(bmin, bmax, start, end):
    s, e = _d3(start), _d3(end)
    ok = lib().orc_box_clip_line_segment(_d3(bmin), _d3(bmax), s, e)
    return (tuple(s), tuple(e)) if ok else None


def area_light_offsets(seed=1234567890, n=100):
    out = np.zeros((n, 3), dtype=np.float64)
    lib().orc_area_light_offsets(seed, n, out.ctypes.data_as(abi.c_double_p))
    return out


def texture3d_sample(tex_id, pos):
    return int(lib().orc_texture3d_sample(tex_id, _d3(pos)))


# ---------------------------------------------------------------- SpatialSubdivision
class Tree:
    """SpatialSubdivision over explicit triangles (n,3,3) float64."""

    def __init__(self, tri_verts, bbox_min, bbox_max, max_tree_depth=15, max_geometry_per_node=25, colors=None):
        tv = np.ascontiguousarray(tri_verts, dtype=np.float64).reshape(-1, 9)
        self.n = tv.shape[0]
        col = None
        if colors is not None:
            col = np.ascontiguousarray(colors, dtype=np.uint32)
        self._h = C.c_void_p()
        rc = lib().orc_tree_build(tv.ctypes.data_as(abi.c_double_p),
                                  col.ctypes.data_as(abi.c_uint32_p) if col is not None else None, self.n,
                                  _d3(bbox_min), _d3(bbox_max), max_tree_depth, max_geometry_per_node, C.byref(self._h))
        if rc != abi.OK:
            raise OracleError(rc, "orc_tree_build")

    def stats(self):
        out = (C.c_int32 * 6)()
        lib().orc_tree_stats(self._h, C.byref(out))
        return dict(depth=out[0], nodes=out[1], leaves=out[2], internal=out[3], refs=out[4], largest_leaf=out[5])

    def intersect(self, start, direction):
        h = Hit()
        return _hit_to_dict(h) if lib().orc_tree_intersect(self._h, _d3(start), _d3(direction), C.byref(h)) else None

    def brute_intersect(self, start, direction):
        h = Hit()
        return _hit_to_dict(h) if lib().orc_tree_brute_intersect(self._h, _d3(start), _d3(direction), C.byref(h)) else None

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.orc_tree_free(self._h)
            self._h = None


# ---------------------------------------------------------------- Model
def _model_to_meshdata(pm):
    m = pm.contents
    nv, nt = m.n_verts, m.n_tris
    verts = np.ctypeslib.as_array(m.verts_xyz, shape=(nv * 3,)).reshape(nv, 3).copy() if nv else np.zeros((0, 3))
    tris = np.ctypeslib.as_array(m.tri_vidx, shape=(nt * 3,)).reshape(nt, 3).copy() if nt else np.zeros((0, 3), np.int32)
    argb = np.ctypeslib.as_array(m.tri_argb, shape=(nt,)).copy() if nt else np.zeros((0,), np.uint32)
    md = MeshData(verts, tris, argb, np.array(list(m.bbox_min)), np.array(list(m.bbox_max)))
    lib().orc_model_free(pm)
    return md


def load_3ds(data: bytes) -> MeshData:
    """Model.Load3ds + PostProcessGeometry (Model.cs:522-653,750-831)."""
    pm = C.POINTER(CModel)()
    rc = lib().orc_model_load_3ds(bytes(data), len(data), C.byref(pm))
    if rc != abi.OK:
        raise OracleError(rc, "orc_model_load_3ds")
    return _model_to_meshdata(pm)


def model_from_arrays(verts, tris, argb=None, normalise=True) -> MeshData:
    """The Cloth.cs pattern: fill lists, CalcExtent(), PostProcessGeometry()."""
    verts = np.ascontiguousarray(verts, dtype=np.float64).reshape(-1, 3)
    tris = np.ascontiguousarray(tris, dtype=np.int32).reshape(-1, 3)
    col = np.ascontiguousarray(argb, dtype=np.uint32) if argb is not None else None
    pm = C.POINTER(CModel)()
    rc = lib().orc_model_from_arrays(verts.ctypes.data_as(abi.c_double_p), verts.shape[0],
                                     tris.ctypes.data_as(abi.c_int32_p),
                                     col.ctypes.data_as(abi.c_uint32_p) if col is not None else None, tris.shape[0],
                                     int(bool(normalise)), C.byref(pm))
    if rc != abi.OK:
        raise OracleError(rc, "orc_model_from_arrays")
    return _model_to_meshdata(pm)


# ---------------------------------------------------------------- Scene / render
def default_options(**kw):
    o = Options()
    lib().orc_options_defaults(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, int(v))
    return o


class Scene:
    def __init__(self, meshes, spheres=None, options=None):
        self.holder = SceneDescHolder(meshes, spheres)
        self.options = options or default_options()
        self._h = C.c_void_p()
        rc = lib().orc_scene_create(C.byref(self.holder.desc), C.byref(self.options), C.byref(self._h))
        if rc != abi.OK:
            raise OracleError(rc, "orc_scene_create")

    def tree_stats(self, mesh=0):
        out = (C.c_int32 * 6)()
        lib().orc_scene_tree_stats(self._h, mesh, C.byref(out))
        return dict(depth=out[0], nodes=out[1], leaves=out[2], internal=out[3], refs=out[4], largest_leaf=out[5])

    def render(self, params: FrameParams, options=None, want_ids=False, want_aux=False, pixels=None):
        """Returns dict(pixels=(H,W) uint32, ids, ray_frac, cos_theta, stats)."""
        f = params.to_c(lib().orc_instance_init)
        W, H = params.width, params.height
        px = pixels if pixels is not None else np.zeros((H, W), dtype=np.uint32)
        ids = np.full((H, W), -1, dtype=np.int32) if want_ids else None
        aux = None
        rf = ct = None
        if want_aux:
            rf = np.full((H, W), np.nan)
            ct = np.full((H, W), np.nan)
            aux = Aux(rf.ctypes.data_as(abi.c_double_p), ct.ctypes.data_as(abi.c_double_p))
        st = abi.Stats()
        opt = options or self.options
        rc = lib().orc_render(self._h, C.byref(f), C.byref(opt), px.ctypes.data_as(abi.c_uint32_p),
                              ids.ctypes.data_as(abi.c_int32_p) if ids is not None else None,
                              C.byref(aux) if aux is not None else None, C.byref(st))
        if rc != abi.OK:
            raise OracleError(rc, "orc_render")
        return dict(pixels=px, ids=ids, ray_frac=rf, cos_theta=ct, stats=st)

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.orc_scene_free(self._h)
            self._h = None


def resolve(src, aa_res=1, style=0, background=0):
    """Renderer.PostProcessImage (Renderer.cs:819-898; Standard / ColorShuffle / Negative) followed by
    Renderer.AntiAliasImage (Renderer.cs:937-978), restated with numpy integer arithmetic.
    src: (H*aa, W*aa) uint32.  TEST INFRASTRUCTURE ONLY."""
    x = np.ascontiguousarray(src, dtype=np.uint32).copy()
    if style == 1:      # ColorShuffle: ((x & 0xffff) << 8) + ((x >> 16) & 0xff)   (:829)
        x = ((x & np.uint32(0xFFFF)) << np.uint32(8)) + ((x >> np.uint32(16)) & np.uint32(0xFF))
    elif style == 2:    # Negative: x == BackgroundColor ? BackgroundColor : 0x00ffffff - x, unchecked uint   (:833)
        with np.errstate(over="ignore"):
            x = np.where(x == np.uint32(background), np.uint32(background), np.uint32(0x00FFFFFF) - x).astype(np.uint32)
    elif style != 0:
        raise OracleError(abi.E_UNSUPPORTED, "resolve style")
    if aa_res < 2:
        return x
    h, w = x.shape[0] // aa_res, x.shape[1] // aa_res
    blocks = x.reshape(h, aa_res, w, aa_res)
    out = np.full((h, w), 255 << 24, dtype=np.uint64)
    for sh in (16, 8, 0):   # Surface.UnpackRgb, int sums, truncating divide, Surface.PackRgb (:946-975)
        s_ = ((blocks >> np.uint32(sh)) & np.uint32(0xFF)).astype(np.int64).sum(axis=(1, 3)) // (aa_res * aa_res)
        out += (s_.astype(np.uint64) & np.uint64(0xFF)) << np.uint64(sh)
    return out.astype(np.uint32)
