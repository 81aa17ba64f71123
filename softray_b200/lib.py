"""ctypes loader for libsoftray_cuda.so and thin object wrappers over its C ABI.

No fallback of any kind lives here: a missing library raises ImportError-like RuntimeError at
load time, a missing CUDA device makes Context() raise SoftRayError(E_NO_DEVICE).
"""
import ctypes as C
import os
import subprocess
import weakref

import numpy as np

from . import abi
from .scene import FrameParams, MeshData, SceneDescHolder, SphereData

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("SOFTRAY_SO") or os.path.join(_HERE, "libsoftray_cuda.so")   # SOFTRAY_SO: an experimental build (scripts/build_variant.py)
CSRC = os.path.join(_HERE, "csrc")
SOURCES = ["sr_api.cu", "sr_render.cu", "sr_diag.cu", "sr_bvh.cpp", "sr_model3ds.cpp", "sr_resolve.cu", "sr_lbvh.cu", "sr_wave.cu"]
HEADERS = ["sr_types.h", "sr_bvh.h", "sr_device.cuh", "sr_wave.h", os.path.join("..", "..", "include", "softray_cuda.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
    "-Xcompiler", "-fPIC,-ffp-contract=off",
]

# every symbol include/softray_cuda.h declares
EXPORTS = [
    "softray_create", "softray_create_multi", "softray_device_count", "softray_destroy", "softray_last_error", "softray_abi_version", "softray_abi_sizeof",
    "softray_scene_create", "softray_scene_destroy", "softray_scene_fingerprint",
    "softray_render", "softray_render_device", "softray_instance_init", "softray_frame_defaults",
    "softray_device_alloc", "softray_device_free", "softray_ipc_export", "softray_ipc_open", "softray_ipc_close",
    "softray_host_register", "softray_host_unregister", "softray_host_barrier",
    "softray_measure_fma_peak", "softray_model_load_3ds", "softray_model_get_mesh", "softray_model_destroy",
    "softray_resolve", "softray_resolve_device",
]

_lib = None


class SoftRayError(RuntimeError):
    """A negative SOFTRAY_E_* code from the library (the shim maps these back to the .NET
    exception types the reference throws: include/softray_cuda.h)."""

    def __init__(self, code, message):
        super().__init__(f"{abi.ERROR_NAMES.get(code, code)}: {message}")
        self.code = code


def build(force=False, verbose=False):
    """Compile libsoftray_cuda.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, h) for h in HEADERS]
    stale = (not os.path.exists(SO_PATH)) or any(os.path.getmtime(p) > os.path.getmtime(SO_PATH) for p in deps)
    if not (force or stale):
        return SO_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO_PATH] + srcs
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return SO_PATH


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise RuntimeError(f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
    L = C.CDLL(SO_PATH)
    vp = C.c_void_p
    L.softray_create.argtypes = [C.c_int32, C.POINTER(vp)]
    L.softray_create_multi.argtypes = [C.c_int32, C.POINTER(vp)]
    L.softray_device_count.argtypes = [vp]
    L.softray_destroy.argtypes = [vp]
    L.softray_destroy.restype = None
    L.softray_last_error.argtypes = [vp]
    L.softray_last_error.restype = C.c_char_p
    L.softray_abi_version.restype = C.c_int
    L.softray_abi_sizeof.argtypes = [C.c_int32]
    L.softray_scene_create.argtypes = [vp, C.POINTER(abi.SceneDesc), C.POINTER(vp)]
    L.softray_scene_destroy.argtypes = [vp]
    L.softray_scene_destroy.restype = None
    L.softray_scene_fingerprint.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.softray_render.argtypes = [vp, vp, C.POINTER(abi.Frame), vp, vp, C.POINTER(abi.Stats)]
    L.softray_render_device.argtypes = [vp, vp, C.POINTER(abi.Frame), vp, vp, vp, C.POINTER(abi.Stats)]
    L.softray_instance_init.argtypes = [C.POINTER(abi.Instance), abi.c_double_p, C.c_double, C.c_double, C.c_double,
                                        C.c_int32]
    L.softray_instance_init.restype = None
    L.softray_frame_defaults.argtypes = [C.POINTER(abi.Frame), C.c_int32, C.c_int32]
    L.softray_frame_defaults.restype = None
    L.softray_device_alloc.argtypes = [vp, C.c_uint64, C.POINTER(vp)]
    L.softray_device_free.argtypes = [vp, vp]
    L.softray_measure_fma_peak.argtypes = [vp, C.c_int32, C.POINTER(C.c_double)]
    L.softray_ipc_export.argtypes = [vp, vp, C.c_char_p]
    L.softray_ipc_open.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
    L.softray_ipc_close.argtypes = [vp, vp]
    L.softray_host_register.argtypes = [vp, vp, C.c_uint64]
    L.softray_host_unregister.argtypes = [vp, vp]
    L.softray_host_barrier.argtypes = [vp, C.c_uint32]
    L.softray_resolve.argtypes = [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_uint32, vp]
    L.softray_resolve_device.argtypes = [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_uint32, vp, vp]
    L.softray_model_load_3ds.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(vp)]
    L.softray_model_get_mesh.argtypes = [vp, C.POINTER(abi.Mesh)]
    L.softray_model_destroy.argtypes = [vp]
    L.softray_model_destroy.restype = None
    if L.softray_abi_version() != abi.ABI_VERSION:
        raise RuntimeError("libsoftray_cuda.so ABI version mismatch")
    for which, name in enumerate(["mesh", "sphere", "scene_desc", "instance", "frame", "stats"]):
        if L.softray_abi_sizeof(which) != abi.EXPECTED_SIZES[name]:
            raise RuntimeError(f"softray_{name} layout mismatch between abi.py and the library")
    _lib = L
    return L


def _check(rc, ctx_handle, what):
    if rc != abi.OK:
        msg = load().softray_last_error(ctx_handle)
        raise SoftRayError(rc, f"{what}: {msg.decode('utf-8', 'replace') if msg else ''}")


def load_3ds(data: bytes) -> MeshData:
    """softray_model_load_3ds: a .3DS byte stream -> MeshData (Model.Load3ds + PostProcessGeometry +
    per-triangle colour packing, in native code; no device needed)."""
    L = load()
    h = C.c_void_p()
    _check(L.softray_model_load_3ds(data, len(data), C.byref(h)), None, "softray_model_load_3ds")
    try:
        m = abi.Mesh()
        _check(L.softray_model_get_mesh(h, C.byref(m)), None, "softray_model_get_mesh")
        nv, nt = m.n_verts, m.n_tris
        verts = np.ctypeslib.as_array(m.verts_xyz, shape=(nv, 3)).copy() if nv else np.zeros((0, 3))
        tris = np.ctypeslib.as_array(m.tri_vidx, shape=(nt, 3)).copy() if nt else np.zeros((0, 3), dtype=np.int32)
        argb = np.ctypeslib.as_array(m.tri_argb, shape=(nt,)).copy() if nt else np.zeros((0,), dtype=np.uint32)
        return MeshData(verts, tris, argb, np.array(list(m.bbox_min)), np.array(list(m.bbox_max)))
    finally:
        L.softray_model_destroy(h)


class Context:
    """softray_ctx: one CUDA device, its stream and scratch buffers."""

    def __init__(self, device=0, n_devices=None):
        """device: one CUDA device (softray_create).  n_devices = k: a group context over the first k devices of this
        process, 0 = all of them (softray_create_multi): scenes are replicated, frames split into row bands."""
        self._h = C.c_void_p()
        if n_devices is None:
            rc = load().softray_create(int(device), C.byref(self._h))
        else:
            rc = load().softray_create_multi(int(n_devices), C.byref(self._h))
        if rc != abi.OK:
            self._h = C.c_void_p()
            _check(rc, None, "softray_create")
        self.device = int(device) if n_devices is None else 0
        self.n_devices = load().softray_device_count(self._h)
        self._scenes = weakref.WeakSet()     # scenes hold device memory of this context: they go first

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            for sc in list(getattr(self, "_scenes", ())):
                sc.close()
            load().softray_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def measure_fma_peak(self, fp64=True):
        """Measured FMA-issue peak of this device, TFLOP/s (FMA = 2 flops)."""
        out = C.c_double()
        _check(load().softray_measure_fma_peak(self._h, int(bool(fp64)), C.byref(out)), self._h, "softray_measure_fma_peak")
        return out.value

    def resolve(self, src, aa_res=1, style=0, background=0):
        """softray_resolve: PostProcessImage + AntiAliasImage of a host (H*aa, W*aa) uint32 surface."""
        src = np.ascontiguousarray(src, dtype=np.uint32)
        k = max(int(aa_res), 1)              # (a bad aa_res is the library's to reject)
        h, w = src.shape[0] // k, src.shape[1] // k
        assert src.shape == (h * k, w * k)
        dst = np.empty((h, w), dtype=np.uint32)
        _check(load().softray_resolve(self._h, src.ctypes.data_as(C.c_void_p), w, h, int(aa_res), int(style),
                                      int(background) & 0xFFFFFFFF, dst.ctypes.data_as(C.c_void_p)), self._h, "softray_resolve")
        return dst

    def resolve_device(self, d_src, d_dst, width, height, aa_res=1, style=0, background=0, stream=None):
        """softray_resolve_device on raw device pointers (ints)."""
        _check(load().softray_resolve_device(self._h, C.c_void_p(d_src), int(width), int(height), int(aa_res), int(style),
                                             int(background) & 0xFFFFFFFF, C.c_void_p(d_dst),
                                             C.c_void_p(stream) if stream else None), self._h, "softray_resolve_device")

    # IPC helpers for the peer-mapped framebuffer (multi-GPU gather fused into the render kernel)
    def device_alloc(self, nbytes):
        out = C.c_void_p()
        _check(load().softray_device_alloc(self._h, int(nbytes), C.byref(out)), self._h, "softray_device_alloc")
        return out.value

    def device_free(self, device_ptr):
        _check(load().softray_device_free(self._h, C.c_void_p(device_ptr)), self._h, "softray_device_free")

    def ipc_export(self, device_ptr):
        buf = C.create_string_buffer(64)
        _check(load().softray_ipc_export(self._h, C.c_void_p(device_ptr), buf), self._h, "softray_ipc_export")
        return buf.raw

    def ipc_open(self, handle_bytes):
        out = C.c_void_p()
        _check(load().softray_ipc_open(self._h, handle_bytes, C.byref(out)), self._h, "softray_ipc_open")
        return out.value

    def host_register(self, array):
        """softray_host_register on a numpy array's memory (page-lock it for zero-copy frames)."""
        _check(load().softray_host_register(self._h, C.c_void_p(array.ctypes.data), int(array.nbytes)), self._h, "softray_host_register")

    def host_unregister(self, array):
        _check(load().softray_host_unregister(self._h, C.c_void_p(array.ctypes.data)), self._h, "softray_host_unregister")

    def ipc_close(self, device_ptr):
        _check(load().softray_ipc_close(self._h, C.c_void_p(device_ptr)), self._h, "softray_ipc_close")


class Scene:
    """softray_scene: device-resident SoA triangles / spheres + BVHs."""

    def __init__(self, ctx: Context, meshes, spheres: SphereData = None, accel=abi.ACCEL_BVH):
        self.ctx = ctx
        self.holder = SceneDescHolder(meshes, spheres, accel)
        self._h = C.c_void_p()
        _check(load().softray_scene_create(ctx._h, C.byref(self.holder.desc), C.byref(self._h)), ctx._h,
               "softray_scene_create")
        ctx._scenes.add(self)

    def fingerprint(self):
        out = C.c_uint64()
        _check(load().softray_scene_fingerprint(self._h, C.byref(out)), self.ctx._h, "softray_scene_fingerprint")
        return out.value

    def render(self, params: FrameParams, want_ids=False, pixels=None, ids=None, want_stats=True):
        """softray_render into host numpy buffers.  Returns dict(pixels, ids, stats)."""
        L = load()
        f = params.to_c(L.softray_instance_init)
        W, H = params.width, params.height
        px = pixels if pixels is not None else np.zeros((H, W), dtype=np.uint32)
        if want_ids and ids is None:
            ids = np.full((H, W), -1, dtype=np.int32)
        st = abi.Stats() if want_stats else None
        rc = L.softray_render(self.ctx._h, self._h, C.byref(f), px.ctypes.data_as(C.c_void_p),
                              ids.ctypes.data_as(C.c_void_p) if ids is not None else None,
                              C.byref(st) if st is not None else None)
        _check(rc, self.ctx._h, "softray_render")
        return dict(pixels=px, ids=ids, stats=st)

    def render_device(self, params: FrameParams, d_pixels, d_ids=None, stream=None, want_stats=False, c_frame=None):
        """softray_render_device: d_pixels / d_ids are raw device pointers (ints), stream a
        cudaStream_t handle (int) or None for the context's own stream."""
        L = load()
        f = c_frame if c_frame is not None else params.to_c(L.softray_instance_init)
        st = abi.Stats() if want_stats else None
        rc = L.softray_render_device(self.ctx._h, self._h, C.byref(f), C.c_void_p(d_pixels),
                                     C.c_void_p(d_ids) if d_ids else None, C.c_void_p(stream) if stream else None,
                                     C.byref(st) if st is not None else None)
        _check(rc, self.ctx._h, "softray_render_device")
        return st

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            load().softray_scene_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
