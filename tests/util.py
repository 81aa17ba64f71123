"""Shared helpers: the reference's RendererSetup/RaytraceScenario (RendererTests.cs:65-96,381-459)
expressed as FrameParams, and image comparison."""
import math

import numpy as np

from softray_b200 import FrameParams, InstanceData, SphereData

DEFAULT_YAW_DEG = 135.0      # RendererTests.cs:47
DEFAULT_PITCH_DEG = -22.0    # RendererTests.cs:48
DEFAULT_ROLL_DEG = 0.0       # RendererTests.cs:49
BACKGROUND = 0xFF00FF        # RendererTests.cs:68


def scenario(resolution=100, shading=True, focal_blur=False, shadows=False, sub_pixel_res=1,
             pitch_deg=DEFAULT_PITCH_DEG, yaw_deg=DEFAULT_YAW_DEG, roll_deg=DEFAULT_ROLL_DEG,
             focal_depth=-1.0, object_depth=1.0, width=None, height=None, **extra):
    """RaytraceScenario's flag handling (RendererTests.cs:381-399)."""
    inst = InstanceData(position=(0.0, 0.0, object_depth),
                        yaw=yaw_deg / 180.0 * math.pi, pitch=pitch_deg / 180.0 * math.pi,
                        roll=roll_deg / 180.0 * math.pi)
    p = FrameParams(width=width or resolution, height=height or resolution, instances=[inst], background=BACKGROUND,
                    shading=shading, focal_blur=focal_blur, shadows=shadows, sub_pixel_res=sub_pixel_res,
                    focal_depth=(object_depth + 0.5 if abs(focal_depth + 1) < 0.001 else focal_depth))
    for k, v in extra.items():
        assert hasattr(p, k), k
        setattr(p, k, v)
    return p


def golden_name(shading=True, focal_blur=False, shadows=False, sub_pixel_res=1, path_tracing=False, n_geometry=0):
    """Test name construction (RendererTests.cs:425-436)."""
    name = ("pathTracing_" if path_tracing else "") + ("shading" if shading else "noShading")
    name += "_shadows" if shadows else ""
    name += "_focalBlur" if focal_blur else ""
    name += ("x%d" % sub_pixel_res) if focal_blur else (("_%dxAA" % sub_pixel_res) if sub_pixel_res > 1 else "")
    name += ("_%d_geometry" % n_geometry) if n_geometry else ""
    return name


def rgb(pixels):
    return np.asarray(pixels, dtype=np.uint32) & 0xFFFFFF


def count_diff(a, b):
    return int((rgb(a) != rgb(b)).sum())


def channel_absdiff(a, b):
    """max over channels of |a-b| per pixel, int array."""
    a = rgb(a)
    b = rgb(b)
    out = np.zeros(a.shape, dtype=np.int32)
    for sh in (0, 8, 16):
        ca = ((a >> sh) & 0xFF).astype(np.int32)
        cb = ((b >> sh) & 0xFF).astype(np.int32)
        out = np.maximum(out, np.abs(ca - cb))
    return out


def path_trace_spheres():
    """PathTracePrimitivesTest's five spheres (RendererTests.cs:251-257); Color.*.ToARGB()."""
    c = np.array([[0, -10000, 0, 9999.5], [-0.5, 0, -0.5, 0.5], [+0.5, 0, +0.5, 0.5], [+0.5, 0, -0.5, 0.5],
                  [-0.5, 0, +0.5, 0.5]], dtype=np.float64)
    argb = np.array([0xFFFFFFFF, 0xFFFF0000, 0xFF00FF00, 0xFF0000FF, 0xFFFFFF00], dtype=np.uint32)
    return SphereData(c, argb)
