#!/usr/bin/env python3
"""Regenerate tests/golden/reference_fixtures.npz from the read-only reference checkout.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_fixtures.py [/root/reference]

What is captured (data fixtures only, no reference source code):
  * model/<name>      raw bytes of the reference's test models
                      (Engine3D/Engine3D-Tests/obj.3ds, obj2.3DS)
  * golden/<WxH>/<name>  uint32 [H,W] 0x00RRGGBB pixels of every raytrace golden BMP whose file
                      name carries no out-of-scope decorator token (_AO, _lightField*,
                      _staticShadows, voxels_) -- see SURVEY.md section 8(c).
The BMPs are 32-bpp bottom-up BGRX (written by System.Drawing from Format32bppRgb,
RendererTests.cs:517-526); we flip them to top-down and drop the X byte so that
pixel [y,x] == renderer pixels[y*W+x] & 0xFFFFFF.
"""
import os
import struct
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
TESTS = os.path.join(REF, "Engine3D", "Engine3D-Tests")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_fixtures.npz")
SKIP_TOKENS = ("_AO", "_lightField", "_staticShadows", "voxels_")


def read_bmp(path):
    with open(path, "rb") as f:
        data = f.read()
    assert data[:2] == b"BM", path
    off = struct.unpack_from("<I", data, 10)[0]
    hdr = struct.unpack_from("<I", data, 14)[0]
    w, h = struct.unpack_from("<ii", data, 18)
    planes, bpp = struct.unpack_from("<HH", data, 26)
    comp = struct.unpack_from("<I", data, 30)[0]
    assert hdr >= 40 and planes == 1 and comp in (0, 3), (path, hdr, planes, comp)
    bottom_up = h > 0
    h = abs(h)
    if bpp == 32:
        px = np.frombuffer(data, dtype="<u4", count=w * h, offset=off).reshape(h, w) & 0xFFFFFF
    elif bpp == 24:
        stride = (w * 3 + 3) & ~3
        raw = np.frombuffer(data, dtype=np.uint8, count=stride * h, offset=off).reshape(h, stride)
        raw = raw[:, : w * 3].reshape(h, w, 3).astype(np.uint32)
        px = raw[..., 0] | (raw[..., 1] << 8) | (raw[..., 2] << 16)
    else:
        raise ValueError(f"{path}: unsupported bpp {bpp}")
    if bottom_up:
        px = px[::-1]
    return np.ascontiguousarray(px.astype(np.uint32))


def main():
    out = {}
    for name in ("obj.3ds", "obj2.3DS"):
        with open(os.path.join(TESTS, name), "rb") as f:
            out["model/" + name.lower()] = np.frombuffer(f.read(), dtype=np.uint8)
    base = os.path.join(TESTS, "baseline images", "raytrace")
    n = 0
    for res in sorted(os.listdir(base)):
        for fn in sorted(os.listdir(os.path.join(base, res))):
            if not fn.endswith(".bmp") or any(t in fn for t in SKIP_TOKENS):
                continue
            out[f"golden/{res}/{fn[:-4]}"] = read_bmp(os.path.join(base, res, fn))
            n += 1
    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT}: {n} goldens, {os.path.getsize(OUT)} bytes")
    for k in sorted(out):
        print("  ", k, out[k].shape)


if __name__ == "__main__":
    main()
