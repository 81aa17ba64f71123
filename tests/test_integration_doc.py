"""INTEGRATION.md's C# binding against the header: the structs a maintainer would paste into Engine3D must have the
sizes and field offsets of include/softray_cuda.h (whose ctypes mirror, softray_b200/abi.py, is itself checked
against the compiled library by tests/test_abi.py), and the DllImport list must name every exported entry point."""
import ctypes as C
import os
import re

from softray_b200 import abi, lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CS_TYPES = {"double": (8, 8), "int": (4, 4), "uint": (4, 4), "ulong": (8, 8), "long": (8, 8), "float": (4, 4), "byte": (1, 1)}
MIRROR = {"softray_mesh": abi.Mesh, "softray_sphere": abi.Sphere, "softray_scene_desc": abi.SceneDesc,
          "softray_instance": abi.Instance, "softray_frame": abi.Frame, "softray_stats": abi.Stats}


def csharp_block():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    return re.search(r"```csharp\nusing System;(.*?)```", text, re.S).group(1)


def layout_of(body):
    """LayoutKind.Sequential: declaration order, natural alignment, size rounded up to the largest alignment."""
    off, max_align, fields = 0, 1, {}
    for decl in re.findall(r"public\s+(.*?);", body, re.S):
        decl = " ".join(decl.split())
        m = re.match(r"fixed (\w+) (\w+)\[(\d+)\]$", decl)
        if m:
            size, align = CS_TYPES[m.group(1)]
            entries = [(m.group(2), size * int(m.group(3)), align)]
        else:
            ty, names = decl.split(" ", 1)
            size, align = (8, 8) if ty.endswith("*") else CS_TYPES[ty]
            entries = [(n.strip(), size, align) for n in names.split(",")]
        for name, size, align in entries:
            off = (off + align - 1) // align * align
            fields[name] = off
            off += size
            max_align = max(max_align, align)
    return (off + max_align - 1) // max_align * max_align, fields


def test_csharp_structs_match_the_header():
    block = csharp_block()
    seen = set()
    for m in re.finditer(r"struct (softray_\w+)\s*\{(.*?)\}\s*\n", block, re.S):
        name, body = m.group(1), m.group(2)
        size, fields = layout_of(body)
        mirror = MIRROR[name]
        assert size == C.sizeof(mirror), f"{name}: C# {size} bytes, header {C.sizeof(mirror)}"
        want = {f[0]: getattr(mirror, f[0]).offset for f in mirror._fields_}
        assert fields == want, f"{name}: {set(fields.items()) ^ set(want.items())}"
        seen.add(name)
    assert seen == set(MIRROR)


def test_csharp_binding_names_every_entry_point():
    block = csharp_block()
    imported = set(re.findall(r"extern \w[\w\.\*]* (softray_\w+)\(", block))
    assert imported == set(lib.EXPORTS), set(lib.EXPORTS) ^ imported
    assert f"softray_abi_version() != {abi.ABI_VERSION}" in block
