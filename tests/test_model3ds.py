"""The native .3DS -> flattened mesh path of the library (softray_model_load_3ds) against
the oracle's restatement of Model.Load3ds / PostProcessGeometry, on the reference's own test models.
Host code only: runs without a GPU."""
import struct

import numpy as np
import pytest

from softray_b200 import abi


@pytest.fixture(scope="module")
def L():
    from softray_b200 import lib

    lib.build()
    return lib


@pytest.mark.parametrize("name,n_verts,n_tris", [("obj.3ds", 269, 152), ("obj2.3ds", 112, 107)])
def test_native_loader_matches_the_oracle_bit_for_bit(L, fixtures, name, n_verts, n_tris):
    import oracle

    data = fixtures["model/" + name].tobytes()
    got = L.load_3ds(data)
    want = oracle.load_3ds(data)
    assert got.n_verts == want.n_verts == n_verts and got.n_tris == want.n_tris == n_tris
    assert np.array_equal(got.verts.view(np.uint64), want.verts.view(np.uint64))
    assert np.array_equal(got.tris, want.tris)
    assert np.array_equal(got.argb, want.argb)
    assert np.array_equal(got.bbox_min.view(np.uint64), want.bbox_min.view(np.uint64))
    assert np.array_equal(got.bbox_max.view(np.uint64), want.bbox_max.view(np.uint64))
    # Model.PostProcessGeometry: the longest axis spans exactly [-0.5, 0.5] (Model.cs:762-790)
    assert (got.bbox_max - got.bbox_min).max() == 1.0 and np.allclose(got.bbox_min + got.bbox_max, 0.0, atol=1e-15)


def test_material_colour_is_truncated_not_rounded(L, fixtures):
    got = L.load_3ds(fixtures["model/obj.3ds"].tobytes())
    assert set(got.argb.tolist()) == {0xFF969696}      # diffuse 0.5882353 * 255 = 149.99... -> 0x96


def test_malformed_streams_are_format_errors(L, fixtures):
    data = fixtures["model/obj.3ds"].tobytes()
    for bad in (b"", b"\x00" * 16, data[:100], data[: len(data) // 2], b"XX" + data[2:],
                struct.pack("<HI", 0x4D4D, 3), struct.pack("<HI", 0x4D4D, 6)):
        with pytest.raises(L.SoftRayError) as e:
            L.load_3ds(bad)
        assert e.value.code == abi.E_FORMAT
    with pytest.raises(L.SoftRayError) as e:
        L.load()  # make sure the library is loaded before poking the raw entry point
        import ctypes as C
        h = C.c_void_p()
        rc = L.load().softray_model_load_3ds(None, 0, C.byref(h))
        if rc != abi.OK:
            raise L.SoftRayError(rc, "null")
    assert e.value.code == abi.E_INVALID_ARG


def test_loader_agrees_with_oracle_on_truncations(L, fixtures):
    """Every prefix either fails in both or yields the same mesh in both."""
    import oracle

    data = fixtures["model/obj2.3ds"].tobytes()
    for cut in range(0, len(data), 97):
        blob = data[:cut]
        try:
            want = oracle.load_3ds(blob)
        except Exception:
            want = None
        try:
            got = L.load_3ds(blob)
        except L.SoftRayError as e:
            assert e.code == abi.E_FORMAT
            got = None
        assert (got is None) == (want is None), cut
        if got is not None:
            assert np.array_equal(got.verts.view(np.uint64), want.verts.view(np.uint64)) and np.array_equal(got.tris, want.tris)
