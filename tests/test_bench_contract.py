"""Host-side logic of bench.py that needs no GPU: the `config` object is identical in both arms (the driver compares
them), every named workload generates at a reduced size, the work model is plain arithmetic over the counters, and the
band partition bench.py asks for covers every row exactly once."""
import argparse

import numpy as np

import bench
from softray_b200 import multi_gpu


def _args(**kw):
    a = argparse.Namespace(workload=bench.HEADLINE, others=bench.OTHERS, scale=1.0, gather="peer", band_height=0)
    for k, v in kw.items():
        setattr(a, k, v)
    return a


def test_headline_is_the_multi_gpu_configuration_and_the_4k_frames_ride_along():
    assert bench.HEADLINE == "config5" and bench.SIZES["config5"] == (7680, 4320)
    assert {"config3", "config4"} <= set(bench.OTHERS.split(","))
    assert bench.SIZES["config3"] == bench.SIZES["config4"] == (3840, 2160)
    assert set(bench.SIZES) == set(bench.DESCRIPTIONS)


def test_config_object_is_deterministic_and_holds_nothing_measured():
    for world in (1, 2, 8):
        a, b = bench.config_of(_args(), world), bench.config_of(_args(), world)
        assert a == b
        assert set(a) == {"workload", "description", "width", "height", "scale", "others", "l2", "partition"}
        assert (a["width"], a["height"]) == (7680, 4320) and a["workload"] == "config5"
        assert ("ranks" in a["partition"]) == (world > 1)
    assert bench.config_of(_args(workload="config3", others=""), 1)["others"] == []


def test_every_workload_generates_at_reduced_size():
    for name in bench.SIZES:
        meshes, spheres, frame, desc = bench.workload(name, scale=0.05)
        assert frame.width == int(bench.SIZES[name][0] * 0.05) and frame.height == int(bench.SIZES[name][1] * 0.05)
        assert desc == bench.DESCRIPTIONS[name] and len(meshes) >= 1
        assert (spheres is not None) == name.startswith("config2")
    m5 = bench.workload("config5", 0.01)[0][0]
    assert m5.n_tris == 10_000_000                       # the scene does not shrink with the frame


def test_work_model_is_the_survey_table():
    class F:
        specular_lighting, shading, sub_pixel_res, texture3d_id, width, height = True, True, 1, 0, 10, 10

    st = dict(rays_primary=100, rays_shadow=1000, rays_secondary=0, node_visits=5000, prim_tests=50, sphere_tests=10,
              shaded_hits=40, filter_tests=300, rays_bundled=600)
    f32, f64, nbytes = bench.algorithmic_work(st, F)
    assert f32 == 5000 * 40 + 300 * 11
    assert f64 == 100 * 22 + 40 * 41 + 10 * 34 + 40 * 112 + 400 * 6
    assert nbytes == 5000 * 64 + 40 * 128 + 300 * 64 + 10 * 48 + 100 * 4


def test_default_bands_cover_every_row_once_and_spread_strips():
    for height, world in ((4320, 8), (2160, 8), (2160, 4), (1080, 2), (512, 8), (100, 3)):
        bh = multi_gpu.default_band_height(height, world)
        assert bh % 4 == 0 and bh >= 4
        rows = np.concatenate([multi_gpu.rows_of_rank(height, world, bh, r) for r in range(world)])
        assert sorted(rows.tolist()) == list(range(height))
        # a strip of 32 consecutive rows (config3's grazing shadow rays) is shared by several ranks
        if height >= 1080:
            strip = set(range(height // 3, height // 3 + 32))
            owners = [r for r in range(world) if strip & set(multi_gpu.rows_of_rank(height, world, bh, r).tolist())]
            assert len(owners) >= min(world, 4)
