"""Host-side multi-GPU logic on CPU: world_size-2 (and 3) `gloo` process groups exercise the row-band
partition and the band gather of softray_b200.multi_gpu exactly as bench.py uses them at N > 1
(--gather nccl), with the CPU oracle standing in for each rank's render kernel.  The reference's
analogue is the row-block fan-out of RaytraceGeometry (Renderer.cs:1655-1680)."""
import os
import socket

import numpy as np
import pytest

from softray_b200 import multi_gpu
from tests.util import scenario


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, band_height, res, rows, out_path):
    import torch
    import torch.distributed as dist

    import oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_fixtures.npz"))
        mesh = oracle.load_3ds(fx["model/obj.3ds"].tobytes())
        params = scenario(resolution=res, shadows=False)
        if rows is not None:
            params.start_row, params.end_row = rows
        my_rows = multi_gpu.apply_partition(params, rank, world, band_height)
        sentinel = 0x11223344 + rank
        local = np.full((res, res), sentinel, dtype=np.uint32)
        oracle.Scene([mesh]).render(params, pixels=local)
        # a rank writes exactly its own rows
        mask = np.zeros(res, dtype=bool)
        mask[my_rows] = True
        assert (local[~mask] == sentinel).all()
        assert (local[mask] != sentinel).all()
        t = torch.from_numpy(local.view(np.int32).copy())
        full = multi_gpu.gather_frame(t, my_rows, res, world, band_height, start_row=params.start_row,
                                      end_row=params.end_row)
        if rank == 0:
            np.save(out_path, full.numpy().view(np.uint32))
        else:
            assert full is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,band_height,rows", [(2, 8, None), (2, 5, (7, 40)), (3, 4, None)])
def test_band_gather_reassembles_the_frame(tmp_path, obj_mesh, world, band_height, rows):
    import torch.multiprocessing as mp

    import oracle

    res = 48
    out = str(tmp_path / "frame.npy")
    mp.spawn(_worker, args=(world, _free_port(), band_height, res, rows, out), nprocs=world, join=True)
    got = np.load(out)
    params = scenario(resolution=res, shadows=False)
    if rows is not None:
        params.start_row, params.end_row = rows
    want = np.full((res, res), 0x11223344, dtype=np.uint32)    # rank 0's sentinel outside the rendered rows
    oracle.Scene([obj_mesh]).render(params, pixels=want)
    assert np.array_equal(got, want)


def test_rows_of_rank_partitions_every_row_once():
    for height, world, bh, s, e in [(100, 4, 8, None, None), (1080, 8, 8, None, None), (37, 3, 5, 4, 30), (16, 8, 4, None, None),
                                    (10, 2, 1, 9, 3)]:
        seen = np.concatenate([multi_gpu.rows_of_rank(height, world, bh, r, s, e) for r in range(world)])
        lo, hi = multi_gpu.clamp_rows(height, s, e)
        assert sorted(seen.tolist()) == list(range(lo, hi + 1))


def test_default_band_height_is_a_tile_multiple():
    for h in (100, 1080, 2160, 4320):
        for w in (2, 4, 8):
            bh = multi_gpu.default_band_height(h, w)
            assert bh % 4 == 0 and bh >= 4
    assert multi_gpu.default_band_height(1080, 1) == 0


def _shared_worker(rank, world, port, band_height, res, out_path):
    import torch.distributed as dist

    import oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_fixtures.npz"))
        mesh = oracle.load_3ds(fx["model/obj.3ds"].tobytes())
        params = scenario(resolution=res, shadows=False)
        multi_gpu.apply_partition(params, rank, world, band_height)
        shared = multi_gpu.SharedHostFramebuffer(res, res)          # no Context: not page-locked on CPU
        oracle.Scene([mesh]).render(params, pixels=shared.pixels)   # every rank writes its own bands only
        for _ in range(200):                                        # the library's spin barrier, many generations
            shared.barrier()
        if rank == 0:
            np.save(out_path, shared.pixels.copy())
        shared.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,band_height", [(2, 8), (3, 4)])
def test_shared_host_framebuffer_collects_every_rank(tmp_path, obj_mesh, world, band_height):
    """multi_gpu "host" variant: the ranks' bands land in one shared-memory surface (bench.py's e2e at N > 1)."""
    import torch.multiprocessing as mp

    import oracle

    res = 48
    out = str(tmp_path / "shared.npy")
    mp.spawn(_shared_worker, args=(world, _free_port(), band_height, res, out), nprocs=world, join=True)
    want = oracle.Scene([obj_mesh]).render(scenario(resolution=res, shadows=False))["pixels"]
    assert np.array_equal(np.load(out), want)
