"""Row-band partition of one frame across ranks (one process per GPU) and the two ways the finished
bands reach rank 0's framebuffer.

The reference splits rows into blocks across 4 tasks that all store into the one shared
surface.Pixels (Renderer.cs:1655-1680) and exposes rayTraceStartRow/EndRow for caller-driven bands
(Renderer.cs:134-136).  Here a rank renders the bands  b % n_ranks == rank  of  band_height  rows
(softray_frame.band_*; interleaved so that cheap background rows spread over all ranks).

  * peer  (default): rank 0 owns the framebuffer, exports it with softray_ipc_export, every other
    rank maps it over NVLink and renders straight into it -- the render kernel's own coalesced
    uchar4 stores are the gather, overlapped with tracing tile by tile.
  * nccl : every rank renders into a local frame, packs its rows and torch.distributed.gather
    moves them to rank 0 (NCCL send/recv over NVLink; gloo on CPU in the tests).

  * host  (the host-buffer entry point, softray_render): the caller's surface lives in a shared-memory section
    every rank process maps and page-locks (softray_host_register); each rank's kernel stores its bands straight
    into it over its own GPU's PCIe link -- no device framebuffer, no NVLink hop, no D2H copy on rank 0.

Only host-side logic lives here; torch.distributed is plumbing.
"""
import numpy as np


def clamp_rows(height, start_row=None, end_row=None):
    """Renderer.cs:1652-1653."""
    s = 0 if start_row is None else min(max(int(start_row), 0), height - 1)
    e = height - 1 if end_row is None else min(max(int(end_row), 0), height - 1)
    return s, e


def rows_of_rank(height, n_ranks, band_height, rank, start_row=None, end_row=None):
    """Sorted row indices rank `rank` renders (same rule as the kernel and the oracle)."""
    s, e = clamp_rows(height, start_row, end_row)
    rows = np.arange(s, e + 1, dtype=np.int64)
    if n_ranks <= 1 or band_height <= 0:
        return rows
    return rows[((rows - s) // band_height) % n_ranks == rank]


def default_band_height(height, n_ranks, target_bands_per_rank=64, tile_rows=4):
    """A multiple of the kernel's 4-row tile that gives every rank ~64 interleaved bands: expensive rows come in
    strips (config3's grazing shadow rays sit in a few dozen rows near the horizon), and a strip must not fall to one
    or two ranks (16 bands per rank: config3 3.0 ms on 8 GPUs, slower than on 4)."""
    if n_ranks <= 1:
        return 0
    bh = height // (n_ranks * target_bands_per_rank)
    bh = max(tile_rows, (bh // tile_rows) * tile_rows)
    return int(bh)


def apply_partition(params, rank, n_ranks, band_height=None):
    """Set the band fields of a FrameParams for this rank (in place) and return its rows."""
    if band_height is None:
        band_height = default_band_height(params.height, n_ranks)
    params.band_height = band_height if n_ranks > 1 else 0
    params.band_count = max(1, n_ranks)
    params.band_index = rank if n_ranks > 1 else 0
    return rows_of_rank(params.height, n_ranks, band_height, rank, params.start_row, params.end_row)


def gather_frame(local_frame, my_rows, height, n_ranks, band_height, dst=0, group=None, start_row=None, end_row=None):
    """nccl/gloo variant.  local_frame: [H, W] int32/uint32-as-int32 tensor whose rows `my_rows`
    are valid.  Returns the assembled [H, W] tensor on rank dst (rows nobody rendered keep
    local_frame's content there), None elsewhere."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group)
    W = local_frame.shape[1]
    counts = [len(rows_of_rank(height, n_ranks, band_height, r, start_row, end_row)) for r in range(n_ranks)]
    pad = max(counts)
    idx = torch.as_tensor(np.asarray(my_rows), device=local_frame.device, dtype=torch.long)
    pack = torch.zeros((pad, W), dtype=local_frame.dtype, device=local_frame.device)
    if len(my_rows):
        pack[: len(my_rows)] = local_frame.index_select(0, idx)
    if rank == dst:
        parts = [torch.empty_like(pack) for _ in range(n_ranks)]
        dist.gather(pack, parts, dst=dst, group=group)
        out = local_frame.clone()
        for r in range(n_ranks):
            rows = rows_of_rank(height, n_ranks, band_height, r, start_row, end_row)
            if len(rows):
                ridx = torch.as_tensor(rows, device=out.device, dtype=torch.long)
                out.index_copy_(0, ridx, parts[r][: len(rows)])
        return out
    dist.gather(pack, None, dst=dst, group=group)
    return None


class PeerFramebuffer:
    """peer variant: rank 0 allocates width*height uint32 in its HBM (softray_device_alloc) and
    every other rank maps it (softray_ipc_open).  .ptr is the device pointer to pass as
    d_pixels_argb to softray_render_device on THIS rank."""

    def __init__(self, ctx, width, height, group=None):
        import torch.distributed as dist

        self.ctx = ctx
        self.group = group
        self.nbytes = int(width) * int(height) * 4
        self.rank = dist.get_rank(group)
        self.owner = self.rank == 0
        box = [None]
        if self.owner:
            self.ptr = ctx.device_alloc(self.nbytes)
            box[0] = ctx.ipc_export(self.ptr)
        dist.broadcast_object_list(box, src=0, group=group)
        if not self.owner:
            self.ptr = ctx.ipc_open(box[0])
        dist.barrier(group)

    def close(self):
        """Collective: the importers unmap first (cudaIpcCloseMemHandle), every rank meets, only then does the
        owner free -- freeing an exported allocation that a peer still has mapped is undefined behaviour."""
        import torch.distributed as dist

        if getattr(self, "ptr", None):
            if not self.owner:
                self.ctx.ipc_close(self.ptr)
            dist.barrier(self.group)
            if self.owner:
                self.ctx.device_free(self.ptr)
            self.ptr = None


class SharedHostFramebuffer:
    """host variant: one width*height uint32 surface in POSIX shared memory, mapped by every rank process.
    .pixels is the [H, W] uint32 numpy view to pass to Scene.render(..., pixels=...) on THIS rank; with a
    Context it is page-locked in this process (softray_host_register) so the kernel writes it directly."""

    def __init__(self, width, height, ctx=None, group=None):
        from multiprocessing import shared_memory

        import torch.distributed as dist

        self.ctx = ctx
        self.rank = dist.get_rank(group)
        self.owner = self.rank == 0
        nbytes = int(width) * int(height) * 4
        self.world = dist.get_world_size(group)
        box = [None]
        if self.owner:
            self.shm = shared_memory.SharedMemory(create=True, size=nbytes + 64)     # + the barrier words
            box[0] = self.shm.name
        dist.broadcast_object_list(box, src=0, group=group)
        if not self.owner:
            self.shm = shared_memory.SharedMemory(name=box[0])
            try:    # the owner unlinks; keep Python's resource tracker from doing it again at exit
                from multiprocessing import resource_tracker

                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:
                pass
        self.pixels = np.ndarray((int(height), int(width)), dtype=np.uint32, buffer=self.shm.buf)
        self._words = np.ndarray((2,), dtype=np.uint32, buffer=self.shm.buf, offset=nbytes)
        if self.owner:
            self.pixels[:] = 0
            self._words[:] = 0
        self.registered = False
        if ctx is not None:
            ctx.host_register(self.pixels)
            self.registered = True
        dist.barrier(group)
        self.group = group

    def barrier(self):
        """End of frame: every rank's softray_render has returned (softray_host_barrier, a spin barrier in the
        shared section; needs the library)."""
        from . import lib

        rc = lib.load().softray_host_barrier(self._words.ctypes.data, self.world)
        if rc != 0:
            raise lib.SoftRayError(rc, "softray_host_barrier")

    def close(self):
        import torch.distributed as dist

        if getattr(self, "shm", None) is None:
            return
        if self.registered:
            self.ctx.host_unregister(self.pixels)
            self.registered = False
        dist.barrier(self.group)
        self.pixels = None
        self._words = None
        self.shm.close()
        if self.owner:
            self.shm.unlink()
        self.shm = None
