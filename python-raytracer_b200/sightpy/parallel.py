"""Frame-level sharding across the GPUs of one node.

Pixels and samples are independent, so the only exchange is the final image.  Two ways to cut a frame
(``SIGHTPY_SHARD`` = ``samples`` | ``tiles`` | ``auto``, or the ``shard`` argument of ``render_frame``):

* ``samples``: rank r of R renders the contiguous sample range [r*spp/R, (r+1)*spp/R) of the whole frame —
  perfect balance whatever the scene looks like (the Cornell box headline);
* ``tiles``: rank r renders the 64x64-pixel tiles r, r + R, r + 2R, ... (row-major tile ids) for all samples —
  interleaving balances scenes whose cost varies across the frame, and works when spp < R (``auto`` picks it then).

Either way every rank accumulates into its own float4 buffer, the buffers are summed onto rank 0 with one NCCL
``reduce`` over NVLink (33 MB at 1080p; for tiles the sum just interleaves, every other rank's pixels being zero) and
rank 0 tonemaps.  The counter-based RNG is keyed by (pixel, *global* sample index), so the union of the shards is the
same set of light paths as a single-GPU render.  Replaces the multiprocessing.Pool fan-out of the reference
(sightpy/scene.py:78-116).
"""
import os

import numpy as np

__all__ = ["sample_range", "tile_ids", "world", "render_frame", "accum_as_tensor", "TILE"]

TILE = 64


def sample_range(spp, rank, world_size):
    """Contiguous, balanced split of range(spp); ranks beyond spp get an empty range."""
    base, extra = divmod(int(spp), int(world_size))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def tile_ids(width, height, rank, world_size, tile=None):
    """Row-major ids of the tiles rank `rank` renders: every world_size-th tile of the frame."""
    tile = tile or TILE
    n = -(-int(width) // tile) * -(-int(height) // tile)
    return np.arange(rank, n, world_size, dtype=np.int32)


def world():
    """(rank, world_size) of an initialised torch.distributed job, else (0, 1)."""
    if int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        return 0, 1
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(), dist.get_world_size()


class _CudaBlob:
    """Minimal __cuda_array_interface__ carrier so torch can view library-owned device memory."""

    def __init__(self, ptr, nfloats):
        self.__cuda_array_interface__ = {
            "shape": (nfloats,), "typestr": "<f4", "data": (ptr, False), "version": 2, "strides": None}


def accum_as_tensor(native):
    """torch.float32 view (no copy) of the scene's device accumulation buffer, on the device the library is bound to."""
    import torch
    ptr, nbytes = native.accum_pointer()
    device = getattr(native, "device_index", None)
    if device is None:
        device = torch.cuda.current_device()
    return torch.as_tensor(_CudaBlob(ptr, nbytes // 4), device=torch.device("cuda", int(device)))


def render_frame(native, spp, seed=0, want_linear=False, shard=None):
    """Render one frame with every rank of the job.  Returns (uint8 H x W x 3, stats) — on ranks
    other than 0 the image is the rank-local (unreduced) resolve and only rank 0's is the frame.

    ``native`` is a backend.NativeScene, or any object with the same ``render`` / ``render_samples`` /
    ``render_tiles`` / ``accum_tensor`` / ``resolve`` / ``use_current_stream`` methods (the CPU tests drive this
    function with a gloo group and a stand-in scene)."""
    rank, size = world()
    if size == 1:
        # one process: a backend.NativeGroup spreads the frame over the node's GPUs inside the library
        kw = {"shard": shard} if shard and hasattr(native, "members") else {}
        srgb, lin, stats = native.render(spp, seed, want_linear=want_linear, **kw)
        return (srgb, stats) if not want_linear else (srgb, lin, stats)
    import torch.distributed as dist
    shard = shard or os.environ.get("SIGHTPY_SHARD", "auto")
    if shard not in ("auto", "samples", "tiles"):
        raise ValueError(f"unknown shard mode {shard!r} (samples | tiles | auto)")
    if shard == "auto":
        shard = "samples" if spp >= size else "tiles"
    native.use_current_stream()      # same stream as the collective: the reduce is ordered after the last level kernel
    if shard == "samples":
        begin, end = sample_range(spp, rank, size)
        stats = native.render_samples(begin, end, seed, clear=True)
    else:
        stats = native.render_tiles(tile_ids(native.width, native.height, rank, size, TILE), TILE, 0, spp, seed, clear=True)
    acc = native.accum_tensor()
    dist.reduce(acc, dst=0, op=dist.ReduceOp.SUM)
    srgb, lin = native.resolve(spp, want_linear=want_linear)     # synchronises the stream first
    return (srgb, stats) if not want_linear else (srgb, lin, stats)
