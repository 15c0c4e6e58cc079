"""CPU tests of the multi-GPU host logic: sample sharding + reduce, with world_size-2 gloo groups.

The stand-in scene evaluates samples with the float64 oracle (test infrastructure) so that the
sharded frame can be compared with the single-process one; on the GPU box the same
``parallel.render_frame`` drives ``backend.NativeScene`` over NCCL (bench.py --gpus N)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parent.parent


def test_sample_range_partitions_every_spp():
    from sightpy.parallel import sample_range
    for spp in (0, 1, 3, 7, 8, 256, 257):
        for world in (1, 2, 3, 4, 8):
            ranges = [sample_range(spp, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == spp
            for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
                assert a1 == b0 and a0 <= a1 and b0 <= b1
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1


def test_interleaved_tiles_partition_the_frame():
    from sightpy.parallel import tile_ids
    for (w, h) in ((1920, 1080), (100, 100), (64, 64), (65, 1)):
        n = -(-w // 64) * -(-h // 64)
        for world in (1, 2, 3, 8):
            parts = [tile_ids(w, h, r, world) for r in range(world)]
            assert sorted(np.concatenate(parts).tolist()) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, spp, shard, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    for p in (REPO, REPO / "python-raytracer_b200", REPO / "tests"):
        sys.path.insert(0, str(p))
    import torch
    import torch.distributed as dist
    import scenes
    import sightpy
    from oracle.sightpy_oracle import Oracle, tonemap_u8
    from sightpy.flatten import flatten_scene
    from sightpy.parallel import render_frame

    dist.init_process_group("gloo", rank=rank, world_size=world)

    class OracleScene:
        """Same surface as backend.NativeScene, evaluated on the CPU by the oracle."""
        width, height = 12, 10

        def __init__(self):
            self.flat = flatten_scene(scenes.cornell(sightpy, width=self.width, height=self.height))
            self.acc = torch.zeros(self.width * self.height * 4, dtype=torch.float32)

        def use_current_stream(self):
            pass

        def render_samples(self, begin, end, seed=0, clear=True):
            return self.render_region(0, self.width * self.height, begin, end, seed, clear)

        def render_region(self, pix_begin, pix_end, begin, end, seed=0, clear=True):
            if clear:
                self.acc.zero_()
            orc = Oracle(self.flat, rng="philox", seed=seed)
            view = self.acc.view(-1, 4)
            pixels = np.arange(pix_begin, pix_end, dtype=np.uint32)
            for s in range(begin, end):
                if len(pixels) == 0:
                    break
                O, D, pix = orc.camera_rays(s, pixels=pixels)
                rgb = orc.trace_columns(O, D, pix, sample=s)["rgb"]
                view[pix_begin:pix_end, :3] += torch.from_numpy(rgb.astype(np.float32))
            return dict(rays_total=orc.rays_total)

        def render_tiles(self, tile_ids, tile_size, begin, end, seed=0, clear=True):
            if clear:
                self.acc.zero_()
            tiles_x = -(-self.width // tile_size)
            total = 0
            for t in tile_ids:                       # a tile = rows of contiguous pixels clipped to the frame
                ty, tx = divmod(int(t), tiles_x)
                for y in range(ty * tile_size, min((ty + 1) * tile_size, self.height)):
                    x0, x1 = tx * tile_size, min((tx + 1) * tile_size, self.width)
                    total += self.render_region(y * self.width + x0, y * self.width + x1, begin, end, seed, clear=False)["rays_total"]
            return dict(rays_total=total)

        def accum_tensor(self):
            return self.acc

        def resolve(self, spp_total, want_linear=True):
            lin = (self.acc.view(-1, 4)[:, :3].numpy().astype(np.float64) / spp_total).T
            return tonemap_u8(lin, self.height, self.width), lin.astype(np.float32).reshape(3, self.height, self.width)

    import sightpy.parallel as par
    par.TILE = 4                                     # a 12 x 10 frame has 3 x 3 tiles of 4 x 4 pixels
    scene = OracleScene()
    srgb, lin, stats = render_frame(scene, spp, seed=4, want_linear=True, shard=shard)
    np.savez(Path(out_dir) / f"rank{rank}.npz", srgb=srgb, lin=lin, rays=stats["rays_total"])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
# auto: 3 spp -> sample-range shards, 1 spp (< world size) -> interleaved tiles; tiles can also be asked for
@pytest.mark.parametrize("spp,shard", [(3, "auto"), (1, "auto"), (3, "tiles")])
def test_two_rank_gloo_frame_equals_single_process(tmp_path, spp, shard):
    import torch.multiprocessing as mp
    world = 2
    mp.start_processes(_worker, args=(world, _free_port(), spp, shard, str(tmp_path)), nprocs=world, join=True,
                       start_method="spawn")
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")

    import scenes
    import sightpy
    from oracle.sightpy_oracle import Oracle, tonemap_u8
    from sightpy.flatten import flatten_scene
    flat = flatten_scene(scenes.cornell(sightpy, width=12, height=10))
    orc = Oracle(flat, rng="philox", seed=4)
    full = orc.render_linear(spp)
    np.testing.assert_allclose(r0["lin"].reshape(3, -1), full, rtol=1e-5, atol=1e-6)   # rank 0 holds the frame
    assert np.array_equal(r0["srgb"], tonemap_u8(r0["lin"].reshape(3, -1).astype(np.float64), 10, 12))
    assert int(r0["rays"]) + int(r1["rays"]) == orc.rays_total                            # shards partition the work
    assert not np.allclose(r1["lin"], r0["lin"])                                          # rank 1 only has its shard
