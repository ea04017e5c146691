"""Host logic of the N > 1 path on CPU: world size 2, gloo backend.

The pixels themselves need a GPU (tests/test_gpu_render.py::test_sharded_render_assembles_to_the_unsharded_frame
renders the shards on one device); what runs here is everything around them -- which tiles a rank owns
(host mirror of ShardMap::pixel_of, checked against the C ABI's rt_shard_float4_count), the single
rank-0 exchange with unequal shard sizes (the same function the NCCL path calls), and the de-interleave
(numpy restatement of k_assemble)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rs_pathtracing_b200 import api
from rs_pathtracing_b200.distributed import assemble_host, exchange_to_rank0, owned_pixel_coords


def _pixel_value(x, y):
    return np.stack([x * 1.0 + 0.25, y * 2.0 + 0.5, (x * 7 + y * 13) % 101 + 1.0], axis=-1)


def _worker(rank, world, port, width, height, tile, spp, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x, y, ok = owned_pixel_coords(width, height, tile, world, rank)
        p = api.render_params(width, height, spp, 8, 0, world, rank, tile=tile)
        counts = [api.shard_float4_count(p, s) for s in range(world)]
        assert len(x) == counts[rank]
        mine = np.zeros((counts[rank], 4), dtype=np.float64)   # (sum rgb, samples); padding stays 0
        mine[ok, :3] = (_pixel_value(x[ok], y[ok]) * spp).astype(np.float64)
        mine[ok, 3] = spp
        shards = exchange_to_rank0(dist, torch.from_numpy(mine), counts, rank, world)
        if rank == 0:
            frame = assemble_host([t.numpy() for t in shards], width, height, tile, world)
            np.save(out_path, frame)
        else:
            assert shards is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("width,height,tile", [(96, 64, 32), (100, 70, 32), (33, 17, 8)])
def test_two_rank_gather_and_assemble(tmp_path, width, height, tile):
    world, spp = 2, 4
    out = str(tmp_path / "frame.npy")
    mp.spawn(_worker, args=(world, _free_port(), width, height, tile, spp, out), nprocs=world, join=True)
    frame = np.load(out)
    yy, xx = np.mgrid[0:height, 0:width]
    want = _pixel_value(xx, yy).astype(np.float64).astype(np.float64)   # f64 accumulator
    assert np.array_equal(frame, want)


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_owned_pixels_partition_the_image(world):
    width, height, tile = 70, 45, 16
    seen = np.zeros((height, width), dtype=np.int32)
    for s in range(world):
        x, y, ok = owned_pixel_coords(width, height, tile, world, s)
        p = api.render_params(width, height, 1, 8, 0, world, s, tile=tile)
        assert len(x) == api.shard_float4_count(p, s)
        np.add.at(seen, (y[ok], x[ok]), 1)
    assert (seen == 1).all()


def _worker_sequence(rank, world, port, width, height, tile, out_path):
    """consecutive frames with DIFFERENT contents through the same (reused) gather buffers, no barrier between"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x, y, ok = owned_pixel_coords(width, height, tile, world, rank)
        p = api.render_params(width, height, 1, 8, 0, world, rank, tile=tile)
        counts = [api.shard_float4_count(p, s) for s in range(world)]
        bufs, frames = None, []
        mine = torch.zeros((counts[rank], 4), dtype=torch.float64)    # one accumulator, overwritten every frame
        for k, spp in enumerate((4, 1, 16, 2)):
            acc = np.zeros((counts[rank], 4), dtype=np.float64)
            acc[ok, :3] = ((_pixel_value(x[ok], y[ok]) + k) * spp).astype(np.float64)
            acc[ok, 3] = spp
            mine.copy_(torch.from_numpy(acc))
            shards = exchange_to_rank0(dist, mine, counts, rank, world, bufs)
            if rank == 0:
                bufs = shards
                frames.append(assemble_host([t.numpy() for t in shards], width, height, tile, world))
        if rank == 0:
            np.save(out_path, np.stack(frames))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_consecutive_frames_through_reused_gather_buffers(tmp_path):
    world, width, height, tile = 2, 100, 70, 32
    out = str(tmp_path / "frames.npy")
    mp.spawn(_worker_sequence, args=(world, _free_port(), width, height, tile, out), nprocs=world, join=True)
    frames = np.load(out)
    yy, xx = np.mgrid[0:height, 0:width]
    for k in range(4):
        want = (_pixel_value(xx, yy) + k).astype(np.float64).astype(np.float64)
        assert np.array_equal(frames[k], want), k


@pytest.mark.parametrize("world", [2, 4, 8])
def test_cfg5_frame_partition_and_balance(world):
    """BASELINE configs[4]: 3840x2160 in 32x32 tiles (120 x 67.5: the last tile row is clipped, and 120 is a multiple
    of every shard count, so the row rotation is what keeps a shard from owning whole tile columns)"""
    width, height, tile = 3840, 2160, 32
    seen = np.zeros((height, width), dtype=np.uint8)
    valid = []
    for s in range(world):
        x, y, ok = owned_pixel_coords(width, height, tile, world, s)
        p = api.render_params(width, height, 1, 8, 0, world, s, tile=tile)
        assert len(x) == api.shard_float4_count(p, s)
        np.add.at(seen, (y[ok], x[ok]), 1)
        valid.append(int(ok.sum()))
        # a shard's tiles spread over (nearly) all tile columns, not width / world of them
        assert len(np.unique(x[ok] // tile)) >= 120 - 1
    assert (seen == 1).all()
    assert max(valid) - min(valid) <= 120 * tile * tile // world   # within one tile row's share of each other
