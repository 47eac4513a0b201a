"""CPU, world_size 2 over gloo: the multi-GPU host logic (stripe partition, slice arithmetic, gather + assembly).
The per-stripe pixels come from the oracle here (no GPU): what is under test is the sharding, not the kernel."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, w, h, s, outdir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist

    from ascendpathtracing_b200 import sharding
    from oracle import oracle as O
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x0, x1 = sharding.stripe(w, rank, world)
        first, count = sharding.path_slice(w, h, s, rank, world)
        n = w * h * 4 * s
        # this rank generates and traces ONLY its slice, with the counter-based stream (global path indices)
        rays = O.gen_rays_from_uniforms(w, h, s, x0, x1, O.philox_uniforms(5, first, count))
        col = O.trace(rays, O.gen_spheres())
        # resolve the stripe as its own (x1-x0)-column image: pixels are whole inside a stripe
        img = O.resolve(col, x1 - x0, h, s) if x1 > x0 else np.zeros((h, 0, 3), np.uint8)
        frame = sharding.gather_stripes(torch.from_numpy(img), w)
        if rank == 0:
            np.save(os.path.join(outdir, "frame.npy"), frame.numpy())
        assert first + count <= n
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("w,h,s", [(10, 6, 2), (7, 5, 1)])
def test_two_rank_stripes_assemble_to_the_single_rank_frame(tmp_path, oracle, w, h, s):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, w, h, s, str(tmp_path)), nprocs=2, join=True)
    frame = np.load(tmp_path / "frame.npy")
    n = w * h * 4 * s
    rays = oracle.gen_rays_from_uniforms(w, h, s, 0, w, oracle.philox_uniforms(5, 0, n))
    whole = oracle.resolve(oracle.trace(rays, oracle.gen_spheres()), w, h, s)
    assert np.array_equal(frame, whole)


def _worker_interleaved(rank, world, port, w, h, parts, outdir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist

    from ascendpathtracing_b200 import sharding
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # pixel (y, x) carries its own coordinates, so the assembled frame is checkable without rendering anything
        pieces = sharding.interleaved_stripes(w, rank, world, parts)
        cols = [x for a, b in pieces for x in range(a, b)]
        local = np.zeros((h, len(cols), 3), np.uint8)
        for k, x in enumerate(cols):
            local[:, k, 0] = x % 251
            local[:, k, 1] = np.arange(h) % 251
            local[:, k, 2] = rank
        frame = sharding.gather_interleaved(torch.from_numpy(local), w, parts)
        if rank == 0:
            np.save(os.path.join(outdir, "frame.npy"), frame.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("w,h,parts", [(37, 4, 3), (5, 3, 4), (64, 2, 1)])
def test_two_rank_interleaved_stripes_assemble(tmp_path, w, h, parts):
    import torch.multiprocessing as mp

    from ascendpathtracing_b200 import sharding
    port = _free_port()
    mp.spawn(_worker_interleaved, args=(2, port, w, h, parts, str(tmp_path)), nprocs=2, join=True)
    frame = np.load(tmp_path / "frame.npy")
    assert np.array_equal(frame[:, :, 0], np.broadcast_to(np.arange(w) % 251, (h, w)))
    assert np.array_equal(frame[:, :, 1], np.broadcast_to((np.arange(h) % 251)[:, None], (h, w)))
    owner = np.empty(w, np.uint8)
    for r in range(2):
        for a, b in sharding.interleaved_stripes(w, r, 2, parts):
            owner[a:b] = r
    assert np.array_equal(frame[0, :, 2], owner)


def _worker_strided(rank, world, port, w, h, outdir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist

    from ascendpathtracing_b200 import sharding
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x0, step, ncols = sharding.strided_columns(w, rank, world)
        local = np.zeros((h, ncols, 3), np.uint8)
        for j in range(ncols):
            local[:, j, 0] = (x0 + j * step) % 251
            local[:, j, 1] = np.arange(h) % 251
            local[:, j, 2] = rank
        frame = sharding.gather_strided(torch.from_numpy(local), w)
        if rank == 0:
            np.save(os.path.join(outdir, "frame.npy"), frame.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("w,h", [(37, 4), (2, 3), (64, 2)])
def test_two_rank_strided_columns_assemble(tmp_path, w, h):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker_strided, args=(2, port, w, h, str(tmp_path)), nprocs=2, join=True)
    frame = np.load(tmp_path / "frame.npy")
    assert np.array_equal(frame[:, :, 0], np.broadcast_to(np.arange(w) % 251, (h, w)))
    assert np.array_equal(frame[:, :, 1], np.broadcast_to((np.arange(h) % 251)[:, None], (h, w)))
    assert np.array_equal(frame[0, :, 2], np.arange(w) % 2)


def test_strided_columns_cover_the_frame():
    from ascendpathtracing_b200 import sharding
    for width in (1, 5, 64, 1920):
        for world in (1, 2, 3, 8):
            seen = np.zeros(width, int)
            for r in range(world):
                x0, step, n = sharding.strided_columns(width, r, world)
                cols = x0 + step * np.arange(n)
                assert (cols < width).all()
                seen[cols] += 1
            assert (seen == 1).all()


def test_interleaved_partition_properties():
    from ascendpathtracing_b200 import sharding
    for width in (1, 7, 64, 1920, 3840):
        for world in (1, 2, 8):
            for parts in (1, 2, 4, 7):
                seen = np.zeros(width, int)
                for r in range(world):
                    pieces = sharding.interleaved_stripes(width, r, world, parts)
                    assert len(pieces) == parts
                    assert all(a[1] <= b[0] for a, b in zip(pieces, pieces[1:]))
                    for a, b in pieces:
                        seen[a:b] += 1
                assert (seen == 1).all()
    assert sharding.interleaved_stripes(100, 1, 4, 1) == [sharding.stripe(100, 1, 4)]
    with pytest.raises(ValueError):
        sharding.interleaved_stripes(8, 0, 2, 0)


def test_stripe_partition_properties():
    from ascendpathtracing_b200 import sharding
    for width in (1, 7, 8, 1024, 3840):
        for world in (1, 2, 3, 4, 8):
            cols = [sharding.stripe(width, r, world) for r in range(world)]
            assert cols[0][0] == 0 and cols[-1][1] == width
            assert all(a[1] == b[0] for a, b in zip(cols, cols[1:]))
            sizes = [b - a for a, b in cols]
            assert max(sizes) - min(sizes) <= 1
            slices = [sharding.path_slice(width, 5, 3, r, world) for r in range(world)]
            assert sum(c for _, c in slices) == width * 5 * 4 * 3
    with pytest.raises(ValueError):
        sharding.stripe(8, 2, 2)
