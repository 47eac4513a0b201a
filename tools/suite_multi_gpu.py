#!/usr/bin/env python3
"""BASELINE configs C3, C4 and C5 in one launch, on 1 GPU or across all ranks of a torchrun (one process group, so the
start-up cost of N processes is paid once).  Every frame is split into column stripes per rank, rendered through the
production entries (device ray generation -> trace -> resolve) and assembled by one NCCL all-gather of the 8-bit stripes;
time = wall clock around render + gather, max over ranks, best of --reps after one warm-up.

    python tools/suite_multi_gpu.py                                        # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/suite_multi_gpu.py
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import ascendpathtracing_b200 as pt  # noqa: E402
from ascendpathtracing_b200 import sharding  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--scale", type=float, default=1.0, help="scales the samples per pixel of every config (quick runs)")
ap.add_argument("--only", default="", help="comma-separated subset of: c3,c4,c5,c5mat")
ap.add_argument("--parts", type=int, default=1, help="column stripes per rank, dealt round-robin (sharding.interleaved_stripes)")
ap.add_argument("--strided", action="store_true", help="rank r renders columns r, r+G, ... in ONE launch (PtParams.column_step)")
ap.add_argument("--out", default="gpurun_out/suite.json")
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
only = set(x for x in a.only.split(",") if x)


def spp_of(full):
    return max(4, int(full * a.scale) // 4 * 4)


def run_strided(name, W, H, spp, make_render):
    """make_render(x0, x1, d_img, d_stats, step) -> callable rendering columns x0, x0+step, ... < x1 into the dense d_img."""
    x0, step, ncols = sharding.strided_columns(W, rank, world)
    d_img = torch.zeros((H, ncols, 3), dtype=torch.uint8, device="cuda")
    d_stats = torch.zeros(2, dtype=torch.int64, device="cuda")
    render = make_render(x0, W, d_img, d_stats, step)

    def frame():
        render()
        return sharding.gather_strided(d_img, W) if world > 1 else d_img

    return measure(name, W, H, spp, frame, d_stats, "strided")


def measure(name, W, H, spp, frame, d_stats, parts):
    frame()
    torch.cuda.synchronize()
    times = []
    for _ in range(a.reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        d_stats.zero_()
        t = time.perf_counter()
        img = frame()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        times.append(float(dt[0]))
    segs = d_stats.clone()
    if world > 1:
        dist.all_reduce(segs)
    n = W * H * spp
    best = min(times)
    row = {"config": name, "frame": f"{W}x{H}", "spp": spp, "paths": n, "n_gpus": world, "stripes_per_rank": parts, "seconds": best, "mpaths_s": n / best / 1e6,
           "segments": int(segs[1]), "grays_s": int(segs[1]) / best / 1e9, "all_times": times, "image_mean": float(img.float().mean()),
           "image_sum": int(img.long().sum())}
    if rank == 0:
        print(json.dumps(row), flush=True)
    return row


def run(name, W, H, spp, make_render):
    """make_render(x0, x1, d_img, d_stats, step) -> callable rendering this rank's columns."""
    if a.strided and world > 1:
        return run_strided(name, W, H, spp, make_render)
    parts = a.parts if world > 1 else 1
    pieces = [(x0, x1) for x0, x1 in sharding.interleaved_stripes(W, rank, world, parts) if x1 > x0]
    d_img = torch.zeros((H, sum(x1 - x0 for x0, x1 in pieces), 3), dtype=torch.uint8, device="cuda")
    d_stats = torch.zeros(2, dtype=torch.int64, device="cuda")
    d_piece_stats = torch.zeros(2, dtype=torch.int64, device="cuda")
    # each piece renders into its own contiguous [H, w, 3] buffer (the C ABI writes a dense image), then lands in its columns
    bufs = [torch.zeros((H, x1 - x0, 3), dtype=torch.uint8, device="cuda") for x0, x1 in pieces]
    renders = [make_render(x0, x1, buf, d_piece_stats, 1) for (x0, x1), buf in zip(pieces, bufs)]

    def frame():
        off = 0
        for (x0, x1), buf, render in zip(pieces, bufs, renders):
            render()
            d_stats.add_(d_piece_stats)
            d_img[:, off:off + x1 - x0] = buf
            off += x1 - x0
        return sharding.gather_interleaved(d_img, W, parts) if world > 1 else d_img

    return measure(name, W, H, spp, frame, d_stats, parts)


rows = []
d_cornell = torch.from_numpy(pt.default_scene()).cuda()
d_smallpt = torch.from_numpy(pt.smallpt_scene()).cuda()

if not only or "c3" in only:
    W, H, spp = 3840, 2160, spp_of(1024)
    rows.append(run("c3 reference-parity depth 5", W, H, spp,
                    lambda x0, x1, img, st, step: (lambda p=pt.default_params(width=W, height=H, samples=spp // 4, depth=5, column_step=step):
                                                   pt.render_image(p, d_cornell, img, x0=x0, x1=x1, seed=2024, stats=st))))

if not only or "c4" in only:
    W, H, spp = 1920, 1080, spp_of(256)
    nsph = 7 + 10000
    d_c4 = torch.from_numpy(pt.random_scene(10000, seed=12345)).cuda()
    bvh = pt.Bvh(d_c4, nsph, nsph)  # every rank builds its own copy of the tree (4 ms)
    mp4 = pt.default_material_params(seed=1, max_depth=64)
    rows.append(run("c4 10k spheres, materials, BVH, depth cap 64", W, H, spp,
                    lambda x0, x1, img, st, step: (lambda p4=pt.default_params(width=W, height=H, samples=spp // 4, column_step=step):
                                                   pt.render_image_mat_bvh(p4, mp4, bvh, img, x0=x0, x1=x1, cam_seed=2024, gamma=True, stats=st))))

if not only or "c5" in only:
    W, H, spp = 1920, 1080, spp_of(512)
    for depth in (5, 10, 50):
        rows.append(run(f"c5 reference-parity depth {depth}", W, H, spp,
                        lambda x0, x1, img, st, step, depth=depth: (
                            lambda p5=pt.default_params(width=W, height=H, samples=spp // 4, depth=depth, column_step=step):
                            pt.render_image(p5, d_cornell, img, x0=x0, x1=x1, seed=2024, stats=st))))

if not only or "c5mat" in only:
    W, H, spp = 1920, 1080, spp_of(512)
    for depth in (5, 10, 50):
        mpm = pt.default_material_params(seed=1, max_depth=depth)
        rows.append(run(f"c5 materials (DIFF/SPEC/REFR + RR) depth cap {depth}", W, H, spp,
                        lambda x0, x1, img, st, step, mpm=mpm: (
                            lambda pm=pt.default_params(width=W, height=H, samples=spp // 4, sphere_count=9, sphere_stride=16, column_step=step):
                            pt.render_image_mat(pm, mpm, d_smallpt, img, x0=x0, x1=x1, cam_seed=2024, gamma=True, stats=st))))

if rank == 0 and a.out:
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump({"n_gpus": world, "rows": rows}, open(a.out, "w"), indent=1)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
