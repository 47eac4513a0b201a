#!/usr/bin/env python3
"""Renders one full frame across all ranks (torchrun) or one GPU: column stripes per rank, the production entry
(device ray generation -> trace -> resolve), NCCL all-gather of the 8-bit stripes.  BASELINE config C3 by default
(3840x2160, 1024 spp = 8 493 465 600 paths).  Strong scaling: the frame is fixed, ranks split it.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/render_frame.py
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
ap.add_argument("--width", type=int, default=3840)
ap.add_argument("--height", type=int, default=2160)
ap.add_argument("--spp", type=int, default=1024)
ap.add_argument("--depth", type=int, default=5)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--materials", action="store_true", help="DIFF/SPEC/REFR + Russian roulette on smallpt's scene (C5, extension kernel)")
ap.add_argument("--c4", action="store_true", help="BASELINE config C4: 10 k random spheres, all three materials, through the GPU-built BVH")
ap.add_argument("--max-depth", type=int, default=64, help="bounce cap of the material kernels")
ap.add_argument("--out", default="")
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
os.dup2(2, 1) if False else None
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W, H, S = a.width, a.height, a.spp // 4
x0, x1 = sharding.stripe(W, rank, world)
d_img = torch.zeros((H, x1 - x0, 3), dtype=torch.uint8, device="cuda")
if a.c4:
    nsph = 7 + 10000
    d_sc = torch.from_numpy(pt.random_scene(10000, seed=12345)).cuda()
    bvh = pt.Bvh(d_sc, nsph, nsph)  # every rank builds its own copy of the tree (4 ms)
    p = pt.default_params(width=W, height=H, samples=S)
    mp = pt.default_material_params(seed=1, max_depth=a.max_depth)
    render = lambda: pt.render_image_mat_bvh(p, mp, bvh, d_img, x0=x0, x1=x1, cam_seed=2024, gamma=True)  # noqa: E731
elif a.materials:
    p = pt.default_params(width=W, height=H, samples=S, sphere_count=9, sphere_stride=16)
    mp = pt.default_material_params(seed=1, max_depth=a.max_depth)
    d_sc = torch.from_numpy(pt.smallpt_scene()).cuda()
    render = lambda: pt.render_image_mat(p, mp, d_sc, d_img, x0=x0, x1=x1, cam_seed=2024, gamma=True)  # noqa: E731
else:
    p = pt.default_params(width=W, height=H, samples=S, depth=a.depth)
    d_sc = torch.from_numpy(pt.default_scene()).cuda()
    render = lambda: pt.render_image(p, d_sc, d_img, x0=x0, x1=x1, seed=2024)  # noqa: E731


def frame():
    render()
    return sharding.gather_stripes(d_img, W) if world > 1 else d_img


frame()  # warm-up: workspace arena, NCCL
torch.cuda.synchronize()
times = []
for _ in range(a.reps):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = time.perf_counter()
    img = frame()
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    times.append(float(dt[0]))
if rank == 0:
    n = W * H * 4 * S
    best = min(times)
    out = {"frame": f"{W}x{H}", "spp": 4 * S, "paths": n, "n_gpus": world, "mode": (f"c4 10k spheres + BVH, depth cap {a.max_depth}" if a.c4 else f"materials depth cap {a.max_depth}" if a.materials
                    else f"reference-parity depth {a.depth}"),
           "seconds": best, "mpaths_s": n / best / 1e6, "all_times": times, "image_mean": float(img.float().mean()),
           "image_sha_head": int(img.view(-1)[:4096].long().sum())}
    print(json.dumps(out), flush=True)
    if a.out:
        os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
        json.dump(out, open(a.out, "w"), indent=1)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
