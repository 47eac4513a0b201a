#!/usr/bin/env python3
"""BASELINE config C4: synthetic 10 k random-sphere scene (all three materials) through the GPU-built BVH, 1920x1080.
Times the BVH build and the production entry (ray generation + material trace + resolve) at --spp samples per pixel."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ascendpathtracing_b200 as pt  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--spp", type=int, default=256)
ap.add_argument("--spheres", type=int, default=10000)
ap.add_argument("--depth", type=int, default=64)
ap.add_argument("--reps", type=int, default=1, help="timed renders; the fastest is reported, all are listed")
a = ap.parse_args()
scene = pt.random_scene(a.spheres, seed=12345)
nsph = 7 + a.spheres
d_scene = torch.from_numpy(scene).cuda()
torch.cuda.synchronize()
t = time.perf_counter()
bvh = pt.Bvh(d_scene, nsph, nsph)
build_ms = (time.perf_counter() - t) * 1e3
t = time.perf_counter()
bvh2 = pt.Bvh(d_scene, nsph, nsph)
build_ms2 = (time.perf_counter() - t) * 1e3
bvh2.close()
p = pt.default_params(width=a.width, height=a.height, samples=a.spp // 4)
mp = pt.default_material_params(seed=1, max_depth=a.depth)
d_img = torch.zeros((a.height, a.width, 3), dtype=torch.uint8, device="cuda")
d_stats = torch.zeros(2, dtype=torch.int64, device="cuda")
small = pt.default_params(width=a.width, height=a.height, samples=1)
pt.render_image_mat_bvh(small, mp, bvh, d_img, cam_seed=3)  # warm-up
torch.cuda.synchronize()
all_ms = []
for _ in range(max(1, a.reps)):
    d_stats.zero_()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    pt.render_image_mat_bvh(p, mp, bvh, d_img, cam_seed=3, stats=d_stats, gamma=True)
    t1.record()
    torch.cuda.synchronize()
    all_ms.append(t0.elapsed_time(t1))
ms = min(all_ms)
n, segs = int(d_stats[0]), int(d_stats[1])
out = {"config": "c4", "width": a.width, "height": a.height, "spp": a.spp, "spheres": nsph, "bvh": bvh.info(), "bvh_build_ms_first": build_ms,
       "bvh_build_ms": build_ms2, "render_ms": ms, "render_ms_all": all_ms, "paths": n, "mpaths_s": n / ms / 1e3, "segments": segs, "segments_per_path": segs / n,
       "grays_s": segs / ms / 1e6, "image_mean": float(d_img.float().mean())}
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/c4.json", "w"), indent=1)
try:
    pt.write_ppm("gpurun_out/c4.ppm", d_img.cpu().numpy()) if a.width * a.height <= 640 * 480 else None
except Exception:
    pass
