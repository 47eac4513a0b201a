#!/usr/bin/env python3
"""The BVH path beyond config C4's 10 k spheres: build time, render rate and a spot check against brute force for
10^4 .. 10^6 random spheres (same generator as C4, radius scaled with n^(-1/3) so that the box stays equally crowded)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import ascendpathtracing_b200 as pt  # noqa: E402
from oracle import oracle as O  # noqa: E402  (checker only)

W, H, SPP = 1920, 1080, 16
rows = []
for n_random in (10_000, 100_000, 1_000_000):
    scene = pt.random_scene(n_random, seed=12345).reshape(11, -1)
    scene[0, 7:] *= (10_000 / n_random) ** (2.0 / 3.0)   # r^2 scaled: radius ~ n^(-1/3)
    scene = np.ascontiguousarray(scene).reshape(-1)
    nsph = 7 + n_random
    d_scene = torch.from_numpy(scene).cuda()
    torch.cuda.synchronize()
    pt.Bvh(d_scene, nsph, nsph).close()
    t = time.perf_counter()
    bvh = pt.Bvh(d_scene, nsph, nsph)
    build_ms = (time.perf_counter() - t) * 1e3
    # spot check: 3000 rays against brute force over all spheres (the CPU checker), same bits and index
    rng = np.random.default_rng(n_random)
    m = 3000
    o = np.stack([rng.uniform(1.5, 98.5, m), rng.uniform(0.5, 81.0, m), rng.uniform(0.5, 169.5, m)])
    d = rng.normal(size=(3, m))
    d /= np.linalg.norm(d, axis=0)
    rays = np.concatenate([o, d]).astype(np.float32)
    d_t = torch.zeros(m, dtype=torch.float32, device="cuda")
    d_i = torch.zeros(m, dtype=torch.int32, device="cuda")
    bvh.first_hit(torch.from_numpy(rays.reshape(-1)).cuda(), m, d_t, d_i, eps=0.1)
    torch.cuda.synchronize()
    want_t, want_i = O.first_hit(rays, scene, nsph=nsph, eps=0.1)
    same = bool(np.array_equal(d_i.cpu().numpy(), want_i) and np.array_equal(d_t.cpu().numpy().view(np.uint32), want_t.view(np.uint32)))
    p = pt.default_params(width=W, height=H, samples=SPP // 4)
    mp = pt.default_material_params(seed=1, max_depth=64)
    d_img = torch.zeros((H, W, 3), dtype=torch.uint8, device="cuda")
    d_stats = torch.zeros(2, dtype=torch.int64, device="cuda")
    pt.render_image_mat_bvh(p, mp, bvh, d_img, cam_seed=3)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    pt.render_image_mat_bvh(p, mp, bvh, d_img, cam_seed=3, stats=d_stats, gamma=True)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1)
    npaths, segs = int(d_stats[0]), int(d_stats[1])
    row = {"spheres": nsph, "bvh_build_ms": build_ms, "first_hit_equals_brute_force_on_3000_rays": same, "frame": f"{W}x{H}", "spp": SPP,
           "render_ms": ms, "mpaths_s": npaths / ms / 1e3, "segments_per_path": segs / npaths, "grays_s": segs / ms / 1e6,
           "small_sphere_hits_in_check": float((want_i >= 7).mean())}
    print(json.dumps(row), flush=True)
    rows.append(row)
    bvh.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/bvh_scale.json", "w"), indent=1)
