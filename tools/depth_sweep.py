#!/usr/bin/env python3
"""BASELINE config C5: max-depth sweep (5 / 10 / 50 bounces) at 1920x1080 on one GPU, for
  (a) the reference-parity kernel (all-mirror, exact early termination vs the reference's fixed depth), and
  (b) the material extension (DIFF/SPEC/REFR + Russian roulette, real divergence).
Reports Mpaths/s, segments actually traced and Grays/s; --spp picks the samples per pixel (default 64)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ascendpathtracing_b200 as pt  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--spp", type=int, default=64)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
W, H, S = a.width, a.height, a.spp // 4


def timed(fn):
    fn()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(a.reps):
        fn()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / a.reps


out = {"width": W, "height": H, "spp": 4 * S, "rows": []}
p = pt.default_params(width=W, height=H, samples=S)
n = p.n_paths
d_rays = torch.empty(6 * n, dtype=torch.float32, device="cuda")
pt.gen_rays(p, d_rays, seed=5)
d_col = torch.empty(3 * n, dtype=torch.float32, device="cuda")
d_sph = torch.from_numpy(pt.default_scene()).cuda()
d_stats = torch.zeros(2, dtype=torch.int64, device="cuda")
d_img = torch.zeros((H, W, 3), dtype=torch.uint8, device="cuda")
for depth in (5, 10, 50):
    for fixed in (0, 1):
        p.depth, p.flags = depth, fixed
        ms = timed(lambda: pt.render_do_ex(p, d_rays, d_sph, d_col))
        pt.render_image(p, d_sph, d_img, seed=5, stats=d_stats)   # same paths through the production entry: counts segments
        segs = int(d_stats[1])
        out["rows"].append({"kernel": "reference-parity", "depth": depth, "mode": "fixed depth (as the reference)" if fixed else "exact early termination",
                            "ms": ms, "mpaths_s": n / ms / 1e3, "segments_traced": segs, "segments_reference": n * depth,
                            "grays_s_traced": segs / ms / 1e6, "grays_s_reference_equivalent": n * depth / ms / 1e6})
        print(out["rows"][-1], flush=True)

pm = pt.default_params(width=W, height=H, samples=S, sphere_count=9, sphere_stride=16)
d_sc = torch.from_numpy(pt.smallpt_scene()).cuda()
for depth in (5, 10, 50):
    mp = pt.default_material_params(max_depth=depth, seed=1)
    d_stats.zero_()
    pt.render_do_mat(pm, mp, d_rays, d_sc, d_col, stats=d_stats)
    torch.cuda.synchronize()
    segs = int(d_stats[0])
    ms = timed(lambda: pt.render_do_mat(pm, mp, d_rays, d_sc, d_col))
    out["rows"].append({"kernel": "materials (DIFF/SPEC/REFR + RR from depth 5)", "depth": depth, "ms": ms, "mpaths_s": n / ms / 1e3,
                        "segments_traced": segs, "segments_per_path": segs / n, "grays_s_traced": segs / ms / 1e6})
    print(out["rows"][-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/depth_sweep.json", "w"), indent=1)
