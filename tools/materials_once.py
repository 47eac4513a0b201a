#!/usr/bin/env python3
"""One C5-materials render (smallpt scene, DIFF/SPEC/REFR + Russian roulette, 1920x1080, --spp, depth cap --depth) through
ptb200_render_image_mat, for ncu captures of trace_materials_kernel."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ascendpathtracing_b200 as pt  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--spp", type=int, default=64)
ap.add_argument("--depth", type=int, default=50)
a = ap.parse_args()
W, H, S = 1920, 1080, a.spp // 4
p = pt.default_params(width=W, height=H, samples=S, sphere_count=9, sphere_stride=16)
mp = pt.default_material_params(seed=1, max_depth=a.depth)
d_sc = torch.from_numpy(pt.smallpt_scene()).cuda()
d_img = torch.zeros((H, W, 3), dtype=torch.uint8, device="cuda")
d_stats = torch.zeros(2, dtype=torch.int64, device="cuda")
for _ in range(2):
    pt.render_image_mat(p, mp, d_sc, d_img, cam_seed=4, gamma=True, stats=d_stats)
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
pt.render_image_mat(p, mp, d_sc, d_img, cam_seed=4, gamma=True, stats=d_stats)
t1.record()
torch.cuda.synchronize()
ms = t0.elapsed_time(t1)
segs = int(d_stats[1])
print(f"materials {W}x{H}x{4*S}spp depth cap {a.depth}: {ms:.3f} ms, {p.n_paths / ms / 1e3:.1f} Mpaths/s, {segs} segments, {segs / ms / 1e6:.2f} Gsegments/s, image mean {d_img.float().mean().item():.2f}")
