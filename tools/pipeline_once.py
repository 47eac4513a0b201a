#!/usr/bin/env python3
"""Runs ptb200_render_image a few times on C2 (for ncu launch lists of the production composition)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ascendpathtracing_b200 as pt  # noqa: E402

w, h, s = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (1024, 768, 16)))
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
p = pt.default_params(width=w, height=h, samples=s)
d_sph = torch.from_numpy(pt.default_scene()).cuda()
d_img = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda")
for _ in range(reps):
    pt.render_image(p, d_sph, d_img, seed=1)
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
pt.render_image(p, d_sph, d_img, seed=1)
t1.record()
torch.cuda.synchronize()
print(f"{w}x{h}x{4*s}spp: {t0.elapsed_time(t1):.3f} ms, {p.n_paths / t0.elapsed_time(t1) / 1e3:.1f} Mpaths/s, image mean {d_img.float().mean().item():.2f}")
