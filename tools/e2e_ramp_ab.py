#!/usr/bin/env python3
"""A/B of ptb200_render_host's chunk schedule on C2 (pinned host buffers): run once per setting of PTB200_HOST_RAMP / PTB200_HOST_CHUNK
(the library reads them once per process), prints ms per call (best and median of 9 after 2 warm-ups)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import ascendpathtracing_b200 as pt  # noqa: E402

p = pt.default_params(width=1024, height=768, samples=16)
n = p.n_paths
d_rays = torch.empty(6 * n, dtype=torch.float32, device="cuda")
pt.gen_rays(p, d_rays, seed=2024)
h_rays = torch.empty(6 * n, dtype=torch.float32).pin_memory()
h_rays.copy_(d_rays)
h_col = torch.empty(3 * n, dtype=torch.float32).pin_memory()
sph = pt.default_scene()
for _ in range(2):
    pt.render_host(p, h_rays, sph, h_col)
ts = []
for _ in range(9):
    t = time.perf_counter()
    pt.render_host(p, h_rays, sph, h_col)
    ts.append((time.perf_counter() - t) * 1e3)
print(f"ramp {os.environ.get('PTB200_HOST_RAMP', 'default')} chunk {os.environ.get('PTB200_HOST_CHUNK', 'default')}: best {min(ts):.3f} median {np.median(ts):.3f} ms")
