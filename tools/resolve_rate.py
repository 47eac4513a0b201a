#!/usr/bin/env python3
"""Resolve kernel throughput against samples per sub-pixel S (spp = 4 S): bytes of colour planes read per second."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ascendpathtracing_b200 as pt

for (w, h, s) in [(1024, 768, 16), (1024, 768, 32), (512, 384, 128), (512, 384, 256), (256, 192, 1024), (1024, 768, 4), (1024, 768, 17)]:
    p = pt.default_params(width=w, height=h, samples=s)
    n = p.n_paths
    d_col = torch.rand(3 * n, dtype=torch.float32, device="cuda")
    d_img = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda")
    pt.resolve(p, d_col, d_img)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        pt.resolve(p, d_col, d_img)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(f"{w}x{h} S={s:5d}: {12 * n / 1e6:8.1f} MB in {ms:7.3f} ms = {12 * n / ms / 1e9:6.2f} TB/s")
