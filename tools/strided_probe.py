#!/usr/bin/env python3
"""What strided columns cost on one GPU: the whole 4K frame in one call against the same frame as eight column sets
(column_step = 8, x0 = 0..7), at --spp samples per pixel.  Sum of the eight should equal the whole frame if striding were free."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ascendpathtracing_b200 as pt  # noqa: E402

W, H = 3840, 2160
S = (int(sys.argv[1]) if len(sys.argv) > 1 else 64) // 4
d_sph = torch.from_numpy(pt.default_scene()).cuda()


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        fn()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / reps


p = pt.default_params(width=W, height=H, samples=S)
d_img = torch.zeros((H, W, 3), dtype=torch.uint8, device="cuda")
whole = timed(lambda: pt.render_image(p, d_sph, d_img, seed=1))
ps = pt.default_params(width=W, height=H, samples=S, column_step=8)
d_part = torch.zeros((H, W // 8, 3), dtype=torch.uint8, device="cuda")
parts = [timed(lambda r=r: pt.render_image(ps, d_sph, d_part, x0=r, x1=W, seed=1)) for r in range(8)]
stripes = [timed(lambda r=r: pt.render_image(p, d_sph, d_part, x0=r * W // 8, x1=(r + 1) * W // 8, seed=1)) for r in range(8)]
print(f"{4*S} spp: whole frame {whole:.2f} ms; eight strided sets {sum(parts):.2f} ms (max {max(parts):.2f}); eight contiguous stripes {sum(stripes):.2f} ms (max {max(stripes):.2f})")
