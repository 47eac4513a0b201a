#!/usr/bin/env python3
"""Would the host-buffer entry be faster if the trace kernel read its rays straight from pinned host memory (and wrote its
colours there) instead of going through staged chunks?  Pinned allocations are device-addressable under UVA, so the probe
needs no new code: render_do_ex is simply handed pinned HOST pointers.  C2 (50 331 648 paths, 1.2 GB in, 0.6 GB out).
    python tools/zero_copy_probe.py        -> gpurun_out/zero_copy_probe.json
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ascendpathtracing_b200 as pt  # noqa: E402

p = pt.default_params(width=1024, height=768, samples=16)
n = p.n_paths
d_rays = torch.empty(6 * n, dtype=torch.float32, device="cuda")
pt.gen_rays(p, d_rays, seed=2024)
d_sph = torch.from_numpy(pt.default_scene()).cuda()
d_col = torch.empty(3 * n, dtype=torch.float32, device="cuda")
h_rays = torch.empty(6 * n, dtype=torch.float32).pin_memory()
h_rays.copy_(d_rays)
h_col = torch.empty(3 * n, dtype=torch.float32).pin_memory()
h_sph = pt.default_scene()
pt.render_do_ex(p, d_rays, d_sph, d_col)
torch.cuda.synchronize()
want = d_col.cpu()


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps * 1e3


out = {}
out["render_host (staged chunks, what ships)"] = timed(lambda: pt.render_host(p, h_rays, h_sph, h_col))
assert torch.equal(h_col.view(torch.int32), want.view(torch.int32))
out["device rays, device colours (kernel alone)"] = timed(lambda: pt.render_do_ex(p, d_rays, d_sph, d_col))
h_col.zero_()
out["zero-copy rays, device colours"] = timed(lambda: pt.render_do_ex(p, h_rays, d_sph, d_col))
out["device rays, zero-copy colours"] = timed(lambda: pt.render_do_ex(p, d_rays, d_sph, h_col))
assert torch.equal(h_col.view(torch.int32), want.view(torch.int32))
h_col.zero_()
out["zero-copy rays and colours"] = timed(lambda: pt.render_do_ex(p, h_rays, d_sph, h_col))
assert torch.equal(h_col.view(torch.int32), want.view(torch.int32))


def rays_zero_copy_colours_copied():
    pt.render_do_ex(p, h_rays, d_sph, d_col)
    h_col.copy_(d_col, non_blocking=True)


out["zero-copy rays, colours by one D2H copy afterwards"] = timed(rays_zero_copy_colours_copied)
for k, v in out.items():
    print(f"{v:8.2f} ms  {n / v / 1e3:8.1f} Mpaths/s  {k}", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/zero_copy_probe.json", "w"), indent=1)
