#!/usr/bin/env python3
"""How evenly does a frame's cost spread over column stripes?  Times every stripe of a `world`-way split (and of the
interleaved split) of config C4 on ONE GPU; the slowest rank's share is what bounds multi-GPU strong scaling."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ascendpathtracing_b200 as pt  # noqa: E402
from ascendpathtracing_b200 import sharding  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--world", type=int, default=8)
ap.add_argument("--parts", type=int, default=4)
ap.add_argument("--spp", type=int, default=256)
a = ap.parse_args()
W, H = 1920, 1080
nsph = 10007
d_sc = torch.from_numpy(pt.random_scene(10000, seed=12345)).cuda()
bvh = pt.Bvh(d_sc, nsph, nsph)
p = pt.default_params(width=W, height=H, samples=a.spp // 4)
mp = pt.default_material_params(seed=1, max_depth=64)


def cost(x0, x1):
    img = torch.zeros((H, x1 - x0, 3), dtype=torch.uint8, device="cuda")
    st = torch.zeros(2, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    t = time.perf_counter()
    pt.render_image_mat_bvh(p, mp, bvh, img, x0=x0, x1=x1, cam_seed=2024, gamma=True, stats=st)
    torch.cuda.synchronize()
    return time.perf_counter() - t, int(st[1])


cost(0, 64)
whole, segs = cost(0, W)
out = {"whole_s": whole, "segments": segs}
for parts in (1, a.parts):
    per_rank = []
    for r in range(a.world):
        ts = [cost(x0, x1) for x0, x1 in sharding.interleaved_stripes(W, r, a.world, parts)]
        per_rank.append({"seconds": sum(t for t, _ in ts), "segments": sum(s for _, s in ts)})
    out[f"parts{parts}"] = per_rank
    tmax, tsum = max(x["seconds"] for x in per_rank), sum(x["seconds"] for x in per_rank)
    smax, ssum = max(x["segments"] for x in per_rank), sum(x["segments"] for x in per_rank)
    print(f"parts={parts}: sum of stripe times {tsum:.4f} s (whole frame {whole:.4f}), slowest rank {tmax:.4f} s = {tmax * a.world / tsum:.3f} x mean; "
          f"segments: slowest rank {smax * a.world / ssum:.3f} x mean")
    print("  per rank ms:", [round(1e3 * x["seconds"], 1) for x in per_rank])
json.dump(out, open("gpurun_out/stripe_costs.json", "w"), indent=1)
