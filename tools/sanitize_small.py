#!/usr/bin/env python3
"""Small renders through every kernel family of the library, for `compute-sanitizer --tool memcheck|racecheck|synccheck`
(one tool per gpurun call): the reference-parity trace kernel (8-sphere and generic instantiations, regeneration and lock step,
ray files and the fused generator), the resolve kernels, the fused-resolve variant, the material kernel, the BVH build, the
warp-local wavefront kernel and the first-hit query.  Sizes are tiny: the sanitizer slows kernels down 10-100x."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import ascendpathtracing_b200 as pt  # noqa: E402


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def main():
    w, h = 24, 16
    d_sph = dev(pt.default_scene())
    for s, depth, flags in [(2, 5, 0), (2, 3, 0), (8, 5, 0), (2, 5, pt.F_FIXED_DEPTH)]:
        p = pt.default_params(width=w, height=h, samples=s, depth=depth, flags=flags)
        n = p.n_paths
        rays = torch.empty(6 * n, dtype=torch.float32, device="cuda")
        pt.gen_rays(p, rays, seed=1)
        col = torch.empty(3 * n, dtype=torch.float32, device="cuda")
        pt.render_do_ex(p, rays, d_sph, col)
        img = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda")
        pt.resolve(p, col, img)
        pt.render_image(p, d_sph, img, seed=1)
        os.environ["PTB200_FUSED_RESOLVE"] = "1"
        pt.render_image(p, d_sph, img, seed=1)
        os.environ["PTB200_FUSED_RESOLVE"] = "0"
    # generic sphere count (open scene)
    rng = np.random.RandomState(3)
    sph = np.zeros((10, 16), dtype=np.float32)
    sph[0, :12] = rng.uniform(5, 30, 12) ** 2
    sph[1:4, :12] = rng.uniform(0, 100, (3, 12))
    sph[7:10, :12] = rng.uniform(0, 1, (3, 12))
    p = pt.default_params(width=w, height=h, samples=2, sphere_count=12, sphere_stride=16, light_index=3)
    img = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda")
    pt.render_image(p, dev(sph.reshape(-1)), img, seed=2)
    # materials, constant-bank scene
    pm = pt.default_params(width=w, height=h, samples=8, sphere_count=9, sphere_stride=16)
    mp = pt.default_material_params(seed=2, max_depth=12)
    pt.render_image_mat(pm, mp, dev(pt.smallpt_scene()), img, cam_seed=4, gamma=True)
    os.environ["PTB200_FUSED_RESOLVE"] = "1"
    pt.render_image_mat(pm, mp, dev(pt.smallpt_scene()), img, cam_seed=4, gamma=True)
    os.environ["PTB200_FUSED_RESOLVE"] = "0"
    # BVH scene: build, first hit, wavefront kernel
    nrand = 300
    scene = pt.random_scene(nrand)
    bvh = pt.Bvh(dev(scene), 7 + nrand, 7 + nrand)
    pb = pt.default_params(width=w, height=h, samples=2)
    n = pb.n_paths
    rays = torch.empty(6 * n, dtype=torch.float32, device="cuda")
    pt.gen_rays(pb, rays, seed=5)
    d_t = torch.empty(n, dtype=torch.float32, device="cuda")
    d_i = torch.empty(n, dtype=torch.int32, device="cuda")
    bvh.first_hit(rays, n, d_t, d_i, eps=0.1)
    pt.render_image_mat_bvh(pb, mp, bvh, img, cam_seed=4)
    torch.cuda.synchronize()
    print("sanitize_small: all kernels ran; image mean", float(img.float().mean()))


if __name__ == "__main__":
    main()
