#!/usr/bin/env python3
"""Two renders of a random-sphere scene of the given size through the BVH at 480x270, 16 spp: the command ncu is pointed at for
profiles of the traversal kernel at 10^5 / 10^6 spheres (profiles/r1_bvh_1m_ncu_summary.txt).  usage: profile_bvh_big.py N_SPHERES"""
import sys, torch, numpy as np
sys.path.insert(0,'/root/repo')
import ascendpathtracing_b200 as pt
n_random=int(sys.argv[1]); W,H,SPP=480,270,16
scene = pt.random_scene(n_random, seed=12345).reshape(11, -1)
scene[0, 7:] *= (10_000 / n_random) ** (2.0 / 3.0)
scene = np.ascontiguousarray(scene).reshape(-1)
nsph=7+n_random
bvh=pt.Bvh(torch.from_numpy(scene).cuda(), nsph, nsph)
p=pt.default_params(width=W,height=H,samples=SPP//4); mp=pt.default_material_params(seed=1,max_depth=64)
img=torch.zeros((H,W,3),dtype=torch.uint8,device='cuda')
for _ in range(2):
    pt.render_image_mat_bvh(p,mp,bvh,img,cam_seed=3)
torch.cuda.synchronize()
