#!/usr/bin/env python3
"""Lock step vs regeneration of the reference-parity kernel at depth 2..6 (run with PTB200_EARLY_FROM_DEPTH=1 so that both
modes can be selected at every depth): where the library's crossover (depth 5) comes from."""
import sys, torch
sys.path.insert(0, '/root/repo')
import ascendpathtracing_b200 as pt
W,H,S=1920,1080,16
p=pt.default_params(width=W,height=H,samples=S)
n=p.n_paths
d_rays=torch.empty(6*n,dtype=torch.float32,device='cuda'); pt.gen_rays(p,d_rays,seed=5)
d_col=torch.empty(3*n,dtype=torch.float32,device='cuda'); d_sph=torch.from_numpy(pt.default_scene()).cuda()
for depth in (2,3,4,5,6):
    for fixed in (0,1):
        p.depth,p.flags=depth,fixed
        pt.render_do_ex(p,d_rays,d_sph,d_col); torch.cuda.synchronize()
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): pt.render_do_ex(p,d_rays,d_sph,d_col)
        b.record(); torch.cuda.synchronize()
        print(depth, 'fixed' if fixed else 'early', round(a.elapsed_time(b)/5,3))
