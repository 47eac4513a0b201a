"""GPU: the material extension (DIFF / SPEC / REFR + Russian roulette).  PARITY UNPINNED by the reference (it has no
materials); checked (1) bit for bit against this repo's own CPU twin oracle/pt_oracle_mat.c, written independently
from the CUDA code, and (2) statistically against a binary64 textbook formulation with an unrelated RNG."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_sincos_polynomial_is_accurate(oracle):
    u = np.linspace(0, 1, 200001, endpoint=False).astype(np.float32)
    s, c = oracle.sincos2pi(u)
    assert np.abs(s - np.sin(2 * np.pi * u.astype(np.float64))).max() < 2e-7
    assert np.abs(c - np.cos(2 * np.pi * u.astype(np.float64))).max() < 2e-7


@pytest.mark.parametrize("w,h,s,max_depth,rr_start,eps", [(64, 48, 4, 64, 5, 0.1), (33, 21, 3, 7, 2, 0.1), (64, 48, 2, 64, 5, 1e-4),
                                                          (16, 16, 1, 1, 5, 0.1), (40, 30, 8, 200, 0, 0.1)])
def test_materials_bit_exact_vs_cpu_twin(pt, cuda, oracle, w, h, s, max_depth, rr_start, eps):
    torch = cuda
    p = pt.default_params(width=w, height=h, samples=s, sphere_count=9, sphere_stride=16)
    mp = pt.default_material_params(max_depth=max_depth, rr_start=rr_start, hit_epsilon=eps, seed=0xfeedbeef12345)
    n = p.n_paths
    rays = oracle.gen_rays(w, h, s, seed=0)
    scene = pt.smallpt_scene()
    d_col = torch.full((3 * n,), float("nan"), dtype=torch.float32, device="cuda")
    d_stats = torch.zeros(1, dtype=torch.int64, device="cuda")
    pt.render_do_mat(p, mp, dev(torch, rays.reshape(-1)), dev(torch, scene), d_col, path0=0, stats=d_stats)
    torch.cuda.synchronize()
    want, segs = oracle.trace_materials(rays, scene, 9, 16, max_depth=max_depth, rr_start=rr_start, eps=eps, seed=mp.seed, return_segments=True)
    got = d_col.cpu().numpy().reshape(3, n)
    assert np.array_equal(bits(got), bits(want))
    assert int(d_stats[0]) == segs


def test_materials_slices_use_global_path_indices(pt, cuda, oracle):
    torch = cuda
    w, h, s = 48, 32, 2
    p = pt.default_params(width=w, height=h, samples=s, sphere_count=9, sphere_stride=16)
    mp = pt.default_material_params(seed=9)
    n = p.n_paths
    rays = oracle.gen_rays(w, h, s, seed=0)
    scene = pt.smallpt_scene()
    want = oracle.trace_materials(rays, scene, 9, 16, seed=9)
    d_rays, d_sc = dev(torch, rays.reshape(-1)), dev(torch, scene)
    d_col = torch.zeros(3 * n, dtype=torch.float32, device="cuda")
    cuts = [0, 1000, 1001, 5000, n]
    for a, b in zip(cuts, cuts[1:]):
        pt.render_do_mat(p, mp, d_rays, d_sc, d_col, first=a, count=b - a, path0=a)
    torch.cuda.synchronize()
    assert np.array_equal(bits(d_col.cpu().numpy().reshape(3, n)), bits(want))


def test_material_image_pipeline_and_stripes(pt, cuda, oracle):
    torch = cuda
    w, h, s = 64, 48, 4
    p = pt.default_params(width=w, height=h, samples=s, sphere_count=9, sphere_stride=16)
    mp = pt.default_material_params(seed=3)
    n = p.n_paths
    scene = pt.smallpt_scene()
    d_sc = dev(torch, scene)
    d_img = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda")
    d_stats = torch.zeros(2, dtype=torch.int64, device="cuda")
    pt.render_image_mat(p, mp, d_sc, d_img, cam_seed=11, stats=d_stats)
    rays = oracle.gen_rays_from_uniforms(w, h, s, 0, w, oracle.philox_uniforms(11, 0, n))
    col, segs = oracle.trace_materials(rays, scene, 9, 16, seed=3, return_segments=True)
    assert np.array_equal(d_img.cpu().numpy(), oracle.resolve(col, w, h, s))
    assert int(d_stats[0]) == n and int(d_stats[1]) == segs
    for x0, x1 in [(0, 20), (20, 21), (21, 64)]:
        d_part = torch.zeros((h, x1 - x0, 3), dtype=torch.uint8, device="cuda")
        pt.render_image_mat(p, mp, d_sc, d_part, x0=x0, x1=x1, cam_seed=11)
        assert np.array_equal(d_part.cpu().numpy(), d_img.cpu().numpy()[:, x0:x1])
    # gamma display transform (smallpt's toInt): within 1 LSB of the double-precision formula
    d_g = torch.zeros_like(d_img)
    pt.render_image_mat(p, mp, d_sc, d_g, cam_seed=11, gamma=True)
    lin = col.reshape(3, w, h, 4, s).astype(np.float32)
    mean = lin.mean(axis=4, dtype=np.float32).astype(np.float64).sum(axis=3) / 4       # [3][w][h]
    ref = (np.clip(mean, 0, 1) ** (1 / 2.2) * 255 + 0.5).astype(np.uint8)             # [3][w][h]
    ref_img = ref.transpose(2, 1, 0)[::-1]                                            # rows top first
    diff = np.abs(d_g.cpu().numpy().astype(int) - ref_img.astype(int))
    assert diff.max() <= 1 and (diff > 0).mean() < 1e-3


def test_materials_converge_to_binary64_formulation(pt, cuda, oracle):
    """Statistical parity: binary32 twin on the GPU vs the binary64 textbook version with an unrelated generator.
    At 64x48, 1024 spp the two means agree to within Monte-Carlo noise; with the reference's epsilon (1e-4) binary32
    leaks through its own 1e5-radius walls and the image is ~40 % too bright -- the reason hit_epsilon defaults to 0.1."""
    torch = cuda
    w, h, s = 64, 48, 256
    p = pt.default_params(width=w, height=h, samples=s, sphere_count=9, sphere_stride=16)
    scene = pt.smallpt_scene()
    n = p.n_paths
    rays = oracle.gen_rays_from_uniforms(w, h, s, 0, w, oracle.philox_uniforms(5, 0, n))
    d_rays, d_sc = dev(torch, rays.reshape(-1)), dev(torch, scene)
    ref = oracle.trace_materials_f64(rays, scene, 9, 16, seed=77)
    ref_px = ref.reshape(3, w * h, 4 * s).mean(axis=2)
    out = {}
    for eps in (0.1, 1e-4):
        mp = pt.default_material_params(seed=6, hit_epsilon=eps)
        d_col = torch.zeros(3 * n, dtype=torch.float32, device="cuda")
        pt.render_do_mat(p, mp, d_rays, d_sc, d_col)
        got_px = d_col.cpu().numpy().reshape(3, w * h, 4 * s).mean(axis=2)
        out[eps] = (got_px.mean() / ref_px.mean(), np.sqrt(((np.clip(got_px, 0, 1) - np.clip(ref_px, 0, 1)) ** 2).mean()))
    assert abs(out[0.1][0] - 1.0) < 0.03, out        # mean radiance within 3 %
    assert out[0.1][1] < 0.08, out                   # per-pixel RMSE (clipped) at 1024 spp: noise level
    assert out[1e-4][0] > 1.2, out                   # the leak


def test_materials_depth_sweep_segments(pt, cuda, oracle):
    """C5-style depth sweep on the material kernel: segments per path grow with the cap and saturate under Russian roulette."""
    torch = cuda
    w, h, s = 64, 48, 8
    p = pt.default_params(width=w, height=h, samples=s, sphere_count=9, sphere_stride=16)
    n = p.n_paths
    rays = oracle.gen_rays(w, h, s, seed=0)
    d_rays, d_sc = dev(torch, rays.reshape(-1)), dev(torch, pt.smallpt_scene())
    per_path = []
    for depth in (5, 10, 50):
        mp = pt.default_material_params(max_depth=depth, seed=1)
        d_col = torch.zeros(3 * n, dtype=torch.float32, device="cuda")
        d_stats = torch.zeros(1, dtype=torch.int64, device="cuda")
        pt.render_do_mat(p, mp, d_rays, d_sc, d_col, stats=d_stats)
        torch.cuda.synchronize()
        per_path.append(int(d_stats[0]) / n)
    assert per_path[0] <= 5 and per_path[0] < per_path[1] <= per_path[2] < 12
