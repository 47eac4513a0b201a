"""GPU: large scenes through the GPU-built sphere BVH (BASELINE config C4).  Acceptance: the tree must
return the brute-force loop's nearest hit -- same t bits, same sphere index, lowest index on ties -- on the same rays;
the material kernel on top of it must then equal the CPU twin, which is brute force over all spheres."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def scene_rays(rng, n):
    """Rays that start inside the box (where every bounce of a closed scene starts), any direction."""
    o = np.stack([rng.uniform(1.5, 98.5, n), rng.uniform(0.5, 81.0, n), rng.uniform(0.5, 169.5, n)])
    d = rng.normal(size=(3, n))
    d /= np.linalg.norm(d, axis=0)
    return np.concatenate([o, d]).astype(np.float32)


@pytest.mark.parametrize("n_random,eps", [(10000, 1e-4), (10000, 0.1), (1000, 1e-4), (1, 1e-4), (0, 1e-4), (2, 0.1), (37, 1e-4)])
def test_bvh_first_hit_equals_brute_force(pt, cuda, oracle, n_random, eps):
    torch = cuda
    rng = np.random.default_rng(n_random + 17)
    scene = pt.random_scene(n_random)
    nsph = 7 + n_random
    bvh = pt.Bvh(dev(torch, scene), nsph, nsph)
    info = bvh.info()
    assert info == {"spheres": nsph, "big": 7, "small": n_random, "nodes": max(n_random - 1, 0)}
    n = 400000
    rays = scene_rays(rng, n)
    # camera rays as well: they start OUTSIDE the box (z = 295.6 - ...), further than any bounce ever is
    cam = oracle.gen_rays(128, 96, 2, seed=0)
    rays = np.concatenate([rays, cam], axis=1)
    n = rays.shape[1]
    d_t = torch.zeros(n, dtype=torch.float32, device="cuda")
    d_i = torch.full((n,), -1, dtype=torch.int32, device="cuda")
    bvh.first_hit(dev(torch, rays.reshape(-1)), n, d_t, d_i, eps=eps)
    torch.cuda.synchronize()
    want_t, want_i = oracle.first_hit(rays, scene, nsph=nsph, eps=eps)
    assert np.array_equal(d_i.cpu().numpy(), want_i)
    assert np.array_equal(bits(d_t.cpu().numpy()), bits(want_t))
    if n_random >= 1000:
        assert (want_i >= 7).mean() > 0.05   # the tree really is exercised: a good share of rays end on small spheres
    bvh.close()


def test_bvh_axis_aligned_grazing_and_far_rays(pt, cuda, oracle):
    """Rays the quantised slab test has to survive: direction components exactly 0 (1 / d = inf, NaN planes), rays aimed at a
    sphere's silhouette (hit or miss decided by rounding noise: the padded boxes must never cull a brute-force hit), origins
    on sphere centres, and origins far outside the tree (beyond 2^21 grid units: the exact fallback over all spheres)."""
    torch = cuda
    rng = np.random.default_rng(5)
    n_random = 4000
    scene = pt.random_scene(n_random)
    nsph = 7 + n_random
    sc = scene.reshape(11, -1)
    bvh = pt.Bvh(dev(torch, scene), nsph, nsph)
    m = 60000
    o = np.stack([rng.uniform(1.5, 98.5, m), rng.uniform(0.5, 81.0, m), rng.uniform(0.5, 169.5, m)])
    # 1. axis-aligned and plane-aligned directions
    axes = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1], [0.6, 0.8, 0], [0, -0.6, 0.8], [0.8, 0, -0.6]], dtype=np.float64).T
    d1 = axes[:, rng.integers(0, axes.shape[1], m)]
    # 2. grazing: aim at a random small sphere's silhouette (offset = radius, perpendicular to the line of sight)
    k = rng.integers(7, nsph, m)
    c = sc[1:4, k].astype(np.float64)
    r = np.sqrt(sc[0, k].astype(np.float64))
    los = c - o
    los /= np.linalg.norm(los, axis=0)
    perp = np.cross(los.T, rng.normal(size=(m, 3))).T
    perp /= np.linalg.norm(perp, axis=0)
    target = c + perp * r * rng.choice([0.999999, 1.0, 1.000001], m)
    d2 = target - o
    d2 /= np.linalg.norm(d2, axis=0)
    # 3. origins exactly on sphere centres (inside a small sphere), random directions
    o3 = sc[1:4, rng.integers(7, nsph, m)].astype(np.float64)
    d3 = rng.normal(size=(3, m))
    d3 /= np.linalg.norm(d3, axis=0)
    # 4. far origins looking back at the scene, and a few absurd ones
    far = rng.choice([3e3, 1e5, 1e7, 1e9], m)
    d4 = rng.normal(size=(3, m))
    d4 /= np.linalg.norm(d4, axis=0)
    centre = np.array([[50.0], [40.0], [85.0]])
    o4 = centre - d4 * far + rng.normal(size=(3, m)) * 20
    rays = np.concatenate([np.concatenate([o, d1]), np.concatenate([o, d2]), np.concatenate([o3, d3]), np.concatenate([o4, d4])], axis=1).astype(np.float32)
    n = rays.shape[1]
    d_t = torch.zeros(n, dtype=torch.float32, device="cuda")
    d_i = torch.full((n,), -1, dtype=torch.int32, device="cuda")
    for eps in (1e-4, 0.1):
        bvh.first_hit(dev(torch, rays.reshape(-1)), n, d_t, d_i, eps=eps)
        torch.cuda.synchronize()
        want_t, want_i = oracle.first_hit(rays, scene, nsph=nsph, eps=eps)
        got_t, got_i = d_t.cpu().numpy(), d_i.cpu().numpy()
        bad = np.flatnonzero((got_i != want_i) | (bits(got_t) != bits(want_t)))
        assert bad.size == 0, (eps, bad[:10], got_i[bad[:10]], want_i[bad[:10]], got_t[bad[:10]], want_t[bad[:10]])
    assert (want_i[m:2 * m] >= 7).mean() > 0.2   # the grazing rays do end on small spheres often enough to matter
    bvh.close()


def test_bvh_clustered_and_duplicate_spheres(pt, cuda, oracle):
    """Degenerate input for an LBVH: many coincident centres (identical Morton codes), touching and nested spheres."""
    torch = cuda
    rng = np.random.default_rng(3)
    n_random = 3000
    scene = pt.random_scene(n_random).reshape(11, -1)
    scene[1:4, 7:1007] = np.array([[50.0], [40.0], [80.0]], dtype=np.float32)       # 1000 spheres on one point...
    scene[0, 7:1007] = np.linspace(0.04, 9.0, 1000, dtype=np.float32)               # ...of growing radius (nested)
    scene[1:4, 1007:2007] = scene[1:4, 2007:3007]                                    # 1000 exact duplicates of other spheres
    scene[0, 1007:2007] = scene[0, 2007:3007]
    scene = np.ascontiguousarray(scene).reshape(-1)
    nsph = 7 + n_random
    bvh = pt.Bvh(dev(torch, scene), nsph, nsph)
    n = 300000
    rays = scene_rays(rng, n)
    d_t = torch.zeros(n, dtype=torch.float32, device="cuda")
    d_i = torch.zeros(n, dtype=torch.int32, device="cuda")
    bvh.first_hit(dev(torch, rays.reshape(-1)), n, d_t, d_i)
    torch.cuda.synchronize()
    want_t, want_i = oracle.first_hit(rays, scene, nsph=nsph)
    assert np.array_equal(d_i.cpu().numpy(), want_i)     # duplicates tie exactly: the lower index must win
    assert np.array_equal(bits(d_t.cpu().numpy()), bits(want_t))


def test_materials_through_bvh_equal_cpu_twin(pt, cuda, oracle):
    torch = cuda
    n_random = 2000
    scene = pt.random_scene(n_random)
    nsph = 7 + n_random
    w, h, s = 48, 32, 2
    p = pt.default_params(width=w, height=h, samples=s)
    mp = pt.default_material_params(seed=4, max_depth=12)
    n = p.n_paths
    rays = oracle.gen_rays(w, h, s, seed=0)
    bvh = pt.Bvh(dev(torch, scene), nsph, nsph)
    d_col = torch.full((3 * n,), float("nan"), dtype=torch.float32, device="cuda")
    d_stats = torch.zeros(1, dtype=torch.int64, device="cuda")
    pt.render_do_mat_bvh(p, mp, bvh, dev(torch, rays.reshape(-1)), d_col, stats=d_stats)
    torch.cuda.synchronize()
    want, segs = oracle.trace_materials(rays, scene, nsph, nsph, max_depth=12, seed=4, return_segments=True)
    assert np.array_equal(bits(d_col.cpu().numpy().reshape(3, n)), bits(want))
    assert int(d_stats[0]) == segs
    # the image entry, in stripes
    d_img = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda")
    pt.render_image_mat_bvh(p, mp, bvh, d_img, cam_seed=8)
    rays2 = oracle.gen_rays_from_uniforms(w, h, s, 0, w, oracle.philox_uniforms(8, 0, n))
    want_img = oracle.resolve(oracle.trace_materials(rays2, scene, nsph, nsph, max_depth=12, seed=4), w, h, s)
    assert np.array_equal(d_img.cpu().numpy(), want_img)
    d_part = torch.zeros((h, 10, 3), dtype=torch.uint8, device="cuda")
    pt.render_image_mat_bvh(p, mp, bvh, d_part, x0=5, x1=15, cam_seed=8)
    assert np.array_equal(d_part.cpu().numpy(), want_img[:, 5:15])
    # a small scene through the tree equals the same scene through the constant-bank kernel
    small = pt.smallpt_scene()
    tree9 = pt.Bvh(dev(torch, small), 9, 16)
    assert tree9.info()["big"] == 7 and tree9.info()["small"] == 2
    p9 = pt.default_params(width=w, height=h, samples=s, sphere_count=9, sphere_stride=16)
    d_a = torch.zeros(3 * n, dtype=torch.float32, device="cuda")
    d_b = torch.zeros(3 * n, dtype=torch.float32, device="cuda")
    pt.render_do_mat(p9, mp, dev(torch, rays.reshape(-1)), dev(torch, small), d_a)
    pt.render_do_mat_bvh(p9, mp, tree9, dev(torch, rays.reshape(-1)), d_b)
    torch.cuda.synchronize()
    assert torch.equal(d_a.view(torch.int32), d_b.view(torch.int32))
