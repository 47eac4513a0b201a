"""Parity at the sizes BASELINE.json quotes for configs C4 and C5 (C2 and C3 are in test_gpu_parity.py), plus the scene
sizes at the edges of each kernel's range.  The full frames are far beyond what a CPU oracle can recompute, so -- as for
C3 -- the frames are rendered whole through the production entries and windows of them are recomputed by the oracle from
the GLOBAL path indices and compared bit for bit; where a whole small stripe is affordable the segment statistics are
compared too.  What this covers that the small tests do not: tiles of 2^29 paths, launch chunking, 32-bit index decode at
real widths, column stripes of a real frame, and the wavefront kernel's chunk dispenser over half a billion paths."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def window_rays(oracle, w, h, s, seed, x, wy, rows):
    """Rays of pixels (x, wy .. wy+rows-1) of the W x H frame under the counter-based RNG keyed by global path indices;
    returns (rays [6][m], global index of the first path)."""
    per_col = h * 4 * s
    first = x * per_col + wy * 4 * s
    m = rows * 4 * s
    ucol = np.zeros(2 * per_col)
    ucol[2 * wy * 4 * s:2 * (wy * 4 * s + m)] = oracle.philox_uniforms(seed, first, m)
    rays = oracle.gen_rays_from_uniforms(w, h, s, x, x + 1, ucol)[:, wy * 4 * s:wy * 4 * s + m]
    return rays, first


def test_c4_full_frame_10k_spheres_bvh_windows(pt, cuda, oracle):
    """BASELINE config C4 at full size: 10 007 spheres (six walls, the light, 10 000 random spheres of all three materials,
    MT19937 seed 12345), 1920 x 1080 at 256 spp = 530 841 600 paths through ptb200_render_image_mat_bvh in one call.
    Three windows of the frame are recomputed by the CPU twin with a BRUTE-FORCE loop over all 10 007 spheres: the tree must
    give the same hits (t bits, index, lowest index on ties) along every path of every sample, or the 8-bit pixels differ."""
    torch = cuda
    oracle.set_threads(os.cpu_count() or 8)
    w, h, s, cam_seed = 1920, 1080, 64, 5
    n_random = 10000
    nsph = 7 + n_random
    scene = pt.random_scene(n_random)            # seed 12345
    assert np.array_equal(scene, oracle.random_scene(n_random))
    p = pt.default_params(width=w, height=h, samples=s)
    mp = pt.default_material_params(seed=11)    # max_depth 64, roulette from depth 5, epsilon 0.1
    tree = pt.Bvh(dev(torch, scene), nsph, nsph)
    assert tree.info() == {"spheres": nsph, "big": 7, "small": n_random, "nodes": n_random - 1}
    d_img = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda")
    d_stats = torch.zeros(2, dtype=torch.int64, device="cuda")
    pt.render_image_mat_bvh(p, mp, tree, d_img, cam_seed=cam_seed, stats=d_stats)
    img = d_img.cpu().numpy()
    tree.close()
    assert int(d_stats[0]) == p.n_paths == 530841600
    assert 4 * p.n_paths < int(d_stats[1]) < 12 * p.n_paths      # ~7.7 segments per path in this scene
    total = 0
    for (wx, wy, cols, rows) in [(952, 532, 16, 16), (0, 0, 8, 8), (1912, 1072, 8, 8)]:   # centre, bottom-left, top-right corners
        win = np.zeros((rows, cols, 3), dtype=np.uint8)
        for cx in range(cols):
            rays, first = window_rays(oracle, w, h, s, cam_seed, wx + cx, wy, rows)
            col, segs = oracle.trace_materials(rays, scene, nsph, nsph, seed=11, path0=first, return_segments=True)
            total += segs
            win[:, cx] = oracle.resolve(col, 1, rows, s)[:, 0]      # a 1-column image: row r = y index rows-1-r
        got = img[h - wy - rows:h - wy, wx:wx + cols]
        assert np.array_equal(got, win), (wx, wy)
    assert total > 0


@pytest.mark.parametrize("depth", [5, 10, 50])
def test_c5_depth_sweep_full_frame_stripes(pt, cuda, oracle, depth):
    """BASELINE config C5: 1920 x 1080 at 512 spp, depth 5 / 10 / 50, reference-parity kernel.  (a) The stripe one of 8 GPUs
    renders (240 columns = 132.7 M paths) through the production entry, two windows recomputed by the oracle at that depth;
    (b) a two-column stripe of the same frame in full: every pixel AND the number of segments traced must equal the
    oracle's count of live segments (exact early termination at work: 22 % / 44 % / 71 % of the bounces are skipped)."""
    torch = cuda
    oracle.set_threads(os.cpu_count() or 8)
    w, h, s, seed = 1920, 1080, 128, 13
    p = pt.default_params(width=w, height=h, samples=s, depth=depth)
    sph = oracle.gen_spheres()
    d_sph = dev(torch, pt.default_scene())
    d_stats = torch.zeros(2, dtype=torch.int64, device="cuda")
    x0, x1 = 960, 1200
    d_img = torch.zeros((h, x1 - x0, 3), dtype=torch.uint8, device="cuda")
    pt.render_image(p, d_sph, d_img, x0=x0, x1=x1, seed=seed, stats=d_stats)
    img = d_img.cpu().numpy()
    assert int(d_stats[0]) == (x1 - x0) * h * 4 * s == 132710400
    assert int(d_stats[1]) <= int(d_stats[0]) * depth
    for (wx, wy, cols, rows) in [(960, 0, 8, 16), (1192, 1064, 8, 16)]:
        win = np.zeros((rows, cols, 3), dtype=np.uint8)
        for cx in range(cols):
            rays, _ = window_rays(oracle, w, h, s, seed, wx + cx, wy, rows)
            win[:, cx] = oracle.resolve(oracle.trace(rays, sph, depth=depth), 1, rows, s)[:, 0]
        got = img[h - wy - rows:h - wy, wx - x0:wx - x0 + cols]
        assert np.array_equal(got, win), (wx, wy)
    # (b) columns 1000..1001 in full
    xa, xb = 1000, 1002
    per_col = h * 4 * s
    d_two = torch.zeros((h, xb - xa, 3), dtype=torch.uint8, device="cuda")
    pt.render_image(p, d_sph, d_two, x0=xa, x1=xb, seed=seed, stats=d_stats)
    u = oracle.philox_uniforms(seed, xa * per_col, (xb - xa) * per_col)
    rays = oracle.gen_rays_from_uniforms(w, h, s, xa, xb, u)
    col, live = oracle.trace(rays, sph, depth=depth, return_live=True)
    assert np.array_equal(d_two.cpu().numpy(), oracle.resolve(col, xb - xa, h, s))
    assert int(d_stats[0]) == (xb - xa) * per_col
    assert int(d_stats[1]) == live
    # the same two columns cut out of the 240-column stripe
    assert np.array_equal(img[:, xa - x0:xb - x0], d_two.cpu().numpy())


def test_first_hit_100k_spheres_equals_brute_force(pt, cuda, oracle):
    """10^5 random spheres (radius scaled to keep the box equally crowded): the tree's nearest hit on 32 768 rays -- camera
    rays and random interior rays -- equals the brute-force loop bit for bit (t, index)."""
    torch = cuda
    oracle.set_threads(os.cpu_count() or 8)
    n_random = 100000
    nsph = 7 + n_random
    scene = pt.random_scene(n_random, seed=777).copy()   # flat SoA [11][nsph]; row 0 = r^2
    scene[7:nsph] *= np.float32(0.1)                     # r -> r / sqrt(10): the same total projected area as the 10 k scene
    rng = np.random.default_rng(5)
    n = 32768
    cam = oracle.gen_rays_from_uniforms(128, 64, 1, 0, 128, oracle.philox_uniforms(3, 0, 128 * 64 * 4))[:, :n // 2]
    o = np.stack([rng.uniform(2, 98, n // 2), rng.uniform(1, 80, n // 2), rng.uniform(1, 169, n // 2)])
    d = rng.normal(size=(3, n // 2))
    d /= np.linalg.norm(d, axis=0)
    rays = np.concatenate([cam, np.concatenate([o, d]).astype(np.float32)], axis=1)
    tree = pt.Bvh(dev(torch, scene), nsph, nsph)
    d_t = torch.zeros(n, dtype=torch.float32, device="cuda")
    d_i = torch.zeros(n, dtype=torch.int32, device="cuda")
    for eps, m in ((1e-4, n), (0.1, n // 4)):   # the reference's epsilon on all rays, the material kernel's on a quarter
        sub = np.ascontiguousarray(rays[:, ::n // m])
        tree.first_hit(dev(torch, sub.reshape(-1)), m, d_t, d_i, eps=eps)
        torch.cuda.synchronize()
        want_t, want_i = oracle.first_hit(sub, scene, nsph=nsph, eps=eps)
        assert np.array_equal(d_i.cpu().numpy()[:m], want_i), eps
        assert np.array_equal(bits(d_t.cpu().numpy()[:m]), bits(want_t)), eps
        assert (want_i >= 7).mean() > 0.15       # the small spheres really are in the way
    tree.close()


@pytest.mark.parametrize("nsph", [768, 769, 1024])
def test_brute_force_kernel_at_its_sphere_limit(pt, cuda, oracle, nsph):
    """check_params admits up to 1024 spheres for the constant-bank kernels; above 768 (mirror) / 512 (materials) the
    kernels need more than the default 48 KB of dynamic shared memory and must opt in -- a bare launch failure before."""
    torch = cuda
    rng = np.random.default_rng(nsph)
    w, h, s, depth = 32, 16, 1, 6
    stride = 1024
    sph = np.zeros(11 * stride, dtype=np.float32)
    r = rng.uniform(1, 6, nsph)
    sph[0:nsph] = (r * r).astype(np.float32)
    sph[1 * stride:1 * stride + nsph] = rng.uniform(0, 100, nsph)
    sph[2 * stride:2 * stride + nsph] = rng.uniform(0, 80, nsph)
    sph[3 * stride:3 * stride + nsph] = rng.uniform(0, 170, nsph)
    for m in (7, 8, 9):
        sph[m * stride:m * stride + nsph] = rng.uniform(0, 1, nsph)
    sph[4 * stride + nsph - 1] = sph[5 * stride + nsph - 1] = sph[6 * stride + nsph - 1] = 12.0   # the last sphere shines
    sph[10 * stride:10 * stride + nsph] = rng.integers(0, 3, nsph)
    p = pt.default_params(width=w, height=h, samples=s, depth=depth, sphere_count=nsph, sphere_stride=stride, light_index=nsph - 1)
    n = p.n_paths
    o = np.stack([rng.uniform(0, 100, n), rng.uniform(0, 80, n), rng.uniform(0, 170, n)])
    d = rng.normal(size=(3, n))
    d /= np.linalg.norm(d, axis=0)
    rays = np.concatenate([o, d]).astype(np.float32)
    d_rays, d_sph = dev(torch, rays.reshape(-1)), dev(torch, sph)
    for flags in (0, 1):
        p.flags = flags
        d_col = torch.full((3 * n,), float("nan"), dtype=torch.float32, device="cuda")
        pt.render_do_ex(p, d_rays, d_sph, d_col)
        torch.cuda.synchronize()
        want = oracle.trace(rays, sph, depth=depth, nsph=nsph, stride=stride, light=nsph - 1)
        assert np.array_equal(bits(d_col.cpu().numpy().reshape(3, n)), bits(want)), flags
    mp = pt.default_material_params(seed=2, max_depth=8)
    d_col = torch.full((3 * n,), float("nan"), dtype=torch.float32, device="cuda")
    pt.render_do_mat(p, mp, d_rays, d_sph, d_col)
    torch.cuda.synchronize()
    want = oracle.trace_materials(rays, sph, nsph, stride, max_depth=8, seed=2)
    assert np.array_equal(bits(d_col.cpu().numpy().reshape(3, n)), bits(want))


def test_gpu_first_hit_matches_reference_test_scene(pt, cuda, oracle, golden_dir):
    """The reference's own first-hit model (scripts/gen_data.py:134-188, fixture made by tests/golden/make_golden_large.py)
    against the nearest-hit stage of the CUDA path: ptb200_bvh_first_hit over the reference's eight spheres (all of them
    end up in the pairwise brute-force list: radius >= 100 or a single leaf) and, independently, one bounce of the radiance
    kernel (the colour after depth 1 is 12 x the hit sphere's colour, 12 x 1 for the light)."""
    torch = cuda
    z = np.load(os.path.join(golden_dir, "w64h64s1_test_scene.npz"))
    want = z["first_hit_index"].astype(np.int32)
    w = h = 64
    rays = oracle.gen_rays(w, h, 1, seed=0)
    n = rays.shape[1]
    scene11 = np.zeros(11 * 8, dtype=np.float32)
    scene11[:80] = pt.default_scene()[:80]
    tree = pt.Bvh(dev(torch, scene11), 8, 8)
    d_t = torch.zeros(n, dtype=torch.float32, device="cuda")
    d_i = torch.zeros(n, dtype=torch.int32, device="cuda")
    tree.first_hit(dev(torch, rays.reshape(-1)), n, d_t, d_i, eps=1e-4)
    torch.cuda.synchronize()
    tree.close()
    got = d_i.cpu().numpy()
    o_t, o_i = oracle.first_hit(rays, oracle.gen_spheres())
    assert np.array_equal(got, o_i) and np.array_equal(bits(d_t.cpu().numpy()), bits(o_t))
    assert (got == want).mean() >= 0.999, (got != want).sum()
    # one bounce of the radiance kernel
    p = pt.default_params(width=w, height=h, samples=1, depth=1)
    d_col = torch.zeros(3 * n, dtype=torch.float32, device="cuda")
    pt.render_do_ex(p, dev(torch, rays.reshape(-1)), dev(torch, pt.default_scene()), d_col)
    col = d_col.cpu().numpy().reshape(3, n)
    s = pt.default_scene()[:80].reshape(10, 8)
    table = np.stack([np.ones(3, dtype=np.float32) if k == 7 else s[7:10, k] for k in range(8)]) * np.float32(12.0)
    agree = (bits(col.T) == bits(table[want])).all(axis=1).mean()
    assert agree >= 0.999, agree
