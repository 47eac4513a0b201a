"""The fused resolve of the production entries (trace_kernels.cu: fuse_reduce_chunk; resolve_kernels.cu: resolve_means_kernel).

With S a power of two in 8..256 and a constant-bank scene, ptb200_render_image* never materialise per-path colours: the warp
that traced a 256-path chunk averages its sub-pixel runs itself, in NumPy's pairwise order (scripts/data_visualization.py:39-45
-> np.mean on a contiguous float32 axis), and a small second kernel finishes the pixels.  The image must equal, bit for bit,
(a) the oracle's trace + resolve and (b) the two-kernel composition (the default: colours to HBM, resolve kernel).  The fused
composition is opt-in (PTB200_FUSED_RESOLVE=1): it needs 48 bytes of workspace per pixel instead of 12 bytes per path, but it is
not faster on B200 (profiles/r2_fused_resolve.md), so the library does not pick it by itself."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


class fused_mode:
    """Selects the fused (on=True) or the two-kernel composition for the calls inside the block (the library reads the variable
    per call; the two-kernel composition is the default)."""

    def __init__(self, on):
        self.on = on

    def __enter__(self):
        self.old = os.environ.get("PTB200_FUSED_RESOLVE")
        os.environ["PTB200_FUSED_RESOLVE"] = "1" if self.on else "0"

    def __exit__(self, *a):
        if self.old is None:
            del os.environ["PTB200_FUSED_RESOLVE"]
        else:
            os.environ["PTB200_FUSED_RESOLVE"] = self.old


@pytest.fixture(autouse=True)
def fused_by_default():
    """Every call in this module runs the fused composition unless it sits inside `with fused_mode(False)`."""
    with fused_mode(True):
        yield


# S = 8 .. 256 (every supported value), frames whose path count is not a multiple of the 256-path chunk where S allows it,
# depth 5 (regeneration), 3 (lock step), 50 (long paths: a chunk's last path retires late)
@pytest.mark.parametrize("w,h,s,depth,fixed", [(24, 20, 8, 5, False), (17, 9, 16, 5, False), (16, 12, 32, 5, False), (9, 7, 64, 5, False),
                                               (6, 5, 128, 5, False), (5, 3, 256, 5, False), (17, 9, 16, 3, False), (12, 10, 16, 50, False),
                                               (7, 3, 256, 50, False), (16, 12, 32, 5, True), (3, 1, 8, 5, False)])
def test_fused_image_equals_oracle_and_two_kernel_composition(pt, cuda, oracle, w, h, s, depth, fixed):
    torch = cuda
    seed = 1234 + s
    p = pt.default_params(width=w, height=h, samples=s, depth=depth, flags=pt.F_FIXED_DEPTH if fixed else 0)
    n = p.n_paths
    d_sph = dev(torch, pt.default_scene())
    d_stats = torch.zeros(2, dtype=torch.int64, device="cuda")
    d_img = torch.full((h, w, 3), 99, dtype=torch.uint8, device="cuda")
    pt.render_image(p, d_sph, d_img, seed=seed, stats=d_stats)
    rays = oracle.gen_rays_from_uniforms(w, h, s, 0, w, oracle.philox_uniforms(seed, 0, n))
    col, live = oracle.trace(rays, oracle.gen_spheres(), depth=depth, return_live=True)
    want = oracle.resolve(col, w, h, s)
    assert np.array_equal(d_img.cpu().numpy(), want)
    st = d_stats.cpu().numpy()
    assert st[0] == n and st[1] == (live if depth >= 5 and not fixed else n * depth)
    with fused_mode(False):
        d_two = torch.full((h, w, 3), 98, dtype=torch.uint8, device="cuda")
        pt.render_image(p, d_sph, d_two, seed=seed)
    assert np.array_equal(d_two.cpu().numpy(), want)
    # column stripes and strided column sets (the multi-GPU partitions) of the same frame
    for x0, x1 in [(0, 1), (1, w), (w // 3, w // 3 + 2)]:
        if not 0 <= x0 < x1 <= w:
            continue
        d_part = torch.full((h, x1 - x0, 3), 97, dtype=torch.uint8, device="cuda")
        pt.render_image(p, d_sph, d_part, x0=x0, x1=x1, seed=seed)
        assert np.array_equal(d_part.cpu().numpy(), want[:, x0:x1]), (x0, x1)
    if w >= 5:
        ps = pt.default_params(width=w, height=h, samples=s, depth=depth, flags=pt.F_FIXED_DEPTH if fixed else 0, column_step=3)
        for r in range(3):
            cols = list(range(r, w, 3))
            d_part = torch.full((h, len(cols), 3), 96, dtype=torch.uint8, device="cuda")
            pt.render_image(ps, d_sph, d_part, x0=r, x1=w, seed=seed)
            assert np.array_equal(d_part.cpu().numpy(), want[:, cols]), r


def test_fused_with_the_replayed_reference_stream(pt, cuda, oracle):
    """MT19937 replay (the reference's own random stream) through the fused path: the reference's image, S = 8 and 16."""
    torch = cuda
    for w, h, s in [(16, 16, 8), (12, 8, 16)]:
        p = pt.default_params(width=w, height=h, samples=s)
        n = p.n_paths
        u = torch.from_numpy(pt.mt19937_uniforms(0, 2 * n)).cuda()
        d_img = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda")
        pt.render_image(p, dev(torch, pt.default_scene()), d_img, uniforms=u)
        rays = oracle.gen_rays(w, h, s, seed=0)
        want = oracle.resolve(oracle.trace(rays, oracle.gen_spheres()), w, h, s)
        assert np.array_equal(d_img.cpu().numpy(), want)


def test_fused_open_scene_generic_sphere_count(pt, cuda, oracle):
    """The generic (NS = 0) instantiation with an open 12-sphere scene: total misses, NaN / inf colours flow through the fused
    average exactly as through np.mean (NaN means clip like NumPy's: compared against the two-kernel composition)."""
    torch = cuda
    rng = np.random.RandomState(5)
    nsph, stride, w, h, s = 12, 16, 20, 10, 16
    sph = np.zeros((10, stride), dtype=np.float32)
    sph[0, :nsph] = rng.uniform(5, 30, nsph) ** 2
    sph[1:4, :nsph] = rng.uniform(0, 100, (3, nsph))
    sph[7:10, :nsph] = rng.uniform(0, 1, (3, nsph))
    p = pt.default_params(width=w, height=h, samples=s, sphere_count=nsph, sphere_stride=stride, light_index=3)
    d_img = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda")
    pt.render_image(p, dev(torch, sph.reshape(-1)), d_img, seed=77)
    with fused_mode(False):
        d_two = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda")
        pt.render_image(p, dev(torch, sph.reshape(-1)), d_two, seed=77)
    assert np.array_equal(d_img.cpu().numpy(), d_two.cpu().numpy())
    assert d_img.float().std().item() > 1


@pytest.mark.parametrize("s,max_depth", [(8, 10), (16, 64), (64, 12), (256, 8)])
def test_fused_material_kernel_equals_two_kernel_composition(pt, cuda, s, max_depth):
    """DIFF / SPEC / REFR + Russian roulette through the same feeder: fused image == two-kernel image (whose per-path colours are
    pinned against the CPU twin in test_gpu_materials.py), incl. gamma, a stripe and strided columns."""
    torch = cuda
    w, h = 14, 9
    p = pt.default_params(width=w, height=h, samples=s, sphere_count=9, sphere_stride=16)
    mp = pt.default_material_params(seed=5, max_depth=max_depth)
    d_sc = dev(torch, pt.smallpt_scene())
    d_stats = torch.zeros(2, dtype=torch.int64, device="cuda")
    d_img = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda")
    pt.render_image_mat(p, mp, d_sc, d_img, cam_seed=21, gamma=True, stats=d_stats)
    st_f = d_stats.cpu().numpy().copy()
    with fused_mode(False):
        d_two = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda")
        pt.render_image_mat(p, mp, d_sc, d_two, cam_seed=21, gamma=True, stats=d_stats)
    assert np.array_equal(d_img.cpu().numpy(), d_two.cpu().numpy())
    assert np.array_equal(st_f, d_stats.cpu().numpy())
    assert d_img.float().std().item() > 1
    d_part = torch.zeros((h, 4, 3), dtype=torch.uint8, device="cuda")
    pt.render_image_mat(p, mp, d_sc, d_part, x0=3, x1=7, cam_seed=21, gamma=True)
    assert np.array_equal(d_part.cpu().numpy(), d_img.cpu().numpy()[:, 3:7])
    ps = pt.default_params(width=w, height=h, samples=s, sphere_count=9, sphere_stride=16, column_step=2)
    d_part = torch.zeros((h, 7, 3), dtype=torch.uint8, device="cuda")
    pt.render_image_mat(ps, mp, d_sc, d_part, x0=1, x1=w, cam_seed=21, gamma=True)
    assert np.array_equal(d_part.cpu().numpy(), d_img.cpu().numpy()[:, 1::2])


def test_fused_c2_frame_equals_two_kernel_composition(pt, cuda):
    """BASELINE config C2 (1024 x 768, 64 spp: 50 331 648 paths, 196 608 chunks over 5 920 warps) and a 1920 x 270 stripe at
    1024 spp, depth 50 (S = 256: one run per chunk, long paths), fused == two-kernel, incl. the segment counts."""
    torch = cuda
    d_sph = dev(torch, pt.default_scene())
    d_stats = torch.zeros(2, dtype=torch.int64, device="cuda")
    for w, h, s, depth, x0, x1 in [(1024, 768, 16, 5, 0, 1024), (1920, 270, 256, 50, 944, 976)]:
        p = pt.default_params(width=w, height=h, samples=s, depth=depth)
        d_img = torch.zeros((h, x1 - x0, 3), dtype=torch.uint8, device="cuda")
        pt.render_image(p, d_sph, d_img, x0=x0, x1=x1, seed=2024, stats=d_stats)
        st_f = d_stats.cpu().numpy().copy()
        with fused_mode(False):
            d_two = torch.zeros_like(d_img)
            pt.render_image(p, d_sph, d_two, x0=x0, x1=x1, seed=2024, stats=d_stats)
        assert torch.equal(d_img, d_two)
        assert np.array_equal(st_f, d_stats.cpu().numpy())
        assert d_img.float().std().item() > 1
