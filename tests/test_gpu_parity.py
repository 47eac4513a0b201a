"""GPU parity tests proper: the CUDA path, called through the C ABI, against the oracle and the golden vectors.

Bar: bit-exact float32 colours, bit-exact uint8 images, bit-exact rays (integer/byte comparisons on the raw
bits).  Nothing here reads /root/reference."""
import hashlib
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def gpu_trace(pt, torch, p, rays, spheres, first=0, count=-1, prefill=None):
    n = p.n_paths
    d_rays, d_sph = dev(torch, rays.reshape(-1)), dev(torch, spheres)
    d_col = torch.full((3 * n,), float("nan") if prefill is None else prefill, dtype=torch.float32, device="cuda")
    pt.render_do_ex(p, d_rays, d_sph, d_col, first=first, count=count)
    torch.cuda.synchronize()
    return d_col.cpu().numpy().reshape(3, n)


# ---- the reference's own known answers ---------------------------------------------------------------

def test_c1_golden_through_legacy_entry_points(pt, cuda, golden_dir, tmp_path):
    """rays.bin / spheres.bin made by the reference's gen_data.py -> render_do -> colour.bin made by the
    reference's kernel; then resolve + PPM against data_visualization.py's output."""
    torch = cuda
    rays = np.fromfile(os.path.join(golden_dir, "w16h16s1d5_rays.bin"), dtype=np.float32)
    sph = np.fromfile(os.path.join(golden_dir, "w16h16s1d5_spheres.bin"), dtype=np.float32)
    want = np.fromfile(os.path.join(golden_dir, "w16h16s1d5_color.bin"), dtype=np.float32)
    pt.set_legacy_config(pt.default_params())
    d_rays, d_sph = dev(torch, rays), dev(torch, sph)
    for entry in ("render", "render_do"):
        d_col = torch.full((3 * 1024,), float("nan"), dtype=torch.float32, device="cuda")
        if entry == "render":
            pt.render(d_rays, d_sph, d_col)
        else:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            pt.render_do(8, None, s, d_rays, d_sph, d_col)
            s.synchronize()
        got = d_col.cpu().numpy()
        assert np.array_equal(bits(got), bits(want)), entry
    p = pt.default_params()
    d_img = torch.zeros((16, 16, 3), dtype=torch.uint8, device="cuda")
    pt.resolve(p, d_col, d_img)
    img = d_img.cpu().numpy()
    ref_img = np.fromfile(os.path.join(golden_dir, "w16h16s1d5_image_u8.bin"), dtype=np.uint8).reshape(16, 16, 3)
    assert np.array_equal(img, ref_img)
    out = tmp_path / "color.ppm"
    pt.write_ppm(str(out), img)
    assert out.read_text() == open(os.path.join(golden_dir, "w16h16s1d5_color.ppm")).read()


@pytest.mark.parametrize("name", ["w64h64s1d5", "w64h64s4d5", "w64h64s1d10", "w64h64s1d50"])
def test_golden_digests_full_device_pipeline(pt, cuda, manifest, golden_dir, name):
    """MT19937 replay -> device ray generation -> trace -> resolve, each stage against the digest of what the
    reference's own scripts and kernel produced."""
    torch = cuda
    m = manifest[name]
    p = pt.default_params(width=m["w"], height=m["h"], samples=m["s"], depth=m["depth"])
    n = p.n_paths
    u = dev(torch, pt.mt19937_uniforms(0, 2 * n))
    d_rays = torch.empty(6 * n, dtype=torch.float32, device="cuda")
    pt.gen_rays(p, d_rays, uniforms=u)
    assert sha(d_rays.cpu().numpy()) == m["rays_sha256"]
    d_sph = dev(torch, pt.default_scene())
    d_col = torch.empty(3 * n, dtype=torch.float32, device="cuda")
    pt.render_do_ex(p, d_rays, d_sph, d_col)
    assert sha(d_col.cpu().numpy()) == m["color_sha256"]
    d_img = torch.zeros((m["h"], m["w"], 3), dtype=torch.uint8, device="cuda")
    pt.resolve(p, d_col, d_img)
    ref = np.fromfile(os.path.join(golden_dir, f"{name}_image_u8.bin"), dtype=np.uint8).reshape(m["h"], m["w"], 3)
    assert np.array_equal(d_img.cpu().numpy(), ref)
    # and the fused production entry gives the same image from the same stream
    d_img2 = torch.zeros_like(d_img)
    pt.render_image(p, d_sph, d_img2, uniforms=u)
    assert np.array_equal(d_img2.cpu().numpy(), ref)


# ---- against the oracle on seeded inputs ----------------------------------------------------------------

@pytest.mark.parametrize("w,h,s,depth", [(16, 16, 1, 5), (64, 48, 2, 5), (256, 256, 1, 5), (33, 17, 3, 7), (64, 64, 1, 50), (3, 5, 1, 1), (128, 96, 2, 10),
                                             (64, 48, 2, 8), (256, 256, 1, 12)])
@pytest.mark.parametrize("fixed_depth", [False, True])
def test_trace_bit_exact_vs_oracle(pt, cuda, oracle, w, h, s, depth, fixed_depth):
    p = pt.default_params(width=w, height=h, samples=s, depth=depth, flags=1 if fixed_depth else 0)
    rays = oracle.gen_rays(w, h, s, seed=0)
    sph = oracle.gen_spheres()
    got = gpu_trace(pt, cuda, p, rays, sph)
    want = oracle.trace(rays, sph, depth=depth)
    assert np.array_equal(bits(got), bits(want))


def test_trace_slices_like_reference_cores(pt, cuda, oracle):
    """The reference gives core b the slice [b*N/8, (b+1)*N/8) (src/render.cpp:24-27); slices must compose and
    must not write outside themselves."""
    w, h, s = 64, 64, 1
    p = pt.default_params(width=w, height=h, samples=s)
    n = p.n_paths
    rays, sph = oracle.gen_rays(w, h, s, seed=0), oracle.gen_spheres()
    want = oracle.trace(rays, sph)
    torch = cuda
    d_rays, d_sph = dev(torch, rays.reshape(-1)), dev(torch, sph)
    d_col = torch.full((3 * n,), -7.0, dtype=torch.float32, device="cuda")
    pt.render_do_ex(p, d_rays, d_sph, d_col, first=2 * n // 8, count=n // 8)
    torch.cuda.synchronize()
    got = d_col.cpu().numpy().reshape(3, n)
    sl = slice(2 * n // 8, 3 * n // 8)
    assert np.array_equal(bits(got[:, sl]), bits(want[:, sl]))
    mask = np.ones(n, dtype=bool)
    mask[sl] = False
    assert np.all(got[:, mask] == -7.0)
    for b in range(8):
        pt.render_do_ex(p, d_rays, d_sph, d_col, first=b * n // 8, count=n // 8)
    pt.render_do_ex(p, d_rays, d_sph, d_col, first=5, count=0)  # empty slice is a no-op
    torch.cuda.synchronize()
    assert np.array_equal(bits(d_col.cpu().numpy().reshape(3, n)), bits(want))


def _random_scene(rng, nsph, stride, signed_colours=False):
    sph = np.zeros(10 * stride, dtype=np.float32)
    r = rng.uniform(3, 25, nsph)
    sph[0 * stride:0 * stride + nsph] = (r * r).astype(np.float32)
    sph[1 * stride:1 * stride + nsph] = rng.uniform(0, 100, nsph)
    sph[2 * stride:2 * stride + nsph] = rng.uniform(0, 80, nsph)
    sph[3 * stride:3 * stride + nsph] = rng.uniform(0, 170, nsph)
    lo = -1.0 if signed_colours else 0.0
    for m in (7, 8, 9):
        sph[m * stride:m * stride + nsph] = rng.uniform(lo, 1, nsph)
    sph[7 * stride + nsph // 2] = 0.0   # one black sphere: exercises the zero-throughput stop
    sph[8 * stride + nsph // 2] = 0.0
    sph[9 * stride + nsph // 2] = 0.0
    return sph


@pytest.mark.parametrize("nsph,stride", [(8, 8), (1, 1), (5, 8), (12, 16), (100, 100)])
@pytest.mark.parametrize("signed_colours", [False, True])
def test_open_random_scenes_generic_sphere_count(pt, cuda, oracle, nsph, stride, signed_colours):
    """Open scenes: most rays miss everything (index 0 / t = 1e20 path), NaN and inf flow through the bounce
    arithmetic; other sphere counts take the run-time-count kernel; signed colours switch the zero stop off."""
    rng = np.random.default_rng(100 * nsph + stride + signed_colours)
    w, h, s, depth = 32, 32, 1, 6
    sph = _random_scene(rng, nsph, stride, signed_colours)
    light = nsph - 1
    p = pt.default_params(width=w, height=h, samples=s, depth=depth, sphere_count=nsph, sphere_stride=stride, light_index=light)
    n = p.n_paths
    o = np.stack([rng.uniform(0, 100, n), rng.uniform(0, 80, n), rng.uniform(0, 170, n)])
    d = rng.normal(size=(3, n))
    d /= np.linalg.norm(d, axis=0)
    rays = np.concatenate([o, d]).astype(np.float32)
    want = oracle.trace(rays, sph, depth=depth, nsph=nsph, stride=stride, light=light)
    for flags in (0, 1):
        p.flags = flags
        got = gpu_trace(pt, cuda, p, rays, sph)
        assert np.array_equal(bits(got), bits(want)), flags


def test_special_values_in_rays(pt, cuda, oracle):
    """Zero directions, huge origins, NaN/inf inputs: same bits out as the reference semantics give."""
    w, h, s = 16, 16, 1
    p = pt.default_params(width=w, height=h, samples=s)
    rays = oracle.gen_rays(w, h, s, seed=0)
    specials = [0.0, -0.0, np.inf, -np.inf, np.nan, 1e30, -1e30, 1e-40, 3.4e38]
    rng = np.random.default_rng(5)
    for j, v in enumerate(specials):
        for c in range(6):
            rays[c, rng.integers(0, rays.shape[1])] = v
    sph = oracle.gen_spheres()
    got = gpu_trace(pt, cuda, p, rays, sph)
    want = oracle.trace(rays, sph)
    assert np.array_equal(bits(got), bits(want))


# ---- ray generation and resolve ------------------------------------------------------------------------

@pytest.mark.parametrize("w,h,s", [(16, 16, 1), (40, 24, 3), (128, 128, 1), (517, 389, 2), (1000, 30, 5)])
def test_gen_rays_mt_replay_bit_exact(pt, cuda, oracle, w, h, s):
    torch = cuda
    p = pt.default_params(width=w, height=h, samples=s)
    n = p.n_paths
    u = pt.mt19937_uniforms(0, 2 * n)
    d_rays = torch.empty(6 * n, dtype=torch.float32, device="cuda")
    pt.gen_rays(p, d_rays, uniforms=dev(torch, u))
    want = oracle.gen_rays(w, h, s, seed=0)
    assert np.array_equal(bits(d_rays.cpu().numpy().reshape(6, n)), bits(want))
    # a column range generated on its own (what one GPU of a multi-GPU job does) equals that slice
    x0, x1 = w // 4, w // 2 + 1
    per_col = h * 4 * s
    m = (x1 - x0) * per_col
    d_part = torch.empty(6 * m, dtype=torch.float32, device="cuda")
    pt.gen_rays(p, d_part, x0=x0, x1=x1, uniforms=dev(torch, u[2 * x0 * per_col:2 * x1 * per_col]))
    assert np.array_equal(bits(d_part.cpu().numpy().reshape(6, m)), bits(want[:, x0 * per_col:x1 * per_col]))


def test_gen_rays_counter_based_rng(pt, cuda, oracle):
    torch = cuda
    w, h, s, seed = 48, 32, 2, 0x1234_5678_9abc_def0
    p = pt.default_params(width=w, height=h, samples=s)
    n = p.n_paths
    d_rays = torch.empty(6 * n, dtype=torch.float32, device="cuda")
    pt.gen_rays(p, d_rays, seed=seed)
    want = oracle.gen_rays_from_uniforms(w, h, s, 0, w, oracle.philox_uniforms(seed, 0, n))
    assert np.array_equal(bits(d_rays.cpu().numpy().reshape(6, n)), bits(want))
    # column sub-range: the counter is the GLOBAL path index, so no stream hand-off is needed
    x0, x1 = 10, 13
    per_col = h * 4 * s
    d_part = torch.empty(6 * (x1 - x0) * per_col, dtype=torch.float32, device="cuda")
    pt.gen_rays(p, d_part, x0=x0, x1=x1, seed=seed)
    assert np.array_equal(bits(d_part.cpu().numpy().reshape(6, -1)), bits(want[:, x0 * per_col:x1 * per_col]))


@pytest.mark.parametrize("w,h,s", [(16, 16, 1), (16, 8, 2), (8, 16, 5), (8, 8, 8), (8, 8, 16), (8, 8, 17), (4, 4, 128), (4, 4, 129), (4, 4, 256), (2, 2, 1000)])
def test_resolve_bit_exact(pt, cuda, oracle, w, h, s):
    torch = cuda
    rng = np.random.default_rng(w * 1000 + s)
    p = pt.default_params(width=w, height=h, samples=s)
    col = (rng.random(3 * p.n_paths) * 1.4 - 0.1).astype(np.float32)   # below 0 and above 1: exercises the clip
    d_img = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda")
    pt.resolve(p, dev(torch, col), d_img)
    assert np.array_equal(d_img.cpu().numpy(), oracle.resolve(col, w, h, s))
    if w >= 4:
        x0, x1 = 1, w - 1
        d_part = torch.zeros((h, x1 - x0, 3), dtype=torch.uint8, device="cuda")
        pt.resolve(p, dev(torch, col), d_part, x0=x0, x1=x1)
        assert np.array_equal(d_part.cpu().numpy(), oracle.resolve(col, w, h, s)[:, x0:x1])


@pytest.mark.parametrize("w,h,s", [(48, 32, 8), (64, 40, 16), (37, 50, 8), (32, 16, 256), (80, 17, 24)])
def test_resolve_tiled_kernel_whole_and_ragged_tiles(pt, cuda, oracle, w, h, s):
    """S % 8 == 0 takes the tiled resolve kernel (128-bit loads, 128-bit row stores): frames with whole 16 x 16 tiles, ragged
    right / bottom edges, rows that are not 16-byte aligned (byte-store fallback) and stripes starting on / off a tile column."""
    torch = cuda
    rng = np.random.default_rng(w * 1000 + h * 10 + s)
    p = pt.default_params(width=w, height=h, samples=s)
    col = (rng.random(3 * p.n_paths) * 1.4 - 0.1).astype(np.float32)
    want = oracle.resolve(col, w, h, s)
    d_col = dev(torch, col)
    d_img = torch.full((h, w, 3), 7, dtype=torch.uint8, device="cuda")
    pt.resolve(p, d_col, d_img)
    assert np.array_equal(d_img.cpu().numpy(), want)
    for x0, x1 in [(16, w), (0, 32), (5, w - 3), (16, 32)]:
        if not 0 <= x0 < x1 <= w:
            continue
        d_part = torch.full((h, x1 - x0, 3), 7, dtype=torch.uint8, device="cuda")
        pt.resolve(p, d_col, d_part, x0=x0, x1=x1)
        assert np.array_equal(d_part.cpu().numpy(), want[:, x0:x1]), (x0, x1)


def test_small_tiles_split_the_frame_mid_column(pt, cuda, tmp_path):
    """The production entry renders a frame in tiles of whole pixels; with the default 512 Mi-path tiles no test frame is ever
    split, so a child process renders with PTB200_TILE_PATHS small enough to start tiles in the middle of image columns
    (pixel ranges the tiled resolve kernel has to clip) and must produce the frame of the unsplit render, all three kernels."""
    torch = cuda
    import subprocess
    import sys
    w, h, s = 40, 36, 8
    script = f"""
import sys, numpy as np, torch
sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})
import ascendpathtracing_b200 as pt
w, h, s = {w}, {h}, {s}
out = []
img = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda")
p = pt.default_params(width=w, height=h, samples=s)
pt.render_image(p, torch.from_numpy(pt.default_scene()).cuda(), img, seed=3); out.append(img.cpu().numpy().copy())
pm = pt.default_params(width=w, height=h, samples=s, sphere_count=9, sphere_stride=16)
mp = pt.default_material_params(seed=2, max_depth=10)
pt.render_image_mat(pm, mp, torch.from_numpy(pt.smallpt_scene()).cuda(), img, cam_seed=4); out.append(img.cpu().numpy().copy())
bvh = pt.Bvh(torch.from_numpy(pt.random_scene(300)).cuda(), 307, 307)
pt.render_image_mat_bvh(p, mp, bvh, img, cam_seed=4); out.append(img.cpu().numpy().copy())
np.save(sys.argv[1], np.stack(out))
"""
    frames = {}
    for name, tile in (("whole", ""), ("split", str(267 * 4 * s))):  # a tile = 267 pixels: not a multiple of h = 36
        env = dict(os.environ)
        if tile:
            env["PTB200_TILE_PATHS"] = tile
        out = tmp_path / f"{name}.npy"
        subprocess.run([sys.executable, "-c", script, str(out)], check=True, env=env, timeout=600)
        frames[name] = np.load(out)
    assert np.array_equal(frames["whole"], frames["split"])
    assert frames["whole"].std() > 1


# ---- whole-job entries -------------------------------------------------------------------------------

def test_render_host_entry(pt, cuda, oracle):
    torch = cuda
    w, h, s = 96, 64, 2
    p = pt.default_params(width=w, height=h, samples=s)
    rays, sph = oracle.gen_rays(w, h, s, seed=0), oracle.gen_spheres()
    out = np.zeros(3 * p.n_paths, dtype=np.float32)
    pt.render_host(p, np.ascontiguousarray(rays.reshape(-1)), sph, out)
    assert np.array_equal(bits(out.reshape(3, -1)), bits(oracle.trace(rays, sph)))
    # pinned host memory, as bench.py uses it
    h_rays = torch.from_numpy(rays.reshape(-1).copy()).pin_memory()
    h_out = torch.zeros(3 * p.n_paths, dtype=torch.float32).pin_memory()
    pt.render_host(p, h_rays, sph, h_out)
    assert np.array_equal(bits(h_out.numpy().reshape(3, -1)), bits(oracle.trace(rays, sph)))


def test_render_image_counter_rng_and_stats(pt, cuda, oracle):
    torch = cuda
    w, h, s, seed = 40, 24, 4, 99
    d_sph = dev(torch, pt.default_scene())
    d_stats = torch.zeros(2, dtype=torch.int64, device="cuda")
    # depth >= 5: regeneration with exact early termination -> segments traced == the oracle's count of live segments;
    # below, the library runs the (bit-identical) lock-step loop because it is faster there -> every segment is traced
    for depth in (10, 5, 3):
        p = pt.default_params(width=w, height=h, samples=s, depth=depth)
        n = p.n_paths
        d_img = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda")
        pt.render_image(p, d_sph, d_img, seed=seed, stats=d_stats)
        rays = oracle.gen_rays_from_uniforms(w, h, s, 0, w, oracle.philox_uniforms(seed, 0, n))
        col, live = oracle.trace(rays, oracle.gen_spheres(), depth=depth, return_live=True)
        assert np.array_equal(d_img.cpu().numpy(), oracle.resolve(col, w, h, s))
        stats = d_stats.cpu().numpy()
        assert stats[0] == n
        assert stats[1] == (live if depth >= 5 else n * depth)
    # column stripes (the multi-GPU partition) tile the same image
    for x0, x1 in [(0, 7), (7, 25), (25, 40)]:
        d_part = torch.zeros((h, x1 - x0, 3), dtype=torch.uint8, device="cuda")
        pt.render_image(p, d_sph, d_part, x0=x0, x1=x1, seed=seed)
        assert np.array_equal(d_part.cpu().numpy(), d_img.cpu().numpy()[:, x0:x1])


def test_concurrent_streams_with_different_scenes(pt, cuda, oracle):
    """The constant-bank scene is a per-device singleton guarded by an event: two streams rendering different
    scenes back to back must not see each other's spheres."""
    torch = cuda
    w, h, s = 128, 128, 1
    p = pt.default_params(width=w, height=h, samples=s)
    rays = oracle.gen_rays(w, h, s, seed=0)
    sph_a = oracle.gen_spheres()
    sph_b = sph_a.copy()
    sph_b[7 * 8 + 0] = 0.9   # recolour the left wall
    sph_b[1 * 8 + 6] = 60.0  # move the mirror ball
    want_a, want_b = oracle.trace(rays, sph_a), oracle.trace(rays, sph_b)
    d_rays = dev(torch, rays.reshape(-1))
    d_a, d_b = dev(torch, sph_a), dev(torch, sph_b)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    outs = []
    for k in range(6):
        st, sp = (s1, d_a) if k % 2 == 0 else (s2, d_b)
        d_col = torch.empty(3 * p.n_paths, dtype=torch.float32, device="cuda")
        pt.render_do_ex(p, d_rays, sp, d_col, stream=st)
        outs.append(d_col)
    torch.cuda.synchronize()
    for k, d_col in enumerate(outs):
        want = want_a if k % 2 == 0 else want_b
        assert np.array_equal(bits(d_col.cpu().numpy().reshape(3, -1)), bits(want)), k


def test_arena_on_device(pt, cuda):
    torch = cuda
    a = pt.Arena(8 << 20)
    p1, p2 = a.alloc(1 << 20), a.alloc(3 << 20)
    assert p2 == p1 + (1 << 20) and a.in_use == 4 << 20
    # the memory is real device memory: run a render out of it
    p = pt.default_params()
    n = p.n_paths
    rays = pt.mt19937_uniforms(0, 1)  # noqa: F841  (touch the host helper)
    a.free(p1)
    a.free(p2)
    assert a.largest_free == a.capacity
    with pytest.raises(pt.PtError):
        a.alloc(16 << 20)
    a.close()


# ---- full-size, size-independent properties -----------------------------------------------------------

def test_c2_full_size_properties(pt, cuda, oracle):
    """BASELINE config C2 (1024x768, 64 spp = 50 331 648 paths), rays from the counter-based RNG on the device:
    (1) early termination and fixed depth agree bit for bit over the whole buffer;
    (2) a strided sample of 100 k paths equals the oracle;
    (3) the frame tiled into three column stripes equals the frame rendered in one go."""
    torch = cuda
    w, h, s, seed = 1024, 768, 16, 2024
    p = pt.default_params(width=w, height=h, samples=s)
    n = p.n_paths
    assert n == 50331648
    d_rays = torch.empty(6 * n, dtype=torch.float32, device="cuda")
    pt.gen_rays(p, d_rays, seed=seed)
    d_sph = dev(torch, pt.default_scene())
    d_a = torch.empty(3 * n, dtype=torch.float32, device="cuda")
    d_b = torch.empty(3 * n, dtype=torch.float32, device="cuda")
    # (1) at depth 10 (regeneration with early termination unless told otherwise), then at the reference's depth 5
    p.depth = 10
    pt.render_do_ex(p, d_rays, d_sph, d_a)
    p.flags = 1
    pt.render_do_ex(p, d_rays, d_sph, d_b)
    p.flags = 0
    torch.cuda.synchronize()
    assert torch.equal(d_a.view(torch.int32), d_b.view(torch.int32))
    idx10 = torch.arange(7, n, 2003, device="cuda")
    assert np.array_equal(bits(d_a.view(3, n)[:, idx10].cpu().numpy()),
                          bits(oracle.trace(d_rays.view(6, n)[:, idx10].cpu().numpy(), oracle.gen_spheres(), depth=10)))
    p.depth = 5
    pt.render_do_ex(p, d_rays, d_sph, d_a)
    p.flags = 1
    pt.render_do_ex(p, d_rays, d_sph, d_b)
    p.flags = 0
    torch.cuda.synchronize()
    assert torch.equal(d_a.view(torch.int32), d_b.view(torch.int32))
    idx = torch.arange(0, n, 503, device="cuda")
    rays_s = d_rays.view(6, n)[:, idx].cpu().numpy()
    got = d_a.view(3, n)[:, idx].cpu().numpy()
    assert np.array_equal(bits(got), bits(oracle.trace(rays_s, oracle.gen_spheres())))
    d_img = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda")
    pt.resolve(p, d_a, d_img)
    d_img2 = torch.zeros_like(d_img)
    for x0, x1 in [(0, 300), (300, 301), (301, 1024)]:
        d_part = torch.zeros((h, x1 - x0, 3), dtype=torch.uint8, device="cuda")
        pt.render_image(p, d_sph, d_part, x0=x0, x1=x1, seed=seed)
        d_img2[:, x0:x1] = d_part
    assert torch.equal(d_img, d_img2)
    # the image is a picture of the Cornell box, not noise: the light column is saturated, the floor is grey-ish
    img = d_img.cpu().numpy()
    assert img.mean() > 20 and img.std() > 20


def test_independent_seeds_converge(pt, cuda, oracle):
    """north_star: converge to the same image under independent seeds.  64x64, exact arithmetic: RMSE between
    renders with different seeds follows ~61/sqrt(spp/4) in 8-bit units (18.3 at 64 spp,
    9.4 at 256 spp).  Reference stream = MT19937 seed 0 through the oracle; ours = counter-based seed 7."""
    torch = cuda
    w = h = 64
    d_sph = dev(torch, pt.default_scene())
    rmse = {}
    for s in (16, 64):
        p = pt.default_params(width=w, height=h, samples=s)
        d_img = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda")
        pt.render_image(p, d_sph, d_img, seed=7)
        ref = oracle.resolve(oracle.trace(oracle.gen_rays(w, h, s, seed=0), oracle.gen_spheres()), w, h, s)
        diff = d_img.cpu().numpy().astype(np.float64) - ref.astype(np.float64)
        rmse[s] = float(np.sqrt((diff ** 2).mean()))
    assert 14.0 < rmse[16] < 23.0, rmse      # expected 18.3
    assert 7.0 < rmse[64] < 12.0, rmse       # expected 9.4
    assert rmse[64] < 0.62 * rmse[16]        # halves when spp quadruples


# ---- the drop-in host and the largest configuration ---------------------------------------------------

def test_host_binary_reproduces_reference_files(pt, cuda, golden_dir, tmp_path):
    """render_gpu = the reference's src/main.cpp flow on CUDA: --gen replays gen_data.py on the device, the kernel entry is
    render_do, --ppm replaces data_visualization.py.  All four files must equal what the reference's own pipeline wrote."""
    import subprocess
    from ascendpathtracing_b200.host import build as host_build
    exe = host_build.build()
    (tmp_path / "input").mkdir()
    (tmp_path / "output").mkdir()
    out = subprocess.run([exe, "--gen", "--ppm"], cwd=tmp_path, capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    for ours, ref in [("input/rays.bin", "w16h16s1d5_rays.bin"), ("input/spheres.bin", "w16h16s1d5_spheres.bin"),
                      ("output/color.bin", "w16h16s1d5_color.bin"), ("output/color.ppm", "w16h16s1d5_color.ppm")]:
        assert (tmp_path / ours).read_bytes() == open(os.path.join(golden_dir, ref), "rb").read(), ours
    # files written by someone else (here: the golden inputs) are read like the reference reads them
    out = subprocess.run([exe, "--ppm"], cwd=tmp_path, capture_output=True, text=True)
    assert out.returncode == 0
    assert (tmp_path / "output/color.bin").read_bytes() == open(os.path.join(golden_dir, "w16h16s1d5_color.bin"), "rb").read()
    # a size outside the reference's tiling rule goes through the run-time entry
    out = subprocess.run([exe, "--gen", "--ppm", "--width", "10", "--height", "6", "--samples", "3"], cwd=tmp_path, capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert (tmp_path / "output/color.ppm").read_text().startswith("P3\n10 6\n255\n")
    # missing input -> failure exit code, like run.sh expects (run.sh:124-127)
    os.remove(tmp_path / "input/rays.bin")
    out = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True)
    assert out.returncode != 0


def test_c3_full_frame_spot_checks(pt, cuda, oracle):
    """BASELINE config C3: 3840x2160 at 1024 spp = 8 493 465 600 paths (more than the reference's int32 ray count can
    hold), rendered through the production entry with the counter-based RNG.  The full-frame oracle is infeasible, so
    16x16-pixel windows are checked bit for bit: the oracle generates that window's rays from the same global path
    indices (inputs C3)."""
    torch = cuda
    w, h, s, seed = 3840, 2160, 256, 77
    p = pt.default_params(width=w, height=h, samples=s)
    d_sph = dev(torch, pt.default_scene())
    d_stats = torch.zeros(2, dtype=torch.int64, device="cuda")
    # one stripe of 480 columns = what one of 8 GPUs renders; then windows inside it
    x0, x1 = 1920, 2400
    d_img = torch.zeros((h, x1 - x0, 3), dtype=torch.uint8, device="cuda")
    pt.render_image(p, d_sph, d_img, x0=x0, x1=x1, seed=seed, stats=d_stats)
    img = d_img.cpu().numpy()
    n_stripe = (x1 - x0) * h * 4 * s
    assert int(d_stats[0]) == n_stripe == 1061683200
    per_col = h * 4 * s
    sph = oracle.gen_spheres()
    for (wx, wy) in [(1920, 0), (2100, 1000), (2399 - 15, 2160 - 16)]:
        win = np.zeros((16, 16, 3), dtype=np.uint8)
        for cx in range(16):
            x = wx + cx
            first = x * per_col + wy * 4 * s                 # global path index of pixel (x, wy), sample 0
            m = 16 * 4 * s                                    # 16 consecutive pixels of this column
            u = oracle.philox_uniforms(seed, first, m)
            # rays of pixels (x, wy .. wy+15): the camera formula needs the true (x, y), so generate the whole column
            # slice through the oracle's column generator and cut the rows out
            ucol = np.zeros(2 * per_col)
            ucol[2 * wy * 4 * s:2 * (wy * 4 * s + m)] = u
            rays = oracle.gen_rays_from_uniforms(w, h, s, x, x + 1, ucol)[:, wy * 4 * s:wy * 4 * s + m]
            col = oracle.trace(rays, sph)
            colimg = oracle.resolve(col, 1, 16, s)            # 16 pixels as a 1-column image: row r = y index 15-r
            win[:, cx] = colimg[:, 0]
        got = img[h - 1 - (wy + 15):h - 1 - (wy + 15) + 16, wx - x0:wx - x0 + 16]
        assert np.array_equal(got, win), (wx, wy)


def test_run_sh_gpu_mode_end_to_end(pt, cuda, golden_dir):
    """`bash run.sh -r gpu` = the reference's run.sh flow (build, generate, run, visualise) on the B200 build; its
    output/color.ppm and output/color.bin must be the files the reference's own pipeline produces for the default size."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    try:
        out = subprocess.run(["bash", os.path.join(root, "run.sh"), "-r", "gpu", "-v", "Ascend310P1"], capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
        assert "execute op on gpu succeed" in out.stdout
        assert open(os.path.join(root, "output/color.ppm")).read() == open(os.path.join(golden_dir, "w16h16s1d5_color.ppm")).read()
        assert open(os.path.join(root, "output/color.bin"), "rb").read() == open(os.path.join(golden_dir, "w16h16s1d5_color.bin"), "rb").read()
        bad = subprocess.run(["bash", os.path.join(root, "run.sh"), "-r", "sim"], capture_output=True, text=True, timeout=60)
        assert bad.returncode != 0
    finally:
        shutil.rmtree(os.path.join(root, "input"), ignore_errors=True)
        shutil.rmtree(os.path.join(root, "output"), ignore_errors=True)


@pytest.mark.parametrize("w,h,s,world", [(37, 20, 2, 3), (64, 16, 8, 8), (5, 7, 1, 8)])
def test_column_step_renders_exactly_those_columns_of_the_frame(pt, cuda, w, h, s, world):
    """PtParams.column_step (multi-GPU: rank r renders columns r, r+G, ... in one launch): every entry of the image family
    must reproduce the whole frame's columns bit for bit -- camera rays see the image column, RNG keys are global path indices."""
    torch = cuda
    import numpy as np
    from ascendpathtracing_b200 import sharding

    def frames(render_full, render_cols):
        full = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda")
        render_full(full)
        got = torch.full((h, w, 3), 255, dtype=torch.uint8, device="cuda")
        for r in range(world):
            x0, step, n = sharding.strided_columns(w, r, world)
            if n == 0:
                continue
            part = torch.zeros((h, n, 3), dtype=torch.uint8, device="cuda")
            render_cols(part, x0, step)
            got[:, x0::step] = part
        torch.cuda.synchronize()
        return full.cpu().numpy(), got.cpu().numpy()

    # reference-parity kernel
    d_sc = torch.from_numpy(pt.default_scene()).cuda()
    p = pt.default_params(width=w, height=h, samples=s)
    ps = pt.default_params(width=w, height=h, samples=s, column_step=world)
    a, b = frames(lambda img: pt.render_image(p, d_sc, img, seed=11), lambda img, x0, step: pt.render_image(ps, d_sc, img, x0=x0, x1=w, seed=11))
    assert np.array_equal(a, b)
    # materials, constant-bank scene
    d_sm = torch.from_numpy(pt.smallpt_scene()).cuda()
    pm = pt.default_params(width=w, height=h, samples=s, sphere_count=9, sphere_stride=16)
    pms = pt.default_params(width=w, height=h, samples=s, sphere_count=9, sphere_stride=16, column_step=world)
    mp = pt.default_material_params(seed=5, max_depth=12)
    a, b = frames(lambda img: pt.render_image_mat(pm, mp, d_sm, img, cam_seed=3, gamma=True),
                  lambda img, x0, step: pt.render_image_mat(pms, mp, d_sm, img, x0=x0, x1=w, cam_seed=3, gamma=True))
    assert np.array_equal(a, b)
    # materials through the BVH
    nsph = 7 + 500
    d_big = torch.from_numpy(pt.random_scene(500)).cuda()
    bvh = pt.Bvh(d_big, nsph, nsph)
    a, b = frames(lambda img: pt.render_image_mat_bvh(p, mp, bvh, img, cam_seed=3),
                  lambda img, x0, step: pt.render_image_mat_bvh(ps, mp, bvh, img, x0=x0, x1=w, cam_seed=3))
    assert np.array_equal(a, b)
    bvh.close()
