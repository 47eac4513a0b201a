"""One process, several GPUs, through the C ABI only (ptb200_render_image_multi / ptb200_render_host_multi, the C++ host
binary's --gpus / --image modes) -- no torch.distributed anywhere on this path.  SURVEY.md 8e; the reference's counterpart
is its kernel's 8-way slice split (src/render.cpp:9-10,24-27) and its device host flow (src/main.cpp:46-92).

A box with one GPU still exercises everything but the NVLink hop: a device may be listed several times, its column sets
are then rendered by concurrent host threads on that device and assembled by the same peer-copy + interleave code."""
import json
import os
import subprocess
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def device_lists(pt, world):
    """[0]*world always; the first `world` real devices as well when the box has them."""
    out = [[0] * world]
    if pt.device_count() >= world > 1:
        out.append(list(range(world)))
    return out


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_image_multi_equals_single_device_frame(pt, cuda, oracle, world):
    """Reference-parity kernel: the frame assembled from `world` strided column sets is bit for bit the frame of one device,
    which is bit for bit the oracle's."""
    torch = cuda
    w, h, s, seed = 50, 24, 2, 21
    p = pt.default_params(width=w, height=h, samples=s)
    scene = pt.default_scene()
    rays = oracle.gen_rays_from_uniforms(w, h, s, 0, w, oracle.philox_uniforms(seed, 0, p.n_paths))
    col, live = oracle.trace(rays, oracle.gen_spheres(), return_live=True)
    want = oracle.resolve(col, w, h, s)
    for devices in device_lists(pt, world):
        img = np.zeros((h, w, 3), dtype=np.uint8)                      # pageable host memory
        stats, ms = pt.render_image_multi(p, devices, scene, img, seed=seed)
        assert np.array_equal(img, want), devices
        assert stats == [p.n_paths, live]
        assert len(ms) == 1 + world and all(m > 0 for m in ms)
        d_img = torch.zeros((h, w, 3), dtype=torch.uint8, device=f"cuda:{devices[0]}")   # memory of devices[0]
        pt.render_image_multi(p, devices, scene, d_img, seed=seed)
        assert np.array_equal(d_img.cpu().numpy(), want), devices


@pytest.mark.parametrize("world", [1, 2, 4])
def test_image_multi_materials_and_bvh(pt, cuda, world):
    """Material kernels (constant-bank scene; 507-sphere scene through the per-device BVH): any device count gives the
    single-device frame -- every random number is keyed by the global path index."""
    torch = cuda
    w, h, s = 36, 20, 2
    mp = pt.default_material_params(seed=5, max_depth=12)
    # smallpt's 9 spheres
    small = pt.smallpt_scene()
    p9 = pt.default_params(width=w, height=h, samples=s, sphere_count=9, sphere_stride=16)
    d_one = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda")
    d_stat = torch.zeros(2, dtype=torch.int64, device="cuda")
    pt.render_image_mat(p9, mp, dev(torch, small), d_one, cam_seed=3, gamma=True, stats=d_stat)
    for devices in device_lists(pt, world):
        img = np.zeros((h, w, 3), dtype=np.uint8)
        stats, _ = pt.render_image_multi(p9, devices, small, img, seed=3, mp=mp, gamma=True)
        assert np.array_equal(img, d_one.cpu().numpy()), devices
        assert stats == [int(d_stat[0]), int(d_stat[1])]
    # 500 random spheres + walls + light through the tree
    nsph = 7 + 500
    big = pt.random_scene(500)
    pb = pt.default_params(width=w, height=h, samples=s, sphere_count=nsph, sphere_stride=nsph)
    tree = pt.Bvh(dev(torch, big), nsph, nsph)
    pt.render_image_mat_bvh(pb, mp, tree, d_one, cam_seed=3)
    tree.close()
    for devices in device_lists(pt, world):
        img = np.zeros((h, w, 3), dtype=np.uint8)
        pt.render_image_multi(pb, devices, big, img, seed=3, mp=mp, use_bvh=True)
        assert np.array_equal(img, d_one.cpu().numpy()), devices


@pytest.mark.parametrize("world", [1, 2, 5])
def test_render_host_multi_slices_like_reference_cores(pt, cuda, oracle, world):
    """Host buffers in and out, the N paths cut into contiguous slices [r*N/g, (r+1)*N/g) like the reference's per-core
    rule (src/render.cpp:24-27): same colours as one device and as the oracle, bit for bit."""
    torch = cuda
    w, h, s = 96, 50, 2   # N = 38 400: not divisible by 5 * 128 -- slices of ragged size
    p = pt.default_params(width=w, height=h, samples=s)
    rays, sph = oracle.gen_rays(w, h, s, seed=0), oracle.gen_spheres()
    want = oracle.trace(rays, sph)
    h_rays = torch.from_numpy(rays.reshape(-1).copy()).pin_memory()
    for devices in device_lists(pt, world):
        h_out = torch.full((3 * p.n_paths,), float("nan"), dtype=torch.float32).pin_memory()
        ms = pt.render_host_multi(p, devices, h_rays, sph, h_out)
        assert np.array_equal(bits(h_out.numpy().reshape(3, -1)), bits(want)), devices
        assert len(ms) == 1 + world


def test_multi_argument_validation(pt, cuda):
    p = pt.default_params(width=8, height=8)
    img = np.zeros((8, 8, 3), dtype=np.uint8)
    with pytest.raises(pt.PtError):
        pt.render_image_multi(p, [pt.device_count()], pt.default_scene(), img)        # no such device
    with pytest.raises(pt.PtError):
        pt.render_image_multi(p, [0] * 17, pt.default_scene(), img)                   # more than 16
    with pytest.raises(pt.PtError):
        pt.render_image_multi(p, [0], pt.default_scene(), img, use_bvh=True)          # a tree needs the material kernel
    with pytest.raises(pt.PtError):
        pt.render_image_multi(pt.default_params(width=2, height=8), [0, 0, 0], pt.default_scene(), img)  # more devices than columns


def test_two_host_threads_share_one_device(pt, cuda, oracle):
    """Two host threads call the production entry on the same device at the same time with jobs of very different size (the
    second needs a far larger workspace than the arena the first one created).  The workspace hands out blocks under one
    lock and adds an arena instead of replacing one that is in use; both frames must be exact, every round."""
    torch = cuda
    jobs = [(40, 24, 4, 99), (256, 192, 16, 7)]
    sph = oracle.gen_spheres()
    want = []
    for (w, h, s, seed) in jobs:
        p = pt.default_params(width=w, height=h, samples=s)
        rays = oracle.gen_rays_from_uniforms(w, h, s, 0, w, oracle.philox_uniforms(seed, 0, p.n_paths))
        want.append(oracle.resolve(oracle.trace(rays, sph), w, h, s))
    d_sph = dev(torch, pt.default_scene())
    torch.cuda.synchronize()
    errors = []

    def run(k, rounds):
        try:
            torch.cuda.set_device(0)
            w, h, s, seed = jobs[k]
            p = pt.default_params(width=w, height=h, samples=s)
            st = torch.cuda.Stream()
            for _ in range(rounds):
                d_img = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda")
                st.wait_stream(torch.cuda.current_stream())
                pt.render_image(p, d_sph, d_img, seed=seed, stream=st)   # synchronous on return
                if not np.array_equal(d_img.cpu().numpy(), want[k]):
                    errors.append(f"job {k}: frame differs")
        except Exception as e:  # noqa: BLE001
            errors.append(f"job {k}: {e}")

    th = [threading.Thread(target=run, args=(0, 12)), threading.Thread(target=run, args=(1, 4))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors, errors


def test_stream_of_another_device_is_refused(pt, cuda):
    torch = cuda
    if pt.device_count() < 2:
        pytest.skip("needs two devices")
    p = pt.default_params()
    with torch.cuda.device(1):
        other = torch.cuda.Stream()
    d = torch.zeros(6 * p.n_paths, dtype=torch.float32, device="cuda:0")
    with pytest.raises(pt.PtError):
        pt.render_do_ex(p, d, dev(torch, pt.default_scene()), torch.zeros(3 * p.n_paths, device="cuda:0"), stream=other)


# ---- the C++ host binary and run.sh ------------------------------------------------------------------------------------

def _run_host(tmp_path, *args):
    from ascendpathtracing_b200.host import build as host_build
    exe = host_build.build()
    (tmp_path / "input").mkdir(exist_ok=True)
    (tmp_path / "output").mkdir(exist_ok=True)
    out = subprocess.run([exe, *args], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "ptb200 error" not in out.stderr and "cudaError" not in out.stderr, out.stderr[-2000:]
    report = [json.loads(ln) for ln in out.stdout.splitlines() if ln.startswith('{"ptb200"')]
    assert len(report) == 1
    return report[0]


def _read_ppm(path):
    raw = open(path, "rb").read()
    if raw[:2] == b"P3":
        vals = raw.decode().split()
        w, h = int(vals[1]), int(vals[2])
        return np.array(vals[4:], dtype=np.int64).astype(np.uint8).reshape(h, w, 3)
    assert raw[:2] == b"P6"
    end = 0
    for _ in range(3):  # "P6\n", "W H\n", "255\n"
        end = raw.index(b"\n", end) + 1
    w, h = (int(v) for v in raw[:end].split()[1:3])
    return np.frombuffer(raw[end:], dtype=np.uint8).reshape(h, w, 3)


def test_host_binary_drop_in_files_over_several_gpus(pt, cuda, golden_dir, tmp_path):
    """render_gpu --gpus N in the reference's file mode: rays.bin / spheres.bin in, color.bin out, the path array cut N ways
    like the reference's cores.  Output files must equal the reference's own, whatever N is; one JSON line reports the run."""
    for n in sorted({1, 2, min(8, max(1, pt.device_count()))}):
        args = ["--gen", "--ppm", "--gpus", str(n)] if n <= pt.device_count() else ["--gen", "--ppm", "--devices", ",".join(["0"] * n)]
        rep = _run_host(tmp_path, *args)
        assert rep["mode"] == "dropin" and rep["gpus"] == n and rep["paths"] == 1024 and rep["segments"] == 5120 and rep["ms"] > 0
        for ours, ref in [("input/rays.bin", "w16h16s1d5_rays.bin"), ("input/spheres.bin", "w16h16s1d5_spheres.bin"),
                          ("output/color.bin", "w16h16s1d5_color.bin"), ("output/color.ppm", "w16h16s1d5_color.ppm")]:
            assert (tmp_path / ours).read_bytes() == open(os.path.join(golden_dir, ref), "rb").read(), (n, ours)


def test_host_binary_image_mode_scene_files_and_materials(pt, cuda, oracle, tmp_path):
    """--image: scene file in, PPM out, nothing per-path leaves the GPU.  (a) the reference's 512-byte scene through the parity
    kernel == oracle; (b) smallpt's 9-sphere scene file (11-row layout, 704 bytes) with materials and gamma; (c) a
    507-sphere scene file through the BVH -- both equal the library's single-device frame; all for 1 and 3 column sets."""
    torch = cuda
    w, h, s = 40, 24, 2
    size = ["--width", str(w), "--height", str(h), "--samples", str(s)]
    p = pt.default_params(width=w, height=h, samples=s)
    rays = oracle.gen_rays_from_uniforms(w, h, s, 0, w, oracle.philox_uniforms(9, 0, p.n_paths))
    col, live = oracle.trace(rays, oracle.gen_spheres(), depth=7, return_live=True)
    want = oracle.resolve(col, w, h, s)
    mp = pt.default_material_params(seed=0, max_depth=10)
    for devs in ("0", "0,0,0"):
        rep = _run_host(tmp_path, "--image", "--gen", "--counter-rng", "9", "--depth", "7", "--devices", devs, *size)
        assert np.array_equal(_read_ppm(tmp_path / "output/color.ppm"), want)
        assert rep["mode"] == "image" and rep["segments"] == live and rep["paths"] == p.n_paths and rep["gpus"] == len(devs.split(","))
        assert (tmp_path / "input/spheres.bin").stat().st_size == 512 and not (tmp_path / "input/rays.bin").exists()
        # (b)
        rep = _run_host(tmp_path, "--image", "--gen", "--scene-kind", "smallpt", "--materials", "--gamma", "--max-depth", "10", "--counter-rng", "4",
                        "--p6", "--devices", devs, *size)
        assert (tmp_path / "input/spheres.bin").stat().st_size == 44 * 16 and rep["spheres"] == 9 and rep["materials"] is True
        p9 = pt.default_params(width=w, height=h, samples=s, sphere_count=9, sphere_stride=16)
        d_img = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda")
        pt.render_image_mat(p9, mp, dev(torch, pt.smallpt_scene()), d_img, cam_seed=4, gamma=True)
        assert np.array_equal(_read_ppm(tmp_path / "output/color.ppm"), d_img.cpu().numpy())
        # (c)
        rep = _run_host(tmp_path, "--image", "--gen", "--scene-kind", "random:500", "--materials", "--bvh", "--max-depth", "10", "--counter-rng", "4",
                        "--devices", devs, *size)
        assert (tmp_path / "input/spheres.bin").stat().st_size == 44 * 507 and rep["spheres"] == 507 and rep["bvh"] is True
        tree = pt.Bvh(dev(torch, pt.random_scene(500)), 507, 507)
        pt.render_image_mat_bvh(p, mp, tree, d_img, cam_seed=4)
        tree.close()
        assert np.array_equal(_read_ppm(tmp_path / "output/color.ppm"), d_img.cpu().numpy())
    # a scene file somebody else wrote is read back by size: 704 bytes = 16 columns of 44, nine of them spheres
    assert pt.scene_layout(704, pt.smallpt_scene()) == (9, 16, 11)
    # a file of a size no layout produces is refused
    (tmp_path / "input/spheres.bin").write_bytes(b"\0" * 100)
    from ascendpathtracing_b200.host import build as host_build
    bad = subprocess.run([host_build.build(), "--image", *size], cwd=tmp_path, capture_output=True, text=True)
    assert bad.returncode != 0


def test_host_binary_drop_in_materials_on_ray_files(pt, cuda, oracle, tmp_path):
    """Drop-in file mode with the material kernel: rays.bin + an 11-row spheres.bin -> color.bin == the CPU twin."""
    w, h, s = 32, 16, 1
    rep = _run_host(tmp_path, "--gen", "--scene-kind", "smallpt", "--materials", "--max-depth", "9", "--mat-seed", "6", "--width", str(w), "--height",
                    str(h), "--samples", str(s))
    rays = np.fromfile(tmp_path / "input/rays.bin", dtype=np.float32).reshape(6, -1)
    assert np.array_equal(bits(rays), bits(oracle.gen_rays(w, h, s, seed=0)))
    want, segs = oracle.trace_materials(rays, oracle.smallpt_scene(), 9, 16, max_depth=9, seed=6, return_segments=True)
    got = np.fromfile(tmp_path / "output/color.bin", dtype=np.float32).reshape(3, -1)
    assert np.array_equal(bits(got), bits(want))
    assert rep["segments"] == segs


def test_run_sh_image_mode_for_the_baseline_configs(pt, cuda):
    """`bash run.sh -r gpu -I ...` reaches BASELINE configs C3 / C4 / C5 (here at a fraction of their size): the flags select
    the scene kind, materials, BVH, gamma, counter-RNG seed and GPU count; output/report.json carries the throughput line."""
    import shutil
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gpus = str(min(2, pt.device_count()))
    try:
        for extra, check in [
            (["-W", "96", "-H", "54", "-S", "8", "-C", "1"], lambda r: r["spheres"] == 8 and not r["materials"]),                                   # C3-shaped
            (["-K", "random:1000", "-M", "-B", "-G", "-W", "96", "-H", "54", "-S", "4", "-C", "1"], lambda r: r["spheres"] == 1007 and r["bvh"]),   # C4-shaped
            (["-W", "96", "-H", "54", "-S", "8", "-D", "50", "-C", "1"], lambda r: r["depth"] == 50),                                                # C5-shaped
        ]:
            out = subprocess.run(["bash", os.path.join(root, "run.sh"), "-r", "gpu", "-I", "-g", gpus, "--p6", *extra], capture_output=True, text=True,
                                 timeout=900)
            assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
            rep = json.load(open(os.path.join(root, "output/report.json")))
            assert rep["mode"] == "image" and rep["gpus"] == int(gpus) and rep["mpaths_per_s"] > 0 and check(rep), rep
            img = _read_ppm(os.path.join(root, "output/color.ppm"))
            assert img.shape == (54, 96, 3) and img.std() > 5
    finally:
        shutil.rmtree(os.path.join(root, "input"), ignore_errors=True)
        shutil.rmtree(os.path.join(root, "output"), ignore_errors=True)
