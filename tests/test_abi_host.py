"""CPU: the C-ABI library loads, exports every symbol include/ptb200.h declares, refuses to compute without a
GPU (no CPU fallback), and its host-side logic (arena bookkeeping, file I/O, RNG replay, scene) is right."""
import ctypes
import os
import subprocess

import numpy as np
import pytest


def test_library_exports_every_declared_symbol(pt):
    L = pt.lib()
    assert len(pt.ABI_SYMBOLS) >= 27
    for name in ("render", "render_do", "render_do_ex", "ptb200_render_host", "ptb200_render_image", "ptb200_gen_rays", "ptb200_resolve"):
        assert name in pt.ABI_SYMBOLS
    missing = [s for s in pt.ABI_SYMBOLS if not hasattr(L, s)]
    assert not missing, f"declared in include/ptb200.h but not exported: {missing}"
    out = subprocess.check_output(["nm", "-D", "--defined-only", pt.lib_path()], text=True)
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert set(pt.ABI_SYMBOLS) <= exported
    # nothing but the declared ABI leaks out of the library
    leaked = {s for s in exported if not s.startswith("_")} - set(pt.ABI_SYMBOLS)
    assert not leaked, leaked
    assert L.ptb200_abi_version() == 2  # round 2: + the multi-device entries and ptb200_scene_layout


def test_scene_file_layout_is_derived_from_its_size(pt):
    """input/spheres.bin beyond the reference's 512 bytes: 512 = the reference's own file, byte
    compatible (count = stride = 8, src/main.cpp:24, common.h:10); otherwise whole 44-byte columns of the 11-row SoA with the
    trailing zero-radius columns counted as padding; anything else is refused."""
    assert pt.scene_layout(512) == (8, 8, 11)
    assert pt.scene_layout(512, pt.default_scene()) == (8, 8, 11)
    assert pt.scene_layout(44 * 16, pt.smallpt_scene()) == (9, 16, 11)
    assert pt.scene_layout(44 * 16) == (16, 16, 11)            # without the contents only the stride is known
    big = pt.random_scene(100, stride=128)
    assert pt.scene_layout(big.nbytes, big) == (107, 128, 11)
    assert pt.scene_layout(44 * 107, pt.random_scene(100)) == (107, 107, 11)
    for bad in (0, 100, 511, 513, 40 * 8):
        with pytest.raises(pt.PtError):
            pt.scene_layout(bad)


def test_library_is_built_for_sm_100a_only(pt):
    out = subprocess.check_output(["cuobjdump", "--list-elf", pt.lib_path()], text=True)
    archs = {ln.split(".")[-2] for ln in out.splitlines() if "sm_" in ln}
    assert archs == {"sm_100a"}, out


def test_product_does_not_touch_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "ascendpathtracing_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.replace("no oracle", ""), f"{f} mentions the oracle: the product must not use it"


def test_no_cpu_fallback(pt):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    p = pt.default_params()
    with pytest.raises(pt.PtError) as ei:
        pt.render_do_ex(p, 16, 32, 48, stream=0)
    assert ei.value.code == -2
    with pytest.raises(pt.PtError):
        pt.render_host(p, np.zeros(6 * 1024, np.float32), np.zeros(128, np.float32), np.zeros(3 * 1024, np.float32))
    assert pt.device_count() == 0


def test_default_params_are_the_reference_constants(pt):
    p = pt.default_params()
    assert (p.width, p.height, p.samples, p.depth, p.sphere_count, p.sphere_stride, p.light_index) == (16, 16, 1, 5, 8, 8, 7)
    assert p.emission_scale == 12.0 and p.n_paths == 1024
    q = pt.get_legacy_config()
    assert (q.width, q.height, q.samples) == (16, 16, 1)


def test_legacy_config_enforces_reference_tiling_rule(pt):
    """src/render.cpp:68-73: N divisible by 8 cores and by 128-ray double tiles."""
    with pytest.raises(pt.PtError):
        pt.set_legacy_config(pt.default_params(width=8, height=8, samples=1))  # N = 256: 32 per core
    pt.set_legacy_config(pt.default_params(width=64, height=64, samples=1))
    assert pt.get_legacy_config().width == 64
    pt.set_legacy_config(pt.default_params())


def test_argument_validation(pt):
    # depth: the kernels keep a path's bounce count in 24 bits of the lane's state word (csrc/trace_kernels.cu)
    for bad in (dict(width=0), dict(depth=0), dict(depth=1 << 24), dict(sphere_count=0), dict(sphere_count=2000, sphere_stride=2000),
                dict(sphere_stride=4)):
        with pytest.raises(pt.PtError) as ei:
            pt.render_do_ex(pt.default_params(**bad), 16, 32, 48, stream=0)
        assert ei.value.code == -1
    with pytest.raises(pt.PtError) as ei:
        pt.render_do_ex(pt.default_params(), 16, 32, 48, first=1000, count=100, stream=0)
    assert ei.value.code == -1


def test_default_scene_and_mt_replay_match_oracle(pt, oracle, golden_dir):
    ref = np.fromfile(os.path.join(golden_dir, "w16h16s1d5_spheres.bin"), dtype=np.float32)
    assert np.array_equal(pt.default_scene().view(np.uint32), ref.view(np.uint32))
    assert np.array_equal(pt.mt19937_uniforms(0, 5000), oracle.mt_doubles(0, 5000))
    assert np.array_equal(pt.mt19937_uniforms(0, 100, skip=777), oracle.mt_doubles(0, 100, skip=777))
    assert np.array_equal(pt.mt19937_uniforms(1234, 10), oracle.mt_doubles(1234, 10))


def test_arena_first_fit_split_and_coalesce(pt):
    """Semantics of the reference Allocator (src/allocator.h:54-151,204-241) on a wrapped (never dereferenced) region."""
    base = 0x7000_0000_0000
    a = pt.Arena(wrap=(base, 1 << 20))
    assert a.capacity == 1 << 20 and a.in_use == 0 and a.largest_free == 1 << 20
    p1 = a.alloc(1000)  # rounded to 1024
    p2 = a.alloc(4096)
    p3 = a.alloc(256)
    assert (p1, p2, p3) == (base, base + 1024, base + 1024 + 4096)
    assert a.in_use == 1024 + 4096 + 256
    a.free(p2)
    assert a.largest_free == (1 << 20) - 1024 - 4096 - 256
    p4 = a.alloc(2048)          # first fit reuses the hole, splitting it
    assert p4 == p2
    p5 = a.alloc(2048)          # the rest of the hole
    assert p5 == p2 + 2048
    with pytest.raises(pt.PtError):
        a.alloc(2 << 20)        # allocator.h:103-105: no block large enough
    with pytest.raises(pt.PtError):
        a.free(base + 8)        # not an allocation start
    a.free(p4)
    with pytest.raises(pt.PtError):
        a.free(p4)              # double free (allocator.h:262-266)
    with pytest.raises(pt.PtError):
        a.free(base + (2 << 20))  # foreign pointer
    for p in (p1, p5, p3):
        a.free(p)
    assert a.in_use == 0 and a.largest_free == 1 << 20  # everything coalesced back into one block
    a.close()


def test_arena_fragmentation_pattern(pt):
    a = pt.Arena(wrap=(0x1000_0000, 64 * 256))
    ptrs = [a.alloc(256) for _ in range(64)]
    with pytest.raises(pt.PtError):
        a.alloc(1)
    for p in ptrs[::2]:
        a.free(p)
    assert a.largest_free == 256            # checkerboard: no two free neighbours
    with pytest.raises(pt.PtError):
        a.alloc(512)
    for p in ptrs[1::2]:
        a.free(p)
    assert a.largest_free == 64 * 256
    a.close()


def test_file_io_semantics(pt, tmp_path):
    """src/data_utils.h:55-122."""
    data = np.arange(1000, dtype=np.float32)
    path = tmp_path / "x.bin"
    pt.write_file(str(path), data)
    assert (os.stat(path).st_mode & 0o777) == (0o600 & ~_umask())
    back = np.zeros(1000, dtype=np.float32)
    assert pt.read_file(str(path), back) == 4000
    assert np.array_equal(back, data)
    big = np.zeros(2000, dtype=np.float32)
    assert pt.read_file(str(path), big) == 4000      # smaller file into a larger buffer is fine
    with pytest.raises(pt.PtError):
        pt.read_file(str(path), np.zeros(10, dtype=np.float32))   # file larger than buffer
    with pytest.raises(pt.PtError):
        pt.read_file(str(tmp_path / "missing.bin"), back)
    (tmp_path / "empty.bin").write_bytes(b"")
    with pytest.raises(pt.PtError):
        pt.read_file(str(tmp_path / "empty.bin"), back)
    with pytest.raises(pt.PtError):
        pt.read_file(str(tmp_path), back)                            # a directory is not a file
    pt.write_file(str(path), data[:10])                               # O_TRUNC
    assert os.path.getsize(path) == 40


def _umask():
    m = os.umask(0)
    os.umask(m)
    return m


def test_write_ppm_matches_reference_text(pt, golden_dir, tmp_path):
    img = np.fromfile(os.path.join(golden_dir, "w16h16s1d5_image_u8.bin"), dtype=np.uint8).reshape(16, 16, 3)
    out = tmp_path / "color.ppm"
    pt.write_ppm(str(out), img)
    assert out.read_text() == open(os.path.join(golden_dir, "w16h16s1d5_color.ppm")).read()
