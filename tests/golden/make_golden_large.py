#!/usr/bin/env python3
"""Second batch of fixtures produced by the REFERENCE ITSELF (dev container only; /root/reference cannot travel):

  * w256h256s1d5_test_soa_delta.npz -- the reference's own NumPy golden model `test_soa` (scripts/gen_data.py:246-429) at
    256 x 256 x 4 spp, the size SURVEY.md 8c quotes (99.45 % of paths bit-identical to the C++ kernel).  The file itself is
    3 MB, so it is stored as what it DIFFERS by from the reference's C++ kernel output: the indices of the differing paths,
    test_soa's colours there, and the sha256 of both complete files.  A test rebuilds test_soa.bin from the oracle's
    colours + this delta and must arrive at the same sha256 -- which pins the oracle on all 262 144 paths and the
    agreement fraction in one go.
  * w64h64s1_test_scene.npz -- the reference's first-hit model `test_scene` (scripts/gen_data.py:134-188) on the 64 x 64
    primary rays: per-ray index of the nearest sphere (recovered from the colour it returns; the eight colours of the
    scene are distinct), plus the sha256 of output/test_scene.bin.  Pins the nearest-hit stage on its own.

Nothing of ours is in the loop except oracle/build_ref.py's compile recipe for the reference's C++ sources.
"""
import hashlib
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/scripts")

from oracle import build_ref  # noqa: E402

import gen_data  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def in_tmp(fn):
    with tempfile.TemporaryDirectory() as tmp:
        cwd = os.getcwd()
        os.chdir(tmp)
        try:
            os.makedirs("input")
            os.makedirs("output")
            return fn()
        finally:
            os.chdir(cwd)


def soa_256():
    w, h, s, depth = 256, 256, 1, 5
    _, exe = build_ref.build(w, h, s, depth)

    def body():
        np.random.seed(0)
        rays = gen_data.gen_rays(w, h, s)
        spheres = gen_data.gen_spheres()
        subprocess.check_call([exe], stdout=subprocess.DEVNULL)
        col = np.fromfile("output/color.bin", dtype=np.float32).reshape(3, -1)
        gen_data.test_soa(rays, spheres)
        soa = np.fromfile("output/test_soa.bin", dtype=np.float32).reshape(3, -1)
        rays_sha = sha(np.fromfile("input/rays.bin", dtype=np.float32))
        return col, soa, rays_sha

    col, soa, rays_sha = in_tmp(body)
    differ = np.nonzero((col.view(np.uint32) != soa.view(np.uint32)).any(axis=0))[0].astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "w256h256s1d5_test_soa_delta.npz"), differ=differ, soa_values=soa[:, differ],
                        color_sha256=np.array(sha(col)), test_soa_sha256=np.array(sha(soa)), rays_sha256=np.array(rays_sha))
    print(f"test_soa 256x256x4spp: {col.shape[1]} paths, {len(differ)} differ from the C++ kernel "
          f"({100 * (1 - len(differ) / col.shape[1]):.3f} % identical)")


def scene_64():
    w, h, s = 64, 64, 1

    def body():
        np.random.seed(0)
        rays = gen_data.gen_rays(w, h, s)
        spheres = gen_data.gen_spheres()
        gen_data.test_scene(rays, spheres)
        return np.fromfile("output/test_scene.bin", dtype=np.float32).reshape(3, -1), spheres.astype(np.float32)

    got, spheres = in_tmp(body)
    # colour -> sphere index (emission for the light, index 7; gen_data.py:176-181); -1 = nothing hit (black, like the front wall:
    # distinguish by construction -- the front wall is index 3 and a miss cannot happen in the closed box, asserted below)
    table = np.array([spheres[k, 4:7] if k == 7 else spheres[k, 7:10] for k in range(8)], dtype=np.float32)
    idx = np.full(got.shape[1], -1, dtype=np.int8)
    for k in range(8):
        idx[(got.T == table[k]).all(axis=1)] = k
    assert (idx >= 0).all()
    np.savez_compressed(os.path.join(HERE, "w64h64s1_test_scene.npz"), first_hit_index=idx, test_scene_sha256=np.array(sha(got)))
    print("test_scene 64x64: hit histogram", np.bincount(idx, minlength=8).tolist())


if __name__ == "__main__":
    scene_64()
    soa_256()
