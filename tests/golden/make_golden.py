#!/usr/bin/env python3
"""Regenerates tests/golden/* from the REFERENCE ITSELF (run in the dev container, where
/root/reference exists; the fixtures are committed because the reference cannot travel to the GPU box).

Sources of truth, nothing of ours in the loop:
  * rays / spheres: the reference's scripts/gen_data.py executed unmodified (module globals width /
    height / samples set per case; np.random.seed(0) as in its __main__, gen_data.py:438)
  * color.bin: the reference's src/main.cpp + src/render.cpp compiled unmodified against
    oracle/shim/ by oracle/build_ref.py  (= `run.sh -r cpu`'s render_cpu binary)
  * color.ppm / u8 image: the reference's scripts/data_visualization.py executed unmodified
  * test_soa.bin: the reference's own NumPy golden model (gen_data.py:246-429)

C1 (16x16, SAMPLES=1, the reference default) is stored in full; larger cases as sha256 digests plus
the resolved 8-bit image.
"""
import hashlib
import json
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

import data_visualization as dv  # noqa: E402
import gen_data  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_case(w, h, s, depth, full):
    _, exe = build_ref.build(w, h, s, depth)
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        cwd = os.getcwd()
        os.chdir(tmp)
        try:
            os.makedirs("input")
            os.makedirs("output")
            np.random.seed(0)
            rays = gen_data.gen_rays(w, h, s)
            spheres = gen_data.gen_spheres()
            subprocess.check_call([exe], stdout=subprocess.DEVNULL)
            dv.samples = s
            img = dv.decode_color("output/color.bin", w, h, s)  # (w, h, 3): [x][row]
            rays_b = np.fromfile("input/rays.bin", dtype=np.float32)
            sph_b = np.fromfile("input/spheres.bin", dtype=np.float32)
            col_b = np.fromfile("output/color.bin", dtype=np.float32)
            ppm = open("output/color.ppm").read()
            name = f"w{w}h{h}s{s}d{depth}"
            out = {"w": w, "h": h, "s": s, "depth": depth, "rays_sha256": sha(rays_b), "spheres_sha256": sha(sph_b),
                   "color_sha256": sha(col_b), "ppm_sha256": hashlib.sha256(ppm.encode()).hexdigest()}
            # image stored row-major top row first: rows[r][x] = img[x][r]
            np.ascontiguousarray(img.transpose(1, 0, 2)).tofile(os.path.join(HERE, f"{name}_image_u8.bin"))
            if full:
                rays_b.tofile(os.path.join(HERE, f"{name}_rays.bin"))
                sph_b.tofile(os.path.join(HERE, f"{name}_spheres.bin"))
                col_b.tofile(os.path.join(HERE, f"{name}_color.bin"))
                open(os.path.join(HERE, f"{name}_color.ppm"), "w").write(ppm)
                if depth == 5:
                    gen_data.test_soa(rays, spheres)
                    np.fromfile("output/test_soa.bin", dtype=np.float32).tofile(os.path.join(HERE, f"{name}_test_soa.bin"))
        finally:
            os.chdir(cwd)
    return name, out


def main():
    manifest = {}
    for (w, h, s, d, full) in [(16, 16, 1, 5, True), (64, 64, 1, 5, False), (64, 64, 4, 5, False), (64, 64, 1, 10, False),
                               (64, 64, 1, 50, False)]:
        name, out = run_case(w, h, s, d, full)
        manifest[name] = out
        print(name, "ok")
    # NumPy legacy RandomState known answers
    np.random.seed(0)
    manifest["numpy_seed0_first_doubles"] = [float(x).hex() for x in np.random.rand(8)]
    st = np.random.RandomState(0).get_state()
    manifest["numpy_seed0_state_head"] = [int(x) for x in st[1][:4]]
    np.random.seed(0)
    d = np.random.rand(2 * 1000003)
    manifest["numpy_seed0_double_2000000"] = float(d[2000000]).hex()
    json.dump(manifest, open(os.path.join(HERE, "manifest.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
