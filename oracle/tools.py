"""TEST INFRASTRUCTURE ONLY -- `run.sh -r cpu`: the reference's cpu mode end to end, for side-by-side comparison.

gen (oracle restatement of gen_data.py, bit-identical to it) -> the reference's own render_cpu binary built by
oracle/build_ref.py -> resolve (oracle restatement of data_visualization.py) -> output/color.ppm.
"""
import argparse
import os
import subprocess
import time

import numpy as np

from . import build_ref
from . import oracle as O


def cpu_mode(w, h, s, d):
    os.makedirs("input", exist_ok=True)
    os.makedirs("output", exist_ok=True)
    O.gen_rays(w, h, s, seed=0).tofile("input/rays.bin")
    O.gen_spheres().tofile("input/spheres.bin")
    exe = build_ref.bin_path(w, h, s, d)
    if not os.path.isfile(exe):
        build_ref.build(w, h, s, d)
    t = time.perf_counter()
    subprocess.check_call([exe], stdout=subprocess.DEVNULL)
    dt = time.perf_counter() - t
    n = w * h * s * 4
    print(f"INFO: reference cpu mode: {n} paths in {dt:.3f} s = {n / dt / 1e6:.3f} Mpaths/s (PT_REF_THREADS={os.environ.get('PT_REF_THREADS', '1')})")
    col = np.fromfile("output/color.bin", dtype=np.float32)
    O.write_ppm("output/color.ppm", O.resolve(col, w, h, s))
    print("Generate Result Image")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("cmd", choices=["cpu-mode"])
    ap.add_argument("--width", type=int, default=16)
    ap.add_argument("--height", type=int, default=16)
    ap.add_argument("--samples", type=int, default=1)
    ap.add_argument("--depth", type=int, default=5)
    a = ap.parse_args()
    cpu_mode(a.width, a.height, a.samples, a.depth)
