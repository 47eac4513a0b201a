#!/usr/bin/env python3
"""What the host link of this box sustains with pinned memory: H2D alone, D2H alone, both at once (the shape of
ptb200_render_host's traffic: 24 B/path in, 12 B/path out), for several copy sizes."""
import time

import torch

dev = torch.device("cuda", 0)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for mib in (16, 64, 256, 1152):
    n_in, n_out = mib << 20, (mib << 20) // 2
    h_in = torch.empty(n_in, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n_out, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n_in, dtype=torch.uint8, device=dev)
    d_out = torch.empty(n_out, dtype=torch.uint8, device=dev)

    def run(h2d, d2h, reps=8):
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        return (time.perf_counter() - t) / reps

    run(True, True, 2)
    a, b, c = run(True, False), run(False, True), run(True, True)
    print(f"{mib:5d} MiB in / {mib // 2} MiB out: H2D alone {n_in / a / 1e9:5.1f} GB/s, D2H alone {n_out / b / 1e9:5.1f} GB/s, "
          f"together {c * 1e3:6.2f} ms = H2D {n_in / c / 1e9:5.1f} + D2H {n_out / c / 1e9:5.1f} GB/s")
