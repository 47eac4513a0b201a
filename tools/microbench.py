#!/usr/bin/env python3
"""Prints what the FP32 issue path of this GPU sustains (ptb200_measure_fp32): the roofline denominators."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ascendpathtracing_b200 as pt  # noqa: E402

KINDS = ["FFMA", "FMUL+FADD alternating", "FFMA2 (packed f32x2)", "FMUL2+FADD2 alternating", "FADD + FSETP/FSEL", "MUFU.RSQ", "sqrt.rn (IEEE)",
         "div.rn (IEEE)", "FMUL2 alone", "FADD2 alone", "FMUL2 + 2 scalar FADD", "FADD2 + 2x(FSETP+FSEL) [FADD2 lanes counted]", "MUFU.RSQ (pure)", "FMNMX x2 (ALU pipe)",
         "FFMA2 + LOP3 [FFMA2 lanes counted]", "FFMA2, three distinct register sources", "FFMA2 + FADD + LOP3 [FP32 lanes counted]",
         "FMUL2, two distinct register sources"]
out = {}
for k, name in enumerate(KINDS):
    best = 0.0
    for _ in range(3):
        g, ms = pt.measure_fp32(k, 2000)
        best = max(best, g)
    out[name] = {"gops": best, "per_sm_per_clk_at_1965MHz": best * 1e9 / 148 / 1.965e9}
    print(f"{name:46s} {best:10.1f} Gop/s   {out[name]['per_sm_per_clk_at_1965MHz']:7.1f} lane-ops/SM/clk @1965MHz", flush=True)
print(json.dumps(out))
