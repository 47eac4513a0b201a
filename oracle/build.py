#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY -- compile oracle/pt_oracle*.c into oracle/libpt_oracle.so (gcc, OpenMP)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRCS = [os.path.join(HERE, "pt_oracle.c"), os.path.join(HERE, "pt_oracle_mat.c")]
LIB = os.path.join(HERE, "libpt_oracle.so")
CFLAGS = ["-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC", "-Wall", "-Wno-unused-label"]


def build(force=False):
    if not force and os.path.isfile(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in SRCS):
        return LIB
    subprocess.check_call(["gcc", *CFLAGS, *SRCS, "-o", LIB, "-lm"])
    return LIB


if __name__ == "__main__":
    print(build(force=True))
