"""Builds the drop-in C++ host (host/main.cpp -> render_gpu) against libptb200.so."""
import os
import subprocess

from .. import build as _lib_build

HERE = os.path.dirname(os.path.abspath(__file__))
EXE = os.path.join(os.path.dirname(HERE), "render_gpu")
DEPS = ["main.cpp", "data_utils.h", "pt_arena.hpp"]


def build(force=False):
    lib = _lib_build.build()
    libdir = os.path.dirname(lib)
    if not force and os.path.isfile(EXE) and all(os.path.getmtime(os.path.join(HERE, d)) <= os.path.getmtime(EXE) for d in DEPS) \
            and os.path.getmtime(lib) <= os.path.getmtime(EXE):
        return EXE
    cmd = [_lib_build.nvcc(), "-Wno-deprecated-gpu-targets", "-O2", "-std=c++17", "-Xcompiler", "-Wall", os.path.join(HERE, "main.cpp"), "-o", EXE,
           f"-L{libdir}", "-lptb200", "-Xlinker", "-rpath=$ORIGIN"]
    subprocess.check_call(cmd)
    return EXE


if __name__ == "__main__":
    print(build(force=True))
