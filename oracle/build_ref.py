#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY -- build the reference's own cpu-mode kernel into oracle/_ref/.

The reference (`/root/reference`, read-only, never copied into this repo) ships its radiance kernel as
Ascend-C sources that its `run.sh -r cpu` builds with CANN's `tikicpulib`.  CANN is not installable
here, so this recipe compiles the reference sources *where they lie* with g++ against the small
stand-in headers in oracle/shim/ (SURVEY.md Appendix D).  The reference's problem size is a set of
compile-time constants (src/common.h:4-6) and its bounce count a literal (src/render.cpp:141), so one
artefact is built per (W, H, SAMPLES, depth):

    oracle/_ref/render_cpu_<tag>     src/main.cpp + src/render.cpp  (what `run.sh -r cpu` would run)
    oracle/_ref/libref_<tag>.so      src/render.cpp + oracle/ref_driver.cpp  (same kernel, in-process)

with tag = w<W>h<H>s<S>d<depth>.  The build happens in a throw-away directory of symlinks under /tmp
plus a generated common.h; when depth != 5 a patched render.cpp is generated there too.  Nothing
but the two artefacts is written under the repo, and oracle/_ref/ is git-ignored.

On the GPU box /root/reference does not exist: `build_all()` then silently keeps whatever prebuilt
artefacts travelled with the snapshot.
"""
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("PT_REFERENCE_SRC", "/root/reference/src")
OUT = os.path.join(HERE, "_ref")
SHIM = os.path.join(HERE, "shim")

# -O2 and -O0 are bit-identical; -ffp-contract=off keeps every op singly rounded.
CXXFLAGS = ["-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-DASCENDC_CPU_DEBUG", "-w", "-pthread"]

# (W, H, SAMPLES, depth) needed by tests/, smoke() and bench.py's cpu_baseline leg.
DEFAULT_CONFIGS = [
    (16, 16, 1, 5),     # C1: the reference's own default (src/common.h:4-6)
    (64, 64, 1, 5),     # 16 384 paths: fast parity case
    (256, 256, 1, 5),   # 262 144 paths: the survey's probe size
    (64, 64, 4, 5),     # SAMPLES > 1 (resolve order)
    (64, 64, 1, 10),    # depth sweep
    (64, 64, 1, 50),
    (1024, 1024, 1, 5), # 4 194 304 paths: bounded sample of C2 for the reference arm at many steps
    (1024, 768, 4, 5),  # 12 582 912 paths: the C2 frame at 16 spp, the CPU baseline's sample (~18 s of CPU work)
    (1024, 768, 16, 5), # 50 331 648 paths: BASELINE config C2 itself, what `bench.py --impl reference` times per step
]
# The reference's own cpu-mode flags are `-g` alone, i.e. -O0 (cmake/cpu/CMakeLists.txt:26-28); BASELINE.md section 3 asks for that
# variant next to -O2.  Same bits either way (tests), very different speed.
REFERENCE_FLAGS_CONFIGS = [
    (256, 256, 1, 5),   # the survey's probe size: -O0 single thread ~3 s
]


def tag(w, h, s, d):
    return f"w{w}h{h}s{s}d{d}"


def have_reference():
    return os.path.isfile(os.path.join(REF_SRC, "render.cpp"))


def _suffix(opt):
    return "" if opt == "-O2" else "_O0g"   # "_O0g" = the reference's own flags: -g, no optimisation


def lib_path(w, h, s, d=5, opt="-O2"):
    return os.path.join(OUT, f"libref_{tag(w, h, s, d)}{_suffix(opt)}.so")


def bin_path(w, h, s, d=5, opt="-O2"):
    return os.path.join(OUT, f"render_cpu_{tag(w, h, s, d)}{_suffix(opt)}")


def build(w, h, s, d=5, force=False, opt="-O2"):
    """Build both artefacts for one configuration; returns (lib, bin).  opt: "-O2" (default) or "-O0" (+ -g: the reference's own flags)."""
    lib, exe = lib_path(w, h, s, d, opt), bin_path(w, h, s, d, opt)
    if not force and os.path.isfile(lib) and os.path.isfile(exe):
        return lib, exe
    if not have_reference():
        raise RuntimeError(f"reference sources not found at {REF_SRC}; cannot build {tag(w, h, s, d)}")
    n = w * h * s * 4
    if n % 8 or (n // 8) % 128:  # src/render.cpp:68-73 with TILING_NUM = BLOCK_LENGTH / 128
        raise ValueError("reference requires W*H*S*4 divisible by 8 cores x 128 rays")
    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="pt_ref_build_")
    try:
        for f in ("main.cpp", "rt_helper.h", "allocator.h", "data_utils.h"):
            os.symlink(os.path.join(REF_SRC, f), os.path.join(tmp, f))
        if d == 5:
            os.symlink(os.path.join(REF_SRC, "render.cpp"), os.path.join(tmp, "render.cpp"))
        else:
            src = open(os.path.join(REF_SRC, "render.cpp")).read()
            patched, k = re.subn(r"while \(depth < 5\)", f"while (depth < {d})", src)
            if k != 1:
                raise RuntimeError("could not find the bounce-count literal in render.cpp")
            open(os.path.join(tmp, "render.cpp"), "w").write(patched)
        # generated configuration header: same names/types as src/common.h, other sizes
        with open(os.path.join(tmp, "common.h"), "w") as f:
            f.write("#pragma once\n#include <stdint.h>\n"
                    f"const int32_t WIDTH = {w};\nconst int32_t HEIGHT = {h};\nconst int32_t SAMPLES = {s};\n"
                    "const float PI = 3.1415926535897932385f;\nconstexpr float EPSILON = 1e-4;\n"
                    "const int32_t SPHERE_NUM = 8;\nconst int32_t SPHERE_MEMBER_NUM = 10;\nusing Float = float;\n"
                    "const int32_t GENERIC_SIZE = 64;\n")
        flags = [f for f in CXXFLAGS if f != "-O2"] + ([opt] if opt == "-O2" else ["-O0", "-g"]) + [f"-I{tmp}", f"-I{SHIM}"]
        subprocess.check_call(["g++", *flags, os.path.join(tmp, "main.cpp"), os.path.join(tmp, "render.cpp"), "-o", exe])
        subprocess.check_call(["g++", *flags, "-fPIC", "-shared", os.path.join(tmp, "render.cpp"),
                               os.path.join(HERE, "ref_driver.cpp"), "-o", lib])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return lib, exe


def build_all(force=False, verbose=False):
    if not have_reference():
        if verbose:
            print(f"[build_ref] {REF_SRC} absent: keeping prebuilt artefacts in {OUT}")
        return []
    built = []
    for cfg in DEFAULT_CONFIGS:
        built.append(build(*cfg, force=force))
        if verbose:
            print("[build_ref]", tag(*cfg), "ok")
    for cfg in REFERENCE_FLAGS_CONFIGS:
        built.append(build(*cfg, force=force, opt="-O0"))
        if verbose:
            print("[build_ref]", tag(*cfg), "-O0 -g ok")
    return built


if __name__ == "__main__":
    if len(sys.argv) == 1:
        build_all(force=False, verbose=True)
    else:
        w, h, s = (int(x) for x in sys.argv[1:4])
        d = int(sys.argv[4]) if len(sys.argv) > 4 else 5
        print(*build(w, h, s, d, force=True))
