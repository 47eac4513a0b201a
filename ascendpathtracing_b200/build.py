"""Builds libptb200.so (hand-written sm_100a CUDA + the C ABI of include/ptb200.h) in-tree with nvcc."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libptb200.so")
SOURCES = ["capi.cu", "arena.cu", "trace_kernels.cu", "raygen_kernels.cu", "resolve_kernels.cu", "fp32_peak.cu", "bvh.cu", "multi.cu"]
HEADERS = ["pt_device.cuh", "pt_material.cuh", "pt_bvh.cuh", "pt_raygen.cuh", "pt_host.h", "philox.h", os.path.join("..", "..", "include", "ptb200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    # Bit-exactness: no FMA contraction anywhere, IEEE division / square root, denormals kept.
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-fvisibility=hidden",
    "-rdc=false", "-shared", "-Xcompiler", "-pthread",
]


def _stale():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isfile(cand) or cand == "nvcc"):
            return cand
    return "nvcc"


def build_variant(name, defines):
    """A/B builds for kernel experiments: libptb200_<name>.so with extra -D flags (select with PTB200_LIB=...)."""
    out = os.path.join(HERE, f"libptb200_{name}.so")
    subprocess.check_call([nvcc(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], *[os.path.join(CSRC, f) for f in SOURCES], "-o", out])
    return out


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    cmd = [nvcc(), *NVCC_FLAGS, *[os.path.join(CSRC, f) for f in SOURCES], "-o", LIB]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force=True, verbose="-v" in sys.argv))
