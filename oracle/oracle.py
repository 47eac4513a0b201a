"""TEST INFRASTRUCTURE ONLY -- ctypes front-end to oracle/libpt_oracle.so and oracle/_ref/libref_*.so.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  The product package (ascendpathtracing_b200) must never import this module.
"""
import ctypes
import os

import numpy as np

from . import build as _build
from . import build_ref as _build_ref

_lib = None

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(_build.build())
        L.pto_trace.restype = ctypes.c_uint64
        L.pto_trace.argtypes = [_f32p, _f32p, _f32p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                ctypes.c_int, ctypes.c_int, ctypes.c_float]
        L.pto_first_hit.restype = None
        L.pto_first_hit.argtypes = [_f32p, _f32p, _f32p, _i32p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_float]
        L.pto_mt_words.restype = None
        L.pto_mt_words.argtypes = [ctypes.c_uint32, _u32p, ctypes.c_int64]
        L.pto_mt_doubles.restype = None
        L.pto_mt_doubles.argtypes = [ctypes.c_uint32, ctypes.c_int64, _f64p, ctypes.c_int64]
        L.pto_philox4x32_10.restype = None
        L.pto_philox4x32_10.argtypes = [_u32p, _u32p, _u32p]
        L.pto_philox_uniforms.restype = None
        L.pto_philox_uniforms.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int64, _f64p]
        L.pto_camera.restype = None
        L.pto_camera.argtypes = [ctypes.c_int, ctypes.c_int, _f64p]
        L.pto_gen_rays_from_uniforms.restype = None
        L.pto_gen_rays_from_uniforms.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f64p, _f32p]
        L.pto_gen_rays.restype = None
        L.pto_gen_rays.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint32, _f32p]
        L.pto_gen_spheres.restype = None
        L.pto_gen_spheres.argtypes = [_f32p]
        L.pto_resolve.restype = None
        L.pto_resolve.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _u8p]
        L.pto_mean_f32.restype = ctypes.c_float
        L.pto_mean_f32.argtypes = [_f32p, ctypes.c_int64]
        L.pm_trace.restype = ctypes.c_uint64
        L.pm_trace.argtypes = [_f32p, _f32p, _f32p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                               ctypes.c_uint64, ctypes.c_uint64]
        L.pm_trace_f64.restype = None
        L.pm_trace_f64.argtypes = [_f32p, _f32p, _f32p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint64]
        L.pm_sincos2pi_array.restype = None
        L.pm_sincos2pi_array.argtypes = [_f32p, _f32p, _f32p, ctypes.c_int64]
        L.pm_random_scene.restype = None
        L.pm_random_scene.argtypes = [ctypes.c_int, ctypes.c_uint32, ctypes.c_int, _f32p]
        L.pm_smallpt_scene.restype = None
        L.pm_smallpt_scene.argtypes = [_f32p]
        L.pto_write_ppm.restype = ctypes.c_int
        L.pto_write_ppm.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, _u8p]
        _lib = L
    return _lib


def set_threads(n):
    os.environ["OMP_NUM_THREADS"] = str(n)


def trace(rays, spheres, depth=5, nsph=8, stride=None, light=7, scale=12.0, first=0, count=None, return_live=False):
    """rays: float32 [6, N]; spheres: float32 flat (>= 10*stride). Returns colors float32 [3, N]."""
    rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(6, -1)
    n = rays.shape[1]
    spheres = np.ascontiguousarray(spheres, dtype=np.float32).reshape(-1)
    stride = nsph if stride is None else stride
    count = n - first if count is None else count
    colors = np.zeros((3, n), dtype=np.float32)
    live = lib().pto_trace(rays.reshape(-1), spheres, colors.reshape(-1), n, first, count, nsph, stride, depth, light, scale)
    return (colors, int(live)) if return_live else colors


def first_hit(rays, spheres, nsph=8, stride=None, eps=1e-4):
    rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(6, -1)
    n = rays.shape[1]
    spheres = np.ascontiguousarray(spheres, dtype=np.float32).reshape(-1)
    t = np.zeros(n, dtype=np.float32)
    idx = np.zeros(n, dtype=np.int32)
    lib().pto_first_hit(rays.reshape(-1), spheres, t, idx, n, nsph, nsph if stride is None else stride, eps)
    return t, idx


def mt_words(seed, n):
    out = np.zeros(n, dtype=np.uint32)
    lib().pto_mt_words(seed, out, n)
    return out


def mt_doubles(seed, n, skip=0):
    out = np.zeros(n, dtype=np.float64)
    lib().pto_mt_doubles(seed, skip, out, n)
    return out


def philox4x32_10(ctr, key):
    out = np.zeros(4, dtype=np.uint32)
    lib().pto_philox4x32_10(np.asarray(ctr, dtype=np.uint32), np.asarray(key, dtype=np.uint32), out)
    return out


def philox_uniforms(seed, first, n):
    out = np.zeros(2 * n, dtype=np.float64)
    lib().pto_philox_uniforms(seed, first, n, out)
    return out


def camera(w, h):
    cam = np.zeros(12, dtype=np.float64)
    lib().pto_camera(w, h, cam)
    return cam


def gen_rays(w, h, s, seed=0):
    rays = np.zeros((6, w * h * 4 * s), dtype=np.float32)
    lib().pto_gen_rays(w, h, s, seed, rays.reshape(-1))
    return rays


def gen_rays_from_uniforms(w, h, s, x0, x1, u):
    m = (x1 - x0) * h * 4 * s
    u = np.ascontiguousarray(u, dtype=np.float64).reshape(-1)
    assert u.size >= 2 * m
    rays = np.zeros((6, m), dtype=np.float32)
    lib().pto_gen_rays_from_uniforms(w, h, s, x0, x1, u, rays.reshape(-1))
    return rays


def gen_spheres():
    out = np.zeros(128, dtype=np.float32)
    lib().pto_gen_spheres(out)
    return out


def resolve(colors, w, h, s):
    colors = np.ascontiguousarray(colors, dtype=np.float32).reshape(-1)
    assert colors.size == 3 * w * h * 4 * s
    img = np.zeros((h, w, 3), dtype=np.uint8)
    lib().pto_resolve(colors, w, h, s, img.reshape(-1))
    return img


def mean_f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return np.float32(lib().pto_mean_f32(a, a.size))


# ---- material extension (parity unpinned by the reference; see pt_oracle_mat.c) --------------------------------
def smallpt_scene():
    out = np.zeros(176, dtype=np.float32)
    lib().pm_smallpt_scene(out)
    return out


def random_scene(n_random, seed=12345, stride=None):
    stride = 7 + n_random if stride is None else stride
    out = np.zeros(11 * stride, dtype=np.float32)
    lib().pm_random_scene(n_random, seed, stride, out)
    return out


def trace_materials(rays, spheres, nsph, stride, max_depth=64, rr_start=5, eps=0.1, seed=0, path0=0, return_segments=False):
    rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(6, -1)
    n = rays.shape[1]
    colors = np.zeros((3, n), dtype=np.float32)
    segs = lib().pm_trace(rays.reshape(-1), np.ascontiguousarray(spheres, dtype=np.float32).reshape(-1), colors.reshape(-1), n, nsph, stride,
                          max_depth, rr_start, eps, seed, path0)
    return (colors, int(segs)) if return_segments else colors


def trace_materials_f64(rays, spheres, nsph, stride, max_depth=64, rr_start=5, seed=0):
    rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(6, -1)
    n = rays.shape[1]
    colors = np.zeros((3, n), dtype=np.float32)
    lib().pm_trace_f64(rays.reshape(-1), np.ascontiguousarray(spheres, dtype=np.float32).reshape(-1), colors.reshape(-1), n, nsph, stride,
                       max_depth, rr_start, seed)
    return colors


def sincos2pi(u):
    u = np.ascontiguousarray(u, dtype=np.float32)
    s, c = np.zeros_like(u), np.zeros_like(u)
    lib().pm_sincos2pi_array(u, s, c, u.size)
    return s, c


def write_ppm(path, img):
    h, w, _ = img.shape
    rc = lib().pto_write_ppm(os.fsencode(path), w, h, np.ascontiguousarray(img).reshape(-1))
    if rc != 0:
        raise OSError(f"cannot write {path}")


# ---- the reference's own kernel, compiled by oracle/build_ref.py -----------------------------------
_ref_libs = {}


def ref_available(w, h, s, depth=5, opt="-O2"):
    return os.path.isfile(_build_ref.lib_path(w, h, s, depth, opt)) or _build_ref.have_reference()


def ref_lib(w, h, s, depth=5, opt="-O2"):
    key = (w, h, s, depth, opt)
    if key not in _ref_libs:
        path = _build_ref.lib_path(w, h, s, depth, opt)
        if not os.path.isfile(path):
            _build_ref.build(w, h, s, depth, opt=opt)
        L = ctypes.CDLL(path)
        L.ref_render.restype = None
        L.ref_render.argtypes = [_u8p, _u8p, _u8p, ctypes.c_int32]
        L.ref_dims.restype = None
        L.ref_dims.argtypes = [ctypes.POINTER(ctypes.c_int32)] * 3
        dims = [ctypes.c_int32() for _ in range(3)]
        L.ref_dims(*[ctypes.byref(d) for d in dims])
        assert tuple(d.value for d in dims) == (w, h, s)
        _ref_libs[key] = L
    return _ref_libs[key]


def ref_render(rays, spheres, w, h, s, depth=5, threads=1, opt="-O2"):
    """Run the reference's own `render` (src/render.cpp:253) over its 8 block slices. Returns colors [3, N]."""
    n = w * h * s * 4
    rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1)
    assert rays.size == 6 * n
    sp = np.zeros(128, dtype=np.float32)
    sp[:] = np.asarray(spheres, dtype=np.float32).reshape(-1)[:128]
    colors = np.zeros(3 * n, dtype=np.float32)
    ref_lib(w, h, s, depth, opt).ref_render(rays.view(np.uint8), sp.view(np.uint8), colors.view(np.uint8), threads)
    return colors.reshape(3, n)
