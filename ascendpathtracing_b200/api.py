"""ctypes binding of include/ptb200.h.  Device buffers are torch CUDA tensors (torch = allocator + streams only)."""
import ctypes
import os
import re

import numpy as np

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))
_HEADER = os.path.join(os.path.dirname(_HERE), "include", "ptb200.h")


class PtParams(ctypes.Structure):
    """Run-time form of the reference's src/common.h constants (see include/ptb200.h)."""
    _fields_ = [("width", ctypes.c_int32), ("height", ctypes.c_int32), ("samples", ctypes.c_int32), ("depth", ctypes.c_int32),
                ("sphere_count", ctypes.c_int32), ("sphere_stride", ctypes.c_int32), ("light_index", ctypes.c_int32),
                ("emission_scale", ctypes.c_float), ("flags", ctypes.c_int32), ("column_step", ctypes.c_int32)]

    @property
    def n_paths(self):
        return self.width * self.height * 4 * self.samples


F_FIXED_DEPTH = 1


class PtMaterialParams(ctypes.Structure):
    """Material extension parameters (include/ptb200.h)."""
    _fields_ = [("max_depth", ctypes.c_int32), ("rr_start", ctypes.c_int32), ("hit_epsilon", ctypes.c_float), ("reserved", ctypes.c_int32),
                ("seed", ctypes.c_uint64)]


class PtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"ptb200 error {code}: {msg}")
        self.code = code


def _declared_symbols():
    """Every function include/ptb200.h declares (the ABI contract the tests check the .so against)."""
    text = open(_HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}()]*\)\s*;", text)
    return sorted(set(n for n in names if n not in ("defined",)))


ABI_SYMBOLS = _declared_symbols()

_lib = None


def lib_path():
    return _build.LIB


def lib():
    """Loads libptb200.so (building it when missing or stale). Raises if it cannot be built/loaded: no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("PTB200_LIB") or _build.build()  # PTB200_LIB: A/B-test another build of the same ABI
    L = ctypes.CDLL(path)
    c = ctypes
    vp, i32, i64, u32, u64, sz = c.c_void_p, c.c_int32, c.c_int64, c.c_uint32, c.c_uint64, c.c_size_t
    PP = c.POINTER(PtParams)
    sig = {
        "ptb200_default_params": (None, [PP]),
        "ptb200_abi_version": (c.c_int, []),
        "ptb200_last_error": (c.c_char_p, []),
        "ptb200_device_count": (c.c_int, []),
        "render": (None, [vp, vp, vp]),
        "render_do": (None, [u32, vp, vp, vp, vp, vp]),
        "ptb200_set_legacy_config": (c.c_int, [PP]),
        "ptb200_get_legacy_config": (None, [PP]),
        "render_do_ex": (c.c_int, [PP, vp, vp, vp, vp, i64, i64]),
        "ptb200_gen_rays": (c.c_int, [PP, vp, vp, u64, i32, i32, vp]),
        "ptb200_mt19937_uniforms": (c.c_int, [u32, u64, u64, vp]),
        "ptb200_default_scene": (c.c_int, [vp]),
        "ptb200_resolve": (c.c_int, [PP, vp, vp, i32, i32, vp]),
        "ptb200_render_image": (c.c_int, [PP, vp, vp, vp, u64, i32, i32, vp, vp]),
        "ptb200_render_host": (c.c_int, [PP, vp, vp, vp]),
        "ptb200_default_material_params": (None, [c.POINTER(PtMaterialParams)]),
        "render_do_mat": (c.c_int, [PP, c.POINTER(PtMaterialParams), vp, vp, vp, vp, i64, i64, u64, vp]),
        "ptb200_smallpt_scene": (c.c_int, [vp]),
        "ptb200_render_image_mat": (c.c_int, [PP, c.POINTER(PtMaterialParams), vp, vp, u64, i32, i32, i32, vp, vp]),
        "ptb200_bvh_build": (c.c_int, [vp, i32, i32, vp, c.POINTER(vp)]),
        "ptb200_bvh_destroy": (c.c_int, [vp]),
        "ptb200_bvh_info": (c.c_int, [vp, c.POINTER(i32), c.POINTER(i32), c.POINTER(i32), c.POINTER(i32)]),
        "ptb200_bvh_first_hit": (c.c_int, [vp, vp, vp, i64, c.c_float, vp, vp]),
        "render_do_mat_bvh": (c.c_int, [PP, c.POINTER(PtMaterialParams), vp, vp, vp, vp, i64, i64, u64, vp]),
        "ptb200_render_image_mat_bvh": (c.c_int, [PP, c.POINTER(PtMaterialParams), vp, vp, u64, i32, i32, i32, vp, vp]),
        "ptb200_random_scene": (c.c_int, [i32, u32, i32, vp]),
        "ptb200_render_host_multi": (c.c_int, [PP, vp, i32, vp, vp, vp, vp]),
        "ptb200_render_image_multi": (c.c_int, [PP, vp, i32, i32, vp, i32, vp, u64, vp, vp, vp]),
        "ptb200_scene_layout": (c.c_int, [vp, sz, c.POINTER(i32), c.POINTER(i32), c.POINTER(i32)]),
        "ptb200_arena_create": (c.c_int, [sz, c.POINTER(vp)]),
        "ptb200_arena_wrap": (c.c_int, [vp, sz, c.POINTER(vp)]),
        "ptb200_arena_destroy": (c.c_int, [vp]),
        "ptb200_arena_alloc": (vp, [vp, sz]),
        "ptb200_arena_free": (c.c_int, [vp, vp]),
        "ptb200_arena_capacity": (sz, [vp]),
        "ptb200_arena_in_use": (sz, [vp]),
        "ptb200_arena_largest_free": (sz, [vp]),
        "ptb200_read_file": (c.c_int, [c.c_char_p, c.POINTER(sz), vp, sz]),
        "ptb200_write_file": (c.c_int, [c.c_char_p, vp, sz]),
        "ptb200_write_ppm": (c.c_int, [c.c_char_p, i32, i32, vp]),
        "ptb200_measure_fp32": (c.c_int, [i32, i32, c.POINTER(c.c_double), c.POINTER(c.c_double)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise PtError(rc, lib().ptb200_last_error().decode(errors="replace"))


def default_params(**over):
    p = PtParams()
    lib().ptb200_default_params(ctypes.byref(p))
    for k, v in over.items():
        setattr(p, k, v)
    return p


def device_count():
    return lib().ptb200_device_count()


def _ptr(t):
    """Device (or host) address of a torch tensor / numpy array / None."""
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        return t.ctypes.data
    if isinstance(t, int):
        return t
    return t.data_ptr()


def _stream_handle(stream):
    if stream is None:
        import torch
        return torch.cuda.current_stream().cuda_stream
    if isinstance(stream, int):
        return stream
    return stream.cuda_stream


# ---- the reference's kernel entry points (src/render.cpp:253,264) -------------------------------------

def set_legacy_config(p):
    _check(lib().ptb200_set_legacy_config(ctypes.byref(p)))


def get_legacy_config():
    p = PtParams()
    lib().ptb200_get_legacy_config(ctypes.byref(p))
    return p


def render(rays, spheres, colors):
    """render(rays, spheres, colors): device tensors, legacy configuration, synchronous (src/main.cpp:37)."""
    lib().render(_ptr(rays), _ptr(spheres), _ptr(colors))


def render_do(block_dim, l2ctrl, stream, rays, spheres, colors):
    """render_do(blockDim, l2ctrl, stream, rays, spheres, colors): asynchronous on stream (src/main.cpp:74)."""
    lib().render_do(block_dim, l2ctrl, _stream_handle(stream), _ptr(rays), _ptr(spheres), _ptr(colors))


def render_do_ex(p, rays, spheres, colors, first=0, count=-1, stream=None):
    _check(lib().render_do_ex(ctypes.byref(p), _stream_handle(stream), _ptr(rays), _ptr(spheres), _ptr(colors), first, count))


# ---- steps either side -----------------------------------------------------------------------------------

def gen_rays(p, rays_out, x0=0, x1=None, uniforms=None, seed=0, stream=None):
    x1 = p.width if x1 is None else x1
    _check(lib().ptb200_gen_rays(ctypes.byref(p), _stream_handle(stream), _ptr(uniforms), seed, x0, x1, _ptr(rays_out)))


def mt19937_uniforms(seed, n, skip=0):
    out = np.empty(n, dtype=np.float64)
    _check(lib().ptb200_mt19937_uniforms(seed, skip, n, out.ctypes.data))
    return out


def default_scene():
    out = np.zeros(128, dtype=np.float32)
    _check(lib().ptb200_default_scene(out.ctypes.data))
    return out


def resolve(p, colors, image_out, x0=0, x1=None, stream=None):
    x1 = p.width if x1 is None else x1
    _check(lib().ptb200_resolve(ctypes.byref(p), _stream_handle(stream), _ptr(colors), x0, x1, _ptr(image_out)))


def render_image(p, spheres, image_out, x0=0, x1=None, uniforms=None, seed=0, stats=None, stream=None):
    x1 = p.width if x1 is None else x1
    _check(lib().ptb200_render_image(ctypes.byref(p), _stream_handle(stream), _ptr(spheres), _ptr(uniforms), seed, x0, x1,
                                     _ptr(image_out), _ptr(stats)))


def render_host(p, rays_host, spheres_host, colors_host):
    """Host buffers (numpy arrays or pinned torch CPU tensors) in and out; synchronous."""
    _check(lib().ptb200_render_host(ctypes.byref(p), _ptr(rays_host), _ptr(spheres_host), _ptr(colors_host)))


# ---- one process, several GPUs ---------------------------------------------------------------------------

def _device_list(devices):
    if devices is None:
        raise ValueError("devices: an int (0..n-1) or a list of device indices")
    if isinstance(devices, int):
        return None, devices
    arr = (ctypes.c_int32 * len(devices))(*devices)
    return arr, len(devices)


def render_host_multi(p, devices, rays_host, spheres_host, colors_host):
    """ptb200_render_host over several GPUs of this process (contiguous path slices). Returns [wall_ms, ms of device 0, ...]."""
    arr, n = _device_list(devices)
    ms = (ctypes.c_double * (1 + n))()
    _check(lib().ptb200_render_host_multi(ctypes.byref(p), arr, n, _ptr(rays_host), _ptr(spheres_host), _ptr(colors_host), ms))
    return list(ms)


def render_image_multi(p, devices, spheres_host, image_out, seed=0, mp=None, use_bvh=False, gamma=False):
    """The whole frame on several GPUs of this process (strided columns, P2P gather on devices[0]).  image_out: numpy array /
    pinned CPU tensor / CUDA tensor on devices[0], [H][W][3] uint8.  Returns (stats [paths, segments], [wall_ms, device ms...])."""
    arr, n = _device_list(devices)
    ms = (ctypes.c_double * (1 + n))()
    st = (ctypes.c_uint64 * 2)()
    _check(lib().ptb200_render_image_multi(ctypes.byref(p), ctypes.byref(mp) if mp is not None else None, 1 if use_bvh else 0, 1 if gamma else 0,
                                           arr, n, _ptr(spheres_host), seed, _ptr(image_out), st, ms))
    return [int(st[0]), int(st[1])], list(ms)


def scene_layout(nbytes, scene=None):
    """(count, stride, rows) of a spheres.bin of nbytes bytes (512 = the reference's file; else 44-byte columns)."""
    c, s, r = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
    _check(lib().ptb200_scene_layout(_ptr(scene), nbytes, ctypes.byref(c), ctypes.byref(s), ctypes.byref(r)))
    return c.value, s.value, r.value


# ---- material extension ----------------------------------------------------------------------------------

def default_material_params(**over):
    mp = PtMaterialParams()
    lib().ptb200_default_material_params(ctypes.byref(mp))
    for k, v in over.items():
        setattr(mp, k, v)
    return mp


def smallpt_scene():
    out = np.zeros(176, dtype=np.float32)
    _check(lib().ptb200_smallpt_scene(out.ctypes.data))
    return out


def render_do_mat(p, mp, rays, spheres, colors, first=0, count=-1, path0=0, stats=None, stream=None):
    _check(lib().render_do_mat(ctypes.byref(p), ctypes.byref(mp), _stream_handle(stream), _ptr(rays), _ptr(spheres), _ptr(colors), first, count,
                               path0, _ptr(stats)))


def render_image_mat(p, mp, spheres, image_out, x0=0, x1=None, cam_seed=0, gamma=False, stats=None, stream=None):
    x1 = p.width if x1 is None else x1
    _check(lib().ptb200_render_image_mat(ctypes.byref(p), ctypes.byref(mp), _stream_handle(stream), _ptr(spheres), cam_seed, x0, x1,
                                         1 if gamma else 0, _ptr(image_out), _ptr(stats)))


# ---- large scenes: BVH -----------------------------------------------------------------------------------

class Bvh:
    """Handle of a GPU-built sphere BVH (ptb200_bvh_*)."""

    def __init__(self, spheres, count, stride, stream=None):
        h = ctypes.c_void_p()
        _check(lib().ptb200_bvh_build(_ptr(spheres), count, stride, _stream_handle(stream), ctypes.byref(h)))
        self._h = h

    def info(self):
        v = [ctypes.c_int32() for _ in range(4)]
        _check(lib().ptb200_bvh_info(self._h, *[ctypes.byref(x) for x in v]))
        return dict(zip(("spheres", "big", "small", "nodes"), (x.value for x in v)))

    def first_hit(self, rays, n, tmin_out, idx_out, eps=1e-4, stream=None):
        _check(lib().ptb200_bvh_first_hit(self._h, _stream_handle(stream), _ptr(rays), n, eps, _ptr(tmin_out), _ptr(idx_out)))

    def close(self):
        if self._h:
            lib().ptb200_bvh_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def random_scene(n_random, seed=12345, stride=None):
    stride = 7 + n_random if stride is None else stride
    out = np.zeros(11 * stride, dtype=np.float32)
    _check(lib().ptb200_random_scene(n_random, seed, stride, out.ctypes.data))
    return out


def render_do_mat_bvh(p, mp, bvh, rays, colors, first=0, count=-1, path0=0, stats=None, stream=None):
    _check(lib().render_do_mat_bvh(ctypes.byref(p), ctypes.byref(mp), bvh._h, _stream_handle(stream), _ptr(rays), _ptr(colors), first, count, path0,
                                   _ptr(stats)))


def render_image_mat_bvh(p, mp, bvh, image_out, x0=0, x1=None, cam_seed=0, gamma=False, stats=None, stream=None):
    x1 = p.width if x1 is None else x1
    _check(lib().ptb200_render_image_mat_bvh(ctypes.byref(p), ctypes.byref(mp), bvh._h, _stream_handle(stream), cam_seed, x0, x1,
                                             1 if gamma else 0, _ptr(image_out), _ptr(stats)))


# ---- arena (src/allocator.h) ---------------------------------------------------------------------------

class Arena:
    """RAII-ish wrapper over PtArena: Alloc/Free with the reference Allocator's split/coalesce semantics."""

    def __init__(self, nbytes=None, wrap=None):
        h = ctypes.c_void_p()
        if wrap is not None:
            base, size = wrap
            _check(lib().ptb200_arena_wrap(base, size, ctypes.byref(h)))
        else:
            _check(lib().ptb200_arena_create(nbytes, ctypes.byref(h)))
        self._h = h

    def alloc(self, nbytes):
        p = lib().ptb200_arena_alloc(self._h, nbytes)
        if not p:
            raise PtError(-3, lib().ptb200_last_error().decode(errors="replace"))
        return p

    def free(self, ptr):
        _check(lib().ptb200_arena_free(self._h, ptr))

    @property
    def capacity(self):
        return lib().ptb200_arena_capacity(self._h)

    @property
    def in_use(self):
        return lib().ptb200_arena_in_use(self._h)

    @property
    def largest_free(self):
        return lib().ptb200_arena_largest_free(self._h)

    def close(self):
        if self._h:
            lib().ptb200_arena_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- file I/O (src/data_utils.h:55-122) ----------------------------------------------------------------

def read_file(path, buf):
    size = ctypes.c_size_t(0)
    _check(lib().ptb200_read_file(os.fsencode(path), ctypes.byref(size), buf.ctypes.data, buf.nbytes))
    return size.value


def write_file(path, buf):
    buf = np.ascontiguousarray(buf)
    _check(lib().ptb200_write_file(os.fsencode(path), buf.ctypes.data, buf.nbytes))


def write_ppm(path, image):
    image = np.ascontiguousarray(image, dtype=np.uint8)
    h, w, _ = image.shape
    _check(lib().ptb200_write_ppm(os.fsencode(path), w, h, image.ctypes.data))


def measure_fp32(kind, iters=2000):
    g, ms = ctypes.c_double(), ctypes.c_double()
    _check(lib().ptb200_measure_fp32(kind, iters, ctypes.byref(g), ctypes.byref(ms)))
    return g.value, ms.value
