"""B200-native drop-in for the per-pixel radiance operator of KVM-Explorer/AscendPathTracing.

The product is libptb200.so (hand-written sm_100a CUDA behind the C ABI of include/ptb200.h) plus the C++
host in host/.  This Python package is plumbing for tests and benchmarks: it builds/loads the library and
mirrors the reference's call surface (render / render_do / ReadFile / WriteFile / the gen_data.py and
data_visualization.py steps) on torch device memory.  There is no CPU fallback.
"""
from .api import (  # noqa: F401
    PtParams, PtError, Arena, default_params, lib, lib_path, device_count, render, render_do, render_do_ex, set_legacy_config,
    get_legacy_config, gen_rays, mt19937_uniforms, default_scene, resolve, render_image, render_host, read_file, write_file,
    write_ppm, measure_fp32, ABI_SYMBOLS, PtMaterialParams, default_material_params, smallpt_scene, render_do_mat, render_image_mat, Bvh, random_scene, render_do_mat_bvh, render_image_mat_bvh,
    render_host_multi, render_image_multi, scene_layout, F_FIXED_DEPTH,
)
