// Internal host-side declarations shared by the .cu files of libptb200.so.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <mutex>

#include "../../include/ptb200.h"

namespace ptb200 {

constexpr int kTraceThreads = 256;

// trace_kernels.cu -- paths [first, first+count) of N-path SoA buffers; stats (nullable, device) gets
// the number of ray segments actually traced added to it.
// gen (nullable): generate the rays inside the kernel instead of reading `rays` (which may then be NULL); element
// `first` of the launch is the generator's element 0.
// fuse (nullable; needs gen, no tree, fuse_supported(samples), whole runs): the kernel resolves on the fly -- instead of
// per-path colours it writes the NumPy-order mean of every run of S consecutive paths: plane c of `means` ([3][n_runs]) at
// (path - first) / S; `colors` is unused then.  resolve_means() finishes the image.
struct RayGenSource;
struct FuseTarget {
    float *means;
    int64_t n_runs;
};
bool fuse_supported(int samples);
cudaError_t trace_paths(cudaStream_t stream, const PtParams &p, const float *rays, const float *spheres, float *colors, int64_t n,
                        int64_t first, int64_t count, unsigned long long *stats, const RayGenSource *gen = nullptr,
                        const FuseTarget *fuse = nullptr);

// trace_kernels.cu -- material extension (pt_material.cuh): spheres is the 11-row SoA; element i of the slice has RNG
// path index path0 + (i - first).
// tree (nullable): a BVH built by ptb200_bvh_build over the same spheres; then `spheres` may be NULL.
cudaError_t trace_materials(cudaStream_t stream, const PtParams &p, const PtMaterialParams &mp, const float *rays, const float *spheres,
                            float *colors, int64_t n, int64_t first, int64_t count, uint64_t path0, unsigned long long *stats,
                            const PtBvh *tree = nullptr, const RayGenSource *gen = nullptr, const FuseTarget *fuse = nullptr);

// bvh.cu
struct BvhScene;
BvhScene bvh_scene(const PtBvh *b);
const float *bvh_big_soa(const PtBvh *b);
int bvh_big_count(const PtBvh *b);
int bvh_sphere_count(const PtBvh *b);

// raygen_kernels.cu -- camera rays of the path range [path0, path0+m) of a W x H x S image into an
// SoA [6][m] buffer.  uniforms (nullable): 2 doubles per ray, uniforms[0] belongs to path0.
cudaError_t gen_rays(cudaStream_t stream, const PtParams &p, const double *uniforms, uint64_t seed, int64_t path0, int64_t m,
                     float *rays);

// resolve_kernels.cu -- pixels [pix0, pix0+npix) in x-major order (pixel = x*H + y) from colours SoA
// [3][cn] whose element 0 is the first sample of pixel `pix0`; image points at the full
// [H][img_w][3] output whose column 0 is image column x_origin.
cudaError_t resolve_pixels(cudaStream_t stream, const PtParams &p, const float *colors, int64_t cn, int64_t pix0, int64_t npix,
                           uint8_t *image, int32_t x_origin, int32_t img_w, int gamma = 0);

// resolve_kernels.cu -- the second half of the fused resolve: run means [3][n_runs] (element 0 = sub-pixel run 0 of pixel
// `pix0`, four runs per pixel) -> 8-bit pixels, same image conventions as resolve_pixels.
cudaError_t resolve_means(cudaStream_t stream, const PtParams &p, const float *means, int64_t n_runs, int64_t pix0, int64_t npix,
                          uint8_t *image, int32_t x_origin, int32_t img_w, int gamma = 0);

// fp32_peak.cu
cudaError_t measure_fp32(int kind, int iters, double *gops, double *ms);

// capi.cu -- the per-device workspace arenas (csrc/arena.cu), grown on demand.  ws_alloc hands out ONE block of `bytes`
// (256-byte aligned) from an arena that can serve it; when none can, all arenas are replaced by a larger one if nobody holds
// a block, otherwise (another host thread, or the caller itself, is using them) an additional arena is created.  The size
// check and the allocation happen under one lock, so concurrent callers on one device never see each other's arena die.
struct WsBlock {
    void *ptr = nullptr;
    PtArena *arena = nullptr;  // the arena the block came from
};
int ws_alloc(size_t bytes, WsBlock *out);
void ws_free(WsBlock *b);  // the caller has synchronised every stream that touched the block

// capi.cu -- the bodies of ptb200_render_image* and ptb200_render_host, shared with the multi-device entries (multi.cu)
int render_image_impl(const char *who, const PtParams *p, const PtMaterialParams *mp, void *stream, const uint8_t *spheres, const double *uniforms,
                      uint64_t seed, int32_t x0, int32_t x1, int gamma, uint8_t *image, uint64_t *stats, const PtBvh *tree = nullptr);
int render_host_slice(const char *who, const PtParams *p, const float *rays_host, const float *spheres_host, float *colors_host, int64_t first,
                      int64_t count);

// argument validation (capi.cu); every failure sets the thread-local error text
int check_params(const PtParams *p, const char *who);
int check_material_params(const PtMaterialParams *mp, const char *who);
int check_device(const char *who);

// error plumbing (capi.cu)
int fail(int code, const char *fmt, ...);
int fail_cuda(cudaError_t e, const char *what);

}  // namespace ptb200
