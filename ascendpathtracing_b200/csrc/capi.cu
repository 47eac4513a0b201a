// extern "C" surface of libptb200.so (declared in include/ptb200.h).  Thin: argument validation, error
// plumbing, workspace management and launch sequencing; the work is in the *_kernels.cu files.
#include <cerrno>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <string>
#include <sys/stat.h>
#include <unistd.h>
#include <vector>

#include "pt_host.h"
#include "pt_raygen.cuh"

namespace ptb200 {

static thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

int fail_cuda(cudaError_t e, const char *what) {
    snprintf(g_err, sizeof g_err, "%s: CUDA error %d (%s)", what, static_cast<int>(e), cudaGetErrorString(e));
    cudaGetLastError();  // clear the sticky-less error state
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver)
        return PTB200_ENODEV;
    return static_cast<int>(e);
}

namespace {

std::mutex g_cfg_mu;
PtParams g_legacy = {16, 16, 1, 5, 8, 8, 7, 12.0f, 0, 0};  // src/common.h:4-6,10; render.cpp:141,194; rt_helper.h:776

int64_t total_paths(const PtParams &p) { return static_cast<int64_t>(p.width) * p.height * 4 * p.samples; }

}  // namespace

int check_params(const PtParams *p, const char *who) {
    if (p == nullptr)
        return fail(PTB200_EINVAL, "%s: params is NULL", who);
    if (p->width < 1 || p->height < 1 || p->samples < 1)
        return fail(PTB200_EINVAL, "%s: width/height/samples must be >= 1 (got %d x %d x %d)", who, p->width, p->height, p->samples);
    if (p->depth < 1 || p->depth > 0xffffff)  // the kernels count a path's bounces in 24 bits of its lane's state word
        return fail(PTB200_EINVAL, "%s: depth must be in [1, 16777215] (got %d)", who, p->depth);
    if (p->sphere_count < 1 || p->sphere_count > 1024)
        return fail(PTB200_EINVAL, "%s: sphere_count must be in [1, 1024] for the brute-force kernel (got %d)", who, p->sphere_count);
    if (p->sphere_stride < p->sphere_count)
        return fail(PTB200_EINVAL, "%s: sphere_stride (%d) < sphere_count (%d)", who, p->sphere_stride, p->sphere_count);
    return PTB200_OK;
}

int check_device(const char *who) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n < 1) {
        cudaGetLastError();
        return fail(PTB200_ENODEV, "%s: no CUDA device (this library has no CPU fallback)", who);
    }
    return PTB200_OK;
}

int check_material_params(const PtMaterialParams *mp, const char *who) {
    if (mp == nullptr)
        return fail(PTB200_EINVAL, "%s: material params is NULL", who);
    if (mp->max_depth < 1 || mp->max_depth > 0xffffff || mp->rr_start < 0 || !(mp->hit_epsilon > 0.0f))
        return fail(PTB200_EINVAL, "%s: need 1 <= max_depth <= 16777215, rr_start >= 0, hit_epsilon > 0", who);
    return PTB200_OK;
}

// ---- per-device workspace arena (grown on demand) ----------------------------------------------------
namespace {
// Per device: a short list of arenas.  One is the normal case; a second (third, ...) appears only when a request does not
// fit while blocks of the existing ones are held -- by another host thread, or by the caller itself (e.g. the multi-device
// entry's staging frame while its worker renders).  When every arena is idle and none fits, they are all replaced by one.
struct Workspace {
    std::vector<PtArena *> arenas;
};
Workspace g_ws[64];
std::mutex g_ws_mu;
constexpr size_t kMinArena = 256u << 20;  // small requests share one arena instead of each forcing a cudaMalloc

// Environment overrides are for experiments; anything unparsable or below `lo` is ignored.
long long env_ll(const char *name, long long dflt, long long lo) {
    const char *e = getenv(name);
    if (e == nullptr || *e == '\0')
        return dflt;
    char *end = nullptr;
    const long long v = strtoll(e, &end, 10);
    return (end == e || v < lo) ? dflt : v;
}
}  // namespace

int ws_alloc(size_t bytes, WsBlock *out) {
    out->ptr = nullptr;
    out->arena = nullptr;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess)
        return fail_cuda(e, "workspace");
    std::lock_guard<std::mutex> lock(g_ws_mu);  // size check and allocation are one critical section
    Workspace &w = g_ws[dev & 63];
    bool all_idle = true;
    for (PtArena *a : w.arenas) {
        if ((out->ptr = ptb200_arena_alloc(a, bytes)) != nullptr) {
            out->arena = a;
            return PTB200_OK;
        }
        all_idle = all_idle && ptb200_arena_in_use(a) == 0;
    }
    if (all_idle) {
        // Nobody holds a block (every user synchronises its streams before ws_free): one larger arena replaces them all.
        // cudaFree waits for the device by itself; no explicit device-wide synchronisation is added for other tenants.
        for (PtArena *a : w.arenas)
            ptb200_arena_destroy(a);
        w.arenas.clear();
    }
    size_t cap = bytes + (bytes >> 3) + (1u << 20);
    if (cap < kMinArena)
        cap = kMinArena;
    PtArena *fresh = nullptr;
    int rc = ptb200_arena_create(cap, &fresh);
    if (rc != PTB200_OK && cap > bytes + 4096) {  // a crowded device: retry with exactly what is needed
        cudaGetLastError();
        rc = ptb200_arena_create(bytes + 4096, &fresh);
    }
    if (rc != PTB200_OK)
        return rc;
    w.arenas.push_back(fresh);
    if ((out->ptr = ptb200_arena_alloc(fresh, bytes)) == nullptr)
        return PTB200_ENOMEM;
    out->arena = fresh;
    return PTB200_OK;
}

void ws_free(WsBlock *b) {
    if (b == nullptr || b->ptr == nullptr)
        return;
    {
        std::lock_guard<std::mutex> lock(g_ws_mu);  // an arena is never replaced while one of its blocks is counted in use
        ptb200_arena_free(b->arena, b->ptr);
    }
    b->ptr = nullptr;
    b->arena = nullptr;
}

}  // namespace ptb200

using namespace ptb200;

extern "C" {

void ptb200_default_params(PtParams *p) {
    if (p == nullptr)
        return;
    const PtParams d = {16, 16, 1, 5, 8, 8, 7, 12.0f, 0, 0};
    *p = d;
}

int ptb200_abi_version(void) { return PTB200_ABI_VERSION; }

const char *ptb200_last_error(void) { return g_err; }

int ptb200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int ptb200_set_legacy_config(const PtParams *p) {
    int rc = check_params(p, "ptb200_set_legacy_config");
    if (rc != PTB200_OK)
        return rc;
    const int64_t n = total_paths(*p);
    // DataFormatCheck, src/render.cpp:68-73 (8 cores, 64-ray tiles, double buffered)
    if (n % 8 != 0 || (n / 8) % 128 != 0)
        return fail(PTB200_EINVAL, "ptb200_set_legacy_config: N = W*H*S*4 = %lld violates the reference's tiling rule (N %% 8 == 0, N/8 %% 128 == 0)",
                    static_cast<long long>(n));
    std::lock_guard<std::mutex> lock(g_cfg_mu);
    g_legacy = *p;
    return PTB200_OK;
}

void ptb200_get_legacy_config(PtParams *p) {
    if (p == nullptr)
        return;
    std::lock_guard<std::mutex> lock(g_cfg_mu);
    *p = g_legacy;
}

int render_do_ex(const PtParams *p, void *stream, const uint8_t *rays, const uint8_t *spheres, uint8_t *colors, int64_t first,
                 int64_t count) {
    int rc = check_params(p, "render_do_ex");
    if (rc != PTB200_OK)
        return rc;
    const int64_t n = total_paths(*p);
    if (count < 0)
        count = n - first;
    if (first < 0 || first + count > n)
        return fail(PTB200_EINVAL, "render_do_ex: slice [%lld, %lld) outside [0, %lld)", static_cast<long long>(first),
                    static_cast<long long>(first + count), static_cast<long long>(n));
    if (count == 0)
        return PTB200_OK;
    if (rays == nullptr || spheres == nullptr || colors == nullptr)
        return fail(PTB200_EINVAL, "render_do_ex: NULL buffer");
    if ((rc = check_device("render_do_ex")) != PTB200_OK)
        return rc;
    cudaError_t e = trace_paths(static_cast<cudaStream_t>(stream), *p, reinterpret_cast<const float *>(rays),
                                reinterpret_cast<const float *>(spheres), reinterpret_cast<float *>(colors), n, first, count, nullptr);
    return e == cudaSuccess ? PTB200_OK : fail_cuda(e, "render_do_ex");
}

void render_do(uint32_t blockDim, void *l2ctrl, void *stream, uint8_t *rays, uint8_t *spheres, uint8_t *colors) {
    (void)blockDim;
    (void)l2ctrl;
    PtParams p;
    ptb200_get_legacy_config(&p);
    int rc = render_do_ex(&p, stream, rays, spheres, colors, 0, -1);
    if (rc != PTB200_OK)  // CHECK_ACL style, src/data_utils.h:41-47: report and carry on
        fprintf(stderr, "%s:%d ptb200 error:%d %s\n", __FILE__, __LINE__, rc, g_err);
}

void render(uint8_t *rays, uint8_t *spheres, uint8_t *colors) {
    render_do(8, nullptr, nullptr, rays, spheres, colors);
    cudaError_t e = cudaStreamSynchronize(nullptr);
    if (e != cudaSuccess)
        fprintf(stderr, "%s:%d ptb200 error:%d %s\n", __FILE__, __LINE__, static_cast<int>(e), cudaGetErrorString(e));
}

int ptb200_gen_rays(const PtParams *p, void *stream, const double *uniforms, uint64_t seed, int32_t x0, int32_t x1, float *rays) {
    int rc = check_params(p, "ptb200_gen_rays");
    if (rc != PTB200_OK)
        return rc;
    if (x0 < 0 || x1 > p->width || x0 > x1)
        return fail(PTB200_EINVAL, "ptb200_gen_rays: columns [%d, %d) outside [0, %d)", x0, x1, p->width);
    if (x0 == x1)
        return PTB200_OK;
    if (rays == nullptr)
        return fail(PTB200_EINVAL, "ptb200_gen_rays: NULL output");
    if ((rc = check_device("ptb200_gen_rays")) != PTB200_OK)
        return rc;
    const int64_t per_col = static_cast<int64_t>(p->height) * 4 * p->samples;
    cudaError_t e = gen_rays(static_cast<cudaStream_t>(stream), *p, uniforms, seed, x0 * per_col, (x1 - x0) * per_col, rays);
    return e == cudaSuccess ? PTB200_OK : fail_cuda(e, "ptb200_gen_rays");
}

int ptb200_resolve(const PtParams *p, void *stream, const float *colors, int32_t x0, int32_t x1, uint8_t *image) {
    int rc = check_params(p, "ptb200_resolve");
    if (rc != PTB200_OK)
        return rc;
    if (x0 < 0 || x1 > p->width || x0 > x1)
        return fail(PTB200_EINVAL, "ptb200_resolve: columns [%d, %d) outside [0, %d)", x0, x1, p->width);
    if (x0 == x1)
        return PTB200_OK;
    if (colors == nullptr || image == nullptr)
        return fail(PTB200_EINVAL, "ptb200_resolve: NULL buffer");
    if ((rc = check_device("ptb200_resolve")) != PTB200_OK)
        return rc;
    const int64_t n = total_paths(*p);
    const int64_t pix0 = static_cast<int64_t>(x0) * p->height;
    const int64_t npix = static_cast<int64_t>(x1 - x0) * p->height;
    cudaError_t e = resolve_pixels(static_cast<cudaStream_t>(stream), *p, colors + pix0 * 4 * p->samples, n, pix0, npix, image, x0, x1 - x0);
    return e == cudaSuccess ? PTB200_OK : fail_cuda(e, "ptb200_resolve");
}

}  // extern "C"

int ptb200::render_image_impl(const char *who, const PtParams *p, const PtMaterialParams *mp, void *stream_, const uint8_t *spheres,
                              const double *uniforms, uint64_t seed, int32_t x0, int32_t x1, int gamma, uint8_t *image, uint64_t *stats,
                              const PtBvh *tree) {
    int rc = check_params(p, who);
    if (rc != PTB200_OK)
        return rc;
    if (x0 < 0 || x1 > p->width || x0 > x1)
        return fail(PTB200_EINVAL, "%s: columns [%d, %d) outside [0, %d)", who, x0, x1, p->width);
    if (x0 == x1)
        return PTB200_OK;
    if ((spheres == nullptr && tree == nullptr) || image == nullptr)
        return fail(PTB200_EINVAL, "%s: NULL buffer", who);
    // column_step k > 1: every k-th column from x0, into a dense [H][ceil((x1-x0)/k)][3] image.  The launch walks that dense
    // frame; ray generation and the RNG keys map back to image columns / global path indices (pt_raygen.cuh).
    const int32_t step = p->column_step > 1 ? p->column_step : 1;
    if (p->column_step < 0)
        return fail(PTB200_EINVAL, "%s: column_step must be >= 0", who);
    if (step > 1 && uniforms != nullptr)
        return fail(PTB200_EINVAL, "%s: a replayed random stream cannot be combined with column_step > 1", who);
    const int32_t x_first = x0;
    if (step > 1) {  // from here on [x0, x1) are columns of the dense frame
        x1 = (x1 - x0 + step - 1) / step;
        x0 = 0;
    }
    if ((rc = check_device(who)) != PTB200_OK)
        return rc;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int64_t spp = 4LL * p->samples;
    const int64_t pix_begin = static_cast<int64_t>(x0) * p->height, pix_end = static_cast<int64_t>(x1) * p->height;
    // Tile = whole pixels, up to 512 Mi paths (6.4 GB of workspace out of 180 GB).  Every launch of a persistent kernel
    // ends in a tail in which all warps run out of fresh paths at about the same time and finish their last ones with
    // ever fewer lanes busy: ~0.1 ms for the mirror kernel but 4-8 ms for the BVH material kernel (paths of up to 64
    // bounces; profiles/r1_c4_tail.md), so a frame should be as few launches as memory allows.  PTB200_TILE_PATHS overrides.
    static const int64_t target_paths = env_ll("PTB200_TILE_PATHS", 512LL << 20, 1024);
    // Fused resolve (trace_kernels.cu, fuse_reduce_chunk), OPT-IN with PTB200_FUSED_RESOLVE=1: with S a power of two in 8..256
    // and a constant-bank scene the trace kernel averages every sub-pixel run itself, in NumPy's order, and writes 12 bytes per
    // RUN instead of 12 bytes per path; the frame is then one pass (launches of 2^30 paths) with 48 bytes of workspace per pixel
    // instead of 6.4 GB.  Same image bit for bit (tests/test_gpu_fused_resolve.py).  It is not the default because it is not
    // faster on B200: the kernel is bound by instruction issue, not by HBM, and the parked colours + in-kernel reduction cost
    // more issue slots (C2: 2.71 ms against 2.56 ms) than the 0.10 ms the separate resolve kernel takes at 5.5 TB/s
    // (profiles/r2_fused_resolve.md).  Read per call: the tests flip it.
    const bool allow_fused = env_ll("PTB200_FUSED_RESOLVE", 0, 0) != 0;
    if (allow_fused && tree == nullptr && fuse_supported(p->samples)) {
        const int64_t npix = pix_end - pix_begin, m = npix * spp;
        WsBlock wsm;
        if ((rc = ws_alloc(sizeof(float) * 3 * 4 * static_cast<size_t>(npix), &wsm)) != PTB200_OK)
            return rc;
        FuseTarget fuse = {static_cast<float *>(wsm.ptr), 4 * npix};
        cudaError_t e = cudaSuccess;
        if (stats != nullptr)
            e = cudaMemsetAsync(stats, 0, 2 * sizeof(uint64_t), stream);
        unsigned long long *seg_stat = stats ? reinterpret_cast<unsigned long long *>(stats) + 1 : nullptr;
        // 32-bit generator indices: pieces of at most 2^30 paths (whole pixels), each its own generator range
        const int64_t piece_pix = (1LL << 30) / spp;
        for (int64_t q = pix_begin; q < pix_end && e == cudaSuccess; q += piece_pix) {
            const int64_t np = (pix_end - q < piece_pix) ? pix_end - q : piece_pix;
            const int64_t mm = np * spp;
            const double *u = uniforms ? uniforms + 2 * (q - pix_begin) * spp : nullptr;
            RayGenSource gen = make_raygen_source(*p, u, seed, q * spp, mm);
            if (step > 1) {
                if (!gen.fast_index) {
                    e = cudaErrorInvalidValue;
                    break;
                }
                gen.x_first = x_first, gen.x_step = step;
            }
            const FuseTarget piece = {fuse.means + (q - pix_begin) * 4, fuse.n_runs};
            if (mp != nullptr)
                e = trace_materials(stream, *p, *mp, nullptr, reinterpret_cast<const float *>(spheres), nullptr, mm, 0, mm,
                                    static_cast<uint64_t>(q * spp), seg_stat, nullptr, &gen, &piece);
            else
                e = trace_paths(stream, *p, nullptr, reinterpret_cast<const float *>(spheres), nullptr, mm, 0, mm, seg_stat, &gen, &piece);
        }
        if (e == cudaSuccess)
            e = resolve_means(stream, *p, fuse.means, fuse.n_runs, pix_begin, npix, image, x0, x1 - x0, gamma);
        const unsigned long long total = static_cast<unsigned long long>(m);
        if (e == cudaSuccess && stats != nullptr)
            e = cudaMemcpyAsync(stats, &total, sizeof total, cudaMemcpyHostToDevice, stream);  // drained by the sync below
        cudaError_t es = cudaStreamSynchronize(stream);
        ws_free(&wsm);
        if (e == cudaSuccess)
            e = es;
        return e == cudaSuccess ? PTB200_OK : fail_cuda(e, who);
    }
    // Rays are generated inside the trace kernel (straight into its shared-memory ring) and never exist in HBM; only the
    // per-path colours of a tile (12 B/path) are materialised between the trace and the resolve kernel.  If the device
    // cannot spare the workspace (other tenants, a smaller GPU) the tile is halved until it fits, down to 16 Mi paths.
    WsBlock ws;
    float *cols = nullptr;
    int64_t tile_pix = 0;
    for (int64_t want = target_paths;; want /= 2) {
        tile_pix = want / spp;
        if (tile_pix < 1)
            tile_pix = 1;
        if (tile_pix > pix_end - pix_begin)
            tile_pix = pix_end - pix_begin;
        const size_t col_bytes = sizeof(float) * 3 * static_cast<size_t>(tile_pix * spp);
        rc = ws_alloc(col_bytes, &ws);
        if (rc == PTB200_OK) {
            cols = static_cast<float *>(ws.ptr);
            break;
        }
        if (rc != PTB200_ENOMEM || want <= (16LL << 20) || tile_pix * spp <= (16LL << 20))
            return rc;
        cudaGetLastError();  // the failed cudaMalloc
    }
    cudaError_t e = cudaSuccess;
    if (stats != nullptr)
        e = cudaMemsetAsync(stats, 0, 2 * sizeof(uint64_t), stream);
    unsigned long long *seg_stat = stats ? reinterpret_cast<unsigned long long *>(stats) + 1 : nullptr;
    for (int64_t q = pix_begin; q < pix_end && e == cudaSuccess; q += tile_pix) {
        const int64_t npix = (pix_end - q < tile_pix) ? pix_end - q : tile_pix;
        const int64_t m = npix * spp;
        const double *u = uniforms ? uniforms + 2 * (q - pix_begin) * spp : nullptr;
        RayGenSource gen = make_raygen_source(*p, u, seed, q * spp, m);
        if (step > 1) {
            if (!gen.fast_index) {  // tiles are whole pixels and below 2^31 paths, so only a frame of 2^31 pixels gets here
                e = cudaErrorInvalidValue;
                break;
            }
            gen.x_first = x_first, gen.x_step = step;
        }
        // the tile is its own m-path problem for the trace kernel
        if (mp != nullptr)
            e = trace_materials(stream, *p, *mp, nullptr, reinterpret_cast<const float *>(spheres), cols, m, 0, m, static_cast<uint64_t>(q * spp), seg_stat,
                                tree, &gen);
        else
            e = trace_paths(stream, *p, nullptr, reinterpret_cast<const float *>(spheres), cols, m, 0, m, seg_stat, &gen);
        if (e != cudaSuccess)
            break;
        if ((e = resolve_pixels(stream, *p, cols, m, q, npix, image, x0, x1 - x0, gamma)) != cudaSuccess)
            break;
    }
    const unsigned long long total = static_cast<unsigned long long>((pix_end - pix_begin) * spp);
    if (e == cudaSuccess && stats != nullptr)
        e = cudaMemcpyAsync(stats, &total, sizeof total, cudaMemcpyHostToDevice, stream);  // drained by the sync below
    // Stream-ordered reuse: the buffers go back to the arena once the work queued above has drained.
    cudaError_t es = cudaStreamSynchronize(stream);
    ws_free(&ws);
    if (e == cudaSuccess)
        e = es;
    return e == cudaSuccess ? PTB200_OK : fail_cuda(e, who);
}

extern "C" {

int ptb200_render_image(const PtParams *p, void *stream, const uint8_t *spheres, const double *uniforms, uint64_t seed, int32_t x0,
                        int32_t x1, uint8_t *image, uint64_t *stats) {
    return render_image_impl("ptb200_render_image", p, nullptr, stream, spheres, uniforms, seed, x0, x1, 0, image, stats);
}

int ptb200_render_image_mat(const PtParams *p, const PtMaterialParams *mp, void *stream, const uint8_t *spheres, uint64_t cam_seed, int32_t x0,
                            int32_t x1, int32_t gamma, uint8_t *image, uint64_t *stats) {
    int rc = check_material_params(mp, "ptb200_render_image_mat");
    if (rc != PTB200_OK)
        return rc;
    return render_image_impl("ptb200_render_image_mat", p, mp, stream, spheres, nullptr, cam_seed, x0, x1, gamma, image, stats);
}

void ptb200_default_material_params(PtMaterialParams *mp) {
    if (mp == nullptr)
        return;
    const PtMaterialParams d = {64, 5, 0.1f, 0, 0};
    *mp = d;
}

int render_do_mat(const PtParams *p, const PtMaterialParams *mp, void *stream, const uint8_t *rays, const uint8_t *spheres, uint8_t *colors,
                  int64_t first, int64_t count, uint64_t path0, uint64_t *stats) {
    int rc = check_params(p, "render_do_mat");
    if (rc != PTB200_OK)
        return rc;
    if ((rc = check_material_params(mp, "render_do_mat")) != PTB200_OK)
        return rc;
    const int64_t n = total_paths(*p);
    if (count < 0)
        count = n - first;
    if (first < 0 || first + count > n)
        return fail(PTB200_EINVAL, "render_do_mat: slice [%lld, %lld) outside [0, %lld)", static_cast<long long>(first),
                    static_cast<long long>(first + count), static_cast<long long>(n));
    if (count == 0)
        return PTB200_OK;
    if (rays == nullptr || spheres == nullptr || colors == nullptr)
        return fail(PTB200_EINVAL, "render_do_mat: NULL buffer");
    if ((rc = check_device("render_do_mat")) != PTB200_OK)
        return rc;
    cudaError_t e = trace_materials(static_cast<cudaStream_t>(stream), *p, *mp, reinterpret_cast<const float *>(rays),
                                    reinterpret_cast<const float *>(spheres), reinterpret_cast<float *>(colors), n, first, count, path0,
                                    reinterpret_cast<unsigned long long *>(stats));
    return e == cudaSuccess ? PTB200_OK : fail_cuda(e, "render_do_mat");
}

// check_params with the sphere fields neutralised: with a tree the scene comes from the handle
static PtParams with_tree_params(const PtParams *p) {
    PtParams q = *p;
    q.sphere_count = 1;
    q.sphere_stride = 1;
    return q;
}

int render_do_mat_bvh(const PtParams *p, const PtMaterialParams *mp, const PtBvh *bvh, void *stream, const uint8_t *rays, uint8_t *colors,
                      int64_t first, int64_t count, uint64_t path0, uint64_t *stats) {
    if (p == nullptr || bvh == nullptr)
        return fail(PTB200_EINVAL, "render_do_mat_bvh: NULL params or tree");
    const PtParams q = with_tree_params(p);
    int rc = check_params(&q, "render_do_mat_bvh");
    if (rc != PTB200_OK)
        return rc;
    if ((rc = check_material_params(mp, "render_do_mat_bvh")) != PTB200_OK)
        return rc;
    const int64_t n = total_paths(q);
    if (count < 0)
        count = n - first;
    if (first < 0 || first + count > n)
        return fail(PTB200_EINVAL, "render_do_mat_bvh: slice outside [0, %lld)", static_cast<long long>(n));
    if (count == 0)
        return PTB200_OK;
    if (rays == nullptr || colors == nullptr)
        return fail(PTB200_EINVAL, "render_do_mat_bvh: NULL buffer");
    if ((rc = check_device("render_do_mat_bvh")) != PTB200_OK)
        return rc;
    cudaError_t e = trace_materials(static_cast<cudaStream_t>(stream), q, *mp, reinterpret_cast<const float *>(rays), nullptr,
                                    reinterpret_cast<float *>(colors), n, first, count, path0, reinterpret_cast<unsigned long long *>(stats), bvh);
    return e == cudaSuccess ? PTB200_OK : fail_cuda(e, "render_do_mat_bvh");
}

int ptb200_render_image_mat_bvh(const PtParams *p, const PtMaterialParams *mp, const PtBvh *bvh, void *stream, uint64_t cam_seed, int32_t x0,
                                int32_t x1, int32_t gamma, uint8_t *image, uint64_t *stats) {
    if (p == nullptr || bvh == nullptr)
        return fail(PTB200_EINVAL, "ptb200_render_image_mat_bvh: NULL params or tree");
    int rc = check_material_params(mp, "ptb200_render_image_mat_bvh");
    if (rc != PTB200_OK)
        return rc;
    const PtParams q = with_tree_params(p);
    return render_image_impl("ptb200_render_image_mat_bvh", &q, mp, stream, nullptr, nullptr, cam_seed, x0, x1, gamma, image, stats, bvh);
}

int ptb200_random_scene(int32_t n_random, uint32_t seed, int32_t stride, float *out) {
    if (out == nullptr || n_random < 0 || stride < 7 + n_random)
        return fail(PTB200_EINVAL, "ptb200_random_scene: need out, n_random >= 0, stride >= 7 + n_random");
    const size_t total = static_cast<size_t>(11) * stride;
    memset(out, 0, total * sizeof(float));
    // scripts/gen_data.py:94-102 without the mirror ball: six walls (DIFF) and the light
    static const double base[7][11] = {
        {1e5, 1e5 + 1, 40.8, 81.6, 0, 0, 0, 0.435, 0.376, 0.667, 0},  {1e5, -1e5 + 99, 40.8, 81.6, 0, 0, 0, 0.667, 0.129, 0.086, 0},
        {1e5, 50, 40.8, 1e5, 0, 0, 0, 0.270, 0.725, 0.486, 0},        {1e5, 50, 40.8, -1e5 + 170, 0, 0, 0, 0, 0, 0, 0},
        {1e5, 50, 1e5, 81.6, 0, 0, 0, 0.5, 0.5, 0.5, 0},              {1e5, 50, -1e5 + 81.6, 81.6, 0, 0, 0, 0.141, 0.408, 0.635, 0},
        {600, 50, 681.6 - 0.27, 81.6, 12, 12, 12, 0, 0, 0, 0}};
    for (int i = 0; i < 7; i++)
        for (int m = 0; m < 11; m++)
            out[static_cast<size_t>(m) * stride + i] = static_cast<float>(m == 0 ? base[i][m] * base[i][m] : base[i][m]);
    std::vector<double> u(static_cast<size_t>(8) * n_random);
    int rc = ptb200_mt19937_uniforms(seed, 0, u.size(), u.data());
    if (rc != PTB200_OK)
        return rc;
    for (int k = 0; k < n_random; k++) {
        const double *q = u.data() + static_cast<size_t>(8) * k;
        const int i = 7 + k;
        const double r = 0.2 + 0.8 * q[3];
        int mat = static_cast<int>(3.0 * q[4]);
        mat = mat > 2 ? 2 : mat;
        const double v[11] = {r * r, 1.0 + 98.0 * q[0], 81.6 * q[1], 170.0 * q[2], 0, 0, 0, 0.2 + 0.75 * q[5], 0.2 + 0.75 * q[6], 0.2 + 0.75 * q[7],
                              static_cast<double>(mat)};
        for (int m = 0; m < 11; m++)
            out[static_cast<size_t>(m) * stride + i] = static_cast<float>(v[m]);
    }
    return PTB200_OK;
}

int ptb200_smallpt_scene(float *out) {
    if (out == nullptr)
        return fail(PTB200_EINVAL, "ptb200_smallpt_scene: NULL output");
    // smallpt's table as quoted in scripts/gen_data.py:77-89: radius, centre, emission, colour, material (0 DIFF, 1 SPEC, 2 REFR)
    static const double tbl[9][11] = {
        {1e5, 1e5 + 1, 40.8, 81.6, 0, 0, 0, .75, .25, .25, 0},   {1e5, -1e5 + 99, 40.8, 81.6, 0, 0, 0, .25, .25, .75, 0},
        {1e5, 50, 40.8, 1e5, 0, 0, 0, .75, .75, .75, 0},         {1e5, 50, 40.8, -1e5 + 170, 0, 0, 0, 0, 0, 0, 0},
        {1e5, 50, 1e5, 81.6, 0, 0, 0, .75, .75, .75, 0},         {1e5, 50, -1e5 + 81.6, 81.6, 0, 0, 0, .75, .75, .75, 0},
        {16.5, 27, 16.5, 47, 0, 0, 0, .999, .999, .999, 1},      {16.5, 73, 16.5, 78, 0, 0, 0, .999, .999, .999, 2},
        {600, 50, 681.6 - .27, 81.6, 12, 12, 12, 0, 0, 0, 0}};
    memset(out, 0, 176 * sizeof(float));
    for (int i = 0; i < 9; i++)
        for (int m = 0; m < 11; m++)
            out[m * 16 + i] = static_cast<float>(m == 0 ? tbl[i][m] * tbl[i][m] : tbl[i][m]);
    return PTB200_OK;
}

int ptb200_render_host(const PtParams *p, const float *rays_host, const float *spheres_host, float *colors_host) {
    return render_host_slice("ptb200_render_host", p, rays_host, spheres_host, colors_host, 0, -1);
}

}  // extern "C"

// Paths [first, first+count) of the N-path HOST buffers on the CURRENT device (the whole job, or one device's share of it:
// the reference's per-core slice, src/render.cpp:24-27).
int ptb200::render_host_slice(const char *who, const PtParams *p, const float *rays_host, const float *spheres_host, float *colors_host, int64_t first,
                              int64_t count) {
    int rc = check_params(p, who);
    if (rc != PTB200_OK)
        return rc;
    if (rays_host == nullptr || spheres_host == nullptr || colors_host == nullptr)
        return fail(PTB200_EINVAL, "%s: NULL buffer", who);
    if ((rc = check_device(who)) != PTB200_OK)
        return rc;
    const int64_t n_total = total_paths(*p);
    if (count < 0)
        count = n_total - first;
    if (first < 0 || first + count > n_total)
        return fail(PTB200_EINVAL, "%s: slice [%lld, %lld) outside [0, %lld)", who, static_cast<long long>(first), static_cast<long long>(first + count),
                    static_cast<long long>(n_total));
    if (count == 0)
        return PTB200_OK;
    const int64_t n = count;  // the job of this call
    rays_host += first;       // plane c of the slice starts at rays_host + c * n_total
    colors_host += first;
    const size_t sph_bytes = sizeof(float) * 10 * static_cast<size_t>(p->sphere_stride);
    const size_t sph_alloc = sph_bytes < 512 ? 512 : sph_bytes;
    // Chunked so that the H2D copy of chunk k+1, the kernel of chunk k and the D2H copy of chunk k-1 overlap
    // (three engines, three streams).  With pinned host memory the copies are truly asynchronous.
    constexpr int kStreams = 4;
    static const int64_t chunk_paths = env_ll("PTB200_HOST_CHUNK", 16LL << 20, 1024);  // experiments only; below 1024 is ignored
    // Peak paths per chunk: 384 MiB in, 192 MiB out.  Measured on C2 (1.2 GB in, 0.6 GB out) with the ramped schedule below
    // (tools/e2e_ramp_ab.py, median of 9 calls): peak 8 Mi, ramps from peak/8: 25.1 ms; from peak/32: 25.0; from peak/128: 25.2;
    // peak 16 Mi from peak/64: 24.9; peak 4 Mi from peak/32: 26.0; equal 4 Mi chunks without ramps 26.7 ms.  The link itself moves
    // the same bytes in 23.0 ms as two giant copies (52.5 GB/s in while 26.3 GB/s go out, tools/pcie_probe.py).
    int64_t chunk = chunk_paths;
    if (chunk > n)
        chunk = n;
    const int nbuf = static_cast<int>((n + chunk - 1) / chunk < kStreams ? (n + chunk - 1) / chunk : kStreams);
    // one workspace block, carved by hand: [spheres | staging buffer 0 | ... | staging buffer nbuf-1], 256-byte aligned pieces
    const size_t per_buf = (sizeof(float) * 9 * static_cast<size_t>(chunk) + 1023) & ~static_cast<size_t>(255);
    const size_t sph_piece = (sph_alloc + 255) & ~static_cast<size_t>(255);
    WsBlock ws;
    if ((rc = ws_alloc(per_buf * nbuf + sph_piece, &ws)) != PTB200_OK)
        return rc;
    float *d_sph = static_cast<float *>(ws.ptr);
    float *d_buf[kStreams] = {nullptr, nullptr, nullptr, nullptr};
    cudaStream_t st[kStreams] = {nullptr, nullptr, nullptr, nullptr};
    cudaError_t e = cudaSuccess;
    bool ok = true;
    for (int b = 0; b < nbuf && ok; b++) {
        d_buf[b] = reinterpret_cast<float *>(static_cast<char *>(ws.ptr) + sph_piece + per_buf * b);
        if ((e = cudaStreamCreateWithFlags(&st[b], cudaStreamNonBlocking)) != cudaSuccess)
            ok = false;
    }
    cudaEvent_t sph_ready = nullptr;
    if (ok && (e = cudaEventCreateWithFlags(&sph_ready, cudaEventDisableTiming)) != cudaSuccess)
        ok = false;
    if (ok) {
        e = cudaMemcpyAsync(d_sph, spheres_host, sph_bytes, cudaMemcpyHostToDevice, st[0]);
        if (e == cudaSuccess)
            e = cudaEventRecord(sph_ready, st[0]);
        for (int b = 1; b < nbuf && e == cudaSuccess; b++)
            e = cudaStreamWaitEvent(st[b], sph_ready, 0);
        PtParams cp = *p;
        // Chunk schedule: the first chunk's upload and the last chunk's kernel + download overlap with nothing, so the
        // schedule ramps up from chunk/64 and down again to chunk/64 (3 ms of exposed transfer with equal chunks on C2, 0.1 ms so).
        std::vector<int64_t> sizes;
        {
            // peak chunk: the largest chunk / 2^j whose two ramps (chunk/64 ... peak/2 each) take at most half of the job
            static const int64_t ramp_div = env_ll("PTB200_HOST_RAMP", 64, 1);  // experiments only
            int64_t peak = chunk, first_c = chunk / ramp_div < 1024 ? chunk : chunk / ramp_div;
            while (peak > first_c && 2 * (peak - first_c) > n / 2)
                peak /= 2;
            int64_t left = n;
            std::vector<int64_t> tail;
            for (int64_t c = first_c; c < peak; c *= 2) {
                sizes.push_back(c);
                tail.push_back(c);
                left -= 2 * c;
            }
            while (left > 0) {
                const int64_t c = left < peak ? left : peak;
                sizes.push_back(c);
                left -= c;
            }
            sizes.insert(sizes.end(), tail.rbegin(), tail.rend());
        }
        int k = 0;
        int64_t a = 0;
        for (size_t i = 0; i < sizes.size() && e == cudaSuccess; a += sizes[i], i++, k++) {
            const int64_t m = sizes[i];
            const int b = k % nbuf;
            float *d_rays = d_buf[b], *d_cols = d_buf[b] + 6 * chunk;
            for (int c = 0; c < 6 && e == cudaSuccess; c++)
                e = cudaMemcpyAsync(d_rays + c * m, rays_host + c * n_total + a, sizeof(float) * m, cudaMemcpyHostToDevice, st[b]);
            if (e == cudaSuccess)
                e = trace_paths(st[b], cp, d_rays, d_sph, d_cols, m, 0, m, nullptr);
            for (int c = 0; c < 3 && e == cudaSuccess; c++)
                e = cudaMemcpyAsync(colors_host + c * n_total + a, d_cols + c * m, sizeof(float) * m, cudaMemcpyDeviceToHost, st[b]);
        }
    }
    for (int b = 0; b < nbuf; b++)
        if (st[b] != nullptr) {
            cudaError_t es = cudaStreamSynchronize(st[b]);
            if (e == cudaSuccess)
                e = es;
            cudaStreamDestroy(st[b]);
        }
    if (sph_ready != nullptr)
        cudaEventDestroy(sph_ready);
    ws_free(&ws);
    if (!ok && e == cudaSuccess)
        return PTB200_ENOMEM;
    return e == cudaSuccess ? PTB200_OK : fail_cuda(e, who);
}

extern "C" {

// ---- host helpers -------------------------------------------------------------------------------------

int ptb200_mt19937_uniforms(uint32_t seed, uint64_t skip, uint64_t n, double *out) {
    if (out == nullptr && n > 0)
        return fail(PTB200_EINVAL, "ptb200_mt19937_uniforms: NULL output");
    // NumPy legacy RandomState (scripts/gen_data.py:438): init_genrand(seed), doubles by genrand_res53.
    uint32_t mt[624];
    mt[0] = seed;
    for (int i = 1; i < 624; i++)
        mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + static_cast<uint32_t>(i);
    int pos = 624;
    auto next = [&]() -> uint32_t {
        if (pos >= 624) {
            for (int k = 0; k < 624; k++) {
                const uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
                mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            pos = 0;
        }
        uint32_t y = mt[pos++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    };
    for (uint64_t i = 0; i < 2 * skip; i++)
        (void)next();
    for (uint64_t i = 0; i < n; i++) {
        const uint32_t a = next() >> 5, b = next() >> 6;
        out[i] = (static_cast<double>(a) * 67108864.0 + static_cast<double>(b)) / 9007199254740992.0;
    }
    return PTB200_OK;
}

int ptb200_default_scene(float *out) {
    if (out == nullptr)
        return fail(PTB200_EINVAL, "ptb200_default_scene: NULL output");
    // scripts/gen_data.py:94-102: radius, centre, emission, colour; float64 table, radius squared, cast (:109,:127)
    static const double tbl[8][10] = {
        {1e5, 1e5 + 1, 40.8, 81.6, 0, 0, 0, 0.435, 0.376, 0.667},  {1e5, -1e5 + 99, 40.8, 81.6, 0, 0, 0, 0.667, 0.129, 0.086},
        {1e5, 50, 40.8, 1e5, 0, 0, 0, 0.270, 0.725, 0.486},        {1e5, 50, 40.8, -1e5 + 170, 0, 0, 0, 0, 0, 0},
        {1e5, 50, 1e5, 81.6, 0, 0, 0, 0.5, 0.5, 0.5},              {1e5, 50, -1e5 + 81.6, 81.6, 0, 0, 0, 0.141, 0.408, 0.635},
        {16.5, 27, 16.5, 47, 0, 0, 0, 0.999, 0.999, 0.999},        {600, 50, 681.6 - 0.27, 81.6, 12, 12, 12, 0, 0, 0}};
    memset(out, 0, 128 * sizeof(float));
    for (int i = 0; i < 8; i++)
        for (int m = 0; m < 10; m++)
            out[m * 8 + i] = static_cast<float>(m == 0 ? tbl[i][m] * tbl[i][m] : tbl[i][m]);
    return PTB200_OK;
}

int ptb200_read_file(const char *path, size_t *file_size, void *buffer, size_t buffer_size) {
    if (path == nullptr || buffer == nullptr)
        return fail(PTB200_EINVAL, "ptb200_read_file: NULL argument");
    struct stat sb;
    if (stat(path, &sb) == -1)
        return fail(PTB200_EIO, "failed to get file %s", path);  // data_utils.h:59-62
    if (!S_ISREG(sb.st_mode))
        return fail(PTB200_EIO, "%s is not a file, please enter a file", path);  // :63-66
    FILE *f = fopen(path, "rb");
    if (f == nullptr)
        return fail(PTB200_EIO, "Open file failed. path = %s", path);  // :70-73
    const size_t size = static_cast<size_t>(sb.st_size);
    if (size == 0) {
        fclose(f);
        return fail(PTB200_EIO, "file size is 0");  // :77-81
    }
    if (size > buffer_size) {
        fclose(f);
        return fail(PTB200_EIO, "file size is larger than buffer size");  // :82-86
    }
    const size_t got = fread(buffer, 1, size, f);
    fclose(f);
    if (got != size)
        return fail(PTB200_EIO, "short read on %s", path);
    if (file_size != nullptr)
        *file_size = size;
    return PTB200_OK;
}

int ptb200_write_file(const char *path, const void *buffer, size_t size) {
    if (buffer == nullptr)
        return fail(PTB200_EINVAL, "Write file failed. buffer is nullptr");  // data_utils.h:103-106
    if (path == nullptr)
        return fail(PTB200_EINVAL, "ptb200_write_file: NULL path");
    const int fd = open(path, O_RDWR | O_CREAT | O_TRUNC, S_IRUSR | S_IWRITE);  // :108
    if (fd < 0)
        return fail(PTB200_EIO, "Open file failed. path = %s", path);
    size_t done = 0;
    const char *src = static_cast<const char *>(buffer);
    while (done < size) {  // the reference issues one write(); large buffers need the loop
        const ssize_t w = write(fd, src + done, size - done);
        if (w < 0) {
            if (errno == EINTR)
                continue;
            break;
        }
        done += static_cast<size_t>(w);
    }
    close(fd);
    if (done != size)
        return fail(PTB200_EIO, "Write file Failed.");  // :116-119
    return PTB200_OK;
}

int ptb200_write_ppm(const char *path, int32_t width, int32_t height, const uint8_t *image) {
    if (path == nullptr || image == nullptr || width < 1 || height < 1)
        return fail(PTB200_EINVAL, "ptb200_write_ppm: bad argument");
    // Build the whole text in memory: "R G B " per pixel, one line per row (data_visualization.py:11-17).
    std::string out;
    out.reserve(static_cast<size_t>(width) * height * 12 + 32);
    char head[64];
    snprintf(head, sizeof head, "P3\n%d %d\n255\n", width, height);
    out += head;
    char num[16];
    for (int r = 0; r < height; r++) {
        for (int x = 0; x < width; x++) {
            const uint8_t *px = image + (static_cast<size_t>(r) * width + x) * 3;
            const int len = snprintf(num, sizeof num, "%d %d %d ", px[0], px[1], px[2]);
            out.append(num, static_cast<size_t>(len));
        }
        out += '\n';
    }
    FILE *f = fopen(path, "w");
    if (f == nullptr)
        return fail(PTB200_EIO, "Open file failed. path = %s", path);
    const size_t w = fwrite(out.data(), 1, out.size(), f);
    const int rc = fclose(f);
    if (w != out.size() || rc != 0)
        return fail(PTB200_EIO, "Write file Failed.");
    return PTB200_OK;
}

int ptb200_measure_fp32(int32_t kind, int32_t iters, double *gops_out, double *ms_out) {
    if (gops_out == nullptr || ms_out == nullptr || iters < 1 || kind < 0 || kind > 17)
        return fail(PTB200_EINVAL, "ptb200_measure_fp32: bad argument");
    int rc = check_device("ptb200_measure_fp32");
    if (rc != PTB200_OK)
        return rc;
    cudaError_t e = measure_fp32(kind, iters, gops_out, ms_out);
    return e == cudaSuccess ? PTB200_OK : fail_cuda(e, "ptb200_measure_fp32");
}

}  // extern "C"
