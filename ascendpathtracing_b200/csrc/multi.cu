// One process, several B200s: the host side of the reference's 8-way split (src/render.cpp:9-10,24-27: blockDim = 8
// AI cores, each on its own contiguous slice) and of its device host flow (src/main.cpp:46-92: set device, stream, H2D,
// launch, synchronise, D2H) for the n GPUs of one box.  One host thread and one stream per device; the scene is replicated;
// nothing is exchanged while rendering.  The production entry gathers the resolved 8-bit column sets on devices[0] with
// peer-to-peer copies over NVLink (cudaMemcpyPeerAsync) and interleaves them there -- no torch, no NCCL communicator to set
// up for one 25 MB exchange per frame.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "pt_host.h"

namespace ptb200 {
namespace {

constexpr int kMaxMulti = 16;

// frame[row][x][c] = set (x mod g) [row][x div g][c]; set r is a dense [h][ceil((w - r) / g)][3] image at stage + r * slot
__global__ void __launch_bounds__(256) interleave_columns_kernel(const uint8_t *__restrict__ stage, uint8_t *__restrict__ frame, int h, int w, int g,
                                                                 size_t slot) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<int64_t>(h) * w)
        return;
    const int row = static_cast<int>(i / w), x = static_cast<int>(i - static_cast<int64_t>(row) * w);
    const int r = x % g, j = x / g;
    const int wr = (w - r + g - 1) / g;
    const uint8_t *src = stage + static_cast<size_t>(r) * slot + (static_cast<size_t>(row) * wr + j) * 3;
    uint8_t *dst = frame + i * 3;
    dst[0] = src[0], dst[1] = src[1], dst[2] = src[2];
}

std::mutex g_peer_mu;
bool g_peer_done[64][64];

// Direct NVLink copies need peer access in the direction of the access; enabled once per ordered pair, failures are not
// fatal (cudaMemcpyPeerAsync then stages through host memory).
void enable_peer(int from, int to) {
    if (from == to || from < 0 || to < 0 || from >= 64 || to >= 64)
        return;
    std::lock_guard<std::mutex> lock(g_peer_mu);
    if (g_peer_done[from][to])
        return;
    g_peer_done[from][to] = true;
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, from, to) != cudaSuccess || !can) {
        cudaGetLastError();
        return;
    }
    int cur = 0;
    cudaGetDevice(&cur);
    cudaSetDevice(from);
    const cudaError_t e = cudaDeviceEnablePeerAccess(to, 0);
    if (e != cudaSuccess)
        cudaGetLastError();  // cudaErrorPeerAccessAlreadyEnabled (the application did it) or unsupported: both fine
    cudaSetDevice(cur);
}

struct Worker {
    int rc = PTB200_OK;
    char err[512] = "";
    double ms = 0.0;
    unsigned long long stats[2] = {0, 0};
};

void keep_error(Worker &w, int rc) {
    w.rc = rc;
    snprintf(w.err, sizeof w.err, "%s", ptb200_last_error());
}

double ms_since(std::chrono::steady_clock::time_point t0) {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

int resolve_devices(const char *who, const int32_t *devices, int32_t n, int (&dev)[kMaxMulti]) {
    if (n < 1 || n > kMaxMulti)
        return fail(PTB200_EINVAL, "%s: n_devices must be in [1, %d] (got %d)", who, kMaxMulti, n);
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have < 1) {
        cudaGetLastError();
        return fail(PTB200_ENODEV, "%s: no CUDA device (this library has no CPU fallback)", who);
    }
    for (int r = 0; r < n; r++) {
        dev[r] = devices ? devices[r] : r;
        if (dev[r] < 0 || dev[r] >= have)
            return fail(PTB200_EINVAL, "%s: device %d does not exist (%d visible)", who, dev[r], have);
    }  // a device may appear more than once: its shares then run one after the other on it (single-GPU tests of the assembly)
    return PTB200_OK;
}

}  // namespace
}  // namespace ptb200

using namespace ptb200;

extern "C" {

int ptb200_render_host_multi(const PtParams *p, const int32_t *devices, int32_t n_devices, const float *rays_host, const float *spheres_host,
                             float *colors_host, double *ms_host) {
    const char *who = "ptb200_render_host_multi";
    int rc = check_params(p, who);
    if (rc != PTB200_OK)
        return rc;
    if (rays_host == nullptr || spheres_host == nullptr || colors_host == nullptr)
        return fail(PTB200_EINVAL, "%s: NULL buffer", who);
    int dev[kMaxMulti];
    if ((rc = resolve_devices(who, devices, n_devices, dev)) != PTB200_OK)
        return rc;
    int home = 0;
    cudaGetDevice(&home);
    const int g = n_devices;
    const int64_t n = static_cast<int64_t>(p->width) * p->height * 4 * p->samples;
    std::vector<Worker> w(g);
    const auto t0 = std::chrono::steady_clock::now();
    auto body = [&](int r) {
        const auto t = std::chrono::steady_clock::now();
        cudaError_t e = cudaSetDevice(dev[r]);
        if (e != cudaSuccess) {
            keep_error(w[r], fail_cuda(e, who));
            return;
        }
        // the reference's per-core slice (src/render.cpp:24-27) with `g` cores: [r*N/g, (r+1)*N/g)
        const int64_t first = n * r / g, last = n * (r + 1) / g;
        const int rcw = render_host_slice(who, p, rays_host, spheres_host, colors_host, first, last - first);
        if (rcw != PTB200_OK)
            keep_error(w[r], rcw);
        w[r].ms = ms_since(t);
    };
    std::vector<std::thread> th;
    for (int r = 1; r < g; r++)
        th.emplace_back(body, r);
    body(0);
    for (auto &t : th)
        t.join();
    cudaSetDevice(home);
    if (ms_host != nullptr) {
        ms_host[0] = ms_since(t0);
        for (int r = 0; r < g; r++)
            ms_host[1 + r] = w[r].ms;
    }
    for (int r = 0; r < g; r++)
        if (w[r].rc != PTB200_OK)
            return fail(w[r].rc, "%s [device %d]", w[r].err, dev[r]);
    return PTB200_OK;
}

int ptb200_render_image_multi(const PtParams *p, const PtMaterialParams *mp, int32_t use_bvh, int32_t gamma, const int32_t *devices, int32_t n_devices,
                              const float *spheres_host, uint64_t seed, uint8_t *image, uint64_t *stats_host, double *ms_host) {
    const char *who = "ptb200_render_image_multi";
    if (p == nullptr)
        return fail(PTB200_EINVAL, "%s: params is NULL", who);
    if (use_bvh && mp == nullptr)
        return fail(PTB200_EINVAL, "%s: the BVH path needs material params (it is the material kernel's scene representation)", who);
    int rc;
    PtParams chk = *p;
    if (use_bvh) {  // the 1024-sphere limit belongs to the brute-force kernels; with a tree the scene comes from the handle
        if (p->sphere_count < 1 || p->sphere_stride < p->sphere_count)
            return fail(PTB200_EINVAL, "%s: need 1 <= sphere_count <= sphere_stride", who);
        chk.sphere_count = 1, chk.sphere_stride = 1;
    }
    if ((rc = check_params(&chk, who)) != PTB200_OK)
        return rc;
    if (mp != nullptr && (rc = check_material_params(mp, who)) != PTB200_OK)
        return rc;
    if (spheres_host == nullptr || image == nullptr)
        return fail(PTB200_EINVAL, "%s: NULL buffer", who);
    int dev[kMaxMulti];
    if ((rc = resolve_devices(who, devices, n_devices, dev)) != PTB200_OK)
        return rc;
    const int g = n_devices, width = p->width, height = p->height;
    if (g > width)
        return fail(PTB200_EINVAL, "%s: more devices (%d) than image columns (%d)", who, g, width);
    int home = 0;
    cudaGetDevice(&home);
    for (int r = 1; r < g; r++) {
        enable_peer(dev[r], dev[0]);
        enable_peer(dev[0], dev[r]);
    }
    const auto t0 = std::chrono::steady_clock::now();
    // devices[0]: staging for the g column sets (equal slots, the widest set's size) and the assembled frame
    const int widest = (width + g - 1) / g;
    const size_t slot = (static_cast<size_t>(height) * widest * 3 + 255) & ~static_cast<size_t>(255);
    const size_t frame_bytes = static_cast<size_t>(height) * width * 3;
    cudaError_t e = cudaSetDevice(dev[0]);
    if (e != cudaSuccess)
        return fail_cuda(e, who);
    WsBlock ws0;
    if ((rc = ws_alloc(slot * g + frame_bytes + 256, &ws0)) != PTB200_OK) {
        cudaSetDevice(home);
        return rc;
    }
    uint8_t *const stage = static_cast<uint8_t *>(ws0.ptr);
    uint8_t *const frame = stage + slot * g;

    const int rows = mp != nullptr ? 11 : 10;
    const size_t scene_bytes = sizeof(float) * rows * static_cast<size_t>(p->sphere_stride);
    std::vector<Worker> w(g);
    auto body = [&](int r) {
        const auto t = std::chrono::steady_clock::now();
        Worker &me = w[r];
        cudaError_t ce = cudaSetDevice(dev[r]);
        cudaStream_t st = nullptr;
        if (ce == cudaSuccess)
            ce = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
        if (ce != cudaSuccess) {
            keep_error(me, fail_cuda(ce, who));
            return;
        }
        const int wr = (width - r + g - 1) / g;  // columns r, r + g, ...
        const size_t img_bytes = static_cast<size_t>(height) * wr * 3;
        const size_t scene_piece = (scene_bytes < 512 ? 512 : scene_bytes + 255) & ~static_cast<size_t>(255);
        const size_t img_piece = (img_bytes + 255) & ~static_cast<size_t>(255);
        WsBlock ws;
        PtBvh *tree = nullptr;
        int rcw = ws_alloc(scene_piece + img_piece + 256, &ws);
        if (rcw == PTB200_OK) {
            uint8_t *d_scene = static_cast<uint8_t *>(ws.ptr), *d_img = d_scene + scene_piece;
            uint64_t *d_stats = reinterpret_cast<uint64_t *>(d_img + img_piece);
            ce = cudaMemcpyAsync(d_scene, spheres_host, scene_bytes, cudaMemcpyHostToDevice, st);
            if (ce == cudaSuccess && use_bvh)
                rcw = ptb200_bvh_build(d_scene, p->sphere_count, p->sphere_stride, st, &tree);
            if (ce == cudaSuccess && rcw == PTB200_OK) {
                PtParams q = *p;
                q.column_step = g > 1 ? g : 0;
                if (use_bvh)
                    q.sphere_count = 1, q.sphere_stride = 1;
                // synchronous on return: the image and the statistics are complete in this device's memory
                rcw = render_image_impl(who, &q, mp, st, use_bvh ? nullptr : d_scene, nullptr, seed, g > 1 ? r : 0, width, gamma, d_img, d_stats, tree);
            }
            if (ce == cudaSuccess && rcw == PTB200_OK) {
                ce = cudaMemcpyAsync(me.stats, d_stats, 2 * sizeof(uint64_t), cudaMemcpyDeviceToHost, st);
                if (ce == cudaSuccess)  // over NVLink into devices[0]'s staging slot (same-device copy for r == 0)
                    ce = cudaMemcpyPeerAsync(stage + slot * r, dev[0], d_img, dev[r], img_bytes, st);
            }
            const cudaError_t es = cudaStreamSynchronize(st);
            if (ce == cudaSuccess)
                ce = es;
        }
        if (rcw != PTB200_OK)
            keep_error(me, rcw);
        else if (ce != cudaSuccess)
            keep_error(me, fail_cuda(ce, who));
        if (tree != nullptr)
            ptb200_bvh_destroy(tree);
        ws_free(&ws);
        cudaStreamDestroy(st);
        me.ms = ms_since(t);
    };
    std::vector<std::thread> th;
    for (int r = 1; r < g; r++)
        th.emplace_back(body, r);
    body(0);
    for (auto &t : th)
        t.join();
    rc = PTB200_OK;
    for (int r = 0; r < g && rc == PTB200_OK; r++)
        if (w[r].rc != PTB200_OK)
            rc = fail(w[r].rc, "%s [device %d]", w[r].err, dev[r]);
    if (rc == PTB200_OK) {
        e = cudaSetDevice(dev[0]);
        const uint8_t *src = stage;  // one device: its only column set IS the frame
        if (e == cudaSuccess && g > 1) {
            const int64_t npix = static_cast<int64_t>(height) * width;
            interleave_columns_kernel<<<static_cast<unsigned>((npix + 255) / 256), 256, 0, cudaStreamPerThread>>>(stage, frame, height, width, g, slot);
            e = cudaGetLastError();
            src = frame;
        }
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(image, src, frame_bytes, cudaMemcpyDefault, cudaStreamPerThread);
        const cudaError_t es = cudaStreamSynchronize(cudaStreamPerThread);
        if (e == cudaSuccess)
            e = es;
        if (e != cudaSuccess)
            rc = fail_cuda(e, who);
    }
    cudaSetDevice(dev[0]);
    ws_free(&ws0);
    cudaSetDevice(home);
    if (ms_host != nullptr) {
        ms_host[0] = ms_since(t0);
        for (int r = 0; r < g; r++)
            ms_host[1 + r] = w[r].ms;
    }
    if (stats_host != nullptr) {
        stats_host[0] = stats_host[1] = 0;
        for (int r = 0; r < g; r++)
            stats_host[0] += w[r].stats[0], stats_host[1] += w[r].stats[1];
    }
    return rc;
}

int ptb200_scene_layout(const float *scene_host, size_t bytes, int32_t *count, int32_t *stride, int32_t *rows) {
    int32_t c = 0, s = 0, r = 0;
    if (bytes == 512) {  // the reference's file: [10][8] + 48 zero floats (src/main.cpp:24, scripts/gen_data.py:120-127)
        c = s = 8;
        r = 11;  // rows 10.. are the zero padding: material 0 = DIFF
    } else if (bytes >= 44 && bytes % 44 == 0 && bytes / 44 <= (1u << 30)) {
        s = static_cast<int32_t>(bytes / 44);
        r = 11;
        c = s;
        if (scene_host != nullptr)
            while (c > 1 && scene_host[c - 1] == 0.0f)  // row 0 = r^2; trailing zero-radius columns are padding
                c--;
    } else {
        return fail(PTB200_EINVAL, "ptb200_scene_layout: %zu bytes is neither the reference's 512-byte scene nor a whole number of 44-byte (11-row) columns",
                    bytes);
    }
    if (count)
        *count = c;
    if (stride)
        *stride = s;
    if (rows)
        *rows = r;
    return PTB200_OK;
}

}  // extern "C"
