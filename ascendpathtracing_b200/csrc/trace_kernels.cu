// Radiance kernel for sm_100a: the B200-native replacement of the reference's KernelRender
// (src/render.cpp:18-223) and its helpers (src/rt_helper.h).
//
// Mapping (not a port): the reference tiles 64 rays through ~30 vector ops per sphere with a scratch
// allocator in between; here one CUDA lane owns one path at a time and keeps its whole state in
// registers.  The kernel is persistent: the grid is sized to what the 148 SMs hold at once, and every
// lane walks its own strided sequence of paths (i, i + lanes, i + 2*lanes, ...).  The loop body is exactly
// one bounce, so lanes of a warp may be at different bounces of different paths without any
// divergence: a lane whose path has settled (light reached / throughput zero -- bit-identical early
// termination) swaps in its next path, already prefetched, while its neighbours carry on.
// The scene lives in the constant bank (immediate operands) plus a shared copy for per-lane lookups.
#include "pt_device.cuh"
#include "pt_host.h"

namespace ptb200 {

// ---- scene staging: device SoA [10][stride] -> constant bank -------------------------------------
// Writes through the global-memory alias of the __constant__ symbols (cudaGetSymbolAddress); the
// constant cache is coherent across kernel launches, and launches on one stream are ordered.
__global__ void pack_scene_kernel(const float *__restrict__ spheres, int nsph, int stride, SceneConst *dst, int *zero_ok) {
    __shared__ int ok;
    if (threadIdx.x == 0)
        ok = 1;
    __syncthreads();
    for (int k = threadIdx.x; k < nsph; k += blockDim.x) {
        dst->r2[k] = spheres[0 * stride + k];
        dst->cx[k] = spheres[1 * stride + k];
        dst->cy[k] = spheres[2 * stride + k];
        dst->cz[k] = spheres[3 * stride + k];
        const float r = spheres[7 * stride + k], g = spheres[8 * stride + k], b = spheres[9 * stride + k];
        dst->kr[k] = r;
        dst->kg[k] = g;
        dst->kb[k] = b;
        // zero-throughput early stop is exact only for finite colours with a clear sign bit
        const unsigned ur = __float_as_uint(r), ug = __float_as_uint(g), ub = __float_as_uint(b);
        if ((ur | ug | ub) >> 31 || !isfinite(r) || !isfinite(g) || !isfinite(b))
            ok = 0;
    }
    __syncthreads();
    if (threadIdx.x == 0)
        *zero_ok = ok;
}

// ---- the trace kernel -----------------------------------------------------------------------------
__device__ __forceinline__ void load_ray(const float *__restrict__ rays, int64_t n, int64_t i, float (&r)[6]) {
#pragma unroll
    for (int c = 0; c < 6; c++)
        r[c] = __ldg(rays + c * n + i);
}

template <int NS, bool EARLY>
__global__ void __launch_bounds__(kTraceThreads) trace_paths_kernel(const float *__restrict__ rays, float *__restrict__ colors, int64_t n,
                                                                    int64_t first, int64_t count, int depth, int nsph, int light,
                                                                    float scale, unsigned long long *__restrict__ stats) {
    extern __shared__ float4 smem[];
    SceneShared sh;
    stage_scene_shared(smem, nsph, sh);
    const bool zero_stop = c_scene_zero_stop_ok != 0;

    const int64_t lanes = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int64_t end = first + count;
    int64_t i = first + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    bool active = i < end;

    PathState p;
    float cur[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 1.f};
    if (active)
        load_ray(rays, n, i, cur);
    p.ox = cur[0], p.oy = cur[1], p.oz = cur[2], p.dx = cur[3], p.dy = cur[4], p.dz = cur[5];
    p.rr = p.rg = p.rb = 1.0f;
    p.alive = true;
    int bounce = 0;

    // software prefetch of this lane's next path
    float nxt[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 1.f};
    int64_t inext = i + lanes;
    bool has_next = inext < end;
    if (has_next)
        load_ray(rays, n, inext, nxt);

    unsigned int segs = 0;
    while (__any_sync(0xffffffffu, active)) {
        float tmin;
        int idx;
        nearest_hit<NS>(p, nsph, tmin, idx);
        bounce_and_shade(p, tmin, idx, light, sh);
        bounce++;
        segs += active ? 1u : 0u;
        const bool fin = (bounce >= depth) || (EARLY && path_settled(p, zero_stop));
        if (active && fin) {
            colors[i] = __fmul_rn(p.rr, scale);  // render.cpp:194-196
            colors[n + i] = __fmul_rn(p.rg, scale);
            colors[2 * n + i] = __fmul_rn(p.rb, scale);
            i = inext;
            active = has_next;
            p.ox = nxt[0], p.oy = nxt[1], p.oz = nxt[2], p.dx = nxt[3], p.dy = nxt[4], p.dz = nxt[5];
            p.rr = p.rg = p.rb = 1.0f;
            p.alive = true;
            bounce = 0;
            inext = i + lanes;
            has_next = inext < end;
            if (has_next)
                load_ray(rays, n, inext, nxt);
        }
    }
    if (stats != nullptr) {
        unsigned int w = segs;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            w += __shfl_xor_sync(0xffffffffu, w, o);
        if ((threadIdx.x & 31) == 0)
            atomicAdd(stats, static_cast<unsigned long long>(w));
    }
}

// ---- host side --------------------------------------------------------------------------------------
namespace {

struct DeviceState {
    bool init = false;
    int sm_count = 0;
    int blocks_per_sm[4] = {0, 0, 0, 0};  // [NS8?][EARLY?]
    SceneConst *scene_alias = nullptr;
    int *zero_ok_alias = nullptr;
    cudaEvent_t scene_free = nullptr;  // recorded after the last kernel that reads the staged scene
    cudaStream_t last_stream = nullptr;
    bool have_last = false;
};

constexpr int kMaxDevices = 64;
DeviceState g_dev[kMaxDevices];
std::mutex g_mu;

template <int NS, bool EARLY> cudaError_t occupancy(int *out, size_t smem) {
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, trace_paths_kernel<NS, EARLY>, kTraceThreads, smem);
}

cudaError_t ensure_device_state(DeviceState **out) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess)
        return e;
    if (dev < 0 || dev >= kMaxDevices)
        return cudaErrorInvalidDevice;
    DeviceState &s = g_dev[dev];
    if (!s.init) {
        if ((e = cudaDeviceGetAttribute(&s.sm_count, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess)
            return e;
        if ((e = cudaGetSymbolAddress(reinterpret_cast<void **>(&s.scene_alias), c_scene)) != cudaSuccess)
            return e;
        if ((e = cudaGetSymbolAddress(reinterpret_cast<void **>(&s.zero_ok_alias), c_scene_zero_stop_ok)) != cudaSuccess)
            return e;
        if ((e = cudaEventCreateWithFlags(&s.scene_free, cudaEventDisableTiming)) != cudaSuccess)
            return e;
        s.init = true;
    }
    *out = &s;
    return cudaSuccess;
}

template <int NS, bool EARLY>
cudaError_t launch_trace(DeviceState &s, cudaStream_t stream, const float *rays, float *colors, int64_t n, int64_t first, int64_t count,
                         const PtParams &p, unsigned long long *stats) {
    const size_t smem = sizeof(float4) * 2 * static_cast<size_t>(p.sphere_count);
    int &occ = s.blocks_per_sm[(NS > 0 ? 2 : 0) + (EARLY ? 1 : 0)];
    if (occ == 0 || NS == 0) {
        cudaError_t e = occupancy<NS, EARLY>(&occ, smem);
        if (e != cudaSuccess)
            return e;
        if (occ < 1)
            occ = 1;
    }
    const int64_t need = (count + kTraceThreads - 1) / kTraceThreads;
    const int64_t cap = static_cast<int64_t>(s.sm_count) * occ;
    const int grid = static_cast<int>(need < cap ? need : cap);
    trace_paths_kernel<NS, EARLY><<<grid, kTraceThreads, smem, stream>>>(rays, colors, n, first, count, p.depth, p.sphere_count,
                                                                         p.light_index, p.emission_scale, stats);
    return cudaGetLastError();
}

}  // namespace

cudaError_t trace_paths(cudaStream_t stream, const PtParams &p, const float *rays, const float *spheres, float *colors, int64_t n,
                        int64_t first, int64_t count, unsigned long long *stats) {
    if (count <= 0)
        return cudaSuccess;
    std::lock_guard<std::mutex> lock(g_mu);
    DeviceState *s = nullptr;
    cudaError_t e = ensure_device_state(&s);
    if (e != cudaSuccess)
        return e;
    // The constant-bank scene is a per-device singleton: a launch sequence on another stream must
    // wait until the previous sequence has finished reading it.
    if (s->have_last && s->last_stream != stream) {
        if ((e = cudaStreamWaitEvent(stream, s->scene_free, 0)) != cudaSuccess)
            return e;
    }
    pack_scene_kernel<<<1, 128, 0, stream>>>(spheres, p.sphere_count, p.sphere_stride, s->scene_alias, s->zero_ok_alias);
    if ((e = cudaGetLastError()) != cudaSuccess)
        return e;
    const bool early = !(p.flags & PTB200_F_FIXED_DEPTH);
    if (p.sphere_count == 8)
        e = early ? launch_trace<8, true>(*s, stream, rays, colors, n, first, count, p, stats)
                  : launch_trace<8, false>(*s, stream, rays, colors, n, first, count, p, stats);
    else
        e = early ? launch_trace<0, true>(*s, stream, rays, colors, n, first, count, p, stats)
                  : launch_trace<0, false>(*s, stream, rays, colors, n, first, count, p, stats);
    if (e != cudaSuccess)
        return e;
    if ((e = cudaEventRecord(s->scene_free, stream)) != cudaSuccess)
        return e;
    s->last_stream = stream;
    s->have_last = true;
    return cudaSuccess;
}

}  // namespace ptb200
