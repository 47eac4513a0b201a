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
#include <cuda_pipeline.h>

#include "pt_device.cuh"
#include "pt_host.h"
#include "pt_material.cuh"

namespace ptb200 {

// ---- scene staging: device SoA [10][stride] -> constant bank -------------------------------------
// Writes through the global-memory alias of the __constant__ symbols (cudaGetSymbolAddress); the
// constant cache is coherent across kernel launches, and launches on one stream are ordered.
__global__ void pack_scene_kernel(const float *__restrict__ spheres, int nsph, int stride, SceneConst *dst, int *zero_ok) {
    __shared__ int ok;
    if (threadIdx.x == 0)
        ok = 1;
    __syncthreads();
    const int padded = (nsph + 1) & ~1;  // the pairwise tests read one sphere past an odd count
    for (int k = threadIdx.x; k < padded; k += blockDim.x) {
        if (k >= nsph) {  // never-hit sphere: NaN centre -> NaN roots -> both compares false
            dst->nr2[k] = 0.0f;
            dst->cx[k] = dst->cy[k] = dst->cz[k] = __int_as_float(0x7fc00000);
            continue;
        }
        dst->nr2[k] = -spheres[0 * stride + k];
        dst->cx[k] = spheres[1 * stride + k];
        dst->cy[k] = spheres[2 * stride + k];
        dst->cz[k] = spheres[3 * stride + k];
        const float r = spheres[7 * stride + k], g = spheres[8 * stride + k], b = spheres[9 * stride + k];
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
// Plane base pointers (already offset to the first path of the launch) travel as kernel parameters: the
// constant bank feeds IMAD.WIDE directly, so a lane forms an address with one instruction from its 32-bit
// path index instead of a 64-bit add chain.
struct TracePlanes {
    const float *ray[6];
    float *col[3];
};

constexpr int kRingBatches = 4;                  // batches of 32 rays in the ring per warp (power of two)
constexpr int kRing = 32 * kRingBatches;         // ring entries per warp
constexpr int kWarpsPerBlock = kTraceThreads / 32;

template <int NS, bool EARLY>
__global__ void __launch_bounds__(kTraceThreads, 5) trace_paths_kernel(const TracePlanes pl, const float *__restrict__ spheres, unsigned int count,
                                                                    int depth, int nsph, int stride, int light, float scale, float one,
                                                                    unsigned long long *__restrict__ stats) {
    extern __shared__ float4 smem[];
    SceneShared sh;
    stage_scene_shared(smem, spheres, nsph, stride, sh);
    const bool zero_stop = c_scene_zero_stop_ok != 0;

    // Persistent warps with path regeneration.  Each warp owns a contiguous range of paths and hands them
    // out in order: lanes whose path finished in this iteration are ranked by a ballot and take the next
    // consecutive indices, so the rays they fetch and (to within the few paths in flight) the colours they
    // store share cache lines even though the lanes are at different bounces of different paths.
    // The range is streamed through a per-warp shared-memory ring, four coalesced cp.async batches of 32
    // rays ahead of consumption, so a swap costs six LDS instead of an exposed HBM round trip.
    const unsigned int lane = threadIdx.x & 31u;
    const unsigned int warp_in_block = threadIdx.x >> 5;
    const unsigned int warp = blockIdx.x * kWarpsPerBlock + warp_in_block;
    const unsigned int nwarps = gridDim.x * kWarpsPerBlock;
    const unsigned int per = ((count + nwarps - 1) / nwarps + 31u) & ~31u;
    const unsigned long long wb = static_cast<unsigned long long>(warp) * per;
    const unsigned int wbeg = wb < count ? static_cast<unsigned int>(wb) : count;
    const unsigned int wcount = (count - wbeg) < per ? (count - wbeg) : per;  // paths of this warp
    float *ring = reinterpret_cast<float *>(smem + 2 * nsph) + warp_in_block * (6 * kRing);

    auto issue_batch = [&](unsigned int b) {  // warp-uniform: every lane commits a (possibly empty) group
        const unsigned int e = b * 32u + lane;
        if (e < wcount) {
            const unsigned int s = e & (kRing - 1);
#pragma unroll
            for (int c = 0; c < 6; c++)
                __pipeline_memcpy_async(ring + c * kRing + s, pl.ray[c] + wbeg + e, sizeof(float));
        }
        __pipeline_commit();
    };
    unsigned int issued = 0;
    for (; issued < kRingBatches; issued++)
        issue_batch(issued);

    PathState p;
    p.ox = p.oy = p.oz = p.dx = p.dy = 0.0f;
    p.dz = 1.0f;
    p.rr = p.rg = p.rb = 1.0f;
    p.alive = true;
    int bounce = 0;
    unsigned int segs = 0;
    unsigned int head = 0;      // next unassigned path of the warp's range (warp-uniform)
    unsigned int mine = 0;      // offset of the path this lane holds
    bool active = false;        // this lane holds a real path
    bool want = true;           // this lane needs a (new) path

    for (;;) {
        const unsigned int wmask = __ballot_sync(0xffffffffu, want);
        if (wmask != 0u) {  // warp-uniform
            if (active && want) {  // lanes that finished a path in the previous iteration
                const unsigned int i = wbeg + mine;
                pl.col[0][i] = __fmul_rn(p.rr, scale);  // render.cpp:194-196
                pl.col[1][i] = __fmul_rn(p.rg, scale);
                pl.col[2][i] = __fmul_rn(p.rb, scale);
                segs += bounce;
            }
            const unsigned int rank = __popc(wmask & ((1u << lane) - 1u));
            const unsigned int e = head + rank;
            const bool take = want && e < wcount;
            __pipeline_wait_prior(1);  // everything but the newest batch has landed ...
            __syncwarp();              // ... and is visible to the other lanes of the warp
            if (take) {
                const unsigned int s = e & (kRing - 1);
                p.ox = ring[0 * kRing + s], p.oy = ring[1 * kRing + s], p.oz = ring[2 * kRing + s];
                p.dx = ring[3 * kRing + s], p.dy = ring[4 * kRing + s], p.dz = ring[5 * kRing + s];
                mine = e;
            }
            if (want) {
                active = take;
                p.rr = p.rg = p.rb = 1.0f;
                p.alive = true;
                bounce = 0;
                want = false;
            }
            head += __popc(wmask);
            head = head < wcount ? head : wcount;
            __syncwarp();  // ring reads done before a slot can be refilled
            if (issued < head / 32u + kRingBatches) {
                issue_batch(issued);
                issued++;
            }
            if (!__any_sync(0xffffffffu, active))
                break;
        }
        float tmin;
        int idx;
        nearest_hit<NS>(p, nsph, one, kEps, tmin, idx);
        bounce_and_shade<EARLY>(p, tmin, idx, light, sh);
        bounce++;
        want = active && ((bounce >= depth) || (EARLY && path_settled(p, zero_stop)));
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


// ---- material extension kernel (pt_material.cuh) ------------------------------------------------------------------
// Same persistent warps, ring and ballot-ranked regeneration; one iteration = one bounce of material_bounce().
template <int NS, bool BVH>
__global__ void __launch_bounds__(kTraceThreads, 3) trace_materials_kernel(const TracePlanes pl, const float *__restrict__ spheres, unsigned int count,
                                                                           int max_depth, int rr_start, int nsph, int stride, float eps, float one,
                                                                           unsigned long long seed, unsigned long long path0,
                                                                           unsigned long long *__restrict__ stats, const BvhScene bvh) {
    extern __shared__ float4 smem[];
    MatShared sh;
    if (BVH) {
        sh.center = sh.color = sh.emission = nullptr;  // per-sphere data comes from global memory
        __syncthreads();
    } else {
        stage_materials_shared(smem, spheres, nsph, stride, sh);
    }
    const unsigned int lane = threadIdx.x & 31u;
    const unsigned int warp_in_block = threadIdx.x >> 5;
    const unsigned int warp = blockIdx.x * kWarpsPerBlock + warp_in_block;
    const unsigned int nwarps = gridDim.x * kWarpsPerBlock;
    const unsigned int per = ((count + nwarps - 1) / nwarps + 31u) & ~31u;
    const unsigned long long wb = static_cast<unsigned long long>(warp) * per;
    const unsigned int wbeg = wb < count ? static_cast<unsigned int>(wb) : count;
    const unsigned int wcount = (count - wbeg) < per ? (count - wbeg) : per;
    float *ring = reinterpret_cast<float *>(smem + (BVH ? 0 : 3 * nsph)) + warp_in_block * (6 * kRing);

    auto issue_batch = [&](unsigned int b) {
        const unsigned int e = b * 32u + lane;
        if (e < wcount) {
            const unsigned int s = e & (kRing - 1);
#pragma unroll
            for (int c = 0; c < 6; c++)
                __pipeline_memcpy_async(ring + c * kRing + s, pl.ray[c] + wbeg + e, sizeof(float));
        }
        __pipeline_commit();
    };
    unsigned int issued = 0;
    for (; issued < kRingBatches; issued++)
        issue_batch(issued);

    MatPath p;
    p.ox = p.oy = p.oz = p.dx = p.dy = 0.0f;
    p.dz = 1.0f;
    p.tr = p.tg = p.tb = 1.0f;
    p.lr = p.lg = p.lb = 0.0f;
    p.depth = 0;
    unsigned int segs = 0, head = 0, mine = 0;
    bool active = false, want = true;

    for (;;) {
        const unsigned int wmask = __ballot_sync(0xffffffffu, want);
        if (wmask != 0u) {
            if (active && want) {
                const unsigned int i = wbeg + mine;
                pl.col[0][i] = p.lr;
                pl.col[1][i] = p.lg;
                pl.col[2][i] = p.lb;
            }
            const unsigned int rank = __popc(wmask & ((1u << lane) - 1u));
            const unsigned int e = head + rank;
            const bool take = want && e < wcount;
            __pipeline_wait_prior(1);
            __syncwarp();
            if (take) {
                const unsigned int s = e & (kRing - 1);
                p.ox = ring[0 * kRing + s], p.oy = ring[1 * kRing + s], p.oz = ring[2 * kRing + s];
                p.dx = ring[3 * kRing + s], p.dy = ring[4 * kRing + s], p.dz = ring[5 * kRing + s];
                mine = e;
            }
            if (want) {
                active = take;
                p.tr = p.tg = p.tb = 1.0f;
                p.lr = p.lg = p.lb = 0.0f;
                p.depth = 0;
                want = false;
            }
            head += __popc(wmask);
            head = head < wcount ? head : wcount;
            __syncwarp();
            if (issued < head / 32u + kRingBatches) {
                issue_batch(issued);
                issued++;
            }
            if (!__any_sync(0xffffffffu, active))
                break;
        }
        if (active) {  // dead lanes of a finished range idle; live ones diverge by material inside
            segs++;
            const bool ended = material_bounce<NS, BVH>(p, nsph, one, eps, rr_start, seed, path0 + wbeg + mine, sh, bvh);
            want = ended || p.depth >= max_depth;
        }
    }
    if (stats != nullptr) {
        unsigned int w = segs;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            w += __shfl_xor_sync(0xffffffffu, w, o);
        if (lane == 0)
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
cudaError_t launch_trace(DeviceState &s, cudaStream_t stream, const float *rays, const float *spheres, float *colors, int64_t n,
                         int64_t first, int64_t count, const PtParams &p, unsigned long long *stats) {
    const size_t smem = sizeof(float4) * 2 * static_cast<size_t>(p.sphere_count) + sizeof(float) * 6 * kRing * kWarpsPerBlock;
    int &occ = s.blocks_per_sm[(NS > 0 ? 2 : 0) + (EARLY ? 1 : 0)];
    if (occ == 0 || NS == 0) {
        cudaError_t e = occupancy<NS, EARLY>(&occ, smem);
        if (e != cudaSuccess)
            return e;
        if (occ < 1)
            occ = 1;
    }
    const int64_t cap = static_cast<int64_t>(s.sm_count) * occ;
    constexpr int64_t kMaxPerLaunch = 1LL << 30;  // 32-bit path indices inside the kernel
    for (int64_t a = first; a < first + count; a += kMaxPerLaunch) {
        const int64_t m = (first + count - a < kMaxPerLaunch) ? first + count - a : kMaxPerLaunch;
        TracePlanes pl;
        for (int c = 0; c < 6; c++)
            pl.ray[c] = rays + c * n + a;
        for (int c = 0; c < 3; c++)
            pl.col[c] = colors + c * n + a;
        const int64_t need = (m + kTraceThreads - 1) / kTraceThreads;
        const int grid = static_cast<int>(need < cap ? need : cap);
        trace_paths_kernel<NS, EARLY><<<grid, kTraceThreads, smem, stream>>>(pl, spheres, static_cast<unsigned int>(m), p.depth, p.sphere_count,
                                                                             p.sphere_stride, p.light_index, p.emission_scale, 1.0f, stats);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess)
            return e;
    }
    return cudaSuccess;
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
    // Early termination and the fixed-depth loop give identical bits, so which one runs is a pure performance choice:
    // lock-step lanes swap paths once per `depth` iterations, regenerating lanes once per iteration.  Measured on B200
    // (profiles/r1_depth_sweep.md): fixed wins up to depth ~7 (5.85 vs 6.27 ms at depth 5), regeneration beyond
    // (30.7 vs 133.6 ms at depth 50).
    const bool early = !(p.flags & PTB200_F_FIXED_DEPTH) && p.depth > 7;
    if (p.sphere_count == 8)
        e = early ? launch_trace<8, true>(*s, stream, rays, spheres, colors, n, first, count, p, stats)
                  : launch_trace<8, false>(*s, stream, rays, spheres, colors, n, first, count, p, stats);
    else
        e = early ? launch_trace<0, true>(*s, stream, rays, spheres, colors, n, first, count, p, stats)
                  : launch_trace<0, false>(*s, stream, rays, spheres, colors, n, first, count, p, stats);
    if (e != cudaSuccess)
        return e;
    if ((e = cudaEventRecord(s->scene_free, stream)) != cudaSuccess)
        return e;
    s->last_stream = stream;
    s->have_last = true;
    return cudaSuccess;
}


cudaError_t trace_materials(cudaStream_t stream, const PtParams &p_in, const PtMaterialParams &mp, const float *rays, const float *spheres_in,
                            float *colors, int64_t n, int64_t first, int64_t count, uint64_t path0, unsigned long long *stats, const PtBvh *tree) {
    if (count <= 0)
        return cudaSuccess;
    // With a tree, the constant bank and the kernel's brute-force loop see only the huge spheres (compacted SoA, stride 1024).
    PtParams p = p_in;
    const float *spheres = spheres_in;
    BvhScene bvh = {};
    if (tree != nullptr) {
        bvh = bvh_scene(tree);
        spheres = bvh_big_soa(tree);
        p.sphere_count = bvh_big_count(tree);
        p.sphere_stride = 1024;
    }
    std::lock_guard<std::mutex> lock(g_mu);
    DeviceState *s = nullptr;
    cudaError_t e = ensure_device_state(&s);
    if (e != cudaSuccess)
        return e;
    if (s->have_last && s->last_stream != stream) {
        if ((e = cudaStreamWaitEvent(stream, s->scene_free, 0)) != cudaSuccess)
            return e;
    }
    if (p.sphere_count > 0) {
        pack_scene_kernel<<<1, 128, 0, stream>>>(spheres, p.sphere_count, p.sphere_stride, s->scene_alias, s->zero_ok_alias);
        if ((e = cudaGetLastError()) != cudaSuccess)
            return e;
    }
    const bool use_tree = tree != nullptr;
    const size_t smem = (use_tree ? 0 : sizeof(float4) * 3 * static_cast<size_t>(p.sphere_count)) + sizeof(float) * 6 * kRing * kWarpsPerBlock;
    const bool ten = !use_tree && (p.sphere_count == 9 || p.sphere_count == 10);  // smallpt's scene: unrolled pairs (index 9 is padding)
    int occ = 0;
    if ((e = use_tree ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, trace_materials_kernel<0, true>, kTraceThreads, smem)
              : ten   ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, trace_materials_kernel<10, false>, kTraceThreads, smem)
                      : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, trace_materials_kernel<0, false>, kTraceThreads, smem)) != cudaSuccess)
        return e;
    if (occ < 1)
        occ = 1;
    const int64_t cap = static_cast<int64_t>(s->sm_count) * occ;
    constexpr int64_t kMaxPerLaunch = 1LL << 30;
    for (int64_t a = first; a < first + count; a += kMaxPerLaunch) {
        const int64_t m = (first + count - a < kMaxPerLaunch) ? first + count - a : kMaxPerLaunch;
        TracePlanes pl;
        for (int c = 0; c < 6; c++)
            pl.ray[c] = rays + c * n + a;
        for (int c = 0; c < 3; c++)
            pl.col[c] = colors + c * n + a;
        const int64_t need = (m + kTraceThreads - 1) / kTraceThreads;
        const int grid = static_cast<int>(need < cap ? need : cap);
        const unsigned int mm = static_cast<unsigned int>(m);
        const unsigned long long pp = path0 + static_cast<uint64_t>(a - first);
        if (use_tree)
            trace_materials_kernel<0, true><<<grid, kTraceThreads, smem, stream>>>(pl, spheres, mm, mp.max_depth, mp.rr_start, p.sphere_count,
                                                                                   p.sphere_stride, mp.hit_epsilon, 1.0f, mp.seed, pp, stats, bvh);
        else if (ten)
            trace_materials_kernel<10, false><<<grid, kTraceThreads, smem, stream>>>(pl, spheres, mm, mp.max_depth, mp.rr_start, p.sphere_count,
                                                                                     p.sphere_stride, mp.hit_epsilon, 1.0f, mp.seed, pp, stats, bvh);
        else
            trace_materials_kernel<0, false><<<grid, kTraceThreads, smem, stream>>>(pl, spheres, mm, mp.max_depth, mp.rr_start, p.sphere_count,
                                                                                    p.sphere_stride, mp.hit_epsilon, 1.0f, mp.seed, pp, stats, bvh);
        if ((e = cudaGetLastError()) != cudaSuccess)
            return e;
    }
    if ((e = cudaEventRecord(s->scene_free, stream)) != cudaSuccess)
        return e;
    s->last_stream = stream;
    s->have_last = true;
    return cudaSuccess;
}

}  // namespace ptb200
