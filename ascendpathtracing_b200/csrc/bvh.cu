// GPU build of the sphere BVH of pt_bvh.cuh (LBVH: Morton codes -> radix sort -> Karras 2012 hierarchy -> bottom-up
// boxes), plus the C-ABI handle around it.  No library calls: the one-time sort of the Morton keys is the small stable LSD
// radix sort below (round 1 called cub::DeviceRadixSort here).

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "pt_bvh.cuh"
#include "pt_host.h"

namespace ptb200 {

struct Aabb {
    float lo[3], hi[3];
};

namespace {

// Which axis each of the 30 key bits halves, most significant first.  A fixed x/y/z interleave gives cells with the aspect
// ratio of the scene's bounds; here every bit halves the axis along which the cells are currently longest (host: morton_plan),
// so the cells - and with them the boxes of the upper tree levels - stay as close to cubes as the bounds allow.
constexpr int kKeyBits = 30;
struct MortonPlan {
    unsigned char axis[kKeyBits];
    int nbits[3];  // bits each axis receives in total
};

// ---- stable LSD radix sort of (30-bit key, int value) pairs: three passes of 10 bits ------------------------------------
// One warp owns one tile of consecutive elements.  Pass = count (per-tile histogram) -> scan (exclusive prefix over
// [bin][tile], one block) -> scatter (the warp walks its tile 32 elements at a time IN ORDER; lanes with the same digit are
// ranked by lane with match.any, the lowest of them bumps the tile's running offset for that digit), which keeps equal keys
// in input order like any LSD radix sort must.  Sized for a one-time build of 10^4 .. 10^7 keys, not for throughput.
constexpr int kRadixBits = 10, kRadixBins = 1 << kRadixBits, kRadixPasses = 3;
static_assert(kRadixBits * kRadixPasses >= kKeyBits, "the passes must cover the key");

__global__ void __launch_bounds__(32) radix_count_kernel(const unsigned int *__restrict__ keys, int n, int tile, int shift, unsigned int *hist,
                                                         int n_tiles) {
    __shared__ unsigned int cnt[kRadixBins];
    const int lane = threadIdx.x, t = blockIdx.x;
    for (int b = lane; b < kRadixBins; b += 32)
        cnt[b] = 0u;
    __syncwarp();
    const int t0 = t * tile, t1 = min(n, t0 + tile);
    for (int i = t0 + lane; i < t1; i += 32)
        atomicAdd(&cnt[(keys[i] >> shift) & (kRadixBins - 1)], 1u);
    __syncwarp();
    for (int b = lane; b < kRadixBins; b += 32)
        hist[static_cast<size_t>(b) * n_tiles + t] = cnt[b];
}

// exclusive prefix sum of hist[0 .. m) in place, one block of 1024 threads
__global__ void __launch_bounds__(1024) radix_scan_kernel(unsigned int *hist, int m) {
    __shared__ unsigned int part[1024];
    const int tid = threadIdx.x;
    const int per = (m + 1023) / 1024, a = tid * per, b = min(m, a + per);
    unsigned int sum = 0u;
    for (int i = a; i < b; i++)
        sum += hist[i];
    part[tid] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {  // Hillis-Steele inclusive scan of the 1024 partial sums
        const unsigned int v = tid >= o ? part[tid - o] : 0u;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    unsigned int run = part[tid] - sum;  // exclusive
    for (int i = a; i < b; i++) {
        const unsigned int v = hist[i];
        hist[i] = run;
        run += v;
    }
}

__global__ void __launch_bounds__(32) radix_scatter_kernel(const unsigned int *__restrict__ keys_in, const int *__restrict__ vals_in, int n, int tile,
                                                           int shift, const unsigned int *__restrict__ hist, int n_tiles, unsigned int *keys_out,
                                                           int *vals_out) {
    __shared__ unsigned int base[kRadixBins];
    const int lane = threadIdx.x, t = blockIdx.x;
    for (int b = lane; b < kRadixBins; b += 32)
        base[b] = hist[static_cast<size_t>(b) * n_tiles + t];
    __syncwarp();
    unsigned int lt;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt));
    const int t0 = t * tile, t1 = min(n, t0 + tile);
    for (int c = t0; c < t1; c += 32) {  // warp-uniform
        const int i = c + lane;
        const bool valid = i < t1;
        const unsigned int key = valid ? keys_in[i] : 0u;
        const int val = valid ? vals_in[i] : 0;
        const unsigned int d = valid ? ((key >> shift) & (kRadixBins - 1)) : static_cast<unsigned int>(kRadixBins + lane);  // no peers
        const unsigned int peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        unsigned int off = 0u;
        if (valid && lane == leader) {
            off = base[d];
            base[d] = off + __popc(peers);
        }
        off = __shfl_sync(0xffffffffu, off, leader);
        if (valid) {
            const unsigned int dst = off + __popc(peers & lt);
            keys_out[dst] = key;
            vals_out[dst] = val;
        }
        __syncwarp();
    }
}

struct RadixPlan {
    int tile, n_tiles;
    size_t hist_bytes;
};
static RadixPlan radix_plan(int n) {
    RadixPlan p;
    int tile = (n + 511) / 512;            // at most 512 tiles (the one-block scan walks bins x tiles entries per pass) ...
    tile = (tile + 31) / 32 * 32;
    p.tile = tile < 256 ? 256 : tile;      // ... of at least 256 elements
    p.n_tiles = (n + p.tile - 1) / p.tile;
    p.hist_bytes = sizeof(unsigned int) * kRadixBins * static_cast<size_t>(p.n_tiles);
    return p;
}

// keys_a / vals_a (input, clobbered) -> keys_b / vals_b (sorted): an odd number of passes ends in the b buffers
static cudaError_t radix_sort_pairs(cudaStream_t stream, unsigned int *keys_a, int *vals_a, unsigned int *keys_b, int *vals_b, int n, unsigned int *hist) {
    static_assert(kRadixPasses % 2 == 1, "the result must land in the b buffers");
    const RadixPlan p = radix_plan(n);
    for (int pass = 0; pass < kRadixPasses; pass++) {
        const int shift = pass * kRadixBits;
        const unsigned int *ki = (pass & 1) ? keys_b : keys_a;
        const int *vi = (pass & 1) ? vals_b : vals_a;
        unsigned int *ko = (pass & 1) ? keys_a : keys_b;
        int *vo = (pass & 1) ? vals_a : vals_b;
        radix_count_kernel<<<p.n_tiles, 32, 0, stream>>>(ki, n, p.tile, shift, hist, p.n_tiles);
        radix_scan_kernel<<<1, 1024, 0, stream>>>(hist, kRadixBins * p.n_tiles);
        radix_scatter_kernel<<<p.n_tiles, 32, 0, stream>>>(ki, vi, n, p.tile, shift, hist, p.n_tiles, ko, vo);
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess)
            return e;
    }
    return cudaSuccess;
}

// AoS copies of the per-sphere data (original index order) + Morton keys of the small spheres.
__global__ void prepare_kernel(const float *__restrict__ sph, int n, int stride, float4 *geom, float4 *color, float4 *emission, const int *small_index,
                               int n_small, float3 lo, float3 inv_extent, MortonPlan plan, unsigned int *keys, int *vals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        geom[i] = make_float4(sph[1 * stride + i], sph[2 * stride + i], sph[3 * stride + i], -sph[0 * stride + i]);
        color[i] = make_float4(sph[7 * stride + i], sph[8 * stride + i], sph[9 * stride + i], sph[10 * stride + i]);
        emission[i] = make_float4(sph[4 * stride + i], sph[5 * stride + i], sph[6 * stride + i], 0.0f);
    }
    if (i < n_small) {
        const int k = small_index[i];
        const float x = (sph[1 * stride + k] - lo.x) * inv_extent.x, y = (sph[2 * stride + k] - lo.y) * inv_extent.y,
                    z = (sph[3 * stride + k] - lo.z) * inv_extent.z;
        const float f[3] = {x, y, z};
        unsigned int q[3];
        int left[3];
        for (int c = 0; c < 3; c++) {
            const float cells = static_cast<float>(1u << plan.nbits[c]);
            q[c] = static_cast<unsigned int>(fminf(fmaxf(f[c] * cells, 0.0f), cells - 1.0f));
            left[c] = plan.nbits[c];
        }
        unsigned int key = 0;
        for (int bit = 0; bit < kKeyBits; bit++) {
            const int c = plan.axis[bit];
            key = (key << 1) | ((q[c] >> --left[c]) & 1u);
        }
        keys[i] = key;
        vals[i] = k;
    }
}

__device__ __forceinline__ int delta(const unsigned int *keys, int n, int i, int j) {
    if (j < 0 || j >= n)
        return -1;
    const unsigned int a = keys[i], b = keys[j];
    if (a == b)
        return 32 + __clz(static_cast<unsigned int>(i) ^ static_cast<unsigned int>(j));
    return __clz(a ^ b);
}

// Karras 2012, one thread per internal node; leaves are referenced as ~(original sphere index).
__global__ void hierarchy_kernel(const unsigned int *__restrict__ keys, const int *__restrict__ vals, int n, BvhNode *nodes, int *leaf_parent) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1)
        return;
    const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin)
        lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (delta(keys, n, i, i + (l + t) * d) > dmin)
            l += t;
    const int j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    int s = 0, t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(keys, n, i, i + (s + t) * d) > dnode)
            s += t;
    } while (t > 1);
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    if (lo == gamma) {
        nodes[i].left = ~vals[gamma];
        leaf_parent[gamma] = i;
    } else {
        nodes[i].left = gamma;
        nodes[gamma].parent = i;
    }
    if (hi == gamma + 1) {
        nodes[i].right = ~vals[gamma + 1];
        leaf_parent[gamma + 1] = i;
    } else {
        nodes[i].right = gamma + 1;
        nodes[gamma + 1].parent = i;
    }
    if (i == 0)
        nodes[0].parent = -1;
}

__device__ __forceinline__ Aabb sphere_box(const float4 g) {
    const float r2 = -g.w;
    const float r = sqrtf(fmaxf(r2, 0.0f));
    const float pad = 1e-3f * r + 1e-3f;  // slack for the slab arithmetic only; the exactness margin is per ray (pt_bvh.cuh)
    Aabb b;
    b.lo[0] = g.x - r - pad, b.lo[1] = g.y - r - pad, b.lo[2] = g.z - r - pad;
    b.hi[0] = g.x + r + pad, b.hi[1] = g.y + r + pad, b.hi[2] = g.z + r + pad;
    return b;
}

__device__ __forceinline__ Aabb merge(const Aabb &a, const Aabb &b) {
    Aabb m;
    for (int c = 0; c < 3; c++) {
        m.lo[c] = fminf(a.lo[c], b.lo[c]);
        m.hi[c] = fmaxf(a.hi[c], b.hi[c]);
    }
    return m;
}

// One thread per leaf climbs towards the root; the second thread to reach a node (atomic counter) owns it, so both
// children are complete by then.  own[] receives every internal node's box.
__global__ void fit_kernel(const int *__restrict__ vals, const int *__restrict__ leaf_parent, int n, const float4 *__restrict__ geom,
                           BvhNode *nodes, Aabb *own, unsigned int *visits) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    int node = leaf_parent[i];
    while (node >= 0) {
        if (atomicAdd(&visits[node], 1u) == 0u)
            return;  // first arrival: the sibling subtree is not done yet
        __threadfence();
        const int l = nodes[node].left, r = nodes[node].right;
        const Aabb bl = l < 0 ? sphere_box(geom[~l]) : own[l];
        const Aabb br = r < 0 ? sphere_box(geom[~r]) : own[r];
        nodes[node].a = make_float4(bl.lo[0], bl.lo[1], bl.lo[2], bl.hi[0]);
        nodes[node].b = make_float4(bl.hi[1], bl.hi[2], br.lo[0], br.lo[1]);
        nodes[node].c = make_float4(br.lo[2], br.hi[0], br.hi[1], br.hi[2]);
        own[node] = merge(bl, br);
        __threadfence();
        node = nodes[node].parent;
    }
}

// Float child boxes -> the 32-byte traversal nodes of pt_bvh.cuh: 16-bit grid planes, rounded outwards and grown by
// kGridGrow units (the slack the one-FFMA slab test needs, see pt_bvh.cuh).  The grid was sized on the host so that no
// plane ever clamps.
struct GridMap {
    float lo[3], scale[3];
};

__device__ __forceinline__ unsigned int quantize_pair(float vlo, float vhi, float glo, float gs) {
    const float a = floorf((vlo - glo) * gs) - static_cast<float>(kGridGrow);
    const float b = ceilf((vhi - glo) * gs) + static_cast<float>(kGridGrow);
    const unsigned int qa = static_cast<unsigned int>(fminf(fmaxf(a, 0.0f), kGridMax));
    const unsigned int qb = static_cast<unsigned int>(fminf(fmaxf(b, 0.0f), kGridMax));
    return qa | (qb << 16);
}

__global__ void quantize_kernel(const BvhNode *__restrict__ nodes, int n_nodes, GridMap g, QNode *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes)
        return;
    const BvhNode nd = nodes[i];
    QNode q;
    q.lx = quantize_pair(nd.a.x, nd.a.w, g.lo[0], g.scale[0]);
    q.ly = quantize_pair(nd.a.y, nd.b.x, g.lo[1], g.scale[1]);
    q.lz = quantize_pair(nd.a.z, nd.b.y, g.lo[2], g.scale[2]);
    q.left = nd.left;
    q.rx = quantize_pair(nd.b.z, nd.c.y, g.lo[0], g.scale[0]);
    q.ry = quantize_pair(nd.b.w, nd.c.z, g.lo[1], g.scale[1]);
    q.rz = quantize_pair(nd.c.x, nd.c.w, g.lo[2], g.scale[2]);
    q.right = nd.right;
    out[i] = q;
}

// Nearest hit only (tests / diagnostics): big spheres brute force + tree, the reference's (t, index) semantics.
__global__ void first_hit_kernel(BvhScene sc, const float *__restrict__ rays, int64_t n, float eps, float *__restrict__ tmin_out,
                                 int *__restrict__ idx_out) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const float ox = rays[i], oy = rays[n + i], oz = rays[2 * n + i], dx = rays[3 * n + i], dy = rays[4 * n + i], dz = rays[5 * n + i];
    float tmin = kMiss;
    int idx = 0;
    for (int k = 0; k < sc.n_big; k++) {
        const int s = sc.big_index[k];
        const float4 g = sc.geom[s];
        const float t = sphere_t(ox, oy, oz, dx, dy, dz, g.x, g.y, g.z, g.w, eps);
        if (t < tmin) {  // ascending original index: strict < keeps the lowest index on ties
            tmin = t;
            idx = s;
        }
    }
    bvh_nearest(sc, ox, oy, oz, dx, dy, dz, eps, tmin, idx);
    tmin_out[i] = tmin;
    idx_out[i] = idx;
}

}  // namespace
}  // namespace ptb200

using namespace ptb200;

struct PtBvh {
    int device = 0;
    int n = 0, stride = 0, n_big = 0, n_small = 0;
    float centre[3] = {0, 0, 0}, radius = 0.0f, rmin = 0.0f, rmax = 0.0f;
    QNode *qnodes = nullptr;
    int *small_index = nullptr;
    float glo[3] = {0, 0, 0}, gscale[3] = {1, 1, 1};
    float4 *geom = nullptr, *color = nullptr, *emission = nullptr;
    int *big_index = nullptr;
    float *big_soa = nullptr;  // [11][1024] SoA of the big spheres for the constant-bank pack
    void *block = nullptr;     // the ONE device allocation all of the arrays above are carved from
    int only_leaf = 0;
    BvhScene scene() const {
        BvhScene s;
        s.qnodes = qnodes, s.geom = geom, s.color = color, s.emission = emission, s.big_index = big_index, s.small_index = small_index;
        s.n_big = n_big, s.n_small = n_small, s.root = 0, s.only_leaf = only_leaf;
        for (int c = 0; c < 3; c++)
            s.glo[c] = glo[c], s.gscale[c] = gscale[c], s.centre[c] = centre[c];
        s.radius = radius, s.rmin = rmin, s.rmax = rmax;
        return s;
    }
};

namespace ptb200 {
BvhScene bvh_scene(const PtBvh *b) { return b->scene(); }
const float *bvh_big_soa(const PtBvh *b) { return b->big_soa; }
int bvh_big_count(const PtBvh *b) { return b->n_big; }
int bvh_sphere_count(const PtBvh *b) { return b->n; }
}  // namespace ptb200

extern "C" {

int ptb200_bvh_destroy(PtBvh *b) {
    if (b == nullptr)
        return PTB200_OK;
    cudaFree(b->block);
    delete b;
    return PTB200_OK;
}

int ptb200_bvh_build(const uint8_t *spheres_, int32_t count, int32_t stride, void *stream_, PtBvh **out) {
    if (out == nullptr)
        return fail(PTB200_EINVAL, "ptb200_bvh_build: NULL out");
    *out = nullptr;
    if (spheres_ == nullptr || count < 1 || stride < count)
        return fail(PTB200_EINVAL, "ptb200_bvh_build: need spheres, count >= 1, stride >= count");
    if (count > (1 << 24))
        return fail(PTB200_EINVAL, "ptb200_bvh_build: at most 2^24 spheres");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        cudaGetLastError();
        return fail(PTB200_ENODEV, "ptb200_bvh_build: no CUDA device (this library has no CPU fallback)");
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const float *sph = reinterpret_cast<const float *>(spheres_);
    // Classification and bounds on the host (4 rows of the SoA, 16 bytes per sphere); the tree is built on the device.
    std::vector<float> rows(static_cast<size_t>(4) * count);
    cudaError_t e = cudaSuccess;
    for (int m = 0; m < 4 && e == cudaSuccess; m++)
        e = cudaMemcpyAsync(rows.data() + static_cast<size_t>(m) * count, sph + static_cast<size_t>(m) * stride, sizeof(float) * count,
                            cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess)
        return fail_cuda(e, "ptb200_bvh_build");
    std::vector<int> big, small;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = 0; i < count; i++) {
        const float r2 = rows[i];
        if (!(r2 < kBigR2)) {  // huge or NaN radius: brute force
            big.push_back(i);
            continue;
        }
        small.push_back(i);
        const float r = std::sqrt(std::max(r2, 0.0f));
        for (int c = 0; c < 3; c++) {
            const float v = rows[static_cast<size_t>(1 + c) * count + i];
            lo[c] = std::min(lo[c], v - r);
            hi[c] = std::max(hi[c], v + r);
        }
    }
    if (big.size() > 1024)
        return fail(PTB200_EINVAL, "ptb200_bvh_build: %zu spheres of radius >= 100 (at most 1024 fit the brute-force list)", big.size());
    PtBvh *b = new PtBvh();
    cudaGetDevice(&b->device);
    b->n = count, b->stride = stride, b->n_big = static_cast<int>(big.size()), b->n_small = static_cast<int>(small.size());
    float ext[3] = {1.0f, 1.0f, 1.0f};
    if (!small.empty()) {
        double rad2 = 0.0, rmin = INFINITY, rmax = 0.0;
        for (int c = 0; c < 3; c++) {
            const double dlt = static_cast<double>(hi[c]) - lo[c];
            ext[c] = static_cast<float>(dlt > 0 ? dlt : 1.0);
            b->centre[c] = static_cast<float>(0.5 * (static_cast<double>(hi[c]) + lo[c]));
            rad2 += 0.25 * dlt * dlt;
        }
        for (int i : small) {
            const double r = std::sqrt(std::max(static_cast<double>(rows[i]), 0.0));
            rmin = std::min(rmin, r);
            rmax = std::max(rmax, r);
        }
        // bounding ball of all boxes (geometric boxes + the 1e-3 r + 1e-3 slack), a little generous; radii rounded the safe way
        b->radius = static_cast<float>(std::sqrt(rad2) * 1.0001 + 0.5);
        b->rmin = static_cast<float>(rmin * 0.9999);
        b->rmax = static_cast<float>(rmax * 1.0001);
        // 16-bit grid of the traversal nodes: covers every box (slack <= 1e-3 * 100 + 1e-3) with room for the outward rounding
        // and the kGridGrow units, so no plane ever clamps.
        const double margin = 1.2;
        for (int c = 0; c < 3; c++) {
            const double glo = static_cast<double>(lo[c]) - margin, ghi = static_cast<double>(hi[c]) + margin;
            b->glo[c] = static_cast<float>(glo);
            b->gscale[c] = static_cast<float>((65535.0 - 4.0 * (kGridGrow + 2)) / (ghi - static_cast<double>(b->glo[c])));
        }
    }

    // Morton key layout: every bit halves the currently longest cell edge (ties: x, y, z).  PTB200_MORTON_INTERLEAVE=1 forces
    // the classic fixed interleave (experiments).
    MortonPlan plan;
    {
        static const bool fixed = [] {
            const char *v = getenv("PTB200_MORTON_INTERLEAVE");
            return v != nullptr && atoi(v) != 0;
        }();
        double cell[3] = {ext[0], ext[1], ext[2]};
        plan.nbits[0] = plan.nbits[1] = plan.nbits[2] = 0;
        for (int bit = 0; bit < kKeyBits; bit++) {
            int c = bit % 3;
            if (!fixed) {
                c = 0;
                for (int a = 1; a < 3; a++)
                    if (cell[a] > cell[c])
                        c = a;
            }
            plan.axis[bit] = static_cast<unsigned char>(c);
            plan.nbits[c]++;
            cell[c] *= 0.5;
        }
    }

    // Device memory: ONE allocation for everything the handle keeps and one block for the build's temporaries, the latter
    // from the library's workspace arena when it can serve it (no driver call at all), else one cudaMalloc.  A build used to
    // make 16 cudaMalloc and 9 cudaFree calls, which on a busy driver cost more than the kernels (2-8 ms for 10^4 spheres,
    // 10-180 ms for 10^5).
    const int ns = b->n_small, nb = b->n_big;
    size_t keep_bytes = 0, temp_bytes = 0;
    auto carve = [](size_t &total, size_t bytes) {  // 256-byte aligned offsets
        const size_t off = total;
        total += (bytes + 255) & ~static_cast<size_t>(255);
        return off;
    };
    const size_t ns1 = static_cast<size_t>(std::max(ns, 1)), nn1 = static_cast<size_t>(std::max(ns - 1, 1));
    const size_t o_geom = carve(keep_bytes, sizeof(float4) * count), o_color = carve(keep_bytes, sizeof(float4) * count),
                 o_emis = carve(keep_bytes, sizeof(float4) * count), o_big = carve(keep_bytes, sizeof(int) * std::max(nb, 1)),
                 o_soa = carve(keep_bytes, sizeof(float) * 11 * 1024), o_qn = carve(keep_bytes, sizeof(QNode) * nn1),
                 o_small = carve(keep_bytes, sizeof(int) * ns1);
    const size_t hist_bytes = ns >= 2 ? radix_plan(ns).hist_bytes : 0;  // the sort's histogram
    const size_t o_nodes = carve(temp_bytes, sizeof(BvhNode) * nn1), o_keys_in = carve(temp_bytes, sizeof(unsigned int) * ns1),
                 o_keys = carve(temp_bytes, sizeof(unsigned int) * ns1), o_vals_in = carve(temp_bytes, sizeof(int) * ns1),
                 o_vals = carve(temp_bytes, sizeof(int) * ns1), o_parent = carve(temp_bytes, sizeof(int) * ns1),
                 o_visits = carve(temp_bytes, sizeof(unsigned int) * ns1), o_own = carve(temp_bytes, sizeof(Aabb) * ns1),
                 o_hist = carve(temp_bytes, hist_bytes ? hist_bytes : 16);
    char *keep = nullptr, *temp = nullptr;
    WsBlock ws;
    if (e == cudaSuccess)
        e = cudaMalloc(reinterpret_cast<void **>(&keep), keep_bytes);
    if (e == cudaSuccess) {
        if (ws_alloc(temp_bytes, &ws) == PTB200_OK)
            temp = static_cast<char *>(ws.ptr);
        else
            e = cudaErrorMemoryAllocation;
    }
    auto release_temp = [&] {
        ws_free(&ws);
        temp = nullptr;
    };
    if (e != cudaSuccess) {
        cudaFree(keep);
        release_temp();
        delete b;
        return e == cudaErrorMemoryAllocation ? fail(PTB200_ENOMEM, "ptb200_bvh_build: out of device memory") : fail_cuda(e, "ptb200_bvh_build");
    }
    b->block = keep;
    b->geom = reinterpret_cast<float4 *>(keep + o_geom), b->color = reinterpret_cast<float4 *>(keep + o_color);
    b->emission = reinterpret_cast<float4 *>(keep + o_emis), b->big_index = reinterpret_cast<int *>(keep + o_big);
    b->big_soa = reinterpret_cast<float *>(keep + o_soa), b->qnodes = reinterpret_cast<QNode *>(keep + o_qn);
    b->small_index = reinterpret_cast<int *>(keep + o_small);
    BvhNode *const d_nodes = reinterpret_cast<BvhNode *>(temp + o_nodes);
    unsigned int *const d_keys_in = reinterpret_cast<unsigned int *>(temp + o_keys_in), *const d_keys = reinterpret_cast<unsigned int *>(temp + o_keys);
    int *const d_vals_in = reinterpret_cast<int *>(temp + o_vals_in), *const d_vals = reinterpret_cast<int *>(temp + o_vals);
    int *const d_leaf_parent = reinterpret_cast<int *>(temp + o_parent);
    unsigned int *const d_visits = reinterpret_cast<unsigned int *>(temp + o_visits);
    Aabb *const d_own = reinterpret_cast<Aabb *>(temp + o_own);
    void *const d_tmp = temp + o_hist;
    int *const d_small = b->small_index;
    if (e == cudaSuccess && nb > 0)
        e = cudaMemcpyAsync(b->big_index, big.data(), sizeof(int) * nb, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess && ns > 0)
        e = cudaMemcpyAsync(d_small, small.data(), sizeof(int) * ns, cudaMemcpyHostToDevice, stream);
    // compacted SoA of the big spheres, stride 1024 rows of the 11-row layout, for pack_scene_kernel / shared staging
    if (e == cudaSuccess)
        e = cudaMemsetAsync(b->big_soa, 0, sizeof(float) * 11 * 1024, stream);
    for (int k = 0; k < nb && e == cudaSuccess; k++)
        e = cudaMemcpy2DAsync(b->big_soa + k, sizeof(float) * 1024, sph + big[k], sizeof(float) * stride, sizeof(float), 11, cudaMemcpyDeviceToDevice,
                              stream);
    if (e == cudaSuccess) {
        const int threads = 256, blocks = (std::max(count, ns) + threads - 1) / threads;
        prepare_kernel<<<blocks, threads, 0, stream>>>(sph, count, stride, b->geom, b->color, b->emission, d_small, ns, make_float3(lo[0], lo[1], lo[2]),
                                                       make_float3(1.0f / ext[0], 1.0f / ext[1], 1.0f / ext[2]), plan, d_keys_in, d_vals_in);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess && ns == 1) {
        b->only_leaf = ~small[0];
    }
    if (e == cudaSuccess && ns >= 2) {
        e = radix_sort_pairs(stream, d_keys_in, d_vals_in, d_keys, d_vals, ns, static_cast<unsigned int *>(d_tmp));
        if (e == cudaSuccess)
            e = cudaMemsetAsync(d_visits, 0, sizeof(unsigned int) * ns, stream);
        if (e == cudaSuccess) {
            hierarchy_kernel<<<(ns + 255) / 256, 256, 0, stream>>>(d_keys, d_vals, ns, d_nodes, d_leaf_parent);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) {
            fit_kernel<<<(ns + 255) / 256, 256, 0, stream>>>(d_vals, d_leaf_parent, ns, b->geom, d_nodes, d_own, d_visits);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) {
            GridMap g;
            for (int c = 0; c < 3; c++)
                g.lo[c] = b->glo[c], g.scale[c] = b->gscale[c];
            quantize_kernel<<<(ns - 1 + 255) / 256, 256, 0, stream>>>(d_nodes, ns - 1, g, b->qnodes);
            e = cudaGetLastError();
        }
    }
    {  // always: the temporaries go back to the arena only once nothing can touch them any more
        const cudaError_t es = cudaStreamSynchronize(stream);
        if (e == cudaSuccess)
            e = es;
    }
    release_temp();
    if (e != cudaSuccess) {
        ptb200_bvh_destroy(b);
        return e == cudaErrorMemoryAllocation ? fail(PTB200_ENOMEM, "ptb200_bvh_build: out of device memory") : fail_cuda(e, "ptb200_bvh_build");
    }
    *out = b;
    return PTB200_OK;
}

int ptb200_bvh_info(const PtBvh *b, int32_t *n_spheres, int32_t *n_big, int32_t *n_small, int32_t *n_nodes) {
    if (b == nullptr)
        return fail(PTB200_EINVAL, "ptb200_bvh_info: NULL handle");
    if (n_spheres)
        *n_spheres = b->n;
    if (n_big)
        *n_big = b->n_big;
    if (n_small)
        *n_small = b->n_small;
    if (n_nodes)
        *n_nodes = b->n_small > 1 ? b->n_small - 1 : 0;
    return PTB200_OK;
}

int ptb200_bvh_first_hit(const PtBvh *b, void *stream, const float *rays, int64_t n, float eps, float *tmin_out, int32_t *idx_out) {
    if (b == nullptr || rays == nullptr || tmin_out == nullptr || idx_out == nullptr || n < 0)
        return fail(PTB200_EINVAL, "ptb200_bvh_first_hit: bad argument");
    if (n == 0)
        return PTB200_OK;
    first_hit_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(b->scene(), rays, n, eps, tmin_out,
                                                                                                         idx_out);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? PTB200_OK : fail_cuda(e, "ptb200_bvh_first_hit");
}

}  // extern "C"
