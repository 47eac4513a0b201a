// Resolve on the device: replaces scripts/data_visualization.py:20-59 (decode_color), bit-exactly.
//
// Per pixel and channel: the 4*S samples are 4 sub-pixel runs of S contiguous float32; each run is averaged
// the way np.mean does on a contiguous float32 axis (NumPy's pairwise summation: fewer than 8 elements
// sequentially, up to 128 with eight interleaved accumulators combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7))
// plus a sequential tail, longer runs split at n/2 rounded down to a multiple of 8), divided by S in
// float32; the four means are summed in binary64, divided by 4, clipped to [0,1], scaled by 255 and
// truncated to uint8 (no gamma, no rounding -- data_visualization.py:54-57).
// Row rule: output row r holds image y = H-1-r; identical to the reference's
// writer for square images, and well defined for the non-square ones its writer cannot handle.
#include <cstdlib>

#include "pt_host.h"

namespace ptb200 {
namespace {

// ---- warp-cooperative form -----------------------------------------------------------------------------
// One warp resolves kItems (pixel, channel) pairs at a time (kPixPerWarp consecutive pixels).  For each pair the 4*S samples are 4 runs of S
// contiguous floats; eight lanes own one run and play NumPy's eight interleaved accumulators (lane j sums a[j],
// a[8+j], ... in order), then combine with three xor-shuffles in exactly NumPy's association
// ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)).  Every step of a run reads one full 32-byte sector, and the kItems
// independent streams keep enough loads in flight to run the colour planes through at HBM speed (a single
// stream per warp is latency-bound at ~1 TB/s).
constexpr int kPixPerWarp = 4;
constexpr int kItems = 3 * kPixPerWarp;  // pixels x channels resolved concurrently by one warp

__device__ __forceinline__ void group_block_sum(const float *const (&a)[kItems], int64_t off, int n, int j, float (&r)[kItems]) {
    // 8 <= n <= 128; lanes j = 0..7 of the group; results valid in lane j == 0
#pragma unroll
    for (int u = 0; u < kItems; u++)
        r[u] = a[u][off + j];
    const int lim = n - (n % 8);
    for (int i = 8; i < lim; i += 8) {
        float v[kItems];
#pragma unroll
        for (int u = 0; u < kItems; u++)
            v[u] = a[u][off + i + j];
#pragma unroll
        for (int u = 0; u < kItems; u++)
            r[u] = __fadd_rn(r[u], v[u]);
    }
#pragma unroll
    for (int u = 0; u < kItems; u++) {
        r[u] = __fadd_rn(r[u], __shfl_xor_sync(0xffffffffu, r[u], 1));  // lane0: r0+r1, lane2: r2+r3, ...
        r[u] = __fadd_rn(r[u], __shfl_xor_sync(0xffffffffu, r[u], 2));  // lane0: (r0+r1)+(r2+r3), lane4: (r4+r5)+(r6+r7)
        r[u] = __fadd_rn(r[u], __shfl_xor_sync(0xffffffffu, r[u], 4));  // lane0: the NumPy association
    }
    for (int i = lim; i < n; i++) {  // sequential tail (only lane 0's value is used)
#pragma unroll
        for (int u = 0; u < kItems; u++)
            r[u] = __fadd_rn(r[u], a[u][off + i]);
    }
}

// NumPy pairwise sums of kItems runs of equal length n; control flow depends on n only, so the warp agrees.
__device__ void group_pairwise_sum(const float *const (&a)[kItems], int64_t n, int j, float (&ret)[kItems]) {
    if (n < 8) {
#pragma unroll
        for (int u = 0; u < kItems; u++) {
            float res = 0.0f;
            for (int64_t i = 0; i < n; i++)
                res = __fadd_rn(res, a[u][i]);
            ret[u] = res;
        }
        return;
    }
    if (n <= 128) {
        group_block_sum(a, 0, static_cast<int>(n), j, ret);
        return;
    }
    int64_t off[32], len[32];
    float left[32][kItems];
    int state[32];
    int sp = 0;
    off[0] = 0, len[0] = n, state[0] = 0;
    while (sp >= 0) {
        if (state[sp] == 0) {
            if (len[sp] <= 128) {
                group_block_sum(a, off[sp], static_cast<int>(len[sp]), j, ret);
                sp--;
            } else {
                int64_t n2 = len[sp] / 2;
                n2 -= n2 % 8;
                state[sp] = 1;
                off[sp + 1] = off[sp], len[sp + 1] = n2, state[sp + 1] = 0;
                sp++;
            }
        } else if (state[sp] == 1) {
#pragma unroll
            for (int u = 0; u < kItems; u++)
                left[sp][u] = ret[u];
            int64_t n2 = len[sp] / 2;
            n2 -= n2 % 8;
            state[sp] = 2;
            off[sp + 1] = off[sp] + n2, len[sp + 1] = len[sp] - n2, state[sp + 1] = 0;
            sp++;
        } else {
#pragma unroll
            for (int u = 0; u < kItems; u++)
                ret[u] = __fadd_rn(left[sp][u], ret[u]);
            sp--;
        }
    }
}

__global__ void __launch_bounds__(256) resolve_kernel(const float *__restrict__ colors, int64_t cn, int64_t pix0, int64_t npix, int h, int s,
                                                      uint8_t *__restrict__ image, int x_origin, int img_w, int gamma) {
    const int lane = threadIdx.x & 31;
    const int k = lane >> 3, j = lane & 7;  // sub-pixel run, accumulator
    const int64_t n_items = (npix + kPixPerWarp - 1) / kPixPerWarp;
    const int64_t warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    // grid-stride over warp items: a few thousand resident warps walk the frame (one block per 16 pixels would spend
    // more time launching 400 k blocks than reading the 600 MB)
    for (int64_t w = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; w < n_items; w += warps) {
        const int64_t q0 = w * kPixPerWarp;
        const float *run[kItems];
        int64_t q[kItems];
#pragma unroll
        for (int u = 0; u < kItems; u++) {
            const int64_t qq = q0 + u / 3;
            q[u] = qq < npix ? qq : q0;  // a ragged tail recomputes pixel q0 (result discarded)
            run[u] = colors + (u % 3) * cn + q[u] * 4 * s + static_cast<int64_t>(k) * s;
        }
        float m[kItems];
        if (s == 1) {
#pragma unroll
            for (int u = 0; u < kItems; u++)
                m[u] = run[u][0];
        } else {
            group_pairwise_sum(run, s, j, m);
#pragma unroll
            for (int u = 0; u < kItems; u++)
                m[u] = __fdiv_rn(m[u], static_cast<float>(s));
        }
        // The four sub-pixel means of item u live in lanes 0, 8, 16, 24.  Hand item u to lane u, then run the scalar
        // epilogue ONCE with lanes 0..kItems-1 active (running it per item with one live lane made this kernel
        // instruction-bound: 690 warp instructions per 1.5 KB).
        float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f;
#pragma unroll
        for (int u = 0; u < kItems; u++) {
            const float a0 = __shfl_sync(0xffffffffu, m[u], 0), a1 = __shfl_sync(0xffffffffu, m[u], 8);
            const float a2 = __shfl_sync(0xffffffffu, m[u], 16), a3 = __shfl_sync(0xffffffffu, m[u], 24);
            if (lane == u)
                m0 = a0, m1 = a1, m2 = a2, m3 = a3;
        }
        if (lane < kItems && q0 + lane / 3 < npix) {
            // sum in binary64 in sub-pixel order (data_visualization.py:39-45)
            double v = __ddiv_rn(__dadd_rn(__dadd_rn(__dadd_rn(static_cast<double>(m0), static_cast<double>(m1)), static_cast<double>(m2)),
                                           static_cast<double>(m3)),
                                 4.0);
            v = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
            if (gamma)  // smallpt's display transform (material extension only; the reference has no gamma)
                v = pow(v, 1.0 / 2.2) * 255.0 + 0.5;
            else
                v = __dmul_rn(v, 255.0);
            const int64_t pix = pix0 + q0 + lane / 3;
            const int x = static_cast<int>(pix / h);
            const int y = static_cast<int>(pix - static_cast<int64_t>(x) * h);
            const int row = h - 1 - y;
            image[(static_cast<int64_t>(row) * img_w + (x - x_origin)) * 3 + (lane % 3)] = static_cast<uint8_t>(static_cast<int>(v));
        }
    }
}


// ---- tiled form: 128-bit loads, 128-bit coalesced framebuffer stores -------------------------------------------------
// The warp-cooperative kernel above stores every pixel channel as a lone byte, and because pixels are x-major while the
// image is row-major, neighbouring stores land img_w * 3 bytes apart.  Here a block owns a tile of kTileCols x kTileRows
// pixels, stages the 8-bit results in shared memory and writes whole 48-byte row segments with three 128-bit stores.
// The sample planes are read with 128-bit loads: two lanes own one sub-pixel run, lane `half` holds NumPy's accumulators
// r[4*half .. 4*half+3] in a float4 (block b of the run contributes a[8b + 4*half ..], in order), and the combine
// ((r0+r1)+(r2+r3)) + ((r4+r5)+(r6+r7)) is two in-lane adds and one shuffle.  Needs S % 8 == 0 (every block of NumPy's
// recursion is then a whole number of 8-float groups, no sequential tails) and 16-byte aligned planes; anything else
// takes the kernel above.  Same bits either way (the parity tests cover both).
constexpr int kTileCols = 16, kTileRows = 16;

__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }

// sum of a[off .. off+len) for three planes at once, 8 <= len <= 128, len % 8 == 0: NumPy's unrolled block
__device__ __forceinline__ void pair_block_sum(const float *const (&a)[3], int64_t off, int len, int half, float (&r)[3]) {
    float4 acc[3];
#pragma unroll
    for (int c = 0; c < 3; c++)
        acc[c] = ldg4(a[c] + off + 4 * half);
    for (int i = 8; i < len; i += 8) {
        float4 v[3];
#pragma unroll
        for (int c = 0; c < 3; c++)
            v[c] = ldg4(a[c] + off + i + 4 * half);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            acc[c].x = __fadd_rn(acc[c].x, v[c].x), acc[c].y = __fadd_rn(acc[c].y, v[c].y);
            acc[c].z = __fadd_rn(acc[c].z, v[c].z), acc[c].w = __fadd_rn(acc[c].w, v[c].w);
        }
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float t = __fadd_rn(__fadd_rn(acc[c].x, acc[c].y), __fadd_rn(acc[c].z, acc[c].w));
        r[c] = __fadd_rn(t, __shfl_xor_sync(0xffffffffu, t, 1));  // float addition commutes: both lanes get the same bits
    }
}

__device__ void pair_pairwise_sum(const float *const (&a)[3], int64_t n, int half, float (&ret)[3]) {
    if (n <= 128) {
        pair_block_sum(a, 0, static_cast<int>(n), half, ret);
        return;
    }
    int64_t off[32], len[32];
    float left[32][3];
    int state[32];
    int sp = 0;
    off[0] = 0, len[0] = n, state[0] = 0;
    while (sp >= 0) {
        if (state[sp] == 0) {
            if (len[sp] <= 128) {
                pair_block_sum(a, off[sp], static_cast<int>(len[sp]), half, ret);
                sp--;
            } else {
                int64_t n2 = len[sp] / 2;
                n2 -= n2 % 8;
                state[sp] = 1;
                off[sp + 1] = off[sp], len[sp + 1] = n2, state[sp + 1] = 0;
                sp++;
            }
        } else if (state[sp] == 1) {
#pragma unroll
            for (int c = 0; c < 3; c++)
                left[sp][c] = ret[c];
            int64_t n2 = len[sp] / 2;
            n2 -= n2 % 8;
            state[sp] = 2;
            off[sp + 1] = off[sp] + n2, len[sp + 1] = len[sp] - n2, state[sp + 1] = 0;
            sp++;
        } else {
#pragma unroll
            for (int c = 0; c < 3; c++)
                ret[c] = __fadd_rn(left[sp][c], ret[c]);
            sp--;
        }
    }
}

__global__ void __launch_bounds__(256) resolve_tiles_kernel(const float *__restrict__ colors, int64_t cn, int64_t pix0, int64_t npix, int h, int s,
                                                            uint8_t *__restrict__ image, int x_origin, int img_w, int gamma, int vector_rows) {
    __shared__ __align__(16) uint8_t tile[kTileRows][kTileCols * 3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int p = lane >> 3, k = (lane >> 1) & 3, half = lane & 1;  // pixel of the item, sub-pixel run, accumulator half
    const int xa = static_cast<int>(pix0 / h), xb = static_cast<int>((pix0 + npix - 1) / h);
    const int tiles_y = (h + kTileRows - 1) / kTileRows;
    const int64_t n_tiles = static_cast<int64_t>((xb - xa) / kTileCols + 1) * tiles_y;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int tx = static_cast<int>(t / tiles_y), ty = static_cast<int>(t - static_cast<int64_t>(tx) * tiles_y);
        const int x_tile = xa + tx * kTileCols, y_tile = ty * kTileRows;
        // 64 items of 4 vertically adjacent pixels per tile, 8 per warp; consecutive items walk down a column (contiguous samples)
        for (int it = 0; it < 8; it++) {
            const int item = warp * 8 + it;
            const int col = item >> 2, rg = item & 3;
            const int x = x_tile + col, y0 = y_tile + rg * 4;
            if (x > xb || y0 >= h)  // warp-uniform
                continue;
            const int64_t q = static_cast<int64_t>(x) * h + y0 + p - pix0;  // pixel of this lane, relative to the range
            const bool valid = y0 + p < h && q >= 0 && q < npix;
            const int64_t qc = valid ? q : (static_cast<int64_t>(x) * h + y0 - pix0 >= 0 && static_cast<int64_t>(x) * h + y0 - pix0 < npix
                                                ? static_cast<int64_t>(x) * h + y0 - pix0
                                                : 0);  // an in-range stand-in, result discarded
            const float *run[3];
#pragma unroll
            for (int c = 0; c < 3; c++)
                run[c] = colors + c * cn + (qc * 4 + k) * s;
            float m[3];
            pair_pairwise_sum(run, s, half, m);
#pragma unroll
            for (int c = 0; c < 3; c++)
                m[c] = __fdiv_rn(m[c], static_cast<float>(s));
            // sub-pixel means of pixel p sit in lanes 8p + 2k (+1); lane 8p + c finishes channel c
            float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f;
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const int base = lane & 24;
                const float a0 = __shfl_sync(0xffffffffu, m[c], base), a1 = __shfl_sync(0xffffffffu, m[c], base + 2);
                const float a2 = __shfl_sync(0xffffffffu, m[c], base + 4), a3 = __shfl_sync(0xffffffffu, m[c], base + 6);
                if ((lane & 7) == c)
                    m0 = a0, m1 = a1, m2 = a2, m3 = a3;
            }
            if ((lane & 7) < 3) {
                // sum in binary64 in sub-pixel order (data_visualization.py:39-45)
                double v = __ddiv_rn(__dadd_rn(__dadd_rn(__dadd_rn(static_cast<double>(m0), static_cast<double>(m1)), static_cast<double>(m2)),
                                               static_cast<double>(m3)),
                                     4.0);
                v = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
                if (gamma)  // smallpt's display transform (material extension only; the reference has no gamma)
                    v = pow(v, 1.0 / 2.2) * 255.0 + 0.5;
                else
                    v = __dmul_rn(v, 255.0);
                // image row = h - 1 - y: tile row index counts DOWN the image, i.e. up in y
                tile[kTileRows - 1 - (rg * 4 + p)][col * 3 + (lane & 7)] = static_cast<uint8_t>(static_cast<int>(v));
            }
        }
        __syncthreads();
        // tile row r holds y = y_tile + kTileRows - 1 - r, image row h - 1 - y
        const int64_t first_q = static_cast<int64_t>(x_tile) * h + y_tile - pix0;
        const int64_t last_q = static_cast<int64_t>(x_tile + kTileCols - 1) * h + y_tile + kTileRows - 1 - pix0;
        const bool whole = vector_rows && x_tile + kTileCols - 1 <= xb && y_tile + kTileRows <= h && first_q >= 0 && last_q < npix;
        if (whole) {
            if (threadIdx.x < kTileRows * 3) {
                const int r = threadIdx.x / 3, part = threadIdx.x - r * 3;
                const int y = y_tile + kTileRows - 1 - r;
                uint8_t *dst = image + (static_cast<int64_t>(h - 1 - y) * img_w + (x_tile - x_origin)) * 3 + part * 16;
                *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(&tile[r][part * 16]);
            }
        } else {  // ragged tile (frame edge, range edge, unaligned rows): byte stores of the valid pixels
            const int r = threadIdx.x >> 4, col = threadIdx.x & 15;
            const int x = x_tile + col, y = y_tile + kTileRows - 1 - r;
            const int64_t q = static_cast<int64_t>(x) * h + y - pix0;
            if (x <= xb && y < h && q >= 0 && q < npix) {
                uint8_t *dst = image + (static_cast<int64_t>(h - 1 - y) * img_w + (x - x_origin)) * 3;
                dst[0] = tile[r][col * 3], dst[1] = tile[r][col * 3 + 1], dst[2] = tile[r][col * 3 + 2];
            }
        }
        __syncthreads();
    }
}

// ---- second half of the fused resolve: run means -> pixels -------------------------------------------------------------
// The trace kernels of the production entries already averaged every sub-pixel run (trace_kernels.cu, fuse_reduce_chunk);
// what is left per pixel and channel is data_visualization.py:39-57: the four means summed in binary64, / 4, clip, x 255,
// truncate.  One thread per pixel (its four means are one 128-bit load per plane; consecutive threads walk down a column, so
// loads coalesce), results staged per 16 x 16 tile and written as 128-bit row segments like resolve_tiles_kernel.
__device__ __forceinline__ uint8_t finish_channel(float4 m, int gamma) {
    double v = __ddiv_rn(__dadd_rn(__dadd_rn(__dadd_rn(static_cast<double>(m.x), static_cast<double>(m.y)), static_cast<double>(m.z)),
                                   static_cast<double>(m.w)),
                         4.0);
    v = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
    if (gamma)  // smallpt's display transform (material extension only; the reference has no gamma)
        v = pow(v, 1.0 / 2.2) * 255.0 + 0.5;
    else
        v = __dmul_rn(v, 255.0);
    return static_cast<uint8_t>(static_cast<int>(v));
}

__global__ void __launch_bounds__(256) resolve_means_kernel(const float *__restrict__ means, int64_t n_runs, int64_t pix0, int64_t npix, int h,
                                                            uint8_t *__restrict__ image, int x_origin, int img_w, int gamma, int vector_rows) {
    __shared__ __align__(16) uint8_t tile[kTileRows][kTileCols * 3];
    const int xa = static_cast<int>(pix0 / h), xb = static_cast<int>((pix0 + npix - 1) / h);
    const int tiles_y = (h + kTileRows - 1) / kTileRows;
    const int64_t n_tiles = static_cast<int64_t>((xb - xa) / kTileCols + 1) * tiles_y;
    const int col = threadIdx.x >> 4, rp = threadIdx.x & 15;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int tx = static_cast<int>(t / tiles_y), ty = static_cast<int>(t - static_cast<int64_t>(tx) * tiles_y);
        const int x_tile = xa + tx * kTileCols, y_tile = ty * kTileRows;
        {
            const int x = x_tile + col, y = y_tile + rp;
            const int64_t q = static_cast<int64_t>(x) * h + y - pix0;
            if (x <= xb && y < h && q >= 0 && q < npix) {
#pragma unroll
                for (int c = 0; c < 3; c++)
                    tile[kTileRows - 1 - rp][col * 3 + c] = finish_channel(__ldg(reinterpret_cast<const float4 *>(means + c * n_runs) + q), gamma);
            }
        }
        __syncthreads();
        const int64_t first_q = static_cast<int64_t>(x_tile) * h + y_tile - pix0;
        const int64_t last_q = static_cast<int64_t>(x_tile + kTileCols - 1) * h + y_tile + kTileRows - 1 - pix0;
        const bool whole = vector_rows && x_tile + kTileCols - 1 <= xb && y_tile + kTileRows <= h && first_q >= 0 && last_q < npix;
        if (whole) {
            if (threadIdx.x < kTileRows * 3) {
                const int r = threadIdx.x / 3, part = threadIdx.x - r * 3;
                const int y = y_tile + kTileRows - 1 - r;
                uint8_t *dst = image + (static_cast<int64_t>(h - 1 - y) * img_w + (x_tile - x_origin)) * 3 + part * 16;
                *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(&tile[r][part * 16]);
            }
        } else {  // ragged tile (frame edge, range edge, unaligned rows): byte stores of the valid pixels
            const int r = threadIdx.x >> 4, c2 = threadIdx.x & 15;
            const int x = x_tile + c2, y = y_tile + kTileRows - 1 - r;
            const int64_t q = static_cast<int64_t>(x) * h + y - pix0;
            if (x <= xb && y < h && q >= 0 && q < npix) {
                uint8_t *dst = image + (static_cast<int64_t>(h - 1 - y) * img_w + (x - x_origin)) * 3;
                dst[0] = tile[r][c2 * 3], dst[1] = tile[r][c2 * 3 + 1], dst[2] = tile[r][c2 * 3 + 2];
            }
        }
        __syncthreads();
    }
}

}  // namespace

cudaError_t resolve_means(cudaStream_t stream, const PtParams &p, const float *means, int64_t n_runs, int64_t pix0, int64_t npix, uint8_t *image,
                          int32_t x_origin, int32_t img_w, int gamma) {
    if (npix <= 0)
        return cudaSuccess;
    if (reinterpret_cast<uintptr_t>(means) % 16 != 0 || n_runs % 4 != 0)  // one 128-bit load per pixel and plane
        return cudaErrorInvalidValue;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int xa = static_cast<int>(pix0 / p.height), xb = static_cast<int>((pix0 + npix - 1) / p.height);
    const int64_t n_tiles = static_cast<int64_t>((xb - xa) / kTileCols + 1) * ((p.height + kTileRows - 1) / kTileRows);
    const int vector_rows = reinterpret_cast<uintptr_t>(image) % 16 == 0 && (static_cast<int64_t>(img_w) * 3) % 16 == 0 &&
                            (static_cast<int64_t>(xa - x_origin) * 3) % 16 == 0;
    const int64_t cap_tiles = static_cast<int64_t>(sms) * 8;
    resolve_means_kernel<<<static_cast<unsigned>(n_tiles < cap_tiles ? n_tiles : cap_tiles), 256, 0, stream>>>(means, n_runs, pix0, npix, p.height,
                                                                                                             image, x_origin, img_w, gamma, vector_rows);
    return cudaGetLastError();
}

cudaError_t resolve_pixels(cudaStream_t stream, const PtParams &p, const float *colors, int64_t cn, int64_t pix0, int64_t npix,
                           uint8_t *image, int32_t x_origin, int32_t img_w, int gamma) {
    if (npix <= 0)
        return cudaSuccess;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // tiled kernel: S a multiple of 8 and every plane / image row 16-byte aligned where it uses 128-bit accesses
    static const bool allow_tiles = getenv("PTB200_RESOLVE_TILES") == nullptr || atoi(getenv("PTB200_RESOLVE_TILES")) != 0;
    if (allow_tiles && p.samples % 8 == 0 && reinterpret_cast<uintptr_t>(colors) % 16 == 0 && cn % 4 == 0) {
        const int xa = static_cast<int>(pix0 / p.height), xb = static_cast<int>((pix0 + npix - 1) / p.height);
        const int64_t n_tiles = static_cast<int64_t>((xb - xa) / kTileCols + 1) * ((p.height + kTileRows - 1) / kTileRows);
        // 128-bit row stores need 16-byte aligned row segments: tiles start at columns xa + 16 j of an image whose column 0 is x_origin
        const int vector_rows = reinterpret_cast<uintptr_t>(image) % 16 == 0 && (static_cast<int64_t>(img_w) * 3) % 16 == 0 &&
                                (static_cast<int64_t>(xa - x_origin) * 3) % 16 == 0;
        const int64_t cap_tiles = static_cast<int64_t>(sms) * 8;
        resolve_tiles_kernel<<<static_cast<unsigned>(n_tiles < cap_tiles ? n_tiles : cap_tiles), 256, 0, stream>>>(
            colors, cn, pix0, npix, p.height, p.samples, image, x_origin, img_w, gamma, vector_rows);
        return cudaGetLastError();
    }
    const int64_t need = ((npix + kPixPerWarp - 1) / kPixPerWarp + 7) / 8;  // 8 warps per block
    const int64_t cap = static_cast<int64_t>(sms) * 16;
    const int64_t blocks = need < cap ? need : cap;
    resolve_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(colors, cn, pix0, npix, p.height, p.samples, image, x_origin, img_w, gamma);
    return cudaGetLastError();
}

}  // namespace ptb200
