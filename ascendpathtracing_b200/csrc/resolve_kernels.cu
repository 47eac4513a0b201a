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
#include "pt_host.h"

namespace ptb200 {
namespace {

__device__ __forceinline__ float sum_block_le128(const float *__restrict__ a, int n) {
    // 8 <= n <= 128
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; j++)
        r[j] = a[j];
    int i = 8;
    const int lim = n - (n % 8);
    for (; i < lim; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; j++)
            r[j] = __fadd_rn(r[j], a[i + j]);
    }
    float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])), __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    for (; i < n; i++)
        res = __fadd_rn(res, a[i]);
    return res;
}

// NumPy pairwise sum without recursion: an explicit stack of pending halves; partial sums are combined in
// exactly the order the recursive formulation would (post-order).
__device__ float pairwise_sum(const float *__restrict__ a, int64_t n) {
    if (n < 8) {
        float res = 0.0f;
        for (int64_t i = 0; i < n; i++)
            res = __fadd_rn(res, a[i]);
        return res;
    }
    int64_t off[40], len[40];
    float left[40];
    int state[40];  // 0 = fresh, 1 = left half done, 2 = right half done
    int sp = 0;
    off[0] = 0, len[0] = n, state[0] = 0;
    float ret = 0.0f;
    while (sp >= 0) {
        if (state[sp] == 0) {
            if (len[sp] <= 128) {
                ret = sum_block_le128(a + off[sp], static_cast<int>(len[sp]));
                sp--;
            } else {
                int64_t n2 = len[sp] / 2;
                n2 -= n2 % 8;
                state[sp] = 1;
                off[sp + 1] = off[sp], len[sp + 1] = n2, state[sp + 1] = 0;
                sp++;
            }
        } else if (state[sp] == 1) {
            left[sp] = ret;
            int64_t n2 = len[sp] / 2;
            n2 -= n2 % 8;
            state[sp] = 2;
            off[sp + 1] = off[sp] + n2, len[sp + 1] = len[sp] - n2, state[sp + 1] = 0;
            sp++;
        } else {
            ret = __fadd_rn(left[sp], ret);
            sp--;
        }
    }
    return ret;
}

__global__ void __launch_bounds__(256) resolve_kernel(const float *__restrict__ colors, int64_t cn, int64_t pix0, int64_t npix, int h, int s,
                                                      uint8_t *__restrict__ image, int x_origin, int img_w) {
    const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= npix * 3)
        return;
    const int64_t q = t / 3;  // pixel within the tile
    const int c = static_cast<int>(t - q * 3);
    const float *px = colors + c * cn + q * 4 * s;
    double sum = 0.0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        float m;
        if (s == 1)
            m = px[k];
        else if (s <= 128 && s >= 8)
            m = __fdiv_rn(sum_block_le128(px + static_cast<int64_t>(k) * s, s), static_cast<float>(s));
        else
            m = __fdiv_rn(pairwise_sum(px + static_cast<int64_t>(k) * s, s), static_cast<float>(s));
        sum = __dadd_rn(sum, static_cast<double>(m));
    }
    double v = __ddiv_rn(sum, 4.0);
    v = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
    v = __dmul_rn(v, 255.0);
    const int64_t pix = pix0 + q;
    const int x = static_cast<int>(pix / h);
    const int y = static_cast<int>(pix - static_cast<int64_t>(x) * h);
    const int row = h - 1 - y;
    image[(static_cast<int64_t>(row) * img_w + (x - x_origin)) * 3 + c] = static_cast<uint8_t>(static_cast<int>(v));
}

}  // namespace

cudaError_t resolve_pixels(cudaStream_t stream, const PtParams &p, const float *colors, int64_t cn, int64_t pix0, int64_t npix,
                           uint8_t *image, int32_t x_origin, int32_t img_w) {
    if (npix <= 0)
        return cudaSuccess;
    const int64_t blocks = (npix * 3 + 255) / 256;
    resolve_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(colors, cn, pix0, npix, p.height, p.samples, image, x_origin, img_w);
    return cudaGetLastError();
}

}  // namespace ptb200
