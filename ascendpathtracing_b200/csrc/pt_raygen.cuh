// Camera-ray generation shared by the stand-alone generator kernel (raygen_kernels.cu) and the fused
// generate-and-trace kernels (trace_kernels.cu).  Arithmetic: scripts/gen_data.py:21-75 in binary64, op for op
// (see raygen_kernels.cu for the contract); random numbers: replayed uniforms or Philox4x32-10.
#pragma once
#include <cmath>
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/ptb200.h"
#include "philox.h"

namespace ptb200 {

struct Camera {
    double pos[3], dir[3], cx[3], cy[3];
};

// Division of x < 2^31 by an invariant d >= 1: q = umulhi(x, mul) >> shift (Granlund-Montgomery), d == 1 special-cased.
struct FastDiv {  // plain aggregate: lives in kernel parameters and in the constant bank
    unsigned int mul, shift, d;
#ifdef __CUDACC__
    __device__ __forceinline__ unsigned int div(unsigned int x) const { return d == 1 ? x : (__umulhi(x, mul) >> shift); }
#endif
};

inline FastDiv make_fastdiv(unsigned int dd) {
    FastDiv f = {0u, 0u, dd};
    if (dd > 1) {
        unsigned int lg = 0;
        while ((1u << lg) < dd)
            lg++;
        const unsigned int p = 31 + lg;
        f.mul = static_cast<unsigned int>(((1ULL << p) + dd - 1) / dd);
        f.shift = p - 32;
    }
    return f;
}

struct RayGenArgs {
    Camera cam;
    double w, h, rw, rh;  // image size as doubles and their correctly rounded reciprocals
    int iw, ih, s;
    FastDiv by_spp, by_s, by_h;
};

// Where the rays of a launch come from when they are generated on the fly: element i of the launch is global path
// path0 + i of the frame; uniforms (nullable) holds 2 doubles per element, else Philox keyed by seed.
struct RayGenSource {
    RayGenArgs a;
    const double *uniforms;
    unsigned long long seed;
    PhiloxKeys keys;       // the round keys of `seed` (host-computed: no key-schedule instructions per ray)
    long long path0;
    int fast_index;        // path0 is a whole number of pixels and all indices fit 32 bits
    unsigned int pix_base; // path0 / (4*S) when fast_index
    // Column stride (PtParams.column_step > 1): the launch walks a DENSE frame of every x_step-th column starting at x_first
    // -- local column j is image column x_first + j * x_step -- and path0 / pix_base count in that dense frame.  The camera
    // sees the image column, and every random number is keyed by the GLOBAL path index, so a strided render reproduces
    // exactly those columns of the whole frame.  x_step == 1: x_first = 0 and local = global.
    int x_first, x_step;
};

// gen_data.py:24-29; np.linalg.norm = sqrt of a left-to-right 3-term dot.  Host code: built with
// -ffp-contract=off so nothing fuses.
inline double raygen_norm3(const double *v) {
    double s = v[0] * v[0];
    s = s + v[1] * v[1];
    s = s + v[2] * v[2];
    return sqrt(s);
}

inline Camera make_camera(int w, int h) {
    Camera c;
    const double pos[3] = {50, 52, 295.6};
    const double raw[3] = {0, -0.042612, -1};
    const double nr = raygen_norm3(raw);
    for (int i = 0; i < 3; i++) {
        c.pos[i] = pos[i];
        c.dir[i] = raw[i] / nr;
    }
    c.cx[0] = static_cast<double>(w) * 0.5135 / static_cast<double>(h);
    c.cx[1] = 0;
    c.cx[2] = 0;
    const double cr[3] = {c.cx[1] * c.dir[2] - c.cx[2] * c.dir[1], c.cx[2] * c.dir[0] - c.cx[0] * c.dir[2],
                          c.cx[0] * c.dir[1] - c.cx[1] * c.dir[0]};
    const double ncr = raygen_norm3(cr);
    for (int i = 0; i < 3; i++)
        c.cy[i] = cr[i] / ncr * 0.5135;
    return c;
}

inline RayGenSource make_raygen_source(const PtParams &p, const double *uniforms, uint64_t seed, int64_t path0, int64_t m) {
    RayGenSource g;
    g.a.cam = make_camera(p.width, p.height);
    g.a.w = static_cast<double>(p.width), g.a.h = static_cast<double>(p.height);
    g.a.rw = 1.0 / g.a.w, g.a.rh = 1.0 / g.a.h;
    g.a.iw = p.width, g.a.ih = p.height, g.a.s = p.samples;
    const int64_t spp = 4LL * p.samples;
    g.a.by_spp = make_fastdiv(static_cast<unsigned int>(spp));
    g.a.by_s = make_fastdiv(static_cast<unsigned int>(p.samples));
    g.a.by_h = make_fastdiv(static_cast<unsigned int>(p.height));
    g.uniforms = uniforms;
    g.seed = seed;
    g.keys = philox_keys(seed);
    g.path0 = path0;
    g.fast_index = (path0 % spp == 0) && m < (1LL << 31) && (path0 + m) / spp < (1LL << 31);
    g.pix_base = g.fast_index ? static_cast<unsigned int>(path0 / spp) : 0u;
    g.x_first = 0;
    g.x_step = 1;
    return g;
}

// The same generator restricted to elements [off, off + m) of its range (a launch chunk).
inline RayGenSource make_raygen_source_shifted(const RayGenSource &g, int64_t off, int64_t m) {
    RayGenSource s = g;
    const int64_t spp = 4LL * g.a.s;
    s.path0 = g.path0 + off;
    if (g.uniforms != nullptr)
        s.uniforms = g.uniforms + 2 * off;
    s.fast_index = (s.path0 % spp == 0) && m < (1LL << 31) && (s.path0 + m) / spp < (1LL << 31);
    s.pix_base = s.fast_index ? static_cast<unsigned int>(s.path0 / spp) : 0u;
    return s;
}

#ifdef __CUDACC__
// Global path index ((((x*H + y)*2 + sy)*2 + sx)*S + k, gen_data.py:32-36) of element i of a strided launch (fast_index).
__device__ __forceinline__ unsigned long long strided_global_path(const RayGenSource &g, unsigned int i) {
    const RayGenArgs &a = g.a;
    const unsigned int spp = 4u * static_cast<unsigned int>(a.s);
    const unsigned int lp = a.by_spp.div(i);
    const unsigned int pix = g.pix_base + lp;
    const unsigned int xl = a.by_h.div(pix);
    const unsigned int y = pix - xl * static_cast<unsigned int>(a.ih);
    const unsigned long long x = static_cast<unsigned long long>(g.x_first) + static_cast<unsigned long long>(xl) * static_cast<unsigned int>(g.x_step);
    return (x * static_cast<unsigned int>(a.ih) + y) * spp + (i - lp * spp);
}

// tent filter, gen_data.py:37-40, of r = 2 * rand(): one square root per call (both branches take the root of a value in [0, 1])
__device__ __forceinline__ double raygen_tent(double r) {
    const bool lo = r < 1.0;
    const double sq = __dsqrt_rn(lo ? r : __dsub_rn(2.0, r));
    return lo ? __dsub_rn(sq, 1.0) : __dsub_rn(1.0, sq);
}

// Correctly rounded a / b given rb = RN(1 / b) (Markstein): q0 = RN(a * rb) is a faithful quotient, the FMA
// residual a - b*q0 is exact, and RN(q0 + r * rb) is the IEEE quotient.  Three DFMA-class instructions instead
// of the ~20 of a full division; bit-identical to NumPy's division (tests compare the rays bit for bit).
__device__ __forceinline__ double raygen_div_by(double a, double b, double rb) {
    const double q0 = __dmul_rn(a, rb);
    const double r = __fma_rn(-b, q0, a);
    return __fma_rn(r, rb, q0);
}

// Ray of element i of the launch (global path g.path0 + i): out = ox, oy, oz, dx, dy, dz as float32.
__device__ __forceinline__ void generate_ray(const RayGenSource &g, long long i, float (&out)[6]) {
    const RayGenArgs &a = g.a;
    // global path index ((((x*H + y)*2 + sy)*2 + sx)*S + k), gen_data.py:32-36
    int sx, sy, x, y;
    uint64_t gp = static_cast<uint64_t>(g.path0 + i);  // global path index = RNG counter (a strided launch: set below)
    if (g.fast_index) {  // no 64-bit divisions
        const unsigned int spp = 4u * static_cast<unsigned int>(a.s);
        const unsigned int ii = static_cast<unsigned int>(i);
        const unsigned int lp = a.by_spp.div(ii);
        const unsigned int pix = g.pix_base + lp;
        const unsigned int rest = ii - lp * spp;
        const unsigned int sub = a.by_s.div(rest);
        sx = static_cast<int>(sub & 1u);
        sy = static_cast<int>(sub >> 1);
        x = static_cast<int>(a.by_h.div(pix));
        y = static_cast<int>(pix - static_cast<unsigned int>(x) * static_cast<unsigned int>(a.ih));
        x = g.x_first + x * g.x_step;  // dense column -> image column (identity unless strided)
        if (g.x_step > 1)              // the strided launch's global index from the decode above (== strided_global_path(g, i))
            gp = (static_cast<uint64_t>(static_cast<unsigned int>(x)) * static_cast<unsigned int>(a.ih) + static_cast<unsigned int>(y)) * spp + rest;
    } else {
        long long r = (g.path0 + i) / a.s;
        sx = static_cast<int>(r & 1);
        r >>= 1;
        sy = static_cast<int>(r & 1);
        r >>= 1;
        y = static_cast<int>(r % a.ih);
        x = static_cast<int>(r / a.ih);
    }
    double r1, r2;  // 2 * rand(), gen_data.py:37,39
    if (g.uniforms != nullptr) {
        r1 = __dmul_rn(2.0, g.uniforms[2 * i]);
        r2 = __dmul_rn(2.0, g.uniforms[2 * i + 1]);
    } else {
        // rand() = ((a >> 5) * 2^26 + (b >> 6)) / 2^53 = m / 2^53 with the 53-bit integer m = (a >> 5) << 26 | (b >> 6): the sum
        // and the division are exact in binary64, and so is the doubling: 2 * rand() = m * 2^-52.  One integer conversion and
        // one exact scaling replace two conversions, a multiply, an add, a divide and the doubling -- same bits.
        uint32_t c[4] = {static_cast<uint32_t>(gp), static_cast<uint32_t>(gp >> 32), 0u, 0u};
        philox4x32_10_keyed(c, g.keys);
        const unsigned long long m1 = (static_cast<unsigned long long>(c[0] >> 5) << 26) | (c[1] >> 6);
        const unsigned long long m2 = (static_cast<unsigned long long>(c[2] >> 5) << 26) | (c[3] >> 6);
        r1 = __dmul_rn(__ull2double_rn(m1), 0x1p-52);
        r2 = __dmul_rn(__ull2double_rn(m2), 0x1p-52);
    }
    const double dx = raygen_tent(r1);
    const double dy = raygen_tent(r2);
    // ((sx + 0.5 + dx) / 2 + x) / w - 0.5, gen_data.py:41-43  (/2 is an exact scaling)
    const double fx = __dsub_rn(raygen_div_by(__dadd_rn(__dmul_rn(__dadd_rn(sx + 0.5, dx), 0.5), static_cast<double>(x)), a.w, a.rw), 0.5);
    const double fy = __dsub_rn(raygen_div_by(__dadd_rn(__dmul_rn(__dadd_rn(sy + 0.5, dy), 0.5), static_cast<double>(y)), a.h, a.rh), 0.5);
    // d = cx * fx + cy * fy + dir (gen_data.py:41-43) for THIS camera: cx = (cx0, +0, +0), cy = (+0, cy1, cy2), dir = (+0, d1, d2)
    // (make_camera; the zeros are exact, cy1 > 0).  The products with those zeros are signed zeros and adding a zero changes no
    // non-zero value; when the surviving term is itself zero (fx or fy exactly +0, the only zero a "q - 0.5" can round to) every
    // variant of the sum is +0.  So the three components reduce, bit for bit, to:
    double d[3];
    d[0] = __dmul_rn(a.cam.cx[0], fx);
    d[1] = __dadd_rn(__dmul_rn(a.cam.cy[1], fy), a.cam.dir[1]);
    d[2] = __dadd_rn(__dmul_rn(a.cam.cy[2], fy), a.cam.dir[2]);
    const double nrm = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(d[0], d[0]), __dmul_rn(d[1], d[1])), __dmul_rn(d[2], d[2])));
    const double rn = __ddiv_rn(1.0, nrm);  // one true division, shared by the three components
#pragma unroll
    for (int c = 0; c < 3; c++) {
        out[c] = __double2float_rn(__dadd_rn(a.cam.pos[c], __dmul_rn(d[c], 140.0)));  // gen_data.py:45
        out[3 + c] = __double2float_rn(raygen_div_by(d[c], nrm, rn));                  // gen_data.py:46
    }
}
#endif  // __CUDACC__

}  // namespace ptb200
