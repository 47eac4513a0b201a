// Camera-ray generation on the device: replaces the reference's pure-Python generator
// (scripts/gen_data.py:21-75, 54 k rays/s) while reproducing its binary64 arithmetic op for op, so that the
// float32 rays are bit-identical to rays.bin when fed the same uniforms.
//
// Random numbers: either a replayed stream (`uniforms`, 2 doubles per ray -- the reference's is NumPy's
// legacy MT19937 seeded with 0, gen_data.py:438, produced on the host by ptb200_mt19937_uniforms) or the
// production counter-based generator: Philox4x32-10 (Salmon et al., SC'11) keyed by the 64-bit seed with
// the global path index as counter, each pair of output words turned into a double exactly like NumPy's
// genrand_res53.  Counter-based means any tile of any GPU can generate its rays with no stream hand-off.
#include "philox.h"
#include "pt_host.h"

namespace ptb200 {

struct Camera {
    double pos[3], dir[3], cx[3], cy[3];
};

namespace {

// gen_data.py:24-29; np.linalg.norm = sqrt of a left-to-right 3-term dot.  Host code: built with
// -ffp-contract=off so nothing fuses.
double norm3(const double *v) {
    double s = v[0] * v[0];
    s = s + v[1] * v[1];
    s = s + v[2] * v[2];
    return sqrt(s);
}

Camera make_camera(int w, int h) {
    Camera c;
    const double pos[3] = {50, 52, 295.6};
    const double raw[3] = {0, -0.042612, -1};
    const double nr = norm3(raw);
    for (int i = 0; i < 3; i++) {
        c.pos[i] = pos[i];
        c.dir[i] = raw[i] / nr;
    }
    c.cx[0] = static_cast<double>(w) * 0.5135 / static_cast<double>(h);
    c.cx[1] = 0;
    c.cx[2] = 0;
    const double cr[3] = {c.cx[1] * c.dir[2] - c.cx[2] * c.dir[1], c.cx[2] * c.dir[0] - c.cx[0] * c.dir[2],
                          c.cx[0] * c.dir[1] - c.cx[1] * c.dir[0]};
    const double ncr = norm3(cr);
    for (int i = 0; i < 3; i++)
        c.cy[i] = cr[i] / ncr * 0.5135;
    return c;
}

// tent filter, gen_data.py:37-40: one square root per call (both branches take the root of a value in [0, 1])
__device__ __forceinline__ double tent(double u) {
    const double r = __dmul_rn(2.0, u);
    const bool lo = r < 1.0;
    const double sq = __dsqrt_rn(lo ? r : __dsub_rn(2.0, r));
    return lo ? __dsub_rn(sq, 1.0) : __dsub_rn(1.0, sq);
}

// Correctly rounded a / b given rb = RN(1 / b) (Markstein): q0 = RN(a * rb) is a faithful quotient, the FMA
// residual a - b*q0 is exact, and RN(q0 + r * rb) is the IEEE quotient.  Three DFMA-class instructions instead
// of the ~20 of a full division; bit-identical to NumPy's division (tests compare the rays bit for bit).
__device__ __forceinline__ double div_by(double a, double b, double rb) {
    const double q0 = __dmul_rn(a, rb);
    const double r = __fma_rn(-b, q0, a);
    return __fma_rn(r, rb, q0);
}

// Division of x < 2^31 by an invariant d >= 1: q = umulhi(x, mul) >> shift (Granlund-Montgomery), d == 1 special-cased.
struct FastDiv {
    unsigned int mul, shift, d;
    __host__ FastDiv() : mul(0), shift(0), d(1) {}
    __host__ explicit FastDiv(unsigned int dd) : mul(0), shift(0), d(dd) {
        if (dd > 1) {
            unsigned int lg = 0;
            while ((1u << lg) < dd)
                lg++;
            const unsigned int p = 31 + lg;
            mul = static_cast<unsigned int>(((1ULL << p) + dd - 1) / dd);
            shift = p - 32;
        }
    }
    __device__ __forceinline__ unsigned int div(unsigned int x) const { return d == 1 ? x : (__umulhi(x, mul) >> shift); }
};

struct RayGenArgs {
    Camera cam;
    double w, h, rw, rh;  // image size as doubles and their correctly rounded reciprocals
    int iw, ih, s;
    FastDiv by_spp, by_s, by_h;
};

__global__ void __launch_bounds__(256) gen_rays_kernel(RayGenArgs a, const double *__restrict__ uniforms, uint64_t seed, int64_t path0, int64_t m,
                                                       float *__restrict__ rays, int fast_index, unsigned int pix_base) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= m)
        return;
    // global path index ((((x*H + y)*2 + sy)*2 + sx)*S + k), gen_data.py:32-36
    int sx, sy, x, y;
    if (fast_index) {  // path0 is a whole number of pixels and everything fits 32 bits: no 64-bit divisions
        const unsigned int spp = 4u * static_cast<unsigned int>(a.s);
        const unsigned int ii = static_cast<unsigned int>(i);
        const unsigned int lp = a.by_spp.div(ii);
        const unsigned int pix = pix_base + lp;
        const unsigned int sub = a.by_s.div(ii - lp * spp);
        sx = static_cast<int>(sub & 1u);
        sy = static_cast<int>(sub >> 1);
        x = static_cast<int>(a.by_h.div(pix));
        y = static_cast<int>(pix - static_cast<unsigned int>(x) * static_cast<unsigned int>(a.ih));
    } else {
        int64_t r = (path0 + i) / a.s;
        sx = static_cast<int>(r & 1);
        r >>= 1;
        sy = static_cast<int>(r & 1);
        r >>= 1;
        y = static_cast<int>(r % a.ih);
        x = static_cast<int>(r / a.ih);
    }

    double u1, u2;
    if (uniforms != nullptr) {
        u1 = uniforms[2 * i];
        u2 = uniforms[2 * i + 1];
    } else {
        philox_uniform2(seed, static_cast<uint64_t>(path0 + i), u1, u2);
    }
    const double dx = tent(u1);
    const double dy = tent(u2);
    // ((sx + 0.5 + dx) / 2 + x) / w - 0.5, gen_data.py:41-43  (/2 is an exact scaling)
    const double fx = __dsub_rn(div_by(__dadd_rn(__dmul_rn(__dadd_rn(sx + 0.5, dx), 0.5), static_cast<double>(x)), a.w, a.rw), 0.5);
    const double fy = __dsub_rn(div_by(__dadd_rn(__dmul_rn(__dadd_rn(sy + 0.5, dy), 0.5), static_cast<double>(y)), a.h, a.rh), 0.5);
    double d[3];
#pragma unroll
    for (int c = 0; c < 3; c++)
        d[c] = __dadd_rn(__dadd_rn(__dmul_rn(a.cam.cx[c], fx), __dmul_rn(a.cam.cy[c], fy)), a.cam.dir[c]);
    const double nrm = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(d[0], d[0]), __dmul_rn(d[1], d[1])), __dmul_rn(d[2], d[2])));
    const double rn = __ddiv_rn(1.0, nrm);  // one true division, shared by the three components
#pragma unroll
    for (int c = 0; c < 3; c++) {
        rays[c * m + i] = __double2float_rn(__dadd_rn(a.cam.pos[c], __dmul_rn(d[c], 140.0)));  // gen_data.py:45
        rays[(3 + c) * m + i] = __double2float_rn(div_by(d[c], nrm, rn));                       // gen_data.py:46
    }
}

}  // namespace

cudaError_t gen_rays(cudaStream_t stream, const PtParams &p, const double *uniforms, uint64_t seed, int64_t path0, int64_t m, float *rays) {
    if (m <= 0)
        return cudaSuccess;
    RayGenArgs a;
    a.cam = make_camera(p.width, p.height);
    a.w = static_cast<double>(p.width), a.h = static_cast<double>(p.height);
    a.rw = 1.0 / a.w, a.rh = 1.0 / a.h;
    a.iw = p.width, a.ih = p.height, a.s = p.samples;
    const int64_t spp = 4LL * p.samples;
    a.by_spp = FastDiv(static_cast<unsigned int>(spp));
    a.by_s = FastDiv(static_cast<unsigned int>(p.samples));
    a.by_h = FastDiv(static_cast<unsigned int>(p.height));
    const int fast_index = (path0 % spp == 0) && m < (1LL << 31) && (path0 + m) / spp < (1LL << 31);
    const int64_t blocks = (m + 255) / 256;
    gen_rays_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(a, uniforms, seed, path0, m, rays, fast_index,
                                                                       fast_index ? static_cast<unsigned int>(path0 / spp) : 0u);
    return cudaGetLastError();
}

}  // namespace ptb200
