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

// tent filter, gen_data.py:37-40
__device__ __forceinline__ double tent(double u) {
    const double r = __dmul_rn(2.0, u);
    return r < 1.0 ? __dsub_rn(__dsqrt_rn(r), 1.0) : __dsub_rn(1.0, __dsqrt_rn(__dsub_rn(2.0, r)));
}

__global__ void __launch_bounds__(256) gen_rays_kernel(Camera cam, int w, int h, int s, const double *__restrict__ uniforms, uint64_t seed,
                                                       int64_t path0, int64_t m, float *__restrict__ rays) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= m)
        return;
    const int64_t g = path0 + i;  // global path index ((((x*H + y)*2 + sy)*2 + sx)*S + k), gen_data.py:32-36
    int64_t r = g / s;
    const int sx = static_cast<int>(r & 1);
    r >>= 1;
    const int sy = static_cast<int>(r & 1);
    r >>= 1;
    const int y = static_cast<int>(r % h);
    const int x = static_cast<int>(r / h);

    double u1, u2;
    if (uniforms != nullptr) {
        u1 = uniforms[2 * i];
        u2 = uniforms[2 * i + 1];
    } else {
        philox_uniform2(seed, static_cast<uint64_t>(g), u1, u2);
    }
    const double dx = tent(u1);
    const double dy = tent(u2);
    // ((sx + 0.5 + dx) / 2 + x) / w - 0.5, gen_data.py:41-43
    const double fx = __dsub_rn(__ddiv_rn(__dadd_rn(__ddiv_rn(__dadd_rn(sx + 0.5, dx), 2.0), static_cast<double>(x)), static_cast<double>(w)), 0.5);
    const double fy = __dsub_rn(__ddiv_rn(__dadd_rn(__ddiv_rn(__dadd_rn(sy + 0.5, dy), 2.0), static_cast<double>(y)), static_cast<double>(h)), 0.5);
    double d[3];
#pragma unroll
    for (int c = 0; c < 3; c++)
        d[c] = __dadd_rn(__dadd_rn(__dmul_rn(cam.cx[c], fx), __dmul_rn(cam.cy[c], fy)), cam.dir[c]);
    const double nrm = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(d[0], d[0]), __dmul_rn(d[1], d[1])), __dmul_rn(d[2], d[2])));
#pragma unroll
    for (int c = 0; c < 3; c++) {
        rays[c * m + i] = __double2float_rn(__dadd_rn(cam.pos[c], __dmul_rn(d[c], 140.0)));  // gen_data.py:45
        rays[(3 + c) * m + i] = __double2float_rn(__ddiv_rn(d[c], nrm));                       // gen_data.py:46
    }
}

}  // namespace

cudaError_t gen_rays(cudaStream_t stream, const PtParams &p, const double *uniforms, uint64_t seed, int64_t path0, int64_t m, float *rays) {
    if (m <= 0)
        return cudaSuccess;
    const Camera cam = make_camera(p.width, p.height);
    const int64_t blocks = (m + 255) / 256;
    gen_rays_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(cam, p.width, p.height, p.samples, uniforms, seed, path0, m, rays);
    return cudaGetLastError();
}

}  // namespace ptb200
