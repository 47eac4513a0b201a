// Camera-ray generation on the device: replaces the reference's pure-Python generator
// (scripts/gen_data.py:21-75, 54 k rays/s) while reproducing its binary64 arithmetic op for op, so that the
// float32 rays are bit-identical to rays.bin when fed the same uniforms.
//
// Random numbers: either a replayed stream (`uniforms`, 2 doubles per ray -- the reference's is NumPy's
// legacy MT19937 seeded with 0, gen_data.py:438, produced on the host by ptb200_mt19937_uniforms) or the
// production counter-based generator: Philox4x32-10 (Salmon et al., SC'11) keyed by the 64-bit seed with
// the global path index as counter, each pair of output words turned into a double exactly like NumPy's
// genrand_res53.  Counter-based means any tile of any GPU can generate its rays with no stream hand-off.
#include "pt_host.h"
#include "pt_raygen.cuh"

namespace ptb200 {
namespace {

__global__ void __launch_bounds__(256) gen_rays_kernel(RayGenSource g, int64_t m, float *__restrict__ rays) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= m)
        return;
    float r[6];
    generate_ray(g, i, r);
#pragma unroll
    for (int c = 0; c < 6; c++)
        rays[c * m + i] = r[c];
}

}  // namespace

cudaError_t gen_rays(cudaStream_t stream, const PtParams &p, const double *uniforms, uint64_t seed, int64_t path0, int64_t m, float *rays) {
    if (m <= 0)
        return cudaSuccess;
    const RayGenSource g = make_raygen_source(p, uniforms, seed, path0, m);
    const int64_t blocks = (m + 255) / 256;
    gen_rays_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(g, m, rays);
    return cudaGetLastError();
}

}  // namespace ptb200
