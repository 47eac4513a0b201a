// Material extension of the radiance kernel: DIFF / SPEC / REFR bounce sampling with Russian
// roulette -- "smallpt in binary32, iterative".  The reference has none of this (its kernel is all-mirror, fixed
// depth, no RNG; it only quotes smallpt's scene table, scripts/gen_data.py:77-89), so parity here is against the
// builder's own CPU specification (same ops in the same order, singly rounded; sin/cos from a fixed polynomial) and,
// statistically, against a binary64 textbook formulation.  DESIGN.md section 8 states the specification.
//
// Structure: the same persistent warps with ballot-ranked path regeneration as the reference-parity kernel; here the
// loop body really diverges (three materials, Russian-roulette exits, refraction branches) and regeneration is what
// keeps the lanes busy: a path that dies is replaced at the next iteration instead of idling until the longest path
// of the warp ends.
#pragma once
#include "philox.h"
#include "pt_bvh.cuh"
#include "pt_device.cuh"

namespace ptb200 {

enum { kMatDiff = 0, kMatSpec = 1, kMatRefr = 2 };

struct MatShared {
    float4 *center;    // x, y, z, material
    float4 *color;     // r, g, b, -
    float4 *emission;  // r, g, b, -
};

__device__ __forceinline__ void stage_materials_shared(float4 *smem, const float *__restrict__ spheres, int nsph, int stride, MatShared &sh) {
    sh.center = smem;
    sh.color = smem + nsph;
    sh.emission = smem + 2 * nsph;
    for (int k = threadIdx.x; k < nsph; k += blockDim.x) {
        sh.center[k] = make_float4(spheres[1 * stride + k], spheres[2 * stride + k], spheres[3 * stride + k], spheres[10 * stride + k]);
        sh.color[k] = make_float4(spheres[7 * stride + k], spheres[8 * stride + k], spheres[9 * stride + k], 0.0f);
        sh.emission[k] = make_float4(spheres[4 * stride + k], spheres[5 * stride + k], spheres[6 * stride + k], 0.0f);
    }
    __syncthreads();
}

// The ten Philox round keys of the material seed, staged per launch next to the scene (trace_materials): the kernels XOR them in
// straight from the constant bank instead of running the key schedule (20 integer adds) at every bounce.
static __constant__ PhiloxKeys c_mat_keys;

struct MatPath {
    float ox, oy, oz, dx, dy, dz;
    float tr, tg, tb;  // throughput
    float lr, lg, lb;  // radiance gathered so far
    int depth;
};

__device__ __forceinline__ float dot_rn(float ax, float ay, float az, float bx, float by, float bz) {
    return __fadd_rn(__fadd_rn(__fmul_rn(ax, bx), __fmul_rn(ay, by)), __fmul_rn(az, bz));
}

// v / |v| with IEEE sqrt and divides (the open-coded correctly rounded forms of pt_device.cuh)
__device__ __forceinline__ void normalize_rn(float &x, float &y, float &z) {
    float ux, uy, uz;
    normalize_fast<true>(x, y, z, dot_rn(x, y, z, x, y, z), ux, uy, uz);
    x = ux, y = uy, z = uz;
}

// sin(2 pi u), cos(2 pi u) for u in [0, 1): nearest quarter turn, exact remainder, fixed polynomials (FMA), rotation.
__device__ __forceinline__ void sincos2pi(float u, float &s_out, float &c_out) {
    const float q = floorf(__fadd_rn(__fmul_rn(u, 4.0f), 0.5f));
    const float r = __fsub_rn(u, __fmul_rn(q, 0.25f));
    const float th = __fmul_rn(6.2831855f, r);
    const float t2 = __fmul_rn(th, th);
    float sp = __fmaf_rn(t2, 2.7557319e-6f, -1.9841270e-4f);
    sp = __fmaf_rn(sp, t2, 8.3333333e-3f);
    sp = __fmaf_rn(sp, t2, -1.6666667e-1f);
    const float s = __fmaf_rn(__fmul_rn(sp, t2), th, th);
    float cp = __fmaf_rn(t2, 2.4801587e-5f, -1.3888889e-3f);
    cp = __fmaf_rn(cp, t2, 4.1666667e-2f);
    cp = __fmaf_rn(cp, t2, -0.5f);
    const float c = __fmaf_rn(cp, t2, 1.0f);
    const int k = static_cast<int>(q) & 3;
    s_out = (k == 0) ? s : (k == 1) ? c : (k == 2) ? -s : -c;
    c_out = (k == 0) ? c : (k == 1) ? -s : (k == 2) ? -c : s;
}

// Per-sphere data of the hit sphere: shared memory (constant-bank scenes) or global memory by original index (BVH scenes).
template <bool BVH> __device__ __forceinline__ void fetch_hit_sphere(int idx, const MatShared &sh, const BvhScene &bvh, float4 &ctr, float4 &col, float4 &emi) {
    if (BVH) {
        const float4 g = __ldg(bvh.geom + idx), cm = __ldg(bvh.color + idx);
        ctr = make_float4(g.x, g.y, g.z, cm.w);
        col = cm;
        emi = __ldg(bvh.emission + idx);
    } else {
        ctr = sh.center[idx];
        col = sh.color[idx];
        emi = sh.emission[idx];
    }
}

// Shading half of a bounce, given the nearest hit (tmin, idx).  Returns true when the path has ended (its radiance is final).
template <bool BVH>
__device__ __forceinline__ bool material_shade(MatPath &p, float tmin, int idx, int rr_start, unsigned long long seed, unsigned long long path,
                                               const MatShared &sh, const BvhScene &bvh) {
    if (!(tmin < kMiss))
        return true;
    uint32_t w[4] = {static_cast<uint32_t>(path), static_cast<uint32_t>(path >> 32), static_cast<uint32_t>(p.depth), 0x4d41u};
    (void)seed;
    philox4x32_10_keyed(w, c_mat_keys);  // == philox4x32_10(w, lo(seed), hi(seed))
    const float u1 = __fmul_rn(static_cast<float>(w[0] >> 8), 5.9604645e-8f), u2 = __fmul_rn(static_cast<float>(w[1] >> 8), 5.9604645e-8f);
    const float u3 = __fmul_rn(static_cast<float>(w[2] >> 8), 5.9604645e-8f), u4 = __fmul_rn(static_cast<float>(w[3] >> 8), 5.9604645e-8f);

    float4 ctr, col, emi;
    fetch_hit_sphere<BVH>(idx, sh, bvh, ctr, col, emi);
    const float xx = __fadd_rn(p.ox, __fmul_rn(p.dx, tmin)), xy = __fadd_rn(p.oy, __fmul_rn(p.dy, tmin)), xz = __fadd_rn(p.oz, __fmul_rn(p.dz, tmin));
    float nx = __fsub_rn(xx, ctr.x), ny = __fsub_rn(xy, ctr.y), nz = __fsub_rn(xz, ctr.z);
    normalize_rn(nx, ny, nz);
    const float dn = dot_rn(nx, ny, nz, p.dx, p.dy, p.dz);
    const bool front = dn < 0.0f;
    const float nlx = front ? nx : -nx, nly = front ? ny : -ny, nlz = front ? nz : -nz;
    p.lr = __fadd_rn(p.lr, __fmul_rn(p.tr, emi.x));
    p.lg = __fadd_rn(p.lg, __fmul_rn(p.tg, emi.y));
    p.lb = __fadd_rn(p.lb, __fmul_rn(p.tb, emi.z));
    float fr = col.x, fg = col.y, fb = col.z;
    float pm = fr > fg ? fr : fg;
    pm = pm > fb ? pm : fb;
    p.depth++;
    if (p.depth > rr_start) {  // Russian roulette
        if (u3 < pm) {
            fr = __fdiv_rn(fr, pm);
            fg = __fdiv_rn(fg, pm);
            fb = __fdiv_rn(fb, pm);
        } else {
            return true;
        }
    }
    p.tr = __fmul_rn(p.tr, fr);
    p.tg = __fmul_rn(p.tg, fg);
    p.tb = __fmul_rn(p.tb, fb);
    const int mat = static_cast<int>(ctr.w);
    // Two of the three materials end in "normalise a vector": the diffuse direction, the refracted direction.  The branches
    // only BUILD that vector (w); one normalisation after them serves both (a warp almost always holds both kinds of lane, and
    // run inside the branches the ~45 instructions were issued twice, the second time for two or three lanes).
    float wx = 0.0f, wy = 0.0f, wz = 1.0f;
    float rx = 0.0f, ry = 0.0f, rz = 0.0f, ddn = 0.0f;
    bool refract = false, into = false;
    if (mat == kMatDiff) {
        float sn, cs;
        sincos2pi(u1, sn, cs);
        const float r2s = __fsqrt_rn(u2);
        float ux, uy, uz;
        if (fabsf(nlx) > 0.1f) {
            ux = nlz, uy = 0.0f, uz = -nlx;
        } else {
            ux = 0.0f, uy = -nlz, uz = nly;
        }
        normalize_rn(ux, uy, uz);
        const float vx = __fsub_rn(__fmul_rn(nly, uz), __fmul_rn(nlz, uy));
        const float vy = __fsub_rn(__fmul_rn(nlz, ux), __fmul_rn(nlx, uz));
        const float vz = __fsub_rn(__fmul_rn(nlx, uy), __fmul_rn(nly, ux));
        const float a = __fmul_rn(cs, r2s), bq = __fmul_rn(sn, r2s), cq = __fsqrt_rn(__fsub_rn(1.0f, u2));
        wx = __fadd_rn(__fadd_rn(__fmul_rn(ux, a), __fmul_rn(vx, bq)), __fmul_rn(nlx, cq));
        wy = __fadd_rn(__fadd_rn(__fmul_rn(uy, a), __fmul_rn(vy, bq)), __fmul_rn(nly, cq));
        wz = __fadd_rn(__fadd_rn(__fmul_rn(uz, a), __fmul_rn(vz, bq)), __fmul_rn(nlz, cq));
    } else {
        const float k2 = __fadd_rn(dn, dn);
        rx = __fsub_rn(p.dx, __fmul_rn(nx, k2)), ry = __fsub_rn(p.dy, __fmul_rn(ny, k2)), rz = __fsub_rn(p.dz, __fmul_rn(nz, k2));
        if (mat != kMatSpec) {  // REFR
            into = dot_rn(nx, ny, nz, nlx, nly, nlz) > 0.0f;
            const float nnt = into ? (1.0f / 1.5f) : 1.5f;
            ddn = dot_rn(p.dx, p.dy, p.dz, nlx, nly, nlz);
            const float cos2t = __fsub_rn(1.0f, __fmul_rn(__fmul_rn(nnt, nnt), __fsub_rn(1.0f, __fmul_rn(ddn, ddn))));
            if (!(cos2t < 0.0f)) {  // else total internal reflection: the mirror direction below
                const float sgn = into ? 1.0f : -1.0f;
                const float kk = __fmul_rn(sgn, __fadd_rn(__fmul_rn(ddn, nnt), __fsqrt_rn(cos2t)));
                wx = __fsub_rn(__fmul_rn(p.dx, nnt), __fmul_rn(nx, kk));
                wy = __fsub_rn(__fmul_rn(p.dy, nnt), __fmul_rn(ny, kk));
                wz = __fsub_rn(__fmul_rn(p.dz, nnt), __fmul_rn(nz, kk));
                refract = true;
            }
        }
    }
    if (mat == kMatDiff || refract)
        normalize_rn(wx, wy, wz);
    if (mat == kMatDiff) {
        p.dx = wx, p.dy = wy, p.dz = wz;
    } else if (!refract) {  // SPEC, or REFR under total internal reflection
        p.dx = rx, p.dy = ry, p.dz = rz;
    } else {
        const float r0 = 0.04f;
        const float c = __fsub_rn(1.0f, into ? -ddn : dot_rn(wx, wy, wz, nx, ny, nz));
        const float c2 = __fmul_rn(c, c);
        const float c5 = __fmul_rn(__fmul_rn(c2, c2), c);
        const float re = __fadd_rn(r0, __fmul_rn(1.0f - 0.04f, c5));
        const float trn = __fsub_rn(1.0f, re);
        const float pr = __fadd_rn(0.25f, __fmul_rn(0.5f, re));
        if (u4 < pr) {
            const float rp = __fdiv_rn(re, pr);
            p.tr = __fmul_rn(p.tr, rp), p.tg = __fmul_rn(p.tg, rp), p.tb = __fmul_rn(p.tb, rp);
            p.dx = rx, p.dy = ry, p.dz = rz;
        } else {
            const float tp = __fdiv_rn(trn, __fsub_rn(1.0f, pr));
            p.tr = __fmul_rn(p.tr, tp), p.tg = __fmul_rn(p.tg, tp), p.tb = __fmul_rn(p.tb, tp);
            p.dx = wx, p.dy = wy, p.dz = wz;
        }
    }
    p.ox = xx, p.oy = xy, p.oz = xz;
    return false;
}

// One bounce of the constant-bank kernel: nearest hit over the nsph spheres of the constant bank, then shading.
template <int NS>
__device__ __forceinline__ bool material_bounce(MatPath &p, int nsph, float one, float eps, int rr_start, unsigned long long seed,
                                                unsigned long long path, const MatShared &sh) {
    PathState ray;
    ray.ox = p.ox, ray.oy = p.oy, ray.oz = p.oz, ray.dx = p.dx, ray.dy = p.dy, ray.dz = p.dz;
    float tmin;
    int idx;
    nearest_hit<NS>(ray, nsph, one, eps, tmin, idx);
    const BvhScene none = {};
    return material_shade<false>(p, tmin, idx, rr_start, seed, path, sh, none);
}

}  // namespace ptb200
