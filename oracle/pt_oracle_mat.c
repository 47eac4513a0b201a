/* TEST INFRASTRUCTURE ONLY -- CPU twin of the material extension (DIFF / SPEC / REFR with
 * Russian roulette).  PARITY UNPINNED BY THE REFERENCE: the reference kernel has no materials, no emission lookup and
 * no RNG; it only quotes smallpt's scene table (scripts/gen_data.py:77-89).  This file is therefore
 * the builder's own specification, "smallpt in binary32, iterative", written independently of the CUDA code
 * (nothing is shared with ascendpathtracing_b200/), op for op:
 *
 *   per bounce (depth = 0, 1, ...):
 *     nearest hit exactly as the reference does it (rt_helper.h:255-370, 397-451) but with a run-time epsilon: with
 *     1e5-radius walls in binary32 the hit point is off by ~1e-2, and at the reference's 1e-4 a bounced ray re-hits its
 *     own wall from outside and escapes the box (mean radiance +40 %); 0.1 agrees with the binary64 formulation.
 *     Stop if nothing below 1e20
 *     x = o + d*t;  n = (x - c)/|x - c|;  nl = (n.d < 0) ? n : -n
 *     L += T * emission;  f = colour;  p = max(f)
 *     depth++;  if depth > rr_start:  if (u3 < p) f = f / p  else stop          (Russian roulette)
 *     T *= f
 *     DIFF: cosine-weighted bounce about nl from (u1, u2);  SPEC: mirror;  REFR: Fresnel, reflect/refract chosen by u4
 *     o = x
 *   colour = L
 *
 * Random numbers: Philox4x32-10, key = seed, counter = (path_lo, path_hi, bounce, 0x4d41); u_i = (word_i >> 8) * 2^-24.
 * sin/cos of 2*pi*u come from pm_sincos2pi below (quadrant reduction + fixed polynomials evaluated with fmaf), so that the
 * CUDA kernel can be bit-identical to this file; everything else is singly rounded + - * / sqrt in the order written.
 *
 * Independent statistical check: pm_trace_f64 (binary64, libm, the textbook formulation) -- images converge.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

void pto_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]); /* pt_oracle.c */

#define PM_MISS 1e20f
enum { PM_DIFF = 0, PM_SPEC = 1, PM_REFR = 2 };

/* sin(2*pi*u), cos(2*pi*u), u in [0,1): q = nearest quarter turn, r = u - q/4 in [-1/8, 1/8] (exact), th = 2*pi*r,
 * odd/even polynomials in th (Taylor coefficients, degree 9 / 8: error < 4e-8 on [-pi/4, pi/4]), quadrant rotation. */
static void pm_sincos2pi(float u, float *s_out, float *c_out) {
    float q = floorf(u * 4.0f + 0.5f);
    float r = u - q * 0.25f;
    float th = 6.2831855f * r;
    float t2 = th * th;
    float sp = fmaf(t2, 2.7557319e-6f, -1.9841270e-4f);
    sp = fmaf(sp, t2, 8.3333333e-3f);
    sp = fmaf(sp, t2, -1.6666667e-1f);
    float s = fmaf(sp * t2, th, th);
    float cp = fmaf(t2, 2.4801587e-5f, -1.3888889e-3f);
    cp = fmaf(cp, t2, 4.1666667e-2f);
    cp = fmaf(cp, t2, -0.5f);
    float c = fmaf(cp, t2, 1.0f);
    int k = ((int)q) & 3;
    float ss = (k == 0) ? s : (k == 1) ? c : (k == 2) ? -s : -c;
    float cc = (k == 0) ? c : (k == 1) ? -s : (k == 2) ? -c : s;
    *s_out = ss;
    *c_out = cc;
}

static inline float pm_dot(float ax, float ay, float az, float bx, float by, float bz) {
    float s = ax * bx;
    s = s + ay * by;
    s = s + az * bz;
    return s;
}

static inline void pm_normalize(float *x, float *y, float *z) {
    float len = sqrtf(pm_dot(*x, *y, *z, *x, *y, *z));
    *x = *x / len;
    *y = *y / len;
    *z = *z / len;
}

/* spheres: float32 SoA [11][stride]: r^2, x, y, z, ex, ey, ez, cr, cg, cb, material */
static void pm_trace_one(float ox, float oy, float oz, float dx, float dy, float dz, const float *sph, int nsph, int stride, int max_depth,
                         int rr_start, float eps, uint64_t seed, uint64_t path, float *out, uint32_t *segs) {
    const float *R2 = sph, *CX = sph + stride, *CY = sph + 2 * stride, *CZ = sph + 3 * stride;
    const float *EX = sph + 4 * stride, *EY = sph + 5 * stride, *EZ = sph + 6 * stride;
    const float *KR = sph + 7 * stride, *KG = sph + 8 * stride, *KB = sph + 9 * stride, *MAT = sph + 10 * stride;
    float Tr = 1, Tg = 1, Tb = 1, Lr = 0, Lg = 0, Lb = 0;
    int depth = 0;
    uint32_t nseg = 0;
    while (depth < max_depth) {
        /* nearest hit, reference arithmetic */
        float tmin = 0;
        int idx = 0;
        for (int k = 0; k < nsph; k++) {
            float ocx = CX[k] - ox, ocy = CY[k] - oy, ocz = CZ[k] - oz;
            float b = ocx * dx;
            b = b + ocy * dy;
            b = b + ocz * dz;
            float c = ocx * ocx;
            c = c + ocy * ocy;
            c = c + ocz * ocz;
            c = c - R2[k];
            float disc = b * b - c;
            float s = sqrtf(disc);
            float t0 = b - s, t1 = b + s;
            float t = (t0 > eps) ? t0 : t1;
            t = (t > eps) ? t : PM_MISS;
            if (k == 0 || t < tmin) {
                tmin = t;
                idx = k;
            }
        }
        nseg++;
        if (!(tmin < PM_MISS))
            break;
        uint32_t ctr[4] = {(uint32_t)path, (uint32_t)(path >> 32), (uint32_t)depth, 0x4d41u}, key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, w[4];
        pto_philox4x32_10(ctr, key, w);
        float u1 = (float)(w[0] >> 8) * 5.9604645e-8f, u2 = (float)(w[1] >> 8) * 5.9604645e-8f;
        float u3 = (float)(w[2] >> 8) * 5.9604645e-8f, u4 = (float)(w[3] >> 8) * 5.9604645e-8f;

        float xx = ox + dx * tmin, xy = oy + dy * tmin, xz = oz + dz * tmin;
        float nx = xx - CX[idx], ny = xy - CY[idx], nz = xz - CZ[idx];
        pm_normalize(&nx, &ny, &nz);
        float dn = pm_dot(nx, ny, nz, dx, dy, dz);
        float nlx = dn < 0 ? nx : -nx, nly = dn < 0 ? ny : -ny, nlz = dn < 0 ? nz : -nz;
        Lr = Lr + Tr * EX[idx];
        Lg = Lg + Tg * EY[idx];
        Lb = Lb + Tb * EZ[idx];
        float fr = KR[idx], fg = KG[idx], fb = KB[idx];
        float p = fr > fg ? fr : fg;
        p = p > fb ? p : fb;
        depth++;
        if (depth > rr_start) {
            if (u3 < p) {
                fr = fr / p;
                fg = fg / p;
                fb = fb / p;
            } else
                break;
        }
        Tr = Tr * fr;
        Tg = Tg * fg;
        Tb = Tb * fb;
        int mat = (int)MAT[idx];
        if (mat == PM_DIFF) {
            float sn, cs;
            pm_sincos2pi(u1, &sn, &cs);
            float r2s = sqrtf(u2);
            float wx = nlx, wy = nly, wz = nlz;
            /* u = normalize(cross(|wx| > .1 ? (0,1,0) : (1,0,0), w)) */
            float ux, uy, uz;
            if (fabsf(wx) > 0.1f) {
                ux = wz;
                uy = 0.0f;
                uz = -wx;
            } else {
                ux = 0.0f;
                uy = -wz;
                uz = wy;
            }
            pm_normalize(&ux, &uy, &uz);
            /* v = cross(w, u) */
            float vx = wy * uz - wz * uy, vy = wz * ux - wx * uz, vz = wx * uy - wy * ux;
            float a = cs * r2s, bq = sn * r2s, cq = sqrtf(1.0f - u2);
            float ndx = ux * a + vx * bq;
            ndx = ndx + wx * cq;
            float ndy = uy * a + vy * bq;
            ndy = ndy + wy * cq;
            float ndz = uz * a + vz * bq;
            ndz = ndz + wz * cq;
            pm_normalize(&ndx, &ndy, &ndz);
            dx = ndx;
            dy = ndy;
            dz = ndz;
        } else if (mat == PM_SPEC) {
            float k2 = dn + dn;
            dx = dx - nx * k2;
            dy = dy - ny * k2;
            dz = dz - nz * k2;
        } else {
            float k2 = dn + dn;
            float rx = dx - nx * k2, ry = dy - ny * k2, rz = dz - nz * k2; /* reflection */
            int into = pm_dot(nx, ny, nz, nlx, nly, nlz) > 0;
            float nnt = into ? (1.0f / 1.5f) : 1.5f;
            float ddn = pm_dot(dx, dy, dz, nlx, nly, nlz);
            float cos2t = 1.0f - nnt * nnt * (1.0f - ddn * ddn);
            if (cos2t < 0) { /* total internal reflection */
                dx = rx;
                dy = ry;
                dz = rz;
            } else {
                float sgn = into ? 1.0f : -1.0f;
                float kk = sgn * (ddn * nnt + sqrtf(cos2t));
                float tx = dx * nnt - nx * kk, ty = dy * nnt - ny * kk, tz = dz * nnt - nz * kk;
                pm_normalize(&tx, &ty, &tz);
                const float R0 = 0.04f; /* (1.5-1)^2 / (1.5+1)^2 */
                float c = 1.0f - (into ? -ddn : pm_dot(tx, ty, tz, nx, ny, nz));
                float c2 = c * c;
                float c5 = c2 * c2 * c;
                float Re = R0 + (1.0f - R0) * c5;
                float Trn = 1.0f - Re;
                float P = 0.25f + 0.5f * Re;
                if (u4 < P) {
                    float RP = Re / P;
                    Tr = Tr * RP;
                    Tg = Tg * RP;
                    Tb = Tb * RP;
                    dx = rx;
                    dy = ry;
                    dz = rz;
                } else {
                    float TP = Trn / (1.0f - P);
                    Tr = Tr * TP;
                    Tg = Tg * TP;
                    Tb = Tb * TP;
                    dx = tx;
                    dy = ty;
                    dz = tz;
                }
            }
        }
        ox = xx;
        oy = xy;
        oz = xz;
    }
    out[0] = Lr;
    out[1] = Lg;
    out[2] = Lb;
    *segs = nseg;
}

/* rays SoA [6][n], colors SoA [3][n]; path index of element i is path0 + i (the RNG counter). */
uint64_t pm_trace(const float *rays, const float *spheres, float *colors, int64_t n, int nsph, int stride, int max_depth, int rr_start,
                  float eps, uint64_t seed, uint64_t path0) {
    uint64_t total = 0;
#pragma omp parallel for schedule(dynamic, 1024) reduction(+ : total)
    for (int64_t i = 0; i < n; i++) {
        float out[3];
        uint32_t segs;
        pm_trace_one(rays[i], rays[n + i], rays[2 * n + i], rays[3 * n + i], rays[4 * n + i], rays[5 * n + i], spheres, nsph, stride, max_depth,
                     rr_start, eps, seed, path0 + (uint64_t)i, out, &segs);
        colors[i] = out[0];
        colors[n + i] = out[1];
        colors[2 * n + i] = out[2];
        total += segs;
    }
    return total;
}

void pm_sincos2pi_array(const float *u, float *s, float *c, int64_t n) {
    for (int64_t i = 0; i < n; i++)
        pm_sincos2pi(u[i], s + i, c + i);
}

/* smallpt's scene (scripts/gen_data.py:77-89 quotes it): 9 spheres, SoA [11][stride=16] -> 176 floats. */
void pm_smallpt_scene(float *out176) {
    static const double tbl[9][11] = {
        {1e5, 1e5 + 1, 40.8, 81.6, 0, 0, 0, .75, .25, .25, PM_DIFF},   {1e5, -1e5 + 99, 40.8, 81.6, 0, 0, 0, .25, .25, .75, PM_DIFF},
        {1e5, 50, 40.8, 1e5, 0, 0, 0, .75, .75, .75, PM_DIFF},         {1e5, 50, 40.8, -1e5 + 170, 0, 0, 0, 0, 0, 0, PM_DIFF},
        {1e5, 50, 1e5, 81.6, 0, 0, 0, .75, .75, .75, PM_DIFF},         {1e5, 50, -1e5 + 81.6, 81.6, 0, 0, 0, .75, .75, .75, PM_DIFF},
        {16.5, 27, 16.5, 47, 0, 0, 0, .999, .999, .999, PM_SPEC},      {16.5, 73, 16.5, 78, 0, 0, 0, .999, .999, .999, PM_REFR},
        {600, 50, 681.6 - .27, 81.6, 12, 12, 12, 0, 0, 0, PM_DIFF}};
    memset(out176, 0, 176 * sizeof(float));
    for (int i = 0; i < 9; i++)
        for (int m = 0; m < 11; m++)
            out176[m * 16 + i] = (float)(m == 0 ? tbl[i][m] * tbl[i][m] : tbl[i][m]);
}

/* BASELINE config C4 scene (inputs C4): reference walls + light, then n_random spheres from NumPy-legacy
 * MT19937(seed) doubles in the order cx, cy, cz, radius, material, cr, cg, cb.  SoA [11][stride]. */
void pto_mt_doubles(uint32_t seed, int64_t skip, double *out, int64_t n); /* pt_oracle.c */
void pm_random_scene(int n_random, uint32_t seed, int stride, float *out) {
    static const double base[7][11] = {
        {1e5, 1e5 + 1, 40.8, 81.6, 0, 0, 0, 0.435, 0.376, 0.667, 0},  {1e5, -1e5 + 99, 40.8, 81.6, 0, 0, 0, 0.667, 0.129, 0.086, 0},
        {1e5, 50, 40.8, 1e5, 0, 0, 0, 0.270, 0.725, 0.486, 0},        {1e5, 50, 40.8, -1e5 + 170, 0, 0, 0, 0, 0, 0, 0},
        {1e5, 50, 1e5, 81.6, 0, 0, 0, 0.5, 0.5, 0.5, 0},              {1e5, 50, -1e5 + 81.6, 81.6, 0, 0, 0, 0.141, 0.408, 0.635, 0},
        {600, 50, 681.6 - 0.27, 81.6, 12, 12, 12, 0, 0, 0, 0}};
    memset(out, 0, sizeof(float) * 11 * (size_t)stride);
    for (int i = 0; i < 7; i++)
        for (int m = 0; m < 11; m++)
            out[(size_t)m * stride + i] = (float)(m == 0 ? base[i][m] * base[i][m] : base[i][m]);
    double *u = (double *)malloc(sizeof(double) * 8 * (size_t)(n_random > 0 ? n_random : 1));
    pto_mt_doubles(seed, 0, u, 8 * (int64_t)n_random);
    for (int k = 0; k < n_random; k++) {
        const double *q = u + 8 * (size_t)k;
        int i = 7 + k;
        double r = 0.2 + 0.8 * q[3];
        int mat = (int)(3.0 * q[4]);
        if (mat > 2)
            mat = 2;
        double v[11] = {r * r, 1.0 + 98.0 * q[0], 81.6 * q[1], 170.0 * q[2], 0, 0, 0, 0.2 + 0.75 * q[5], 0.2 + 0.75 * q[6], 0.2 + 0.75 * q[7], (double)mat};
        for (int m = 0; m < 11; m++)
            out[(size_t)m * stride + i] = (float)v[m];
    }
    free(u);
}

/* ---- independent binary64 formulation (libm sin/cos, textbook order) for the statistical check ------------------- */
static double rnd01(uint64_t *st) { /* splitmix64: an unrelated generator on purpose */
    uint64_t z = (*st += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    z ^= z >> 31;
    return (double)(z >> 11) / 9007199254740992.0;
}

static void pm_trace_one_f64(double ox, double oy, double oz, double dx, double dy, double dz, const float *sph, int nsph, int stride,
                             int max_depth, int rr_start, uint64_t *rng, double *out) {
    double T[3] = {1, 1, 1}, L[3] = {0, 0, 0};
    for (int depth = 0; depth < max_depth;) {
        double tmin = 1e20;
        int idx = -1;
        for (int k = 0; k < nsph; k++) {
            double opx = sph[stride + k] - ox, opy = sph[2 * stride + k] - oy, opz = sph[3 * stride + k] - oz;
            double b = opx * dx + opy * dy + opz * dz, det = b * b - (opx * opx + opy * opy + opz * opz) + sph[k];
            if (det < 0)
                continue;
            det = sqrt(det);
            double t = b - det > 1e-4 ? b - det : (b + det > 1e-4 ? b + det : 0);
            if (t > 0 && t < tmin) {
                tmin = t;
                idx = k;
            }
        }
        if (idx < 0)
            break;
        double x[3] = {ox + dx * tmin, oy + dy * tmin, oz + dz * tmin};
        double n[3] = {x[0] - sph[stride + idx], x[1] - sph[2 * stride + idx], x[2] - sph[3 * stride + idx]};
        double ln = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        for (int c = 0; c < 3; c++)
            n[c] /= ln;
        double dn = n[0] * dx + n[1] * dy + n[2] * dz;
        double nl[3] = {dn < 0 ? n[0] : -n[0], dn < 0 ? n[1] : -n[1], dn < 0 ? n[2] : -n[2]};
        double f[3];
        for (int c = 0; c < 3; c++) {
            L[c] += T[c] * sph[(4 + c) * stride + idx];
            f[c] = sph[(7 + c) * stride + idx];
        }
        double p = fmax(f[0], fmax(f[1], f[2]));
        if (++depth > rr_start) {
            if (rnd01(rng) < p)
                for (int c = 0; c < 3; c++)
                    f[c] /= p;
            else
                break;
        }
        for (int c = 0; c < 3; c++)
            T[c] *= f[c];
        int mat = (int)sph[10 * stride + idx];
        double nd[3];
        if (mat == PM_DIFF) {
            double r1 = 2 * M_PI * rnd01(rng), r2 = rnd01(rng), r2s = sqrt(r2);
            double u[3];
            if (fabs(nl[0]) > .1) {
                u[0] = nl[2];
                u[1] = 0;
                u[2] = -nl[0];
            } else {
                u[0] = 0;
                u[1] = -nl[2];
                u[2] = nl[1];
            }
            double lu = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
            for (int c = 0; c < 3; c++)
                u[c] /= lu;
            double v[3] = {nl[1] * u[2] - nl[2] * u[1], nl[2] * u[0] - nl[0] * u[2], nl[0] * u[1] - nl[1] * u[0]};
            for (int c = 0; c < 3; c++)
                nd[c] = u[c] * cos(r1) * r2s + v[c] * sin(r1) * r2s + nl[c] * sqrt(1 - r2);
        } else {
            double refl[3] = {dx - n[0] * 2 * dn, dy - n[1] * 2 * dn, dz - n[2] * 2 * dn};
            memcpy(nd, refl, sizeof nd);
            if (mat == PM_REFR) {
                int into = n[0] * nl[0] + n[1] * nl[1] + n[2] * nl[2] > 0;
                double nnt = into ? 1 / 1.5 : 1.5, ddn = dx * nl[0] + dy * nl[1] + dz * nl[2], cos2t = 1 - nnt * nnt * (1 - ddn * ddn);
                if (cos2t >= 0) {
                    double kk = (into ? 1 : -1) * (ddn * nnt + sqrt(cos2t));
                    double t[3] = {dx * nnt - n[0] * kk, dy * nnt - n[1] * kk, dz * nnt - n[2] * kk};
                    double lt = sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
                    for (int c = 0; c < 3; c++)
                        t[c] /= lt;
                    double c1 = 1 - (into ? -ddn : t[0] * n[0] + t[1] * n[1] + t[2] * n[2]);
                    double Re = 0.04 + 0.96 * c1 * c1 * c1 * c1 * c1, P = .25 + .5 * Re;
                    if (rnd01(rng) < P) {
                        for (int c = 0; c < 3; c++)
                            T[c] *= Re / P;
                    } else {
                        for (int c = 0; c < 3; c++)
                            T[c] *= (1 - Re) / (1 - P);
                        memcpy(nd, t, sizeof nd);
                    }
                }
            }
        }
        double lnd = sqrt(nd[0] * nd[0] + nd[1] * nd[1] + nd[2] * nd[2]);
        dx = nd[0] / lnd;
        dy = nd[1] / lnd;
        dz = nd[2] / lnd;
        ox = x[0];
        oy = x[1];
        oz = x[2];
    }
    out[0] = L[0];
    out[1] = L[1];
    out[2] = L[2];
}

void pm_trace_f64(const float *rays, const float *spheres, float *colors, int64_t n, int nsph, int stride, int max_depth, int rr_start,
                  uint64_t seed) {
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t i = 0; i < n; i++) {
        uint64_t rng = seed * 0x2545f4914f6cdd1dull + (uint64_t)i * 0x9e3779b97f4a7c15ull;
        double out[3];
        pm_trace_one_f64(rays[i], rays[n + i], rays[2 * n + i], rays[3 * n + i], rays[4 * n + i], rays[5 * n + i], spheres, nsph, stride,
                         max_depth, rr_start, &rng, out);
        colors[i] = (float)out[0];
        colors[n + i] = (float)out[1];
        colors[2 * n + i] = (float)out[2];
    }
}
