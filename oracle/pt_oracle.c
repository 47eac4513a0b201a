/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's path for the parity tests.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this file's library; the product (ascendpathtracing_b200/) never does and has no CPU fallback.
 *
 * Every function restates, op for op and rounding for rounding, a piece of
 * KVM-Explorer/AscendPathTracing (citations are file:line into /root/reference):
 *
 *   pto_trace          src/render.cpp:104-207 (Compute) with src/rt_helper.h:255-370 (SphereHitInfo),
 *                      :397-451 (ReduceMinInfo), :504-709 (GenerateNewRays), :711-830
 *                      (AccumulateIntervalColor)
 *   pto_mt_*           numpy.random.seed / numpy.random.rand of scripts/gen_data.py:37-39,438
 *                      (NumPy legacy MT19937: init_genrand + genrand_res53)
 *   pto_camera / pto_gen_rays   scripts/gen_data.py:21-75 (gen_rays)
 *   pto_gen_spheres    scripts/gen_data.py:92-132 (gen_spheres)
 *   pto_resolve        scripts/data_visualization.py:20-59 (decode_color) + :11-17 (write_ppm) with the
 *                      row/column rule of SURVEY.md section 5 for non-square images
 *
 * Pinning: tests/test_oracle.py checks pto_trace bit-for-bit against the reference's own sources
 * compiled by oracle/build_ref.py (oracle/_ref/libref_*.so) and against committed golden vectors in
 * tests/golden/ produced by the reference's own gen_data.py / data_visualization.py.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp -shared -fPIC   (see oracle/build.py).
 * The float arithmetic below relies on SSE2 binary32/binary64 ops rounding once each.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define PTO_EPS 1e-4f   /* src/common.h:9  */
#define PTO_MISS 1e20f  /* src/rt_helper.h:363 */

/* ------------------------------------------------------------------------------------------------
 * Radiance kernel.  Layouts: rays float32 SoA [6][n] = ox,oy,oz,dx,dy,dz;
 * spheres float32 SoA [10][stride] = r^2,x,y,z,ex,ey,ez,cr,cg,cb; colors float32 SoA [3][n].
 * The reference is nsph = stride = 8, depth = 5, light = 7, scale = 12.
 * first/count select a slice of paths (the reference's per-core slice, src/render.cpp:24-27).
 * ---------------------------------------------------------------------------------------------- */
static inline void pto_trace_one(const float *ox_, const float *oy_, const float *oz_, const float *dx_, const float *dy_,
                                 const float *dz_, const float *sph, int nsph, int stride, int depth, int light, float scale,
                                 float *out_r, float *out_g, float *out_b, uint32_t *segs) {
    const float *R2 = sph, *CX = sph + stride, *CY = sph + 2 * stride, *CZ = sph + 3 * stride;
    const float *KR = sph + 7 * stride, *KG = sph + 8 * stride, *KB = sph + 9 * stride;
    float ox = *ox_, oy = *oy_, oz = *oz_, dx = *dx_, dy = *dy_, dz = *dz_;
    float rr = 1.0f, rg = 1.0f, rb = 1.0f; /* render.cpp:113-118 */
    int alive = 1;                          /* render.cpp:120-121 */
    uint32_t live_segments = 0;
    for (int bounce = 0; bounce < depth; bounce++) { /* render.cpp:141 */
        if (alive && (rr != 0.0f || rg != 0.0f || rb != 0.0f))
            live_segments++; /* statistics only: segments an early-terminating tracer would need */
        float tmin = 0.0f;
        int idx = 0;
        for (int k = 0; k < nsph; k++) { /* rt_helper.h:457-474 */
            /* rt_helper.h:263-268: Adds(oc, o, -c); Muls(oc, oc, -1) */
            float ocx = (ox + (-CX[k])) * -1.0f;
            float ocy = (oy + (-CY[k])) * -1.0f;
            float ocz = (oz + (-CZ[k])) * -1.0f;
            /* rt_helper.h:273-280: b = 0; b += ocx*dx; b += ocy*dy; b += ocz*dz (mul then add) */
            float b = 0.0f, tmp;
            tmp = ocx * dx; b = b + tmp;
            tmp = ocy * dy; b = b + tmp;
            tmp = ocz * dz; b = b + tmp;
            /* rt_helper.h:297-304 */
            float c = 0.0f;
            tmp = ocx * ocx; c = c + tmp;
            tmp = ocy * ocy; c = c + tmp;
            tmp = ocz * ocz; c = c + tmp;
            c = c + (-R2[k]);
            /* rt_helper.h:314-315 */
            float disc = b * b;
            disc = disc - c;
            float s = sqrtf(disc); /* rt_helper.h:325: NaN for disc < 0 */
            float t0 = b - s, t1 = b + s; /* rt_helper.h:330-331 */
            float t = (t0 > PTO_EPS) ? t0 : t1;   /* FakeSelect rt_helper.h:207-213,346 */
            t = (t > PTO_EPS) ? t : PTO_MISS;      /* FakeCompare + Select rt_helper.h:357-364; NaN -> miss */
            /* rt_helper.h:402-443: block min, then lowest index whose value equals the min */
            if (k == 0 || t < tmin) {
                tmin = t;
                idx = k;
            }
        }
        /* rt_helper.h:513-518 */
        float hx = dx * tmin; hx = ox + hx;
        float hy = dy * tmin; hy = oy + hy;
        float hz = dz * tmin; hz = oz + hz;
        /* rt_helper.h:563-565,635-637 */
        float nx = hx - CX[idx], ny = hy - CY[idx], nz = hz - CZ[idx];
        /* rt_helper.h:641-658 */
        float len = 0.0f;
        {
            float q;
            q = nx * nx; len = len + q;
            q = ny * ny; len = len + q;
            q = nz * nz; len = len + q;
        }
        len = sqrtf(len);
        /* rt_helper.h:664-666 */
        float ux = nx / len, uy = ny / len, uz = nz / len;
        /* rt_helper.h:690-697 */
        float dot = 0.0f;
        {
            float q;
            q = dx * ux; dot = dot + q;
            q = dy * uy; dot = dot + q;
            q = dz * uz; dot = dot + q;
        }
        dot = dot * 2.0f;
        /* rt_helper.h:698-703 */
        float px = ux * dot, py = uy * dot, pz = uz * dot;
        dx = dx - px; dy = dy - py; dz = dz - pz;
        /* rt_helper.h:706-708 */
        ox = hx * 1.0f; oy = hy * 1.0f; oz = hz * 1.0f;
        /* rt_helper.h:773-810 */
        alive = alive && ((float)idx != (float)light);
        float cr = alive ? KR[idx] : 1.0f, cg = alive ? KG[idx] : 1.0f, cb = alive ? KB[idx] : 1.0f;
        rr = cr * rr; rg = cg * rg; rb = cb * rb;
    }
    *out_r = rr * scale; *out_g = rg * scale; *out_b = rb * scale; /* render.cpp:194-196 */
    if (segs)
        *segs = live_segments;
}

/* Returns the number of "live" segments (see above) summed over the slice; the reference itself always
 * traces count*depth segments. */
uint64_t pto_trace(const float *rays, const float *spheres, float *colors, int64_t n, int64_t first, int64_t count, int nsph,
                   int stride, int depth, int light, float scale) {
    uint64_t live = 0;
#pragma omp parallel for schedule(static) reduction(+ : live)
    for (int64_t i = first; i < first + count; i++) {
        uint32_t segs;
        pto_trace_one(rays + i, rays + n + i, rays + 2 * n + i, rays + 3 * n + i, rays + 4 * n + i, rays + 5 * n + i, spheres, nsph,
                      stride, depth, light, scale, colors + i, colors + n + i, colors + 2 * n + i, &segs);
        live += segs;
    }
    return live;
}

/* First-hit only (stage check): writes tmin and idx of bounce 0 for each path. */
void pto_first_hit(const float *rays, const float *spheres, float *tmin_out, int32_t *idx_out, int64_t n, int nsph, int stride, float eps) {
    const float *R2 = spheres, *CX = spheres + stride, *CY = spheres + 2 * stride, *CZ = spheres + 3 * stride;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        float ox = rays[i], oy = rays[n + i], oz = rays[2 * n + i], dx = rays[3 * n + i], dy = rays[4 * n + i], dz = rays[5 * n + i];
        float tmin = 0.0f;
        int idx = 0;
        for (int k = 0; k < nsph; k++) {
            float ocx = (ox + (-CX[k])) * -1.0f, ocy = (oy + (-CY[k])) * -1.0f, ocz = (oz + (-CZ[k])) * -1.0f;
            float b = 0.0f, tmp;
            tmp = ocx * dx; b = b + tmp;
            tmp = ocy * dy; b = b + tmp;
            tmp = ocz * dz; b = b + tmp;
            float c = 0.0f;
            tmp = ocx * ocx; c = c + tmp;
            tmp = ocy * ocy; c = c + tmp;
            tmp = ocz * ocz; c = c + tmp;
            c = c + (-R2[k]);
            float disc = b * b;
            disc = disc - c;
            float s = sqrtf(disc);
            float t0 = b - s, t1 = b + s;
            float t = (t0 > eps) ? t0 : t1;
            t = (t > eps) ? t : PTO_MISS;
            if (k == 0 || t < tmin) {
                tmin = t;
                idx = k;
            }
        }
        tmin_out[i] = tmin;
        idx_out[i] = idx;
    }
}

/* ------------------------------------------------------------------------------------------------
 * NumPy legacy RandomState: MT19937 seeded by init_genrand, doubles by genrand_res53.
 * np.random.seed(0) (gen_data.py:438) == init_genrand(0); rand() = ((a>>5)*2^26 + (b>>6)) / 2^53.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    uint32_t mt[624];
    int pos;
} pto_mt_t;

void pto_mt_seed(pto_mt_t *st, uint32_t seed) {
    st->mt[0] = seed;
    for (int i = 1; i < 624; i++)
        st->mt[i] = 1812433253u * (st->mt[i - 1] ^ (st->mt[i - 1] >> 30)) + (uint32_t)i;
    st->pos = 624;
}

static void pto_mt_refill(pto_mt_t *st) {
    uint32_t *mt = st->mt;
    for (int k = 0; k < 624; k++) {
        uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
        mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    st->pos = 0;
}

uint32_t pto_mt_next(pto_mt_t *st) {
    if (st->pos >= 624)
        pto_mt_refill(st);
    uint32_t y = st->mt[st->pos++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

double pto_mt_double(pto_mt_t *st) {
    uint32_t a = pto_mt_next(st) >> 5, b = pto_mt_next(st) >> 6;
    return ((double)a * 67108864.0 + (double)b) / 9007199254740992.0;
}

/* Fill `out` with the first n tempered words after seeding (ray i uses words 4i..4i+3). */
void pto_mt_words(uint32_t seed, uint32_t *out, int64_t n) {
    pto_mt_t st;
    pto_mt_seed(&st, seed);
    for (int64_t i = 0; i < n; i++)
        out[i] = pto_mt_next(&st);
}

void pto_mt_doubles(uint32_t seed, int64_t skip, double *out, int64_t n) {
    pto_mt_t st;
    pto_mt_seed(&st, seed);
    for (int64_t i = 0; i < skip; i++)
        (void)pto_mt_double(&st);
    for (int64_t i = 0; i < n; i++)
        out[i] = pto_mt_double(&st);
}


/* ------------------------------------------------------------------------------------------------
 * Counter-based generator of the production path: Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11).
 * NOT part of the reference ("parity unpinned" by it): pinned instead by the Random123 known-answer
 * vectors in tests/test_oracle.py.  Uniforms are formed from word pairs exactly like genrand_res53.
 * ---------------------------------------------------------------------------------------------- */
void pto_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* uniforms for global path indices [first, first+n): out[2i], out[2i+1] */
void pto_philox_uniforms(uint64_t seed, uint64_t first, int64_t n, double *out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        uint64_t g = first + (uint64_t)i;
        uint32_t ctr[4] = {(uint32_t)g, (uint32_t)(g >> 32), 0, 0}, key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, w[4];
        pto_philox4x32_10(ctr, key, w);
        out[2 * i] = ((double)(w[0] >> 5) * 67108864.0 + (double)(w[1] >> 6)) / 9007199254740992.0;
        out[2 * i + 1] = ((double)(w[2] >> 5) * 67108864.0 + (double)(w[3] >> 6)) / 9007199254740992.0;
    }
}

/* ------------------------------------------------------------------------------------------------
 * Camera (gen_data.py:24-29), all binary64.  cam[0..2] = position, [3..5] = direction, [6..8] = cx,
 * [9..11] = cy.  np.linalg.norm(v) = sqrt(v.dot(v)); the 3-term dot is accumulated left to right.
 * ---------------------------------------------------------------------------------------------- */
static double pto_norm3(const double *v) {
    double s = v[0] * v[0];
    s = s + v[1] * v[1];
    s = s + v[2] * v[2];
    return sqrt(s);
}

void pto_camera(int w, int h, double *cam) {
    const double pos[3] = {50, 52, 295.6};
    const double raw[3] = {0, -0.042612, -1};
    double nr = pto_norm3(raw);
    double dir[3] = {raw[0] / nr, raw[1] / nr, raw[2] / nr};
    double cx[3] = {(double)w * 0.5135 / (double)h, 0, 0};
    /* np.cross(cx, dir) */
    double cr[3] = {cx[1] * dir[2] - cx[2] * dir[1], cx[2] * dir[0] - cx[0] * dir[2], cx[0] * dir[1] - cx[1] * dir[0]};
    double ncr = pto_norm3(cr);
    double cy[3] = {cr[0] / ncr * 0.5135, cr[1] / ncr * 0.5135, cr[2] / ncr * 0.5135};
    memcpy(cam, pos, sizeof pos);
    memcpy(cam + 3, dir, sizeof dir);
    memcpy(cam + 6, cx, sizeof cx);
    memcpy(cam + 9, cy, sizeof cy);
}

/* One camera ray from its two uniform doubles (gen_data.py:37-47); writes 6 float32 (pos, dir). */
static inline void pto_ray_from_uniforms(const double *cam, int w, int h, int x, int y, int sx, int sy, double u1, double u2,
                                         float *out6) {
    double r1 = 2 * u1;
    double dx = r1 < 1 ? sqrt(r1) - 1 : 1 - sqrt(2 - r1);
    double r2 = 2 * u2;
    double dy = r2 < 1 ? sqrt(r2) - 1 : 1 - sqrt(2 - r2);
    double fx = ((sx + 0.5 + dx) / 2 + x) / w - 0.5;
    double fy = ((sy + 0.5 + dy) / 2 + y) / h - 0.5;
    double d[3];
    for (int c = 0; c < 3; c++) {
        double a = cam[6 + c] * fx;
        double b = cam[9 + c] * fy;
        d[c] = (a + b) + cam[3 + c];
    }
    double nrm = pto_norm3(d);
    for (int c = 0; c < 3; c++) {
        out6[c] = (float)(cam[c] + d[c] * 140);
        out6[3 + c] = (float)(d[c] / nrm);
    }
}

/* Rays of image columns [x0, x1) of a w x h image, s samples per sub-pixel, from uniforms `u`
 * (2 doubles per ray, ray order x -> y -> sy -> sx -> k, gen_data.py:32-36; u[0] belongs to the
 * first ray of column x0).  Output: float32 SoA [6][m], m = (x1-x0)*h*4*s. */
void pto_gen_rays_from_uniforms(int w, int h, int s, int x0, int x1, const double *u, float *rays) {
    double cam[12];
    pto_camera(w, h, cam);
    int64_t m = (int64_t)(x1 - x0) * h * 4 * s;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < m; i++) {
        int64_t r = i;
        int k = (int)(r % s); r /= s; (void)k;
        int sx = (int)(r % 2); r /= 2;
        int sy = (int)(r % 2); r /= 2;
        int y = (int)(r % h); r /= h;
        int x = x0 + (int)r;
        float o6[6];
        pto_ray_from_uniforms(cam, w, h, x, y, sx, sy, u[2 * i], u[2 * i + 1], o6);
        for (int c = 0; c < 6; c++)
            rays[(int64_t)c * m + i] = o6[c];
    }
}

/* gen_data.py main: np.random.seed(seed); gen_rays(w, h, s) for the whole image. */
void pto_gen_rays(int w, int h, int s, uint32_t seed, float *rays) {
    int64_t n = (int64_t)w * h * 4 * s;
    double *u = (double *)malloc(sizeof(double) * 2 * (size_t)n);
    pto_mt_doubles(seed, 0, u, 2 * n);
    pto_gen_rays_from_uniforms(w, h, s, 0, w, u, rays);
    free(u);
}

/* gen_data.py:92-132: 8 spheres, float64 table -> radius squared -> transpose -> pad to 512 B. */
void pto_gen_spheres(float *out128) {
    static const double tbl[8][10] = {
        {1e5, 1e5 + 1, 40.8, 81.6, 0, 0, 0, 0.435, 0.376, 0.667},  {1e5, -1e5 + 99, 40.8, 81.6, 0, 0, 0, 0.667, 0.129, 0.086},
        {1e5, 50, 40.8, 1e5, 0, 0, 0, 0.270, 0.725, 0.486},        {1e5, 50, 40.8, -1e5 + 170, 0, 0, 0, 0, 0, 0},
        {1e5, 50, 1e5, 81.6, 0, 0, 0, 0.5, 0.5, 0.5},              {1e5, 50, -1e5 + 81.6, 81.6, 0, 0, 0, 0.141, 0.408, 0.635},
        {16.5, 27, 16.5, 47, 0, 0, 0, 0.999, 0.999, 0.999},        {600, 50, 681.6 - 0.27, 81.6, 12, 12, 12, 0, 0, 0}};
    memset(out128, 0, 128 * sizeof(float));
    for (int i = 0; i < 8; i++)
        for (int m = 0; m < 10; m++) {
            double v = tbl[i][m];
            if (m == 0)
                v = v * v;
            out128[m * 8 + i] = (float)v;
        }
}

/* ------------------------------------------------------------------------------------------------
 * Resolve (data_visualization.py:20-59).  colors: float32 SoA [3][w][h][4s].
 * np.mean over a float32 slice of length s along its contiguous axis = NumPy pairwise float32 sum
 * (numpy/_core/src/umath/loops_utils.h.src, pairwise sum: <8 sequential from 0; <=128: 8 accumulators
 * + tail; else split at n/2 rounded down to a multiple of 8) followed by one float32 divide by s.
 * The four sub-pixel means are summed in binary64, /4, clip to [0,1], *255, truncate to uint8.
 * Output `img`: uint8 [h rows, top row first][w][3]; file row r = image y = h-1-r (SURVEY.md section 5:
 * identical to the reference writer whenever w == h, the only case the reference's writer handles).
 * ---------------------------------------------------------------------------------------------- */
static float pto_pairwise_sum_f32(const float *a, int64_t n) {
    if (n < 8) {
        float res = 0.0f; /* numpy starts from 0. (-0 -> +0 is immaterial here: values are >= 0) */
        for (int64_t i = 0; i < n; i++)
            res += a[i];
        return res;
    } else if (n <= 128) {
        float r[8];
        for (int j = 0; j < 8; j++)
            r[j] = a[j];
        int64_t i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++)
                r[j] += a[i + j];
        float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++)
            res += a[i];
        return res;
    } else {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        return pto_pairwise_sum_f32(a, n2) + pto_pairwise_sum_f32(a + n2, n - n2);
    }
}

/* np.mean of a contiguous float32 run (exported so tests can pin the summation order on floats). */
float pto_mean_f32(const float *a, int64_t n) { return n == 1 ? a[0] : pto_pairwise_sum_f32(a, n) / (float)n; }

void pto_resolve(const float *colors, int w, int h, int s, uint8_t *img) {
    int64_t n = (int64_t)w * h * 4 * s;
#pragma omp parallel for schedule(static)
    for (int r = 0; r < h; r++) {
        int u = h - 1 - r; /* data_visualization.py:40 with j == r */
        for (int i = 0; i < w; i++) {
            for (int c = 0; c < 3; c++) {
                const float *px = colors + (int64_t)c * n + ((int64_t)i * h + u) * 4 * s;
                double sum = 0.0;
                for (int k = 0; k < 4; k++) {
                    float m;
                    if (s == 1)
                        m = px[k]; /* mean of one element */
                    else
                        m = pto_pairwise_sum_f32(px + (int64_t)k * s, s) / (float)s;
                    sum += (double)m;
                }
                double v = sum / 4;
                v = v < 0 ? 0 : (v > 1 ? 1 : v);
                v = v * 255;
                img[((int64_t)r * w + i) * 3 + c] = (uint8_t)v;
            }
        }
    }
}

/* data_visualization.py:11-17 (P3).  One text line per image row, "R G B " per pixel. */
int pto_write_ppm(const char *path, int w, int h, const uint8_t *img) {
    FILE *f = fopen(path, "w");
    if (!f)
        return -1;
    fprintf(f, "P3\n%d %d\n255\n", w, h);
    for (int r = 0; r < h; r++) {
        for (int i = 0; i < w; i++) {
            const uint8_t *p = img + ((int64_t)r * w + i) * 3;
            fprintf(f, "%d %d %d ", p[0], p[1], p[2]);
        }
        fprintf(f, "\n");
    }
    return fclose(f);
}
