// Sphere BVH for scenes beyond the constant bank (BASELINE config C4: 10 k spheres).
//
// Requirement: the nearest hit found through the BVH equals the brute-force loop's (same t bits, same index, lowest
// index on ties), for every ray.  The brute-force test is decided by rounding noise near grazing incidence
// (disc = b^2 - c carries an absolute error of a few ulp(|oc|^2)), so a sphere can be "hit" by a ray that
// geometrically misses it by up to sqrt(r^2 + E) - r.  Hence
//   * huge spheres (r^2 >= kBigR2: the 1e5-radius walls, the 600-radius light) never enter the tree: they stay in a
//     brute-force list in the constant bank, tested pairwise with packed f32x2 like the 8-sphere scene;
//   * a sphere's box is its geometric box plus a little slack (1e-3 r + 1e-3) for the slab arithmetic; the margin that
//     exactness needs is applied PER RAY, because it depends on the ray: the reference's discriminant carries an absolute
//     error proportional to |centre - origin|^2, plus |d|^2 - 1 times that when the direction is not exactly unit (mirror
//     bounces are not renormalised), so a ray can "hit" a sphere its line misses by up to m = sqrt(r^2 + E) - r with
//     E = (2^-19 + ||d|^2 - 1|) L^2, L = the farthest a sphere that can still beat the ray's current best hit can be.
//     m is largest for the smallest sphere of the tree; the walk grows every box by that m, folded into the near / far
//     constants of the slab FFMA (no instruction per step).  A fixed, scene-wide pad (the first version: 0.29 units in C4)
//     swamps small spheres; the per-ray margin is 0.04 at 100 units and 0.004 at 30;
//   * leaves run the exact scalar test, candidates are merged with the (t, index) order of the reference.
//
// What bounds traversal on B200 (ncu, profiles/r2_c4_*): incoherent rays make every node fetch touch one cache line per
// lane, and L1 looks up one line per clock per SM -- the tag stage, not the arithmetic, is the limit.  So a node is
// 32 bytes = ONE LDG.256 (per child: 6 x 16-bit box planes + reference), a quarter of the float layout's lookups:
//   * planes are quantised to a 16-bit grid over the tree's bounds, rounded outwards and grown by kGridGrow units;
//   * decoding costs nothing: PRMT drops the 16 bits into the mantissa of 2^23 (0x4B000000 | q = 2^23 + q exactly), and
//     the slab distance is ONE FFMA, t = (2^23 + q) * id + c with id = 1 / (d * scale) and c = -(2^23 + round(og)) * id
//     (og = ray origin in grid units).  Rounding og to an integer and rounding c each move a plane by at most 0.504 grid
//     units, the FMA's own rounding by less than 0.13 (origins within 2^21 units) and folding the per-ray margin into c
//     by another 0.5, together < kGridGrow = 3 minus the slack spent on the float evaluation of floor/ceil at build
//     time: the test stays conservative.  A relative error of
//     id scales both terms alike, so MUFU.RCP is good enough for it;
//   * the near/far plane of each axis is picked by the PRMT selector (sign of d), not by min/max afterwards.
// Origins further than 2^21 grid units from the tree (64 scene widths), directions more than 1 % off unit length and NaNs
// take an exact loop over all spheres (run by the whole warp for one ray in the wavefront kernel).
//
// Layout: LBVH (Morton order, Karras 2012); a reference < 0 is a leaf (~ref = sphere).  Traversal: near child first,
// leaf tests and stack policy supplied by the caller.
#pragma once
#include "pt_device.cuh"

namespace ptb200 {

constexpr float kBigR2 = 1.0e4f;    // radius >= 100 stays out of the tree
constexpr int kBvhStack = 64;       // Karras depth <= 30 key bits + 24 index bits (ptb200_bvh_build caps the count at 2^24)
constexpr int kGridGrow = 3;        // grid units added on every side of a quantised box (see above)
constexpr float kGridMax = 65535.0f;
constexpr float kGridFar = 2097152.0f;  // 2^21: |og| beyond this -> exact fallback

struct BvhNode {  // build-time node, float boxes (bvh.cu), 64 bytes
    float4 a;     // l.min.x l.min.y l.min.z l.max.x
    float4 b;     // l.max.y l.max.z r.min.x r.min.y
    float4 c;     // r.min.z r.max.x r.max.y r.max.z
    int left, right;
    int parent, pad;
};

struct QNode {  // traversal node, 32 bytes: word c of a child = lo_c | hi_c << 16 in grid units
    unsigned int lx, ly, lz;
    int left;
    unsigned int rx, ry, rz;
    int right;
};

struct BvhScene {
    const QNode *qnodes;     // n_small - 1 internal nodes (none when n_small <= 1)
    const float4 *geom;      // per sphere (original index): x, y, z, -r^2
    const float4 *color;     // r, g, b, material
    const float4 *emission;  // r, g, b, -
    const int *big_index;    // original indices of the spheres kept in the constant bank (ascending)
    const int *small_index;  // original indices of the spheres in the tree (ascending): the exact fallback's list
    int n_big, n_small, root, only_leaf;  // root: internal node 0 or, when n_small == 1, the leaf reference only_leaf
    float glo[3], gscale[3];              // grid = (world - glo) * gscale
    float centre[3], radius;              // bounding ball of the tree's boxes (world units)
    float rmin, rmax;                     // smallest / largest sphere radius in the tree
};

// Per-ray traversal constants (12 registers).
struct BvhRay {
    float idx, idy, idz;          // 1 / (d * gscale)
    float nx, ny, nz;             // near planes: -(2^23 + round(og)) * id - margin * |id|
    float fx, fy, fz;             // far planes:  -(2^23 + round(og)) * id + margin * |id|
    unsigned int sx, sy, sz;      // PRMT selectors of the near plane per axis (far = near ^ 0x22)
    bool far_origin;              // not coverable by the quantised walk (far / NaN origin, direction far from unit): exact loop
};

// tbest: the best hit so far (the brute-force list's, or 1e20): bounds how far a sphere that still matters can be.
__device__ __forceinline__ BvhRay bvh_ray(const BvhScene &sc, float ox, float oy, float oz, float dx, float dy, float dz, float tbest) {
    BvhRay r;
    const float ogx = __fmul_rn(__fsub_rn(ox, sc.glo[0]), sc.gscale[0]);
    const float ogy = __fmul_rn(__fsub_rn(oy, sc.glo[1]), sc.gscale[1]);
    const float ogz = __fmul_rn(__fsub_rn(oz, sc.glo[2]), sc.gscale[2]);
    // Exactness margin (header).  The computed discriminant of a sphere at distance a = |centre - origin| differs from
    // r^2 - (distance of the centre from the ray's line)^2 by at most 16 u a^2 + 2 u r^2 of rounding (u = 2^-24: three
    // subtractions, two 3-term dot products, one square, two sums) plus | |d|^2 - 1 | a^2 of direction length; twice the
    // rounding bound is used.  A sphere that can still replace the best hit t has a <= t |d| + r + m, and none is further
    // than `reach`.
    const float ex = ox - sc.centre[0], ey = oy - sc.centre[1], ez = oz - sc.centre[2];
    const float reach = sqrtf(ex * ex + ey * ey + ez * ez) + sc.radius;
    const float eta = fabsf((dx * dx + dy * dy + dz * dz) - 1.0f) + 4.0e-7f;  // | |d|^2 - 1 |, with room for its own rounding
    const float k2 = 1.9073486e-6f + eta;                                       // 2^-19 + eta
    const float by_hit = (tbest * (1.001f + 0.5f * eta) + sc.rmax + 0.01f) / (1.0f - sqrtf(k2)) * 1.001f;
    const float L = fminf(by_hit, reach);
    const float E = k2 * L * L + 2.3841858e-7f * sc.rmax * sc.rmax;             // ... + 2^-22 rmax^2
    const float margin = 1.001f * (E / (sqrtf(sc.rmin * sc.rmin + E) + sc.rmin)) + 1e-6f;  // sqrt(rmin^2 + E) - rmin, rounded up
    r.far_origin = !(fmaxf(fmaxf(fabsf(ogx), fabsf(ogy)), fabsf(ogz)) <= kGridFar) || !(eta <= 0.01f) || !(margin <= 64.0f);
    r.idx = mufu_rcp(__fmul_rn(dx, sc.gscale[0]));
    r.idy = mufu_rcp(__fmul_rn(dy, sc.gscale[1]));
    r.idz = mufu_rcp(__fmul_rn(dz, sc.gscale[2]));
    const float cx = -__fmul_rn(__fadd_rn(ogx, 8388608.0f), r.idx);
    const float cy = -__fmul_rn(__fadd_rn(ogy, 8388608.0f), r.idy);
    const float cz = -__fmul_rn(__fadd_rn(ogz, 8388608.0f), r.idz);
    const float mx = __fmul_rn(__fmul_rn(margin, sc.gscale[0]), fabsf(r.idx));  // margin in grid units times |id|
    const float my = __fmul_rn(__fmul_rn(margin, sc.gscale[1]), fabsf(r.idy));
    const float mz = __fmul_rn(__fmul_rn(margin, sc.gscale[2]), fabsf(r.idz));
    r.nx = __fsub_rn(cx, mx), r.ny = __fsub_rn(cy, my), r.nz = __fsub_rn(cz, mz);
    r.fx = __fadd_rn(cx, mx), r.fy = __fadd_rn(cy, my), r.fz = __fadd_rn(cz, mz);
    r.sx = dx < 0.0f ? 0x7632u : 0x7610u;
    r.sy = dy < 0.0f ? 0x7632u : 0x7610u;
    r.sz = dz < 0.0f ? 0x7632u : 0x7610u;
    return r;
}

// Slab test of one quantised child box.  NaN planes (0 * inf, inf - inf on an axis the ray does not move along) drop out
// of the 3-input min/max, which only makes the test more permissive.
__device__ __forceinline__ bool hit_qbox(const BvhRay &r, unsigned int wx, unsigned int wy, unsigned int wz, float tbest, float &tnear) {
#ifdef PTB_BVH_SCALAR_SLABS
    const float nx = __fmaf_rn(__uint_as_float(__byte_perm(wx, 0x4B000000u, r.sx)), r.idx, r.nx);
    const float ny = __fmaf_rn(__uint_as_float(__byte_perm(wy, 0x4B000000u, r.sy)), r.idy, r.ny);
    const float nz = __fmaf_rn(__uint_as_float(__byte_perm(wz, 0x4B000000u, r.sz)), r.idz, r.nz);
    const float fx = __fmaf_rn(__uint_as_float(__byte_perm(wx, 0x4B000000u, r.sx ^ 0x22u)), r.idx, r.fx);
    const float fy = __fmaf_rn(__uint_as_float(__byte_perm(wy, 0x4B000000u, r.sy ^ 0x22u)), r.idy, r.fy);
    const float fz = __fmaf_rn(__uint_as_float(__byte_perm(wz, 0x4B000000u, r.sz ^ 0x22u)), r.idz, r.fz);
#else
    // near and far plane of an axis as one packed FFMA2 (same two roundings as two FFMAs, half the issue slots)
    const float2 px = __ffma2_rn(make_float2(__uint_as_float(__byte_perm(wx, 0x4B000000u, r.sx)), __uint_as_float(__byte_perm(wx, 0x4B000000u, r.sx ^ 0x22u))),
                                 make_float2(r.idx, r.idx), make_float2(r.nx, r.fx));
    const float2 py = __ffma2_rn(make_float2(__uint_as_float(__byte_perm(wy, 0x4B000000u, r.sy)), __uint_as_float(__byte_perm(wy, 0x4B000000u, r.sy ^ 0x22u))),
                                 make_float2(r.idy, r.idy), make_float2(r.ny, r.fy));
    const float2 pz = __ffma2_rn(make_float2(__uint_as_float(__byte_perm(wz, 0x4B000000u, r.sz)), __uint_as_float(__byte_perm(wz, 0x4B000000u, r.sz ^ 0x22u))),
                                 make_float2(r.idz, r.idz), make_float2(r.nz, r.fz));
    const float nx = px.x, ny = py.x, nz = pz.x, fx = px.y, fy = py.y, fz = pz.y;
#endif
    const float tn = fmaxf(fmaxf(nx, ny), nz);
    const float tf = fminf(fminf(fx, fy), fz);
    tnear = tn;
    return tf >= fmaxf(tn, 0.0f) && tn <= tbest;
}

// The whole 32-byte node in ONE 256-bit load (LDG.E.256, new on sm_100): one L1 tag lookup per lane per step.
__device__ __forceinline__ void ldg_node(const QNode *n, uint4 &L, uint4 &R) {
    asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(L.x), "=r"(L.y), "=r"(L.z), "=r"(L.w), "=r"(R.x), "=r"(R.y), "=r"(R.z), "=r"(R.w)
        : "l"(n));
}

// Exact leaf test (sphere_t's arithmetic, rt_helper.h:255-370) with the reference's (t, index) merge: smaller t, then
// lower index.  A negative or NaN discriminant is a miss (sqrt gives NaN, both compares fail, t = 1e20) and a miss can
// never replace the current best, so the square root is skipped for it.
__device__ __forceinline__ void bvh_leaf(const BvhScene &sc, int sphere, float ox, float oy, float oz, float dx, float dy, float dz, float eps,
                                         float &tmin, int &idx) {
    PTB_CHECK(sphere >= 0 && sphere < sc.n_big + sc.n_small);
    const float4 g = __ldg(sc.geom + sphere);
    const float ocx = __fsub_rn(g.x, ox);
    const float ocy = __fsub_rn(g.y, oy);
    const float ocz = __fsub_rn(g.z, oz);
    const float b = __fadd_rn(__fadd_rn(__fmul_rn(ocx, dx), __fmul_rn(ocy, dy)), __fmul_rn(ocz, dz));
    const float c = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(ocx, ocx), __fmul_rn(ocy, ocy)), __fmul_rn(ocz, ocz)), g.w);
    const float disc = __fsub_rn(__fmul_rn(b, b), c);
    if (disc >= 0.0f) {
        const float s = __fsqrt_rn(disc);
        const float t0 = __fsub_rn(b, s);
        const float t1 = __fadd_rn(b, s);
        float t = (t0 > eps) ? t0 : t1;
        t = (t > eps) ? t : kMiss;
        if (t < tmin || (t == tmin && t < kMiss && sphere < idx)) {
            tmin = t;
            idx = sphere;
        }
    }
}

// One traversal step at internal node `node`: tests both child boxes, reports hit leaves in leaf_a / leaf_b (sphere index
// or -1; leaf_b is only set when leaf_a is) and leaves the next internal node in `node` (-1 when the stack ran empty:
// traversal finished).  The caller runs the exact test on the leaves -- at once, or batched across the warp.
// Stack: int step(bool push, bool pop, int x, int otherwise) -- a traversal step pushes x, pops (-1 when empty), or neither.
template <class Stack>
__device__ __forceinline__ void bvh_step_boxes(const BvhScene &sc, const BvhRay &r, float tmin, int &node, int &leaf_a, int &leaf_b, Stack &st) {
    uint4 L, R;
    PTB_CHECK(node >= 0 && node < sc.n_small - 1);
    ldg_node(sc.qnodes + node, L, R);
    const int left = static_cast<int>(L.w), right = static_cast<int>(R.w);
    PTB_CHECK((left >= 0 ? left : ~left) < sc.n_big + sc.n_small && (right >= 0 ? right : ~right) < sc.n_big + sc.n_small);
#ifdef PTB_BVH_PREFETCH  // both children towards L1 while this node's boxes are tested
    if (left >= 0)
        asm volatile("prefetch.global.L1 [%0];" ::"l"(sc.qnodes + left));
    if (right >= 0)
        asm volatile("prefetch.global.L1 [%0];" ::"l"(sc.qnodes + right));
#endif
    float tl, tr;
    const bool bl = hit_qbox(r, L.x, L.y, L.z, tmin, tl);
    const bool br = hit_qbox(r, R.x, R.y, R.z, tmin, tr);
    const bool ll = bl && left < 0, lr = br && right < 0;
    leaf_a = ll ? ~left : (lr ? ~right : -1);
    leaf_b = (ll && lr) ? ~right : -1;
    const bool hl = bl && left >= 0, hr = br && right >= 0;
    // near child first, far child on the stack; nothing hit: pop
    const bool lfirst = tl <= tr, both = hl && hr;
    const int next = (both ? lfirst : hl) ? left : right;
    node = st.step(both, !(hl || hr), lfirst ? right : left, next);
}

// Step with the leaves tested at once (the plain per-lane walk).
template <class Stack, class Ray>
__device__ __forceinline__ void bvh_step(const BvhScene &sc, const BvhRay &r, const Ray &ray, float eps, int &node, float &tmin, int &idx, Stack &st) {
    int leaf_a, leaf_b;
    bvh_step_boxes(sc, r, tmin, node, leaf_a, leaf_b, st);
    if (leaf_a >= 0)
        bvh_leaf(sc, leaf_a, ray.ox(), ray.oy(), ray.oz(), ray.dx(), ray.dy(), ray.dz(), eps, tmin, idx);
    if (leaf_b >= 0)
        bvh_leaf(sc, leaf_b, ray.ox(), ray.oy(), ray.oz(), ray.dx(), ray.dy(), ray.dz(), eps, tmin, idx);
}

// Exact fallback for rays the quantised traversal does not cover: every sphere of the tree, ascending index.
static __device__ __noinline__ void bvh_all_leaves(const BvhScene &sc, float ox, float oy, float oz, float dx, float dy, float dz, float eps,
                                                   float &tmin, int &idx) {
    for (int k = 0; k < sc.n_small; k++)
        bvh_leaf(sc, __ldg(sc.small_index + k), ox, oy, oz, dx, dy, dz, eps, tmin, idx);
}

// The same, by a whole warp for ONE ray (all 32 lanes must call it with that ray's values): lanes test interleaved spheres
// with coalesced loads and the best (t, index) is reduced in the reference's order.  A leaked path that bounces off the
// outside of a 1e5-radius wall ends up 10^5 units away, where the binary32 test reports hits all over the scene, so nothing
// can be culled for it; alone, one lane took 0.15 s for such a ray in a 10^6-sphere scene (dependent loads), the warp 1 ms.
__device__ __forceinline__ void bvh_all_leaves_warp(const BvhScene &sc, unsigned int lane, float ox, float oy, float oz, float dx, float dy,
                                                    float dz, float eps, float &tmin, int &idx) {
    const float t_in = tmin;
    const int i_in = idx;
    for (int k = static_cast<int>(lane); k < sc.n_small; k += 32)
        bvh_leaf(sc, __ldg(sc.small_index + k), ox, oy, oz, dx, dy, dz, eps, tmin, idx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float t2 = __shfl_xor_sync(0xffffffffu, tmin, o);
        const int i2 = __shfl_xor_sync(0xffffffffu, idx, o);
        if (t2 < tmin || (t2 == tmin && i2 < idx)) {
            tmin = t2;
            idx = i2;
        }
    }
    if (!(tmin < t_in) && !(tmin == t_in && tmin < kMiss && idx < i_in)) {  // nothing beat what the ray came in with
        tmin = t_in;
        idx = i_in;
    }
}

struct LocalStack {  // per-lane stack in local memory
    int *v;
    int sp;
    __device__ __forceinline__ void push_if(bool c, int x) {
        if (c)
            v[sp++] = x;
    }
    __device__ __forceinline__ int pop_if(bool c, int fallback) {
        if (!c)
            return fallback;
        return sp == 0 ? -1 : v[--sp];
    }
    __device__ __forceinline__ int step(bool push, bool pop, int x, int otherwise) {  // a step pushes, pops, or neither
        push_if(push, x);
        return pop_if(pop, otherwise);
    }
};

struct RegRay {
    float o[3], d[3];
    __device__ __forceinline__ float ox() const { return o[0]; }
    __device__ __forceinline__ float oy() const { return o[1]; }
    __device__ __forceinline__ float oz() const { return o[2]; }
    __device__ __forceinline__ float dx() const { return d[0]; }
    __device__ __forceinline__ float dy() const { return d[1]; }
    __device__ __forceinline__ float dz() const { return d[2]; }
};

// Refines (tmin, idx) -- already holding the best of the brute-force list, or (1e20, 0) -- with the tree's spheres.
// Plain per-lane loop: used by the first-hit diagnostic kernel and by the lock-step material kernel of small launches.
__device__ __forceinline__ void bvh_nearest(const BvhScene &sc, float ox, float oy, float oz, float dx, float dy, float dz, float eps, float &tmin,
                                            int &idx) {
    if (sc.n_small <= 0)
        return;
    if (sc.n_small == 1) {
        bvh_leaf(sc, ~sc.only_leaf, ox, oy, oz, dx, dy, dz, eps, tmin, idx);
        return;
    }
    const BvhRay r = bvh_ray(sc, ox, oy, oz, dx, dy, dz, tmin);
    if (r.far_origin) {
        bvh_all_leaves(sc, ox, oy, oz, dx, dy, dz, eps, tmin, idx);
        return;
    }
    RegRay ray = {{ox, oy, oz}, {dx, dy, dz}};
    int stack_mem[kBvhStack];
    LocalStack st = {stack_mem, 0};
    int node = 0;
    while (node >= 0)
        bvh_step(sc, r, ray, eps, node, tmin, idx, st);
}

}  // namespace ptb200
