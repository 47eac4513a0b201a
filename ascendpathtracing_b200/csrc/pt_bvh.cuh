// Sphere BVH for scenes beyond the constant bank (BASELINE config C4: 10 k spheres).
//
// Requirement: the nearest hit found through the BVH equals the brute-force loop's (same t bits, same index, lowest
// index on ties), for every ray whose origin lies in the scene.  The brute-force test is decided by rounding noise near
// grazing incidence (disc = b^2 - c carries an absolute error of a few ulp(|oc|^2)), so a sphere can be "hit" by a ray that
// geometrically misses it by up to sqrt(r^2 + E) - r.  Hence
//   * huge spheres (r^2 >= kBigR2: the 1e5-radius walls, the 600-radius light) never enter the tree: they stay in a
//     brute-force list in the constant bank, tested pairwise with packed f32x2 like the 8-sphere scene;
//   * every other sphere's box is its geometric box grown by pad = sqrt(r^2 + E) - r + 1e-3 r + 1e-3, E = 2^-19 * D^2 with
//     D the diagonal of the scene's bounds (a 16x margin over the worst-case discriminant error at that distance);
//   * leaves run the exact scalar test (sphere_t), candidates are merged with the (t, index) order of the reference.
//
// Layout: LBVH (Morton order, Karras 2012).  An internal node stores BOTH child boxes and both child references in
// 64 bytes (4 x LDG.128 served from L1/L2), so one fetch decides two subtrees; a reference < 0 is a leaf (~ref = sphere).
// Traversal: per-lane stack of 32 entries in local memory, near child first.
#pragma once
#include "pt_device.cuh"

namespace ptb200 {

constexpr float kBigR2 = 1.0e4f;  // radius >= 100 stays out of the tree
constexpr int kBvhStack = 48;

struct BvhNode {  // 64 bytes
    float4 a;     // l.min.x l.min.y l.min.z l.max.x
    float4 b;     // l.max.y l.max.z r.min.x r.min.y
    float4 c;     // r.min.z r.max.x r.max.y r.max.z
    int left, right;
    int parent, pad;
};

struct BvhScene {
    const BvhNode *nodes;    // n_small - 1 internal nodes (none when n_small <= 1)
    const float4 *geom;      // per sphere (original index): x, y, z, -r^2
    const float4 *color;     // r, g, b, material
    const float4 *emission;  // r, g, b, -
    const int *big_index;    // original indices of the spheres kept in the constant bank (ascending)
    int n_big, n_small, root, only_leaf;  // root: internal node 0 or, when n_small == 1, the leaf reference only_leaf
};

// Slab test against a padded box; NaN-tolerant min/max order (a NaN from 0 * inf falls out).
__device__ __forceinline__ bool hit_box(float ox, float oy, float oz, float ix, float iy, float iz, float lx, float ly, float lz, float hx,
                                        float hy, float hz, float tbest, float &tnear) {
    const float tx1 = (lx - ox) * ix, tx2 = (hx - ox) * ix;
    const float ty1 = (ly - oy) * iy, ty2 = (hy - oy) * iy;
    const float tz1 = (lz - oz) * iz, tz2 = (hz - oz) * iz;
    float tn = fminf(tx1, tx2), tf = fmaxf(tx1, tx2);
    tn = fmaxf(tn, fminf(ty1, ty2)), tf = fminf(tf, fmaxf(ty1, ty2));
    tn = fmaxf(tn, fminf(tz1, tz2)), tf = fminf(tf, fmaxf(tz1, tz2));
    tnear = tn;
    return tf >= fmaxf(tn, 0.0f) && tn <= tbest;
}

__device__ __forceinline__ void bvh_leaf(const BvhScene &sc, int sphere, float ox, float oy, float oz, float dx, float dy, float dz, float eps,
                                         float &tmin, int &idx) {
    const float4 g = __ldg(sc.geom + sphere);
    const float t = sphere_t(ox, oy, oz, dx, dy, dz, g.x, g.y, g.z, g.w, eps);
    if (t < tmin || (t == tmin && t < kMiss && sphere < idx)) {  // the reference's order: smaller t, then lower index
        tmin = t;
        idx = sphere;
    }
}

// Refines (tmin, idx) -- already holding the best of the brute-force list, or (1e20, 0) -- with the tree's spheres.
__device__ __forceinline__ void bvh_nearest(const BvhScene &sc, float ox, float oy, float oz, float dx, float dy, float dz, float eps, float &tmin,
                                            int &idx) {
    if (sc.n_small <= 0)
        return;
    if (sc.n_small == 1) {
        bvh_leaf(sc, ~sc.only_leaf, ox, oy, oz, dx, dy, dz, eps, tmin, idx);
        return;
    }
    const float ix = 1.0f / dx, iy = 1.0f / dy, iz = 1.0f / dz;
    int stack[kBvhStack];
    int sp = 0;
    int node = 0;
    for (;;) {
        const BvhNode *nd = sc.nodes + node;
        const float4 a = __ldg(&nd->a), b = __ldg(&nd->b), c = __ldg(&nd->c);
        const int left = __ldg(&nd->left), right = __ldg(&nd->right);
        float tl, tr;
        bool hl = hit_box(ox, oy, oz, ix, iy, iz, a.x, a.y, a.z, a.w, b.x, b.y, tmin, tl);
        bool hr = hit_box(ox, oy, oz, ix, iy, iz, b.z, b.w, c.x, c.y, c.z, c.w, tmin, tr);
        if (hl && left < 0) {
            bvh_leaf(sc, ~left, ox, oy, oz, dx, dy, dz, eps, tmin, idx);
            hl = false;
        }
        if (hr && right < 0) {
            bvh_leaf(sc, ~right, ox, oy, oz, dx, dy, dz, eps, tmin, idx);
            hr = false;
        }
        if (hl && hr) {  // near child first, far child on the stack
            const bool lfirst = tl <= tr;
            if (sp < kBvhStack)
                stack[sp++] = lfirst ? right : left;
            node = lfirst ? left : right;
        } else if (hl) {
            node = left;
        } else if (hr) {
            node = right;
        } else {
            if (sp == 0)
                break;
            node = stack[--sp];
        }
    }
}

}  // namespace ptb200
