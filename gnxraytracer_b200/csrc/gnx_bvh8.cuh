// gnx_bvh8.cuh — compressed 8-wide BVH for the ORDER-INDEPENDENT queries (any-hit: shadow rays and the environment's
// MIS rays, VisibilityTester::Unoccluded / Scene::IntersectP, core/Light.cpp:14-31, accelerator/BVHAccel.cpp:689-729).
//
// A closest-hit query must visit the leaves in the reference's order to return the reference's hit, so it keeps the
// two-child tree of gnx_bvh.cuh.  An any-hit query only asks "is there a triangle with 0 < t < tMax": any tree over the
// same ordered triangles, visited in any order, gives the same answer as long as its box tests are conservative — the
// triangle test (Triangle::IntersectP's arithmetic, gnx_bvh.cuh) decides, exactly as in the reference.  That freedom is
// spent on a shorter dependent-load chain: eight children per node, child boxes quantised to 8 bits on a per-node
// power-of-two grid (Ylitie, Karras, Laine 2017), 96 bytes per node instead of 8 x 32, one stack entry per NODE
// (node, mask of the children still to visit).
//
// Node8 = six 16-byte words (read with LDG.128):
//   w0  origin.x origin.y origin.z (float bits)   ex | ey << 8 | ez << 16 | valid << 24
//       (e* = biased IEEE exponent of the grid step 2^(e - 127) on that axis; valid = mask of the occupied slots)
//   w1  child references 0..3      w2  child references 4..7
//       (>= 0: Node8 index; < 0: leaf, ~ref = primitive offset | (count - 1) << 27, the encoding of gnx_bvh.cuh)
//   w3  qlo.x[0..7] qlo.y[0..7]    w4  qlo.z[0..7] qhi.x[0..7]    w5  qhi.y[0..7] qhi.z[0..7]   (one byte per child)
// Child k's box is [origin + qlo * step, origin + qhi * step], a superset of the box the two-child tree holds for it.
// Children are stored by decreasing surface area (the likeliest occluder first).
#pragma once
#include "gnx_bvh.cuh"

namespace gnx {

constexpr int kNode8Words = 6;  // uint4 words per node

GNX_D int ffs32(int v) {
#ifdef __CUDA_ARCH__
    return __ffs(v);
#else
    return __builtin_ffs(v);
#endif
}

// byte k (0..7) of the pair (v.x, v.y) as a float, exactly: 0x4B0000bb is 2^23 + bb
GNX_D float node8_byte(uint2 v, int k) {
    const uint32_t w = k < 4 ? v.x : v.y;
#ifdef __CUDA_ARCH__
    const uint32_t bits = __byte_perm(w, 0x4B000000u, 0x7440u | (uint32_t)(k & 3));
#else
    const uint32_t bits = 0x4B000000u | ((w >> (8 * (k & 3))) & 0xffu);
#endif
    return u2f(bits) - 8388608.f;
}

// Slab tests of the (up to) eight children of node `np` against the ray of `t` over (0, t.tMax); returns the mask of the
// children whose box the ray may touch.
// With inv = 1 / d as the traversal holds it, plane q of an axis lies at t* = (origin + q * step - o) * inv.  Computed
// as fma(q, A, B) with A = step * inv (exact: a power of two) and B = (origin - o) * inv (two roundings), the result is
// within 2^-23 |B| + 2^-24 |t*| of t*: the near planes take B - 2^-21 |B|, the far planes B + 2^-21 |B|, and the exit
// distance is widened by 1 + 2^-21 before the comparison, so a box the exact ray touches is never rejected.  A NaN
// (0 * inf, inf - inf: a ray parallel to the slab) drops out of fmaxf / fminf, i.e. that slab does not constrain.
GNX_D uint32_t node8_test(const uint4 *np, const Trav &t) {
    const uint4 w0 = ldg(np), q0 = ldg(np + 3), q1 = ldg(np + 4), q2 = ldg(np + 5);
    const float kEps = 4.76837158203125e-7f, kWiden = 1.f + 4.76837158203125e-7f;  // 2^-21
    const float Ax = u2f((w0.w & 0xffu) << 23) * t.invDir.x, Ay = u2f(((w0.w >> 8) & 0xffu) << 23) * t.invDir.y,
                Az = u2f(((w0.w >> 16) & 0xffu) << 23) * t.invDir.z;
    const float Bx = (u2f(w0.x) - t.o.x) * t.invDir.x, By = (u2f(w0.y) - t.o.y) * t.invDir.y, Bz = (u2f(w0.z) - t.o.z) * t.invDir.z;
    const float Ex = fabsf(Bx) * kEps, Ey = fabsf(By) * kEps, Ez = fabsf(Bz) * kEps;
    const float Bnx = Bx - Ex, Bfx = Bx + Ex, Bny = By - Ey, Bfy = By + Ey, Bnz = Bz - Ez, Bfz = Bz + Ez;
    const uint2 lox = make_uint2(q0.x, q0.y), loy = make_uint2(q0.z, q0.w), loz = make_uint2(q1.x, q1.y);
    const uint2 hix = make_uint2(q1.z, q1.w), hiy = make_uint2(q2.x, q2.y), hiz = make_uint2(q2.z, q2.w);
    const bool n0 = t.neg & 1, n1 = t.neg & 2, n2 = t.neg & 4;
    const uint2 nx = n0 ? hix : lox, fx = n0 ? lox : hix, ny = n1 ? hiy : loy, fy = n1 ? loy : hiy, nz = n2 ? hiz : loz, fz = n2 ? loz : hiz;
    uint32_t mask = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float tn = fmaxf(fmaxf(fmaf(node8_byte(nx, k), Ax, Bnx), fmaf(node8_byte(ny, k), Ay, Bny)),
                               fmaxf(fmaf(node8_byte(nz, k), Az, Bnz), 0.f));
        const float tf = fminf(fminf(fmaf(node8_byte(fx, k), Ax, Bfx), fmaf(node8_byte(fy, k), Ay, Bfy)),
                               fminf(fmaf(node8_byte(fz, k), Az, Bfz), t.tMax));
        if (tn <= tf * kWiden) mask |= 1u << k;
    }
    return mask & (w0.w >> 24);
}

GNX_D int node8_child(const uint4 *nodes8, int node, int k) { return ldg((const int *)(nodes8 + (size_t)kNode8Words * node) + 4 + k); }

// Any-hit traversal state on top of Trav (o, invDir, neg, rs, tMax, hit, sp, spill):
//   t.cur    the child to look at next: a Node8 index, a leaf reference, kRefPop ("take the next one") or kRefNone (over)
//   gnode / gmask   the node whose children are being visited and the ones still to visit; stack entries are such pairs
struct Trav8 : Trav {
    int gnode, gmask;
};
GNX_D void trav8_init(const DeviceScene &sc, Trav8 &t) {  // after trav_init (ray set up)
    t.gnode = 0;
    t.gmask = 0;
    t.sp = 0;
    t.cur = sc.n_nodes8 > 0 ? 0 : kRefNone;
}
// next child of the current node, or of the node on top of the stack
GNX_D void trav8_next(const DeviceScene &sc, Trav8 &t, const int2 *stack, int stride, uint32_t sb = 0) {
    if (t.gmask == 0) {
        if (t.sp == 0) { t.cur = kRefNone; return; }
        --t.sp;
        const int2 e = stack_load(t, stack, stride, t.sp, sb);
        t.gnode = e.x;
        t.gmask = e.y;
    }
    const int k = ffs32(t.gmask) - 1;
    t.gmask &= t.gmask - 1;
    t.cur = node8_child(sc.nodes8, t.gnode, k);
}
// one interior node: test its children, park the rest of the current node on the stack, go on with the first child hit
GNX_D void trav8_interior(const DeviceScene &sc, Trav8 &t, int2 *stack, int stride, TraversalCounters &cnt, uint32_t sb = 0) {
    const int node = t.cur;
    const uint32_t m = node8_test(sc.nodes8 + (size_t)kNode8Words * node, t);
    cnt.nodes += 3;  // 96 bytes = three 32-byte node words (the unit of the algorithmic-bytes count)
    if (t.gmask) {
        stack_store(t, stack, stride, t.sp, make_int2(t.gnode, t.gmask), sb);
        ++t.sp;
    }
    t.gnode = node;
    t.gmask = (int)m;
    if (m) {
        const int k = ffs32(t.gmask) - 1;
        t.gmask &= t.gmask - 1;
        t.cur = node8_child(sc.nodes8, node, k);
    } else t.cur = kRefPop;
}
// Whole any-hit query (sequential callers: CPU emulation, tests).
GNX_D bool traverse8_any(const DeviceScene &sc, V3 o, V3 d, float tMax, int2 *stack, int stride, TraversalCounters &cnt) {
    Trav8 t;
    int2 store[kSpillStack];
    t.spill = store;
    trav_init(sc, t, o, d, tMax);
    trav8_init(sc, t);
    while (!trav_done(t)) {
        if (trav_needs_pop(t)) trav8_next(sc, t, stack, stride);
        else if (trav_is_leaf(t)) t.cur = trav_leaf_ref<true>(sc, t, t.cur, cnt) ? kRefNone : kRefPop;
        else trav8_interior(sc, t, stack, stride, cnt);
    }
    return t.hit;
}

}  // namespace gnx
