// gnx_bvh8.cuh — compressed 8-wide BVH (second half of SURVEY §8 f-2; node to beat: accelerator/BVHAccel.cpp:54-65).
//
// ANY-HIT queries (shadow rays and the environment's MIS rays, VisibilityTester::Unoccluded / Scene::IntersectP,
// core/Light.cpp:14-31, accelerator/BVHAccel.cpp:689-729) only ask "is there a triangle with 0 < t < tMax": any tree
// over the same ordered triangles, visited in any order, gives the same answer as long as its box tests are
// conservative — the triangle test (Triangle::IntersectP's arithmetic, gnx_bvh.cuh) decides, exactly as in the reference.
//
// CLOSEST-HIT queries must return the reference's hit.  When ONE triangle is clearly the closest, every conservative
// traversal finds it, and its (t, b0, b1, b2) come from the same triangle arithmetic whatever the path to the leaf.
// Only when two candidates lie within kTieBand of each other does the reference's answer depend on its visiting order
// (it accepts t <= tMax in a scaled space, Triangle.cpp:137-140, and culls boxes with tMin < tMax): such a ray is FLAGGED
// here (Trav8::tie) and traced again by the caller on the reference-order two-child tree (gnx_bvh.cuh).  Ties are rare
// (coplanar duplicates, rays through a shared edge), the retrace keeps the result bit-equal to the reference's.
//
// That freedom is spent on a shorter dependent-load chain: eight children per node, child boxes quantised to 8 bits on a
// per-node power-of-two grid (Ylitie, Karras, Laine 2017), 96 bytes per node instead of 8 x 32, all of it fetched with six
// independent 16-byte loads, one stack entry per NODE (node, mask of the children still to visit).  Children sit in
// OCTANT slots: slot s holds a child lying towards (s&1 ? +x : -x, s&2 ? +y : -y, s&4 ? +z : -z) of the node's centre,
// so visiting the slots in increasing (s XOR ray octant) is a front-to-back order without sorting.
//
// Node8 = six 16-byte words (read with LDG.128):
//   w0  origin.x origin.y origin.z (float bits)   ex | ey << 8 | ez << 16 | valid << 24
//       (e* = biased IEEE exponent of the grid step 2^(e - 127) on that axis; valid = mask of the occupied slots)
//   w1  child references 0..3      w2  child references 4..7
//       (>= 0: Node8 index; < 0: leaf, ~ref = primitive offset | (count - 1) << 27, the encoding of gnx_bvh.cuh)
//   w3  qlo.x[0..7] qlo.y[0..7]    w4  qlo.z[0..7] qhi.x[0..7]    w5  qhi.y[0..7] qhi.z[0..7]   (one byte per child)
// Child k's box is [origin + qlo * step, origin + qhi * step], a superset of the box the two-child tree holds for it.
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

// Slab tests of the (up to) eight children of node `np` against the ray of `t` over (0, tLimit).  Returns the mask of the
// children whose box the ray may touch IN VISITING ORDER: bit j stands for slot j ^ t.neg, so the lowest set bit is the
// front-most child (see the octant slots above).  r0 / r1 receive the eight child references.
// With inv = 1 / d as the traversal holds it, plane q of an axis lies at t* = (origin + q * step - o) * inv.  Computed
// as fma(q, A, B) with A = step * inv (exact: a power of two) and B = (origin - o) * inv (two roundings), the result is
// within 2^-23 |B| + 2^-24 |t*| of t*: the near planes take B - 2^-21 |B|, the far planes B + 2^-21 |B|, and the exit
// distance is widened by 1 + 2^-21 before the comparison, so a box the exact ray touches is never rejected.  A NaN
// (0 * inf, inf - inf: a ray parallel to the slab) drops out of fmaxf / fminf, i.e. that slab does not constrain.
GNX_D uint32_t node8_test(const uint4 *np, const Trav &t, float tLimit, uint4 &r0, uint4 &r1) {
    const uint4 w0 = ldg(np), q0 = ldg(np + 3), q1 = ldg(np + 4), q2 = ldg(np + 5);
    r0 = ldg(np + 1);
    r1 = ldg(np + 2);
    const float kEps = 4.76837158203125e-7f, kWiden = 1.f + 4.76837158203125e-7f;  // 2^-21
    // |1 / d| is capped at 2^100: with an infinite (or overflowing) factor both products below are inf / NaN and the slab
    // would not constrain at all — the ray through the exact image centre (d = (0, 0, -1)) then visited every node.  With
    // the cap, a ray parallel to a slab sees both planes at +-1e30 x distance: inside the slab they straddle 0, outside
    // they lie on one side, beyond any tLimit or behind the origin.  Scaling 1 / d DOWN keeps the signs, so it stays conservative.
    const float kInvCap = 1.2676506e30f;
    const float ix = fminf(fmaxf(t.invDir.x, -kInvCap), kInvCap), iy = fminf(fmaxf(t.invDir.y, -kInvCap), kInvCap),
                iz = fminf(fmaxf(t.invDir.z, -kInvCap), kInvCap);
    const float Ax = u2f((w0.w & 0xffu) << 23) * ix, Ay = u2f(((w0.w >> 8) & 0xffu) << 23) * iy, Az = u2f(((w0.w >> 16) & 0xffu) << 23) * iz;
    const float Bx = (u2f(w0.x) - t.o.x) * ix, By = (u2f(w0.y) - t.o.y) * iy, Bz = (u2f(w0.z) - t.o.z) * iz;
    const float Ex = fabsf(Bx) * kEps, Ey = fabsf(By) * kEps, Ez = fabsf(Bz) * kEps;
    const float Bnx = Bx - Ex, Bfx = Bx + Ex, Bny = By - Ey, Bfy = By + Ey, Bnz = Bz - Ez, Bfz = Bz + Ez;
    const uint2 lox = make_uint2(q0.x, q0.y), loy = make_uint2(q0.z, q0.w), loz = make_uint2(q1.x, q1.y);
    const uint2 hix = make_uint2(q1.z, q1.w), hiy = make_uint2(q2.x, q2.y), hiz = make_uint2(q2.z, q2.w);
    const bool n0 = t.neg & 1, n1 = t.neg & 2, n2 = t.neg & 4;
    const uint2 nx = n0 ? hix : lox, fx = n0 ? lox : hix, ny = n1 ? hiy : loy, fy = n1 ? loy : hiy, nz = n2 ? hiz : loz, fz = n2 ? loz : hiz;
    const uint32_t valid = w0.w >> 24;
    uint32_t mask = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float tn = fmaxf(fmaxf(fmaf(node8_byte(nx, k), Ax, Bnx), fmaf(node8_byte(ny, k), Ay, Bny)),
                               fmaxf(fmaf(node8_byte(nz, k), Az, Bnz), 0.f));
        const float tf = fminf(fminf(fmaf(node8_byte(fx, k), Ax, Bfx), fmaf(node8_byte(fy, k), Ay, Bfy)),
                               fminf(fmaf(node8_byte(fz, k), Az, Bfz), tLimit));
        if (tn <= tf * kWiden && ((valid >> k) & 1u)) mask |= 1u << (k ^ t.neg);
    }
    return mask;
}

GNX_D int node8_child(const uint4 *nodes8, int node, int k) { return ldg((const int *)(nodes8 + (size_t)kNode8Words * node) + 4 + k); }
// reference k (0..7) out of the two words already in registers
GNX_D int node8_pick(const uint4 &r0, const uint4 &r1, int k) {
    const uint32_t a0 = (k & 1) ? r0.y : r0.x, a1 = (k & 1) ? r0.w : r0.z, a2 = (k & 1) ? r1.y : r1.x, a3 = (k & 1) ? r1.w : r1.z;
    const uint32_t b0 = (k & 2) ? a1 : a0, b1 = (k & 2) ? a3 : a2;
    return (int)((k & 4) ? b1 : b0);
}

// Two candidates of a closest-hit query closer together than this (relative) make the reference's answer depend on its
// visiting order: the ray is flagged and retraced in reference order.  2^-16: three orders of magnitude above the
// rounding of the acceptance tests involved, far below the spacing of distinct surfaces.
constexpr float kTieBand = 1.52587890625e-5f;

// Traversal state on top of Trav (o, invDir, neg, rs, tMax, hit, sp, spill):
//   t.cur    the child to look at next: a Node8 index, a leaf reference, kRefPop ("take the next one") or kRefNone (over)
//   gnode / gmask   the node whose children are being visited and the ones still to visit (visiting-order bits);
//                   stack entries are such pairs
//   tie      closest-hit queries: the best hit so far has a rival within kTieBand
struct Trav8 : Trav {
    int gnode, gmask;
    bool tie;
};
GNX_D void trav8_init(const DeviceScene &sc, Trav8 &t) {  // after trav_init (ray set up)
    t.gnode = 0;
    t.gmask = 0;
    t.sp = 0;
    t.tie = false;
    // a ray with a NaN in it fails every comparison of the reference's slab test (no hit); here a NaN drops out of
    // fminf / fmaxf and nothing would ever be culled
    const bool bad = (t.invDir.x != t.invDir.x) | (t.invDir.y != t.invDir.y) | (t.invDir.z != t.invDir.z) |
                     !(fabsf(t.o.x) + fabsf(t.o.y) + fabsf(t.o.z) < GNX_INF);
    t.cur = (sc.n_nodes8 > 0 && !bad) ? 0 : kRefNone;
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
    const int j = ffs32(t.gmask) - 1;
    t.gmask &= t.gmask - 1;
    t.cur = node8_child(sc.nodes8, t.gnode, j ^ t.neg);
}
// one interior node: test its children, park the rest of the current node on the stack, go on with the front-most child hit.
// CLOSEST: boxes are culled against the best hit so far, widened by the tie band (a rival inside the band must be seen).
template <bool CLOSEST>
GNX_D void trav8_interior(const DeviceScene &sc, Trav8 &t, int2 *stack, int stride, TraversalCounters &cnt, uint32_t sb = 0) {
    const int node = t.cur;
    uint4 r0, r1;
    const float tLimit = CLOSEST ? t.tMax * (1.f + 2.f * kTieBand) : t.tMax;
    const uint32_t m = node8_test(sc.nodes8 + (size_t)kNode8Words * node, t, tLimit, r0, r1);
    cnt.nodes += 3;  // 96 bytes = three 32-byte node words (the unit of the algorithmic-bytes count)
    if (t.gmask) {
        stack_store(t, stack, stride, t.sp, make_int2(t.gnode, t.gmask), sb);
        ++t.sp;
    }
    t.gnode = node;
    t.gmask = (int)m;
    if (m) {
        const int j = ffs32(t.gmask) - 1;
        t.gmask &= t.gmask - 1;
        t.cur = node8_pick(r0, r1, j ^ t.neg);
    } else t.cur = kRefPop;
}
// Closest-hit leaf: the triangles of leaf `ref` against the best hit so far, with the tie bookkeeping described at the
// top of the file.  Accepts candidates up to the upper edge of the band so that a rival behind the best hit is seen too.
GNX_D void trav8_leaf_closest(const DeviceScene &sc, Trav8 &t, int ref, TraversalCounters &cnt) {
    const int x = ~ref, offset = x & 0x7ffffff, count = (x >> 27) + 1;
    for (int i = 0; i < count; ++i) {
        const int prim = offset + i;
        const TriVerts tv = load_tri(sc.tris, prim);
        ++cnt.tris;
        TriHit h;
        const float lim = t.hit ? t.tMax * (1.f + kTieBand) : t.tMax;
        if (intersect_tri(tv, t.o, t.rs, lim, &h) && !tri_degenerate(tv)) {
            if (t.hit && h.t >= t.tMax * (1.f - kTieBand)) {
                t.tie = true;
                if (h.t >= t.tMax) continue;  // the best hit stays (its rival is on record)
            } else t.tie = false;              // clearly in front of everything seen so far
            t.hit = true;
            t.tMax = h.t;
            t.h = h;
            t.prim = prim;
        }
    }
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
        else trav8_interior<false>(sc, t, stack, stride, cnt);
    }
    return t.hit;
}
// Whole closest-hit query; *tieOut tells the caller to repeat it on the two-child tree (traverse<false>).
GNX_D bool traverse8_closest(const DeviceScene &sc, V3 o, V3 d, float tMax, int2 *stack, int stride, int *primOut, TriHit *hitOut,
                             bool *tieOut, TraversalCounters &cnt) {
    Trav8 t;
    int2 store[kSpillStack];
    t.spill = store;
    trav_init(sc, t, o, d, tMax);
    trav8_init(sc, t);
    while (!trav_done(t)) {
        if (trav_needs_pop(t)) trav8_next(sc, t, stack, stride);
        else if (trav_is_leaf(t)) { trav8_leaf_closest(sc, t, t.cur, cnt); t.cur = kRefPop; }
        else trav8_interior<true>(sc, t, stack, stride, cnt);
    }
    if (t.hit) { *primOut = t.prim; *hitOut = t.h; }
    *tieOut = t.tie;
    return t.hit;
}

// Closest hit for the per-lane kernels (Whitted / DirectLighting recursion, VolPath): GNX_PERLANE_CLOSEST8=1 sends them
// through the wide tree first (same tie rule).
#ifndef GNX_PERLANE_CLOSEST8
#define GNX_PERLANE_CLOSEST8 0
#endif
GNX_D bool closest_hit(const DeviceScene &sc, V3 o, V3 d, float tMax, int2 *stack, int stride, int *primOut, TriHit *hitOut,
                       TraversalCounters &cnt) {
#if GNX_PERLANE_CLOSEST8
    if (sc.nodes8) {
        bool tie;
        const bool found = traverse8_closest(sc, o, d, tMax, stack, stride, primOut, hitOut, &tie, cnt);
        if (!tie) return found;
    }
#endif
    return traverse<false>(sc, o, d, tMax, stack, stride, primOut, hitOut, cnt);
}

// Closest hit of the ray `t` has been set up for (trav_init): the 8-wide tree when the scene routes closest-hit queries
// there, the reference-order two-child tree otherwise and for a flagged ray (sequential callers).
GNX_D void closest_hit_run(const DeviceScene &sc, TravLocal &t, int2 *stack, int stride, TraversalCounters &cnt) {
    if (sc.nodes8 && sc.wide_closest) {
        Trav8 t8;
#ifdef GNX_DEBUG_LONG_RAYS
        const unsigned n0 = cnt.nodes;
#endif
        static_cast<Trav &>(t8) = t;
        trav8_init(sc, t8);
        while (!trav_done(t8)) {
            if (trav_needs_pop(t8)) trav8_next(sc, t8, stack, stride);
            else if (trav_is_leaf(t8)) { trav8_leaf_closest(sc, t8, t8.cur, cnt); t8.cur = kRefPop; }
            else trav8_interior<true>(sc, t8, stack, stride, cnt);
        }
#ifdef GNX_DEBUG_LONG_RAYS
        if (cnt.nodes - n0 > 30000) fprintf(stderr, "long ray: %u nodes o=(%g %g %g) inv=(%g %g %g) tMax=%g hit=%d\n", cnt.nodes - n0, t.o.x, t.o.y, t.o.z, t.invDir.x, t.invDir.y, t.invDir.z, t8.tMax, (int)t8.hit);
#endif
        if (!t8.tie) { t.hit = t8.hit; t.h = t8.h; t.prim = t8.prim; t.tMax = t8.tMax; t.cur = kRefNone; return; }
    }
    while (!trav_step<false>(sc, t, stack, stride, cnt)) {}
}

}  // namespace gnx
