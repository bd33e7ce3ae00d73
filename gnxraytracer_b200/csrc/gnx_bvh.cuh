// gnx_bvh.cuh — BVH traversal and the watertight ray-triangle test.
//
// Traversal follows BVHAccel::Intersect / IntersectP (accelerator/BVHAccel.cpp:653-729): depth-first,
// near child first by the ray's sign on the node's split axis, slab test of
// Bounds3::IntersectP(ray, invDir, dirIsNeg) (core/Geometry.h:1380-1406) including the
// 1 + 2*gamma(3) widening.  The triangle test is Triangle::Intersect up to the point where the
// reference starts building the SurfaceInteraction (shape/Triangle.cpp:71-168), including the
// double-precision fallback for zero edge functions; the interaction itself is rebuilt once, for
// the final hit only, in the shade stage.
//
// Layout: a device node is four float4 (64 B: an interior LinearBVHNode plus both children's bounds, see
// "Node2" below), a triangle is three float4 (48 B, vertices pre-gathered); both are read with 16-byte
// __ldg loads.  The traversal stack lives in shared memory, one column per thread (stack[level][thread]:
// conflict-free), and spills to a local array only beyond kSmemStack levels.
#pragma once
#include "gnx_scene.cuh"

namespace gnx {

#ifndef GNX_SMEM_STACK
#define GNX_SMEM_STACK 24
#endif
constexpr int kSmemStack = GNX_SMEM_STACK;   // levels kept in shared memory
#ifndef GNX_BVH_WIDTH
#define GNX_BVH_WIDTH 2  // 4 measured no faster (profiles/README.md)
#endif
constexpr int kSpillStack = GNX_BVH_WIDTH == 4 ? 72 : 40;  // further levels in local memory (reference total: 64; a 4-wide node pushes up to 3)

struct TriHit {
    float t, b0, b1, b2;
};

struct TriVerts { V3 p0, p1, p2; };

GNX_D TriVerts load_tri(const float4 *tris, int prim, float4 *cOut = nullptr) {
    const float4 a = ldg(tris + 3 * prim), b = ldg(tris + 3 * prim + 1), c = ldg(tris + 3 * prim + 2);
    if (cOut) *cOut = c;
    TriVerts t;
    t.p0 = V3(a.x, a.y, a.z);
    t.p1 = V3(a.w, b.x, b.y);
    t.p2 = V3(b.z, b.w, c.x);
    return t;
}

GNX_D float permute_get(V3 v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }

// Ray-space constants that the reference recomputes per triangle (shape/Triangle.cpp:88-101).
struct RayShear {
    int kz;  // kx = (kz + 1) % 3, ky = (kx + 1) % 3 are re-derived where needed (one register instead of three)
    float Sx, Sy, Sz;
    GNX_D int kx() const { return kz == 2 ? 0 : kz + 1; }
    GNX_D int ky() const { return kz == 0 ? 2 : kz - 1; }
};
GNX_D RayShear make_shear(V3 d) {
    RayShear r;
    r.kz = max_dimension(vabs(d));
    float dx = permute_get(d, r.kx()), dy = permute_get(d, r.ky()), dz = permute_get(d, r.kz);
    r.Sx = -dx / dz;
    r.Sy = -dy / dz;
    r.Sz = 1.f / dz;
    return r;
}

// shape/Triangle.cpp:82-168.  Returns true and fills `h` when the triangle is hit inside (0, tMax).
GNX_D bool intersect_tri(const TriVerts &tv, V3 o, const RayShear &rs, float tMax, TriHit *h) {
    V3 q0 = tv.p0 - o, q1 = tv.p1 - o, q2 = tv.p2 - o;
    const int kx = rs.kx(), ky = rs.ky();
    float p0x = permute_get(q0, kx), p0y = permute_get(q0, ky), p0z = permute_get(q0, rs.kz);
    float p1x = permute_get(q1, kx), p1y = permute_get(q1, ky), p1z = permute_get(q1, rs.kz);
    float p2x = permute_get(q2, kx), p2y = permute_get(q2, ky), p2z = permute_get(q2, rs.kz);
    p0x += rs.Sx * p0z; p0y += rs.Sy * p0z;
    p1x += rs.Sx * p1z; p1y += rs.Sy * p1z;
    p2x += rs.Sx * p2z; p2y += rs.Sy * p2z;
    float e0 = p1x * p2y - p1y * p2x;
    float e1 = p2x * p0y - p2y * p0x;
    float e2 = p0x * p1y - p0y * p1x;
    if (e0 == 0.0f || e1 == 0.0f || e2 == 0.0f) {
        double p2txp1ty = (double)p2x * (double)p1y, p2typ1tx = (double)p2y * (double)p1x;
        e0 = (float)(p2typ1tx - p2txp1ty);
        double p0txp2ty = (double)p0x * (double)p2y, p0typ2tx = (double)p0y * (double)p2x;
        e1 = (float)(p0typ2tx - p0txp2ty);
        double p1txp0ty = (double)p1x * (double)p0y, p1typ0tx = (double)p1y * (double)p0x;
        e2 = (float)(p1typ0tx - p1txp0ty);
    }
    if ((e0 < 0 || e1 < 0 || e2 < 0) && (e0 > 0 || e1 > 0 || e2 > 0)) return false;
    float det = e0 + e1 + e2;
    if (det == 0) return false;
    p0z *= rs.Sz; p1z *= rs.Sz; p2z *= rs.Sz;
    float tScaled = e0 * p0z + e1 * p1z + e2 * p2z;
    if (det < 0 && (tScaled >= 0 || tScaled < tMax * det)) return false;
    else if (det > 0 && (tScaled <= 0 || tScaled > tMax * det)) return false;
    float invDet = 1 / det;
    float b0 = e0 * invDet, b1 = e1 * invDet, b2 = e2 * invDet;
    float t = tScaled * invDet;
    // conservative t > 0 check (Triangle.cpp:150-168)
    float maxZt = fmaxf(fabsf(p0z), fmaxf(fabsf(p1z), fabsf(p2z)));
    float deltaZ = gamma_n(3) * maxZt;
    float maxXt = fmaxf(fabsf(p0x), fmaxf(fabsf(p1x), fabsf(p2x)));
    float maxYt = fmaxf(fabsf(p0y), fmaxf(fabsf(p1y), fabsf(p2y)));
    float deltaX = gamma_n(5) * (maxXt + maxZt);
    float deltaY = gamma_n(5) * (maxYt + maxZt);
    float deltaE = 2 * (gamma_n(2) * maxXt * maxYt + deltaY * maxXt + deltaX * maxYt);
    float maxE = fmaxf(fabsf(e0), fmaxf(fabsf(e1), fabsf(e2)));
    float deltaT = 3 * (gamma_n(3) * maxE * maxZt + deltaE * maxZt + deltaZ * maxE) * fabsf(invDet);
    if (t <= deltaT) return false;
    h->t = t; h->b0 = b0; h->b1 = b1; h->b2 = b2;
    return true;
}

// A truly degenerate triangle is rejected by the reference when it builds dpdu/dpdv
// (Triangle.cpp:186-196).
GNX_D bool tri_degenerate(const TriVerts &tv) {
    V3 ng = cross(tv.p2 - tv.p0, tv.p1 - tv.p0);
    return length_sq(ng) == 0;
}

struct TraversalCounters { unsigned nodes, tris; };

// ---- traversal -------------------------------------------------------------------------------------
// Device node ("Node2", 64 B = four float4): an INTERIOR node of the reference's LinearBVHNode array
// together with the bounds of both of its children, so one fetch feeds two slab tests and the
// dependent-load chain is half as long:
//   n0 = (c0.lo.xyz, c0.hi.x)  n1 = (c0.hi.yz, c1.lo.xy)  n2 = (c1.lo.z, c1.hi.xyz)  n3 = (ref0, ref1, axis, -)
// c0 is the reference's first child (index + 1), c1 its second child.  A child ref >= 0 is another
// Node2; a negative ref is a leaf, ~ref = primitive offset | (count - 1) << 27.  axis 0..2 picks the near
// child from the ray's sign exactly as BVHAccel::Intersect does; axis 3 means "c0 first" (chained
// oversized leaves).  The order in which leaves are intersected is the reference's order, so the
// closest hit is the same hit (not merely an equally close one).
constexpr int kRefNone = 0x7fffffff;
constexpr int kRefPop = 0x7ffffffe;   // traversal state: "take the next entry from the stack" (see trav_pop)
constexpr int kLeafMaxPrims = 16;
GNX_HD int leaf_ref(int offset, int count) { return ~(offset | ((count - 1) << 27)); }

struct Trav {
    V3 o, invDir;
    RayShear rs;
    float tMax;
    int cur, sp, neg;  // neg: bit k set when invDir[k] < 0
    bool hit;
    TriHit h;
    int prim;
    int2 *spill;  // kSpillStack entries of the caller's local memory (kept out of the struct so that cur / sp /
                  // tMax stay in registers: with the array inside, every step stored them to the local frame)
};
// Trav with its own spill storage, for the sequential callers.
struct TravLocal : Trav {
    int2 store[kSpillStack];
    GNX_D TravLocal() { spill = store; }
};

GNX_D void trav_init(const DeviceScene &sc, Trav &t, V3 o, V3 d, float tMax) {
    t.o = o;
    t.invDir = V3(1.f / d.x, 1.f / d.y, 1.f / d.z);
    t.neg = (t.invDir.x < 0 ? 1 : 0) | (t.invDir.y < 0 ? 2 : 0) | (t.invDir.z < 0 ? 4 : 0);
    t.rs = make_shear(d);
    t.tMax = tMax;
    t.sp = 0;
    t.hit = false;
    t.prim = -1;
    t.cur = sc.n_nodes2 > 0 ? 0 : kRefNone;
}

// Bounds3::IntersectP(ray, invDir, dirIsNeg), core/Geometry.h:1380-1406; *tminOut is the entry distance
// the reference compares with ray.tMax.
GNX_D bool slab_test(const Trav &t, float lox, float loy, float loz, float hix, float hiy, float hiz, float *tminOut) {
    // Same operations and comparisons as the reference, without its early returns: a warp gathers 32
    // unrelated nodes, so some lane always needs the z slab, and the branches only added divergence
    // bookkeeping.  (Selects, not fminf/fmaxf: a NaN from 0 * inf must propagate as it does in the reference.)
    const float widen = 1 + 2 * gamma_n(3);
    const bool n0 = t.neg & 1, n1 = t.neg & 2, n2 = t.neg & 4;
    float tmin = ((n0 ? hix : lox) - t.o.x) * t.invDir.x, tmax = ((n0 ? lox : hix) - t.o.x) * t.invDir.x;
    float tymin = ((n1 ? hiy : loy) - t.o.y) * t.invDir.y, tymax = ((n1 ? loy : hiy) - t.o.y) * t.invDir.y;
    tmax *= widen;
    tymax *= widen;
    const bool missXY = (tmin > tymax) | (tymin > tmax);
    tmin = tymin > tmin ? tymin : tmin;
    tmax = tymax < tmax ? tymax : tmax;
    float tzmin = ((n2 ? hiz : loz) - t.o.z) * t.invDir.z, tzmax = ((n2 ? loz : hiz) - t.o.z) * t.invDir.z;
    tzmax *= widen;
    const bool missZ = (tmin > tzmax) | (tzmin > tmax);
    tmin = tzmin > tmin ? tzmin : tmin;
    tmax = tzmax < tmax ? tzmax : tmax;
    *tminOut = tmin;
    return !missXY & !missZ & (tmin < t.tMax) & (tmax > 0);
}

// The traversal is split the "while-while" way so that the lanes of a warp do the same kind of work at
// the same time: trav_interior handles one interior node (two slab tests, push far, go near),
// trav_leaf intersects one leaf and pops.  t.cur is an interior index (0 <= cur < kRefNone), a leaf
// reference (negative) or kRefNone when the traversal is over.  `stack` is the calling thread's column
// of a shared-memory stack (entry k at stack[k * stride]); deeper levels spill into t.spill.
GNX_D bool trav_is_interior(const Trav &t) { return (unsigned)t.cur < (unsigned)kRefPop; }
GNX_D bool trav_needs_pop(const Trav &t) { return t.cur == kRefPop; }
GNX_D bool trav_is_leaf(const Trav &t) { return t.cur < 0; }
GNX_D bool trav_done(const Trav &t) { return t.cur == kRefNone; }

// Stack entry k of the calling thread lives at stack[k * stride] for the first kSmemStack levels and in t.spill
// beyond.  `sb`, when non-zero, is the shared-state-space address of stack[0], computed once by the kernel: through
// the generic pointer every access re-derived the shared window base (S2R CgaCtaId, S2R TID, two LEAs).
GNX_D void stack_store(Trav &t, int2 *stack, int stride, int level, int2 e, uint32_t sb = 0) {
    if (level < kSmemStack) {
#ifdef __CUDA_ARCH__
        if (sb) {
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(sb + (uint32_t)(level * stride) * 8u), "r"(e.x), "r"(e.y) : "memory");
            return;
        }
#endif
        stack[level * stride] = e;
    } else {
        t.spill[level - kSmemStack] = e;
    }
}
GNX_D int2 stack_load(const Trav &t, const int2 *stack, int stride, int level, uint32_t sb = 0) {
    if (level < kSmemStack) {
#ifdef __CUDA_ARCH__
        if (sb) {
            int2 e;
            asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(e.x), "=r"(e.y) : "r"(sb + (uint32_t)(level * stride) * 8u) : "memory");
            return e;
        }
#endif
        return stack[level * stride];
    }
    return t.spill[level - kSmemStack];
}
// shared-state-space address of p, pinned in a register (the compiler would otherwise rematerialize it)
GNX_D uint32_t stack_shared_base(const int2 *p) {
#ifdef __CUDA_ARCH__
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p), b;
    asm volatile("mov.u32 %0, %1;" : "=r"(b) : "r"(a));
    return b;
#else
    return 0;
#endif
}

// pop; the reference re-tests a popped node's box against the current tMax (tMin < ray.tMax).  An any-hit
// ray's tMax never shrinks, so its entries stay valid.  trav_interior / trav_leaf only request the pop
// (cur = kRefPop): the caller runs it for all requesting lanes of the warp together.
template <bool ANY>
GNX_D void trav_pop(Trav &t, const int2 *stack, int stride, uint32_t sb = 0) {
    while (true) {
        if (t.sp == 0) { t.cur = kRefNone; return; }
        --t.sp;
        const int2 e = stack_load(t, stack, stride, t.sp, sb);
        if (ANY || i2f(e.y) < t.tMax) { t.cur = e.x; return; }
    }
}

// One pop attempt without divergence, for the warp-synchronous kernel: every lane reads the entry under its stack
// top (a shared-memory load whether it needs it or not) and lanes with cur == kRefPop take it.  An entry that the
// current tMax culls leaves the lane in the kRefPop state: it pops again in the next round.  Levels beyond the shared
// part of the stack take the (rare) branch.
template <bool ANY>
GNX_D void trav_pop_round(Trav &t, const int2 *stack, int stride, uint32_t sb, bool active) {
    const bool need = active && t.cur == kRefPop;
    const int top = t.sp - 1;
    if (need && top >= kSmemStack) {
        const int2 e = t.spill[top - kSmemStack];
        t.sp = top;
        if (ANY || i2f(e.y) < t.tMax) t.cur = e.x;
        return;
    }
    const int2 e = stack_load(t, stack, stride, top < 0 ? 0 : (top >= kSmemStack ? kSmemStack - 1 : top), sb);
    if (need) {
        if (top < 0) t.cur = kRefNone;
        else {
            t.sp = top;
            if (ANY || i2f(e.y) < t.tMax) t.cur = e.x;
        }
    }
}

#if GNX_BVH_WIDTH == 4
// Node4 (128 B = eight float4): lo.x[4] lo.y[4] lo.z[4] hi.x[4] hi.y[4] hi.z[4] refs[4] (axTop, ax0, ax1, -).
// Slots 0-1 are the children of the source node's first child, slots 2-3 of its second child (a child
// that is a leaf occupies the first slot of its pair).  Visiting order = the reference's depth-first order:
// near pair first by the sign on axTop, near slot first inside a pair by ax0 / ax1 (axis 3: slot order).
GNX_D void trav_interior(const DeviceScene &sc, Trav &t, int2 *stack, int stride, TraversalCounters &cnt, uint32_t sb = 0) {
    const float4 *np = sc.nodes2 + 8 * (size_t)t.cur;
    const float4 lx = ldg(np), ly = ldg(np + 1), lz = ldg(np + 2), hx = ldg(np + 3), hy = ldg(np + 4), hz = ldg(np + 5);
    const float4 rf = ldg(np + 6), ax = ldg(np + 7);
    const int ref[4] = {f2i(rf.x), f2i(rf.y), f2i(rf.z), f2i(rf.w)};
    const float lox[4] = {lx.x, lx.y, lx.z, lx.w}, loy[4] = {ly.x, ly.y, ly.z, ly.w}, loz[4] = {lz.x, lz.y, lz.z, lz.w};
    const float hix[4] = {hx.x, hx.y, hx.z, hx.w}, hiy[4] = {hy.x, hy.y, hy.z, hy.w}, hiz[4] = {hz.x, hz.y, hz.z, hz.w};
    float tmin[4];
    bool hit[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        hit[k] = false;
        tmin[k] = 0;
        if (ref[k] != kRefNone) { ++cnt.nodes; hit[k] = slab_test(t, lox[k], loy[k], loz[k], hix[k], hiy[k], hiz[k], &tmin[k]); }
    }
    const int axTop = f2i(ax.x), ax0 = f2i(ax.y), ax1 = f2i(ax.z);
    const int swapTop = axTop < 3 && ((t.neg >> axTop) & 1);
    const int swap0 = ax0 < 3 && ((t.neg >> ax0) & 1), swap1 = ax1 < 3 && ((t.neg >> ax1) & 1);
    // visiting order of the four slots
    int order[4];
    order[swapTop ? 2 : 0] = swap0 ? 1 : 0;
    order[swapTop ? 3 : 1] = swap0 ? 0 : 1;
    order[swapTop ? 0 : 2] = swap1 ? 3 : 2;
    order[swapTop ? 1 : 3] = swap1 ? 2 : 3;
    // push the hit slots in reverse visiting order, then continue with the first one
    int first = -1;
#pragma unroll
    for (int k = 3; k >= 0; --k) {
        const int sl = order[k];
        if (hit[sl]) {
            if (first >= 0) {
                const int2 e = make_int2(ref[first], f2i(tmin[first]));
                stack_store(t, stack, stride, t.sp, e, sb);
                ++t.sp;
            }
            first = sl;
        }
    }
    t.cur = first >= 0 ? ref[first] : kRefPop;
}
#else
GNX_D void trav_interior(const DeviceScene &sc, Trav &t, int2 *stack, int stride, TraversalCounters &cnt, uint32_t sb = 0) {
    const float4 *np = sc.nodes2 + 4 * (size_t)t.cur;
    float4 n0, n1, n2, n3;
    ldg256(np, &n0, &n1);
    ldg256(np + 2, &n2, &n3);
    const int ref0 = f2i(n3.x), ref1 = f2i(n3.y), axis = f2i(n3.z);
    float tmin0, tmin1;
    const bool hit0 = slab_test(t, n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, &tmin0) & (ref0 != kRefNone);
    const bool hit1 = slab_test(t, n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, &tmin1) & (ref1 != kRefNone);
    cnt.nodes += 2;
    if (ref1 == kRefNone) --cnt.nodes;  // single-child link of a chained oversized leaf (ref0 is never empty)
    const bool nearIs1 = (axis < 3) & ((t.neg >> axis) & 1);
    const int refN = nearIs1 ? ref1 : ref0, refF = nearIs1 ? ref0 : ref1;
    const bool hitN = nearIs1 ? hit1 : hit0, hitF = nearIs1 ? hit0 : hit1;
    if (hitN & hitF) {
        const int2 e = make_int2(refF, f2i(nearIs1 ? tmin0 : tmin1));
        stack_store(t, stack, stride, t.sp, e, sb);
        ++t.sp;
    }
    t.cur = hitN ? refN : (hitF ? refF : kRefPop);
}

#endif

// Intersects the triangles of leaf `ref` in order.  Returns true when an ANY-hit query is over.
template <bool ANY>
GNX_D bool trav_leaf_ref(const DeviceScene &sc, Trav &t, int ref, TraversalCounters &cnt) {
    const int x = ~ref, offset = x & 0x7ffffff, count = (x >> 27) + 1;
    for (int i = 0; i < count; ++i) {
        const int prim = offset + i;
        const TriVerts tv = load_tri(sc.tris, prim);
        ++cnt.tris;
        TriHit h;
        if (intersect_tri(tv, t.o, t.rs, t.tMax, &h) && !tri_degenerate(tv)) {
            t.hit = true;
            t.tMax = h.t;
            t.h = h;
            t.prim = prim;
            if (ANY) return true;
        }
    }
    return false;
}
template <bool ANY>
GNX_D void trav_leaf(const DeviceScene &sc, Trav &t, const int2 *stack, int stride, TraversalCounters &cnt) {
    t.cur = trav_leaf_ref<ANY>(sc, t, t.cur, cnt) ? kRefNone : kRefPop;
}

// One step for sequential callers: an interior node and whatever leaves follow it.  Returns true when
// the traversal is finished (t.hit / t.h / t.prim hold the result).
template <bool ANY>
GNX_D bool trav_step(const DeviceScene &sc, Trav &t, int2 *stack, int stride, TraversalCounters &cnt) {
    if (trav_is_interior(t)) trav_interior(sc, t, stack, stride, cnt);
    if (trav_needs_pop(t)) trav_pop<ANY>(t, stack, stride);
    while (trav_is_leaf(t)) {
        trav_leaf<ANY>(sc, t, stack, stride, cnt);
        if (trav_needs_pop(t)) trav_pop<ANY>(t, stack, stride);
    }
    return trav_done(t);
}

// Whole traversal in one call (parity hooks, light-table build, CPU emulation).
template <bool ANY>
GNX_D bool traverse(const DeviceScene &sc, V3 o, V3 d, float tMax, int2 *stack, int stride, int *primOut, TriHit *hitOut,
                    TraversalCounters &cnt) {
    TravLocal t;
    trav_init(sc, t, o, d, tMax);
    while (!trav_step<ANY>(sc, t, stack, stride, cnt)) {}
    if (t.hit) { *primOut = t.prim; *hitOut = t.h; }
    return t.hit;
}

}  // namespace gnx
