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
// Layout: a node is two float4 (32 B, the reference's LinearBVHNode bit for bit), a triangle is three
// float4 (48 B, vertices pre-gathered); both are read with 16-byte __ldg loads.  The traversal stack
// lives in shared memory, one column per thread (stack[level][thread]: conflict-free), and spills
// to a local array only beyond kSmemStack levels.
#pragma once
#include "gnx_scene.cuh"

namespace gnx {

constexpr int kSmemStack = 24;   // levels kept in shared memory
constexpr int kSpillStack = 40;  // further levels in local memory (reference total: 64)

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
    int kx, ky, kz;
    float Sx, Sy, Sz;
};
GNX_D RayShear make_shear(V3 d) {
    RayShear r;
    r.kz = max_dimension(vabs(d));
    r.kx = r.kz + 1; if (r.kx == 3) r.kx = 0;
    r.ky = r.kx + 1; if (r.ky == 3) r.ky = 0;
    float dx = permute_get(d, r.kx), dy = permute_get(d, r.ky), dz = permute_get(d, r.kz);
    r.Sx = -dx / dz;
    r.Sy = -dy / dz;
    r.Sz = 1.f / dz;
    return r;
}

// shape/Triangle.cpp:82-168.  Returns true and fills `h` when the triangle is hit inside (0, tMax).
GNX_D bool intersect_tri(const TriVerts &tv, V3 o, const RayShear &rs, float tMax, TriHit *h) {
    V3 q0 = tv.p0 - o, q1 = tv.p1 - o, q2 = tv.p2 - o;
    float p0x = permute_get(q0, rs.kx), p0y = permute_get(q0, rs.ky), p0z = permute_get(q0, rs.kz);
    float p1x = permute_get(q1, rs.kx), p1y = permute_get(q1, rs.ky), p1z = permute_get(q1, rs.kz);
    float p2x = permute_get(q2, rs.kx), p2y = permute_get(q2, rs.ky), p2z = permute_get(q2, rs.kz);
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

// Closest-hit (ANY=false) or any-hit (ANY=true).  `stack` is the calling thread's column of the
// block's shared-memory stack: entry k lives at stack[k * stride].
template <bool ANY>
GNX_D bool traverse(const DeviceScene &sc, V3 o, V3 d, float tMax, int *stack, int stride, int *primOut,
                    TriHit *hitOut, TraversalCounters &cnt) {
    if (sc.n_nodes == 0) return false;
    const V3 invDir(1.f / d.x, 1.f / d.y, 1.f / d.z);
    const int neg0 = invDir.x < 0, neg1 = invDir.y < 0, neg2 = invDir.z < 0;
    const RayShear rs = make_shear(d);
    const float widen = 1 + 2 * gamma_n(3);
    int spill[kSpillStack];
    int sp = 0, cur = 0;
    bool hit = false;
    while (true) {
        const float4 n0 = ldg(sc.nodes + 2 * cur), n1 = ldg(sc.nodes + 2 * cur + 1);
        ++cnt.nodes;
        // bounds[dirIsNeg] selects pMin (0) or pMax (1)
        const float bx0 = neg0 ? n0.w : n0.x, bx1 = neg0 ? n0.x : n0.w;
        const float by0 = neg1 ? n1.x : n0.y, by1 = neg1 ? n0.y : n1.x;
        const float bz0 = neg2 ? n1.y : n0.z, bz1 = neg2 ? n0.z : n1.y;
        float tmin = (bx0 - o.x) * invDir.x, tmax = (bx1 - o.x) * invDir.x;
        float tymin = (by0 - o.y) * invDir.y, tymax = (by1 - o.y) * invDir.y;
        tmax *= widen;
        tymax *= widen;
        bool inside = !(tmin > tymax || tymin > tmax);
        if (inside) {
            if (tymin > tmin) tmin = tymin;
            if (tymax < tmax) tmax = tymax;
            float tzmin = (bz0 - o.z) * invDir.z, tzmax = (bz1 - o.z) * invDir.z;
            tzmax *= widen;
            inside = !(tmin > tzmax || tzmin > tmax);
            if (inside) {
                if (tzmin > tmin) tmin = tzmin;
                if (tzmax < tmax) tmax = tzmax;
                inside = (tmin < tMax) && (tmax > 0);
            }
        }
        bool pop = true;
        if (inside) {
            const int offset = f2i(n1.z);
            const unsigned meta = f2u(n1.w);
            const int nPrims = (int)(meta & 0xffffu);
            if (nPrims > 0) {
                for (int i = 0; i < nPrims; ++i) {
                    const int prim = offset + i;
                    const TriVerts tv = load_tri(sc.tris, prim);
                    ++cnt.tris;
                    TriHit h;
                    if (intersect_tri(tv, o, rs, tMax, &h) && !tri_degenerate(tv)) {
                        if (ANY) return true;
                        hit = true;
                        tMax = h.t;
                        *hitOut = h;
                        *primOut = prim;
                    }
                }
            } else {
                const int axis = (int)((meta >> 16) & 0xffu);
                const int negAxis = axis == 0 ? neg0 : (axis == 1 ? neg1 : neg2);
                int far;
                if (negAxis) { far = cur + 1; cur = offset; } else { far = offset; cur = cur + 1; }
                if (sp < kSmemStack) stack[sp * stride] = far; else spill[sp - kSmemStack] = far;
                ++sp;
                pop = false;
            }
        }
        if (pop) {
            if (sp == 0) break;
            --sp;
            cur = sp < kSmemStack ? stack[sp * stride] : spill[sp - kSmemStack];
        }
    }
    return hit;
}

}  // namespace gnx
