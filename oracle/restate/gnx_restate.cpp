// gnx_restate.cpp — ORACLE (TEST INFRASTRUCTURE ONLY): a plain, scalar, single-path-at-a-time CPU
// restatement of the reference's path-tracing hot path over the flattened scene buffers of
// include/gnxrt.h.  It exists to check the CUDA kernels; only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline leg may load it.  The product never does.
//
// Independent of the product code: it includes nothing from gnxraytracer_b200/ (only the ABI header
// that defines the scene buffers), walks the reference's 32-byte LinearBVHNode array with the
// reference's own loop (not the product's two-child nodes), keeps the light distribution in a lazily
// filled voxel map like SpatialLightDistribution, and follows PathIntegrator::Li as one loop per path
// (not a wavefront).
//
// PINNING.  The reference has no tests, golden vectors or fixtures of its own.  This
// restatement is pinned against outputs of the reference itself: tests/test_restate.py compares it with
// oracle/_ref (the unmodified reference compiled here) where that library exists, and with the committed
// fixtures tests/golden/*.npz (generated from the reference by tests/golden/make_golden.py) everywhere.
//
// Coverage: HaltonSampler, PerspectiveCamera, BVHAccel::Intersect/IntersectP, Triangle, Matte (Lambert,
// Oren-Nayar), Mirror, Glass (smooth), Plastic, Metal, DiffuseAreaLight, InfiniteAreaLight, uniform and
// spatial light distributions, PathIntegrator::Li, SamplerIntegrator::Render's box film.
// Not covered (returns -4): rough glass, Disney, textures, media.
#include <omp.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "gnxrt.h"

namespace {

typedef float Float;
const Float Pi = 3.14159265358979323846f, InvPi = 0.31830988618379067154f, Inv2Pi = 0.15915494309189533577f;
const Float PiOver2 = 1.57079632679489661923f, PiOver4 = 0.78539816339744830961f;
const Float Infinity = INFINITY, MachineEpsilon = 5.9604644775390625e-08f;   // core/GNXRayTracer.h:137-148
const Float OneMinusEpsilon = 0.99999994f, ShadowEpsilon = 0.0001f;           // core/RNG.h:14
inline Float gammaN(int n) { return (n * MachineEpsilon) / (1 - n * MachineEpsilon); }  // GNXRayTracer.h:354-357

struct Vec {
    Float x = 0, y = 0, z = 0;
    Vec() {}
    Vec(Float a, Float b, Float c) : x(a), y(b), z(c) {}
    Float operator[](int i) const { return i == 0 ? x : i == 1 ? y : z; }
    Vec operator+(const Vec &o) const { return Vec(x + o.x, y + o.y, z + o.z); }
    Vec operator-(const Vec &o) const { return Vec(x - o.x, y - o.y, z - o.z); }
    Vec operator-() const { return Vec(-x, -y, -z); }
    Vec operator*(Float s) const { return Vec(x * s, y * s, z * s); }
};
inline Vec operator*(Float s, const Vec &v) { return v * s; }
inline Float Dot(const Vec &a, const Vec &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Float AbsDot(const Vec &a, const Vec &b) { return std::abs(Dot(a, b)); }
inline Vec Cross(const Vec &a, const Vec &b) {  // core/Geometry.h:925-931 (double)
    double ax = a.x, ay = a.y, az = a.z, bx = b.x, by = b.y, bz = b.z;
    return Vec((Float)((ay * bz) - (az * by)), (Float)((az * bx) - (ax * bz)), (Float)((ax * by) - (ay * bx)));
}
inline Float LengthSquared(const Vec &v) { return v.x * v.x + v.y * v.y + v.z * v.z; }
inline Vec DivV(const Vec &v, Float f) { Float inv = (Float)1 / f; return Vec(v.x * inv, v.y * inv, v.z * inv); }  // Geometry.h:206-210
inline Vec Normalize(const Vec &v) { return DivV(v, std::sqrt(LengthSquared(v))); }
inline Vec Abs(const Vec &v) { return Vec(std::abs(v.x), std::abs(v.y), std::abs(v.z)); }
inline Vec Faceforward(const Vec &n, const Vec &v) { return Dot(n, v) < 0.f ? -n : n; }
inline void CoordinateSystem(const Vec &v1, Vec *v2, Vec *v3) {  // Geometry.h:988-995
    if (std::abs(v1.x) > std::abs(v1.y)) *v2 = DivV(Vec(-v1.z, 0, v1.x), std::sqrt(v1.x * v1.x + v1.z * v1.z));
    else *v2 = DivV(Vec(0, v1.z, -v1.y), std::sqrt(v1.y * v1.y + v1.z * v1.z));
    *v3 = Cross(v1, *v2);
}

struct Spectrum {  // RGBSpectrum, core/Spectrum.h:395
    Float c[3] = {0, 0, 0};
    Spectrum() {}
    explicit Spectrum(Float v) { c[0] = c[1] = c[2] = v; }
    Spectrum(Float r, Float g, Float b) { c[0] = r; c[1] = g; c[2] = b; }
    bool IsBlack() const { return c[0] == 0 && c[1] == 0 && c[2] == 0; }
    Float y() const { return 0.212671f * c[0] + 0.715160f * c[1] + 0.072169f * c[2]; }
    Float MaxComponentValue() const { return std::max(c[0], std::max(c[1], c[2])); }
    Spectrum operator+(const Spectrum &o) const { return Spectrum(c[0] + o.c[0], c[1] + o.c[1], c[2] + o.c[2]); }
    Spectrum operator-(const Spectrum &o) const { return Spectrum(c[0] - o.c[0], c[1] - o.c[1], c[2] - o.c[2]); }
    Spectrum operator*(const Spectrum &o) const { return Spectrum(c[0] * o.c[0], c[1] * o.c[1], c[2] * o.c[2]); }
    Spectrum operator/(const Spectrum &o) const { return Spectrum(c[0] / o.c[0], c[1] / o.c[1], c[2] / o.c[2]); }
    Spectrum operator*(Float s) const { return Spectrum(c[0] * s, c[1] * s, c[2] * s); }
    Spectrum operator/(Float s) const { return Spectrum(c[0] / s, c[1] / s, c[2] / s); }
    Spectrum &operator+=(const Spectrum &o) { *this = *this + o; return *this; }
    Spectrum &operator*=(const Spectrum &o) { *this = *this * o; return *this; }
    Spectrum Clamp0() const { return Spectrum(std::max(c[0], 0.f), std::max(c[1], 0.f), std::max(c[2], 0.f)); }
};
inline Spectrum operator*(Float s, const Spectrum &v) { return v * s; }
inline Spectrum Sqrt(const Spectrum &s) { return Spectrum(std::sqrt(s.c[0]), std::sqrt(s.c[1]), std::sqrt(s.c[2])); }

inline float NextFloatUp(float v) {  // GNXRayTracer.h:179-192
    if (std::isinf(v) && v > 0.) return v;
    if (v == -0.f) v = 0.f;
    uint32_t ui; memcpy(&ui, &v, 4);
    if (v >= 0) ++ui; else --ui;
    memcpy(&v, &ui, 4); return v;
}
inline float NextFloatDown(float v) {  // GNXRayTracer.h:194-205
    if (std::isinf(v) && v < 0.) return v;
    if (v == 0.f) v = -0.f;
    uint32_t ui; memcpy(&ui, &v, 4);
    if (v > 0) --ui; else ++ui;
    memcpy(&v, &ui, 4); return v;
}
Vec OffsetRayOrigin(const Vec &p, const Vec &pError, const Vec &n, const Vec &w) {  // Geometry.h:1408-1422
    Float d = Dot(Abs(n), pError);
    Vec offset = d * n;
    if (Dot(w, n) < 0) offset = -offset;
    Vec po = p + offset;
    Float *q = &po.x; const Float *o = &offset.x;
    for (int i = 0; i < 3; ++i) { if (o[i] > 0) q[i] = NextFloatUp(q[i]); else if (o[i] < 0) q[i] = NextFloatDown(q[i]); }
    return po;
}

// ---- Halton (samplers/HaltonSampler.cpp, samplers/LowDiscrepancy.cpp) ------------------------------------
struct Halton {
    std::vector<int> primes, sums;
    std::vector<uint16_t> perms;
    gnx_sampler s;
    void Init(const gnx_sampler &smp) {
        s = smp;
        std::vector<char> sieve(8200, 1);
        for (int i = 2; i < 8200 && primes.size() < 1000; ++i) {
            if (!sieve[i]) continue;
            primes.push_back(i);
            for (int j = 2 * i; j < 8200; j += i) sieve[j] = 0;
        }
        int acc = 0;
        for (int p : primes) { sums.push_back(acc); acc += p; }
        if (smp.perms) perms.assign(smp.perms, smp.perms + smp.n_perm_entries);
        else {  // ComputeRadicalInversePermutations with a default RNG (LowDiscrepancy.cpp:2459-2473, core/RNG.h)
            perms.resize(acc);
            uint64_t state = 0x853c49e6748fea9bULL, inc = 0xda3e39cb94b95bdbULL;
            auto next = [&]() {
                uint64_t old = state; state = old * 0x5851f42d4c957f2dULL + inc;
                uint32_t xs = (uint32_t)(((old >> 18u) ^ old) >> 27u), rot = (uint32_t)(old >> 59u);
                return (xs >> rot) | (xs << ((~rot + 1u) & 31));
            };
            uint16_t *p = perms.data();
            for (int prime : primes) {
                for (int j = 0; j < prime; ++j) p[j] = (uint16_t)j;
                for (int i = 0; i < prime; ++i) {
                    uint32_t b = (uint32_t)(prime - i), threshold = (~b + 1u) % b, r;
                    do r = next(); while (r < threshold);
                    std::swap(p[i], p[i + (int)(r % b)]);
                }
                p += prime;
            }
        }
    }
    static uint64_t InverseRadicalInverse(uint64_t inverse, int base, int nDigits) {  // LowDiscrepancy.h:47-56
        uint64_t index = 0;
        for (int i = 0; i < nDigits; ++i) { uint64_t digit = inverse % base; inverse /= base; index = index * base + digit; }
        return index;
    }
    int64_t IndexForSample(int px, int py, int64_t sampleNum) const {  // HaltonSampler.cpp:63-82
        int64_t offset = 0;
        if (s.sample_stride > 1) {
            int pm[2] = {px % 128, py % 128};
            for (int i = 0; i < 2; ++i) {
                uint64_t dimOffset = InverseRadicalInverse(pm[i], i == 0 ? 2 : 3, s.base_exponents[i]);
                offset += dimOffset * (s.sample_stride / s.base_scales[i]) * s.mult_inverse[i];
            }
            offset %= s.sample_stride;
        }
        return offset + sampleNum * s.sample_stride;
    }
    Float RadicalInverse(int baseIndex, uint64_t a) const {  // LowDiscrepancy.cpp:358-405
        if (baseIndex == 0) {
            uint64_t r = 0, n = a;
            for (int i = 0; i < 64; ++i) { r = (r << 1) | (n & 1); n >>= 1; }
            return (Float)(r * 5.4210108624275222e-20);
        }
        const int base = primes[baseIndex];
        const Float invBase = (Float)1 / (Float)base;
        uint64_t reversedDigits = 0;
        Float invBaseN = 1;
        while (a) { uint64_t next = a / base, digit = a - next * base; reversedDigits = reversedDigits * base + digit; invBaseN *= invBase; a = next; }
        return std::min(reversedDigits * invBaseN, OneMinusEpsilon);
    }
    Float Scrambled(int baseIndex, uint64_t a) const {  // LowDiscrepancy.cpp:373-392
        const int base = primes[baseIndex];
        const uint16_t *perm = &perms[sums[baseIndex]];
        const Float invBase = (Float)1 / (Float)base;
        uint64_t reversedDigits = 0;
        Float invBaseN = 1;
        while (a) { uint64_t next = a / base, digit = a - next * base; reversedDigits = reversedDigits * base + perm[digit]; invBaseN *= invBase; a = next; }
        return std::min(invBaseN * (reversedDigits + invBase * perm[0] / (1 - invBase)), OneMinusEpsilon);
    }
    Float SampleDimension(int64_t index, int dim) const {  // HaltonSampler.cpp:85-94
        if (s.sample_at_pixel_center && (dim == 0 || dim == 1)) return 0.5f;
        if (dim == 0) return RadicalInverse(0, index >> s.base_exponents[0]);
        if (dim == 1) return RadicalInverse(1, index / s.base_scales[1]);
        if (dim >= 1000) return 0;
        return Scrambled(dim, index);
    }
};
struct Sampler {  // GlobalSampler::Get1D/Get2D, core/Sampler.cpp:162-179
    const Halton *h; int64_t index; int dim = 0;
    Float Get1D() { return h->SampleDimension(index, dim++); }
    void Get2D(Float u[2]) { u[0] = h->SampleDimension(index, dim); u[1] = h->SampleDimension(index, dim + 1); dim += 2; }
};

// ---- scene view ----------------------------------------------------------------------------------------------
struct Ray { Vec o, d; mutable Float tMax = Infinity; };
struct Hit { int prim = -1; Float b0, b1, b2, t; };

struct Scene {
    const gnx_scene_desc *d;
    Halton halton;
    std::mutex voxMutex;
    std::unordered_map<uint64_t, std::vector<Float>> voxels;  // packed voxel -> [func(n) | cdf(n+1) | funcInt]
    int nVoxels[3];
    std::vector<Float> uniformTab, powerTab;  // func[n] | cdf[n+1] | funcInt

    Vec P(int prim, int v) const { const float *p = d->geom.prim_p + (size_t)prim * 9 + 3 * v; return Vec(p[0], p[1], p[2]); }

    // Triangle::Intersect up to the t test (shape/Triangle.cpp:71-168)
    bool TriangleTest(int prim, const Ray &ray, Hit *h) const {
        Vec p0 = P(prim, 0), p1 = P(prim, 1), p2 = P(prim, 2);
        Vec p0t = p0 - ray.o, p1t = p1 - ray.o, p2t = p2 - ray.o;
        Vec ad = Abs(ray.d);
        int kz = (ad.x > ad.y) ? ((ad.x > ad.z) ? 0 : 2) : ((ad.y > ad.z) ? 1 : 2);
        int kx = kz + 1; if (kx == 3) kx = 0;
        int ky = kx + 1; if (ky == 3) ky = 0;
        Vec dd(ray.d[kx], ray.d[ky], ray.d[kz]);
        p0t = Vec(p0t[kx], p0t[ky], p0t[kz]); p1t = Vec(p1t[kx], p1t[ky], p1t[kz]); p2t = Vec(p2t[kx], p2t[ky], p2t[kz]);
        Float Sx = -dd.x / dd.z, Sy = -dd.y / dd.z, Sz = 1.f / dd.z;
        p0t.x += Sx * p0t.z; p0t.y += Sy * p0t.z;
        p1t.x += Sx * p1t.z; p1t.y += Sy * p1t.z;
        p2t.x += Sx * p2t.z; p2t.y += Sy * p2t.z;
        Float e0 = p1t.x * p2t.y - p1t.y * p2t.x, e1 = p2t.x * p0t.y - p2t.y * p0t.x, e2 = p0t.x * p1t.y - p0t.y * p1t.x;
        if (e0 == 0.0f || e1 == 0.0f || e2 == 0.0f) {
            e0 = (float)((double)p2t.y * (double)p1t.x - (double)p2t.x * (double)p1t.y);
            e1 = (float)((double)p0t.y * (double)p2t.x - (double)p0t.x * (double)p2t.y);
            e2 = (float)((double)p1t.y * (double)p0t.x - (double)p1t.x * (double)p0t.y);
        }
        if ((e0 < 0 || e1 < 0 || e2 < 0) && (e0 > 0 || e1 > 0 || e2 > 0)) return false;
        Float det = e0 + e1 + e2;
        if (det == 0) return false;
        p0t.z *= Sz; p1t.z *= Sz; p2t.z *= Sz;
        Float tScaled = e0 * p0t.z + e1 * p1t.z + e2 * p2t.z;
        if (det < 0 && (tScaled >= 0 || tScaled < ray.tMax * det)) return false;
        else if (det > 0 && (tScaled <= 0 || tScaled > ray.tMax * det)) return false;
        Float invDet = 1 / det;
        Float b0 = e0 * invDet, b1 = e1 * invDet, b2 = e2 * invDet, t = tScaled * invDet;
        Float maxZt = std::max(std::abs(p0t.z), std::max(std::abs(p1t.z), std::abs(p2t.z)));
        Float deltaZ = gammaN(3) * maxZt;
        Float maxXt = std::max(std::abs(p0t.x), std::max(std::abs(p1t.x), std::abs(p2t.x)));
        Float maxYt = std::max(std::abs(p0t.y), std::max(std::abs(p1t.y), std::abs(p2t.y)));
        Float deltaX = gammaN(5) * (maxXt + maxZt), deltaY = gammaN(5) * (maxYt + maxZt);
        Float deltaE = 2 * (gammaN(2) * maxXt * maxYt + deltaY * maxXt + deltaX * maxYt);
        Float maxE = std::max(std::abs(e0), std::max(std::abs(e1), std::abs(e2)));
        Float deltaT = 3 * (gammaN(3) * maxE * maxZt + deltaE * maxZt + deltaZ * maxE) * std::abs(invDet);
        if (t <= deltaT) return false;
        if (LengthSquared(Cross(p2 - p0, p1 - p0)) == 0) return false;  // degenerate (Triangle.cpp:186-196)
        h->prim = prim; h->b0 = b0; h->b1 = b1; h->b2 = b2; h->t = t;
        return true;
    }

    // Bounds3::IntersectP(ray, invDir, dirIsNeg), core/Geometry.h:1380-1406
    static bool BoxTest(const gnx_bvh_node &n, const Ray &ray, const Vec &invDir, const int neg[3]) {
        const float *b[2] = {n.bmin, n.bmax};
        Float tMin = (b[neg[0]][0] - ray.o.x) * invDir.x, tMax = (b[1 - neg[0]][0] - ray.o.x) * invDir.x;
        Float tyMin = (b[neg[1]][1] - ray.o.y) * invDir.y, tyMax = (b[1 - neg[1]][1] - ray.o.y) * invDir.y;
        tMax *= 1 + 2 * gammaN(3); tyMax *= 1 + 2 * gammaN(3);
        if (tMin > tyMax || tyMin > tMax) return false;
        if (tyMin > tMin) tMin = tyMin;
        if (tyMax < tMax) tMax = tyMax;
        Float tzMin = (b[neg[2]][2] - ray.o.z) * invDir.z, tzMax = (b[1 - neg[2]][2] - ray.o.z) * invDir.z;
        tzMax *= 1 + 2 * gammaN(3);
        if (tMin > tzMax || tzMin > tMax) return false;
        if (tzMin > tMin) tMin = tzMin;
        if (tzMax < tMax) tMax = tzMax;
        return (tMin < ray.tMax) && (tMax > 0);
    }
    // BVHAccel::Intersect / IntersectP, accelerator/BVHAccel.cpp:653-729
    bool Intersect(const Ray &ray, Hit *hit, bool any, uint64_t *nodes = nullptr, uint64_t *tris = nullptr) const {
        const gnx_geometry &g = d->geom;
        if (g.n_nodes == 0) return false;
        bool found = false;
        Vec invDir(1 / ray.d.x, 1 / ray.d.y, 1 / ray.d.z);
        int neg[3] = {invDir.x < 0, invDir.y < 0, invDir.z < 0};
        int toVisit = 0, cur = 0, stack[64];
        while (true) {
            const gnx_bvh_node &n = g.nodes[cur];
            if (nodes) ++*nodes;
            if (BoxTest(n, ray, invDir, neg)) {
                if (n.n_prims > 0) {
                    for (int i = 0; i < n.n_prims; ++i) {
                        if (tris) ++*tris;
                        Hit h;
                        if (TriangleTest(n.offset + i, ray, &h)) {
                            if (any) return true;
                            found = true; ray.tMax = h.t; *hit = h;
                        }
                    }
                    if (toVisit == 0) break;
                    cur = stack[--toVisit];
                } else if (neg[n.axis]) { stack[toVisit++] = cur + 1; cur = n.offset; }
                else { stack[toVisit++] = n.offset; cur = cur + 1; }
            } else {
                if (toVisit == 0) break;
                cur = stack[--toVisit];
            }
        }
        return found;
    }
};

// ---- surface interaction, BSDF --------------------------------------------------------------------------
struct Surf { Vec p, pError, n, ns, dpdu, dpdv, wo; int prim, material, light; };

Surf MakeSurface(const Scene &sc, const Hit &h, const Vec &rayD) {  // Triangle.cpp:170-303, Interaction.cpp:8-54
    const gnx_geometry &g = sc.d->geom;
    Surf s;
    s.prim = h.prim; s.material = g.prim_material[h.prim]; s.light = g.prim_light ? g.prim_light[h.prim] : -1;
    Vec p0 = sc.P(h.prim, 0), p1 = sc.P(h.prim, 1), p2 = sc.P(h.prim, 2);
    Float uv[3][2] = {{0, 0}, {1, 0}, {1, 1}};
    if (g.prim_uv) memcpy(uv, g.prim_uv + (size_t)h.prim * 6, 24);
    Float duv02[2] = {uv[0][0] - uv[2][0], uv[0][1] - uv[2][1]}, duv12[2] = {uv[1][0] - uv[2][0], uv[1][1] - uv[2][1]};
    Vec dp02 = p0 - p2, dp12 = p1 - p2;
    Float determinant = duv02[0] * duv12[1] - duv02[1] * duv12[0];
    bool degenerateUV = std::abs(determinant) < 1e-8f;
    Vec dpdu, dpdv;
    if (!degenerateUV) {
        Float invdet = 1 / determinant;
        dpdu = (duv12[1] * dp02 - duv02[1] * dp12) * invdet;
        dpdv = (-duv12[0] * dp02 + duv02[0] * dp12) * invdet;
    }
    if (degenerateUV || LengthSquared(Cross(dpdu, dpdv)) == 0) CoordinateSystem(Normalize(Cross(p2 - p0, p1 - p0)), &dpdu, &dpdv);
    s.pError = gammaN(7) * Vec(std::abs(h.b0 * p0.x) + std::abs(h.b1 * p1.x) + std::abs(h.b2 * p2.x),
                               std::abs(h.b0 * p0.y) + std::abs(h.b1 * p1.y) + std::abs(h.b2 * p2.y),
                               std::abs(h.b0 * p0.z) + std::abs(h.b1 * p1.z) + std::abs(h.b2 * p2.z));
    s.p = h.b0 * p0 + h.b1 * p1 + h.b2 * p2;
    s.wo = Normalize(-rayD);
    s.n = Normalize(Cross(dp02, dp12));
    unsigned flags = g.prim_flags ? g.prim_flags[h.prim] : 0;
    if (flags & GNX_PRIM_FLIP_N) s.n = -s.n;
    s.ns = s.n; s.dpdu = dpdu; s.dpdv = dpdv;
    if (g.prim_n && g.prim_has_n && g.prim_has_n[h.prim]) {
        const float *q = g.prim_n + (size_t)h.prim * 9;
        Vec ns = h.b0 * Vec(q[0], q[1], q[2]) + h.b1 * Vec(q[3], q[4], q[5]) + h.b2 * Vec(q[6], q[7], q[8]);
        ns = LengthSquared(ns) > 0 ? Normalize(ns) : s.n;
        Vec ss = Normalize(dpdu), ts = Cross(ss, ns);
        if (LengthSquared(ts) > 0.f) { ts = Normalize(ts); ss = Cross(ts, ns); } else CoordinateSystem(ns, &ss, &ts);
        if (flags & GNX_PRIM_REVERSE_ORI) ts = -ts;
        s.ns = Normalize(Cross(ss, ts)); s.n = Faceforward(s.n, s.ns); s.dpdu = ss; s.dpdv = ts;
    }
    return s;
}

enum { REFL = 1, TRANS = 2, DIFFUSE = 4, GLOSSY = 8, SPECULAR = 16, ALL = 31 };
enum Kind { LAMBERT, OREN, SPECR, FRSPEC, MICRO };
struct Lobe {
    Kind kind; int type;
    Spectrum R, T, eta, k;
    Float A = 0, B = 0, ax = 0, ay = 0, etaI = 1, etaT = 1;
    bool conductor = false;
};
inline Float CosTheta(const Vec &w) { return w.z; }
inline Float AbsCosTheta(const Vec &w) { return std::abs(w.z); }
inline Float Cos2Theta(const Vec &w) { return w.z * w.z; }
inline Float Sin2Theta(const Vec &w) { return std::max((Float)0, (Float)1 - Cos2Theta(w)); }
inline Float SinTheta(const Vec &w) { return std::sqrt(Sin2Theta(w)); }
inline Float TanTheta(const Vec &w) { return SinTheta(w) / CosTheta(w); }
inline Float Tan2Theta(const Vec &w) { return Sin2Theta(w) / Cos2Theta(w); }
inline Float Clampf(Float v, Float lo, Float hi) { return v < lo ? lo : (v > hi ? hi : v); }
inline Float CosPhi(const Vec &w) { Float s = SinTheta(w); return (s == 0) ? 1 : Clampf(w.x / s, -1, 1); }
inline Float SinPhi(const Vec &w) { Float s = SinTheta(w); return (s == 0) ? 0 : Clampf(w.y / s, -1, 1); }
inline bool SameHemisphere(const Vec &w, const Vec &wp) { return w.z * wp.z > 0; }

Float FrDielectric(Float cosThetaI, Float etaI, Float etaT) {  // core/Reflection.cpp:16-38
    cosThetaI = Clampf(cosThetaI, -1, 1);
    if (!(cosThetaI > 0.f)) { std::swap(etaI, etaT); cosThetaI = std::abs(cosThetaI); }
    Float sinThetaI = std::sqrt(std::max((Float)0, 1 - cosThetaI * cosThetaI));
    Float sinThetaT = etaI / etaT * sinThetaI;
    if (sinThetaT >= 1) return 1;
    Float cosThetaT = std::sqrt(std::max((Float)0, 1 - sinThetaT * sinThetaT));
    Float Rparl = ((etaT * cosThetaI) - (etaI * cosThetaT)) / ((etaT * cosThetaI) + (etaI * cosThetaT));
    Float Rperp = ((etaI * cosThetaI) - (etaT * cosThetaT)) / ((etaI * cosThetaI) + (etaT * cosThetaT));
    return (Rparl * Rparl + Rperp * Rperp) / 2;
}
Spectrum FrConductor(Float cosThetaI, const Spectrum &etai, const Spectrum &etat, const Spectrum &k) {  // Reflection.cpp:41-64
    cosThetaI = Clampf(cosThetaI, -1, 1);
    Spectrum eta = etat / etai, etak = k / etai;
    Float cosThetaI2 = cosThetaI * cosThetaI, sinThetaI2 = 1.f - cosThetaI2;
    Spectrum eta2 = eta * eta, etak2 = etak * etak;
    Spectrum t0 = eta2 - etak2 - Spectrum(sinThetaI2);
    Spectrum a2plusb2 = Sqrt(t0 * t0 + 4 * eta2 * etak2);
    Spectrum t1 = a2plusb2 + Spectrum(cosThetaI2);
    Spectrum a = Sqrt(0.5f * (a2plusb2 + t0));
    Spectrum t2 = (Float)2 * cosThetaI * a;
    Spectrum Rs = (t1 - t2) / (t1 + t2);
    Spectrum t3 = cosThetaI2 * a2plusb2 + Spectrum(sinThetaI2 * sinThetaI2);
    Spectrum t4 = t2 * sinThetaI2;
    Spectrum Rp = Rs * (t3 - t4) / (t3 + t4);
    return 0.5f * (Rp + Rs);
}
// TrowbridgeReitzDistribution, core/MicroFacet.cpp:129-136,150-159,215-324
Float TR_D(const Lobe &l, const Vec &wh) {
    Float tan2Theta = Tan2Theta(wh);
    if (std::isinf(tan2Theta)) return 0.;
    const Float cos4Theta = Cos2Theta(wh) * Cos2Theta(wh);
    Float e = (CosPhi(wh) * CosPhi(wh) / (l.ax * l.ax) + SinPhi(wh) * SinPhi(wh) / (l.ay * l.ay)) * tan2Theta;
    return 1 / (Pi * l.ax * l.ay * cos4Theta * (1 + e) * (1 + e));
}
Float TR_Lambda(const Lobe &l, const Vec &w) {
    Float absTanTheta = std::abs(TanTheta(w));
    if (std::isinf(absTanTheta)) return 0.;
    Float alpha = std::sqrt(CosPhi(w) * CosPhi(w) * l.ax * l.ax + SinPhi(w) * SinPhi(w) * l.ay * l.ay);
    Float a2t2 = (alpha * absTanTheta) * (alpha * absTanTheta);
    return (-1 + std::sqrt(1.f + a2t2)) / 2;
}
Float TR_Pdf(const Lobe &l, const Vec &wo, const Vec &wh) { return TR_D(l, wh) * (1 / (1 + TR_Lambda(l, wo))) * AbsDot(wo, wh) / AbsCosTheta(wo); }
void TR_Sample11(Float cosTheta, Float U1, Float U2, Float *slope_x, Float *slope_y) {
    if (cosTheta > .9999) {
        Float r = std::sqrt(U1 / (1 - U1)), phi = 6.28318530718 * U2;
        *slope_x = r * std::cos(phi); *slope_y = r * std::sin(phi);
        return;
    }
    Float sinTheta = std::sqrt(std::max((Float)0, (Float)1 - cosTheta * cosTheta));
    Float tanTheta = sinTheta / cosTheta, a = 1 / tanTheta;
    Float G1 = 2 / (1 + std::sqrt(1.f + 1.f / (a * a)));
    Float A = 2 * U1 / G1 - 1, tmp = 1.f / (A * A - 1.f);
    if (tmp > 1e10) tmp = 1e10;
    Float B = tanTheta;
    Float D = std::sqrt(std::max(Float(B * B * tmp * tmp - (A * A - B * B) * tmp), Float(0)));
    Float s1 = B * tmp - D, s2 = B * tmp + D;
    *slope_x = (A < 0 || s2 > 1.f / tanTheta) ? s1 : s2;
    Float S;
    if (U2 > 0.5f) { S = 1.f; U2 = 2.f * (U2 - .5f); } else { S = -1.f; U2 = 2.f * (.5f - U2); }
    Float z = (U2 * (U2 * (U2 * 0.27385f - 0.73369f) + 0.46341f)) / (U2 * (U2 * (U2 * 0.093073f + 0.309420f) - 1.000000f) + 0.597999f);
    *slope_y = S * z * std::sqrt(1.f + *slope_x * *slope_x);
}
Vec TR_Sample_wh(const Lobe &l, const Vec &wo, const Float u[2]) {
    bool flip = wo.z < 0;
    Vec wi = flip ? -wo : wo;
    Vec wiS = Normalize(Vec(l.ax * wi.x, l.ay * wi.y, wi.z));
    Float sx, sy;
    TR_Sample11(CosTheta(wiS), u[0], u[1], &sx, &sy);
    Float tmp = CosPhi(wiS) * sx - SinPhi(wiS) * sy;
    sy = SinPhi(wiS) * sx + CosPhi(wiS) * sy; sx = tmp;
    sx = l.ax * sx; sy = l.ay * sy;
    Vec wh = Normalize(Vec(-sx, -sy, 1.));
    return flip ? -wh : wh;
}
void ConcentricSampleDisk(const Float u[2], Float *dx, Float *dy) {  // core/Sampling.cpp:87-105
    Float ox = 2.f * u[0] - 1, oy = 2.f * u[1] - 1;
    if (ox == 0 && oy == 0) { *dx = *dy = 0; return; }
    Float theta, r;
    if (std::abs(ox) > std::abs(oy)) { r = ox; theta = PiOver4 * (oy / ox); } else { r = oy; theta = PiOver2 - PiOver4 * (ox / oy); }
    *dx = r * std::cos(theta); *dy = r * std::sin(theta);
}
Vec CosineSampleHemisphere(const Float u[2]) {
    Float dx, dy; ConcentricSampleDisk(u, &dx, &dy);
    return Vec(dx, dy, std::sqrt(std::max((Float)0, 1 - dx * dx - dy * dy)));
}

struct BSDF {  // core/Reflection.h:96-130, core/Reflection.cpp:440-563
    Vec ns, ng, ss, ts; Float eta = 1; int n = 0; Lobe lobes[4];
    Vec ToLocal(const Vec &v) const { return Vec(Dot(v, ss), Dot(v, ts), Dot(v, ns)); }
    Vec ToWorld(const Vec &v) const { return Vec(ss.x * v.x + ts.x * v.y + ns.x * v.z, ss.y * v.x + ts.y * v.y + ns.y * v.z, ss.z * v.x + ts.z * v.y + ns.z * v.z); }
    static bool Match(const Lobe &l, int flags) { return (l.type & flags) == l.type; }
    int NumComponents(int flags) const { int c = 0; for (int i = 0; i < n; ++i) c += Match(lobes[i], flags); return c; }
    static Spectrum Fresnel(const Lobe &l, Float cosI) { return l.conductor ? FrConductor(std::abs(cosI), Spectrum(1.f), l.eta, l.k) : Spectrum(FrDielectric(cosI, l.etaI, l.etaT)); }
    static Spectrum LobeF(const Lobe &l, const Vec &wo, const Vec &wi) {
        switch (l.kind) {
        case LAMBERT: return l.R * InvPi;
        case OREN: {  // Reflection.cpp:173-198
            Float sinThetaI = SinTheta(wi), sinThetaO = SinTheta(wo), maxCos = 0;
            if (sinThetaI > 1e-4 && sinThetaO > 1e-4) maxCos = std::max((Float)0, CosPhi(wi) * CosPhi(wo) + SinPhi(wi) * SinPhi(wo));
            Float sinAlpha, tanBeta;
            if (AbsCosTheta(wi) > AbsCosTheta(wo)) { sinAlpha = sinThetaO; tanBeta = sinThetaI / AbsCosTheta(wi); }
            else { sinAlpha = sinThetaI; tanBeta = sinThetaO / AbsCosTheta(wo); }
            return l.R * InvPi * (l.A + l.B * maxCos * sinAlpha * tanBeta);
        }
        case MICRO: {  // Reflection.cpp:223-240
            Float cosThetaO = AbsCosTheta(wo), cosThetaI = AbsCosTheta(wi);
            Vec wh = wi + wo;
            if (cosThetaI == 0 || cosThetaO == 0) return Spectrum(0.);
            if (wh.x == 0 && wh.y == 0 && wh.z == 0) return Spectrum(0.);
            wh = Normalize(wh);
            Spectrum F = Fresnel(l, Dot(wi, Faceforward(wh, Vec(0, 0, 1))));
            Float G = 1 / (1 + TR_Lambda(l, wo) + TR_Lambda(l, wi));
            return l.R * TR_D(l, wh) * G * F / (4 * cosThetaI * cosThetaO);
        }
        default: return Spectrum(0.f);
        }
    }
    static Float LobePdf(const Lobe &l, const Vec &wo, const Vec &wi) {
        switch (l.kind) {
        case LAMBERT: case OREN: return SameHemisphere(wo, wi) ? AbsCosTheta(wi) * InvPi : 0;
        case MICRO: { if (!SameHemisphere(wo, wi)) return 0; Vec wh = Normalize(wo + wi); return TR_Pdf(l, wo, wh) / (4 * Dot(wo, wh)); }
        default: return 0;
        }
    }
    Spectrum f(const Vec &woW, const Vec &wiW, int flags) const {
        Vec wi = ToLocal(wiW), wo = ToLocal(woW);
        if (wo.z == 0) return Spectrum(0.);
        bool reflect = Dot(wiW, ng) * Dot(woW, ng) > 0;
        Spectrum r(0.f);
        for (int i = 0; i < n; ++i)
            if (Match(lobes[i], flags) && ((reflect && (lobes[i].type & REFL)) || (!reflect && (lobes[i].type & TRANS)))) r += LobeF(lobes[i], wo, wi);
        return r;
    }
    Float Pdf(const Vec &woW, const Vec &wiW, int flags) const {
        if (n == 0) return 0.f;
        Vec wo = ToLocal(woW), wi = ToLocal(wiW);
        if (wo.z == 0) return 0.;
        Float pdf = 0; int m = 0;
        for (int i = 0; i < n; ++i) if (Match(lobes[i], flags)) { ++m; pdf += LobePdf(lobes[i], wo, wi); }
        return m > 0 ? pdf / m : 0.f;
    }
    Spectrum Sample_f(const Vec &woW, Vec *wiW, const Float u[2], Float *pdf, int type, int *sampledType) const {
        int matching = NumComponents(type);
        *pdf = 0; *sampledType = 0;
        if (matching == 0) return Spectrum(0);
        int comp = std::min((int)std::floor(u[0] * matching), matching - 1);
        const Lobe *l = nullptr; int count = comp, chosen = -1;
        for (int i = 0; i < n; ++i) if (Match(lobes[i], type) && count-- == 0) { l = &lobes[i]; chosen = i; break; }
        Float ur[2] = {std::min(u[0] * matching - comp, OneMinusEpsilon), u[1]};
        Vec wi, wo = ToLocal(woW);
        if (wo.z == 0) return Spectrum(0.);
        *sampledType = l->type;
        Spectrum f(0.f);
        switch (l->kind) {
        case SPECR:  // Reflection.cpp:89-97
            wi = Vec(-wo.x, -wo.y, wo.z); *pdf = 1;
            f = (l->etaI == l->etaT && !l->conductor ? Spectrum(1.f) : Fresnel(*l, CosTheta(wi))) * l->R / AbsCosTheta(wi);
            break;
        case FRSPEC: {  // Reflection.cpp:346-380
            Float F = FrDielectric(CosTheta(wo), l->etaI, l->etaT);
            if (ur[0] < F) { wi = Vec(-wo.x, -wo.y, wo.z); *sampledType = SPECULAR | REFL; *pdf = F; f = F * l->R / AbsCosTheta(wi); }
            else {
                bool entering = CosTheta(wo) > 0;
                Float etaI = entering ? l->etaI : l->etaT, etaT = entering ? l->etaT : l->etaI;
                Vec nn = Faceforward(Vec(0, 0, 1), wo);
                Float eta = etaI / etaT, cosThetaI = Dot(nn, wo);  // Refract, Reflection.h:68-80
                Float sin2ThetaI = std::max(Float(0), Float(1 - cosThetaI * cosThetaI)), sin2ThetaT = eta * eta * sin2ThetaI;
                if (sin2ThetaT >= 1) return Spectrum(0);
                Float cosThetaT = std::sqrt(1 - sin2ThetaT);
                wi = eta * -wo + (eta * cosThetaI - cosThetaT) * nn;
                Spectrum ft = l->T * (1 - F);
                ft = ft * ((etaI * etaI) / (etaT * etaT));
                *sampledType = SPECULAR | TRANS; *pdf = 1 - F; f = ft / AbsCosTheta(wi);
            }
            break;
        }
        case MICRO: {  // Reflection.cpp:206-221
            if (wo.z == 0) return Spectrum(0.);
            Vec wh = TR_Sample_wh(*l, wo, ur);
            if (Dot(wo, wh) < 0) return Spectrum(0.);
            wi = -wo + 2 * Dot(wo, wh) * wh;
            if (!SameHemisphere(wo, wi)) return Spectrum(0.f);
            *pdf = TR_Pdf(*l, wo, wh) / (4 * Dot(wo, wh));
            f = LobeF(*l, wo, wi);
            break;
        }
        default:  // BxDF::Sample_f, Reflection.cpp:394-402
            wi = CosineSampleHemisphere(ur);
            if (wo.z < 0) wi.z *= -1;
            *pdf = LobePdf(*l, wo, wi); f = LobeF(*l, wo, wi);
        }
        if (*pdf == 0) { *sampledType = 0; return Spectrum(0); }
        *wiW = ToWorld(wi);
        if (!(l->type & SPECULAR) && matching > 1)
            for (int i = 0; i < n; ++i) if (i != chosen && Match(lobes[i], type)) *pdf += LobePdf(lobes[i], wo, wi);
        if (matching > 1) *pdf /= matching;
        if (!(l->type & SPECULAR)) {
            bool reflect = Dot(*wiW, ng) * Dot(woW, ng) > 0;
            f = Spectrum(0.);
            for (int i = 0; i < n; ++i)
                if (Match(lobes[i], type) && ((reflect && (lobes[i].type & REFL)) || (!reflect && (lobes[i].type & TRANS)))) f += LobeF(lobes[i], wo, wi);
        }
        return f;
    }
};

Float RoughnessToAlpha(Float roughness) {  // core/MicroFacet.h:97-103
    roughness = std::max(roughness, (Float)1e-3);
    Float x = std::log(roughness);
    return 1.62142f + 0.819955f * x + 0.1734f * x * x + 0.0171201f * x * x * x + 0.000640711f * x * x * x * x;
}

// materials/*.cpp ComputeScatteringFunctions; returns false for materials outside the restatement
bool BuildBSDF(const Scene &sc, Surf &s, BSDF *b) {
    const gnx_material &m = sc.d->materials[s.material];
    for (int i = 0; i < GNX_MAT_MAX_RGB; ++i) if (m.rgb_tex[i] >= 0) return false;
    for (int i = 0; i < GNX_MAT_MAX_F; ++i) if (m.f_tex[i] >= 0) return false;
    if (m.flags & GNX_MATF_BUMP_IDENTITY) { s.ns = Normalize(Cross(s.dpdu, s.dpdv)); s.ns = Faceforward(s.ns, s.n); }  // Material.cpp:45-51
    b->ns = s.ns; b->ng = s.n; b->ss = Normalize(s.dpdu); b->ts = Cross(b->ns, b->ss);
    auto rgb = [&](int i) { return Spectrum(m.rgb[i][0], m.rgb[i][1], m.rgb[i][2]); };
    auto micro = [&](const Spectrum &R, Float ur, Float vr) { Lobe l; l.kind = MICRO; l.type = REFL | GLOSSY; l.R = R; l.ax = std::max(Float(0.001), ur); l.ay = std::max(Float(0.001), vr); return l; };
    switch (m.type) {
    case GNX_MAT_MATTE: {
        Spectrum r = rgb(0).Clamp0(); Float sig = Clampf(m.f[0], 0, 90);
        if (!r.IsBlack()) {
            Lobe l; l.type = REFL | DIFFUSE; l.R = r;
            if (sig == 0) l.kind = LAMBERT;
            else { l.kind = OREN; Float sg = (Pi / 180) * sig, s2 = sg * sg; l.A = 1.f - (s2 / (2.f * (s2 + 0.33f))); l.B = 0.45f * s2 / (s2 + 0.09f); }
            b->lobes[b->n++] = l;
        }
        return true;
    }
    case GNX_MAT_MIRROR: {
        Spectrum R = rgb(0).Clamp0();
        if (!R.IsBlack()) { Lobe l; l.kind = SPECR; l.type = REFL | SPECULAR; l.R = R; b->lobes[b->n++] = l; }  // FresnelNoOp: etaI == etaT
        return true;
    }
    case GNX_MAT_PLASTIC: {
        Spectrum kd = rgb(0).Clamp0(), ks = rgb(1).Clamp0();
        if (!kd.IsBlack()) { Lobe l; l.kind = LAMBERT; l.type = REFL | DIFFUSE; l.R = kd; b->lobes[b->n++] = l; }
        if (!ks.IsBlack()) {
            Float rough = m.f[0];
            if (m.flags & GNX_MATF_REMAP_ROUGHNESS) rough = RoughnessToAlpha(rough);
            Lobe l = micro(ks, rough, rough); l.etaI = 1.5f; l.etaT = 1.f;
            b->lobes[b->n++] = l;
        }
        return true;
    }
    case GNX_MAT_METAL: {
        Float ur = m.f[0], vr = m.f[1];
        if (m.flags & GNX_MATF_REMAP_ROUGHNESS) { ur = RoughnessToAlpha(ur); vr = RoughnessToAlpha(vr); }
        Lobe l = micro(Spectrum(1.f), ur, vr); l.conductor = true; l.eta = rgb(0); l.k = rgb(1);
        b->lobes[b->n++] = l;
        return true;
    }
    case GNX_MAT_GLASS: {
        Spectrum R = rgb(0).Clamp0(), T = rgb(1).Clamp0();
        b->eta = m.f[2];
        if (R.IsBlack() && T.IsBlack()) return true;
        if (!(m.f[0] == 0 && m.f[1] == 0)) return false;  // rough glass: not restated
        Lobe l; l.kind = FRSPEC; l.type = REFL | TRANS | SPECULAR; l.R = R; l.T = T; l.etaI = 1.f; l.etaT = m.f[2];
        b->lobes[b->n++] = l;
        return true;
    }
    default: return false;
    }
}

// ---- lights ----------------------------------------------------------------------------------------------------
Spectrum EnvLookup(const gnx_envmap &e, Float su, Float sv) {  // MIPMap::triangle(0, st), core/MIPMap.h:245-256
    Float s = su * e.width - 0.5f, t = sv * e.height - 0.5f;
    int s0 = std::floor(s), t0 = std::floor(t);
    Float ds = s - s0, dt = t - t0;
    auto tx = [&](int x, int y) {
        x %= e.width; if (x < 0) x += e.width;
        y %= e.height; if (y < 0) y += e.height;
        const float *q = e.texels + ((size_t)y * e.width + x) * 3;
        return Spectrum(q[0], q[1], q[2]);
    };
    return (1 - ds) * (1 - dt) * tx(s0, t0) + (1 - ds) * dt * tx(s0, t0 + 1) + ds * (1 - dt) * tx(s0 + 1, t0) + ds * dt * tx(s0 + 1, t0 + 1);
}
Vec XformVec(const float m[16], const Vec &v) { return Vec(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z, m[8] * v.x + m[9] * v.y + m[10] * v.z); }
Spectrum EnvLe(const gnx_envmap &e, const Vec &d) {  // InfiniteAreaLight::Le, lights/InfiniteAreaLight.cpp:91-96
    Vec w = Normalize(XformVec(e.world_to_light, d));
    Float phi = std::atan2(w.y, w.x); if (phi < 0) phi += 2 * Pi;
    return EnvLookup(e, phi * Inv2Pi, std::acos(Clampf(w.z, -1, 1)) * InvPi);
}
int FindInterval(const Float *cdf, int size, Float u) {  // GNXRayTracer.h:336-349
    int first = 0, len = size;
    while (len > 0) { int half = len >> 1, middle = first + half; if (cdf[middle] <= u) { first = middle + 1; len -= half + 1; } else len = half; }
    return std::max(0, std::min(first - 1, size - 2));
}
Float SampleContinuous(const Float *func, const Float *cdf, Float funcInt, int n, Float u, Float *pdf, int *off) {  // Sampling.h:40-59
    int offset = FindInterval(cdf, n + 1, u);
    *off = offset;
    Float du = u - cdf[offset];
    if ((cdf[offset + 1] - cdf[offset]) > 0) du /= (cdf[offset + 1] - cdf[offset]);
    *pdf = (funcInt > 0) ? func[offset] / funcInt : 0;
    return (offset + du) / n;
}
struct LightSample { Vec wi, pl, nl, plError; Spectrum Li; Float pdf = 0; bool env = false; };
Spectrum AreaL(const gnx_light &l, const Vec &n, const Vec &w) {  // DiffuseAreaLight::L's bool truncation, lights/DiffuseAreaLight.h:22-27
    bool dotNW = Dot(n, w);
    return (l.two_sided || dotNW > 0) ? Spectrum(l.L[0], l.L[1], l.L[2]) : Spectrum(0.f);
}
bool SampleLi(const Scene &sc, const gnx_light &l, const Vec &refP, const Float u[2], LightSample *ls) {
    if (l.type == GNX_LIGHT_INFINITE) {  // InfiniteAreaLight::Sample_Li, InfiniteAreaLight.cpp:98-121
        const gnx_envmap &e = sc.d->env;
        Float pdf0, pdf1; int v, dummy;
        Float d1 = SampleContinuous(e.marg_func, e.marg_cdf, e.marg_int, e.dist_h, u[1], &pdf1, &v);
        Float d0 = SampleContinuous(e.cond_func + (size_t)v * e.dist_w, e.cond_cdf + (size_t)v * (e.dist_w + 1), e.cond_int[v], e.dist_w, u[0], &pdf0, &dummy);
        Float mapPdf = pdf0 * pdf1;
        if (mapPdf == 0) return false;
        Float theta = d1 * Pi, phi = d0 * 2 * Pi;
        Float cosTheta = std::cos(theta), sinTheta = std::sin(theta), sinPhi = std::sin(phi), cosPhi = std::cos(phi);
        ls->wi = XformVec(e.light_to_world, Vec(sinTheta * cosPhi, sinTheta * sinPhi, cosTheta));
        ls->pdf = mapPdf / (2 * Pi * Pi * sinTheta);
        if (sinTheta == 0) ls->pdf = 0;
        ls->Li = EnvLookup(e, d0, d1); ls->env = true;
        return true;
    }
    // DiffuseAreaLight::Sample_Li -> Shape::Sample(ref, u) -> Triangle::Sample(u)
    Vec p0 = sc.P(l.prim, 0), p1 = sc.P(l.prim, 1), p2 = sc.P(l.prim, 2);
    Float su0 = std::sqrt(u[0]), b0 = 1 - su0, b1 = u[1] * su0;
    Vec p = b0 * p0 + b1 * p1 + (1 - b0 - b1) * p2;
    Vec n = Normalize(Cross(p1 - p0, p2 - p0));
    const gnx_geometry &g = sc.d->geom;
    if (g.prim_n && g.prim_has_n && g.prim_has_n[l.prim]) {
        const float *q = g.prim_n + (size_t)l.prim * 9;
        n = Faceforward(n, b0 * Vec(q[0], q[1], q[2]) + b1 * Vec(q[3], q[4], q[5]) + (1 - b0 - b1) * Vec(q[6], q[7], q[8]));
    } else if (g.prim_flags && (g.prim_flags[l.prim] & GNX_PRIM_FLIP_N)) n = -n;
    ls->plError = gammaN(6) * (Abs(b0 * p0) + Abs(b1 * p1) + Abs((1 - b0 - b1) * p2));
    Float pdf = 1 / l.area;
    Vec wi = p - refP;
    if (LengthSquared(wi) == 0) pdf = 0;
    else { wi = Normalize(wi); pdf *= LengthSquared(refP - p) / AbsDot(n, -wi); if (std::isinf(pdf)) pdf = 0.f; }
    ls->pl = p; ls->nl = n;
    if (pdf == 0 || LengthSquared(p - refP) == 0) return false;
    ls->wi = Normalize(p - refP); ls->pdf = pdf; ls->Li = AreaL(l, n, -ls->wi);
    return true;
}

// LightDistribution::Lookup: uniform, or SpatialLightDistribution with lazily computed voxels
// (core/LightDistribution.cpp:109-274).  Returns pointers to func[n], cdf[n+1] and funcInt.
void LookupDistribution(Scene &sc, const Vec &p, int strategy, const Float **func, const Float **cdf, Float *funcInt) {
    const int n = sc.d->n_lights;
    if (strategy == GNX_LIGHTS_UNIFORM || n == 1) { *func = sc.uniformTab.data(); *cdf = sc.uniformTab.data() + n; *funcInt = sc.uniformTab[2 * n + 1]; return; }
    // PowerLightDistribution::Lookup (core/LightDistribution.cpp:44-50)
    if (strategy == GNX_LIGHTS_POWER) { *func = sc.powerTab.data(); *cdf = sc.powerTab.data() + n; *funcInt = sc.powerTab[2 * n + 1]; return; }
    const float *wb = sc.d->geom.world_bound;
    int pi[3];
    for (int i = 0; i < 3; ++i) {
        Float o = p[i] - wb[i];
        if (wb[3 + i] > wb[i]) o /= wb[3 + i] - wb[i];
        pi[i] = std::max(0, std::min((int)(o * sc.nVoxels[i]), sc.nVoxels[i] - 1));
    }
    uint64_t key = (uint64_t(pi[0]) << 40) | (uint64_t(pi[1]) << 20) | pi[2];
    std::lock_guard<std::mutex> lock(sc.voxMutex);
    auto it = sc.voxels.find(key);
    if (it == sc.voxels.end()) {  // ComputeDistribution, LightDistribution.cpp:206-274
        std::vector<Float> tab(2 * n + 2, 0.f);
        Vec lo, hi;
        Float *plo = &lo.x, *phi = &hi.x;
        for (int a = 0; a < 3; ++a) {
            Float t0 = Float(pi[a]) / Float(sc.nVoxels[a]), t1 = Float(pi[a] + 1) / Float(sc.nVoxels[a]);
            plo[a] = (1 - t0) * wb[a] + t0 * wb[3 + a]; phi[a] = (1 - t1) * wb[a] + t1 * wb[3 + a];
        }
        for (int i = 0; i < 128; ++i) {
            Float r[5];
            for (int k = 0; k < 5; ++k) r[k] = sc.halton.RadicalInverse(k, i);
            Vec po((1 - r[0]) * lo.x + r[0] * hi.x, (1 - r[1]) * lo.y + r[1] * hi.y, (1 - r[2]) * lo.z + r[2] * hi.z);
            Float u[2] = {r[3], r[4]};
            for (int j = 0; j < n; ++j) { LightSample ls; if (SampleLi(sc, sc.d->lights[j], po, u, &ls) && ls.pdf > 0) tab[j] += ls.Li.y() / ls.pdf; }
        }
        Float sum = 0; for (int j = 0; j < n; ++j) sum += tab[j];
        Float avg = sum / (128 * n), minC = (avg > 0) ? .001 * avg : 1;
        for (int j = 0; j < n; ++j) tab[j] = std::max(tab[j], minC);
        Float *c = tab.data() + n;
        c[0] = 0;
        for (int j = 1; j < n + 1; ++j) c[j] = c[j - 1] + tab[j - 1] / n;
        Float fi = c[n];
        for (int j = 1; j < n + 1; ++j) c[j] = fi == 0 ? Float(j) / Float(n) : c[j] / fi;
        tab[2 * n + 1] = fi;
        it = sc.voxels.emplace(key, std::move(tab)).first;
    }
    *func = it->second.data(); *cdf = it->second.data() + n; *funcInt = it->second[2 * n + 1];
}

// ---- camera (camera/Perspective.cpp:35-60, core/Transform.h:198-282) ----------------------------------------
Ray CameraRay(const Scene &sc, int px, int py, int64_t index) {
    const gnx_camera &c = sc.d->camera;
    Float fx = (Float)px + sc.halton.SampleDimension(index, 0), fy = (Float)py + sc.halton.SampleDimension(index, 1);
    const float *m = c.raster_to_camera;
    Float xp = m[0] * fx + m[1] * fy + m[2] * 0 + m[3], yp = m[4] * fx + m[5] * fy + m[6] * 0 + m[7];
    Float zp = m[8] * fx + m[9] * fy + m[10] * 0 + m[11], wp = m[12] * fx + m[13] * fy + m[14] * 0 + m[15];
    Vec pCamera(xp, yp, zp);
    if (wp != 1) { Float inv = (Float)1 / wp; pCamera = Vec(inv * xp, inv * yp, inv * zp); }
    Vec ro(0, 0, 0), rd = Normalize(pCamera);
    if (c.lens_radius > 0) {
        Float ul[2] = {sc.halton.SampleDimension(index, 3), sc.halton.SampleDimension(index, 4)}, lx, ly;
        ConcentricSampleDisk(ul, &lx, &ly);
        lx *= c.lens_radius; ly *= c.lens_radius;
        Float ft = c.focal_distance / rd.z;
        Vec pFocus = ro + rd * ft;
        ro = Vec(lx, ly, 0); rd = Normalize(pFocus - ro);
    }
    const float *w = c.camera_to_world;
    Ray r;
    Float ox = (w[0] * ro.x + w[1] * ro.y) + (w[2] * ro.z + w[3]), oy = (w[4] * ro.x + w[5] * ro.y) + (w[6] * ro.z + w[7]);
    Float oz = (w[8] * ro.x + w[9] * ro.y) + (w[10] * ro.z + w[11]);
    Vec oErr = gammaN(3) * Vec(std::abs(w[0] * ro.x) + std::abs(w[1] * ro.y) + std::abs(w[2] * ro.z) + std::abs(w[3]),
                               std::abs(w[4] * ro.x) + std::abs(w[5] * ro.y) + std::abs(w[6] * ro.z) + std::abs(w[7]),
                               std::abs(w[8] * ro.x) + std::abs(w[9] * ro.y) + std::abs(w[10] * ro.z) + std::abs(w[11]));
    r.o = Vec(ox, oy, oz);
    r.d = XformVec(w, rd);
    Float lsq = LengthSquared(r.d);
    if (lsq > 0) { Float dt = Dot(Abs(r.d), oErr) / lsq; r.o = r.o + r.d * dt; r.tMax -= dt; }
    return r;
}

// ---- PathIntegrator::Li (integrators/PathIntegrator.cpp:62-208) with EstimateDirect (core/Integrator.cpp:93-210)
struct Counters { uint64_t nodes = 0, tris = 0, extend = 0, shadow = 0, mis = 0; };

int Li(Scene &sc, const gnx_render_params &prm, Ray ray, Sampler &smp, Spectrum *Lout, Counters &cn) {
    Spectrum L(0.f), beta(1.f);
    bool specularBounce = false;
    Float etaScale = 1;
    const int kNonSpec = ALL & ~SPECULAR;
    for (int bounces = 0;; ++bounces) {
        Hit hit;
        ++cn.extend;
        bool found = sc.Intersect(ray, &hit, false, &cn.nodes, &cn.tris);
        Surf s;
        if (found) s = MakeSurface(sc, hit, ray.d);
        if (bounces == 0 || specularBounce) {
            if (found) { if (s.light >= 0) L += beta * AreaL(sc.d->lights[s.light], s.n, -ray.d); }
            else if (sc.d->env.present) L += beta * EnvLe(sc.d->env, ray.d);
        }
        if (!found || bounces >= prm.max_depth) break;
        if (s.material < 0) { ray.o = OffsetRayOrigin(s.p, s.pError, s.n, ray.d); ray.tMax = Infinity; bounces--; continue; }
        BSDF bsdf;
        if (!BuildBSDF(sc, s, &bsdf)) return GNX_ERR_UNSUPPORTED;
        if (bsdf.NumComponents(kNonSpec) > 0 && sc.d->n_lights > 0) {
            // UniformSampleOneLight, core/Integrator.cpp:57-79
            const Float *func, *cdf; Float funcInt;
            LookupDistribution(sc, s.p, prm.light_strategy, &func, &cdf, &funcInt);
            const int nL = sc.d->n_lights;
            int lightNum = FindInterval(cdf, nL + 1, smp.Get1D());
            Float lightPdfSel = (funcInt > 0) ? func[lightNum] / (funcInt * nL) : 0;
            if (lightPdfSel != 0) {
                Float uLight[2], uScat[2];
                smp.Get2D(uLight); smp.Get2D(uScat);
                const gnx_light &light = sc.d->lights[lightNum];
                Spectrum Ld(0.f);
                LightSample ls;
                Float lightPdf = 0, scatteringPdf = 0;
                if (SampleLi(sc, light, s.p, uLight, &ls)) lightPdf = ls.pdf;
                if (lightPdf > 0 && !ls.Li.IsBlack()) {
                    Spectrum f = bsdf.f(s.wo, ls.wi, kNonSpec) * AbsDot(ls.wi, bsdf.ns);
                    scatteringPdf = bsdf.Pdf(s.wo, ls.wi, kNonSpec);
                    if (!f.IsBlack()) {
                        Ray sh;  // VisibilityTester / Interaction::SpawnRayTo, core/Interaction.h:39-53
                        if (ls.env) { Vec p1 = s.p + ls.wi * (2 * sc.d->env.world_radius); sh.o = OffsetRayOrigin(s.p, s.pError, s.n, p1 - s.p); sh.d = p1 - sh.o; }
                        else { sh.o = OffsetRayOrigin(s.p, s.pError, s.n, ls.pl - s.p); Vec target = OffsetRayOrigin(ls.pl, ls.plError, ls.nl, sh.o - ls.pl); sh.d = target - sh.o; }
                        sh.tMax = 1 - ShadowEpsilon;
                        Hit dummy;
                        ++cn.shadow;
                        if (!sc.Intersect(sh, &dummy, true, &cn.nodes, &cn.tris)) {
                            Float weight = (lightPdf * lightPdf) / (lightPdf * lightPdf + scatteringPdf * scatteringPdf);
                            Ld += f * ls.Li * weight / lightPdf;
                        }
                    }
                }
                {   // BSDF-sampling half (lights here are never delta lights)
                    Vec wi; int sampledType;
                    Spectrum f = bsdf.Sample_f(s.wo, &wi, uScat, &scatteringPdf, kNonSpec, &sampledType);
                    f = f * AbsDot(wi, bsdf.ns);
                    bool done = false;
                    if (!f.IsBlack() && scatteringPdf > 0) {
                        Ray pr; pr.o = OffsetRayOrigin(s.p, s.pError, s.n, wi); pr.d = wi;
                        Float lp;
                        if (light.type == GNX_LIGHT_INFINITE) {  // InfiniteAreaLight::Pdf_Li, InfiniteAreaLight.cpp:123-132
                            const gnx_envmap &e = sc.d->env;
                            Vec w = XformVec(e.world_to_light, wi);
                            Float theta = std::acos(Clampf(w.z, -1, 1)), phi = std::atan2(w.y, w.x); if (phi < 0) phi += 2 * Pi;
                            Float sinTheta = std::sin(theta);
                            if (sinTheta == 0) lp = 0;
                            else {
                                int iu = std::max(0, std::min(int(phi * Inv2Pi * e.dist_w), e.dist_w - 1)), iv = std::max(0, std::min(int(theta * InvPi * e.dist_h), e.dist_h - 1));
                                lp = (e.cond_func[(size_t)iv * e.dist_w + iu] / e.marg_int) / (2 * Pi * Pi * sinTheta);
                            }
                        } else {  // Shape::Pdf(ref, wi): hit the light's own triangle, core/Shape.cpp:37-53
                            Hit lh; Ray own = pr;
                            if (!sc.TriangleTest(light.prim, own, &lh)) lp = 0;
                            else {
                                Vec p0 = sc.P(light.prim, 0), p1 = sc.P(light.prim, 1), p2 = sc.P(light.prim, 2);
                                Vec pl = lh.b0 * p0 + lh.b1 * p1 + lh.b2 * p2, nl = Normalize(Cross(p0 - p2, p1 - p2));
                                lp = LengthSquared(s.p - pl) / (AbsDot(nl, -wi) * light.area);
                                if (std::isinf(lp)) lp = 0.f;
                            }
                        }
                        if (lp == 0) done = true;
                        if (!done) {
                            Float weight = (scatteringPdf * scatteringPdf) / (scatteringPdf * scatteringPdf + lp * lp);
                            Hit lh;
                            ++cn.mis;
                            bool fh = sc.Intersect(pr, &lh, false, &cn.nodes, &cn.tris);
                            Spectrum Lr(0.f);
                            if (fh) {
                                if (light.type == GNX_LIGHT_AREA_TRI && lh.prim == light.prim) {
                                    Surf ls2 = MakeSurface(sc, lh, wi);
                                    Lr = AreaL(light, ls2.n, -wi);
                                }
                            } else if (light.type == GNX_LIGHT_INFINITE) Lr = EnvLe(sc.d->env, wi);
                            if (!Lr.IsBlack()) Ld += f * Lr * weight / scatteringPdf;
                        }
                    }
                }
                L += beta * (Ld / lightPdfSel);
            }
        }
        Vec wo = -ray.d, wi;
        Float pdf, u[2]; int flags;
        smp.Get2D(u);
        Spectrum f = bsdf.Sample_f(wo, &wi, u, &pdf, ALL, &flags);
        if (f.IsBlack() || pdf == 0.f) break;
        beta *= f * AbsDot(wi, bsdf.ns) / pdf;
        specularBounce = (flags & SPECULAR) != 0;
        if ((flags & SPECULAR) && (flags & TRANS)) { Float eta = bsdf.eta; etaScale *= (Dot(wo, s.n) > 0) ? (eta * eta) : 1 / (eta * eta); }
        ray.o = OffsetRayOrigin(s.p, s.pError, s.n, wi); ray.d = wi; ray.tMax = Infinity;
        Spectrum rrBeta = beta * etaScale;
        if (rrBeta.MaxComponentValue() < prm.rr_threshold && bounces > 3) {
            Float q = std::max((Float).05, 1 - rrBeta.MaxComponentValue());
            if (smp.Get1D() < q) break;
            beta = beta / (1 - q);
        }
    }
    *Lout = L;
    return 0;
}

struct Handle { Scene sc; };

}  // namespace

extern "C" {

void *gnxr_create(const gnx_scene_desc *d) {
    auto *h = new Handle;
    Scene &sc = h->sc;
    sc.d = d;
    sc.halton.Init(d->sampler);
    const int n = d->n_lights;
    sc.uniformTab.assign(2 * n + 2, 0.f);  // UniformLightDistribution, LightDistribution.cpp:35-38
    for (int i = 0; i < n; ++i) sc.uniformTab[i] = 1;
    Float *c = sc.uniformTab.data() + n;
    for (int i = 1; i < n + 1; ++i) c[i] = c[i - 1] + Float(1) / n;
    Float fi = n ? c[n] : 0;
    for (int i = 1; i < n + 1; ++i) c[i] /= fi;
    if (n) sc.uniformTab[2 * n + 1] = fi;
    // ComputeLightPowerDistribution (core/Integrator.cpp:212-220): Distribution1D over Light::Power().y(),
    // taken from the description's light_power table (the restatement does not re-derive Light::Power)
    sc.powerTab.assign(2 * n + 2, 0.f);
    if (n && d->light_power) {
        Float *pc = sc.powerTab.data() + n;
        for (int i = 0; i < n; ++i) sc.powerTab[i] = d->light_power[i];
        for (int i = 1; i < n + 1; ++i) pc[i] = pc[i - 1] + sc.powerTab[i - 1] / n;
        Float pfi = pc[n];
        if (pfi == 0) for (int i = 1; i < n + 1; ++i) pc[i] = Float(i) / Float(n);
        else for (int i = 1; i < n + 1; ++i) pc[i] /= pfi;
        sc.powerTab[2 * n + 1] = pfi;
    }
    const float *wb = d->geom.world_bound;  // SpatialLightDistribution ctor, LightDistribution.cpp:70-87
    Float diag[3] = {wb[3] - wb[0], wb[4] - wb[1], wb[5] - wb[2]};
    int mx = (diag[0] > diag[1] && diag[0] > diag[2]) ? 0 : (diag[1] > diag[2] ? 1 : 2);
    for (int i = 0; i < 3; ++i) sc.nVoxels[i] = std::max(1, int(std::round(diag[i] / diag[mx] * 64)));
    return h;
}
void gnxr_destroy(void *h) { delete (Handle *)h; }

// SamplerIntegrator::Render's loops and box film (core/Integrator.cpp:256-311).  counts: 5 uint64 or NULL.
int gnxr_render(void *hh, const gnx_render_params *p, float *rgba, uint64_t *counts) {
    Scene &sc = ((Handle *)hh)->sc;
    int rc = 0;
    Counters total;
#pragma omp parallel
    {
        Counters cn;
#pragma omp for schedule(dynamic, 16)
        for (int pixel = 0; pixel < p->width * p->height; ++pixel) {
            int px = pixel % p->width, py = pixel / p->width;
            Spectrum col(0.f);
            for (int s = 0; s < p->spp; ++s) {
                Sampler smp{&sc.halton, sc.halton.IndexForSample(px, py, p->first_sample + s), 5};
                Spectrum L;
                int r = Li(sc, *p, CameraRay(sc, px, py, smp.index), smp, &L, cn);
                if (r) { rc = r; break; }
                col += L;
            }
            Float norm = (Float)(p->spp_normalize > 0 ? p->spp_normalize : p->spp);
            col = col / norm;
            for (int c = 0; c < 3; ++c) rgba[4 * pixel + c] = col.c[c];
            rgba[4 * pixel + 3] = 1.f;
        }
#pragma omp critical
        { total.nodes += cn.nodes; total.tris += cn.tris; total.extend += cn.extend; total.shadow += cn.shadow; total.mis += cn.mis; }
    }
    if (counts) { counts[0] = total.extend; counts[1] = total.shadow; counts[2] = total.mis; counts[3] = total.nodes; counts[4] = total.tris; }
    return rc;
}

int gnxr_samples(void *hh, const gnx_render_params *p, int n, const int *px, const int *py, const int *sample, float *rgb) {
    Scene &sc = ((Handle *)hh)->sc;
    int rc = 0;
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < n; ++i) {
        Counters cn;
        Sampler smp{&sc.halton, sc.halton.IndexForSample(px[i], py[i], sample[i]), 5};
        Spectrum L;
        int r = Li(sc, *p, CameraRay(sc, px[i], py[i], smp.index), smp, &L, cn);
        if (r) rc = r;
        for (int c = 0; c < 3; ++c) rgb[3 * i + c] = L.c[c];
    }
    return rc;
}

int gnxr_primary_hits(void *hh, int width, int height, int sample, int *out) {
    Scene &sc = ((Handle *)hh)->sc;
#pragma omp parallel for schedule(dynamic, 256)
    for (int pixel = 0; pixel < width * height; ++pixel) {
        int px = pixel % width, py = pixel / width;
        Ray r = CameraRay(sc, px, py, sc.halton.IndexForSample(px, py, sample));
        Hit h;
        out[pixel] = sc.Intersect(r, &h, false) ? (sc.d->geom.prim_id ? sc.d->geom.prim_id[h.prim] : h.prim) : -1;
    }
    return 0;
}

int gnxr_sample_dims(void *hh, int n, const int64_t *index, const int *dim, float *out) {
    Scene &sc = ((Handle *)hh)->sc;
    for (int i = 0; i < n; ++i) out[i] = sc.halton.SampleDimension(index[i], dim[i]);
    return 0;
}
int64_t gnxr_sample_index(void *hh, int px, int py, int sample) { return ((Handle *)hh)->sc.halton.IndexForSample(px, py, sample); }

}  // extern "C"
