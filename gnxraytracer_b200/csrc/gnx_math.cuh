// gnx_math.cuh — fp32 vector math for the wavefront kernels.
//
// The arithmetic follows the association order of the reference's Vector3/Point3/Normal3 operators
// (core/Geometry.h) so that discrete decisions (which triangle is hit, which lobe/light is
// chosen) agree with the CPU oracle on all but grazing cases: Dot is ((x*x)+(y*y))+(z*z)
// (Geometry.h:905), Normalize multiplies by the reciprocal length (Geometry.h:206-210,952),
// Cross is evaluated in double (Geometry.h:925-931).  The library is compiled with -fmad=false so
// nvcc does not contract a*b+c into FMA, which x86 -O2 does not do either.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <math.h>

namespace gnx {

#ifdef __CUDACC__
#define GNX_HD __host__ __device__ __forceinline__
#define GNX_D __host__ __device__ __forceinline__
#else
// host compilers (the scene kit, the CPU emulation under tests/): plain inline — with always_inline everywhere g++ -O2
// took up to 15 minutes over the emulation's single translation unit
#define GNX_HD inline
#define GNX_D inline
#endif
// One out-of-line copy for the large switch-over-lobe-kind functions: the shade kernel calls each of them from
// several places, and fully inlined it grew to 54 K SASS instructions (870 KB) against a 32 KB L1.5 / ~128 KB
// instruction cache — ncu showed 46 % icache hit rate and stall_no_instruction as the top stall.
#ifdef __CUDACC__
#define GNX_NOINLINE static __host__ __device__ __noinline__
#else
#define GNX_NOINLINE static inline
#endif
#ifndef GNX_LOBE_NOINLINE
#define GNX_LOBE_NOINLINE 0
#endif
#if GNX_LOBE_NOINLINE
#define GNX_LOBE_FN GNX_NOINLINE
#else
#define GNX_LOBE_FN GNX_D
#endif

constexpr float kPi = 3.14159265358979323846f;
constexpr float kInvPi = 0.31830988618379067154f;
constexpr float kInv2Pi = 0.15915494309189533577f;
constexpr float kInv4Pi = 0.07957747154594766788f;
constexpr float kPiOver2 = 1.57079632679489661923f;
constexpr float kPiOver4 = 0.78539816339744830961f;
constexpr float kMachineEpsilon = 5.9604644775390625e-08f;  // 2^-24, core/GNXRayTracer.h:140
constexpr float kShadowEpsilon = 0.0001f;                   // core/GNXRayTracer.h:141
constexpr float kOneMinusEpsilon = 0.99999994f;             // core/RNG.h:14
#define GNX_INF (__builtin_huge_valf())

// Bit casts and read-only loads that also compile for the host: every per-path function below is
// __host__ __device__ so that tests/emul can run the very same code on the CPU against the
// reference (the product never executes the host instantiation).
GNX_HD float i2f(int v) {
#ifdef __CUDA_ARCH__
    return __int_as_float(v);
#else
    float f; memcpy(&f, &v, 4); return f;
#endif
}
GNX_HD int f2i(float v) {
#ifdef __CUDA_ARCH__
    return __float_as_int(v);
#else
    int i; memcpy(&i, &v, 4); return i;
#endif
}
GNX_HD uint32_t f2u(float v) { return (uint32_t)f2i(v); }
GNX_HD float u2f(uint32_t v) { return i2f((int)v); }
template <typename T>
GNX_HD T ldg(const T *p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}
// 32-byte read-only load (LDG.E.256 on sm_100a; the pointer must be 32-byte aligned).  A traversal step
// gathers a 64-byte node per lane from unrelated addresses, so the L1 wavefront count — not bytes — is what
// the load path pays for: two 256-bit loads cost half the wavefronts of four 128-bit ones.
GNX_HD void ldg256(const float4 *p, float4 *a, float4 *b) {
#ifdef __CUDA_ARCH__
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a->x), "=f"(a->y), "=f"(a->z), "=f"(a->w), "=f"(b->x), "=f"(b->y), "=f"(b->z), "=f"(b->w)
                 : "l"(p));
#else
    *a = p[0]; *b = p[1];
#endif
}
GNX_HD uint32_t brev32(uint32_t n) {
#ifdef __CUDA_ARCH__
    return __brev(n);
#else
    uint32_t r = 0;
    for (int i = 0; i < 32; ++i) { r = (r << 1) | (n & 1); n >>= 1; }
    return r;
#endif
}
GNX_HD uint64_t brev64(uint64_t n) {
#ifdef __CUDA_ARCH__
    return __brevll(n);
#else
    uint64_t r = 0;
    for (int i = 0; i < 64; ++i) { r = (r << 1) | (n & 1); n >>= 1; }
    return r;
#endif
}
GNX_HD bool finf(float v) { return v == __builtin_huge_valf() || v == -__builtin_huge_valf(); }

GNX_HD constexpr float gamma_n(int n) { return (n * kMachineEpsilon) / (1 - n * kMachineEpsilon); }

struct V3 {
    float x, y, z;
    GNX_HD V3() : x(0), y(0), z(0) {}
    GNX_HD V3(float a, float b, float c) : x(a), y(b), z(c) {}
    GNX_HD explicit V3(float a) : x(a), y(a), z(a) {}
    GNX_HD float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
GNX_HD V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
GNX_HD V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
GNX_HD V3 operator-(V3 a) { return V3(-a.x, -a.y, -a.z); }
GNX_HD V3 operator*(V3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
GNX_HD V3 operator*(float s, V3 a) { return V3(a.x * s, a.y * s, a.z * s); }
GNX_HD V3 operator*(V3 a, V3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
GNX_HD V3 operator/(V3 a, V3 b) { return V3(a.x / b.x, a.y / b.y, a.z / b.z); }
// Vector3::operator/ multiplies by the reciprocal (core/Geometry.h:206-210)
GNX_HD V3 div_recip(V3 a, float f) { float inv = 1.0f / f; return V3(a.x * inv, a.y * inv, a.z * inv); }
// RGBSpectrum::operator/(Float) divides each sample (core/Spectrum.h CoefficientSpectrum)
GNX_HD V3 div_each(V3 a, float f) { return V3(a.x / f, a.y / f, a.z / f); }
GNX_HD V3 &operator+=(V3 &a, V3 b) { a = a + b; return a; }
GNX_HD V3 &operator*=(V3 &a, V3 b) { a = a * b; return a; }
GNX_HD V3 &operator*=(V3 &a, float s) { a = a * s; return a; }
GNX_HD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
GNX_HD float absdot(V3 a, V3 b) { return fabsf(dot(a, b)); }
GNX_HD float length_sq(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
GNX_HD float length(V3 a) { return sqrtf(length_sq(a)); }
GNX_HD V3 normalize(V3 a) { return div_recip(a, length(a)); }
GNX_HD V3 vabs(V3 a) { return V3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
GNX_HD V3 cross(V3 a, V3 b) {
    double ax = a.x, ay = a.y, az = a.z, bx = b.x, by = b.y, bz = b.z;
    return V3((float)((ay * bz) - (az * by)), (float)((az * bx) - (ax * bz)), (float)((ax * by) - (ay * bx)));
}
GNX_HD V3 faceforward(V3 n, V3 v) { return (dot(n, v) < 0.f) ? -n : n; }
GNX_HD float max_component(V3 a) { return fmaxf(a.x, fmaxf(a.y, a.z)); }
GNX_HD int max_dimension(V3 v) { return (v.x > v.y) ? ((v.x > v.z) ? 0 : 2) : ((v.y > v.z) ? 1 : 2); }
GNX_HD bool is_black(V3 a) { return a.x == 0.f && a.y == 0.f && a.z == 0.f; }
GNX_HD float lum_y(V3 c) { return 0.212671f * c.x + 0.715160f * c.y + 0.072169f * c.z; }  // core/Spectrum.h:429-432
GNX_HD float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
GNX_HD V3 clamp0(V3 c) { return V3(clampf(c.x, 0.f, GNX_INF), clampf(c.y, 0.f, GNX_INF), clampf(c.z, 0.f, GNX_INF)); }
GNX_HD V3 vsqrt(V3 a) { return V3(sqrtf(a.x), sqrtf(a.y), sqrtf(a.z)); }
GNX_HD float lerpf(float t, float a, float b) { return (1 - t) * a + t * b; }

// core/Geometry.h:988-995
GNX_HD void coordinate_system(V3 v1, V3 *v2, V3 *v3) {
    if (fabsf(v1.x) > fabsf(v1.y))
        *v2 = div_recip(V3(-v1.z, 0, v1.x), sqrtf(v1.x * v1.x + v1.z * v1.z));
    else
        *v2 = div_recip(V3(0, v1.z, -v1.y), sqrtf(v1.y * v1.y + v1.z * v1.z));
    *v3 = cross(v1, *v2);
}

// core/GNXRayTracer.h:179-205
GNX_D float next_float_up(float v) {
    if (finf(v) && v > 0.f) return v;
    if (v == -0.f) v = 0.f;
    uint32_t ui = f2u(v);
    if (v >= 0) ++ui; else --ui;
    return u2f(ui);
}
GNX_D float next_float_down(float v) {
    if (finf(v) && v < 0.f) return v;
    if (v == 0.f) v = -0.f;
    uint32_t ui = f2u(v);
    if (v > 0) --ui; else ++ui;
    return u2f(ui);
}

// OffsetRayOrigin, core/Geometry.h:1408-1422
GNX_D V3 offset_ray_origin(V3 p, V3 pError, V3 n, V3 w) {
    float d = dot(vabs(n), pError);
    V3 offset = d * n;
    if (dot(w, n) < 0) offset = -offset;
    V3 po = p + offset;
    if (offset.x > 0) po.x = next_float_up(po.x); else if (offset.x < 0) po.x = next_float_down(po.x);
    if (offset.y > 0) po.y = next_float_up(po.y); else if (offset.y < 0) po.y = next_float_down(po.y);
    if (offset.z > 0) po.z = next_float_up(po.z); else if (offset.z < 0) po.z = next_float_down(po.z);
    return po;
}

// Row-major 4x4 (core/Transform.h Matrix4x4::m[i][j] == m[4*i+j]).
struct M44 { float m[16]; };

// Transform::operator()(Point3f) with the homogeneous divide, core/Transform.h:198-210
GNX_HD V3 xform_point(const M44 &M, V3 p) {
    const float *m = M.m;
    float xp = m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3];
    float yp = m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7];
    float zp = m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11];
    float wp = m[12] * p.x + m[13] * p.y + m[14] * p.z + m[15];
    if (wp == 1) return V3(xp, yp, zp);
    float inv = 1.0f / wp;
    return V3(inv * xp, inv * yp, inv * zp);
}
// Transform::operator()(Point3f, Vector3f *pError), core/Transform.h:259-282
GNX_HD V3 xform_point_err(const M44 &M, V3 p, V3 *err) {
    const float *m = M.m;
    float xp = (m[0] * p.x + m[1] * p.y) + (m[2] * p.z + m[3]);
    float yp = (m[4] * p.x + m[5] * p.y) + (m[6] * p.z + m[7]);
    float zp = (m[8] * p.x + m[9] * p.y) + (m[10] * p.z + m[11]);
    float wp = (m[12] * p.x + m[13] * p.y) + (m[14] * p.z + m[15]);
    float xs = (fabsf(m[0] * p.x) + fabsf(m[1] * p.y) + fabsf(m[2] * p.z) + fabsf(m[3]));
    float ys = (fabsf(m[4] * p.x) + fabsf(m[5] * p.y) + fabsf(m[6] * p.z) + fabsf(m[7]));
    float zs = (fabsf(m[8] * p.x) + fabsf(m[9] * p.y) + fabsf(m[10] * p.z) + fabsf(m[11]));
    *err = gamma_n(3) * V3(xs, ys, zs);
    if (wp == 1) return V3(xp, yp, zp);
    float inv = 1.0f / wp;
    return V3(inv * xp, inv * yp, inv * zp);
}
// Transform::operator()(Vector3f), core/Transform.h:213-219
GNX_HD V3 xform_vector(const M44 &M, V3 v) {
    const float *m = M.m;
    return V3(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z,
              m[8] * v.x + m[9] * v.y + m[10] * v.z);
}

// FindInterval over a cdf: largest index with cdf[index] <= u, clamped (core/GNXRayTracer.h:336-349)
GNX_HD int find_interval_cdf(const float *cdf, int size, float u) {
    int first = 0, len = size;
    while (len > 0) {
        int half = len >> 1, middle = first + half;
        if (cdf[middle] <= u) { first = middle + 1; len -= half + 1; } else len = half;
    }
    int r = first - 1;
    return r < 0 ? 0 : (r > size - 2 ? size - 2 : r);
}

// The same answer through a guide table: guide[b] = find_interval_cdf(cdf, size, b / G) for b = 0..G (G a power of
// two, so b = floor(u * G) is exact and b / G <= u < (b + 1) / G).  The answer is monotone in u, hence it lies in
// [guide[b], guide[b + 1]]: a 12-step chain of dependent loads becomes one guide load and ~log2(size / G) steps.
GNX_HD int find_interval_guided(const float *cdf, int size, float u, const uint16_t *guide, int G) {
    if (!guide) return find_interval_cdf(cdf, size, u);
    int b = (int)(u * (float)G);
    b = b < 0 ? 0 : (b > G - 1 ? G - 1 : b);
    int lo = guide[b], hi = guide[b + 1];
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (cdf[mid] <= u) lo = mid; else hi = mid - 1;
    }
    return lo;
}

}  // namespace gnx
