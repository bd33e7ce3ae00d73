// gnx_shade.cuh — device functions of the shade stage: rebuild the SurfaceInteraction of the final
// hit, evaluate textures, build the lobe list of a material, sample lights.
#pragma once
#include "gnx_bsdf.cuh"
#include "gnx_bvh.cuh"
#include "gnx_bvh8.cuh"
#include "gnx_sampler.cuh"

namespace gnx {

struct Surface {
    V3 p, pError, n;        // geometric
    V3 ns, dpdu_s, dpdv_s;  // shading.n, shading.dpdu, shading.dpdv
    V3 wo;                  // Normalize(-ray.d)  (Interaction ctor, core/Interaction.h:24-32)
    float u, v;
    int material, light, prim;
    unsigned flags;
    // ray differentials (integrators that carry a RayDifferential: Whitted, DirectLighting, VolPath's camera vertex)
    V3 dpdu, dpdv;          // geometric partial derivatives (SurfaceInteraction::dpdu / dpdv)
    V3 dndu, dndv;          // shading.dndu / dndv (zero without vertex normals)
    V3 dpdx, dpdy;
    float dudx, dvdx, dudy, dvdy;
};

// RayDifferential's extra members (core/Geometry.h:856-890)
struct RayDiff {
    bool has;
    V3 rxo, ryo, rxd, ryd;
};

// Triangle::Intersect from "Compute triangle partial derivatives" on (shape/Triangle.cpp:170-303),
// for the accepted hit only.
template <bool DIFF = false>
GNX_D Surface make_surface(const DeviceScene &sc, int prim, float b0, float b1, float b2, V3 rayD) {
    float4 c;
    const TriVerts tv = load_tri(sc.tris, prim, &c);
    const unsigned mf = f2u(c.y);
    Surface s;
    s.prim = prim;
    s.material = (int)(mf & 0xfffffu) - 1;  // bits 0-19: material index + 1 (0 = none), 20-23: shade type
    s.flags = mf >> 24;
    s.light = f2i(c.z);
    const V3 p0 = tv.p0, p1 = tv.p1, p2 = tv.p2;
    float uv0x = 0, uv0y = 0, uv1x = 1, uv1y = 0, uv2x = 1, uv2y = 1;  // shape/Triangle.h:60-74
    if (sc.tri_uv) {
        const float *q = sc.tri_uv + 6 * (size_t)prim;
        uv0x = q[0]; uv0y = q[1]; uv1x = q[2]; uv1y = q[3]; uv2x = q[4]; uv2y = q[5];
    }
    const float duv02x = uv0x - uv2x, duv02y = uv0y - uv2y, duv12x = uv1x - uv2x, duv12y = uv1y - uv2y;
    const V3 dp02 = p0 - p2, dp12 = p1 - p2;
    const float determinant = duv02x * duv12y - duv02y * duv12x;
    const bool degenerateUV = fabsf(determinant) < 1e-8f;
    V3 dpdu, dpdv;
    if (!degenerateUV) {
        float invdet = 1 / determinant;
        dpdu = (duv12y * dp02 - duv02y * dp12) * invdet;
        dpdv = (-duv12x * dp02 + duv02x * dp12) * invdet;
    }
    if (degenerateUV || length_sq(cross(dpdu, dpdv)) == 0) {
        V3 ng = cross(p2 - p0, p1 - p0);
        coordinate_system(normalize(ng), &dpdu, &dpdv);
    }
    const float xAbsSum = (fabsf(b0 * p0.x) + fabsf(b1 * p1.x) + fabsf(b2 * p2.x));
    const float yAbsSum = (fabsf(b0 * p0.y) + fabsf(b1 * p1.y) + fabsf(b2 * p2.y));
    const float zAbsSum = (fabsf(b0 * p0.z) + fabsf(b1 * p1.z) + fabsf(b2 * p2.z));
    s.pError = gamma_n(7) * V3(xAbsSum, yAbsSum, zAbsSum);
    s.u = b0 * uv0x + b1 * uv1x + b2 * uv2x;
    s.v = b0 * uv0y + b1 * uv1y + b2 * uv2y;
    s.p = b0 * p0 + b1 * p1 + b2 * p2;
    s.wo = normalize(-rayD);
    s.n = normalize(cross(dp02, dp12));
    if (s.flags & GNX_PRIM_FLIP_N) s.n = -s.n;
    s.ns = s.n;
    s.dpdu_s = dpdu;
    s.dpdv_s = dpdv;
    s.dpdu = dpdu; s.dpdv = dpdv;
    s.dndu = V3(0.f); s.dndv = V3(0.f);
    s.dpdx = V3(0.f); s.dpdy = V3(0.f);
    s.dudx = s.dvdx = s.dudy = s.dvdy = 0.f;
    if (sc.tri_has_n && sc.tri_has_n[prim]) {
        const float *q = sc.tri_n + 9 * (size_t)prim;
        V3 n0(q[0], q[1], q[2]), n1(q[3], q[4], q[5]), n2(q[6], q[7], q[8]);
        if (DIFF) {
            // dndu / dndv of the triangle's shading geometry (shape/Triangle.cpp:262-293)
            const V3 dn1 = n0 - n2, dn2 = n1 - n2;
            if (degenerateUV) {
                V3 dn = cross(n2 - n0, n1 - n0);
                if (length_sq(dn) != 0) coordinate_system(dn, &s.dndu, &s.dndv);
            } else {
                float invDet = 1 / determinant;
                s.dndu = (duv12y * dn1 - duv02y * dn2) * invDet;
                s.dndv = (-duv12x * dn1 + duv02x * dn2) * invDet;
            }
        }
        V3 ns = (b0 * n0 + b1 * n1 + b2 * n2);
        if (length_sq(ns) > 0) ns = normalize(ns); else ns = s.n;
        V3 ss = normalize(dpdu);
        V3 ts = cross(ss, ns);
        if (length_sq(ts) > 0.f) { ts = normalize(ts); ss = cross(ts, ns); }
        else coordinate_system(ns, &ss, &ts);
        if (s.flags & GNX_PRIM_REVERSE_ORI) ts = -ts;
        // SetShadingGeometry(ss, ts, ..., orientationIsAuthoritative = true), core/Interaction.cpp:36-54
        s.ns = normalize(cross(ss, ts));
        s.n = faceforward(s.n, s.ns);
        s.dpdu_s = ss;
        s.dpdv_s = ts;
    }
    return s;
}

// MIPMap::triangle(0, st) with MIPMap::Texel wrap handling (core/MIPMap.h:201-256); PathIntegrator
// reaches no other branch of MIPMap::Lookup because its rays carry no differentials.
GNX_D V3 texel_fetch(const DevTexture &t, int s, int tt) {
    if (t.wrap == GNX_WRAP_REPEAT) {
        s = s % t.w; if (s < 0) s += t.w;
        tt = tt % t.h; if (tt < 0) tt += t.h;
    } else if (t.wrap == GNX_WRAP_CLAMP) {
        s = s < 0 ? 0 : (s > t.w - 1 ? t.w - 1 : s);
        tt = tt < 0 ? 0 : (tt > t.h - 1 ? t.h - 1 : tt);
    } else if (s < 0 || s >= t.w || tt < 0 || tt >= t.h) {
        return V3(0.f);
    }
    const float *q = t.texels + ((size_t)tt * t.w + s) * t.nch;
    return t.nch == 3 ? V3(ldg(q), ldg(q + 1), ldg(q + 2)) : V3(ldg(q));
}
GNX_D V3 texture_bilinear(const DevTexture &t, float su, float sv) {
    float s = su * t.w - 0.5f, tt = sv * t.h - 0.5f;
    int s0 = (int)floorf(s), t0 = (int)floorf(tt);
    float ds = s - s0, dt = tt - t0;
    return (1 - ds) * (1 - dt) * texel_fetch(t, s0, t0) + (1 - ds) * dt * texel_fetch(t, s0, t0 + 1) +
           ds * (1 - dt) * texel_fetch(t, s0 + 1, t0) + ds * dt * texel_fetch(t, s0 + 1, t0 + 1);
}
// ---- MIPMap::Lookup with texture-space differentials (core/MIPMap.h:226-337): trilinear or EWA over the pyramid ----
GNX_D int mip_w(const DevTexture &t, int level) { int w = t.w >> level; return w < 1 ? 1 : w; }
GNX_D int mip_h(const DevTexture &t, int level) { int h = t.h >> level; return h < 1 ? 1 : h; }
// MIPMap::Texel(level, s, t)
GNX_D V3 mip_texel(const DevTexture &t, int level, int s, int tt) {
    const int w = mip_w(t, level), h = mip_h(t, level);
    if (t.wrap == GNX_WRAP_REPEAT) {
        s = s % w; if (s < 0) s += w;
        tt = tt % h; if (tt < 0) tt += h;
    } else if (t.wrap == GNX_WRAP_CLAMP) {
        s = s < 0 ? 0 : (s > w - 1 ? w - 1 : s);
        tt = tt < 0 ? 0 : (tt > h - 1 ? h - 1 : tt);
    } else if (s < 0 || s >= w || tt < 0 || tt >= h) {
        return V3(0.f);
    }
    const float *q = t.texels + ((size_t)t.level_off[level] + (size_t)tt * w + s) * t.nch;
    return t.nch == 3 ? V3(ldg(q), ldg(q + 1), ldg(q + 2)) : V3(ldg(q));
}
// MIPMap::triangle(level, st)
GNX_D V3 mip_triangle(const DevTexture &t, int level, float su, float sv) {
    level = level < 0 ? 0 : (level > t.n_levels - 1 ? t.n_levels - 1 : level);
    float s = su * mip_w(t, level) - 0.5f, tt = sv * mip_h(t, level) - 0.5f;
    int s0 = (int)floorf(s), t0 = (int)floorf(tt);
    float ds = s - s0, dt = tt - t0;
    return (1 - ds) * (1 - dt) * mip_texel(t, level, s0, t0) + (1 - ds) * dt * mip_texel(t, level, s0, t0 + 1) +
           ds * (1 - dt) * mip_texel(t, level, s0 + 1, t0) + ds * dt * mip_texel(t, level, s0 + 1, t0 + 1);
}
GNX_D float log2_ref(float x) { return logf(x) * 1.442695040888963387004650940071f; }  // Log2(), core/GNXRayTracer.h:259-262
// MIPMap::EWA(level, st, dst0, dst1)
GNX_D V3 mip_ewa(const DevTexture &t, const float *lut, int level, float st0, float st1, float d00, float d01, float d10, float d11) {
    if (level >= t.n_levels) return mip_texel(t, t.n_levels - 1, 0, 0);
    const int w = mip_w(t, level), h = mip_h(t, level);
    st0 = st0 * w - 0.5f; st1 = st1 * h - 0.5f;
    d00 *= w; d01 *= h; d10 *= w; d11 *= h;
    float A = d01 * d01 + d11 * d11 + 1;
    float B = -2 * (d00 * d01 + d10 * d11);
    float C = d00 * d00 + d10 * d10 + 1;
    float invF = 1 / (A * C - B * B * 0.25f);
    A *= invF; B *= invF; C *= invF;
    float det = -B * B + 4 * A * C;
    float invDet = 1 / det;
    float uSqrt = sqrtf(det * C), vSqrt = sqrtf(A * det);
    int s0 = (int)ceilf(st0 - 2 * invDet * uSqrt), s1 = (int)floorf(st0 + 2 * invDet * uSqrt);
    int t0 = (int)ceilf(st1 - 2 * invDet * vSqrt), t1 = (int)floorf(st1 + 2 * invDet * vSqrt);
    V3 sum(0.f);
    float sumWts = 0;
    for (int it = t0; it <= t1; ++it) {
        float tt = it - st1;
        for (int is = s0; is <= s1; ++is) {
            float ss = is - st0;
            float r2 = A * ss * ss + B * ss * tt + C * tt * tt;
            if (r2 < 1) {
                int index = (int)(r2 * 128);
                if (index > 127) index = 127;
                float weight = ldg(lut + index);
                sum += mip_texel(t, level, is, it) * weight;
                sumWts += weight;
            }
        }
    }
    return div_each(sum, sumWts);
}
// MIPMap::Lookup(st, dst0, dst1)
GNX_D V3 mip_lookup(const DevTexture &t, const float *lut, float su, float sv, float d00, float d01, float d10, float d11) {
    if (t.do_trilinear) {
        float width = fmaxf(fmaxf(fabsf(d00), fabsf(d01)), fmaxf(fabsf(d10), fabsf(d11)));
        float level = t.n_levels - 1 + log2_ref(fmaxf(width, 1e-8f));
        if (level < 0) return mip_triangle(t, 0, su, sv);
        if (level >= t.n_levels - 1) return mip_texel(t, t.n_levels - 1, 0, 0);
        int iLevel = (int)floorf(level);
        float delta = level - iLevel;
        return (1 - delta) * mip_triangle(t, iLevel, su, sv) + delta * mip_triangle(t, iLevel + 1, su, sv);
    }
    if (d00 * d00 + d01 * d01 < d10 * d10 + d11 * d11) { float a = d00, b = d01; d00 = d10; d01 = d11; d10 = a; d11 = b; }
    float majorLength = sqrtf(d00 * d00 + d01 * d01), minorLength = sqrtf(d10 * d10 + d11 * d11);
    if (minorLength * t.max_aniso < majorLength && minorLength > 0) {
        float scale = majorLength / (minorLength * t.max_aniso);
        d10 *= scale; d11 *= scale;
        minorLength *= scale;
    }
    if (minorLength == 0) return mip_triangle(t, 0, su, sv);
    float lod = fmaxf(0.f, t.n_levels - 1.f + log2_ref(minorLength));
    int ilod = (int)floorf(lod);
    float dl = lod - ilod;
    return (1 - dl) * mip_ewa(t, lut, ilod, su, sv, d00, d01, d10, d11) + dl * mip_ewa(t, lut, ilod + 1, su, sv, d00, d01, d10, d11);
}
// ImageTexture::Evaluate (textures/ImageTexture.h:55-62) through UVMapping2D::Map (core/Texture.cpp:168-175).  DIFF =
// false: the integrator carries no ray differentials (PathIntegrator slices them off, integrators/PathIntegrator.cpp:67),
// every lookup is MIPMap::triangle(0, st).
template <bool DIFF>
GNX_D V3 eval_texture(const DeviceScene &sc, const DevTexture &t, const Surface &s) {
    const float su = t.su * s.u + t.du, sv = t.sv * s.v + t.dv;
    if (!DIFF) return texture_bilinear(t, su, sv);
    return mip_lookup(t, sc.ewa_lut, su, sv, t.su * s.dudx, t.sv * s.dvdx, t.su * s.dudy, t.sv * s.dvdy);
}
template <bool DIFF = false>
GNX_D V3 eval_rgb(const DeviceScene &sc, const gnx_material &m, int slot, const Surface &s) {
    int tex = m.rgb_tex[slot];
    if (tex < 0) return V3(m.rgb[slot][0], m.rgb[slot][1], m.rgb[slot][2]);
    return eval_texture<DIFF>(sc, sc.textures[tex], s);
}
template <bool DIFF = false>
GNX_D float eval_f(const DeviceScene &sc, const gnx_material &m, int slot, const Surface &s) {
    int tex = m.f_tex[slot];
    if (tex < 0) return m.f[slot];
    return eval_texture<DIFF>(sc, sc.textures[tex], s).x;
}

// SurfaceInteraction::ComputeDifferentials (core/Interaction.cpp:65-114)
GNX_D void compute_differentials(Surface &s, const RayDiff &rd) {
    s.dudx = s.dvdx = s.dudy = s.dvdy = 0.f;
    s.dpdx = V3(0.f); s.dpdy = V3(0.f);
    if (!rd.has) return;
    const float d = dot(s.n, s.p);
    const float tx = -(dot(s.n, rd.rxo) - d) / dot(s.n, rd.rxd);
    if (finf(tx) || tx != tx) return;
    const V3 px = rd.rxo + tx * rd.rxd;
    const float ty = -(dot(s.n, rd.ryo) - d) / dot(s.n, rd.ryd);
    if (finf(ty) || ty != ty) return;
    const V3 py = rd.ryo + ty * rd.ryd;
    s.dpdx = px - s.p;
    s.dpdy = py - s.p;
    int d0, d1;
    if (fabsf(s.n.x) > fabsf(s.n.y) && fabsf(s.n.x) > fabsf(s.n.z)) { d0 = 1; d1 = 2; }
    else if (fabsf(s.n.y) > fabsf(s.n.z)) { d0 = 0; d1 = 2; }
    else { d0 = 0; d1 = 1; }
    const float A00 = s.dpdu[d0], A01 = s.dpdv[d0], A10 = s.dpdu[d1], A11 = s.dpdv[d1];
    const float Bx0 = px[d0] - s.p[d0], Bx1 = px[d1] - s.p[d1], By0 = py[d0] - s.p[d0], By1 = py[d1] - s.p[d1];
    // SolveLinearSystem2x2 (core/Transform.cpp:12-20)
    const float det = A00 * A11 - A01 * A10;
    if (fabsf(det) < 1e-10f) return;
    float x0 = (A11 * Bx0 - A01 * Bx1) / det, x1 = (A00 * Bx1 - A10 * Bx0) / det;
    if (!(x0 != x0 || x1 != x1)) { s.dudx = x0; s.dvdx = x1; }
    x0 = (A11 * By0 - A01 * By1) / det; x1 = (A00 * By1 - A10 * By0) / det;
    if (!(x0 != x0 || x1 != x1)) { s.dudy = x0; s.dvdy = x1; }
}

GNX_D Lobe make_lobe(int kind, int type, V3 R) {
    Lobe l;
    l.kind = kind; l.type = type; l.fresnel = FR_NOOP; l.distrib = DK_TROWBRIDGE;
    l.R = R; l.a = V3(0.f); l.b = V3(0.f);
    l.p0 = l.p1 = 0; l.e0 = l.e1 = 1;
    return l;
}
GNX_D float clamp_alpha(float a) { return fmaxf(0.001f, a); }  // TrowbridgeReitzDistribution ctor, MicroFacet.h:82-83

// <Material>::ComputeScatteringFunctions (materials/*.cpp), allowMultipleLobes == true,
// TransportMode::Radiance.  Also applies Material::Bump's re-derivation of shading.n when a
// (constant) bump map is attached, and fills the BSDF frame (core/Reflection.h:106-111).
// multiLobes: the allowMultipleLobes argument (true for Path / VolPath, false for Whitted / DirectLighting; only
// GlassMaterial looks at it, materials/GlassMaterial.cpp:31).
template <int MAXL, bool DIFF = false>
GNX_D void build_bsdf(const DeviceScene &sc, const gnx_material &m, Surface &s, Bsdf<MAXL> &b, bool multiLobes = true) {
    if (m.flags & GNX_MATF_BUMP_IDENTITY) {
        // SetShadingGeometry(dpdu, dpdv, ..., false) with unchanged dpdu/dpdv (core/Material.cpp:45-51)
        s.ns = normalize(cross(s.dpdu_s, s.dpdv_s));
        s.ns = faceforward(s.ns, s.n);
    }
    b.n = 0;
    b.eta = 1;
    b.ns = s.ns;
    b.ng = s.n;
    b.ss = normalize(s.dpdu_s);
    b.ts = cross(b.ns, b.ss);
    switch (m.type) {
    case GNX_MAT_MATTE: {
        V3 r = clamp0(eval_rgb<DIFF>(sc, m, 0, s));
        float sig = clampf(eval_f<DIFF>(sc, m, 0, s), 0, 90);
        if (!is_black(r)) {
            if (sig == 0) b.add(make_lobe(LK_LAMBERT_R, BSDF_REFLECTION | BSDF_DIFFUSE, r));
            else {
                Lobe l = make_lobe(LK_OREN_NAYAR, BSDF_REFLECTION | BSDF_DIFFUSE, r);
                float sg = (kPi / 180.f) * sig;  // Radians(), core/GNXRayTracer.h
                float sigma2 = sg * sg;
                l.p0 = 1.f - (sigma2 / (2.f * (sigma2 + 0.33f)));
                l.p1 = 0.45f * sigma2 / (sigma2 + 0.09f);
                b.add(l);
            }
        }
        break;
    }
    case GNX_MAT_MIRROR: {
        V3 R = clamp0(eval_rgb<DIFF>(sc, m, 0, s));
        if (!is_black(R)) b.add(make_lobe(LK_SPEC_R, BSDF_REFLECTION | BSDF_SPECULAR, R));
        break;
    }
    case GNX_MAT_PLASTIC: {
        V3 kd = clamp0(eval_rgb<DIFF>(sc, m, 0, s));
        if (!is_black(kd)) b.add(make_lobe(LK_LAMBERT_R, BSDF_REFLECTION | BSDF_DIFFUSE, kd));
        V3 ks = clamp0(eval_rgb<DIFF>(sc, m, 1, s));
        if (!is_black(ks)) {
            float rough = eval_f<DIFF>(sc, m, 0, s);
            if (m.flags & GNX_MATF_REMAP_ROUGHNESS) rough = roughness_to_alpha(rough);
            Lobe l = make_lobe(LK_MICRO_R, BSDF_REFLECTION | BSDF_GLOSSY, ks);
            l.fresnel = FR_DIELECTRIC; l.e0 = 1.5f; l.e1 = 1.f;
            l.p0 = l.p1 = clamp_alpha(rough);
            b.add(l);
        }
        break;
    }
    case GNX_MAT_METAL: {
        float ur = eval_f<DIFF>(sc, m, 0, s), vr = eval_f<DIFF>(sc, m, 1, s);
        if (m.flags & GNX_MATF_REMAP_ROUGHNESS) { ur = roughness_to_alpha(ur); vr = roughness_to_alpha(vr); }
        Lobe l = make_lobe(LK_MICRO_R, BSDF_REFLECTION | BSDF_GLOSSY, V3(1.f));
        l.fresnel = FR_CONDUCTOR; l.e0 = 1.f;
        l.a = eval_rgb<DIFF>(sc, m, 0, s);
        l.b = eval_rgb<DIFF>(sc, m, 1, s);
        l.p0 = clamp_alpha(ur); l.p1 = clamp_alpha(vr);
        b.add(l);
        break;
    }
    case GNX_MAT_GLASS: {
        float eta = eval_f<DIFF>(sc, m, 2, s);
        float ur = eval_f<DIFF>(sc, m, 0, s), vr = eval_f<DIFF>(sc, m, 1, s);
        V3 R = clamp0(eval_rgb<DIFF>(sc, m, 0, s)), T = clamp0(eval_rgb<DIFF>(sc, m, 1, s));
        b.eta = eta;
        if (is_black(R) && is_black(T)) break;
        bool isSpecular = ur == 0 && vr == 0;
        if (isSpecular && multiLobes) {
            Lobe l = make_lobe(LK_FRESNEL_SPEC, BSDF_REFLECTION | BSDF_TRANSMISSION | BSDF_SPECULAR, R);
            l.a = T; l.e0 = 1.f; l.e1 = eta;
            b.add(l);
        } else if (isSpecular) {
            if (!is_black(R)) {
                Lobe l = make_lobe(LK_SPEC_R, BSDF_REFLECTION | BSDF_SPECULAR, R);
                l.fresnel = FR_DIELECTRIC; l.e0 = 1.f; l.e1 = eta;
                b.add(l);
            }
            if (!is_black(T)) {
                Lobe l = make_lobe(LK_SPEC_T, BSDF_TRANSMISSION | BSDF_SPECULAR, T);
                l.e0 = 1.f; l.e1 = eta;
                b.add(l);
            }
        } else {
            if (m.flags & GNX_MATF_REMAP_ROUGHNESS) { ur = roughness_to_alpha(ur); vr = roughness_to_alpha(vr); }
            if (!is_black(R)) {
                Lobe l = make_lobe(LK_MICRO_R, BSDF_REFLECTION | BSDF_GLOSSY, R);
                l.fresnel = FR_DIELECTRIC; l.e0 = 1.f; l.e1 = eta;
                l.p0 = clamp_alpha(ur); l.p1 = clamp_alpha(vr);
                b.add(l);
            }
            if (!is_black(T)) {
                Lobe l = make_lobe(LK_MICRO_T, BSDF_TRANSMISSION | BSDF_GLOSSY, T);
                l.e0 = 1.f; l.e1 = eta;
                l.p0 = clamp_alpha(ur); l.p1 = clamp_alpha(vr);
                b.add(l);
            }
        }
        break;
    }
    case GNX_MAT_DISNEY: {  // materials/DisneyMaterial.cpp:467-581
        const V3 c = clamp0(eval_rgb<DIFF>(sc, m, 0, s));
        const float metallicWeight = eval_f<DIFF>(sc, m, 0, s), e = eval_f<DIFF>(sc, m, 1, s), strans = eval_f<DIFF>(sc, m, 9, s);
        const float diffuseWeight = (1 - metallicWeight) * (1 - strans);
        const float dt = eval_f<DIFF>(sc, m, 11, s) / 2;
        const float rough = eval_f<DIFF>(sc, m, 2, s);
        const float lum = lum_y(c);
        const V3 Ctint = lum > 0 ? div_each(c, lum) : V3(1.f);
        const float sheenWeight = eval_f<DIFF>(sc, m, 5, s);
        V3 Csheen(0.f);
        if (sheenWeight > 0) { float stint = eval_f<DIFF>(sc, m, 6, s); Csheen = (1 - stint) * V3(1.f) + stint * Ctint; }
        const bool thin = (m.flags & GNX_MATF_THIN) != 0;
        if (diffuseWeight > 0) {
            if (thin) {
                float flat = eval_f<DIFF>(sc, m, 10, s);
                b.add(make_lobe(LK_DISNEY_DIFFUSE, BSDF_REFLECTION | BSDF_DIFFUSE, (diffuseWeight * (1 - flat) * (1 - dt)) * c));
                Lobe l = make_lobe(LK_DISNEY_FAKESS, BSDF_REFLECTION | BSDF_DIFFUSE, (diffuseWeight * flat * (1 - dt)) * c);
                l.p0 = rough;
                b.add(l);
            } else {
                V3 sd = eval_rgb<DIFF>(sc, m, 1, s);
                if (is_black(sd)) b.add(make_lobe(LK_DISNEY_DIFFUSE, BSDF_REFLECTION | BSDF_DIFFUSE, diffuseWeight * c));
                else {
                    // the BSSRDF itself is compiled out of both integrators (PathIntegrator.cpp:165-192); the
                    // SpecularTransmission lobe that goes with it is still added
                    Lobe l = make_lobe(LK_SPEC_T, BSDF_TRANSMISSION | BSDF_SPECULAR, V3(1.f));
                    l.e0 = 1.f; l.e1 = e;
                    b.add(l);
                }
            }
            Lobe r = make_lobe(LK_DISNEY_RETRO, BSDF_REFLECTION | BSDF_DIFFUSE, diffuseWeight * c);
            r.p0 = rough;
            b.add(r);
            if (sheenWeight > 0) b.add(make_lobe(LK_DISNEY_SHEEN, BSDF_REFLECTION | BSDF_DIFFUSE, (diffuseWeight * sheenWeight) * Csheen));
        }
        const float aspect = (float)sqrt(1 - eval_f<DIFF>(sc, m, 4, s) * .9);
        const float ax = fmaxf(.001f, (rough * rough) / aspect), ay = fmaxf(.001f, (rough * rough) * aspect);
        const float specTint = eval_f<DIFF>(sc, m, 3, s);
        const float r0 = ((e - 1) * (e - 1)) / ((e + 1) * (e + 1));  // SchlickR0FromEta
        const V3 tintMix = (1 - specTint) * V3(1.f) + specTint * Ctint;
        const V3 Cspec0 = (1 - metallicWeight) * (r0 * tintMix) + metallicWeight * c;
        {
            Lobe l = make_lobe(LK_MICRO_R, BSDF_REFLECTION | BSDF_GLOSSY, V3(1.f));
            l.fresnel = FR_DISNEY; l.distrib = DK_DISNEY;
            l.a = Cspec0; l.e0 = metallicWeight; l.e1 = e;
            l.p0 = clamp_alpha(ax); l.p1 = clamp_alpha(ay);
            b.add(l);
        }
        const float cc = eval_f<DIFF>(sc, m, 7, s);
        if (cc > 0) {
            Lobe l = make_lobe(LK_DISNEY_CLEARCOAT, BSDF_REFLECTION | BSDF_GLOSSY, V3(0.f));
            l.p0 = cc;
            l.p1 = lerpf(eval_f<DIFF>(sc, m, 8, s), .1f, .001f);
            b.add(l);
        }
        if (strans > 0) {
            V3 T = strans * vsqrt(c);
            Lobe l = make_lobe(LK_MICRO_T, BSDF_TRANSMISSION | BSDF_GLOSSY, T);
            l.e0 = 1.f; l.e1 = e;
            if (thin) {
                float rscaled = (0.65f * e - 0.35f) * rough;
                l.p0 = clamp_alpha(fmaxf(.001f, (rscaled * rscaled) / aspect));
                l.p1 = clamp_alpha(fmaxf(.001f, (rscaled * rscaled) * aspect));
            } else {
                l.distrib = DK_DISNEY;
                l.p0 = clamp_alpha(ax); l.p1 = clamp_alpha(ay);
            }
            b.add(l);
        }
        if (thin) b.add(make_lobe(LK_LAMBERT_T, BSDF_TRANSMISSION | BSDF_DIFFUSE, dt * c));
        break;
    }
    default: break;
    }
}

// ---- lights ----------------------------------------------------------------------------------------------------
struct LightSample {
    V3 wi, Li;
    float pdf;
    V3 pl, nl, plError;  // point on the light, its normal and error bound (VisibilityTester p1)
};

// DiffuseAreaLight::L with the reference's bool truncation (lights/DiffuseAreaLight.h:22-27):
// `bool dotNW = Dot(n, w)` is true for any non-zero dot product, so the light emits on both sides.
GNX_D V3 area_light_L(const gnx_light &l, V3 n, V3 w) {
    bool dotNW = dot(n, w) != 0.f;
    return (l.two_sided || dotNW) ? V3(l.L[0], l.L[1], l.L[2]) : V3(0.f);
}

// DiffuseAreaLight::Sample_Li -> Shape::Sample(ref,u) -> Triangle::Sample(u)
// (lights/DiffuseAreaLight.cpp:37-52, core/Shape.cpp:21-35, shape/Triangle.cpp:464-492)
GNX_D bool area_sample_li(const DeviceScene &sc, const gnx_light &l, V3 refP, float u0, float u1, LightSample *ls) {
    float4 c;
    const TriVerts tv = load_tri(sc.tris, l.prim, &c);
    const unsigned flags = f2u(c.y) >> 24;
    float su0 = sqrtf(u0);
    float bb0 = 1 - su0, bb1 = u1 * su0;
    V3 p = bb0 * tv.p0 + bb1 * tv.p1 + (1 - bb0 - bb1) * tv.p2;
    V3 n = normalize(cross(tv.p1 - tv.p0, tv.p2 - tv.p0));
    if (sc.tri_has_n && sc.tri_has_n[l.prim]) {
        const float *q = sc.tri_n + 9 * (size_t)l.prim;
        V3 ns = bb0 * V3(q[0], q[1], q[2]) + bb1 * V3(q[3], q[4], q[5]) + (1 - bb0 - bb1) * V3(q[6], q[7], q[8]);
        n = faceforward(n, ns);
    } else if (flags & GNX_PRIM_FLIP_N) n = -n;
    V3 pAbsSum = vabs(bb0 * tv.p0) + vabs(bb1 * tv.p1) + vabs((1 - bb0 - bb1) * tv.p2);
    ls->plError = gamma_n(6) * pAbsSum;
    float pdf = 1 / l.area;
    V3 wi = p - refP;
    if (length_sq(wi) == 0) pdf = 0;
    else {
        wi = normalize(wi);
        V3 dd = refP - p;
        pdf *= length_sq(dd) / absdot(n, -wi);
        if (finf(pdf)) pdf = 0.f;
    }
    ls->pl = p; ls->nl = n;
    if (pdf == 0 || length_sq(p - refP) == 0) { ls->pdf = 0; ls->Li = V3(0.f); return false; }
    ls->wi = normalize(p - refP);
    ls->pdf = pdf;
    ls->Li = area_light_L(l, n, -ls->wi);
    return true;
}

// DiffuseAreaLight::Pdf_Li -> Shape::Pdf(ref, wi): intersects the light's own triangle
// (core/Shape.cpp:37-53).  `o` is ref.SpawnRay(wi).o.
GNX_D float area_pdf_li(const DeviceScene &sc, const gnx_light &l, V3 refP, V3 o, V3 wi) {
    const TriVerts tv = load_tri(sc.tris, l.prim);
    TriHit h;
    const RayShear rs = make_shear(wi);
    if (!intersect_tri(tv, o, rs, GNX_INF, &h) || tri_degenerate(tv)) return 0;
    V3 pl = h.b0 * tv.p0 + h.b1 * tv.p1 + h.b2 * tv.p2;
    V3 nl = normalize(cross(tv.p0 - tv.p2, tv.p1 - tv.p2));
    V3 dd = refP - pl;
    float pdf = length_sq(dd) / (absdot(nl, -wi) * l.area);
    if (finf(pdf)) pdf = 0.f;
    return pdf;
}

// InfiniteAreaLight (lights/InfiniteAreaLight.cpp:91-132) ------------------------------------------------
GNX_D V3 env_lookup(const DevEnv &e, float su, float sv) {
    // Lmap->Lookup(st) -> triangle(0, st), wrap Repeat (the MIPMap default)
    float s = su * e.w - 0.5f, t = sv * e.h - 0.5f;
    int s0 = (int)floorf(s), t0 = (int)floorf(t);
    float ds = s - s0, dt = t - t0;
    // Repeat wrap; the MIPMap resamples to powers of two, where Mod() is a mask (also for negatives)
    const bool pow2 = ((e.w & (e.w - 1)) | (e.h & (e.h - 1))) == 0;
    auto tx = [&](int x, int y) {
        if (pow2) { x &= e.w - 1; y &= e.h - 1; }
        else {
            x = x % e.w; if (x < 0) x += e.w;
            y = y % e.h; if (y < 0) y += e.h;
        }
        const float4 q = ldg(e.texels + (size_t)y * e.w + x);
        return V3(q.x, q.y, q.z);
    };
    return (1 - ds) * (1 - dt) * tx(s0, t0) + (1 - ds) * dt * tx(s0, t0 + 1) + ds * (1 - dt) * tx(s0 + 1, t0) +
           ds * dt * tx(s0 + 1, t0 + 1);
}
GNX_D float spherical_theta(V3 v) { return acosf(clampf(v.z, -1, 1)); }
GNX_D float spherical_phi(V3 v) { float p = atan2f(v.y, v.x); return (p < 0) ? (p + 2 * kPi) : p; }

GNX_D V3 env_Le(const DevEnv &e, V3 rayD) {
    V3 w = normalize(xform_vector(e.w2l, rayD));
    return env_lookup(e, spherical_phi(w) * kInv2Pi, spherical_theta(w) * kInvPi);
}
// Distribution1D::SampleContinuous, core/Sampling.h:40-59
GNX_D float dist1d_sample_continuous(const float *func, const float *cdf, float funcInt, int n, float u, float *pdf, int *off,
                                     const uint16_t *guide = nullptr, int G = 0) {
    int offset = find_interval_guided(cdf, n + 1, u, guide, G);
    *off = offset;
    float du = u - cdf[offset];
    if ((cdf[offset + 1] - cdf[offset]) > 0) du /= (cdf[offset + 1] - cdf[offset]);
    *pdf = (funcInt > 0) ? func[offset] / funcInt : 0;
    return (offset + du) / n;
}
GNX_D bool env_sample_li(const DevEnv &e, float u0, float u1, LightSample *ls) {
    float pdf1, pdf0;
    int v, dummy;
    float d1 = dist1d_sample_continuous(e.marg_func, e.marg_cdf, e.marg_int, e.dh, u1, &pdf1, &v, e.marg_guide, e.marg_g);
    float d0 = dist1d_sample_continuous(e.cond_func + (size_t)v * e.dw, e.cond_cdf + (size_t)v * (e.dw + 1),
                                        e.cond_int[v], e.dw, u0, &pdf0, &dummy,
                                        e.cond_guide ? e.cond_guide + (size_t)v * (e.cond_g + 1) : nullptr, e.cond_g);
    float mapPdf = pdf0 * pdf1;
    ls->pdf = 0;
    ls->Li = V3(0.f);
    if (mapPdf == 0) return false;
    float theta = d1 * kPi, phi = d0 * 2 * kPi;
    float cosTheta = cosf(theta), sinTheta = sinf(theta);
    float sinPhi = sinf(phi), cosPhi = cosf(phi);
    ls->wi = xform_vector(e.l2w, V3(sinTheta * cosPhi, sinTheta * sinPhi, cosTheta));
    ls->pdf = mapPdf / (2 * kPi * kPi * sinTheta);
    if (sinTheta == 0) ls->pdf = 0;
    ls->Li = env_lookup(e, d0, d1);
    return true;
}
GNX_D float env_pdf_li(const DevEnv &e, V3 w) {
    V3 wi = xform_vector(e.w2l, w);
    float theta = spherical_theta(wi), phi = spherical_phi(wi);
    float sinTheta = sinf(theta);
    if (sinTheta == 0) return 0;
    float pu = phi * kInv2Pi, pv = theta * kInvPi;
    int iu = (int)(pu * e.dw); iu = iu < 0 ? 0 : (iu > e.dw - 1 ? e.dw - 1 : iu);
    int iv = (int)(pv * e.dh); iv = iv < 0 ? 0 : (iv > e.dh - 1 ? e.dh - 1 : iv);
    return (e.cond_func[(size_t)iv * e.dw + iu] / e.marg_int) / (2 * kPi * kPi * sinTheta);
}

// Light choice: Distribution1D::SampleDiscrete over the distribution that LightDistribution::Lookup(p)
// returns (core/Sampling.h:60-70, core/LightDistribution.cpp:109-204).
GNX_D int choose_light(const DeviceScene &sc, V3 p, float u, float *pdf) {
    const DevLightDistrib &ld = sc.ld;
    const int n = sc.n_lights;
    const float *func, *cdf;
    float funcInt;
    if (ld.mode == GNX_LIGHTS_SPATIAL) {
        // Bounds3::Offset, core/Geometry.h
        float o[3] = {p.x - sc.wb_min[0], p.y - sc.wb_min[1], p.z - sc.wb_min[2]};
        int pi[3];
        for (int i = 0; i < 3; ++i) {
            if (sc.wb_max[i] > sc.wb_min[i]) o[i] /= sc.wb_max[i] - sc.wb_min[i];
            int v = (int)(o[i] * ld.nvox[i]);
            pi[i] = v < 0 ? 0 : (v > ld.nvox[i] - 1 ? ld.nvox[i] - 1 : v);
        }
        size_t vox = ((size_t)pi[2] * ld.nvox[1] + pi[1]) * ld.nvox[0] + pi[0];
        func = ld.sp_func + vox * n;
        cdf = ld.sp_cdf + vox * (n + 1);
        funcInt = ld.sp_int[vox];
    } else {
        func = ld.uni_func; cdf = ld.uni_cdf; funcInt = ld.uni_int;
    }
    int offset = find_interval_cdf(cdf, n + 1, u);
    *pdf = (funcInt > 0) ? func[offset] / (funcInt * n) : 0;
    return offset;
}

// ---- every light type behind one interface (Light::Sample_Li / Pdf_Li / Le, lights/*.cpp): DiffuseAreaLight and
// InfiniteAreaLight above, plus PointLight, SpotLight, DistantLight and SkyBoxLight.  Used by the wavefront
// PathIntegrator's shade stage when the scene holds one of the latter four, by the spatial light-table build, and by the
// Whitted / DirectLighting kernels (gnx_whitted.cuh).
struct WLightSample {
    V3 wi, Li;
    float pdf;
    V3 target, targetN, targetErr;  // VisibilityTester p1 (normal and error zero for points in space)
    bool delta;                     // IsDeltaLight(flags)
};

// SpotLight::Falloff, lights/SpotLight.cpp:33-43
GNX_D float spot_falloff(const gnx_light &l, V3 w) {
    M44 w2l;
    for (int i = 0; i < 16; ++i) w2l.m[i] = l.world_to_light[i];
    V3 wl = normalize(xform_vector(w2l, w));
    float cosTheta = wl.z;
    if (cosTheta < l.cos_total) return 0;
    if (cosTheta >= l.cos_falloff) return 1;
    float delta = (cosTheta - l.cos_total) / (l.cos_falloff - l.cos_total);
    return (delta * delta) * (delta * delta);
}

// SkyBoxLight::getLightValue, lights/SkyBoxLight.cpp:27-43
GNX_D V3 skybox_value(const DevSkybox &sb, float u, float v) {
    V3 Lv(0.f);
    if (sb.data) {
        int w = (int)(u * sb.w), h = (int)(v * sb.h);
        int offset = (w + h * sb.w) * sb.nc;
        const float scale = 1.0f / 10.0f;
        Lv = V3(ldg(sb.data + offset) * scale, ldg(sb.data + offset + 1) * scale, ldg(sb.data + offset + 2) * scale);
    }
    return Lv;
}
// SkyBoxLight::Le, lights/SkyBoxLight.cpp:57-86 (the reference mixes float and double here: b, t)
GNX_D V3 skybox_le(const DevSkybox &sb, V3 o, V3 d) {
    V3 oc = o - sb.center;
    float a = dot(d, d);
    float b = (float)(2.0 * (double)dot(oc, d));
    float c = dot(oc, oc) - sb.radius * sb.radius;
    float discriminant = b * b - 4 * a * c;
    if (discriminant < 0) return V3(0.f);
    float t = (float)(((double)(-b) + sqrt((double)discriminant)) / (2.0 * (double)a));
    V3 hitPos = o + t * d;
    V3 hp = hitPos - sb.center;
    V3 q = div_each(hp, sb.radius);
    // get_sphere_uv: atan2 / asin are the double overloads in the reference (float arguments promoted)
    float phi = (float)atan2((double)q.z, (double)q.x);
    float theta = (float)asin((double)q.y);
    float u = 1 - (phi + kPi) * kInv2Pi;
    float v = (theta + kPiOver2) * kInvPi;
    if (sb.data) return skybox_value(sb, u, v);
    return V3((hp.x + sb.radius) / (2.f * sb.radius), (hp.y + sb.radius) / (2.f * sb.radius), (hp.z + sb.radius) / (2.f * sb.radius));
}

// Light::Le summed over the scene's lights for a ray that escapes (only the two infinite kinds return non-zero)
GNX_D V3 scene_le(const DeviceScene &sc, V3 o, V3 d) {
    V3 L(0.f);
    // scene.lights order: whichever infinite light comes first is added first
    const bool envFirst = !sc.skybox.present || (sc.env.present && sc.env.light_index < sc.skybox.light_index);
    if (envFirst) {
        if (sc.env.present) L += env_Le(sc.env, d);
        if (sc.skybox.present) L += skybox_le(sc.skybox, o, d);
    } else {
        L += skybox_le(sc.skybox, o, d);
        if (sc.env.present) L += env_Le(sc.env, d);
    }
    return L;
}

// Light::Sample_Li for every light type (lights/*.cpp)
GNX_D bool w_sample_li(const DeviceScene &sc, const gnx_light &l, V3 refP, float u0, float u1, WLightSample *o) {
    o->delta = false;
    o->targetN = V3(0.f);
    o->targetErr = V3(0.f);
    switch (l.type) {
    case GNX_LIGHT_AREA_TRI: {
        LightSample ls;
        bool ok = area_sample_li(sc, l, refP, u0, u1, &ls);
        o->wi = ls.wi; o->Li = ls.Li; o->pdf = ls.pdf;
        o->target = ls.pl; o->targetN = ls.nl; o->targetErr = ls.plError;
        return ok;
    }
    case GNX_LIGHT_INFINITE: {
        LightSample ls;
        bool ok = env_sample_li(sc.env, u0, u1, &ls);
        o->wi = ls.wi; o->Li = ls.Li; o->pdf = ls.pdf;
        o->target = refP + ls.wi * (2 * sc.env.world_radius);
        return ok;
    }
    case GNX_LIGHT_POINT:
    case GNX_LIGHT_SPOT: {
        const V3 pLight(l.p[0], l.p[1], l.p[2]);
        o->wi = normalize(pLight - refP);
        o->pdf = 1.f;
        o->target = pLight;
        o->delta = true;
        V3 I(l.L[0], l.L[1], l.L[2]);
        if (l.type == GNX_LIGHT_SPOT) I = I * spot_falloff(l, -o->wi);
        o->Li = div_each(I, length_sq(pLight - refP));
        return true;
    }
    case GNX_LIGHT_DISTANT: {
        const V3 wLight(l.p[0], l.p[1], l.p[2]);
        o->wi = wLight;
        o->pdf = 1;
        o->target = refP + wLight * (2 * l.area);  // area: the scene's bounding-sphere radius (Preprocess)
        o->delta = true;
        o->Li = V3(l.L[0], l.L[1], l.L[2]);
        return true;
    }
    case GNX_LIGHT_SKYBOX: {
        float theta = u1 * kPi, phi = u0 * 2 * kPi;
        float cosTheta = cosf(theta), sinTheta = sinf(theta);
        float sinPhi = sinf(phi), cosPhi = cosf(phi);
        M44 l2w;
        for (int i = 0; i < 16; ++i) l2w.m[i] = l.world_to_light[i];  // SKYBOX: LightToWorld
        o->wi = xform_vector(l2w, V3(sinTheta * cosPhi, sinTheta * sinPhi, cosTheta));
        o->pdf = 1.f / (4 * kPi);
        o->target = refP + o->wi * (2 * sc.skybox.radius);
        o->Li = 16 * skybox_value(sc.skybox, u0, u1);
        return true;
    }
    default:
        o->pdf = 0; o->Li = V3(0.f);
        return false;
    }
}

// Light::Pdf_Li for the BSDF-sampling half of EstimateDirect (0 for SkyBoxLight and the delta lights)
GNX_D float w_pdf_li(const DeviceScene &sc, const gnx_light &l, V3 refP, V3 rayO, V3 wi) {
    if (l.type == GNX_LIGHT_AREA_TRI) return area_pdf_li(sc, l, refP, rayO, wi);
    if (l.type == GNX_LIGHT_INFINITE) return env_pdf_li(sc.env, wi);
    return 0.f;
}

}  // namespace gnx
