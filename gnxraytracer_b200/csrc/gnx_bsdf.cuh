// gnx_bsdf.cuh — BSDF::f / Sample_f / Pdf over a compact lobe list built from a material record.
//
// The reference allocates a BSDF plus 1..8 BxDF objects in a per-pixel arena at every vertex and
// dispatches virtually (materials/*.cpp, core/Reflection.cpp).  Here a vertex's lobes are a small
// register/local array of tagged records; the semantics that decide parity are kept exactly:
//   lobe choice          comp = min(floor(u0 * n), n - 1), u0 remapped   core/Reflection.cpp:494,509
//   pdf                  averaged over matching lobes                    core/Reflection.cpp:526-530
//   f                    summed over lobes on the geometric-normal side  core/Reflection.cpp:445,535
#pragma once
#include "gnx_scene.cuh"

namespace gnx {

enum : int {
    BSDF_REFLECTION = 1, BSDF_TRANSMISSION = 2, BSDF_DIFFUSE = 4, BSDF_GLOSSY = 8, BSDF_SPECULAR = 16,
    BSDF_ALL = 31
};

enum LobeKind : int {
    LK_LAMBERT_R = 0, LK_OREN_NAYAR, LK_SPEC_R, LK_SPEC_T, LK_FRESNEL_SPEC, LK_MICRO_R, LK_MICRO_T, LK_LAMBERT_T,
    LK_DISNEY_DIFFUSE, LK_DISNEY_FAKESS, LK_DISNEY_RETRO, LK_DISNEY_SHEEN, LK_DISNEY_CLEARCOAT
};
enum FresnelKind : int { FR_NOOP = 0, FR_DIELECTRIC, FR_CONDUCTOR, FR_DISNEY };
enum DistribKind : int { DK_TROWBRIDGE = 0, DK_DISNEY = 1 };

struct Lobe {
    int kind, type, fresnel, distrib;
    V3 R;            // R, or T for transmissive lobes
    V3 a, b;         // FR_CONDUCTOR: eta, k.  LK_FRESNEL_SPEC: a = T.  FR_DISNEY: a = R0
    float p0, p1;    // alphax, alphay | OrenNayar A, B | clearcoat weight, gloss | roughness
    float e0, e1;    // etaI/etaA, etaT/etaB | FR_DISNEY: metallic, eta
};

template <int MAXL>
struct Bsdf {
    V3 ns, ng, ss, ts;
    float eta;
    int n;
    Lobe lobes[MAXL];
    GNX_D V3 to_local(V3 v) const { return V3(dot(v, ss), dot(v, ts), dot(v, ns)); }
    GNX_D V3 to_world(V3 v) const {
        return V3(ss.x * v.x + ts.x * v.y + ns.x * v.z, ss.y * v.x + ts.y * v.y + ns.y * v.z,
                  ss.z * v.x + ts.z * v.y + ns.z * v.z);
    }
    GNX_D void add(const Lobe &l) { if (n < MAXL) lobes[n++] = l; }
    GNX_D int num_components(int flags) const {
        int c = 0;
        for (int i = 0; i < n; ++i) if ((lobes[i].type & flags) == lobes[i].type) ++c;
        return c;
    }
};

// ---- core/Reflection.h:18-60 ---------------------------------------------------------------------
GNX_D float cos_theta(V3 w) { return w.z; }
GNX_D float cos2_theta(V3 w) { return w.z * w.z; }
GNX_D float abs_cos_theta(V3 w) { return fabsf(w.z); }
GNX_D float sin2_theta(V3 w) { return fmaxf(0.f, 1.f - cos2_theta(w)); }
GNX_D float sin_theta(V3 w) { return sqrtf(sin2_theta(w)); }
GNX_D float tan_theta(V3 w) { return sin_theta(w) / cos_theta(w); }
GNX_D float tan2_theta(V3 w) { return sin2_theta(w) / cos2_theta(w); }
GNX_D float cos_phi(V3 w) { float s = sin_theta(w); return (s == 0) ? 1 : clampf(w.x / s, -1, 1); }
GNX_D float sin_phi(V3 w) { float s = sin_theta(w); return (s == 0) ? 0 : clampf(w.y / s, -1, 1); }
GNX_D float cos2_phi(V3 w) { return cos_phi(w) * cos_phi(w); }
GNX_D float sin2_phi(V3 w) { return sin_phi(w) * sin_phi(w); }
GNX_D bool same_hemisphere(V3 w, V3 wp) { return w.z * wp.z > 0; }
GNX_D V3 reflect(V3 wo, V3 n) { return -wo + 2 * dot(wo, n) * n; }
GNX_D bool refract(V3 wi, V3 n, float eta, V3 *wt) {
    float cosThetaI = dot(n, wi);
    float sin2ThetaI = fmaxf(0.f, 1 - cosThetaI * cosThetaI);
    float sin2ThetaT = eta * eta * sin2ThetaI;
    if (sin2ThetaT >= 1) return false;
    float cosThetaT = sqrtf(1 - sin2ThetaT);
    *wt = eta * -wi + (eta * cosThetaI - cosThetaT) * n;
    return true;
}

// ---- sampling warps, core/Sampling.cpp:87-105, core/Sampling.h:140-145 ------------------------------
GNX_D void concentric_sample_disk(float u0, float u1, float *dx, float *dy) {
    float ox = 2.f * u0 - 1, oy = 2.f * u1 - 1;
    if (ox == 0 && oy == 0) { *dx = 0; *dy = 0; return; }
    float theta, r;
    if (fabsf(ox) > fabsf(oy)) { r = ox; theta = kPiOver4 * (oy / ox); }
    else { r = oy; theta = kPiOver2 - kPiOver4 * (ox / oy); }
    *dx = r * cosf(theta);
    *dy = r * sinf(theta);
}
GNX_D V3 cosine_sample_hemisphere(float u0, float u1) {
    float dx, dy;
    concentric_sample_disk(u0, u1, &dx, &dy);
    float z = sqrtf(fmaxf(0.f, 1 - dx * dx - dy * dy));
    return V3(dx, dy, z);
}

// ---- Fresnel, core/Reflection.cpp:16-64 -------------------------------------------------------------
GNX_LOBE_FN float fr_dielectric(float cosThetaI, float etaI, float etaT) {
    cosThetaI = clampf(cosThetaI, -1, 1);
    bool entering = cosThetaI > 0.f;
    if (!entering) { float t = etaI; etaI = etaT; etaT = t; cosThetaI = fabsf(cosThetaI); }
    float sinThetaI = sqrtf(fmaxf(0.f, 1 - cosThetaI * cosThetaI));
    float sinThetaT = etaI / etaT * sinThetaI;
    if (sinThetaT >= 1) return 1;
    float cosThetaT = sqrtf(fmaxf(0.f, 1 - sinThetaT * sinThetaT));
    float Rparl = ((etaT * cosThetaI) - (etaI * cosThetaT)) / ((etaT * cosThetaI) + (etaI * cosThetaT));
    float Rperp = ((etaI * cosThetaI) - (etaT * cosThetaT)) / ((etaI * cosThetaI) + (etaT * cosThetaT));
    return (Rparl * Rparl + Rperp * Rperp) / 2;
}
GNX_LOBE_FN V3 fr_conductor(float cosThetaI, V3 etai, V3 etat, V3 k) {
    cosThetaI = clampf(cosThetaI, -1, 1);
    V3 eta = etat / etai, etak = k / etai;
    float cosThetaI2 = cosThetaI * cosThetaI;
    float sinThetaI2 = 1.f - cosThetaI2;
    V3 eta2 = eta * eta, etak2 = etak * etak;
    V3 t0 = eta2 - etak2 - V3(sinThetaI2);
    V3 a2plusb2 = vsqrt(t0 * t0 + 4 * eta2 * etak2);
    V3 t1 = a2plusb2 + V3(cosThetaI2);
    V3 a = vsqrt(0.5f * (a2plusb2 + t0));
    V3 t2 = (2.f * cosThetaI) * a;
    V3 Rs = (t1 - t2) / (t1 + t2);
    V3 t3 = cosThetaI2 * a2plusb2 + V3(sinThetaI2 * sinThetaI2);
    V3 t4 = t2 * sinThetaI2;
    V3 Rp = Rs * (t3 - t4) / (t3 + t4);
    return 0.5f * (Rp + Rs);
}

// ---- Disney helpers, materials/DisneyMaterial.cpp:28-48 ------------------------------------------------
GNX_D float schlick_weight(float cosTheta) {
    float m = clampf(1 - cosTheta, 0, 1);
    return (m * m) * (m * m) * m;
}
GNX_D float fr_schlick(float R0, float cosTheta) { return lerpf(schlick_weight(cosTheta), R0, 1); }
GNX_D V3 fr_schlick3(V3 R0, float cosTheta) {
    float w = schlick_weight(cosTheta);
    return (1 - w) * R0 + w * V3(1.f);
}
GNX_D float schlick_r0_from_eta(float eta) { return ((eta - 1) * (eta - 1)) / ((eta + 1) * (eta + 1)); }

GNX_D V3 lobe_fresnel(const Lobe &l, float cosI) {
    switch (l.fresnel) {
    case FR_DIELECTRIC: return V3(fr_dielectric(cosI, l.e0, l.e1));
    case FR_CONDUCTOR: return fr_conductor(fabsf(cosI), V3(l.e0), l.a, l.b);
    case FR_DISNEY:   // DisneyFresnel::Evaluate, materials/DisneyMaterial.cpp:308-327
        return (1 - l.e0) * V3(fr_dielectric(cosI, 1, l.e1)) + l.e0 * fr_schlick3(l.a, cosI);
    default: return V3(1.f);
    }
}

// ---- TrowbridgeReitzDistribution, core/MicroFacet.cpp:129-136,150-159,215-316 ----------------------
GNX_LOBE_FN float tr_D(V3 wh, float ax, float ay) {
    float tan2Theta = tan2_theta(wh);
    if (finf(tan2Theta)) return 0.f;
    const float cos4Theta = cos2_theta(wh) * cos2_theta(wh);
    float e = (cos2_phi(wh) / (ax * ax) + sin2_phi(wh) / (ay * ay)) * tan2Theta;
    return 1 / (kPi * ax * ay * cos4Theta * (1 + e) * (1 + e));
}
GNX_LOBE_FN float tr_lambda(V3 w, float ax, float ay) {
    float absTanTheta = fabsf(tan_theta(w));
    if (finf(absTanTheta)) return 0.f;
    float alpha = sqrtf(cos2_phi(w) * ax * ax + sin2_phi(w) * ay * ay);
    float alpha2Tan2Theta = (alpha * absTanTheta) * (alpha * absTanTheta);
    return (-1 + sqrtf(1.f + alpha2Tan2Theta)) / 2;
}
GNX_D float tr_G1(V3 w, float ax, float ay) { return 1 / (1 + tr_lambda(w, ax, ay)); }
GNX_D float distrib_G(const Lobe &l, V3 wo, V3 wi) {
    // DisneyMicrofacetDistribution::G is the separable product (DisneyMaterial.cpp:332-343)
    if (l.distrib == DK_DISNEY) return tr_G1(wo, l.p0, l.p1) * tr_G1(wi, l.p0, l.p1);
    return 1 / (1 + tr_lambda(wo, l.p0, l.p1) + tr_lambda(wi, l.p0, l.p1));
}
GNX_D void tr_sample11(float cosTheta, float U1, float U2, float *slope_x, float *slope_y) {
    if (cosTheta > .9999f) {
        float r = sqrtf(U1 / (1 - U1));
        float phi = (float)(6.28318530718 * (double)U2);
        *slope_x = r * cosf(phi);
        *slope_y = r * sinf(phi);
        return;
    }
    float sinTheta = sqrtf(fmaxf(0.f, 1.f - cosTheta * cosTheta));
    float tanTheta = sinTheta / cosTheta;
    float a = 1 / tanTheta;
    float G1 = 2 / (1 + sqrtf(1.f + 1.f / (a * a)));
    float A = 2 * U1 / G1 - 1;
    float tmp = 1.f / (A * A - 1.f);
    if (tmp > 1e10f) tmp = 1e10f;
    float B = tanTheta;
    float D = sqrtf(fmaxf(B * B * tmp * tmp - (A * A - B * B) * tmp, 0.f));
    float slope_x_1 = B * tmp - D;
    float slope_x_2 = B * tmp + D;
    *slope_x = (A < 0 || slope_x_2 > 1.f / tanTheta) ? slope_x_1 : slope_x_2;
    float S;
    if (U2 > 0.5f) { S = 1.f; U2 = 2.f * (U2 - .5f); } else { S = -1.f; U2 = 2.f * (.5f - U2); }
    float z = (U2 * (U2 * (U2 * 0.27385f - 0.73369f) + 0.46341f)) /
              (U2 * (U2 * (U2 * 0.093073f + 0.309420f) - 1.000000f) + 0.597999f);
    *slope_y = S * z * sqrtf(1.f + *slope_x * *slope_x);
}
GNX_LOBE_FN V3 tr_sample(V3 wi, float ax, float ay, float U1, float U2) {
    V3 wiS = normalize(V3(ax * wi.x, ay * wi.y, wi.z));
    float sx, sy;
    tr_sample11(cos_theta(wiS), U1, U2, &sx, &sy);
    float tmp = cos_phi(wiS) * sx - sin_phi(wiS) * sy;
    sy = sin_phi(wiS) * sx + cos_phi(wiS) * sy;
    sx = tmp;
    sx = ax * sx;
    sy = ay * sy;
    return normalize(V3(-sx, -sy, 1.f));
}
GNX_D V3 tr_sample_wh(V3 wo, float ax, float ay, float u0, float u1) {
    bool flip = wo.z < 0;
    V3 wh = tr_sample(flip ? -wo : wo, ax, ay, u0, u1);
    if (flip) wh = -wh;
    return wh;
}
GNX_D float tr_pdf(V3 wo, V3 wh, float ax, float ay) {  // sampleVisibleArea == true always (MicroFacet.h:79-83)
    return tr_D(wh, ax, ay) * tr_G1(wo, ax, ay) * absdot(wo, wh) / abs_cos_theta(wo);
}
GNX_D float roughness_to_alpha(float roughness) {  // core/MicroFacet.h:97-103
    roughness = fmaxf(roughness, 1e-3f);
    float x = logf(roughness);
    return 1.62142f + 0.819955f * x + 0.1734f * x * x + 0.0171201f * x * x * x + 0.000640711f * x * x * x * x;
}

// ---- Disney lobes, materials/DisneyMaterial.cpp:50-295 ---------------------------------------------------
GNX_D float gtr1(float cosTheta, float alpha) {
    float alpha2 = alpha * alpha;
    return (alpha2 - 1) / (kPi * logf(alpha2) * (1 + (alpha2 - 1) * cosTheta * cosTheta));
}
GNX_D float smith_g_ggx(float cosTheta, float alpha) {
    float alpha2 = alpha * alpha;
    float cosTheta2 = cosTheta * cosTheta;
    return 1 / (cosTheta + sqrtf(alpha2 + cosTheta2 - alpha2 * cosTheta2));
}

// ---- per-lobe f / pdf / sample ----------------------------------------------------------------------------
GNX_LOBE_FN V3 lobe_f(const Lobe &l, V3 wo, V3 wi) {
    switch (l.kind) {
    case LK_LAMBERT_R:
    case LK_LAMBERT_T:
        return l.R * kInvPi;
    case LK_OREN_NAYAR: {  // core/Reflection.cpp:173-198
        float sinThetaI = sin_theta(wi), sinThetaO = sin_theta(wo);
        float maxCos = 0;
        if (sinThetaI > 1e-4f && sinThetaO > 1e-4f) {
            float sinPhiI = sin_phi(wi), cosPhiI = cos_phi(wi);
            float sinPhiO = sin_phi(wo), cosPhiO = cos_phi(wo);
            float dCos = cosPhiI * cosPhiO + sinPhiI * sinPhiO;
            maxCos = fmaxf(0.f, dCos);
        }
        float sinAlpha, tanBeta;
        if (abs_cos_theta(wi) > abs_cos_theta(wo)) { sinAlpha = sinThetaO; tanBeta = sinThetaI / abs_cos_theta(wi); }
        else { sinAlpha = sinThetaI; tanBeta = sinThetaO / abs_cos_theta(wo); }
        return l.R * kInvPi * (l.p0 + l.p1 * maxCos * sinAlpha * tanBeta);
    }
    case LK_MICRO_R: {  // core/Reflection.cpp:223-240
        float cosThetaO = abs_cos_theta(wo), cosThetaI = abs_cos_theta(wi);
        V3 wh = wi + wo;
        if (cosThetaI == 0 || cosThetaO == 0) return V3(0.f);
        if (wh.x == 0 && wh.y == 0 && wh.z == 0) return V3(0.f);
        wh = normalize(wh);
        V3 F = lobe_fresnel(l, dot(wi, faceforward(wh, V3(0, 0, 1))));
        return div_each(l.R * tr_D(wh, l.p0, l.p1) * distrib_G(l, wo, wi) * F, 4 * cosThetaI * cosThetaO);
    }
    case LK_MICRO_T: {  // core/Reflection.cpp:275-302
        if (same_hemisphere(wo, wi)) return V3(0.f);
        float cosThetaO = cos_theta(wo), cosThetaI = cos_theta(wi);
        if (cosThetaI == 0 || cosThetaO == 0) return V3(0.f);
        float eta = cos_theta(wo) > 0 ? (l.e1 / l.e0) : (l.e0 / l.e1);
        V3 wh = normalize(wo + wi * eta);
        if (wh.z < 0) wh = -wh;
        if (dot(wo, wh) * dot(wi, wh) > 0) return V3(0.f);
        float F = fr_dielectric(dot(wo, wh), l.e0, l.e1);
        float sqrtDenom = dot(wo, wh) + eta * dot(wi, wh);
        float factor = 1 / eta;  // TransportMode::Radiance
        return (V3(1.f) - V3(F)) * l.R *
               fabsf(tr_D(wh, l.p0, l.p1) * distrib_G(l, wo, wi) * eta * eta * absdot(wi, wh) * absdot(wo, wh) *
                     factor * factor / (cosThetaI * cosThetaO * sqrtDenom * sqrtDenom));
    }
    case LK_DISNEY_DIFFUSE: {  // DisneyMaterial.cpp:64-72
        float Fo = schlick_weight(abs_cos_theta(wo)), Fi = schlick_weight(abs_cos_theta(wi));
        return l.R * kInvPi * (1 - Fo / 2) * (1 - Fi / 2);
    }
    case LK_DISNEY_FAKESS: {  // DisneyMaterial.cpp:105-122
        V3 wh = wi + wo;
        if (wh.x == 0 && wh.y == 0 && wh.z == 0) return V3(0.f);
        wh = normalize(wh);
        float cosThetaD = dot(wi, wh);
        float Fss90 = cosThetaD * cosThetaD * l.p0;
        float Fo = schlick_weight(abs_cos_theta(wo)), Fi = schlick_weight(abs_cos_theta(wi));
        float Fss = lerpf(Fo, 1.0f, Fss90) * lerpf(Fi, 1.0f, Fss90);
        float ss = 1.25f * (Fss * (1 / (abs_cos_theta(wo) + abs_cos_theta(wi)) - .5f) + .5f);
        return l.R * kInvPi * ss;
    }
    case LK_DISNEY_RETRO: {  // DisneyMaterial.cpp:151-164
        V3 wh = wi + wo;
        if (wh.x == 0 && wh.y == 0 && wh.z == 0) return V3(0.f);
        wh = normalize(wh);
        float cosThetaD = dot(wi, wh);
        float Fo = schlick_weight(abs_cos_theta(wo)), Fi = schlick_weight(abs_cos_theta(wi));
        float Rr = 2 * l.p0 * cosThetaD * cosThetaD;
        return l.R * kInvPi * Rr * (Fo + Fi + Fo * Fi * (Rr - 1));
    }
    case LK_DISNEY_SHEEN: {  // DisneyMaterial.cpp:189-197
        V3 wh = wi + wo;
        if (wh.x == 0 && wh.y == 0 && wh.z == 0) return V3(0.f);
        wh = normalize(wh);
        float cosThetaD = dot(wi, wh);
        return l.R * schlick_weight(cosThetaD);
    }
    case LK_DISNEY_CLEARCOAT: {  // DisneyMaterial.cpp:239-258
        V3 wh = wi + wo;
        if (wh.x == 0 && wh.y == 0 && wh.z == 0) return V3(0.f);
        wh = normalize(wh);
        float Dr = gtr1(abs_cos_theta(wh), l.p1);
        float Fr = fr_schlick(.04f, dot(wo, wh));
        float Gr = smith_g_ggx(abs_cos_theta(wo), .25f) * smith_g_ggx(abs_cos_theta(wi), .25f);
        return V3(l.p0 * Gr * Fr * Dr / 4);
    }
    default:
        return V3(0.f);  // specular lobes: f == 0
    }
}

GNX_LOBE_FN float lobe_pdf(const Lobe &l, V3 wo, V3 wi) {
    switch (l.kind) {
    case LK_LAMBERT_R:
    case LK_OREN_NAYAR:
    case LK_DISNEY_DIFFUSE:
    case LK_DISNEY_FAKESS:
    case LK_DISNEY_RETRO:
    case LK_DISNEY_SHEEN:
        return same_hemisphere(wo, wi) ? abs_cos_theta(wi) * kInvPi : 0;  // BxDF::Pdf
    case LK_LAMBERT_T:
        return !same_hemisphere(wo, wi) ? abs_cos_theta(wi) * kInvPi : 0;
    case LK_MICRO_R: {
        if (!same_hemisphere(wo, wi)) return 0;
        V3 wh = normalize(wo + wi);
        return tr_pdf(wo, wh, l.p0, l.p1) / (4 * dot(wo, wh));
    }
    case LK_MICRO_T: {
        if (same_hemisphere(wo, wi)) return 0;
        float eta = cos_theta(wo) > 0 ? (l.e1 / l.e0) : (l.e0 / l.e1);
        V3 wh = normalize(wo + wi * eta);
        if (dot(wo, wh) * dot(wi, wh) > 0) return 0;
        float sqrtDenom = dot(wo, wh) + eta * dot(wi, wh);
        float dwh_dwi = fabsf((eta * eta * dot(wi, wh)) / (sqrtDenom * sqrtDenom));
        return tr_pdf(wo, wh, l.p0, l.p1) * dwh_dwi;
    }
    case LK_DISNEY_CLEARCOAT: {  // DisneyMaterial.cpp:285-295
        if (!same_hemisphere(wo, wi)) return 0;
        V3 wh = wi + wo;
        if (wh.x == 0 && wh.y == 0 && wh.z == 0) return 0;
        wh = normalize(wh);
        float Dr = gtr1(abs_cos_theta(wh), l.p1);
        return Dr * abs_cos_theta(wh) / (4 * dot(wo, wh));
    }
    default:
        return 0;
    }
}

// *pdf == 0 means "no sample".  *sampledType is preset to l.type by the caller.  The return value is f for the
// SPECULAR lobes only: for every other lobe BSDF::Sample_f discards the sampled lobe's f and re-evaluates the
// sum over all matching lobes (core/Reflection.cpp:545-556), so it is not computed here (it was 10 % of the
// shade kernel, at half-empty warps).
GNX_LOBE_FN V3 lobe_sample(const Lobe &l, V3 wo, float u0, float u1, V3 *wi, float *pdf, int *sampledType) {
    switch (l.kind) {
    case LK_SPEC_R: {  // core/Reflection.cpp:89-97
        *wi = V3(-wo.x, -wo.y, wo.z);
        *pdf = 1;
        return div_each(lobe_fresnel(l, cos_theta(*wi)) * l.R, abs_cos_theta(*wi));
    }
    case LK_SPEC_T: {  // core/Reflection.cpp:105-122
        bool entering = cos_theta(wo) > 0;
        float etaI = entering ? l.e0 : l.e1, etaT = entering ? l.e1 : l.e0;
        if (!refract(wo, faceforward(V3(0, 0, 1), wo), etaI / etaT, wi)) return V3(0.f);
        *pdf = 1;
        V3 ft = l.R * (V3(1.f) - V3(fr_dielectric(cos_theta(*wi), l.e0, l.e1)));
        ft *= (etaI * etaI) / (etaT * etaT);
        return div_each(ft, abs_cos_theta(*wi));
    }
    case LK_FRESNEL_SPEC: {  // core/Reflection.cpp:346-380
        float F = fr_dielectric(cos_theta(wo), l.e0, l.e1);
        if (u0 < F) {
            *wi = V3(-wo.x, -wo.y, wo.z);
            *sampledType = BSDF_SPECULAR | BSDF_REFLECTION;
            *pdf = F;
            return div_each(F * l.R, abs_cos_theta(*wi));
        } else {
            bool entering = cos_theta(wo) > 0;
            float etaI = entering ? l.e0 : l.e1, etaT = entering ? l.e1 : l.e0;
            if (!refract(wo, faceforward(V3(0, 0, 1), wo), etaI / etaT, wi)) return V3(0.f);
            V3 ft = l.a * (1 - F);
            ft *= (etaI * etaI) / (etaT * etaT);
            *sampledType = BSDF_SPECULAR | BSDF_TRANSMISSION;
            *pdf = 1 - F;
            return div_each(ft, abs_cos_theta(*wi));
        }
    }
    case LK_MICRO_R: {  // core/Reflection.cpp:206-221
        if (wo.z == 0) return V3(0.f);
        V3 wh = tr_sample_wh(wo, l.p0, l.p1, u0, u1);
        if (dot(wo, wh) < 0) return V3(0.f);
        *wi = reflect(wo, wh);
        if (!same_hemisphere(wo, *wi)) return V3(0.f);
        *pdf = tr_pdf(wo, wh, l.p0, l.p1) / (4 * dot(wo, wh));
        return V3(0.f);
    }
    case LK_MICRO_T: {  // core/Reflection.cpp:249-258
        if (wo.z == 0) return V3(0.f);
        V3 wh = tr_sample_wh(wo, l.p0, l.p1, u0, u1);
        if (dot(wo, wh) < 0) return V3(0.f);
        float eta = cos_theta(wo) > 0 ? (l.e0 / l.e1) : (l.e1 / l.e0);
        if (!refract(wo, wh, eta, wi)) return V3(0.f);
        *pdf = lobe_pdf(l, wo, *wi);
        return V3(0.f);
    }
    case LK_LAMBERT_T: {  // core/Reflection.cpp:146-154
        *wi = cosine_sample_hemisphere(u0, u1);
        if (wo.z > 0) wi->z *= -1;
        *pdf = lobe_pdf(l, wo, *wi);
        return V3(0.f);
    }
    case LK_DISNEY_CLEARCOAT: {  // DisneyMaterial.cpp:260-283
        if (wo.z == 0) return V3(0.f);
        float alpha2 = l.p1 * l.p1;
        float cosTheta = sqrtf(fmaxf(0.f, (1 - powf(alpha2, 1 - u0)) / (1 - alpha2)));
        float sinTheta = sqrtf(fmaxf(0.f, 1 - cosTheta * cosTheta));
        float phi = 2 * kPi * u1;
        V3 wh(sinTheta * cosf(phi), sinTheta * sinf(phi), cosTheta);
        if (!same_hemisphere(wo, wh)) wh = -wh;
        *wi = reflect(wo, wh);
        if (!same_hemisphere(wo, *wi)) return V3(0.f);
        *pdf = lobe_pdf(l, wo, *wi);
        return V3(0.f);
    }
    default: {  // BxDF::Sample_f, core/Reflection.cpp:394-402
        *wi = cosine_sample_hemisphere(u0, u1);
        if (wo.z < 0) wi->z *= -1;
        *pdf = lobe_pdf(l, wo, *wi);
        return V3(0.f);
    }
    }
}

// f and pdf of one lobe for the same pair of directions.  The microfacet reflection lobe shares the half vector,
// D(wh) and Lambda(wo) between the two (identical inputs, hence identical bits); every other lobe is the two
// separate evaluations.
GNX_D void lobe_eval(const Lobe &l, V3 wo, V3 wi, bool needF, bool needPdf, V3 *f, float *pdf) {
    *f = V3(0.f);
    *pdf = 0.f;
    if (l.kind == LK_MICRO_R) {
        const float cosThetaO = abs_cos_theta(wo), cosThetaI = abs_cos_theta(wi);
        const V3 whs = wi + wo;
        const bool fOk = needF && !(cosThetaI == 0 || cosThetaO == 0) && !(whs.x == 0 && whs.y == 0 && whs.z == 0);
        const bool pdfOk = needPdf && same_hemisphere(wo, wi);
        if (!(fOk || pdfOk)) return;
        const V3 wh = normalize(whs);
        const float D = tr_D(wh, l.p0, l.p1);
        const float lambdaO = tr_lambda(wo, l.p0, l.p1);
        if (pdfOk) *pdf = (D * (1 / (1 + lambdaO)) * absdot(wo, wh) / abs_cos_theta(wo)) / (4 * dot(wo, wh));
        if (fOk) {
            const V3 F = lobe_fresnel(l, dot(wi, faceforward(wh, V3(0, 0, 1))));
            const float lambdaI = tr_lambda(wi, l.p0, l.p1);
            const float G = l.distrib == DK_DISNEY ? (1 / (1 + lambdaO)) * (1 / (1 + lambdaI)) : 1 / (1 + lambdaO + lambdaI);
            *f = div_each(l.R * D * G * F, 4 * cosThetaI * cosThetaO);
        }
        return;
    }
    if (needF) *f = lobe_f(l, wo, wi);
    if (needPdf) *pdf = lobe_pdf(l, wo, wi);
}

// ---- BSDF::f / Pdf / Sample_f, core/Reflection.cpp:440-563 -----------------------------------------------------
template <int MAXL>
GNX_D V3 bsdf_f(const Bsdf<MAXL> &b, V3 woW, V3 wiW, int flags) {
    V3 wi = b.to_local(wiW), wo = b.to_local(woW);
    if (wo.z == 0) return V3(0.f);
    bool refl = dot(wiW, b.ng) * dot(woW, b.ng) > 0;
    V3 f(0.f);
    for (int i = 0; i < b.n; ++i) {
        const Lobe &l = b.lobes[i];
        if ((l.type & flags) == l.type &&
            ((refl && (l.type & BSDF_REFLECTION)) || (!refl && (l.type & BSDF_TRANSMISSION))))
            f += lobe_f(l, wo, wi);
    }
    return f;
}

template <int MAXL>
GNX_D float bsdf_pdf(const Bsdf<MAXL> &b, V3 woW, V3 wiW, int flags) {
    if (b.n == 0) return 0.f;
    V3 wo = b.to_local(woW), wi = b.to_local(wiW);
    if (wo.z == 0) return 0.f;
    float pdf = 0.f;
    int matching = 0;
    for (int i = 0; i < b.n; ++i)
        if ((b.lobes[i].type & flags) == b.lobes[i].type) { ++matching; pdf += lobe_pdf(b.lobes[i], wo, wi); }
    return matching > 0 ? pdf / matching : 0.f;
}

// The two halves of BSDF::Sample_f (core/Reflection.cpp:500-563), separately callable so that the shading kernel
// can run one copy of each for its three uses (light-sample evaluation, MIS sample, continuation sample).
//
// bsdf_sample_dir: lobe choice and the chosen lobe's own Sample_f.  wo / *wi are in the shading frame.  Returns
// false when there is no sample (*pdf == 0).  *fSpec is the sampled lobe's f, meaningful for SPECULAR lobes only.
template <int MAXL>
GNX_D bool bsdf_sample_dir(const Bsdf<MAXL> &b, V3 wo, float u0, float u1, int flags, int matching, int *chosenOut, V3 *wi,
                           float *pdf, int *sampledType, V3 *fSpec) {
    *pdf = 0;
    *sampledType = 0;
    *chosenOut = -1;
    *fSpec = V3(0.f);
    if (matching == 0) return false;
    int comp = (int)floorf(u0 * matching);
    if (comp > matching - 1) comp = matching - 1;
    int chosen = -1, count = comp;
    for (int i = 0; i < b.n; ++i)
        if ((b.lobes[i].type & flags) == b.lobes[i].type && count-- == 0) { chosen = i; break; }
    const Lobe &l = b.lobes[chosen];
    float ur0 = fminf(u0 * matching - comp, kOneMinusEpsilon);
    if (wo.z == 0) return false;
    *sampledType = l.type;
    *fSpec = lobe_sample(l, wo, ur0, u1, wi, pdf, sampledType);
    if (*pdf == 0) { *sampledType = 0; return false; }
    *chosenOut = chosen;
    return true;
}
// bsdf_eval_sum: f summed over the matching lobes on the side of the geometric normal that (wo, wi) are on, and
// the pdf of every matching lobe except `chosen` added to *pdfAcc — both in lobe order, as the reference adds them.
template <int MAXL>
GNX_D void bsdf_eval_sum(const Bsdf<MAXL> &b, V3 wo, V3 wi, bool refl, int flags, int chosen, V3 *fOut, float *pdfAcc) {
    V3 f(0.f);
    float pdf = *pdfAcc;
    for (int i = 0; i < b.n; ++i) {
        const Lobe &l = b.lobes[i];
        if ((l.type & flags) != l.type) continue;
        const bool needF = (refl && (l.type & BSDF_REFLECTION)) || (!refl && (l.type & BSDF_TRANSMISSION));
        const bool needPdf = i != chosen;
        if (!(needF || needPdf)) continue;
        V3 fi;
        float pi;
        lobe_eval(l, wo, wi, needF, needPdf, &fi, &pi);
        if (needF) f += fi;
        if (needPdf) pdf += pi;
    }
    *fOut = f;
    *pdfAcc = pdf;
}

// BSDF::f and BSDF::Pdf for the same directions (EstimateDirect's light-sampling half calls both)
template <int MAXL>
GNX_D void bsdf_f_pdf(const Bsdf<MAXL> &b, V3 woW, V3 wiW, int flags, V3 *fOut, float *pdfOut) {
    *fOut = V3(0.f);
    *pdfOut = 0.f;
    V3 wi = b.to_local(wiW), wo = b.to_local(woW);
    if (wo.z == 0 || b.n == 0) return;
    const int matching = b.num_components(flags);
    const bool refl = dot(wiW, b.ng) * dot(woW, b.ng) > 0;
    float pdf = 0.f;
    bsdf_eval_sum(b, wo, wi, refl, flags, -1, fOut, &pdf);
    *pdfOut = matching > 0 ? pdf / matching : 0.f;
}

template <int MAXL>
GNX_D V3 bsdf_sample(const Bsdf<MAXL> &b, V3 woW, V3 *wiW, float u0, float u1, float *pdf, int flags, int *sampledType) {
    const int matching = b.num_components(flags);
    const V3 wo = b.to_local(woW);
    int chosen;
    V3 wi, f;
    if (!bsdf_sample_dir(b, wo, u0, u1, flags, matching, &chosen, &wi, pdf, sampledType, &f)) { *pdf = 0; return V3(0.f); }
    *wiW = b.to_world(wi);
    if (!(b.lobes[chosen].type & BSDF_SPECULAR)) {
        const bool refl = dot(*wiW, b.ng) * dot(woW, b.ng) > 0;
        bsdf_eval_sum(b, wo, wi, refl, flags, chosen, &f, pdf);
    }
    *pdf /= matching;  // "if (matchingComps > 1) *pdf /= matchingComps": x / 1 == x
    return f;
}

}  // namespace gnx
