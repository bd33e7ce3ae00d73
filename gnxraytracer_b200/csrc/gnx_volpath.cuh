// gnx_volpath.cuh — VolPathIntegrator::Li (integrators/VolPathIntegrator.cpp:24-159) with participating
// media: HomogeneousMedium (media/HomogeneousMedium.cpp:11-43), GridDensityMedium delta / ratio tracking
// (media/GridDensityMedium.cpp:14-87), Henyey-Greenstein (core/Medium.cpp:164-188, core/Medium.h:34-38),
// VisibilityTester::Tr (core/Light.cpp:33-53), Scene::IntersectTr (core/Scene.cpp:26-40) and
// EstimateDirect with handleMedia = true (core/Integrator.cpp:93-210).
//
// Unlike PathIntegrator, the number of sampler draws between two decisions of a path is data dependent
// here (tracking loops inside the shadow-ray transmittance come BEFORE the draws of the next direction), so
// one path is carried from camera to termination by one lane ("one path per lane" kernel, k_volpath)
// instead of being cut into wavefront stages.  Everything is __host__ __device__ like gnx_path.cuh.
#pragma once
#include "gnx_path.cuh"

namespace gnx {

constexpr float kMaxFloat = 3.402823466e+38f;

struct VRay {
    V3 o, d;
    float tMax;
    int medium;  // -1 = vacuum
};

// What SpawnRay / SpawnRayTo / GetMedium need from an Interaction (core/Interaction.h:33-71)
struct VPoint {
    V3 p, pError, n;   // n == 0 for medium interactions and for the far point of an infinite light
    int mIn, mOut;     // MediumInterface
};
GNX_D int medium_toward(const VPoint &it, V3 w) { return dot(w, it.n) > 0 ? it.mOut : it.mIn; }
GNX_D VRay spawn_ray(const VPoint &it, V3 d) {
    VRay r;
    r.o = offset_ray_origin(it.p, it.pError, it.n, d);
    r.d = d;
    r.tMax = GNX_INF;
    r.medium = medium_toward(it, d);
    return r;
}
GNX_D VRay spawn_ray_to(const VPoint &it, const VPoint &to) {
    VRay r;
    r.o = offset_ray_origin(it.p, it.pError, it.n, to.p - it.p);
    V3 target = offset_ray_origin(to.p, to.pError, to.n, r.o - to.p);
    r.d = target - r.o;
    r.tMax = 1 - kShadowEpsilon;
    r.medium = medium_toward(it, r.d);
    return r;
}

// Transform::operator()(Ray), core/Transform.h:230-244
GNX_D void xform_ray(const M44 &M, V3 o, V3 d, float tMax, V3 *oOut, V3 *dOut, float *tMaxOut) {
    V3 oErr;
    V3 no = xform_point_err(M, o, &oErr);
    V3 nd = xform_vector(M, d);
    float lsq = length_sq(nd);
    if (lsq > 0) {
        float dt = dot(vabs(nd), oErr) / lsq;
        no = no + nd * dt;
        tMax -= dt;
    }
    *oOut = no; *dOut = nd; *tMaxOut = tMax;
}

// Bounds3f((0,0,0),(1,1,1)).IntersectP(ray, &t0, &t1), core/Geometry.h:1356-1377
GNX_D bool unit_box_interval(V3 o, V3 d, float tMax, float *hit0, float *hit1) {
    float t0 = 0, t1 = tMax;
    for (int i = 0; i < 3; ++i) {
        float invRayDir = 1 / d[i];
        float tNear = (0.f - o[i]) * invRayDir, tFar = (1.f - o[i]) * invRayDir;
        if (tNear > tFar) { float t = tNear; tNear = tFar; tFar = t; }
        tFar *= 1 + 2 * gamma_n(3);
        t0 = tNear > t0 ? tNear : t0;
        t1 = tFar < t1 ? tFar : t1;
        if (t0 > t1) return false;
    }
    *hit0 = t0; *hit1 = t1;
    return true;
}

// GridDensityMedium::D / Density, media/GridDensityMedium.h:45-49, media/GridDensityMedium.cpp:14-29
GNX_D float grid_D(const DevMedium &m, int x, int y, int z) {
    if (x < 0 || y < 0 || z < 0 || x >= m.nx || y >= m.ny || z >= m.nz) return 0;
    return ldg(m.density + ((size_t)z * m.ny + y) * m.nx + x);
}
GNX_D float grid_density(const DevMedium &m, V3 p) {
    float sx = p.x * m.nx - .5f, sy = p.y * m.ny - .5f, sz = p.z * m.nz - .5f;
    int ix = (int)floorf(sx), iy = (int)floorf(sy), iz = (int)floorf(sz);
    float dx = sx - ix, dy = sy - iy, dz = sz - iz;
    float d00 = lerpf(dx, grid_D(m, ix, iy, iz), grid_D(m, ix + 1, iy, iz));
    float d10 = lerpf(dx, grid_D(m, ix, iy + 1, iz), grid_D(m, ix + 1, iy + 1, iz));
    float d01 = lerpf(dx, grid_D(m, ix, iy, iz + 1), grid_D(m, ix + 1, iy, iz + 1));
    float d11 = lerpf(dx, grid_D(m, ix, iy + 1, iz + 1), grid_D(m, ix + 1, iy + 1, iz + 1));
    float d0 = lerpf(dy, d00, d10), d1 = lerpf(dy, d01, d11);
    return lerpf(dz, d0, d1);
}

GNX_D V3 vexp(V3 a) { return V3(expf(a.x), expf(a.y), expf(a.z)); }

// ---- GridDensityMedium tracking loops, one step at a time ------------------------------------------------------
// Delta tracking (GridDensityMedium::Sample, media/GridDensityMedium.cpp:31-55) and ratio tracking
// (GridDensityMedium::Tr, :57-87) share the walk: t -= log(1 - u) * invMaxDensity / sigma_t, density lookup at
// ray(t).  Written as begin / step so that the wavefront tracking kernel (k_vp_track) can run one step per lane and
// refill lanes whose walk has ended; the sequential callers below loop over the same functions.
struct TrackState {
    V3 o, d;          // the ray in medium space (direction normalised in world space before the transform)
    float t, tMax;
    float Tr;         // ratio tracking: transmittance so far
    int mode;         // 0 = delta tracking (sample an interaction), 1 = ratio tracking (transmittance)
    bool sampled;     // delta tracking: an interaction was sampled at parameter t
};
// false: the ray misses the medium's unit box (no interaction / Tr = 1, no sampler draws)
GNX_D bool track_begin(const DevMedium &m, V3 ro, V3 rd, float rtMax, int mode, TrackState &ts) {
    float tMaxM, tMin, tMax;
    xform_ray(m.w2m, ro, normalize(rd), rtMax * length(rd), &ts.o, &ts.d, &tMaxM);
    ts.mode = mode; ts.Tr = 1; ts.sampled = false;
    if (!unit_box_interval(ts.o, ts.d, tMaxM, &tMin, &tMax)) return false;
    ts.t = tMin; ts.tMax = tMax;
    return true;
}
// One tracking step; true when the walk has ended (ts.sampled / ts.t, or ts.Tr, hold the result).
GNX_D bool track_step(const DevMedium &m, TrackState &ts, PathSampler &smp) {
    ts.t -= logf(1 - smp.get1d()) * m.inv_max_density / m.sigma_t_scalar;
    if (ts.t >= ts.tMax) return true;
    const float density = grid_density(m, ts.o + ts.d * ts.t);
    if (ts.mode == 0) {
        if (density * m.inv_max_density > smp.get1d()) { ts.sampled = true; return true; }
        return false;
    }
    ts.Tr *= 1 - fmaxf(0.f, density * m.inv_max_density);
    const float rrThreshold = .1f;
    if (ts.Tr < rrThreshold) {
        float q = fmaxf(.05f, 1 - ts.Tr);
        if (smp.get1d() < q) { ts.Tr = 0.f; return true; }
        ts.Tr /= 1 - q;
    }
    return false;
}

// HomogeneousMedium::Tr (media/HomogeneousMedium.cpp:11-15): closed form, no sampler draws
GNX_D V3 homogeneous_tr(const DevMedium &m, V3 rd, float rtMax) {
    const V3 st(m.sigma_t[0], m.sigma_t[1], m.sigma_t[2]);
    return vexp(-st * fminf(rtMax * length(rd), kMaxFloat));
}
// HomogeneousMedium::Sample (media/HomogeneousMedium.cpp:17-43): the weight; *tOut >= 0 when an interaction was sampled
// at ray(t) (t in the ray's own parametrisation), -1 otherwise.  Two sampler draws.
GNX_D V3 homogeneous_sample(const DevMedium &m, V3 rd, float rtMax, PathSampler &smp, float *tOut) {
    const V3 st(m.sigma_t[0], m.sigma_t[1], m.sigma_t[2]), ss(m.sigma_s[0], m.sigma_s[1], m.sigma_s[2]);
    int channel = (int)(smp.get1d() * 3);
    if (channel > 2) channel = 2;
    float dist = -logf(1 - smp.get1d()) / st[channel];
    float len = length(rd);
    float t = fminf(dist / len, rtMax);
    bool sampledMedium = t < rtMax;
    *tOut = sampledMedium ? t : -1.f;
    V3 Tr = vexp(-st * fminf(t, kMaxFloat) * len);
    V3 density = sampledMedium ? (st * Tr) : Tr;
    float pdf = 0;
    for (int i = 0; i < 3; ++i) pdf += density[i];
    pdf *= 1 / (float)3;
    if (pdf == 0) pdf = 1;
    return sampledMedium ? div_each(Tr * ss, pdf) : div_each(Tr, pdf);
}

// Medium::Tr
GNX_D V3 medium_tr(const DevMedium &m, const VRay &ray, PathSampler &smp) {
    if (m.type == GNX_MEDIUM_HOMOGENEOUS) return homogeneous_tr(m, ray.d, ray.tMax);
    TrackState ts;
    if (!track_begin(m, ray.o, ray.d, ray.tMax, 1, ts)) return V3(1.f);
    while (!track_step(m, ts, smp)) {}
    return V3(ts.Tr);
}

// Medium::Sample; *sampled / *pMi describe the MediumInteraction when one is created
GNX_D V3 medium_sample(const DevMedium &m, const VRay &ray, PathSampler &smp, bool *sampled, V3 *pMi) {
    *sampled = false;
    if (m.type == GNX_MEDIUM_HOMOGENEOUS) {
        float t;
        V3 w = homogeneous_sample(m, ray.d, ray.tMax, smp, &t);
        if (t >= 0) { *sampled = true; *pMi = ray.o + ray.d * t; }
        return w;
    }
    TrackState ts;
    if (!track_begin(m, ray.o, ray.d, ray.tMax, 0, ts)) return V3(1.f);
    while (!track_step(m, ts, smp)) {}
    if (ts.sampled) {
        *sampled = true;
        *pMi = ray.o + ray.d * ts.t;  // rWorld(t), as the reference writes it
        return div_each(V3(m.sigma_s[0], m.sigma_s[1], m.sigma_s[2]), m.sigma_t_scalar);
    }
    return V3(1.f);
}

// Henyey-Greenstein
GNX_D float phase_hg(float cosTheta, float g) {
    float denom = 1 + g * g + 2 * g * cosTheta;
    return kInv4Pi * (1 - g * g) / (denom * sqrtf(denom));
}
GNX_D float hg_sample_p(V3 wo, V3 *wi, float u0, float u1, float g) {
    float cosTheta;
    if (fabsf(g) < 1e-3f) cosTheta = 1 - 2 * u0;
    else {
        float sqrTerm = (1 - g * g) / (1 + g - 2 * g * u0);
        cosTheta = -(1 + g * g - sqrTerm * sqrTerm) / (2 * g);
    }
    float sinTheta = sqrtf(fmaxf(0.f, 1 - cosTheta * cosTheta));
    float phi = 2 * kPi * u1;
    V3 v1, v2;
    coordinate_system(wo, &v1, &v2);
    *wi = sinTheta * cosf(phi) * v1 + sinTheta * sinf(phi) * v2 + cosTheta * wo;  // SphericalDirection(.., x, y, z)
    return phase_hg(cosTheta, g);
}

struct VHit { int prim; TriHit h; };

GNX_D bool vol_intersect(const DeviceScene &sc, VRay &ray, VHit *hit, int2 *stack, int stride, TraversalCounters &cnt) {
    bool found = closest_hit(sc, ray.o, ray.d, ray.tMax, stack, stride, &hit->prim, &hit->h, cnt);
    if (found) ray.tMax = hit->h.t;
    return found;
}
// GeometricPrimitive::Intersect's medium interface rule (core/Primitive.cpp:41-44)
GNX_D VPoint surface_point(const DeviceScene &sc, const Surface &s, int rayMedium) {
    VPoint v;
    v.p = s.p; v.pError = s.pError; v.n = s.n;
    v.mIn = v.mOut = rayMedium;
    if (sc.tri_media && sc.tri_transition && sc.tri_transition[s.prim]) { int2 m = sc.tri_media[s.prim]; v.mIn = m.x; v.mOut = m.y; }
    return v;
}

// VisibilityTester::Tr, core/Light.cpp:33-53
GNX_D V3 visibility_tr(const DeviceScene &sc, const VPoint &p0, const VPoint &p1, PathSampler &smp, int2 *stack, int stride,
                       TraversalCounters &cnt, unsigned &rays) {
    VRay ray = spawn_ray_to(p0, p1);
    V3 Tr(1.f);
    while (true) {
        VHit hit;
        ++rays;
        bool hitSurface = vol_intersect(sc, ray, &hit, stack, stride, cnt);
        Surface s;
        if (hitSurface) {
            s = make_surface(sc, hit.prim, hit.h.b0, hit.h.b1, hit.h.b2, ray.d);
            if (s.material >= 0) return V3(0.f);
        }
        if (ray.medium >= 0) Tr *= medium_tr(sc.media[ray.medium], ray, smp);
        if (!hitSurface) break;
        ray = spawn_ray_to(surface_point(sc, s, ray.medium), p1);
    }
    return Tr;
}

// Scene::IntersectTr, core/Scene.cpp:26-40
GNX_D bool intersect_tr(const DeviceScene &sc, VRay ray, PathSampler &smp, Surface *sOut, V3 *Tr, int2 *stack, int stride,
                        TraversalCounters &cnt, unsigned &rays) {
    *Tr = V3(1.f);
    while (true) {
        VHit hit;
        ++rays;
        bool hitSurface = vol_intersect(sc, ray, &hit, stack, stride, cnt);
        if (ray.medium >= 0) *Tr *= medium_tr(sc.media[ray.medium], ray, smp);
        if (!hitSurface) return false;
        *sOut = make_surface(sc, hit.prim, hit.h.b0, hit.h.b1, hit.h.b2, ray.d);
        if (sOut->material >= 0) return true;
        ray = spawn_ray(surface_point(sc, *sOut, ray.medium), ray.d);
    }
}

// UniformSampleOneLight + EstimateDirect with handleMedia = true, for a surface (bsdf != null) or a
// medium interaction (bsdf == null; phase function HG(g), wo = -ray.d as MediumInteraction stores it).
template <int MAXL>
GNX_D V3 vol_sample_one_light(const DeviceScene &sc, const VPoint &it, const Bsdf<MAXL> *bsdf, V3 woSurf, V3 woMedium, float g,
                              PathSampler &smp, int2 *stack, int stride, TraversalCounters &cnt, unsigned &raysShadow,
                              unsigned &raysMis) {
    const int kNonSpec = BSDF_ALL & ~BSDF_SPECULAR;
    if (sc.n_lights == 0) return V3(0.f);
    float selPdf;
    const int lightNum = choose_light(sc, it.p, smp.get1d(), &selPdf);
    if (selPdf == 0) return V3(0.f);
    float ul0, ul1, us0, us1;
    smp.get2d(&ul0, &ul1);
    smp.get2d(&us0, &us1);
    const gnx_light &light = sc.lights[lightNum];
    const bool isEnv = light.type == GNX_LIGHT_INFINITE;
    V3 Ld(0.f);
    LightSample ls;
    bool ok = isEnv ? env_sample_li(sc.env, ul0, ul1, &ls) : area_sample_li(sc, light, it.p, ul0, ul1, &ls);
    if (ok && ls.pdf > 0 && !is_black(ls.Li)) {
        V3 f;
        float scatteringPdf;
        if (bsdf) {
            bsdf_f_pdf(*bsdf, woSurf, ls.wi, kNonSpec, &f, &scatteringPdf);
            f = f * absdot(ls.wi, bsdf->ns);
        } else {
            float p = phase_hg(dot(woMedium, ls.wi), g);
            f = V3(p);
            scatteringPdf = p;
        }
        if (!is_black(f)) {
            VPoint p1;
            if (isEnv) { p1.p = it.p + ls.wi * (2 * sc.env.world_radius); p1.pError = V3(0.f); p1.n = V3(0.f); }
            else { p1.p = ls.pl; p1.pError = ls.plError; p1.n = ls.nl; }
            p1.mIn = p1.mOut = light.medium;
            V3 Li = ls.Li * visibility_tr(sc, it, p1, smp, stack, stride, cnt, raysShadow);
            if (!is_black(Li)) {
                float weight = (ls.pdf * ls.pdf) / (ls.pdf * ls.pdf + scatteringPdf * scatteringPdf);
                Ld += div_each(f * Li * weight, ls.pdf);
            }
        }
    }
    {
        V3 wi, f;
        float scatteringPdf;
        if (bsdf) {
            int sampledType;
            f = bsdf_sample(*bsdf, woSurf, &wi, us0, us1, &scatteringPdf, kNonSpec, &sampledType);
            f = f * absdot(wi, bsdf->ns);
        } else {
            float p = hg_sample_p(woMedium, &wi, us0, us1, g);
            f = V3(p);
            scatteringPdf = p;
        }
        if (!is_black(f) && scatteringPdf > 0) {
            VRay ray = spawn_ray(it, wi);
            float lightPdf = isEnv ? env_pdf_li(sc.env, wi) : area_pdf_li(sc, light, it.p, ray.o, wi);
            if (lightPdf == 0) return div_each(Ld, selPdf);
            float weight = (scatteringPdf * scatteringPdf) / (scatteringPdf * scatteringPdf + lightPdf * lightPdf);
            Surface ls2;
            V3 Tr;
            bool found = intersect_tr(sc, ray, smp, &ls2, &Tr, stack, stride, cnt, raysMis);
            V3 Li(0.f);
            if (found) { if (!isEnv && ls2.prim == light.prim) Li = area_light_L(light, ls2.n, -wi); }
            else if (isEnv) Li = env_Le(sc.env, ray.d);
            if (!is_black(Li)) Ld += div_each(f * Li * Tr * weight, scatteringPdf);
        }
    }
    return div_each(Ld, selPdf);
}

struct VolCounters { unsigned extend, shadow, mis; };

// VolPathIntegrator::Li for camera sample `sample` of pixel (px, py).
template <bool TEX = false>
GNX_D V3 volpath_li(const DeviceScene &sc, const RenderConsts &rc, int px, int py, int sample, int2 *stack, int stride,
                    TraversalCounters &cnt, VolCounters &vc) {
    const bool pcg = sc.smp.type == GNX_SAMPLER_PCG32;
    PathSampler smp = pcg ? PathSampler::stream(sc.smp, ((uint64_t)(rc.width * py + px) << 20) | (uint64_t)sample)
                          : PathSampler(sc.smp, sampler_index(sc.smp, px, py, (uint64_t)sample), 0);
    // Sampler::GetCameraSample: film (2), time (1), lens (2)
    float u0, u1, tm, l0, l1;
    smp.get_film(px, py, &u0, &u1);
    tm = smp.get1d();
    smp.get2d(&l0, &l1);
    (void)tm;
    VRay ray;
    camera_ray_uv(sc, px, py, u0, u1, l0, l1, &ray.o, &ray.d, &ray.tMax);
    ray.medium = sc.cam.medium;
    V3 L(0.f), beta(1.f);
    bool specularBounce = false;
    float etaScale = 1;
    RayDiff camDiff;
    camDiff.has = false;
    if (TEX) camDiff = camera_ray_differentials(sc, px, py, u0, u1, l0, l1, ray.o, ray.d);
    for (int bounces = 0;; ++bounces) {
        VHit hit;
        ++vc.extend;
        const bool found = vol_intersect(sc, ray, &hit, stack, stride, cnt);
        bool miValid = false;
        V3 pMi;
        if (ray.medium >= 0) beta *= medium_sample(sc.media[ray.medium], ray, smp, &miValid, &pMi);
        if (is_black(beta)) break;
        // The medium-vertex and surface-vertex cases share ONE call of the direct-lighting code: it is most of the
        // work of an iteration (shadow rays with their transmittance walks, the MIS ray) and lanes of both kinds are
        // present in a warp; called from two places, the warp ran it twice with the lanes split between the calls.
        VPoint it;
        Surface s;
        Bsdf<8> bsdf;
        bsdf.n = 0;
        const V3 wo = -ray.d;
        float g = 0;
        if (miValid) {
            if (bounces >= rc.max_depth) break;
            g = sc.media[ray.medium].g;
            it.p = pMi; it.pError = V3(0.f); it.n = V3(0.f);
            it.mIn = it.mOut = ray.medium;
        } else {
            if (found) s = make_surface<TEX>(sc, hit.prim, hit.h.b0, hit.h.b1, hit.h.b2, ray.d);
            if (bounces == 0 || specularBounce) {
                if (found) { if (s.light >= 0) L += beta * area_light_L(sc.lights[s.light], s.n, -ray.d); }
                else if (sc.env.present) L += beta * env_Le(sc.env, ray.d);
            }
            if (!found || bounces >= rc.max_depth) break;
            it = surface_point(sc, s, ray.medium);
            if (s.material < 0) {
                ray = spawn_ray(it, ray.d);
                camDiff.has = false;
                bounces--;
                continue;
            }
            if (TEX) compute_differentials(s, camDiff);
            build_bsdf<8, TEX>(sc, sc.materials[s.material], s, bsdf);
        }
        camDiff.has = false;  // every ray spawned from here on is a plain Ray
        L += beta * vol_sample_one_light<8>(sc, it, miValid ? nullptr : &bsdf, miValid ? V3(0.f) : s.wo, miValid ? wo : V3(0.f), g, smp,
                                            stack, stride, cnt, vc.shadow, vc.mis);
        if (miValid) {
            float s0, s1;
            smp.get2d(&s0, &s1);
            V3 wi;
            hg_sample_p(wo, &wi, s0, s1, g);
            ray = spawn_ray(it, wi);
            specularBounce = false;
        } else {
            V3 wi;
            float pdf, b0, b1;
            int flags;
            smp.get2d(&b0, &b1);
            V3 f = bsdf_sample(bsdf, wo, &wi, b0, b1, &pdf, BSDF_ALL, &flags);
            if (is_black(f) || pdf == 0.f) break;
            beta *= div_each(f * absdot(wi, bsdf.ns), pdf);
            specularBounce = (flags & BSDF_SPECULAR) != 0;
            if ((flags & BSDF_SPECULAR) && (flags & BSDF_TRANSMISSION)) {
                float eta = bsdf.eta;
                etaScale *= (dot(wo, s.n) > 0) ? (eta * eta) : 1 / (eta * eta);
            }
            ray = spawn_ray(it, wi);
        }
        V3 rrBeta = beta * etaScale;
        float mx = max_component(rrBeta);
        if (mx < rc.rr_threshold && bounces > 3) {
            float q = fmaxf(.05f, 1 - mx);
            if (smp.get1d() < q) break;
            beta = div_each(beta, 1 - q);
        }
    }
    return L;
}

}  // namespace gnx
