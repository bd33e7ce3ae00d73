// gnx_whitted.cuh — WhittedIntegrator and DirectLightingIntegrator, with the lights the
// reference's UI uses with them: PointLight, SpotLight, DistantLight, SkyBoxLight (next to DiffuseAreaLight and
// InfiniteAreaLight).
//
// Both integrators are recursive (SpecularReflect / SpecularTransmit, core/Integrator.cpp:321-420) and draw their
// Halton dimensions in depth-first order: the reflection subtree of a vertex consumes its dimensions before the
// transmission sample of that vertex is drawn.  A breadth-first wavefront would have to know the size of every
// subtree in advance, so — like VolPath — the recursion runs per lane, as an explicit depth-first stack of frames;
// a pending transmission frame keeps only the hit record and is re-shaded when it is popped.
//
// Ray differentials (RayDifferential, SurfaceInteraction::ComputeDifferentials, the reflected / refracted offset rays of
// SpecularReflect / SpecularTransmit) are carried only when the scene has image textures: they feed MIPMap::Lookup's
// trilinear / EWA filtering and nothing else (TEX template parameter).
#pragma once
#include "gnx_volpath.cuh"

namespace gnx {

struct RecCounters { unsigned extend, shadow, mis; };

// VisibilityTester::Unoccluded (core/Light.cpp:22-25): !scene.IntersectP(p0.SpawnRayTo(p1))
GNX_D bool w_unoccluded(const DeviceScene &sc, const Surface &s, const WLightSample &ls, int2 *stack, int stride,
                        TraversalCounters &cnt, RecCounters &rcnt) {
    V3 origin = offset_ray_origin(s.p, s.pError, s.n, ls.target - s.p);
    V3 target = ls.target;
    if (ls.targetN.x != 0 || ls.targetN.y != 0 || ls.targetN.z != 0 || ls.targetErr.x != 0 || ls.targetErr.y != 0 || ls.targetErr.z != 0)
        target = offset_ray_origin(ls.target, ls.targetErr, ls.targetN, origin - ls.target);
    V3 d = target - origin;
    int prim;
    TriHit h;
    ++rcnt.shadow;
    // (the two-child tree: inside this per-lane kernel the wide tree measured SLOWER — U1w 68 -> 89 ms, W1 23.9 -> 27.1 ms;
    // the staged first vertex sends its shadow rays through k_anyhit8 instead)
    return !traverse<true>(sc, origin, d, 1 - kShadowEpsilon, stack, stride, &prim, &h, cnt);
}

// EstimateDirect (core/Integrator.cpp:99-215) for a surface, handleMedia = false, specular = false
template <int MAXL>
GNX_D V3 w_estimate_direct(const DeviceScene &sc, const Surface &s, const Bsdf<MAXL> &bsdf, const gnx_light &light, float ul0, float ul1,
                           float us0, float us1, int2 *stack, int stride, TraversalCounters &cnt, RecCounters &rcnt) {
    const int kNonSpec = BSDF_ALL & ~BSDF_SPECULAR;
    V3 Ld(0.f);
    WLightSample ls;
    w_sample_li(sc, light, s.p, ul0, ul1, &ls);
    if (ls.pdf > 0 && !is_black(ls.Li)) {
        V3 f;
        float scatteringPdf;
        bsdf_f_pdf(bsdf, s.wo, ls.wi, kNonSpec, &f, &scatteringPdf);
        f = f * absdot(ls.wi, bsdf.ns);
        if (!is_black(f)) {
            V3 Li = ls.Li;
            if (!w_unoccluded(sc, s, ls, stack, stride, cnt, rcnt)) Li = V3(0.f);
            if (!is_black(Li)) {
                if (ls.delta) Ld += div_each(f * Li, ls.pdf);
                else {
                    float weight = (ls.pdf * ls.pdf) / (ls.pdf * ls.pdf + scatteringPdf * scatteringPdf);
                    Ld += div_each(f * Li * weight, ls.pdf);
                }
            }
        }
    }
    if (!ls.delta) {
        V3 wi;
        float scatteringPdf;
        int sampledType;
        V3 f = bsdf_sample(bsdf, s.wo, &wi, us0, us1, &scatteringPdf, kNonSpec, &sampledType);
        f = f * absdot(wi, bsdf.ns);
        const bool sampledSpecular = (sampledType & BSDF_SPECULAR) != 0;
        if (!is_black(f) && scatteringPdf > 0) {
            float weight = 1;
            V3 o = offset_ray_origin(s.p, s.pError, s.n, wi);
            if (!sampledSpecular) {
                float lightPdf = w_pdf_li(sc, light, s.p, o, wi);
                if (lightPdf == 0) return Ld;
                weight = (scatteringPdf * scatteringPdf) / (scatteringPdf * scatteringPdf + lightPdf * lightPdf);
            }
            int prim;
            TriHit h;
            ++rcnt.mis;
            bool found = closest_hit(sc, o, wi, GNX_INF, stack, stride, &prim, &h, cnt);
            V3 Li(0.f);
            if (found) {
                if (light.type == GNX_LIGHT_AREA_TRI && prim == light.prim) {
                    Surface ls2 = make_surface(sc, prim, h.b0, h.b1, h.b2, wi);
                    Li = area_light_L(light, ls2.n, -wi);
                }
            } else if (light.type == GNX_LIGHT_INFINITE) Li = env_Le(sc.env, wi);
            else if (light.type == GNX_LIGHT_SKYBOX) Li = skybox_le(sc.skybox, o, wi);
            if (!is_black(Li)) Ld += div_each(f * Li * weight, scatteringPdf);
        }
    }
    return Ld;
}

// One pending piece of the recursion.
struct RecFrame {
    RayDiff rd;       // the ray's differentials (kind 1: of the ray that produced the hit)
    V3 o, d;          // kind 0: the ray to trace.  kind 1: d = direction of the ray that produced the hit
    V3 weight;        // product of f * |cos| / pdf down to here
    float hb0, hb1, hb2;
    int prim;         // kind 1: the hit to re-shade
    int depth;
    int kind;         // 0 = Li(ray, depth), 1 = SpecularTransmit of an already lit vertex
};
constexpr int kMaxRecDepth = 16;

// WhittedIntegrator::Li / DirectLightingIntegrator::Li for camera sample `sample` of pixel (px, py).
//   direct = 0: integrators/WhittedIntegrator.cpp:14-67      direct = 1 / 2: integrators/DirectLightingIntegrator.cpp:28-63
//   with LightStrategy::UniformSampleOne / UniformSampleAll.
//
// UniformSampleAll draws from the sampler's 2-D sample ARRAYS (Preprocess requests, per depth and light, one array for the
// light samples and one for the BSDF samples, DirectLightingIntegrator.cpp:13-27).  GlobalSampler::StartPixel fills element
// k of pixel sample s of array a with dimensions (5 + 2a, 5 + 2a + 1) of Halton index GetIndexForSample(s n + k)
// (core/Sampler.cpp:134-144), so the arrays need no storage here: an element is two radical inverses.  The arrays are
// handed out in the order the vertices ask for them (depth first); a sample whose recursion visits more vertices than
// maxDepth finds none left and falls back to one Get2D pair per light (core/Integrator.cpp:38-43).  Ordinary dimensions
// start behind the arrays' (GlobalSampler::Get1D / Get2D skip [arrayStartDim, arrayEndDim), core/Sampler.cpp).
// (direct is a template parameter: with the three integrators in one body the Whitted render measured 12 % slower —
// the UniformSampleAll code costs registers in the hot loop even when it never runs.)
template <int MAXL, int direct, bool TEX = false>
GNX_D V3 recursive_li(const DeviceScene &sc, const RenderConsts &rc, int px, int py, int sample, int2 *stack, int stride,
                      TraversalCounters &cnt, RecCounters &rcnt) {
    const uint64_t hidx = sampler_index(sc.smp, px, py, (uint64_t)sample);
    const int nArrays = direct == 2 ? 2 * rc.max_depth * sc.n_lights : 0;
    int arrayCursor = 0;               // Sampler::array2DOffset
    PathSampler smp(sc.smp, hidx, 5 + 2 * nArrays);  // dimensions 0-4 belong to the camera sample, then the arrays'
    RecFrame frames[kMaxRecDepth + 2];
    int nf = 0;
    {
        V3 o, d;
        float tMax;
        camera_ray(sc, px, py, hidx, &o, &d, &tMax);
        RecFrame &f = frames[nf++];
        f.o = o; f.d = d; f.weight = V3(1.f); f.depth = 0; f.kind = 0; f.prim = -1;
        f.hb0 = f.hb1 = f.hb2 = 0;
        f.rd.has = false;
        if (TEX) {
            float u0, u1;
            sampler_film_dimensions(sc.smp, hidx, px, py, &u0, &u1);
            float l0 = 0, l1 = 0;
            if (sc.cam.lens_radius > 0) { l0 = halton_sample_dimension(sc.smp, hidx, 3); l1 = halton_sample_dimension(sc.smp, hidx, 4); }
            f.rd = camera_ray_differentials(sc, px, py, u0, u1, l0, l1, o, d);
        }
    }
    V3 L(0.f);
    const int maxDepth = rc.max_depth < kMaxRecDepth ? rc.max_depth : kMaxRecDepth;
    while (nf > 0) {
        RecFrame fr = frames[--nf];
        Surface s;
        Bsdf<MAXL> bsdf;
        bool lit = false;  // true: a freshly hit vertex (lights + reflection), false: the postponed transmission
        if (fr.kind == 0) {
            // ---- Li(ray, depth): closest hit, skipping surfaces without a material at the same depth
            bool escaped = false;
            for (int guard = 0; guard < 4096; ++guard) {
                int prim;
                TriHit h;
                ++rcnt.extend;
                if (!closest_hit(sc, fr.o, fr.d, GNX_INF, stack, stride, &prim, &h, cnt)) { escaped = true; break; }
                s = make_surface<TEX>(sc, prim, h.b0, h.b1, h.b2, fr.d);
                fr.prim = prim; fr.hb0 = h.b0; fr.hb1 = h.b1; fr.hb2 = h.b2;
                if (s.material >= 0) break;
                fr.o = offset_ray_origin(s.p, s.pError, s.n, fr.d);  // isect.SpawnRay(ray.d): a plain Ray, the differentials end here
                fr.rd.has = false;
            }
            if (escaped) { L += fr.weight * scene_le(sc, fr.o, fr.d); continue; }
            if (s.material < 0) continue;
            lit = true;
        } else {
            s = make_surface<TEX>(sc, fr.prim, fr.hb0, fr.hb1, fr.hb2, fr.d);
        }
        if (TEX) compute_differentials(s, fr.rd);  // isect.ComputeScatteringFunctions(ray, ...) -> ComputeDifferentials(ray)
        build_bsdf<MAXL, TEX>(sc, sc.materials[s.material], s, bsdf, false);
        const V3 ns = bsdf.ns;
        if (lit) {
            // ---- emitted light, then the direct illumination of the integrator
            if (s.light >= 0) L += fr.weight * area_light_L(sc.lights[s.light], s.n, s.wo);
            V3 lightL(0.f);
            if constexpr (direct == 0) {
                for (int j = 0; j < sc.n_lights; ++j) {
                    float u0, u1;
                    smp.get2d(&u0, &u1);
                    WLightSample ls;
                    w_sample_li(sc, sc.lights[j], s.p, u0, u1, &ls);
                    if (is_black(ls.Li) || ls.pdf == 0) continue;
                    V3 f;
                    float pdfUnused;
                    bsdf_f_pdf(bsdf, s.wo, ls.wi, BSDF_ALL, &f, &pdfUnused);
                    if (!is_black(f) && w_unoccluded(sc, s, ls, stack, stride, cnt, rcnt))
                        lightL += div_each(f * ls.Li * absdot(ls.wi, ns), ls.pdf);
                }
            } else if constexpr (direct == 2) {
                // UniformSampleAllLights (core/Integrator.cpp:25-55)
                for (int j = 0; j < sc.n_lights; ++j) {
                    const int n = sc.light_nsamples ? sc.light_nsamples[j] : 1;
                    if (arrayCursor + 2 > nArrays) {
                        float ul0, ul1, us0, us1;
                        smp.get2d(&ul0, &ul1);
                        smp.get2d(&us0, &us1);
                        lightL += w_estimate_direct<MAXL>(sc, s, bsdf, sc.lights[j], ul0, ul1, us0, us1, stack, stride, cnt, rcnt);
                    } else {
                        const int dimL = 5 + 2 * arrayCursor, dimS = dimL + 2;
                        arrayCursor += 2;
                        V3 Ld(0.f);
                        for (int k = 0; k < n; ++k) {
                            const uint64_t idx = sampler_index(sc.smp, px, py, (uint64_t)sample * (uint64_t)n + (uint64_t)k);
                            const float ul0 = halton_sample_dimension(sc.smp, idx, dimL), ul1 = halton_sample_dimension(sc.smp, idx, dimL + 1);
                            const float us0 = halton_sample_dimension(sc.smp, idx, dimS), us1 = halton_sample_dimension(sc.smp, idx, dimS + 1);
                            Ld += w_estimate_direct<MAXL>(sc, s, bsdf, sc.lights[j], ul0, ul1, us0, us1, stack, stride, cnt, rcnt);
                        }
                        lightL += div_each(Ld, (float)n);
                    }
                }
            } else if (sc.n_lights > 0) {
                // UniformSampleOneLight without a light distribution (core/Integrator.cpp:70-79)
                int lightNum = (int)(smp.get1d() * sc.n_lights);
                if (lightNum > sc.n_lights - 1) lightNum = sc.n_lights - 1;
                const float lightPdf = 1.f / sc.n_lights;
                float ul0, ul1, us0, us1;
                smp.get2d(&ul0, &ul1);
                smp.get2d(&us0, &us1);
                lightL = div_each(w_estimate_direct<MAXL>(sc, s, bsdf, sc.lights[lightNum], ul0, ul1, us0, us1, stack, stride, cnt, rcnt), lightPdf);
            }
            L += fr.weight * lightL;
            if (!(fr.depth + 1 < maxDepth)) continue;
        }
        // ---- SpecularReflect (lit vertex) or SpecularTransmit (postponed frame): BSDF::Sample_f restricted to the
        // specular lobes of one hemisphere
        float u0, u1, pdf;
        smp.get2d(&u0, &u1);
        V3 wi;
        int sampledType;
        const int lobeFlags = (lit ? BSDF_REFLECTION : BSDF_TRANSMISSION) | BSDF_SPECULAR;
        V3 f = bsdf_sample(bsdf, s.wo, &wi, u0, u1, &pdf, lobeFlags, &sampledType);
        if (lit && nf < kMaxRecDepth + 1) {
            // the transmission sample of this vertex is drawn after the whole reflection subtree
            RecFrame &t = frames[nf++];
            t = fr;
            t.kind = 1;
        }
        if (pdf > 0.f && !is_black(f) && absdot(wi, ns) != 0.f && nf < kMaxRecDepth + 2) {
            RecFrame &c = frames[nf++];
            c.o = offset_ray_origin(s.p, s.pError, s.n, wi);
            c.d = wi;
            c.weight = fr.weight * div_each(f * absdot(wi, ns), pdf);
            c.depth = fr.depth + 1;
            c.kind = 0;
            c.prim = -1;
            c.hb0 = c.hb1 = c.hb2 = 0;
            c.rd.has = false;
            if (TEX && fr.rd.has) {
                // differentials of the reflected / refracted ray (core/Integrator.cpp:336-356, 381-437)
                c.rd.has = true;
                c.rd.rxo = s.p + s.dpdx;
                c.rd.ryo = s.p + s.dpdy;
                V3 nsd = ns;
                V3 dndx = s.dndu * s.dudx + s.dndv * s.dvdx, dndy = s.dndu * s.dudy + s.dndv * s.dvdy;
                const V3 wo = s.wo;
                const V3 dwodx = -fr.rd.rxd - wo, dwody = -fr.rd.ryd - wo;
                if (lit) {
                    const float dDNdx = dot(dwodx, nsd) + dot(wo, dndx), dDNdy = dot(dwody, nsd) + dot(wo, dndy);
                    c.rd.rxd = wi - dwodx + 2.f * (dot(wo, nsd) * dndx + dDNdx * nsd);
                    c.rd.ryd = wi - dwody + 2.f * (dot(wo, nsd) * dndy + dDNdy * nsd);
                } else {
                    float eta = 1 / bsdf.eta;
                    if (dot(wo, nsd) < 0) { eta = 1 / eta; nsd = -nsd; dndx = -dndx; dndy = -dndy; }
                    const float dDNdx = dot(dwodx, nsd) + dot(wo, dndx), dDNdy = dot(dwody, nsd) + dot(wo, dndy);
                    const float mu = eta * dot(wo, nsd) - absdot(wi, nsd);
                    const float dmudx = (eta - (eta * eta * dot(wo, nsd)) / absdot(wi, nsd)) * dDNdx;
                    const float dmudy = (eta - (eta * eta * dot(wo, nsd)) / absdot(wi, nsd)) * dDNdy;
                    c.rd.rxd = wi - eta * dwodx + (mu * dndx + dmudx * nsd);
                    c.rd.ryd = wi - eta * dwody + (mu * dndy + dmudy * nsd);
                }
            }
        }
    }
    return L;
}

// ---- WhittedIntegrator, first vertex as a wavefront stage ---------------------------------------------------
// Most camera samples of a Whitted render end at their first vertex: emitted light plus one shadow ray per light
// (WhittedIntegrator.cpp:36-58); only surfaces with specular lobes recurse (:59-64).  The staged form (gnx_render.cu):
// k_trace<3> finds the camera rays' hits, k_whitted_vertex runs this function over them, the any-hit kernel answers the
// shadow rays of ALL lights at once, k_whitted_sum adds the surviving contributions in light order, and the samples this
// function turns down (specular lobes, surfaces without a material) go through recursive_li as before.
//
// Per light j a vertex yields either nothing (contribution zero) or a shadow item whose contribution f * Li * |cos| / pdf
// the any-hit kernel stores in plane j of the light planes when the light is visible (d_path.w = j * capacity + slot).
template <int MAXL>
struct WhittedVertex {
    Surface s;
    Bsdf<MAXL> bsdf;
    uint64_t hidx;
};
// The vertex of camera sample (px, py, sample) at hit (prim, b0 b1 b2).  False: the sample needs the recursion.
template <int MAXL, bool TEX>
GNX_D bool whitted_vertex_begin(const DeviceScene &sc, int px, int py, int sample, int prim, float b0, float b1, float b2,
                                WhittedVertex<MAXL> &v, V3 *LeOut) {
    v.hidx = sampler_index(sc.smp, px, py, (uint64_t)sample);
    V3 o, d;
    float tMax;
    camera_ray(sc, px, py, v.hidx, &o, &d, &tMax);
    v.s = make_surface<TEX>(sc, prim, b0, b1, b2, d);
    if (v.s.material < 0) return false;
    if (TEX) {
        float u0, u1;
        sampler_film_dimensions(sc.smp, v.hidx, px, py, &u0, &u1);
        float l0 = 0, l1 = 0;
        if (sc.cam.lens_radius > 0) { l0 = halton_sample_dimension(sc.smp, v.hidx, 3); l1 = halton_sample_dimension(sc.smp, v.hidx, 4); }
        const RayDiff rd = camera_ray_differentials(sc, px, py, u0, u1, l0, l1, o, d);
        compute_differentials(v.s, rd);
    }
    build_bsdf<MAXL, TEX>(sc, sc.materials[v.s.material], v.s, v.bsdf, false);
    for (int i = 0; i < v.bsdf.n; ++i) if (v.bsdf.lobes[i].type & BSDF_SPECULAR) return false;  // SpecularReflect / Transmit would recurse
    *LeOut = v.s.light >= 0 ? area_light_L(sc.lights[v.s.light], v.s.n, v.s.wo) : V3(0.f);
    return true;
}
// Light j of that vertex (WhittedIntegrator.cpp:44-57; the j-th Get2D of the sample = dimensions 5 + 2j, 5 + 2j + 1).
// True: *it is the shadow ray to trace (d_path.w left for the caller's plane index).
template <int MAXL>
GNX_D bool whitted_vertex_light(const DeviceScene &sc, const WhittedVertex<MAXL> &v, int j, ShadowItem *it) {
    const Surface &s = v.s;
    const float u0 = halton_sample_dimension(sc.smp, v.hidx, 5 + 2 * j), u1 = halton_sample_dimension(sc.smp, v.hidx, 5 + 2 * j + 1);
    WLightSample ls;
    w_sample_li(sc, sc.lights[j], s.p, u0, u1, &ls);
    if (is_black(ls.Li) || ls.pdf == 0) return false;
    V3 f;
    float pdfUnused;
    bsdf_f_pdf(v.bsdf, s.wo, ls.wi, BSDF_ALL, &f, &pdfUnused);
    if (is_black(f)) return false;
    // VisibilityTester::Unoccluded (w_unoccluded above): the segment from the offset origin to the offset target
    V3 origin = offset_ray_origin(s.p, s.pError, s.n, ls.target - s.p);
    V3 target = ls.target;
    if (ls.targetN.x != 0 || ls.targetN.y != 0 || ls.targetN.z != 0 || ls.targetErr.x != 0 || ls.targetErr.y != 0 || ls.targetErr.z != 0)
        target = offset_ray_origin(ls.target, ls.targetErr, ls.targetN, origin - ls.target);
    const V3 sd = target - origin;
    const V3 c = div_each(f * ls.Li * absdot(ls.wi, v.bsdf.ns), ls.pdf);
    it->o_tmax = make_float4(origin.x, origin.y, origin.z, 1 - kShadowEpsilon);
    it->d_path = make_float4(sd.x, sd.y, sd.z, 0.f);
    it->contrib = make_float4(c.x, c.y, c.z, 0.f);
    return true;
}

}  // namespace gnx

namespace gnx {
// run-time selection of the integrator (host emulation, tests)
template <int MAXL>
GNX_D V3 recursive_li(const DeviceScene &sc, const RenderConsts &rc, int direct, int px, int py, int sample, int2 *stack, int stride,
                      TraversalCounters &cnt, RecCounters &rcnt, bool tex = false) {
    if (tex) {
        if (direct == 0) return recursive_li<MAXL, 0, true>(sc, rc, px, py, sample, stack, stride, cnt, rcnt);
        if (direct == 1) return recursive_li<MAXL, 1, true>(sc, rc, px, py, sample, stack, stride, cnt, rcnt);
        return recursive_li<MAXL, 2, true>(sc, rc, px, py, sample, stack, stride, cnt, rcnt);
    }
    if (direct == 0) return recursive_li<MAXL, 0>(sc, rc, px, py, sample, stack, stride, cnt, rcnt);
    if (direct == 1) return recursive_li<MAXL, 1>(sc, rc, px, py, sample, stack, stride, cnt, rcnt);
    return recursive_li<MAXL, 2>(sc, rc, px, py, sample, stack, stride, cnt, rcnt);
}
}  // namespace gnx
