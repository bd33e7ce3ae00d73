// gnx_volwave.cuh — VolPathIntegrator::Li (integrators/VolPathIntegrator.cpp:24-159) as a staged wavefront.
//
// The per-lane kernel (k_volpath, gnx_volpath.cuh) carried one path from camera to termination in one lane; ncu showed
// 4.6 active lanes per instruction, most of the time in the tracking loops of GridDensityMedium (delta tracking for the
// medium sample, ratio tracking for every shadow / MIS segment, media/GridDensityMedium.cpp:31-87) and in the direct
// lighting of medium vertices, with the lanes of a warp in different loops.  Here a path is a state machine that YIELDS
// whenever it needs a tracking walk:
//
//   k_vp_logic   advances paths from phase to phase (extend, vertex + light sample, shadow-walk segment, MIS-walk
//                segment, continuation) until the path ends or needs a walk through a grid medium; launched once per
//                resume phase, so the lanes of a warp start in the same phase
//   k_vp_track   persistent kernel over the queued walks: ONE tracking step per lane and iteration, lanes whose walk has
//                ended are refilled from the queue (lane refill per tracking step, not per path); on completion the path
//                goes to the queue of its resume phase
//
// The sampler travels with the path (PCG32 state, or the Halton dimension counter), so every draw is the draw the
// reference makes: results equal the per-lane kernel's and the reference's sample by sample.
#pragma once
#include "gnx_volpath.cuh"

namespace gnx {

// The eleven fields of a path's walk state share ONE 160-byte record (five full 32-byte sectors; SlotField, gnx_scene.cuh):
// with one array per field the vertex kernel moved 2.2 x its algorithmic bytes through DRAM (ncu), and the shadow walk
// took 19.3 ms against 13.1 ms now (C4).
template <class T>
using VolField = SlotField<T>;
constexpr int kVolRecordBytes = 160;
struct VolWave {        // per path slot, next to PathState
    VolField<uint2> rng;       // PCG32 state (lo, hi); Halton: x = dimension counter (index in PathState::hidx)
    VolField<float> tmi;       // medium sample on the current path segment: t of the interaction (ray parametrisation), -1 none
    VolField<float4> sub_o;    // walk (shadow / MIS) ray: origin, tMax
    VolField<float4> sub_d;    //   direction, bits(medium)
    VolField<float4> sub_hit;  //   hit of its current segment: b0 b1 b2 bits(prim), prim -1 = none
    VolField<float4> sub_tr;   //   transmittance so far xyz, bits(light number)
    VolField<float4> w0;       // shadow walk: f xyz, MIS weight      MIS walk: f xyz, MIS weight
    VolField<float4> w1;       // shadow walk: Li xyz, light pdf      MIS walk: Ld of the light-sampling half xyz, scattering pdf
    VolField<float4> w2;       // shadow walk: target point xyz       both: w = pdf of the light choice
    VolField<float4> w3;       // shadow walk: target normal xyz, uScattering.x
    VolField<float4> w4;       // shadow walk: target error xyz, uScattering.y
    // device layout: sub_o 0, sub_d 16, sub_hit 32, sub_tr 48, w0 64, w1 80, w2 96, w3 112, w4 128, rng 144, tmi 152
    void bind(void *records) {
        VolField<float4> *f4[] = {&sub_o, &sub_d, &sub_hit, &sub_tr, &w0, &w1, &w2, &w3, &w4};
        for (int k = 0; k < 9; ++k) *f4[k] = VolField<float4>(records, 16 * k, kVolRecordBytes);
        rng = VolField<uint2>(records, 144, kVolRecordBytes);
        tmi = VolField<float>(records, 152, kVolRecordBytes);
    }
};

// Phases of a path, and which logic kernel owns them.  A path that reaches a phase of ANOTHER kernel is stored and
// queued for it, so that every launch runs one kind of work with all its lanes:
//   KE  k_vp_logic<VK_EXTEND>   VP_EXTEND         closest hit of the path segment, HomogeneousMedium::Sample, boundary crossings
//   KV  k_vp_logic<VK_VERTEX>   VP_VERTEX         emission, termination, light choice + light sample, f and pdf towards it
//   KS  k_vp_logic<VK_SHADOW>   VP_SHADOW_*       VisibilityTester::Tr: the shadow walk through media and boundaries
//   KM  k_vp_logic<VK_MIS>      VP_MIS_* VP_CONT  BSDF / phase sample, Scene::IntersectTr walk, then the next direction + RR
// (VP_CONT also runs in KV for vertices that need no direct lighting).  A walk segment inside a GridDensityMedium yields
// to k_vp_track and comes back through the RESUME queue of its kernel.
enum VolPhase { VP_START = 0, VP_EXTEND, VP_VERTEX, VP_SHADOW_SEG, VP_SHADOW_AFTER_TR, VP_MIS_SETUP, VP_MIS_SEG, VP_MIS_AFTER_TR, VP_CONT };
enum VolKernel { VK_EXTEND = 0, VK_VERTEX, VK_SHADOW, VK_MIS, VK_ANY };
// return codes of vol_advance: done, a walk for k_vp_track (path segment / walk segment), or "queue for logic queue q"
enum VolYield { VY_DONE = 0, VY_TRACK_MAIN = 1, VY_TRACK_SUB = 2, VY_QUEUE0 = 8 };
// logic queues (slices of Queues::shade_q): entry phase of each
enum VolQueue { VQ_EXTEND = 0, VQ_VERTEX, VQ_SHADOW, VQ_SHADOW_RESUME, VQ_MIS, VQ_MIS_RESUME, VQ_COUNT };
constexpr uint32_t kVolPhaseMask = 0xffu;
GNX_HD int vol_queue_phase(int q) {
    return q == VQ_EXTEND ? VP_EXTEND : q == VQ_VERTEX ? VP_VERTEX : q == VQ_SHADOW ? VP_SHADOW_SEG
           : q == VQ_SHADOW_RESUME ? VP_SHADOW_AFTER_TR : q == VQ_MIS ? VP_MIS_SETUP : VP_MIS_AFTER_TR;
}
GNX_HD int vol_queue_kernel(int q) { return q == VQ_EXTEND ? VK_EXTEND : q == VQ_VERTEX ? VK_VERTEX : (q <= VQ_SHADOW_RESUME ? VK_SHADOW : VK_MIS); }
// resume queue of a path whose tracking walk has ended (its resume phase is in the low byte of PathState::meta)
GNX_HD int vol_resume_queue(int mode, int resumePhase) {
    return mode == 0 ? VQ_VERTEX : (resumePhase == VP_SHADOW_AFTER_TR ? VQ_SHADOW_RESUME : VQ_MIS_RESUME);
}

GNX_D PathSampler vol_load_sampler(const DeviceScene &sc, const RenderConsts &rc, const PathState &ps, const VolWave &vw, int slot, int px,
                                   int py, int sample) {
    const uint2 r = vw.rng[slot];
    if (sc.smp.type == GNX_SAMPLER_PCG32) {
        PathSampler s(sc.smp, 0, 0);
        s.pcg = true;
        s.rng.state = ((uint64_t)r.y << 32) | r.x;
        s.rng.inc = (((((uint64_t)(rc.width * py + px)) << 20) | (uint64_t)sample) << 1u) | 1u;
        return s;
    }
    return PathSampler(sc.smp, (uint64_t)ps.hidx[slot], (int)r.x);
}
GNX_D void vol_store_sampler(const VolWave &vw, int slot, const PathSampler &s) {
    vw.rng[slot] = s.pcg ? make_uint2((uint32_t)s.rng.state, (uint32_t)(s.rng.state >> 32)) : make_uint2((uint32_t)s.dim, 0u);
}

// The vertex of the current path segment, rebuilt from the path state whenever a phase needs it (between kernels
// nothing of it is kept but the hit record and the medium sample).
template <int MAXL>
struct VolVertex {
    bool built = false, medium = false;
    VPoint it;
    Surface s;
    Bsdf<MAXL> bsdf;
    V3 wo;       // -ray.d
    float g = 0;
};

// Advances path `slot` from phase `phase` inside logic kernel `kernel` (VK_ANY: the sequential driver, every phase in
// place) until the path ends (VY_DONE), needs a tracking walk through a grid medium (VY_TRACK_MAIN: delta tracking
// along the path segment, VY_TRACK_SUB: ratio tracking along the walk segment; the resume phase is left in the low
// byte of ps.meta), or reaches a phase of another kernel (VY_QUEUE0 + queue).
// TEX: the scene has image textures — the camera segment keeps its ray differentials (VolPathIntegrator::Li takes a
// RayDifferential and hands it to ComputeScatteringFunctions at the first vertex; every spawned ray is a plain Ray,
// integrators/VolPathIntegrator.cpp:30,97-105,130), so a textured surface the camera sees directly is filtered with
// MIPMap::Lookup's trilinear / EWA path.
template <int MAXL, bool TEX = false>
GNX_D int vol_advance(const DeviceScene &sc, const RenderConsts &rc, const PathState &ps, const VolWave &vw, int slot, int phase,
                      int kernel, int2 *stack, int stride, TraversalCounters &cnt, VolCounters &vc) {
    const int kNonSpec = BSDF_ALL & ~BSDF_SPECULAR;
    int pixel, sample, px, py;
    slot_to_sample(rc, slot, &pixel, &sample);
    if (!pixel_xy(rc, pixel, &px, &py)) { ps.L[slot] = make_float4(0.f, 0.f, 0.f, 0.f); return VY_DONE; }
    // ---- path state
    VRay ray;
    V3 L(0.f), beta(1.f);
    float etaScale = 1;
    int bounces = 0;
    bool specularBounce = false, cameraDiff = false;
    VHit hit;
    hit.prim = -1;
    bool found = false;
    PathSampler smp(sc.smp, 0, 0);
    if (phase == VP_START) {
        cameraDiff = TEX;
        if (sc.smp.type == GNX_SAMPLER_PCG32) smp.take(PathSampler::stream(sc.smp, ((uint64_t)(rc.width * py + px) << 20) | (uint64_t)sample));
        else smp.take(PathSampler(sc.smp, sampler_index(sc.smp, px, py, (uint64_t)sample), 0));
        float u0, u1, l0, l1;
        smp.get_film(px, py, &u0, &u1);
        smp.get1d();  // time
        smp.get2d(&l0, &l1);
        camera_ray_uv(sc, px, py, u0, u1, l0, l1, &ray.o, &ray.d, &ray.tMax);
        ray.medium = sc.cam.medium;
        if (!smp.pcg) ps.hidx[slot] = (uint32_t)smp.index;
        phase = VP_EXTEND;
    } else {
        const float4 ro = ps.ray_o[slot], rd = ps.ray_d[slot], b4 = ps.beta[slot], L4 = ps.L[slot], h4 = ps.hit[slot];
        ray.o = V3(ro.x, ro.y, ro.z); ray.tMax = ro.w;
        ray.d = V3(rd.x, rd.y, rd.z); etaScale = rd.w;
        ray.medium = ps.medium[slot];
        beta = V3(b4.x, b4.y, b4.z);
        L = V3(L4.x, L4.y, L4.z);
        const uint32_t meta = ps.meta[slot];
        bounces = (int)((meta >> 16) & 0xff);
        specularBounce = ((meta >> 24) & kFlagSpecular) != 0;
        cameraDiff = TEX && ((meta >> 24) & kFlagCameraDiff) != 0;
        hit.prim = f2i(h4.w); hit.h.b0 = h4.x; hit.h.b1 = h4.y; hit.h.b2 = h4.z; hit.h.t = ray.tMax;
        found = hit.prim >= 0;
        smp.take(vol_load_sampler(sc, rc, ps, vw, slot, px, py, sample));
    }
    // ---- walk state (kept in the VolWave arrays between kernels)
    VRay sub;
    sub.o = V3(0.f); sub.d = V3(0.f); sub.tMax = 0; sub.medium = -1;
    VHit subHit;
    subHit.prim = -1;
    // wf / wWeight / wPdf: f, MIS weight and pdf of the walk's estimator; w1v: Li of the light sample during the shadow
    // walk, Ld of the light-sampling half afterwards
    V3 Tr(1.f), wf(0.f), w1v(0.f), p1p(0.f), p1n(0.f), p1e(0.f);
    float wWeight = 0, wPdf = 0, selPdf = 0, us0 = 0, us1 = 0;
    int lightNum = 0;
    if (phase == VP_SHADOW_SEG || phase == VP_SHADOW_AFTER_TR || phase == VP_MIS_SETUP || phase == VP_MIS_AFTER_TR) {
        const float4 so = vw.sub_o[slot], sd = vw.sub_d[slot], sh = vw.sub_hit[slot], st = vw.sub_tr[slot];
        const float4 a0 = vw.w0[slot], a1 = vw.w1[slot], a2 = vw.w2[slot], a3 = vw.w3[slot], a4 = vw.w4[slot];
        sub.o = V3(so.x, so.y, so.z); sub.tMax = so.w;
        sub.d = V3(sd.x, sd.y, sd.z); sub.medium = f2i(sd.w);
        subHit.prim = f2i(sh.w); subHit.h.b0 = sh.x; subHit.h.b1 = sh.y; subHit.h.b2 = sh.z; subHit.h.t = sub.tMax;
        Tr = V3(st.x, st.y, st.z); lightNum = f2i(st.w);
        wf = V3(a0.x, a0.y, a0.z); wWeight = a0.w;
        w1v = V3(a1.x, a1.y, a1.z); wPdf = a1.w;
        p1p = V3(a2.x, a2.y, a2.z); selPdf = a2.w;
        p1n = V3(a3.x, a3.y, a3.z); us0 = a3.w;
        p1e = V3(a4.x, a4.y, a4.z); us1 = a4.w;
    }
    VolVertex<MAXL> vx;
    auto ensureVertex = [&]() {
        if (vx.built) return;
        vx.built = true;
        vx.wo = -ray.d;
        const float tmi = vw.tmi[slot];
        vx.medium = tmi >= 0.f;
        vx.bsdf.n = 0;
        if (vx.medium) {
            vx.g = sc.media[ray.medium].g;
            vx.it.p = ray.o + ray.d * tmi; vx.it.pError = V3(0.f); vx.it.n = V3(0.f);
            vx.it.mIn = vx.it.mOut = ray.medium;
        } else {
            vx.s = make_surface<TEX>(sc, hit.prim, hit.h.b0, hit.h.b1, hit.h.b2, ray.d);
            vx.it = surface_point(sc, vx.s, ray.medium);
            if (TEX && cameraDiff && vx.s.material >= 0) {
                // the camera sample again (its five draws open the stream), for the offset rays
                float u0, u1, l0, l1;
                if (sc.smp.type == GNX_SAMPLER_PCG32) {
                    PathSampler cs = PathSampler::stream(sc.smp, ((uint64_t)(rc.width * py + px) << 20) | (uint64_t)sample);
                    cs.get2d(&u0, &u1); cs.get1d(); cs.get2d(&l0, &l1);
                } else {
                    const uint64_t hi = (uint64_t)ps.hidx[slot];
                    sampler_film_dimensions(sc.smp, hi, px, py, &u0, &u1);
                    l0 = halton_sample_dimension(sc.smp, hi, 3); l1 = halton_sample_dimension(sc.smp, hi, 4);
                }
                compute_differentials(vx.s, camera_ray_differentials(sc, px, py, u0, u1, l0, l1, ray.o, ray.d));
            }
            if (vx.s.material >= 0) build_bsdf<MAXL, TEX>(sc, sc.materials[vx.s.material], vx.s, vx.bsdf);
        }
    };
    auto storeMain = [&](int resume) {
        ps.ray_o[slot] = make_float4(ray.o.x, ray.o.y, ray.o.z, ray.tMax);
        ps.ray_d[slot] = make_float4(ray.d.x, ray.d.y, ray.d.z, etaScale);
        ps.medium[slot] = ray.medium;
        ps.beta[slot] = make_float4(beta.x, beta.y, beta.z, 0.f);
        ps.L[slot] = make_float4(L.x, L.y, L.z, 0.f);
        ps.hit[slot] = make_float4(hit.h.b0, hit.h.b1, hit.h.b2, i2f(found ? hit.prim : -1));
        ps.meta[slot] = (uint32_t)resume | ((uint32_t)bounces << 16) | (((specularBounce ? kFlagSpecular : 0u) | (cameraDiff ? kFlagCameraDiff : 0u)) << 24);
        vol_store_sampler(vw, slot, smp);
    };
    auto storeSub = [&]() {
        vw.sub_o[slot] = make_float4(sub.o.x, sub.o.y, sub.o.z, sub.tMax);
        vw.sub_d[slot] = make_float4(sub.d.x, sub.d.y, sub.d.z, i2f(sub.medium));
        vw.sub_hit[slot] = make_float4(subHit.h.b0, subHit.h.b1, subHit.h.b2, i2f(subHit.prim));
        vw.sub_tr[slot] = make_float4(Tr.x, Tr.y, Tr.z, i2f(lightNum));
        vw.w0[slot] = make_float4(wf.x, wf.y, wf.z, wWeight);
        vw.w1[slot] = make_float4(w1v.x, w1v.y, w1v.z, wPdf);
        vw.w2[slot] = make_float4(p1p.x, p1p.y, p1p.z, selPdf);
        vw.w3[slot] = make_float4(p1n.x, p1n.y, p1n.z, us0);
        vw.w4[slot] = make_float4(p1e.x, p1e.y, p1e.z, us1);
    };
    // a phase owned by another kernel: park the path in that kernel's queue
    auto foreign = [&](int q) { return kernel != VK_ANY && vol_queue_kernel(q) != kernel; };
    bool misFound = false;
    Surface misSurface;
    misSurface.n = V3(0.f); misSurface.prim = -1;

    while (true) {
        switch (phase) {
        case VP_EXTEND: {
            if (foreign(VQ_EXTEND)) { storeMain(VP_EXTEND); return VY_QUEUE0 + VQ_EXTEND; }
            ++vc.extend;
            found = vol_intersect(sc, ray, &hit, stack, stride, cnt);
            vx.built = false;
            float tmi = -1.f;
            if (ray.medium >= 0) {
                const DevMedium &m = sc.media[ray.medium];
                if (m.type == GNX_MEDIUM_GRID) {
                    vw.tmi[slot] = -1.f;
                    storeMain(VP_VERTEX);
                    return VY_TRACK_MAIN;  // the tracker applies sigma_s / sigma_t to beta when it samples an interaction
                }
                beta *= homogeneous_sample(m, ray.d, ray.tMax, smp, &tmi);
            }
            vw.tmi[slot] = tmi;
            // A medium boundary (surface without a material) reached without a medium interaction is crossed right here —
            // VolPathIntegrator.cpp:101-106: same direction, the bounce does not count — when the vertex logic has nothing
            // else to do for it: no emission to add and no termination to decide.
            if (tmi < 0.f && found && !is_black(beta) && bounces < rc.max_depth) {
                float4 c;
                load_tri(sc.tris, hit.prim, &c);
                const bool noMaterial = (f2u(c.y) & 0xfffffu) == 0u, emits = f2i(c.z) >= 0;
                if (noMaterial && !emits) {
                    const Surface s = make_surface(sc, hit.prim, hit.h.b0, hit.h.b1, hit.h.b2, ray.d);
                    ray = spawn_ray(surface_point(sc, s, ray.medium), ray.d);
                    cameraDiff = false;
                    break;  // phase stays VP_EXTEND
                }
            }
            phase = VP_VERTEX;
            break;
        }
        case VP_VERTEX: {
            if (foreign(VQ_VERTEX)) { storeMain(VP_VERTEX); return VY_QUEUE0 + VQ_VERTEX; }
            if (is_black(beta)) { storeMain(VP_CONT); return VY_DONE; }
            const bool miValid = vw.tmi[slot] >= 0.f;
            if (miValid) {
                if (bounces >= rc.max_depth) { storeMain(VP_CONT); return VY_DONE; }
                ensureVertex();
            } else {
                if (found) ensureVertex();
                if (bounces == 0 || specularBounce) {
                    if (found) { if (vx.s.light >= 0) L += beta * area_light_L(sc.lights[vx.s.light], vx.s.n, -ray.d); }
                    else if (sc.env.present) L += beta * env_Le(sc.env, ray.d);
                }
                if (!found || bounces >= rc.max_depth) { storeMain(VP_CONT); return VY_DONE; }
                if (vx.s.material < 0) {  // medium boundary: same direction, the bounce does not count
                    ray = spawn_ray(vx.it, ray.d);
                    cameraDiff = false;
                    phase = VP_EXTEND;
                    break;
                }
            }
            // ---- UniformSampleOneLight / EstimateDirect (handleMedia = true), light-sampling half
            w1v = V3(0.f);
            selPdf = 0;
            phase = VP_CONT;
            if (sc.n_lights == 0) break;
            lightNum = choose_light(sc, vx.it.p, smp.get1d(), &selPdf);
            if (selPdf == 0) break;
            float ul0, ul1;
            smp.get2d(&ul0, &ul1);
            smp.get2d(&us0, &us1);
            const gnx_light &light = sc.lights[lightNum];
            const bool isEnv = light.type == GNX_LIGHT_INFINITE;
            LightSample ls;
            const bool ok = isEnv ? env_sample_li(sc.env, ul0, ul1, &ls) : area_sample_li(sc, light, vx.it.p, ul0, ul1, &ls);
            phase = VP_MIS_SETUP;
            if (ok && ls.pdf > 0 && !is_black(ls.Li)) {
                V3 f;
                float scatteringPdf;
                if (!vx.medium) {
                    bsdf_f_pdf(vx.bsdf, vx.s.wo, ls.wi, kNonSpec, &f, &scatteringPdf);
                    f = f * absdot(ls.wi, vx.bsdf.ns);
                } else {
                    float p = phase_hg(dot(vx.wo, ls.wi), vx.g);
                    f = V3(p);
                    scatteringPdf = p;
                }
                if (!is_black(f)) {
                    VPoint p1;
                    if (isEnv) { p1.p = vx.it.p + ls.wi * (2 * sc.env.world_radius); p1.pError = V3(0.f); p1.n = V3(0.f); }
                    else { p1.p = ls.pl; p1.pError = ls.plError; p1.n = ls.nl; }
                    p1.mIn = p1.mOut = light.medium;
                    p1p = p1.p; p1n = p1.n; p1e = p1.pError;
                    wf = f; w1v = ls.Li; wPdf = ls.pdf;
                    wWeight = (ls.pdf * ls.pdf) / (ls.pdf * ls.pdf + scatteringPdf * scatteringPdf);
                    sub = spawn_ray_to(vx.it, p1);
                    Tr = V3(1.f);
                    phase = VP_SHADOW_SEG;
                }
            }
            break;
        }
        case VP_SHADOW_SEG: {  // VisibilityTester::Tr, one segment (core/Light.cpp:33-53)
            if (foreign(VQ_SHADOW)) { storeMain(VP_SHADOW_SEG); storeSub(); return VY_QUEUE0 + VQ_SHADOW; }
            ++vc.shadow;
            const bool hitSurface = vol_intersect(sc, sub, &subHit, stack, stride, cnt);
            if (!hitSurface) subHit.prim = -1;
            bool blocked = false;
            if (hitSurface) {
                float4 c;
                load_tri(sc.tris, subHit.prim, &c);
                blocked = (f2u(c.y) & 0xfffffu) != 0u;  // a surface with a material
            }
            if (blocked) { Tr = V3(0.f); goto shadow_done; }
            if (sub.medium >= 0) {
                const DevMedium &m = sc.media[sub.medium];
                if (m.type == GNX_MEDIUM_GRID) {
                    storeMain(VP_SHADOW_AFTER_TR);
                    storeSub();
                    return VY_TRACK_SUB;  // the tracker multiplies sub_tr by the walk's transmittance
                }
                Tr *= homogeneous_tr(m, sub.d, sub.tMax);
            }
            phase = VP_SHADOW_AFTER_TR;
            break;
        }
        case VP_SHADOW_AFTER_TR: {
            if (kernel != VK_ANY && kernel != VK_SHADOW) return VY_DONE;  // (not reached: lets the other kernels drop the code)
            if (subHit.prim < 0) goto shadow_done;
            const Surface s = make_surface(sc, subHit.prim, subHit.h.b0, subHit.h.b1, subHit.h.b2, sub.d);
            VPoint p1;
            p1.p = p1p; p1.n = p1n; p1.pError = p1e;
            p1.mIn = p1.mOut = sc.lights[lightNum].medium;
            sub = spawn_ray_to(surface_point(sc, s, sub.medium), p1);
            phase = VP_SHADOW_SEG;
            break;
        }
        case VP_MIS_SETUP: {
            // ---- BSDF / phase-function sampling half of EstimateDirect; w1v holds the light-sampling half's Ld
            if (foreign(VQ_MIS)) { storeMain(VP_MIS_SETUP); storeSub(); return VY_QUEUE0 + VQ_MIS; }
            ensureVertex();
            const gnx_light &light = sc.lights[lightNum];
            const bool isEnv = light.type == GNX_LIGHT_INFINITE;
            V3 wi, f;
            float scatteringPdf;
            if (!vx.medium) {
                int sampledType;
                f = bsdf_sample(vx.bsdf, vx.s.wo, &wi, us0, us1, &scatteringPdf, kNonSpec, &sampledType);
                f = f * absdot(wi, vx.bsdf.ns);
            } else {
                float p = hg_sample_p(vx.wo, &wi, us0, us1, vx.g);
                f = V3(p);
                scatteringPdf = p;
            }
            bool walk = false;
            if (!is_black(f) && scatteringPdf > 0) {
                sub = spawn_ray(vx.it, wi);
                float lightPdf = isEnv ? env_pdf_li(sc.env, wi) : area_pdf_li(sc, light, vx.it.p, sub.o, wi);
                if (lightPdf != 0) {
                    wWeight = (scatteringPdf * scatteringPdf) / (scatteringPdf * scatteringPdf + lightPdf * lightPdf);
                    wf = f; wPdf = scatteringPdf;
                    Tr = V3(1.f);
                    walk = true;
                }
            }
            if (walk) { phase = VP_MIS_SEG; break; }
            L += beta * div_each(w1v, selPdf);
            phase = VP_CONT;
            break;
        }
        case VP_MIS_SEG: {  // Scene::IntersectTr, one segment (core/Scene.cpp:26-40)
            if (kernel != VK_ANY && kernel != VK_MIS) return VY_DONE;  // (not reached)
            ++vc.mis;
            const bool hitSurface = vol_intersect(sc, sub, &subHit, stack, stride, cnt);
            if (!hitSurface) subHit.prim = -1;
            if (sub.medium >= 0) {
                const DevMedium &m = sc.media[sub.medium];
                if (m.type == GNX_MEDIUM_GRID) {
                    storeMain(VP_MIS_AFTER_TR);
                    storeSub();
                    return VY_TRACK_SUB;
                }
                Tr *= homogeneous_tr(m, sub.d, sub.tMax);
            }
            phase = VP_MIS_AFTER_TR;
            break;
        }
        case VP_MIS_AFTER_TR: {
            if (kernel != VK_ANY && kernel != VK_MIS) return VY_DONE;  // (not reached)
            if (subHit.prim < 0) { misFound = false; goto mis_done; }
            misSurface = make_surface(sc, subHit.prim, subHit.h.b0, subHit.h.b1, subHit.h.b2, sub.d);
            if (misSurface.material >= 0) { misFound = true; goto mis_done; }
            sub = spawn_ray(surface_point(sc, misSurface, sub.medium), sub.d);
            phase = VP_MIS_SEG;
            break;
        }
        case VP_CONT: {
            if (kernel == VK_EXTEND || kernel == VK_SHADOW) return VY_DONE;  // (not reached)
            ensureVertex();
            if (vx.medium) {
                float s0, s1;
                smp.get2d(&s0, &s1);
                V3 wi;
                hg_sample_p(vx.wo, &wi, s0, s1, vx.g);
                ray = spawn_ray(vx.it, wi);
                specularBounce = false;
            } else {
                V3 wi;
                float pdf, b0, b1;
                int flags;
                smp.get2d(&b0, &b1);
                V3 f = bsdf_sample(vx.bsdf, vx.wo, &wi, b0, b1, &pdf, BSDF_ALL, &flags);
                if (is_black(f) || pdf == 0.f) { storeMain(VP_CONT); return VY_DONE; }
                beta *= div_each(f * absdot(wi, vx.bsdf.ns), pdf);
                specularBounce = (flags & BSDF_SPECULAR) != 0;
                if ((flags & BSDF_SPECULAR) && (flags & BSDF_TRANSMISSION)) {
                    float eta = vx.bsdf.eta;
                    etaScale *= (dot(vx.wo, vx.s.n) > 0) ? (eta * eta) : 1 / (eta * eta);
                }
                ray = spawn_ray(vx.it, wi);
            }
            V3 rrBeta = beta * etaScale;
            float mx = max_component(rrBeta);
            if (mx < rc.rr_threshold && bounces > 3) {
                float q = fmaxf(.05f, 1 - mx);
                if (smp.get1d() < q) { storeMain(VP_CONT); return VY_DONE; }
                beta = div_each(beta, 1 - q);
            }
            ++bounces;
            cameraDiff = false;
            phase = VP_EXTEND;
            break;
        }
        default: storeMain(VP_CONT); return VY_DONE;
        }
        continue;

    shadow_done : {
        // Li *= visibility.Tr; Ld += f * Li * weight / lightPdf
        const V3 Li = w1v * Tr;
        w1v = V3(0.f);
        if (!is_black(Li)) w1v += div_each(wf * Li * wWeight, wPdf);
        phase = VP_MIS_SETUP;
        continue;
    }
    mis_done : {
        const gnx_light &light = sc.lights[lightNum];
        const bool isEnv = light.type == GNX_LIGHT_INFINITE;
        V3 Li(0.f);
        if (misFound) { if (!isEnv && misSurface.prim == light.prim) Li = area_light_L(light, misSurface.n, -sub.d); }
        else if (isEnv) Li = env_Le(sc.env, sub.d);
        V3 Ld = w1v;
        if (!is_black(Li)) Ld += div_each(wf * Li * Tr * wWeight, wPdf);
        L += beta * div_each(Ld, selPdf);
        phase = VP_CONT;
        continue;
    }
    }
}

// One queued tracking walk: loads the ray (path segment or walk segment) and the sampler, and on completion applies the
// result — delta tracking: the interaction's t and sigma_s / sigma_t on beta; ratio tracking: the walk's transmittance.
struct TrackLane {
    TrackState ts;
    int slot, mode, medium;
};
GNX_D bool vol_track_begin(const DeviceScene &sc, const PathState &ps, const VolWave &vw, int slot, int mode, TrackLane &tl) {
    tl.slot = slot; tl.mode = mode;
    V3 o, d;
    float tMax;
    if (mode == 0) {
        const float4 ro = ps.ray_o[slot], rd = ps.ray_d[slot];
        o = V3(ro.x, ro.y, ro.z); d = V3(rd.x, rd.y, rd.z); tMax = ro.w;
        tl.medium = ps.medium[slot];
    } else {
        const float4 so = vw.sub_o[slot], sd = vw.sub_d[slot];
        o = V3(so.x, so.y, so.z); d = V3(sd.x, sd.y, sd.z); tMax = so.w;
        tl.medium = f2i(sd.w);
    }
    return track_begin(sc.media[tl.medium], o, d, tMax, mode, tl.ts);
}
GNX_D void vol_track_finish(const DeviceScene &sc, const PathState &ps, const VolWave &vw, const TrackLane &tl, const PathSampler &smp) {
    const int slot = tl.slot;
    if (tl.mode == 0) {
        if (tl.ts.sampled) {
            const DevMedium &m = sc.media[tl.medium];
            vw.tmi[slot] = tl.ts.t;  // MediumInteraction(rWorld(t)): t of the normalised walk on the UN-normalised world ray, as the reference writes it
            float4 b = ps.beta[slot];
            const V3 w = div_each(V3(m.sigma_s[0], m.sigma_s[1], m.sigma_s[2]), m.sigma_t_scalar);
            b.x *= w.x; b.y *= w.y; b.z *= w.z;
            ps.beta[slot] = b;
        }
    } else {
        float4 t4 = vw.sub_tr[slot];
        t4.x *= tl.ts.Tr; t4.y *= tl.ts.Tr; t4.z *= tl.ts.Tr;
        vw.sub_tr[slot] = t4;
    }
    vol_store_sampler(vw, slot, smp);
}

// Sequential driver (CPU emulation, tests): the same state machine for one path.  staged = true follows the path
// through the kernels exactly as the device does — every hand-over between two logic kernels and every tracking walk goes
// through the stored state — staged = false runs all phases in place.
template <int MAXL, bool TEX = false>
GNX_D V3 volwave_li(const DeviceScene &sc, const RenderConsts &rc, const PathState &ps, const VolWave &vw, int slot, int2 *stack, int stride,
                    TraversalCounters &cnt, VolCounters &vc, bool staged = true) {
    int phase = VP_START, kernel = staged ? VK_EXTEND : VK_ANY;
    while (true) {
        const int y = vol_advance<MAXL, TEX>(sc, rc, ps, vw, slot, phase, kernel, stack, stride, cnt, vc);
        if (y == VY_DONE) break;
        if (y >= VY_QUEUE0) {
            phase = vol_queue_phase(y - VY_QUEUE0);
            kernel = vol_queue_kernel(y - VY_QUEUE0);
            continue;
        }
        int pixel, sample, px, py;
        slot_to_sample(rc, slot, &pixel, &sample);
        pixel_xy(rc, pixel, &px, &py);
        PathSampler smp = vol_load_sampler(sc, rc, ps, vw, slot, px, py, sample);
        TrackLane tl;
        const int mode = y == VY_TRACK_MAIN ? 0 : 1;
        if (vol_track_begin(sc, ps, vw, slot, mode, tl))
            while (!track_step(sc.media[tl.medium], tl.ts, smp)) {}
        vol_track_finish(sc, ps, vw, tl, smp);
        const int q = vol_resume_queue(mode, (int)(ps.meta[slot] & kVolPhaseMask));
        phase = vol_queue_phase(q);
        if (staged) kernel = vol_queue_kernel(q);
    }
    const float4 L = ps.L[slot];
    return V3(L.x, L.y, L.z);
}

}  // namespace gnx
