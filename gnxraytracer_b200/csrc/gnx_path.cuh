// gnx_path.cuh — what ONE path does in each wavefront stage, as __host__ __device__ functions.
// The kernels in gnx_kernels.cuh are thin wrappers (pick the slot, call these, push to the queues
// with warp-aggregated atomics); tests/emul runs the very same functions sequentially on the CPU
// to check the logic against the reference without a GPU.
//
//   raygen_slot    GetCameraSample + GenerateRayDifferential          (Sampler.cpp:14-20, Perspective.cpp:62-112)
//   extend_slot    scene.Intersect; escaped rays pick up InfiniteAreaLight::Le (PathIntegrator.cpp:98-117)
//   shade_slot     Le at the vertex, material -> lobes, UniformSampleOneLight / EstimateDirect sample
//                  generation, BSDF sampling, Russian roulette              (PathIntegrator.cpp:101-204)
//   shadow_item    VisibilityTester::Unoccluded, and the "ray escapes" half of MIS for the environment
//   probe_item     closest-hit MIS probe for area lights                    (core/Integrator.cpp:191-207)
#pragma once
#include "gnx_shade.cuh"

namespace gnx {

struct RenderConsts {
    int width, height, npix;
    int max_depth;
    float rr_threshold;
    int batch_spp;       // samples per pixel in this batch
    int first_sample;    // sample number of the batch's first sample
    int capacity;        // path slots
    // Interleaved-tile partition of an N-device job (SURVEY 8e mode 2): this device renders the kTile x kTile tiles
    // number tile_dev, tile_dev + tile_n, ... of the skewed row-major tile order; npix is then its LOCAL pixel count
    // (tiles x kTile^2, pixels of edge tiles outside the image included and skipped).  tile_n == 0: the whole image.
    int tile_n, tile_dev, tiles_x, tiles_y;
    int rec_list;        // k_recursive: 0 = every camera sample of the batch, 1 + q = the samples listed in extend queue q
};
constexpr int kTile = 32;

// Local pixel index -> image coordinates; false for the pixels of an edge tile that lie outside the image.
// Tile number t of the order sits in tile row t / tiles_x, shifted by its row number (a skew, so that one device's
// tiles form diagonals instead of columns).
GNX_D bool pixel_xy(const RenderConsts &rc, int pixel, int *px, int *py) {
    if (rc.tile_n == 0) { *px = pixel % rc.width; *py = pixel / rc.width; return true; }
    const int lt = pixel / (kTile * kTile), li = pixel % (kTile * kTile);
    const int gt = lt * rc.tile_n + rc.tile_dev;
    const int ty = gt / rc.tiles_x, tx = (gt % rc.tiles_x + ty) % rc.tiles_x;
    *px = tx * kTile + li % kTile;
    *py = ty * kTile + li / kTile;
    return gt < rc.tiles_x * rc.tiles_y && *px < rc.width && *py < rc.height;
}
// Tiles of the order that device `dev` of `n` owns.
GNX_HD int local_tile_count(int tiles_x, int tiles_y, int n, int dev) {
    const int total = tiles_x * tiles_y;
    return total > dev ? (total - dev + n - 1) / n : 0;
}

// Camera ray for (pixel, halton index): Sampler::GetCameraSample + PerspectiveCamera::GenerateRay +
// Transform::operator()(Ray) (core/Transform.h:230-244).
GNX_D void camera_ray_uv(const DeviceScene &sc, int px, int py, float u0, float u1, float l0, float l1, V3 *o, V3 *d,
                         float *tMax) {
    V3 pFilm((float)px + u0, (float)py + u1, 0.f);
    V3 pCamera = xform_point(sc.cam.r2c, pFilm);
    V3 ro(0.f, 0.f, 0.f);
    V3 rd = normalize(pCamera);
    if (sc.cam.lens_radius > 0) {
        float lx, ly;
        concentric_sample_disk(l0, l1, &lx, &ly);
        lx *= sc.cam.lens_radius; ly *= sc.cam.lens_radius;
        float ft = sc.cam.focal_distance / rd.z;
        V3 pFocus = ro + rd * ft;
        ro = V3(lx, ly, 0);
        rd = normalize(pFocus - ro);
    }
    V3 oErr;
    V3 wo = xform_point_err(sc.cam.c2w, ro, &oErr);
    V3 wd = xform_vector(sc.cam.c2w, rd);
    float lsq = length_sq(wd);
    float tm = GNX_INF;
    if (lsq > 0) {
        float dt = dot(vabs(wd), oErr) / lsq;
        wo = wo + wd * dt;
        tm -= dt;
    }
    *o = wo; *d = wd; *tMax = tm;
}
// The offset rays of PerspectiveCamera::GenerateRayDifferential (camera/Perspective.cpp:62-112), carried to world space
// by Transform::operator()(RayDifferential) (core/Transform.h:246-256: the main ray with its error-bound shift, the
// offset rays as plain points / vectors) and scaled by 1 / sqrt(samplesPerPixel) as Render does (core/Integrator.cpp:277).
// (o, d) is the world ray camera_ray_uv returned for the same sample.
GNX_D RayDiff camera_ray_differentials(const DeviceScene &sc, int px, int py, float u0, float u1, float l0, float l1, V3 o, V3 d) {
    RayDiff rd;
    rd.has = true;
    const V3 pFilm((float)px + u0, (float)py + u1, 0.f);
    const V3 pCamera = xform_point(sc.cam.r2c, pFilm);
    V3 rxo(0.f), ryo(0.f), rxd, ryd;
    if (sc.cam.lens_radius > 0) {
        float lx, ly;
        concentric_sample_disk(l0, l1, &lx, &ly);
        lx *= sc.cam.lens_radius; ly *= sc.cam.lens_radius;
        const V3 dx = normalize(pCamera + sc.cam.dx_camera);
        float ft = sc.cam.focal_distance / dx.z;
        V3 pFocus = V3(0.f) + (ft * dx);
        rxo = V3(lx, ly, 0);
        rxd = normalize(pFocus - rxo);
        const V3 dy = normalize(pCamera + sc.cam.dy_camera);
        ft = sc.cam.focal_distance / dy.z;
        pFocus = V3(0.f) + (ft * dy);
        ryo = V3(lx, ly, 0);
        ryd = normalize(pFocus - ryo);
    } else {
        rxd = normalize(pCamera + sc.cam.dx_camera);
        ryd = normalize(pCamera + sc.cam.dy_camera);
    }
    rxo = xform_point(sc.cam.c2w, rxo); ryo = xform_point(sc.cam.c2w, ryo);
    rxd = xform_vector(sc.cam.c2w, rxd); ryd = xform_vector(sc.cam.c2w, ryd);
    const float sscale = 1 / sqrtf((float)sc.smp.spp);
    rd.rxo = o + (rxo - o) * sscale; rd.ryo = o + (ryo - o) * sscale;
    rd.rxd = d + (rxd - d) * sscale; rd.ryd = d + (ryd - d) * sscale;
    return rd;
}
GNX_D void camera_ray(const DeviceScene &sc, int px, int py, uint64_t hidx, V3 *o, V3 *d, float *tMax) {
    float u0, u1;
    sampler_film_dimensions(sc.smp, hidx, px, py, &u0, &u1);
    float l0 = 0, l1 = 0;
    if (sc.cam.lens_radius > 0) { l0 = halton_sample_dimension(sc.smp, hidx, 3); l1 = halton_sample_dimension(sc.smp, hidx, 4); }
    camera_ray_uv(sc, px, py, u0, u1, l0, l1, o, d, tMax);
}

// Slot s of a batch holds sample (s % batch_spp) of pixel (s / batch_spp): the 32 lanes of a warp trace samples
// of the same pixel (or of neighbouring pixels when the batch holds fewer than 32 samples per pixel).  Their rays
// visit the same nodes and triangles, so the per-lane node gathers collapse into a few L1 wavefronts and the
// lanes stay in step — the traversal kernels are bound by exactly those two things.
// (Walking the pixels in 8x8 tiles instead of rows changed nothing measurable.)
GNX_D void slot_to_sample(const RenderConsts &rc, int slot, int *pixel, int *sample) {
    *pixel = slot / rc.batch_spp;
    *sample = rc.first_sample + slot % rc.batch_spp;
}

GNX_D int shade_type_of(unsigned matWord) { return (int)((matWord >> 20) & 0xfu); }
constexpr int kPendEscape = -2;  // primary_finish / extend_finish: "queue the slot for k_escape" (scenes with a SkyBoxLight)

// Camera sample `sample` of pixel (px, py): the ray is generated in registers and traversed at once
// (no ray round trip through HBM for the ~90 % of C2's camera rays that never touch the mesh).
GNX_D void primary_begin(const DeviceScene &sc, int px, int py, int sample, uint32_t *hidxOut, V3 *dOut, Trav &t) {
    uint64_t hidx = sampler_index(sc.smp, px, py, (uint64_t)sample);
    V3 o, d;
    float tMax;
    camera_ray(sc, px, py, hidx, &o, &d, &tMax);
    trav_init(sc, t, o, d, tMax);
    *hidxOut = (uint32_t)hidx;
    *dOut = d;
}
// Path state is written only for camera rays that hit something; an escaped ray's radiance is
// beta (= 1) * Le, the first bounce of PathIntegrator::Li (PathIntegrator.cpp:101-117).
// Returns the shade-queue type or -1.
GNX_D int primary_finish(const DeviceScene &sc, const PathState &ps, const RenderConsts &rc, int slot, uint32_t hidx, V3 d,
                         const Trav &t) {
    if (!t.hit) {
        if (sc.skybox.present) {
            // SkyBoxLight::Le needs the ray and double-precision atan2 / asin: escaped rays of such scenes are queued
            // for k_escape instead of paying for that code (and its registers) inside the traversal kernel
            ps.L[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
            ps.ray_o[slot] = make_float4(t.o.x, t.o.y, t.o.z, GNX_INF);
            ps.ray_d[slot] = make_float4(d.x, d.y, d.z, 1.f);
            ps.beta[slot] = make_float4(1.f, 1.f, 1.f, 0.f);
            return kPendEscape;
        }
        V3 Le = sc.env.present ? env_Le(sc.env, d) : V3(0.f);
        ps.L[slot] = make_float4(Le.x, Le.y, Le.z, 0.f);
        return -1;
    }
    ps.L[slot] = make_float4(0.f, 0.f, 0.f, (ps.Lb || ps.La) ? 1.f : 0.f);
    if (ps.Lb) ps.Lb[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ps.La) ps.La[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
    ps.ray_d[slot] = make_float4(d.x, d.y, d.z, 1.f);  // w: etaScale = 1
    ps.beta[slot] = make_float4(1.f, 1.f, 1.f, 0.f);
    ps.hidx[slot] = hidx;
    ps.meta[slot] = 5u;  // dimension 5 (film 0-1, time 2, lens 3-4), bounces 0, flags 0
    ps.hit[slot] = make_float4(t.h.b0, t.h.b1, t.h.b2, i2f(t.prim));
    return shade_type_of(f2u(ldg(&sc.tris[3 * t.prim + 2].y)));
}


// Extension ray of a path slot: begin loads the ray and arms the traversal, finish consumes the result.
GNX_D void extend_begin(const DeviceScene &sc, const PathState &ps, int slot, Trav &t) {
    const float4 ro = ps.ray_o[slot], rd = ps.ray_d[slot];
    trav_init(sc, t, V3(ro.x, ro.y, ro.z), V3(rd.x, rd.y, rd.z), ro.w);
}
// Returns the shade-queue type of the hit, or -1 when the path ends here.
GNX_D int extend_finish(const DeviceScene &sc, const PathState &ps, const RenderConsts &rc, int slot, const Trav &t) {
    const uint32_t meta = ps.meta[slot];
    const int bounces = (meta >> 16) & 0xff;
    const bool emitOk = bounces == 0 || ((meta >> 24) & kFlagSpecular);
    if (t.hit) {
        // beyond maxDepth only the emission term is left (PathIntegrator.cpp:101-117)
        if (bounces >= rc.max_depth && !emitOk) return -1;
        const unsigned matWord = f2u(ldg(&sc.tris[3 * t.prim + 2].y));
        ps.hit[slot] = make_float4(t.h.b0, t.h.b1, t.h.b2, i2f(t.prim));
        return shade_type_of(matWord);
    }
    if (emitOk && sc.skybox.present) return kPendEscape;  // -> k_escape (ray and throughput are in the path state)
    if (emitOk && sc.env.present) {
        // for (light : scene.infiniteLights) L += beta * light->Le(ray)
        const float4 b = ps.beta[slot], rd = ps.ray_d[slot];
        float4 L = ps.L[slot];
        V3 add = V3(b.x, b.y, b.z) * env_Le(sc.env, V3(rd.x, rd.y, rd.z));
        L.x += add.x; L.y += add.y; L.z += add.z;
        ps.L[slot] = L;
    }
    return -1;
}
// An escaped ray of a scene with a SkyBoxLight: for (light : scene.infiniteLights) L += beta * light->Le(ray)
// (PathIntegrator.cpp:112-116), both infinite kinds in the order of scene.lights.
GNX_D void escape_slot(const DeviceScene &sc, const PathState &ps, int slot) {
    const float4 b = ps.beta[slot], ro = ps.ray_o[slot], rd = ps.ray_d[slot];
    float4 L = ps.L[slot];
    V3 add = V3(b.x, b.y, b.z) * scene_le(sc, V3(ro.x, ro.y, ro.z), V3(rd.x, rd.y, rd.z));
    L.x += add.x; L.y += add.y; L.z += add.z;
    ps.L[slot] = L;
}
GNX_D int extend_slot(const DeviceScene &sc, const PathState &ps, const RenderConsts &rc, int slot, int2 *stack, int stride,
                      TraversalCounters &cnt) {
    TravLocal t;
    extend_begin(sc, ps, slot, t);
    closest_hit_run(sc, t, stack, stride, cnt);  // the 8-wide tree when the scene routes closest hits there
    int type = extend_finish(sc, ps, rc, slot, t);
    if (type == kPendEscape) { escape_slot(sc, ps, slot); type = -1; }  // sequential callers: at once
    return type;
}

// Surfaces without a material are medium boundaries: PathIntegrator re-spawns the ray in the same
// direction and does not count the bounce (PathIntegrator.cpp:121-126).  Returns "still alive".
GNX_D bool shade_null_slot(const DeviceScene &sc, const PathState &ps, const RenderConsts &rc, int slot) {
    const float4 rd = ps.ray_d[slot], hit = ps.hit[slot];
    const V3 d(rd.x, rd.y, rd.z);
    Surface s = make_surface(sc, f2i(hit.w), hit.x, hit.y, hit.z, d);
    const uint32_t meta = ps.meta[slot];
    const int bounces = (meta >> 16) & 0xff;
    const bool emitOk = bounces == 0 || ((meta >> 24) & kFlagSpecular);
    if (emitOk && s.light >= 0) {
        const float4 b = ps.beta[slot];
        float4 L = ps.L[slot];
        V3 add = V3(b.x, b.y, b.z) * area_light_L(sc.lights[s.light], s.n, -d);
        L.x += add.x; L.y += add.y; L.z += add.z;
        ps.L[slot] = L;
    }
    if (bounces >= rc.max_depth) return false;
    V3 o = offset_ray_origin(s.p, s.pError, s.n, d);
    ps.ray_o[slot] = make_float4(o.x, o.y, o.z, GNX_INF);
    return true;
}

struct ShadeOut {
    bool alive, haveShadowA, haveShadowB, haveProbe;
    ShadowItem shA, shB;
    ProbeItem pr;
};

// The function has no early exit and a fixed sequence of stages: with SYNC the thread block meets at a barrier
// between stages (every thread of the block must call it, `valid` false for threads without an item).  The
// shading code is ~50 KB of straight-line instructions per item against a 32 KB instruction cache; warps drifting
// through it independently made the kernel instruction-fetch bound (ncu: stall_no_instruction first, 26-30 % issue
// utilisation).  Marching in step, the warps of a block share each fetched line.
#if defined(__CUDA_ARCH__)
#define GNX_STAGE_SYNC() do { if (SYNC) __syncthreads(); } while (0)
#else
#define GNX_STAGE_SYNC() do { } while (0)
#endif
// NEXT: the scene holds PointLight / SpotLight / DistantLight / SkyBoxLight records as well (EstimateDirect's delta
// branch, core/Integrator.cpp:148,159; SkyBoxLight::Pdf_Li is 0, so its BSDF-sampling half contributes nothing).  A
// template parameter so that the area + environment scenes of the BASELINE configs keep their code size.
template <int MAXL, bool SYNC = false, bool NEXT = false>
GNX_D void shade_slot(const DeviceScene &sc, const PathState &ps, const RenderConsts &rc, int slot, ShadeOut &out, bool valid = true) {
    const int kNonSpec = BSDF_ALL & ~BSDF_SPECULAR;
    out.alive = out.haveShadowA = out.haveShadowB = out.haveProbe = false;
    if (!valid) slot = 0;
    const float4 rd4 = ps.ray_d[slot], hit = ps.hit[slot], b4 = ps.beta[slot];
    const V3 rayD(rd4.x, rd4.y, rd4.z);
    V3 beta(b4.x, b4.y, b4.z);
    float etaScale = rd4.w;
    const uint32_t meta = ps.meta[slot];
    const int bounces = (meta >> 16) & 0xff;
    const bool specularBounce = ((meta >> 24) & kFlagSpecular) != 0;
    PathSampler smp(sc.smp, (uint64_t)ps.hidx[slot], (int)(meta & 0xffff));
    Surface s;
    Bsdf<MAXL> bsdf;
    bsdf.n = 0;
    bool live = valid;
    if (live) {
        s = make_surface(sc, f2i(hit.w), hit.x, hit.y, hit.z, rayD);
        // ---- emitted light at the vertex (PathIntegrator.cpp:101-111)
        if ((bounces == 0 || specularBounce) && s.light >= 0) {
            float4 L4 = ps.L[slot];
            V3 add = beta * area_light_L(sc.lights[s.light], s.n, -rayD);
            L4.x += add.x; L4.y += add.y; L4.z += add.z;
            ps.L[slot] = L4;
        }
        live = bounces < rc.max_depth;
    }
    GNX_STAGE_SYNC();
    if (live) build_bsdf<MAXL>(sc, sc.materials[s.material], s, bsdf);
    const V3 ns = bsdf.ns;
    GNX_STAGE_SYNC();
    // ---- UniformSampleOneLight (core/Integrator.cpp:57-79), EstimateDirect (core/Integrator.cpp:107-208) and the
    // continuation sample (PathIntegrator.cpp:143-163) as three passes over ONE copy of the BSDF code:
    //   P_LIGHT  direction from the light sample     -> f, pdf for it                -> shadow ray A
    //   P_MIS    direction from BSDF::Sample_f(kNonSpec) -> f, pdf                   -> shadow ray B / probe
    //   P_CONT   direction from BSDF::Sample_f(BSDF_ALL) -> f, pdf                   -> next path segment
    // Inlined three times the lobe code made the kernel 54 K instructions, with a hot footprint of ~107 KB against
    // a 32 KB instruction cache (46 % hit rate, stall_no_instruction on top).
    enum { P_LIGHT = 0, P_MIS = 1, P_CONT = 2 };
    const int nNonSpec = bsdf.num_components(kNonSpec);
    float selPdf = 0, ul0 = 0, ul1 = 0, us0 = 0, us1 = 0;
    int lightNum = 0;
    if (live && nNonSpec > 0 && sc.n_lights > 0) {
        lightNum = choose_light(sc, s.p, smp.get1d(), &selPdf);
        if (selPdf != 0) {
            smp.get2d(&ul0, &ul1);
            smp.get2d(&us0, &us1);
        }
    }
    const gnx_light &light = sc.lights[lightNum];
    const bool isEnv = sc.n_lights > 0 && light.type == GNX_LIGHT_INFINITE;
    const bool isArea = !NEXT ? !isEnv : (sc.n_lights > 0 && light.type == GNX_LIGHT_AREA_TRI);
    bool lsDelta = false;
    V3 wi(0.f), f(0.f);
    float pdf = 0;
    int flags = 0;
#pragma unroll 1
    for (int pass = P_LIGHT; pass <= P_CONT; ++pass) {
        GNX_STAGE_SYNC();
        const bool run = live && (pass == P_CONT || (selPdf != 0 && (pass == P_LIGHT || isEnv || isArea)));
        const int lobeFlags = pass == P_CONT ? BSDF_ALL : kNonSpec;
        const int matching = pass == P_CONT ? bsdf.num_components(BSDF_ALL) : nNonSpec;
        const V3 woW = pass == P_CONT ? -rayD : s.wo;  // isect.wo is normalized, PathIntegrator's wo = -ray.d is not
        const V3 wo = bsdf.to_local(woW);
        V3 wiL;
        int chosen = -1;
        bool have = false;
        LightSample ls;
        pdf = 0;
        f = V3(0.f);
        flags = 0;
        if (!run) {
        } else if (pass == P_LIGHT) {
            bool ok;
            if (NEXT && !isEnv && !isArea) {
                WLightSample wl;
                ok = w_sample_li(sc, light, s.p, ul0, ul1, &wl);
                ls.wi = wl.wi; ls.Li = wl.Li; ls.pdf = wl.pdf; ls.pl = wl.target; ls.nl = wl.targetN; ls.plError = wl.targetErr;
                lsDelta = wl.delta;
            } else ok = isEnv ? env_sample_li(sc.env, ul0, ul1, &ls) : area_sample_li(sc, light, s.p, ul0, ul1, &ls);
            have = ok && ls.pdf > 0 && !is_black(ls.Li) && wo.z != 0;
            wi = ls.wi;
            wiL = bsdf.to_local(wi);
        } else {
            float ua = us0, ub = us1;
            if (pass == P_CONT) smp.get2d(&ua, &ub);
            have = bsdf_sample_dir(bsdf, wo, ua, ub, lobeFlags, matching, &chosen, &wiL, &pdf, &flags, &f);
            if (have) wi = bsdf.to_world(wiL);
        }
        GNX_STAGE_SYNC();
        if (have) {
            if (chosen < 0 || !(bsdf.lobes[chosen].type & BSDF_SPECULAR)) {
                const bool refl = dot(wi, bsdf.ng) * dot(woW, bsdf.ng) > 0;
                bsdf_eval_sum(bsdf, wo, wiL, refl, lobeFlags, chosen, &f, &pdf);
            }
            pdf /= matching;
        }
        GNX_STAGE_SYNC();
        if (!run) {
        } else if (pass == P_LIGHT) {
            if (!have) continue;
            // ---- EstimateDirect, light-sampling half
            f = f * absdot(ls.wi, ns);
            const float scatteringPdf = pdf;
            if (!is_black(f)) {
                V3 origin, dirv;
                if (isEnv || (NEXT && !isArea)) {
                    // VisibilityTester(ref, Interaction(ref.p + wi * (2 * worldRadius), ...)): the far
                    // point has neither normal nor error bound, so SpawnRayTo's target is the point itself
                    // (likewise the position of a point / spot light and the far point of a distant / skybox light)
                    V3 p1 = isEnv ? s.p + ls.wi * (2 * sc.env.world_radius) : ls.pl;
                    origin = offset_ray_origin(s.p, s.pError, s.n, p1 - s.p);
                    dirv = p1 - origin;
                } else {
                    origin = offset_ray_origin(s.p, s.pError, s.n, ls.pl - s.p);
                    V3 target = offset_ray_origin(ls.pl, ls.plError, ls.nl, origin - ls.pl);
                    dirv = target - origin;
                }
                // IsDeltaLight: no MIS weight (core/Integrator.cpp:157-163)
                float weight = (ls.pdf * ls.pdf) / (ls.pdf * ls.pdf + scatteringPdf * scatteringPdf);
                V3 Ld = (NEXT && lsDelta) ? div_each(f * ls.Li, ls.pdf) : div_each(f * ls.Li * weight, ls.pdf);
                V3 c = beta * div_each(Ld, selPdf);
                out.shA.o_tmax = make_float4(origin.x, origin.y, origin.z, 1 - kShadowEpsilon);
                out.shA.d_path = make_float4(dirv.x, dirv.y, dirv.z, i2f(slot));
                out.shA.contrib = make_float4(c.x, c.y, c.z, 0.f);
                out.haveShadowA = true;
            }
        } else if (pass == P_MIS) {
            if (!have) continue;
            // ---- EstimateDirect, BSDF-sampling half
            const float scatteringPdf = pdf;
            f = f * absdot(wi, ns);
            if (!is_black(f) && scatteringPdf > 0) {
                V3 o = offset_ray_origin(s.p, s.pError, s.n, wi);
                float lightPdf = isEnv ? env_pdf_li(sc.env, wi) : area_pdf_li(sc, light, s.p, o, wi);
                if (lightPdf != 0) {
                    float weight = (scatteringPdf * scatteringPdf) / (scatteringPdf * scatteringPdf + lightPdf * lightPdf);
                    if (isEnv) {
                        // light.Le(ray) depends on the direction only: the probe just has to find no surface
                        V3 Li = env_Le(sc.env, wi);
                        if (!is_black(Li)) {
                            V3 Ld = div_each(f * Li * weight, scatteringPdf);
                            V3 c = beta * div_each(Ld, selPdf);
                            out.shB.o_tmax = make_float4(o.x, o.y, o.z, GNX_INF);
                            out.shB.d_path = make_float4(wi.x, wi.y, wi.z, i2f(slot));
                            out.shB.contrib = make_float4(c.x, c.y, c.z, 0.f);
                            out.haveShadowB = true;
                        }
                    } else {
                        // lightIsect.Le(-wi), counted only if the closest hit is this light's triangle
                        const TriVerts lt = load_tri(sc.tris, light.prim);
                        V3 nl = normalize(cross(lt.p0 - lt.p2, lt.p1 - lt.p2));
                        V3 Li = area_light_L(light, nl, -wi);
                        if (!is_black(Li)) {
                            V3 Ld = div_each(f * Li * weight, scatteringPdf);
                            V3 c = beta * div_each(Ld, selPdf);
                            out.pr.o_tmax = make_float4(o.x, o.y, o.z, GNX_INF);
                            out.pr.d_path = make_float4(wi.x, wi.y, wi.z, i2f(slot));
                            out.pr.contrib_expect = make_float4(c.x, c.y, c.z, i2f(light.prim));
                            out.haveProbe = true;
                        }
                    }
                }
            }
        } else if (!have) {
            live = false;
        }
    }
    GNX_STAGE_SYNC();
    // ---- the next direction (PathIntegrator.cpp:143-163): wi, f, pdf, flags are the P_CONT pass's
    const V3 wo = -rayD;
    if (!live || is_black(f) || pdf == 0.f) return;
    beta *= div_each(f * absdot(wi, ns), pdf);
    const bool spec = (flags & BSDF_SPECULAR) != 0;
    if (spec && (flags & BSDF_TRANSMISSION)) {
        float eta = bsdf.eta;
        etaScale *= (dot(wo, s.n) > 0) ? (eta * eta) : 1 / (eta * eta);
    }
    V3 o = offset_ray_origin(s.p, s.pError, s.n, wi);
    // ---- Russian roulette (PathIntegrator.cpp:198-204)
    V3 rrBeta = beta * etaScale;
    float mx = max_component(rrBeta);
    if (mx < rc.rr_threshold && bounces > 3) {
        float qq = fmaxf(.05f, 1 - mx);
        if (smp.get1d() < qq) return;
        beta = div_each(beta, 1 - qq);
    }
    out.alive = true;
    ps.ray_o[slot] = make_float4(o.x, o.y, o.z, GNX_INF);
    ps.ray_d[slot] = make_float4(wi.x, wi.y, wi.z, etaScale);
    ps.beta[slot] = make_float4(beta.x, beta.y, beta.z, 0.f);
    ps.meta[slot] = (uint32_t)(smp.dim & 0xffff) | ((uint32_t)(bounces + 1) << 16) | ((spec ? kFlagSpecular : 0u) << 24);
}

// Shadow / MIS-escape item (any-hit): the contribution is added when nothing is hit.
GNX_D void shadow_begin(const DeviceScene &sc, const ShadowItem *item, Trav &t) {
    const float4 o4 = ldg(&item->o_tmax), d4 = ldg(&item->d_path);
    trav_init(sc, t, V3(o4.x, o4.y, o4.z), V3(d4.x, d4.y, d4.z), o4.w);
}
GNX_D void shadow_finish(const PathState &ps, const ShadowItem *item, const Trav &t, bool toLb = false) {
    if (t.hit) return;
    const float4 c = ldg(&item->contrib);
    const int slot = f2i(ldg(&item->d_path).w);
    float4 *acc = toLb ? (ps.Lb ? ps.Lb : ps.L) : (ps.La ? ps.La : ps.L);
    float4 L = acc[slot];
    L.x += c.x; L.y += c.y; L.z += c.z;
    acc[slot] = L;
}
GNX_D void shadow_item(const DeviceScene &sc, const PathState &ps, const ShadowItem *item, int2 *stack, int stride,
                       TraversalCounters &cnt) {
    if (sc.nodes8 && sc.wide_any) {  // the compressed 8-wide tree answers the any-hit queries (gnx_bvh8.cuh)
        Trav8 t8;
        int2 store[kSpillStack];
        t8.spill = store;
        shadow_begin(sc, item, t8);
        trav8_init(sc, t8);
        while (!trav_done(t8)) {
            if (trav_needs_pop(t8)) trav8_next(sc, t8, stack, stride);
            else if (trav_is_leaf(t8)) t8.cur = trav_leaf_ref<true>(sc, t8, t8.cur, cnt) ? kRefNone : kRefPop;
            else trav8_interior<false>(sc, t8, stack, stride, cnt);
        }
        shadow_finish(ps, item, t8);
        return;
    }
    TravLocal t;
    shadow_begin(sc, item, t);
    while (!trav_step<true>(sc, t, stack, stride, cnt)) {}
    shadow_finish(ps, item, t);
}

// Area-light MIS probe (closest hit): counted only if the closest hit is the sampled light's triangle.
GNX_D void probe_begin(const DeviceScene &sc, const ProbeItem *item, Trav &t) {
    const float4 o4 = ldg(&item->o_tmax), d4 = ldg(&item->d_path);
    trav_init(sc, t, V3(o4.x, o4.y, o4.z), V3(d4.x, d4.y, d4.z), o4.w);
}
GNX_D void probe_finish(const PathState &ps, const ProbeItem *item, const Trav &t) {
    const float4 c = ldg(&item->contrib_expect);
    if (!(t.hit && t.prim == f2i(c.w))) return;
    const int slot = f2i(ldg(&item->d_path).w);
    float4 L = ps.L[slot];
    L.x += c.x; L.y += c.y; L.z += c.z;
    ps.L[slot] = L;
}
GNX_D void probe_item(const DeviceScene &sc, const PathState &ps, const ProbeItem *item, int2 *stack, int stride,
                      TraversalCounters &cnt) {
    TravLocal t;
    probe_begin(sc, item, t);
    while (!trav_step<false>(sc, t, stack, stride, cnt)) {}
    probe_finish(ps, item, t);
}

GNX_D int primary_hit_id(const DeviceScene &sc, int width, int px, int py, int sample, int2 *stack, int stride) {
    V3 o, d;
    float tMax;
    if (sc.smp.type == GNX_SAMPLER_PCG32) {
        PathSampler smp = PathSampler::stream(sc.smp, ((uint64_t)(width * py + px) << 20) | (uint64_t)sample);
        float u0, u1, l0, l1;
        smp.get2d(&u0, &u1);
        smp.get1d();
        smp.get2d(&l0, &l1);
        camera_ray_uv(sc, px, py, u0, u1, l0, l1, &o, &d, &tMax);
    } else {
        uint64_t hidx = sampler_index(sc.smp, px, py, (uint64_t)sample);
        camera_ray(sc, px, py, hidx, &o, &d, &tMax);
    }
    int prim = -1;
    TriHit h;
    TraversalCounters cnt{0, 0};
    bool hit = traverse<false>(sc, o, d, tMax, stack, stride, &prim, &h, cnt);
    return hit ? f2i(ldg(&sc.tris[3 * prim + 2].w)) : -1;
}

// SpatialLightDistribution::ComputeDistribution for one voxel (core/LightDistribution.cpp:206-274):
// 128 Halton points, each light's Li.y()/pdf summed, floor at 0.1% of the mean, then the
// Distribution1D construction of core/Sampling.h:22-35.
GNX_D void build_spatial_voxel(const DeviceScene &sc, int vox, float *func, float *cdf, float *fint) {
    const int nL = sc.n_lights;
    const int pi[3] = {vox % sc.ld.nvox[0], (vox / sc.ld.nvox[0]) % sc.ld.nvox[1], vox / (sc.ld.nvox[0] * sc.ld.nvox[1])};
    float lo[3], hi[3];
    for (int a = 0; a < 3; ++a) {
        float t0 = (float)pi[a] / (float)sc.ld.nvox[a], t1 = (float)(pi[a] + 1) / (float)sc.ld.nvox[a];
        lo[a] = lerpf(t0, sc.wb_min[a], sc.wb_max[a]);
        hi[a] = lerpf(t1, sc.wb_min[a], sc.wb_max[a]);
    }
    float *fv = func + (size_t)vox * nL;
    for (int j = 0; j < nL; ++j) fv[j] = 0.f;
    for (int i = 0; i < 128; ++i) {
        float r0 = radical_inverse(sc.smp, 0, i), r1 = radical_inverse(sc.smp, 1, i), r2 = radical_inverse(sc.smp, 2, i);
        V3 po(lerpf(r0, lo[0], hi[0]), lerpf(r1, lo[1], hi[1]), lerpf(r2, lo[2], hi[2]));
        float u0 = radical_inverse(sc.smp, 3, i), u1 = radical_inverse(sc.smp, 4, i);
        for (int j = 0; j < nL; ++j) {
            const gnx_light &l = sc.lights[j];
            LightSample ls;
            bool ok;
            if (l.type == GNX_LIGHT_INFINITE) ok = env_sample_li(sc.env, u0, u1, &ls);
            else if (l.type == GNX_LIGHT_AREA_TRI) ok = area_sample_li(sc, l, po, u0, u1, &ls);
            else {
                WLightSample wl;
                ok = w_sample_li(sc, l, po, u0, u1, &wl);
                ls.pdf = wl.pdf; ls.Li = wl.Li;
            }
            if (ok && ls.pdf > 0) fv[j] += lum_y(ls.Li) / ls.pdf;
        }
    }
    float sum = 0.f;
    for (int j = 0; j < nL; ++j) sum += fv[j];
    float avg = sum / (128 * nL);
    float minContrib = (avg > 0) ? .001f * avg : 1;
    for (int j = 0; j < nL; ++j) fv[j] = fmaxf(fv[j], minContrib);
    float *cv = cdf + (size_t)vox * (nL + 1);
    cv[0] = 0;
    for (int j = 1; j < nL + 1; ++j) cv[j] = cv[j - 1] + fv[j - 1] / nL;
    float funcInt = cv[nL];
    if (funcInt == 0) for (int j = 1; j < nL + 1; ++j) cv[j] = (float)j / (float)nL;
    else for (int j = 1; j < nL + 1; ++j) cv[j] /= funcInt;
    fint[vox] = funcInt;
}

}  // namespace gnx
