// host_emul.cpp — TEST TOOL: runs the product's per-path device functions (gnx_path.cuh and below,
// all __host__ __device__) sequentially on the CPU over a gnx_scene_desc.
//
// Purpose: the CPU-only test tier (`pytest -m "not gpu"`) can check the LOGIC of the kernels — Halton
// indexing, camera rays, BVH traversal, BSDF/light sampling, MIS, Russian roulette — against the
// reference without a GPU.  It is built with g++ (no nvcc) into tests/emul/_build/libgnxemul.so and
// is loaded only by tests/.  The product library neither links nor loads it, and nothing here
// is a fallback for libgnxrt.so.
#include <omp.h>

#include <cstdio>
#include <string>
#include <vector>

#include "gnx_pack.h"
#include "gnx_whitted.cuh"
#include "gnx_volwave.cuh"
#include "gnx_film.cuh"

using namespace gnx;

namespace {

struct EmulScene {
    DeviceScene sc{};
    std::vector<float4> tris, nodes2, env_texels;
    std::vector<uint4> nodes8;
    std::vector<int2> media;
    std::vector<DevMedium> dev_media;
    std::vector<DevTexture> textures;
    std::vector<float> uni_func, uni_cdf, pow_func, pow_cdf, power, sp_func, sp_cdf, sp_int, skybox_img;
    std::vector<int> primes, sums;
    std::vector<uint4> dims;
    std::vector<uint16_t> perms;
    unsigned typeMask = 0;
    bool hasNext = false;  // point / spot / distant / skybox lights present
    bool hasTex = false;   // image textures: the integrators with a RayDifferential carry it
    std::vector<std::vector<float>> pyramids;
    float ewa_lut[128];
    float wb[6];
    std::string err;
};

bool build(const gnx_scene_desc *d, EmulScene &e) {
    DeviceScene &sc = e.sc;
    const gnx_geometry &g = d->geom;
    if (!build_nodes(g.nodes, g.n_nodes, e.nodes2, &sc.n_nodes2, &e.err)) return false;
    sc.nodes2 = e.nodes2.data();
#if GNX_BVH_WIDTH == 2
    // the compressed 8-wide tree of the product (gnx_bvh8.cuh); GNX_ANYHIT_BVH8=0 keeps the two-child tree for the
    // any-hit queries, GNX_CLOSEST_BVH8=1 sends the closest-hit queries through the wide tree as well
    const char *a8 = getenv("GNX_ANYHIT_BVH8"), *c8 = getenv("GNX_CLOSEST_BVH8");
    const bool any8 = !(a8 && a8[0] == '0'), close8 = c8 && c8[0] == '1';  // the product's defaults
    if ((any8 || close8) && build_node8(e.nodes2.data(), sc.n_nodes2, e.nodes8) && !e.nodes8.empty()) {
        sc.nodes8 = e.nodes8.data();
        sc.n_nodes8 = (int)(e.nodes8.size() / kNode8Words);
        sc.wide_any = any8;
        sc.wide_closest = close8;
    }
#endif
    sc.n_nodes = g.n_nodes;
    sc.n_prims = g.n_prims;
    if (!pack_triangles(*d, e.tris, &e.typeMask, &e.err)) return false;
    sc.tris = e.tris.data();
    sc.tri_uv = g.prim_uv;
    sc.tri_n = (g.prim_n && g.prim_has_n) ? g.prim_n : nullptr;
    sc.tri_has_n = (g.prim_n && g.prim_has_n) ? g.prim_has_n : nullptr;
    for (int c = 0; c < 3; ++c) { sc.wb_min[c] = g.world_bound[c]; sc.wb_max[c] = g.world_bound[3 + c]; }
    memcpy(e.wb, g.world_bound, sizeof(e.wb));
    sc.materials = d->materials;
    sc.n_materials = d->n_materials;
    e.textures.resize(d->n_textures);
    e.pyramids.resize(d->n_textures);
    for (int i = 0; i < d->n_textures; ++i) {
        const gnx_texture &t = d->textures[i];
        std::vector<int> offs;
        int nLevels = 1;
        build_mip_pyramid(t, e.pyramids[i], offs, &nLevels);
        fill_dev_texture(t, e.pyramids[i].data(), offs, nLevels, e.textures[i]);
    }
    sc.textures = e.textures.data();
    make_ewa_lut(e.ewa_lut);
    sc.ewa_lut = e.ewa_lut;
    e.hasTex = d->n_textures > 0;
    sc.lights = d->lights;
    sc.n_lights = d->n_lights;
    for (int i = 0; i < d->n_lights; ++i) e.hasNext |= d->lights[i].type >= GNX_LIGHT_POINT;
    sc.light_nsamples = d->light_n_samples;
    if (g.prim_medium_in && g.prim_medium_out) {
        e.media.resize(g.n_prims);
        for (int k = 0; k < g.n_prims; ++k) e.media[k] = make_int2(g.prim_medium_in[k], g.prim_medium_out[k]);
        sc.tri_media = e.media.data();
        sc.tri_transition = g.prim_is_transition;
    }
    e.dev_media.resize(d->n_media);
    for (int i = 0; i < d->n_media; ++i) fill_dev_medium(d->media[i], d->media[i].density, e.dev_media[i]);
    sc.media = e.dev_media.data();
    sc.n_media = d->n_media;
    if (d->env.present) {
        const gnx_envmap &v = d->env;
        DevEnv &de = sc.env;
        de.present = 1; de.light_index = v.light_index;
        de.w = v.width; de.h = v.height; de.dw = v.dist_w; de.dh = v.dist_h;
        pack_env_texels(v.texels, (size_t)v.width * v.height, e.env_texels);
        de.texels = e.env_texels.data(); de.cond_func = v.cond_func; de.cond_cdf = v.cond_cdf; de.cond_int = v.cond_int;
        de.marg_func = v.marg_func; de.marg_cdf = v.marg_cdf; de.marg_int = v.marg_int;
        de.cond_guide = de.marg_guide = nullptr; de.cond_g = de.marg_g = 0;
        memcpy(de.l2w.m, v.light_to_world, 64);
        memcpy(de.w2l.m, v.world_to_light, 64);
        de.world_radius = v.world_radius;
    }
    sc.skybox.present = 0;
    if (d->skybox.present) {
        const gnx_skybox &sb = d->skybox;
        sc.skybox.present = 1; sc.skybox.light_index = sb.light_index;
        sc.skybox.w = sb.width; sc.skybox.h = sb.height; sc.skybox.nc = sb.channels;
        sc.skybox.center = V3(sb.center[0], sb.center[1], sb.center[2]);
        sc.skybox.radius = sb.radius;
        sc.skybox.data = nullptr;
        if (sb.data && sb.width > 0 && sb.height > 0 && sb.channels >= 3) {
            e.skybox_img.assign((size_t)sb.width * (sb.height + 1) * sb.channels + 4, 0.f);
            memcpy(e.skybox_img.data(), sb.data, sizeof(float) * (size_t)sb.width * sb.height * sb.channels);
            sc.skybox.data = e.skybox_img.data();
        }
    }
    if (d->n_lights > 0) {
        sc.ld.uni_int = uniform_light_distribution(d->n_lights, e.uni_func, e.uni_cdf);
        sc.ld.uni_func = e.uni_func.data();
        sc.ld.uni_cdf = e.uni_cdf.data();
        e.power.assign((size_t)d->n_lights, 0.f);
        for (int i = 0; i < d->n_lights; ++i) {
            if (d->light_power) e.power[i] = d->light_power[i];
            else derive_light_power(d->lights[i], &e.power[i]);
        }
    }
    sc.ld.mode = GNX_LIGHTS_UNIFORM;
    memcpy(sc.cam.r2c.m, d->camera.raster_to_camera, 64);
    memcpy(sc.cam.c2w.m, d->camera.camera_to_world, 64);
    sc.cam.lens_radius = d->camera.lens_radius;
    sc.cam.focal_distance = d->camera.focal_distance;
    sc.cam.medium = d->camera.medium;
    sc.cam.dx_camera = V3(d->camera.dx_camera[0], d->camera.dx_camera[1], d->camera.dx_camera[2]);
    sc.cam.dy_camera = V3(d->camera.dy_camera[0], d->camera.dy_camera[1], d->camera.dy_camera[2]);
    const gnx_sampler &s = d->sampler;
    sc.smp.spp = s.samples_per_pixel > 0 ? s.samples_per_pixel : 1;
    sc.smp.sobol32 = s.sobol_matrices32; sc.smp.sobol_vdc = s.sobol_vdc; sc.smp.sobol_vdc_inv = s.sobol_vdc_inv;
    sc.smp.sobol_dims = s.n_sobol_dimensions; sc.smp.sobol_log2res = s.sobol_log2_resolution; sc.smp.sobol_res = s.sobol_resolution;
    sc.smp.type = s.type;
    sc.smp.base_scale0 = s.base_scales[0]; sc.smp.base_scale1 = s.base_scales[1];
    sc.smp.base_exp0 = s.base_exponents[0]; sc.smp.base_exp1 = s.base_exponents[1];
    sc.smp.stride = s.sample_stride;
    sc.smp.mult_inv0 = s.mult_inverse[0]; sc.smp.mult_inv1 = s.mult_inverse[1];
    sc.smp.at_center = s.sample_at_pixel_center;
    sc.smp.stride_over_scale0 = s.base_scales[0] > 0 ? s.sample_stride / s.base_scales[0] : 0;
    sc.smp.stride_over_scale1 = s.base_scales[1] > 0 ? s.sample_stride / s.base_scales[1] : 0;
    make_primes(e.primes, e.sums);
    make_dim_table(e.primes, e.sums, e.dims);
    if (s.perms) e.perms.assign(s.perms, s.perms + s.n_perm_entries);
    else make_permutations(e.primes, e.perms);
    sc.smp.perms = e.perms.data();
    sc.smp.primes = e.primes.data();
    sc.smp.dims = e.dims.data();
    sc.smp.n_primes = (int)e.primes.size();
    return true;
}

void ensure_spatial(EmulScene &e, int strategy) {
    DeviceScene &sc = e.sc;
    if (sc.n_lights > 0) {
        sc.ld.uni_int = uniform_light_distribution(sc.n_lights, e.uni_func, e.uni_cdf);
        sc.ld.uni_func = e.uni_func.data(); sc.ld.uni_cdf = e.uni_cdf.data();
    }
    if (strategy == GNX_LIGHTS_UNIFORM || sc.n_lights <= 1) { sc.ld.mode = GNX_LIGHTS_UNIFORM; return; }
    if (strategy == GNX_LIGHTS_POWER) {  // the test scenes always carry light_power or derivable lights
        sc.ld.uni_int = power_light_distribution(sc.n_lights, e.power.data(), e.pow_func, e.pow_cdf);
        sc.ld.uni_func = e.pow_func.data(); sc.ld.uni_cdf = e.pow_cdf.data();
        sc.ld.mode = GNX_LIGHTS_POWER;
        return;
    }
    size_t nv = spatial_voxel_resolution(e.wb, sc.ld.nvox);
    e.sp_func.assign(nv * sc.n_lights, 0.f);
    e.sp_cdf.assign(nv * (sc.n_lights + 1), 0.f);
    e.sp_int.assign(nv, 0.f);
    sc.ld.mode = GNX_LIGHTS_SPATIAL;
    sc.ld.sp_func = e.sp_func.data(); sc.ld.sp_cdf = e.sp_cdf.data(); sc.ld.sp_int = e.sp_int.data();
#pragma omp parallel for schedule(dynamic, 256)
    for (long long v = 0; v < (long long)nv; ++v) build_spatial_voxel(sc, (int)v, e.sp_func.data(), e.sp_cdf.data(), e.sp_int.data());
}

// One camera sample carried through every wavefront stage, one slot.
V3 trace_sample(const EmulScene &e, const gnx_render_params &p, int px, int py, int sample, TraversalCounters &cnt,
                unsigned long long rays[3]) {
    const DeviceScene &sc = e.sc;
    if (p.integrator == GNX_INTEGRATOR_VOLPATH) {
        RenderConsts rcv{};
        rcv.width = p.width; rcv.height = p.height; rcv.max_depth = p.max_depth; rcv.rr_threshold = p.rr_threshold;
        int2 vstack[kSmemStack];
        VolCounters vc{0, 0, 0};
        V3 Lv;
        if (getenv("GNX_VOLPATH_MEGAKERNEL")) Lv = e.hasTex ? volpath_li<true>(sc, rcv, px, py, sample, vstack, 1, cnt, vc) : volpath_li<false>(sc, rcv, px, py, sample, vstack, 1, cnt, vc);
        else {
            // the staged wavefront's state machine (gnx_volwave.cuh), one slot: pixel / sample go in through the batch mapping
            float4 ro, rd, be, Lq, hq, so, sd, sh, st, w0, w1, w2, w3, w4;
            uint32_t hidx = 0, meta = 0;
            int32_t med = -1;
            uint2 rng = make_uint2(0, 0);
            float tmi = -1;
            PathState ps1{&ro, &rd, &be, &Lq, &hq, &hidx, &meta, &med};
            VolWave vw1{&rng, &tmi, &so, &sd, &sh, &st, &w0, &w1, &w2, &w3, &w4};
            rcv.batch_spp = 1; rcv.npix = p.width * p.height;
            // slot_to_sample maps slot -> (pixel = slot, sample = first_sample): render "slot" = pixel through offset pointers
            const int pix = py * p.width + px;
            rcv.first_sample = sample;
            PathState psq = ps1;
            psq.ray_o -= pix; psq.ray_d -= pix; psq.beta -= pix; psq.L -= pix; psq.hit -= pix; psq.hidx -= pix; psq.meta -= pix; psq.medium -= pix;
            VolWave vwq = vw1;
            vwq.rng -= pix; vwq.tmi -= pix; vwq.sub_o -= pix; vwq.sub_d -= pix; vwq.sub_hit -= pix; vwq.sub_tr -= pix;
            vwq.w0 -= pix; vwq.w1 -= pix; vwq.w2 -= pix; vwq.w3 -= pix; vwq.w4 -= pix;
            const bool staged = getenv("GNX_VOLWAVE_INPLACE") == nullptr;
            Lv = e.hasTex ? volwave_li<8, true>(sc, rcv, psq, vwq, pix, vstack, 1, cnt, vc, staged) : volwave_li<8, false>(sc, rcv, psq, vwq, pix, vstack, 1, cnt, vc, staged);
        }
        rays[0] += vc.extend; rays[1] += vc.shadow; rays[2] += vc.mis;
        return Lv;
    }
    if (p.integrator >= GNX_INTEGRATOR_WHITTED) {
        RenderConsts rcw{};
        rcw.width = p.width; rcw.height = p.height; rcw.max_depth = p.max_depth; rcw.rr_threshold = p.rr_threshold;
        int2 wstack[kSmemStack];
        RecCounters rcnt{0, 0, 0};
        V3 Lw = recursive_li<8>(sc, rcw, p.integrator - GNX_INTEGRATOR_WHITTED, px, py, sample, wstack, 1, cnt, rcnt, e.hasTex);
        rays[0] += rcnt.extend; rays[1] += rcnt.shadow; rays[2] += rcnt.mis;
        return Lw;
    }
    float4 ray_o, ray_d, beta, L, hit;
    uint32_t hidx, meta;
    int32_t medium = -1;
    PathState ps{&ray_o, &ray_d, &beta, &L, &hit, &hidx, &meta, &medium};
    RenderConsts rc{};
    rc.width = p.width; rc.height = p.height;
    rc.npix = 1;  // slot 0 only: pixel/sample are injected through first_sample below
    rc.max_depth = p.max_depth;
    rc.rr_threshold = p.rr_threshold;
    rc.batch_spp = 1;
    int2 stack[kSmemStack];
    // camera ray: generated and traversed in registers, like the fused primary kernel
    TravLocal t;
    V3 d;
    primary_begin(sc, px, py, sample, &hidx, &d, t);
    ++rays[0];
    closest_hit_run(sc, t, stack, 1, cnt);
    int type = primary_finish(sc, ps, rc, 0, hidx, d, t);
    if (type == kPendEscape) { escape_slot(sc, ps, 0); type = -1; }
    for (int iter = 0; iter < 100000 && type >= 0; ++iter) {
        ShadeOut out;
        out.alive = out.haveShadowA = out.haveShadowB = out.haveProbe = false;
        if (type == kNumShadeTypes - 1) out.alive = shade_null_slot(sc, ps, rc, 0);
        else if (e.hasNext) {
            if (type == GNX_MAT_DISNEY) shade_slot<8, false, true>(sc, ps, rc, 0, out);
            else shade_slot<2, false, true>(sc, ps, rc, 0, out);
        }
        else if (type == GNX_MAT_DISNEY) shade_slot<8>(sc, ps, rc, 0, out);
        else shade_slot<2>(sc, ps, rc, 0, out);
        if (out.haveShadowA) { ++rays[1]; shadow_item(sc, ps, &out.shA, stack, 1, cnt); }
        if (out.haveShadowB) { ++rays[2]; shadow_item(sc, ps, &out.shB, stack, 1, cnt); }
        if (out.haveProbe) { ++rays[2]; probe_item(sc, ps, &out.pr, stack, 1, cnt); }
        if (!out.alive) break;
        ++rays[0];
        type = extend_slot(sc, ps, rc, 0, stack, 1, cnt);
    }
    return V3(L.x, L.y, L.z);
}

}  // namespace

extern "C" {

void *gnxe_create(const gnx_scene_desc *d) {
    auto *e = new EmulScene;
    if (!build(d, *e)) fprintf(stderr, "[gnxe] %s\n", e->err.c_str());
    return e;
}
void gnxe_destroy(void *h) { delete (EmulScene *)h; }

int gnxe_render(void *h, const gnx_render_params *p, float *rgba_out, gnx_stats *stats) {
    auto *e = (EmulScene *)h;
    if (p->integrator < GNX_INTEGRATOR_WHITTED) ensure_spatial(*e, p->light_strategy);
    unsigned long long nodes = 0, tris = 0, r0 = 0, r1 = 0, r2 = 0;
    // Gaussian film (gnx_film.cuh): per-sample radiance and film offsets of the whole image as ONE batch, then the same
    // gather the kernels run
    const bool gaussian = p->film != GNX_FILM_BOX;
    std::vector<float4> Ls, offs;
    if (gaussian) { Ls.resize((size_t)p->width * p->height * p->spp); offs.resize(Ls.size()); }
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : nodes, tris, r0, r1, r2)
    for (int pixel = 0; pixel < p->width * p->height; ++pixel) {
        int px = pixel % p->width, py = pixel / p->width;
        V3 sum(0.f);
        TraversalCounters cnt{0, 0};
        unsigned long long rays[3] = {0, 0, 0};
        for (int s = 0; s < p->spp; ++s) {
            V3 L = trace_sample(*e, *p, px, py, p->first_sample + s, cnt, rays);
            sum += L;
            if (gaussian) {
                float u0, u1;
                film_sample_offset(e->sc, p->width, px, py, p->first_sample + s, &u0, &u1);
                Ls[(size_t)pixel * p->spp + s] = make_float4(L.x, L.y, L.z, 0.f);
                offs[(size_t)pixel * p->spp + s] = make_float4(u0, u1, 0.f, 0.f);
            }
        }
        float norm = (float)(p->spp_normalize > 0 ? p->spp_normalize : p->spp);
        rgba_out[4 * pixel + 0] = sum.x / norm;
        rgba_out[4 * pixel + 1] = sum.y / norm;
        rgba_out[4 * pixel + 2] = sum.z / norm;
        rgba_out[4 * pixel + 3] = 1.f;
        nodes += cnt.nodes; tris += cnt.tris; r0 += rays[0]; r1 += rays[1]; r2 += rays[2];
    }
    if (gaussian) {
        RenderConsts rc{};
        rc.width = p->width; rc.height = p->height; rc.npix = p->width * p->height; rc.batch_spp = p->spp;
        FilmFilter f{p->filter_radius, p->filter_alpha, std::exp(-p->filter_alpha * p->filter_radius * p->filter_radius),
                     (int)std::floor(p->filter_radius + 0.5f)};
#pragma omp parallel for schedule(dynamic, 16)
        for (int pixel = 0; pixel < rc.npix; ++pixel) {
            float4 a = gaussian_gather(Ls.data(), offs.data(), rc, f, pixel % rc.width, pixel / rc.width);
            if (p->film == GNX_FILM_GAUSSIAN) a = gaussian_resolve(a);
            memcpy(rgba_out + 4 * (size_t)pixel, &a, 16);
        }
    }
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->paths = (uint64_t)p->width * p->height * p->spp;
        stats->rays_extend = r0; stats->rays_shadow = r1; stats->rays_mis = r2;
        stats->nodes_visited = nodes; stats->tris_tested = tris;
    }
    return 0;
}

// One device's share of an N-device job, with the library's own partition arithmetic (gnx_path.cuh pixel_xy /
// local_tile_count, the sample-range split of gnx_render.cu): the frame a device hands to the reduce.  The sum of the
// shares' frames must be the single-device image — bit for bit with the tile partition.
int gnxe_render_share(void *h, const gnx_render_params *p, int n_shares, int share, float *rgba_out) {
    auto *e = (EmulScene *)h;
    if (p->integrator < GNX_INTEGRATOR_WHITTED) ensure_spatial(*e, p->light_strategy);
    const int W = p->width, H = p->height;
    memset(rgba_out, 0, sizeof(float) * 4 * (size_t)W * H);
    const float norm = (float)(p->spp_normalize > 0 ? p->spp_normalize : p->spp);
    if (p->partition == GNX_PARTITION_TILES) {
        RenderConsts rc{};
        rc.width = W; rc.height = H;
        rc.tile_n = n_shares; rc.tile_dev = share;
        rc.tiles_x = (W + kTile - 1) / kTile; rc.tiles_y = (H + kTile - 1) / kTile;
        const int npix = local_tile_count(rc.tiles_x, rc.tiles_y, n_shares, share) * kTile * kTile;
#pragma omp parallel for schedule(dynamic, 16)
        for (int lp = 0; lp < npix; ++lp) {
            int px, py;
            if (!pixel_xy(rc, lp, &px, &py)) continue;
            V3 sum(0.f);
            TraversalCounters cnt{0, 0};
            unsigned long long rays[3] = {0, 0, 0};
            for (int s = 0; s < p->spp; ++s) sum += trace_sample(*e, *p, px, py, p->first_sample + s, cnt, rays);
            float *o = rgba_out + 4 * ((size_t)py * W + px);
            o[0] = sum.x / norm; o[1] = sum.y / norm; o[2] = sum.z / norm; o[3] = 1.f;
        }
        return 0;
    }
    const int base = p->spp / n_shares, rem = p->spp % n_shares;
    const int count = base + (share < rem ? 1 : 0), first = p->first_sample + share * base + (share < rem ? share : rem);
#pragma omp parallel for schedule(dynamic, 16)
    for (int pixel = 0; pixel < W * H; ++pixel) {
        V3 sum(0.f);
        TraversalCounters cnt{0, 0};
        unsigned long long rays[3] = {0, 0, 0};
        for (int s = 0; s < count; ++s) sum += trace_sample(*e, *p, pixel % W, pixel / W, first + s, cnt, rays);
        float *o = rgba_out + 4 * (size_t)pixel;
        o[0] = sum.x / norm; o[1] = sum.y / norm; o[2] = sum.z / norm; o[3] = share == 0 ? 1.f : 0.f;
    }
    return 0;
}

int gnxe_samples(void *h, const gnx_render_params *p, int n, const int *px, const int *py, const int *sample, float *rgb_out) {
    auto *e = (EmulScene *)h;
    if (p->integrator < GNX_INTEGRATOR_WHITTED) ensure_spatial(*e, p->light_strategy);
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < n; ++i) {
        TraversalCounters cnt{0, 0};
        unsigned long long rays[3] = {0, 0, 0};
        V3 L = trace_sample(*e, *p, px[i], py[i], sample[i], cnt, rays);
        rgb_out[3 * i] = L.x; rgb_out[3 * i + 1] = L.y; rgb_out[3 * i + 2] = L.z;
    }
    return 0;
}

// gaussian_1d(x) * gaussian_1d(y) of the film code (gnx_film.cuh) at n points
int gnxe_gaussian_eval(float radius, float alpha, int n, const float *x, const float *y, float *out) {
    FilmFilter f{radius, alpha, std::exp(-alpha * radius * radius), (int)std::floor(radius + 0.5f)};
    for (int i = 0; i < n; ++i) out[i] = gaussian_1d(f, x[i]) * gaussian_1d(f, y[i]);
    return 0;
}

int gnxe_primary_hits(void *h, int width, int height, int sample, int *out) {
    auto *e = (EmulScene *)h;
#pragma omp parallel for schedule(dynamic, 256)
    for (int pixel = 0; pixel < width * height; ++pixel) {
        int2 stack[kSmemStack];
        out[pixel] = primary_hit_id(e->sc, width, pixel % width, pixel / width, sample, stack, 1);
    }
    return 0;
}

int gnxe_sample_dims(void *h, int n, const int64_t *index, const int *dim, float *out) {
    auto *e = (EmulScene *)h;
    for (int i = 0; i < n; ++i) out[i] = halton_sample_dimension(e->sc.smp, (uint64_t)index[i], dim[i]);
    return 0;
}

int64_t gnxe_sample_index(void *h, int px, int py, int sample) {
    auto *e = (EmulScene *)h;
    return (int64_t)sampler_index(e->sc.smp, px, py, (uint64_t)sample);
}

}  // extern "C"
