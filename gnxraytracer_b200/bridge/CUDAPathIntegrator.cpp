// CUDAPathIntegrator.cpp — flattens a live pbr::Scene into include/gnxrt.h buffers and drives
// libgnxrt.so.  See CUDAPathIntegrator.h and INTEGRATION.md.
//
// Nearly every field the flattening needs is private in the reference, and
// LinearBVHNode is defined inside accelerator/BVHAccel.cpp:54-65.  This one translation unit
// therefore includes the standard headers first and then re-includes the reference headers with
// the access keywords neutralised.  Access specifiers do not change the Itanium-ABI layout, so
// the objects stay link-compatible with the unmodified reference build.

// --- every standard header the reference pulls in, BEFORE the access override -----------------
#include <algorithm>
#include <array>
#include <atomic>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <iterator>
#include <limits>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <type_traits>
#include <typeinfo>
#include <unordered_map>
#include <utility>
#include <vector>
#include <omp.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <malloc.h>

#define private public
#define protected public
#include "accelerator/BVHAccel.h"
#include "camera/Perspective.h"
#include "core/Camera.h"
#include "core/Integrator.h"
#include "core/Light.h"
#include "core/MIPMap.h"
#include "core/Medium.h"
#include "core/Primitive.h"
#include "core/Sampling.h"
#include "core/Scene.h"
#include "core/Texture.h"
#include "lights/DiffuseAreaLight.h"
#include "lights/DistantLight.h"
#include "lights/InfiniteAreaLight.h"
#include "lights/PointLight.h"
#include "lights/SkyBoxLight.h"
#include "lights/SpotLight.h"
#include "materials/DisneyMaterial.h"
#include "materials/GlassMaterial.h"
#include "materials/MatteMaterial.h"
#include "materials/MetalMaterial.h"
#include "materials/MirrorMaterial.h"
#include "materials/PlasticMaterial.h"
#include "media/GridDensityMedium.h"
#include "media/HomogeneousMedium.h"
#include "samplers/HaltonSampler.h"
#include "shape/Triangle.h"
#include "textures/ConstantTexture.h"
#include "textures/ImageTexture.h"
#undef private
#undef protected

#include "gnxraytracer_b200/bridge/CUDAPathIntegrator.h"
#include "gnxraytracer_b200/bridge/SobolSampler.h"

namespace gnx {

using namespace pbr;

static_assert(sizeof(gnx_bvh_node) == 32, "gnx_bvh_node must mirror LinearBVHNode (32 bytes)");

namespace {

void CopyMatrix(const Matrix4x4 &m, float out[16]) {
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) out[4 * i + j] = m.m[i][j];
}

// The DFS layout puts a node's whole subtree in [i, end(i)) (accelerator/BVHAccel.cpp:628-646);
// BVHAccel does not keep the node count, so it is recovered by walking the right spine.
int CountNodes(const gnx_bvh_node *nodes) {
    int i = 0;
    while (nodes[i].n_prims == 0) i = nodes[i].offset;
    return i + 1;
}

struct Flattener {
    FlatScene *out;
    std::map<const Material *, int> materialIndex;
    std::map<const void *, int> textureIndex;
    std::map<const Medium *, int> mediumIndex;

    bool Fail(const std::string &why) {
        out->error = why;
        return false;
    }

    int MediumIndex(const Medium *m) {
        if (!m) return -1;
        auto it = mediumIndex.find(m);
        if (it != mediumIndex.end()) return it->second;
        gnx_medium gm{};
        if (auto *h = dynamic_cast<const HomogeneousMedium *>(m)) {
            gm.type = GNX_MEDIUM_HOMOGENEOUS;
            for (int c = 0; c < 3; ++c) { gm.sigma_a[c] = h->sigma_a[c]; gm.sigma_s[c] = h->sigma_s[c]; }
            gm.g = h->g;
        } else if (auto *g = dynamic_cast<const GridDensityMedium *>(m)) {
            gm.type = GNX_MEDIUM_GRID;
            for (int c = 0; c < 3; ++c) { gm.sigma_a[c] = g->sigma_a[c]; gm.sigma_s[c] = g->sigma_s[c]; }
            gm.g = g->g;
            gm.nx = g->nx; gm.ny = g->ny; gm.nz = g->nz;
            out->media_density.emplace_back(g->density.get(), g->density.get() + (size_t)g->nx * g->ny * g->nz);
            CopyMatrix(g->WorldToMedium.m, gm.world_to_medium);
            gm.inv_max_density = g->invMaxDensity;
        } else {
            out->error = "unsupported Medium subclass";
            return -2;
        }
        if (gm.type != GNX_MEDIUM_GRID) out->media_density.emplace_back();
        int idx = (int)out->media.size();
        out->media.push_back(gm);
        mediumIndex[m] = idx;
        return idx;
    }

    template <typename Tmem>
    int ImageTextureIndex(const void *key, const MIPMap<Tmem> *mip, const TextureMapping2D *mapping, int nChannels) {
        auto it = textureIndex.find(key);
        if (it != textureIndex.end()) return it->second;
        auto *uvmap = dynamic_cast<const UVMapping2D *>(mapping);
        if (!uvmap) { out->error = "only UVMapping2D texture mappings are supported"; return -2; }
        gnx_texture t{};
        t.width = mip->Width();
        t.height = mip->Height();
        t.n_channels = nChannels;
        t.n_levels = mip->Levels();
        t.wrap = mip->wrapMode == ImageWrap::Repeat ? GNX_WRAP_REPEAT
                 : mip->wrapMode == ImageWrap::Black ? GNX_WRAP_BLACK : GNX_WRAP_CLAMP;
        t.do_trilinear = mip->doTrilinear;
        t.max_aniso = mip->maxAnisotropy;
        t.su = uvmap->su; t.sv = uvmap->sv; t.du = uvmap->du; t.dv = uvmap->dv;
        std::vector<float> texels;
        for (int l = 0; l < t.n_levels; ++l) {
            const BlockedArray<Tmem> &lvl = *mip->pyramid[l];
            std::vector<Tmem> lin((size_t)lvl.uSize() * lvl.vSize());
            lvl.GetLinearArray(lin.data());
            const float *f = reinterpret_cast<const float *>(lin.data());
            texels.insert(texels.end(), f, f + lin.size() * nChannels);
        }
        out->texture_texels.push_back(std::move(texels));
        int idx = (int)out->textures.size();
        out->textures.push_back(t);
        textureIndex[key] = idx;
        return idx;
    }

    bool SpectrumParam(const std::shared_ptr<Texture<Spectrum>> &tex, float rgb[3], int32_t *texIdx) {
        *texIdx = -1;
        rgb[0] = rgb[1] = rgb[2] = 0;
        if (!tex) return true;
        if (auto *c = dynamic_cast<const ConstantTexture<Spectrum> *>(tex.get())) {
            for (int i = 0; i < 3; ++i) rgb[i] = c->value[i];
            return true;
        }
        if (auto *im = dynamic_cast<const ImageTexture<RGBSpectrum, Spectrum> *>(tex.get())) {
            static_assert(sizeof(RGBSpectrum) == 3 * sizeof(float), "RGBSpectrum must be 3 floats");
            *texIdx = ImageTextureIndex<RGBSpectrum>(tex.get(), im->mipmap, im->mapping.get(), 3);
            return *texIdx >= 0;
        }
        return Fail("unsupported Texture<Spectrum> subclass");
    }

    bool FloatParam(const std::shared_ptr<Texture<Float>> &tex, float *v, int32_t *texIdx) {
        *texIdx = -1;
        *v = 0;
        if (!tex) return true;
        if (auto *c = dynamic_cast<const ConstantTexture<Float> *>(tex.get())) {
            *v = c->value;
            return true;
        }
        if (auto *im = dynamic_cast<const ImageTexture<Float, Float> *>(tex.get())) {
            *texIdx = ImageTextureIndex<Float>(tex.get(), im->mipmap, im->mapping.get(), 1);
            return *texIdx >= 0;
        }
        return Fail("unsupported Texture<Float> subclass");
    }

    // A bump map that is a ConstantTexture displaces nothing, but Material::Bump still re-derives
    // the shading normal (core/Material.cpp:16-52); anything else is outside the hot path.
    bool BumpFlag(const std::shared_ptr<Texture<Float>> &bump, uint32_t *flags) {
        if (!bump) return true;
        if (dynamic_cast<const ConstantTexture<Float> *>(bump.get())) {
            *flags |= GNX_MATF_BUMP_IDENTITY;
            return true;
        }
        return Fail("non-constant bump maps are not supported");
    }

    int MaterialIndex(const Material *m) {
        if (!m) return -1;
        auto it = materialIndex.find(m);
        if (it != materialIndex.end()) return it->second;
        gnx_material g{};
        for (int i = 0; i < GNX_MAT_MAX_RGB; ++i) g.rgb_tex[i] = -1;
        for (int i = 0; i < GNX_MAT_MAX_F; ++i) g.f_tex[i] = -1;
        bool ok = true;
        if (auto *mt = dynamic_cast<const MatteMaterial *>(m)) {
            g.type = GNX_MAT_MATTE;
            ok = SpectrumParam(mt->Kd, g.rgb[0], &g.rgb_tex[0]) && FloatParam(mt->sigma, &g.f[0], &g.f_tex[0]) &&
                 BumpFlag(mt->bumpMap, &g.flags);
        } else if (auto *mr = dynamic_cast<const MirrorMaterial *>(m)) {
            g.type = GNX_MAT_MIRROR;
            ok = SpectrumParam(mr->Kr, g.rgb[0], &g.rgb_tex[0]) && BumpFlag(mr->bumpMap, &g.flags);
        } else if (auto *pl = dynamic_cast<const PlasticMaterial *>(m)) {
            g.type = GNX_MAT_PLASTIC;
            ok = SpectrumParam(pl->Kd, g.rgb[0], &g.rgb_tex[0]) && SpectrumParam(pl->Ks, g.rgb[1], &g.rgb_tex[1]) &&
                 FloatParam(pl->roughness, &g.f[0], &g.f_tex[0]) && BumpFlag(pl->bumpMap, &g.flags);
            if (pl->remapRoughness) g.flags |= GNX_MATF_REMAP_ROUGHNESS;
        } else if (auto *me = dynamic_cast<const MetalMaterial *>(m)) {
            g.type = GNX_MAT_METAL;
            // uRoughness / vRoughness fall back to roughness when absent (materials/MetalMaterial.cpp:36-39)
            ok = SpectrumParam(me->eta, g.rgb[0], &g.rgb_tex[0]) && SpectrumParam(me->k, g.rgb[1], &g.rgb_tex[1]) &&
                 FloatParam(me->uRoughness ? me->uRoughness : me->roughness, &g.f[0], &g.f_tex[0]) &&
                 FloatParam(me->vRoughness ? me->vRoughness : me->roughness, &g.f[1], &g.f_tex[1]) &&
                 BumpFlag(me->bumpMap, &g.flags);
            if (me->remapRoughness) g.flags |= GNX_MATF_REMAP_ROUGHNESS;
        } else if (auto *gl = dynamic_cast<const GlassMaterial *>(m)) {
            g.type = GNX_MAT_GLASS;
            ok = SpectrumParam(gl->Kr, g.rgb[0], &g.rgb_tex[0]) && SpectrumParam(gl->Kt, g.rgb[1], &g.rgb_tex[1]) &&
                 FloatParam(gl->uRoughness, &g.f[0], &g.f_tex[0]) && FloatParam(gl->vRoughness, &g.f[1], &g.f_tex[1]) &&
                 FloatParam(gl->index, &g.f[2], &g.f_tex[2]) && BumpFlag(gl->bumpMap, &g.flags);
            if (gl->remapRoughness) g.flags |= GNX_MATF_REMAP_ROUGHNESS;
        } else if (auto *ds = dynamic_cast<const DisneyMaterial *>(m)) {
            g.type = GNX_MAT_DISNEY;
            const std::shared_ptr<Texture<Float>> *fl[12] = {
                &ds->metallic, &ds->eta, &ds->roughness, &ds->specularTint, &ds->anisotropic, &ds->sheen,
                &ds->sheenTint, &ds->clearcoat, &ds->clearcoatGloss, &ds->specTrans, &ds->flatness, &ds->diffTrans};
            ok = SpectrumParam(ds->color, g.rgb[0], &g.rgb_tex[0]) &&
                 SpectrumParam(ds->scatterDistance, g.rgb[1], &g.rgb_tex[1]) && BumpFlag(ds->bumpMap, &g.flags);
            for (int i = 0; ok && i < 12; ++i) ok = FloatParam(*fl[i], &g.f[i], &g.f_tex[i]);
            if (ds->thin) g.flags |= GNX_MATF_THIN;
        } else {
            out->error = "unsupported Material subclass";
            return -2;
        }
        if (!ok) return -2;
        int idx = (int)out->materials.size();
        out->materials.push_back(g);
        materialIndex[m] = idx;
        return idx;
    }
};

}  // namespace

bool FlattenScene(const Scene &scene, const Camera &camera, const Sampler &sampler, FlatScene *out) {
    *out = FlatScene();
    Flattener fl{out};
    gnx_scene_desc &d = out->desc;
    std::memset(&d, 0, sizeof(d));
    d.abi_version = GNX_ABI_VERSION;

    // ---- geometry: BVHAccel::nodes / ::primitives ------------------------------------------
    auto *bvh = dynamic_cast<const BVHAccel *>(scene.aggregate.get());
    if (!bvh) return fl.Fail("Scene aggregate is not a BVHAccel");
    if (!bvh->nodes) return fl.Fail("empty BVH");
    const gnx_bvh_node *nodes = reinterpret_cast<const gnx_bvh_node *>(bvh->nodes);
    int nNodes = CountNodes(nodes);
    out->nodes.assign(nodes, nodes + nNodes);

    const size_t nPrims = bvh->primitives.size();
    out->prim_p.resize(nPrims * 9);
    out->prim_material.resize(nPrims);
    out->prim_light.assign(nPrims, -1);
    out->prim_medium_in.resize(nPrims);
    out->prim_medium_out.resize(nPrims);
    out->prim_is_transition.resize(nPrims);
    out->prim_flags.resize(nPrims);
    out->prim_id.resize(nPrims);
    out->prim_ptr.resize(nPrims);
    bool anyUV = false, anyN = false, anyMedia = false;
    std::map<const Shape *, int> shapeToPrim;
    for (size_t k = 0; k < nPrims; ++k) {
        auto *gp = dynamic_cast<const GeometricPrimitive *>(bvh->primitives[k].get());
        if (!gp) return fl.Fail("only GeometricPrimitive is supported");
        auto *tri = dynamic_cast<const Triangle *>(gp->shape.get());
        if (!tri) return fl.Fail("only Triangle shapes are supported (the reference Sphere is a stub)");
        const TriangleMesh &mesh = *tri->mesh;
        if (mesh.alphaMask || mesh.shadowAlphaMask) return fl.Fail("alpha masks are not supported");
        if (mesh.s) return fl.Fail("per-vertex tangents are not supported");
        if (mesh.uv) anyUV = true;
        if (mesh.n) anyN = true;
        for (int v = 0; v < 3; ++v)
            for (int c = 0; c < 3; ++c) out->prim_p[k * 9 + v * 3 + c] = mesh.p[tri->v[v]][c];
        int mi = fl.MaterialIndex(gp->material.get());
        if (mi == -2) return false;
        out->prim_material[k] = mi;
        int min = fl.MediumIndex(gp->mediumInterface.inside), mout = fl.MediumIndex(gp->mediumInterface.outside);
        if (min == -2 || mout == -2) return false;
        out->prim_medium_in[k] = min;
        out->prim_medium_out[k] = mout;
        out->prim_is_transition[k] = gp->mediumInterface.IsMediumTransition();
        if (min >= 0 || mout >= 0) anyMedia = true;
        out->prim_flags[k] = (uint8_t)(((tri->reverseOrientation ^ tri->transformSwapsHandedness) ? GNX_PRIM_FLIP_N : 0) |
                                       (tri->reverseOrientation ? GNX_PRIM_REVERSE_ORI : 0));
        out->prim_id[k] = (int32_t)k;
        out->prim_ptr[k] = bvh->primitives[k].get();
        shapeToPrim[gp->shape.get()] = (int)k;
    }
    if (anyUV) {
        out->prim_uv.resize(nPrims * 6);
        for (size_t k = 0; k < nPrims; ++k) {
            auto *tri = static_cast<const Triangle *>(static_cast<const GeometricPrimitive *>(bvh->primitives[k].get())->shape.get());
            Point2f uv[3];
            tri->GetUVs(uv);
            for (int v = 0; v < 3; ++v) { out->prim_uv[k * 6 + v * 2] = uv[v].x; out->prim_uv[k * 6 + v * 2 + 1] = uv[v].y; }
        }
    }
    if (anyN) {
        out->prim_n.assign(nPrims * 9, 0.f);
        out->prim_has_n.assign(nPrims, 0);
        for (size_t k = 0; k < nPrims; ++k) {
            auto *tri = static_cast<const Triangle *>(static_cast<const GeometricPrimitive *>(bvh->primitives[k].get())->shape.get());
            if (!tri->mesh->n) continue;
            out->prim_has_n[k] = 1;
            for (int v = 0; v < 3; ++v)
                for (int c = 0; c < 3; ++c) out->prim_n[k * 9 + v * 3 + c] = tri->mesh->n[tri->v[v]][c];
        }
    }
    gnx_geometry &g = d.geom;
    g.n_nodes = nNodes;
    g.nodes = out->nodes.data();
    g.n_prims = (int32_t)nPrims;
    g.prim_p = out->prim_p.data();
    g.prim_uv = anyUV ? out->prim_uv.data() : nullptr;
    g.prim_n = anyN ? out->prim_n.data() : nullptr;
    g.prim_has_n = anyN ? out->prim_has_n.data() : nullptr;
    g.prim_material = out->prim_material.data();
    g.prim_light = out->prim_light.data();
    g.prim_medium_in = anyMedia ? out->prim_medium_in.data() : nullptr;
    g.prim_medium_out = anyMedia ? out->prim_medium_out.data() : nullptr;
    g.prim_is_transition = anyMedia ? out->prim_is_transition.data() : nullptr;
    g.prim_flags = out->prim_flags.data();
    g.prim_id = out->prim_id.data();
    const Bounds3f &wb = scene.WorldBound();
    for (int c = 0; c < 3; ++c) { g.world_bound[c] = wb.pMin[c]; g.world_bound[3 + c] = wb.pMax[c]; }

    // ---- lights ---------------------------------------------------------------------------
    d.env.present = 0;
    d.skybox.present = 0;
    for (size_t i = 0; i < scene.lights.size(); ++i) {
        const Light *l = scene.lights[i].get();
        gnx_light gl{};
        gl.prim = -1;
        // SkyBoxLight hands its own, not yet constructed, mediumInterface member to the Light base class
        // (lights/SkyBoxLight.h:17): the base copy holds indeterminate pointers and must not be looked at.
        const bool skybox = dynamic_cast<const SkyBoxLight *>(l) != nullptr;
        gl.medium = skybox ? -1 : fl.MediumIndex(l->mediumInterface.inside);
        if (gl.medium == -2) return false;
        if (auto *al = dynamic_cast<const DiffuseAreaLight *>(l)) {
            gl.type = GNX_LIGHT_AREA_TRI;
            auto it = shapeToPrim.find(al->shape.get());
            if (it == shapeToPrim.end()) return fl.Fail("area light shape is not a scene primitive");
            gl.prim = it->second;
            gl.two_sided = al->twoSided;
            for (int c = 0; c < 3; ++c) gl.L[c] = al->Lemit[c];
            gl.area = al->area;
            out->prim_light[gl.prim] = (int32_t)i;
        } else if (auto *il = dynamic_cast<const InfiniteAreaLight *>(l)) {
            if (d.env.present) return fl.Fail("more than one InfiniteAreaLight");
            gl.type = GNX_LIGHT_INFINITE;
            gnx_envmap &e = d.env;
            e.present = 1;
            e.light_index = (int32_t)i;
            const BlockedArray<RGBSpectrum> &l0 = *il->Lmap->pyramid[0];
            e.width = l0.uSize();
            e.height = l0.vSize();
            std::vector<RGBSpectrum> lin((size_t)e.width * e.height);
            l0.GetLinearArray(lin.data());
            const float *f = reinterpret_cast<const float *>(lin.data());
            out->env_texels.assign(f, f + lin.size() * 3);
            const Distribution2D &dist = *il->distribution;
            e.dist_h = (int32_t)dist.pConditionalV.size();
            e.dist_w = dist.pConditionalV[0]->Count();
            for (int v = 0; v < e.dist_h; ++v) {
                const Distribution1D &row = *dist.pConditionalV[v];
                out->env_cond_func.insert(out->env_cond_func.end(), row.func.begin(), row.func.end());
                out->env_cond_cdf.insert(out->env_cond_cdf.end(), row.cdf.begin(), row.cdf.end());
                out->env_cond_int.push_back(row.funcInt);
            }
            out->env_marg_func = dist.pMarginal->func;
            out->env_marg_cdf = dist.pMarginal->cdf;
            e.marg_int = dist.pMarginal->funcInt;
            e.texels = out->env_texels.data();
            e.cond_func = out->env_cond_func.data();
            e.cond_cdf = out->env_cond_cdf.data();
            e.cond_int = out->env_cond_int.data();
            e.marg_func = out->env_marg_func.data();
            e.marg_cdf = out->env_marg_cdf.data();
            CopyMatrix(il->LightToWorld.m, e.light_to_world);
            CopyMatrix(il->WorldToLight.m, e.world_to_light);
            for (int c = 0; c < 3; ++c) e.world_center[c] = il->worldCenter[c];
            e.world_radius = il->worldRadius;
        } else if (auto *sl = dynamic_cast<const SpotLight *>(l)) {
            gl.type = GNX_LIGHT_SPOT;
            for (int c = 0; c < 3; ++c) { gl.L[c] = sl->I[c]; gl.p[c] = sl->pLight[c]; }
            gl.cos_total = sl->cosTotalWidth;
            gl.cos_falloff = sl->cosFalloffStart;
            CopyMatrix(sl->WorldToLight.m, gl.world_to_light);
        } else if (auto *pl = dynamic_cast<const PointLight *>(l)) {
            gl.type = GNX_LIGHT_POINT;
            for (int c = 0; c < 3; ++c) { gl.L[c] = pl->I[c]; gl.p[c] = pl->pLight[c]; }
        } else if (auto *dl = dynamic_cast<const DistantLight *>(l)) {
            gl.type = GNX_LIGHT_DISTANT;
            for (int c = 0; c < 3; ++c) { gl.L[c] = dl->L[c]; gl.p[c] = dl->wLight[c]; }
            gl.area = dl->worldRadius;  // set by DistantLight::Preprocess (Scene constructor)
        } else if (auto *sb = dynamic_cast<const SkyBoxLight *>(l)) {
            if (d.skybox.present) return fl.Fail("more than one SkyBoxLight");
            gl.type = GNX_LIGHT_SKYBOX;
            CopyMatrix(sb->LightToWorld.m, gl.world_to_light);
            gnx_skybox &k = d.skybox;
            k.present = 1;
            k.light_index = (int32_t)i;
            k.width = sb->imageWidth; k.height = sb->imageHeight; k.channels = sb->nrComponents;
            k.data = sb->data;  // owned by the light, alive as long as the scene
            for (int c = 0; c < 3; ++c) k.center[c] = sb->worldCenter[c];
            k.radius = sb->worldRadius;
        } else {
            return fl.Fail("unsupported Light subclass");
        }
        out->lights.push_back(gl);
        // the input of ComputeLightPowerDistribution (core/Integrator.cpp:216-217), from the light itself
        out->light_power.push_back(l->Power().y());
        out->light_n_samples.push_back(std::max(1, l->nSamples));  // Light::Light clamps the same way (core/Light.cpp)
    }
    d.n_lights = (int32_t)out->lights.size();
    d.lights = out->lights.data();
    d.light_power = out->light_power.data();
    d.light_n_samples = out->light_n_samples.data();

    d.n_materials = (int32_t)out->materials.size();
    d.materials = out->materials.data();

    // ---- camera ---------------------------------------------------------------------------
    auto *pc = dynamic_cast<const PerspectiveCamera *>(&camera);
    if (!pc) return fl.Fail("only PerspectiveCamera is supported");
    if (pc->CameraToWorld.actuallyAnimated) return fl.Fail("animated cameras are not supported");
    CopyMatrix(pc->RasterToCamera.m, d.camera.raster_to_camera);
    CopyMatrix(pc->CameraToWorld.startTransform->m, d.camera.camera_to_world);
    d.camera.lens_radius = pc->lensRadius;
    d.camera.focal_distance = pc->focalDistance;
    d.camera.shutter_open = pc->shutterOpen;
    d.camera.shutter_close = pc->shutterClose;
    for (int c = 0; c < 3; ++c) { d.camera.dx_camera[c] = pc->dxCamera[c]; d.camera.dy_camera[c] = pc->dyCamera[c]; }
    d.camera.medium = fl.MediumIndex(pc->medium);
    if (d.camera.medium == -2) return false;

    // media / textures are complete only now (camera and lights may reference media)
    for (size_t i = 0; i < out->media.size(); ++i)
        out->media[i].density = out->media_density[i].empty() ? nullptr : out->media_density[i].data();
    d.n_media = (int32_t)out->media.size();
    d.media = out->media.data();
    for (size_t i = 0; i < out->textures.size(); ++i) out->textures[i].texels = out->texture_texels[i].data();
    d.n_textures = (int32_t)out->textures.size();
    d.textures = out->textures.data();

    // ---- sampler --------------------------------------------------------------------------
    gnx_sampler &s = d.sampler;
    s.samples_per_pixel = (int32_t)sampler.samplesPerPixel;
    if (auto *hs = dynamic_cast<const HaltonSampler *>(&sampler)) {
        s.type = GNX_SAMPLER_HALTON;
        for (int i = 0; i < 2; ++i) {
            s.base_scales[i] = hs->baseScales[i];
            s.base_exponents[i] = hs->baseExponents[i];
            s.mult_inverse[i] = hs->multInverse[i];
        }
        s.sample_stride = hs->sampleStride;
        s.sample_at_pixel_center = hs->sampleAtPixelCenter;
        out->perms = HaltonSampler::radicalInversePermutations;
        s.n_perm_entries = (int32_t)out->perms.size();
        s.perms = out->perms.data();
    } else if (auto *sb = dynamic_cast<const SobolSampler *>(&sampler)) {
        // the reference's own tables, by pointer (samplers/SobolMatrices.cpp); nothing of them is copied into this repository
        if (sb->sampleBounds.pMin.x != 0 || sb->sampleBounds.pMin.y != 0) return fl.Fail("SobolSampler: sample bounds must start at (0, 0)");
        s.type = GNX_SAMPLER_SOBOL;
        s.sobol_resolution = sb->resolution;
        s.sobol_log2_resolution = sb->log2Resolution;
        s.n_sobol_dimensions = NumSobolDimensions;
        s.sobol_matrices32 = SobolMatrices32;
        s.sobol_vdc = sb->log2Resolution > 0 ? VdCSobolMatrices[sb->log2Resolution - 1] : nullptr;
        s.sobol_vdc_inv = sb->log2Resolution > 0 ? VdCSobolMatricesInv[sb->log2Resolution - 1] : nullptr;
    } else {
        // Any other sampler is mapped onto the per-pixel PCG32 stream; parity is then statistical.
        s.type = GNX_SAMPLER_PCG32;
    }
    return true;
}

// ------------------------------------------------------------------------------------------------

CUDAPathIntegrator::CUDAPathIntegrator(int maxDepth, std::shared_ptr<const Camera> camera,
                                       std::shared_ptr<Sampler> sampler, const Bounds2i &pixelBounds,
                                       FrameBuffer *pFrameBuffer, Float rrThreshold,
                                       const std::string &lightSampleStrategy, bool volumetric, int device,
                                       const std::vector<int> &devices)
    : maxDepth_(maxDepth),
      camera_(std::move(camera)),
      sampler_(std::move(sampler)),
      pixelBounds_(pixelBounds),
      fb_(pFrameBuffer),
      rrThreshold_(rrThreshold),
      lightSampleStrategy_(lightSampleStrategy),
      integrator_(volumetric ? GNX_INTEGRATOR_VOLPATH : GNX_INTEGRATOR_PATH) {
    // the GPUs of this integrator: the constructor argument, else GNX_DEVICES ("all" or a comma-separated list), else `device`
    std::vector<int> ids = devices;
    if (ids.empty())
        if (const char *env = getenv("GNX_DEVICES")) {
            if (!strcmp(env, "all")) for (int g = 0; g < gnx_device_count(); ++g) ids.push_back(g);
            else for (const char *c = env; *c;) {
                char *end = nullptr;
                long v = strtol(c, &end, 10);
                if (end == c) break;
                ids.push_back((int)v);
                c = *end ? end + 1 : end;
            }
        }
    int rc = ids.size() > 1 ? gnx_create_multi(&ctx_, ids.data(), (int)ids.size()) : gnx_create(&ctx_, ids.empty() ? device : ids[0]);
    if (rc != GNX_OK) {
        error_ = std::string("gnx_create failed: ") + gnx_last_error(nullptr);
        ctx_ = nullptr;
    }
}

CUDAPathIntegrator::~CUDAPathIntegrator() {
    if (ctx_) gnx_destroy(ctx_);
}

gnx_render_params CUDAPathIntegrator::MakeParams() const {
    gnx_render_params p{};
    p.width = pixelBounds_.pMax.x;
    p.height = pixelBounds_.pMax.y;
    p.spp = (int32_t)sampler_->samplesPerPixel;
    p.first_sample = 0;
    p.spp_normalize = 0;
    p.max_depth = maxDepth_;
    p.rr_threshold = rrThreshold_;
    p.integrator = integrator_;
    // CreateLightSampleDistribution (core/LightDistribution.cpp:15-33); the single-light override
    // is applied inside the library, which knows the light count.
    p.light_strategy = lightSampleStrategy_ == "uniform" ? GNX_LIGHTS_UNIFORM
                       : lightSampleStrategy_ == "power" ? GNX_LIGHTS_POWER : GNX_LIGHTS_SPATIAL;
    p.film = filterRadius_ > 0 ? GNX_FILM_GAUSSIAN : GNX_FILM_BOX;
    p.filter_radius = filterRadius_; p.filter_alpha = filterAlpha_;
    p.partition = partition_;
    return p;
}

// What identifies an uploaded scene besides its address: a Scene rebuilt at the same address, another light list,
// another sampler or camera all change it (edits INSIDE those objects need Invalidate()).
static size_t SceneFingerprint(const Scene &scene, const Camera *camera, const Sampler *sampler) {
    size_t h = 1469598103934665603ull;
    auto mix = [&](size_t v) { h = (h ^ v) * 1099511628211ull; };
    mix((size_t)scene.aggregate.get());
    mix(scene.lights.size());
    for (const auto &l : scene.lights) mix((size_t)l.get());
    mix((size_t)camera);
    mix((size_t)sampler);
    mix((size_t)sampler->samplesPerPixel);
    return h;
}

bool CUDAPathIntegrator::EnsureUploaded(const Scene &scene) {
    if (!ctx_) return false;
    const size_t fp = SceneFingerprint(scene, camera_.get(), sampler_.get());
    if (uploaded_ == &scene && fingerprint_ == fp) return true;
    uploaded_ = nullptr;
    flat_.reset(new FlatScene);
    if (!FlattenScene(scene, *camera_, *sampler_, flat_.get())) {
        error_ = "FlattenScene: " + flat_->error;
        return false;
    }
    if (gnx_upload_scene(ctx_, &flat_->desc) != GNX_OK) {
        error_ = std::string("gnx_upload_scene: ") + gnx_last_error(ctx_);
        return false;
    }
    uploaded_ = &scene;
    fingerprint_ = fp;
    return true;
}

void CUDAPathIntegrator::Render(const Scene &scene, double &timeConsume) {
    timeConsume = 0;
    if (!EnsureUploaded(scene)) {
        // The reference has no error channel (void Render); leave the FrameBuffer untouched.
        fprintf(stderr, "[CUDAPathIntegrator] %s\n", error_.c_str());
        return;
    }
    gnx_render_params p = MakeParams();
    auto t0 = std::chrono::steady_clock::now();
    // Same sink as core/Integrator.cpp:230,307-310: one more pass in the FrameBuffer's running mean.  The mean and its
    // tonemapped 8-bit copy are computed on the device and land in FrameBuffer::fbuffer / ubuffer directly (this
    // translation unit sees the private members); a FrameBuffer of another shape goes through the setters.
    fb_->renderCountIncrease();
    const int pass = fb_->curRenderCount;
    if (progressive_) p.first_sample = (pass - 1) * p.spp;
    const bool direct = fb_->fbuffer && fb_->ubuffer && fb_->width == p.width && fb_->height >= p.height && fb_->channals == 4;
    int rc;
    std::vector<float> rgba;
    if (direct) rc = gnx_render_framebuffer(ctx_, &p, pass, fb_->fbuffer, fb_->ubuffer, &stats_);
    else {
        rgba.resize((size_t)p.width * p.height * 4);
        rc = gnx_render(ctx_, &p, rgba.data(), &stats_);
    }
    if (rc != GNX_OK) {
        error_ = std::string(direct ? "gnx_render_framebuffer: " : "gnx_render: ") + gnx_last_error(ctx_);
        fprintf(stderr, "[CUDAPathIntegrator] %s\n", error_.c_str());
        fb_->curRenderCount = pass - 1;  // the pass did not happen
        return;
    }
    if (!direct)
        for (int j = 0; j < p.height; ++j)
            for (int i = 0; i < p.width; ++i) {
                const float *px = &rgba[((size_t)i + (size_t)j * p.width) * 4];
                fb_->update_f_u_c(i, j, 0, px[0]);
                fb_->update_f_u_c(i, j, 1, px[1]);
                fb_->update_f_u_c(i, j, 2, px[2]);
                fb_->set_uc(i, j, 3, 255);
            }
    timeConsume = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

bool CUDAPathIntegrator::PrimaryHits(const Scene &scene, int sample, std::vector<int32_t> *ordered) {
    if (!EnsureUploaded(scene)) return false;
    gnx_render_params p = MakeParams();
    ordered->assign((size_t)p.width * p.height, -1);
    if (gnx_primary_hits(ctx_, &p, sample, ordered->data()) != GNX_OK) {
        error_ = std::string("gnx_primary_hits: ") + gnx_last_error(ctx_);
        return false;
    }
    return true;
}

}  // namespace gnx
