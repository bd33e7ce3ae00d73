// ref_harness.cpp — TEST INFRASTRUCTURE ONLY (oracle side).
//
// Builds the BASELINE.json configs through the reference's own class API (the UI's scene code in
// ui/ModelList.cpp, ui/MaterialList.cpp and ui/RenderThread.cpp cannot be compiled without Qt, so
// its recipes are re-expressed here with the same numbers), renders them with the UNMODIFIED
// reference integrators and exposes parity hooks over a small C interface used by tests/ and by
// bench.py's cpu_baseline / --impl reference legs via ctypes.
//
// Linked into oracle/_ref/libgnxref.so together with the reference objects (oracle/Makefile); that library links nothing
// of the product (the drop-in class is instantiated on these scenes by oracle/bridge_harness.cpp, libgnxbridge.so).
// Nothing in the product library depends on this file.
#include <dlfcn.h>
#include <omp.h>

#include <cstdio>
#include <cstring>
#include <fstream>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "accelerator/BVHAccel.h"
#include "camera/Perspective.h"
#include "core/Integrator.h"
#include "core/Scene.h"
#include "core/Transform.h"
#include "filters/GaussianFilter.h"
#include "integrators/PathIntegrator.h"
#include "integrators/VolPathIntegrator.h"
#include "integrators/DirectLightingIntegrator.h"
#include "integrators/WhittedIntegrator.h"
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
#include "shape/plyRead.h"
#include "textures/ConstantTexture.h"
#include "textures/ImageTexture.h"

#include "ref_harness.h"
#include "gnxraytracer_b200/bridge/SobolSampler.h"  // header-only, built from the reference's own Sobol helpers and tables
#include "gnxraytracer_b200/host/scenekit_mesh.h"
#include "gnxraytracer_b200/host/scenekit_io.h"

using namespace pbr;

// lights/SkyBoxLight.cpp:21 switches stb_image's process-wide "flip vertically" flag on and never off again; every
// image the reference loads afterwards (environment maps, textures) would come out upside down.  The harness builds
// many scenes per process, so it puts the flag back once a SkyBoxLight has been constructed.
extern "C" void stbi_set_flip_vertically_on_load(int flag_true_if_should_flip);

namespace {

std::string ResourceDir() {
    if (const char *e = getenv("GNX_RESOURCES")) return std::string(e) + "/";
    Dl_info info;
    if (dladdr((void *)&ResourceDir, &info) && info.dli_fname) {
        std::string p = info.dli_fname;
        size_t s = p.find_last_of('/');
        return p.substr(0, s) + "/Resources/";
    }
    return "Resources/";
}

// A Sampler for integrators whose number of draws per sample is unbounded (VolPath + GridDensityMedium:
// the reference's Halton tables stop at dimension 1000 and are read out of range beyond that, SURVEY.md
// §8a-14).  One PCG32 stream (core/RNG.h) per camera sample, sequence id = (seed << 20) | sampleNumber where
// seed is what SamplerIntegrator::Render passes to Clone(): W * y + x (core/Integrator.cpp:264,268).
// The product's GNX_SAMPLER_PCG32 uses the same convention (gnx_sampler.cuh, PathSampler::stream).
class PcgStreamSampler : public Sampler {
  public:
    PcgStreamSampler(int64_t spp, int seed = 0) : Sampler(spp), base(seed) {}
    void StartPixel(const Point2i &p) override { Sampler::StartPixel(p); Reseed(0); }
    bool StartNextSample() override { bool more = Sampler::StartNextSample(); Reseed(currentPixelSampleIndex); return more; }
    bool SetSampleNumber(int64_t n) override { bool ok = Sampler::SetSampleNumber(n); Reseed(n); return ok; }
    Float Get1D() override { return rng.UniformFloat(); }
    Point2f Get2D() override { Float a = rng.UniformFloat(); Float b = rng.UniformFloat(); return Point2f(a, b); }
    std::unique_ptr<Sampler> Clone(int seed) override { return std::unique_ptr<Sampler>(new PcgStreamSampler(samplesPerPixel, seed)); }

  private:
    void Reseed(int64_t s) { rng.SetSequence(((uint64_t)base << 20) | (uint64_t)s); }
    int base;
    RNG rng;
};

std::shared_ptr<Texture<Spectrum>> ConstSpec(float r, float g, float b) {
    Spectrum s;
    s[0] = r; s[1] = g; s[2] = b;
    return std::make_shared<ConstantTexture<Spectrum>>(s);
}
std::shared_ptr<Texture<Float>> ConstF(float v) { return std::make_shared<ConstantTexture<Float>>(v); }

// TriangleMesh + one Triangle/GeometricPrimitive per face, as ui/ModelList.cpp:49-69 does.
void AddMesh(HarnessScene &hs, const gnxsk::Mesh &m, const Transform &o2w, std::shared_ptr<Material> material,
             const Spectrum *emit, const MediumInterface &mi = MediumInterface()) {
    const Transform *O2W = hs.keep(o2w), *W2O = hs.keep(Inverse(o2w));
    int nv = m.nVerts(), nt = m.nTris();
    std::vector<Point3f> P(nv);
    for (int i = 0; i < nv; ++i) P[i] = Point3f(m.P[3 * i], m.P[3 * i + 1], m.P[3 * i + 2]);
    std::vector<Normal3f> N;
    if (!m.N.empty()) {
        N.resize(nv);
        for (int i = 0; i < nv; ++i) N[i] = Normal3f(m.N[3 * i], m.N[3 * i + 1], m.N[3 * i + 2]);
    }
    std::vector<Point2f> UV;
    if (!m.UV.empty()) {
        UV.resize(nv);
        for (int i = 0; i < nv; ++i) UV[i] = Point2f(m.UV[2 * i], m.UV[2 * i + 1]);
    }
    auto mesh = std::make_shared<TriangleMesh>(*O2W, nt, m.idx.data(), nv, P.data(), nullptr,
                                               N.empty() ? nullptr : N.data(), UV.empty() ? nullptr : UV.data(), nullptr);
    for (int i = 0; i < nt; ++i) {
        std::shared_ptr<Shape> tri = std::make_shared<Triangle>(O2W, W2O, false, mesh, i);
        std::shared_ptr<AreaLight> area;
        if (emit) {
            // one DiffuseAreaLight per triangle, nSamples 5, one-sided (ui/ModelList.cpp:140-146)
            area = std::make_shared<DiffuseAreaLight>(*O2W, MediumInterface(), *emit, 5, tri, false);
            hs.lights.push_back(area);
        }
        hs.prims.push_back(std::make_shared<GeometricPrimitive>(tri, material, area, mi));
    }
}

void SetupCamera(HarnessScene &hs, const Point3f &eye, const Point3f &look) {
    // ui/RenderThread.cpp:60-68
    Transform lookat = LookAt(eye, look, Vector3f(0.0f, 1.0f, 0.0f));
    hs.cam2world.reset(new Transform(Inverse(lookat)));
    hs.animated.reset(new AnimatedTransform(hs.cam2world.get(), 0.0f, hs.cam2world.get(), 1.0f));
    hs.camera = std::shared_ptr<Camera>(CreatePerspectiveCamera(hs.width, hs.height, *hs.animated));
}

void Finish(HarnessScene &hs, bool pcgSampler = false) {
    double t0 = omp_get_wtime();
    auto bvh = std::make_shared<BVHAccel>(hs.prims, 1);  // ui/RenderThread.cpp:155
    hs.bvhSeconds = omp_get_wtime() - t0;
    hs.scene.reset(new Scene(bvh, hs.lights));
    Bounds2i bounds(Point2i(0, 0), Point2i(hs.width, hs.height));
    if (pcgSampler) hs.sampler = std::make_shared<PcgStreamSampler>(hs.spp);
    else hs.sampler = std::make_shared<HaltonSampler>(hs.spp, bounds, false);  // ui/RenderThread.cpp:159
    hs.fb.reset(new FrameBuffer);
    hs.fb->InitBuffer(hs.width, hs.height, 4);
    for (size_t i = 0; i < hs.prims.size(); ++i) hs.originalIndex[hs.prims[i].get()] = (int)i;
}

std::shared_ptr<Material> Matte(float r, float g, float b, float sigma) {
    // constant-0 bump map, as every UI material has (ui/RenderThread.cpp:90-99)
    return std::make_shared<MatteMaterial>(ConstSpec(r, g, b), ConstF(sigma), ConstF(0.0f));
}

// Config 1: Cornell box, two icospheres (Mirror, Glass), two-triangle area light.
// variant: 0 = Lambert walls (sigma 0), 1 = the UI's Oren-Nayar sigma 60, 2 = Lambert + a second emitter; sphere subdivision in p1.
void BuildCornell(HarnessScene &hs, int variant, int subdiv) {
    float sigma = variant == 1 ? 60.0f : 0.0f;
    auto white = Matte(0.91f, 0.91f, 0.91f, sigma);
    auto red = Matte(0.9f, 0.1f, 0.17f, sigma);
    auto blue = Matte(0.14f, 0.21f, 0.87f, sigma);
    if (subdiv >= 0) {
        auto mirror = std::make_shared<MirrorMaterial>(ConstSpec(0.9f, 0.9f, 0.9f), ConstF(0.0f));
        auto glass = std::make_shared<GlassMaterial>(ConstSpec(0.98f, 0.98f, 0.98f), ConstSpec(0.98f, 0.98f, 0.98f),
                                                     ConstF(0.0f), ConstF(0.0f), ConstF(1.5f), ConstF(0.0f), false);
        AddMesh(hs, gnxsk::icosphere(subdiv, 0.8f, -1.0f, -1.7f, -0.5f), Transform(), mirror, nullptr);
        AddMesh(hs, gnxsk::icosphere(subdiv, 0.8f, 1.0f, -1.7f, 0.8f), Transform(), glass, nullptr);
    }
    // walls: triangles 6,7 red, 8,9 blue, others white (ui/ModelList.cpp:116-124)
    gnxsk::Mesh walls = gnxsk::cornell_walls(5.0f);
    Transform box2world = Translate(Vector3f(-2.5f, -2.5f, -2.5f));
    for (int i = 0; i < 10; ++i) {
        gnxsk::Mesh one;
        for (int v = 0; v < 3; ++v) {
            one.P.insert(one.P.end(), {walls.P[9 * i + 3 * v], walls.P[9 * i + 3 * v + 1], walls.P[9 * i + 3 * v + 2]});
            one.idx.push_back(v);
        }
        AddMesh(hs, one, box2world, (i == 6 || i == 7) ? red : (i == 8 || i == 9) ? blue : white, nullptr);
    }
    Spectrum Le(5.0f);
    AddMesh(hs, gnxsk::area_light_quad(1.4f), Translate(Vector3f(0.0f, 2.45f, 0.0f)), white, &Le);
    if (variant == 2) {  // a second, smaller and differently coloured emitter: unequal Light::Power() for "power" sampling
        Spectrum Le2; Le2[0] = 9.0f; Le2[1] = 3.0f; Le2[2] = 1.0f;
        AddMesh(hs, gnxsk::area_light_quad(0.6f), Translate(Vector3f(-1.5f, 2.45f, 1.2f)), white, &Le2);
    }
    SetupCamera(hs, Point3f(0.f, 0.f, 5.0f), Point3f(0.f, 0.f, 0.0f));
    Finish(hs);
}

// the Cornell room (open towards the camera) with a Mirror and a Glass sphere and a plastic
// box, lit by a PointLight, a SpotLight, a DistantLight, the area light and a SkyBoxLight (visible through the
// opening and in the mirror).  lightMask selects the lights (bit 0 area, 1 point, 2 spot, 3 distant, 4 skybox,
// 5 = skybox with the awesomeface.jpg image instead of the procedural colours).
// textured: 0 = constant colours; 1 / 2 = the white walls (floor, ceiling, back wall) carry awesomeface.jpg as an
// ImageTexture filtered with EWA (1) or trilinearly (2), and the spheres get per-vertex normals — Whitted / DirectLighting
// then exercise ray differentials: ComputeDifferentials, MIPMap::Lookup, the offset rays through SpecularReflect /
// SpecularTransmit (core/Interaction.cpp:65-114, core/MIPMap.h:226-337, core/Integrator.cpp:321-442).
void BuildLightsRoom(HarnessScene &hs, int lightMask, int subdiv, int textured = 0) {
    std::shared_ptr<Material> white = Matte(0.91f, 0.91f, 0.91f, 0.f);
    if (textured) {
        std::unique_ptr<TextureMapping2D> map = std::make_unique<UVMapping2D>(2.f, 3.f, 0.1f, 0.2f);
        std::shared_ptr<Texture<Spectrum>> face = std::make_shared<ImageTexture<RGBSpectrum, Spectrum>>(
            std::move(map), ResourceDir() + "awesomeface.jpg", textured == 2, 8.f, ImageWrap::Repeat, 1.f, false);
        white = std::make_shared<MatteMaterial>(face, ConstF(0.f), ConstF(0.0f));
    }
    auto red = Matte(0.9f, 0.1f, 0.17f, 0.f);
    auto blue = Matte(0.14f, 0.21f, 0.87f, 30.f);
    auto mirror = std::make_shared<MirrorMaterial>(ConstSpec(0.9f, 0.9f, 0.9f), ConstF(0.0f));
    auto glass = std::make_shared<GlassMaterial>(ConstSpec(0.98f, 0.98f, 0.98f), ConstSpec(0.98f, 0.98f, 0.98f), ConstF(0.0f),
                                                 ConstF(0.0f), ConstF(1.5f), ConstF(0.0f), false);
    auto plastic = std::make_shared<PlasticMaterial>(ConstSpec(0.35f, 0.12f, 0.48f), ConstSpec(0.65f, 0.88f, 0.52f), ConstF(0.1f),
                                                     ConstF(0.0f), true);
    AddMesh(hs, gnxsk::icosphere(subdiv, 0.8f, -1.0f, -1.7f, -0.5f, textured != 0), Transform(), mirror, nullptr);
    AddMesh(hs, gnxsk::icosphere(subdiv, 0.8f, 1.0f, -1.7f, 0.8f, textured != 0), Transform(), glass, nullptr);
    AddMesh(hs, gnxsk::icosphere(subdiv > 1 ? subdiv - 1 : subdiv, 0.5f, 0.0f, -2.0f, 1.6f), Transform(), plastic, nullptr);
    gnxsk::Mesh walls = gnxsk::cornell_walls(5.0f);
    Transform box2world = Translate(Vector3f(-2.5f, -2.5f, -2.5f));
    for (int i = 0; i < 10; ++i) {
        gnxsk::Mesh one;
        for (int v = 0; v < 3; ++v) {
            one.P.insert(one.P.end(), {walls.P[9 * i + 3 * v], walls.P[9 * i + 3 * v + 1], walls.P[9 * i + 3 * v + 2]});
            one.idx.push_back(v);
        }
        AddMesh(hs, one, box2world, (i == 6 || i == 7) ? red : (i == 8 || i == 9) ? blue : white, nullptr);
    }
    if (lightMask & 1) {
        Spectrum Le(3.0f);
        AddMesh(hs, gnxsk::area_light_quad(1.4f), Translate(Vector3f(0.0f, 2.45f, 0.0f)), white, &Le);
    }
    if (lightMask & 2) {
        Spectrum I;
        I[0] = 6.f; I[1] = 5.f; I[2] = 4.f;
        hs.lights.push_back(std::make_shared<PointLight>(Translate(Vector3f(1.5f, 1.8f, 1.0f)), MediumInterface(), I));
    }
    if (lightMask & 4) {
        Spectrum I;
        I[0] = 9.f; I[1] = 14.f; I[2] = 18.f;
        Transform l2w = Inverse(LookAt(Point3f(-1.8f, 2.0f, 1.8f), Point3f(0.2f, -2.0f, 0.0f), Vector3f(0.f, 1.f, 0.f)));
        hs.lights.push_back(std::make_shared<SpotLight>(l2w, MediumInterface(), I, 32.f, 22.f));
    }
    if (lightMask & 8) {
        Spectrum L;
        L[0] = 0.5f; L[1] = 0.45f; L[2] = 0.35f;
        hs.lights.push_back(std::make_shared<DistantLight>(RotateY(15), L, Vector3f(0.3f, 0.4f, 1.0f)));
    }
    if (lightMask & (16 | 32)) {
        std::string img = (lightMask & 32) ? ResourceDir() + "awesomeface.jpg" : std::string("/nonexistent");
        // ui/RenderThread.cpp:145: SkyBoxLight(transform, worldCenter, worldRadius, file, nSamples)
        hs.lights.push_back(std::make_shared<SkyBoxLight>(RotateX(10), Point3f(0.f, 0.f, 0.f), 50.f, img.c_str(), 1));
        stbi_set_flip_vertically_on_load(0);
    }
    SetupCamera(hs, Point3f(0.f, 0.f, 6.5f), Point3f(0.f, -0.4f, 0.0f));
    Finish(hs);
}

std::shared_ptr<Material> PurplePlastic() {  // ui/MaterialList.cpp:48-56
    return std::make_shared<PlasticMaterial>(ConstSpec(0.35f, 0.12f, 0.48f), ConstSpec(1.f - 0.35f, 1.f - 0.12f, 1.f - 0.48f),
                                             ConstF(0.1f), ConstF(0.0f), true);
}
std::shared_ptr<Material> YellowMetal() {  // ui/MaterialList.cpp:58-69
    return std::make_shared<MetalMaterial>(ConstSpec(0.2f, 0.2f, 0.8f), ConstSpec(0.11f, 0.11f, 0.11f), ConstF(0.15f),
                                           ConstF(0.15f), ConstF(0.15f), ConstF(0.0f), false);
}

// Config 2: dragon-class mesh (torus knot stand-in for the stripped dragon.3d) under the MonValley
// environment light.  variant 0 = Plastic, 1 = Metal; nu x nv quads.
// file3d non-empty: the mesh is read by the reference's OWN reader, plyInfo, exactly as ui/ModelList.cpp:49-69 does.
bool BuildDragon(HarnessScene &hs, int variant, int nu, int nv, const std::string &hdr, const std::string &file3d = "") {
    if (!file3d.empty()) {
        FILE *fp = fopen(file3d.c_str(), "rb");
        if (!fp) { hs.error = "missing mesh " + file3d; return false; }
        fclose(fp);
        plyInfo plyi(file3d);
        gnxsk::Mesh m;
        for (int i = 0; i < plyi.nVertices; ++i) for (int c = 0; c < 3; ++c) m.P.push_back(plyi.vertexArray[i][c]);
        m.idx.assign(plyi.vertexIndices, plyi.vertexIndices + 3 * plyi.nTriangles);
        plyi.Release();
        AddMesh(hs, m, Translate(Vector3f(0.f, -2.9f, 0.f)), variant == 1 ? YellowMetal() : PurplePlastic(), nullptr);
    } else {
        gnxsk::Mesh knot = gnxsk::torus_knot(nu, nv);
        for (float &x : knot.P) x *= 20;  // plyInfo scales every vertex by 20 (shape/plyRead.h:38)
        AddMesh(hs, knot, Translate(Vector3f(0.f, -2.9f, 0.f)), variant == 1 ? YellowMetal() : PurplePlastic(), nullptr);
    }
    std::string path = ResourceDir() + hdr;
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) { hs.error = "missing resource " + path; return false; }
    fclose(f);
    Transform l2w = RotateX(20) * RotateY(-90) * RotateX(-90);  // ui/ModelList.cpp:172-178
    hs.lights.push_back(std::make_shared<InfiniteAreaLight>(l2w, Spectrum(1.0f), 10, path));
    SetupCamera(hs, Point3f(0.f, 0.f, 5.0f), Point3f(0.f, 0.f, 0.0f));
    Finish(hs);
    return true;
}

// A Wavefront OBJ (+ MTL) through the reference's classes: the reference has no OBJ loader, so the
// file is read by the scene kit's reader and handed to TriangleMesh / GeometricPrimitive exactly like ui/ModelList.cpp:49-69
// does for dragon.3d — one TriangleMesh per material, the MTL entries mapped onto MatteMaterial / PlasticMaterial /
// MirrorMaterial / GlassMaterial by the same recipe the kit uses (gnxsk::material_recipe).  Placement and light as the kit's
// "obj:" scene: fitted into a sphere of radius 2.5 around (0, -0.4, 0), MonValley environment.
bool BuildObj(HarnessScene &hs, int variant, const std::string &path) {
    gnxsk::Mesh file;
    std::vector<gnxsk::ObjMaterial> objMats;
    if (!gnxsk::load_obj(path, &file, &hs.error, &objMats)) return false;
    if (file.nTris() == 0) { hs.error = "mesh file without triangles"; return false; }
    const float centre[3] = {0.f, -0.4f, 0.f};
    gnxsk::fit_to_sphere(&file, 2.5f, centre);
    std::vector<std::shared_ptr<Material>> mats;
    mats.push_back(variant == 1 ? YellowMetal() : PurplePlastic());  // faces without usemtl
    for (const gnxsk::ObjMaterial &om : objMats) {
        const gnxsk::MaterialRecipe r = gnxsk::material_recipe(om);
        if (r.kind == gnxsk::MaterialRecipe::Matte) mats.push_back(Matte(r.kd[0], r.kd[1], r.kd[2], 0.f));
        else if (r.kind == gnxsk::MaterialRecipe::Plastic)
            mats.push_back(std::make_shared<PlasticMaterial>(ConstSpec(r.kd[0], r.kd[1], r.kd[2]), ConstSpec(r.ks[0], r.ks[1], r.ks[2]), ConstF(r.roughness), ConstF(0.0f), false));
        else if (r.kind == gnxsk::MaterialRecipe::Mirror) mats.push_back(std::make_shared<MirrorMaterial>(ConstSpec(r.ks[0], r.ks[1], r.ks[2]), ConstF(0.0f)));
        else mats.push_back(std::make_shared<GlassMaterial>(ConstSpec(r.ks[0], r.ks[1], r.ks[2]), ConstSpec(r.ks[0], r.ks[1], r.ks[2]), ConstF(0.0f), ConstF(0.0f), ConstF(r.index), ConstF(0.0f), false));
    }
    // one sub-mesh per material, faces in file order (the scene's primitive order is file order within each material)
    for (size_t k = 0; k < mats.size(); ++k) {
        gnxsk::Mesh sub;
        sub.P = file.P; sub.N = file.N; sub.UV = file.UV;
        for (int f = 0; f < file.nTris(); ++f) {
            const int mi = file.tri_material.empty() ? 0 : file.tri_material[f] + 1;
            if (mi == (int)k) for (int v = 0; v < 3; ++v) sub.idx.push_back(file.idx[3 * f + v]);
        }
        if (sub.nTris() > 0) AddMesh(hs, sub, Transform(), mats[k], nullptr);
    }
    std::string hdr = ResourceDir() + "MonValley1000.hdr";
    FILE *f = fopen(hdr.c_str(), "rb");
    if (!f) { hs.error = "missing resource " + hdr; return false; }
    fclose(f);
    Transform l2w = RotateX(20) * RotateY(-90) * RotateX(-90);
    hs.lights.push_back(std::make_shared<InfiniteAreaLight>(l2w, Spectrum(1.0f), 10, hdr));
    SetupCamera(hs, Point3f(0.f, 0.f, 5.0f), Point3f(0.f, 0.f, 0.0f));
    Finish(hs);
    return true;
}

// Config 3: textured mesh with per-vertex UVs and normals (stand-in for the stripped nanosuit: there is no
// OBJ loader in the reference), DisneyMaterial whose colour is an ImageTexture on awesomeface.jpg exactly
// as getSmileFacePlasticMaterial builds it (ui/MaterialList.cpp:31-46), TropicalRuins environment.
// variant 0 = the SURVEY §8d constants, 1 = thin surface with transmission lobes; nu x nv knot quads.
bool BuildNano(HarnessScene &hs, int variant, int nu, int nv) {
    std::string tex = ResourceDir() + "awesomeface.jpg", hdr = ResourceDir() + "TropicalRuins1000.hdr";
    for (const std::string &f : {tex, hdr}) {
        FILE *fp = fopen(f.c_str(), "rb");
        if (!fp) { hs.error = "missing resource " + f; return false; }
        fclose(fp);
    }
    std::unique_ptr<TextureMapping2D> map = std::make_unique<UVMapping2D>(1.f, 1.f, 0.f, 0.f);
    std::shared_ptr<Texture<Spectrum>> color =
        std::make_shared<ImageTexture<RGBSpectrum, Spectrum>>(std::move(map), tex, false, 8.f, ImageWrap::Repeat, 1.f, false);
    bool thin = variant == 1;
    auto disney = std::make_shared<DisneyMaterial>(
        color, ConstF(0.2f) /*metallic*/, ConstF(1.5f) /*eta*/, ConstF(0.4f) /*roughness*/, ConstF(0.f) /*specularTint*/,
        ConstF(thin ? 0.3f : 0.f) /*anisotropic*/, ConstF(0.5f) /*sheen*/, ConstF(0.5f) /*sheenTint*/, ConstF(0.5f) /*clearcoat*/,
        ConstF(0.8f) /*clearcoatGloss*/, ConstF(thin ? 0.4f : 0.f) /*specTrans*/, ConstSpec(0.f, 0.f, 0.f) /*scatterDistance*/, thin,
        ConstF(thin ? 0.3f : 0.f) /*flatness*/, ConstF(thin ? 0.5f : 0.f) /*diffTrans*/, nullptr /*bumpMap*/);
    gnxsk::Mesh knot = gnxsk::torus_knot(nu, nv, 1.0f, true, true);
    for (float &x : knot.P) x *= 20;
    AddMesh(hs, knot, Translate(Vector3f(0.f, -2.9f, 0.f)), disney, nullptr);
    AddMesh(hs, gnxsk::uv_sphere(std::max(8, nu / 4), std::max(6, nv), 0.9f, -2.6f, -1.2f, 0.6f), Transform(), disney, nullptr);
    AddMesh(hs, gnxsk::uv_sphere(std::max(8, nu / 4), std::max(6, nv), 0.7f, 2.7f, 1.4f, -0.4f), Transform(), disney, nullptr);
    Transform l2w = RotateX(20) * RotateY(-90) * RotateX(-90);
    hs.lights.push_back(std::make_shared<InfiniteAreaLight>(l2w, Spectrum(1.0f), 10, hdr));
    SetupCamera(hs, Point3f(0.f, 0.f, 5.0f), Point3f(0.f, 0.f, 0.0f));
    Finish(hs);
    return true;
}

// Config 4: VolPathIntegrator, GridDensityMedium(density_render.70.volume) inside a HomogeneousMedium "fog"
// box, both bounded by material-less triangles carrying MediumInterfaces, a Matte ground quad, MonValley
// environment.  There is no .volume loader in the reference: the text file is parsed here (header tokens
// nx ny nz / p0 / p1 / sigma_a / sigma_s, then nx*ny*nz floats, x fastest).
// variant 0 = grid + fog (PCG stream sampler), 1 = fog only with the Halton sampler (bounded dimensions).
bool BuildSmoke(HarnessScene &hs, int variant) {
    std::string hdr = ResourceDir() + "MonValley1000.hdr", vol = ResourceDir() + "density_render.70.volume";
    FILE *fp = fopen(hdr.c_str(), "rb");
    if (!fp) { hs.error = "missing resource " + hdr; return false; }
    fclose(fp);
    auto fog = std::make_shared<HomogeneousMedium>(Spectrum(0.02f), Spectrum(0.08f), 0.5f);
    hs.media.push_back(fog);
    const Medium *inner = fog.get();
    if (variant == 0) {
        std::ifstream f(vol);
        if (!f) { hs.error = "missing resource " + vol; return false; }
        std::string tok;
        int nx = 0, ny = 0, nz = 0;
        float p0[3], p1[3], sa[3], ss[3];
        f >> tok >> nx >> tok >> ny >> tok >> nz;
        f >> tok >> p0[0] >> p0[1] >> p0[2] >> tok >> p1[0] >> p1[1] >> p1[2];
        f >> tok >> sa[0] >> sa[1] >> sa[2] >> tok >> ss[0] >> ss[1] >> ss[2];
        std::vector<Float> dens((size_t)nx * ny * nz);
        for (Float &d : dens) f >> d;
        if (!f || nx <= 0) { hs.error = "cannot parse " + vol; return false; }
        // centre the medium box on the origin
        Transform m2w = Translate(Vector3f(-1.f, -1.f, -0.4f)) * Translate(Vector3f(p0[0], p0[1], p0[2])) *
                        Scale(p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]);
        auto grid = std::make_shared<GridDensityMedium>(Spectrum(sa[0]), Spectrum(ss[0]), 0.f, nx, ny, nz, m2w, dens.data());
        hs.media.push_back(grid);
        const float lo[3] = {-1.f + p0[0] - 0.01f, -1.f + p0[1] - 0.01f, -0.4f + p0[2] - 0.01f};
        const float hi[3] = {-1.f + p1[0] + 0.01f, -1.f + p1[1] + 0.01f, -0.4f + p1[2] + 0.01f};
        AddMesh(hs, gnxsk::box(lo, hi), Transform(), nullptr, nullptr, MediumInterface(grid.get(), fog.get()));
    }
    const float flo[3] = {-2.4f, -2.4f, -2.4f}, fhi[3] = {2.4f, 2.4f, 2.4f};
    AddMesh(hs, gnxsk::box(flo, fhi), Transform(), nullptr, nullptr, MediumInterface(inner, nullptr));
    AddMesh(hs, gnxsk::ground_quad(2.2f, -1.5f), Transform(), Matte(0.5f, 0.5f, 0.5f, 0.f), nullptr, MediumInterface(inner));
    if (variant >= 2) {
        // variant 2 / 3: a large textured floor below the fog box, seen by the camera DIRECTLY in front of the box: the one
        // place where VolPathIntegrator still holds the camera's RayDifferential (EWA / trilinear MIPMap::Lookup)
        std::unique_ptr<TextureMapping2D> map = std::make_unique<UVMapping2D>(3.f, 3.f, 0.f, 0.f);
        std::shared_ptr<Texture<Spectrum>> face = std::make_shared<ImageTexture<RGBSpectrum, Spectrum>>(
            std::move(map), ResourceDir() + "awesomeface.jpg", variant == 3, 8.f, ImageWrap::Repeat, 1.f, false);
        AddMesh(hs, gnxsk::ground_quad(9.0f, -2.45f), Transform(), std::make_shared<MatteMaterial>(face, ConstF(0.f), ConstF(0.0f)), nullptr);
    }
    Transform l2w = RotateX(20) * RotateY(-90) * RotateX(-90);
    hs.lights.push_back(std::make_shared<InfiniteAreaLight>(l2w, Spectrum(1.0f), 10, hdr));
    SetupCamera(hs, Point3f(0.f, 0.f, 5.0f), Point3f(0.f, 0.f, 0.0f));
    hs.integrator = 1;
    Finish(hs, variant == 0);  // the grid medium needs the unbounded PCG stream; the fog-only variants run on Halton
    return true;
}

// The UI's live scene (ui/RenderThread.cpp:60-164): the mesh (Matte sigma 60, green) INSIDE the Cornell box whose five
// walls are Oren-Nayar sigma 60, the two-triangle area light carrying the mesh's material (ui/ModelList.cpp:140-146), a
// SkyBoxLight of radius 10 whose image file "1" does not exist (procedural colours through the box's open front),
// HaltonSampler, and the integrator of ui/RenderThread.cpp:163 (Whitted, maxDepth 5) or :164 (Path, maxDepth 15).
// The camera looks into the open front of the box: every camera ray hits a surface (100 % coverage).
// dragon.3d is stripped from the reference snapshot: the mesh is the same torus-knot stand-in as config 2.
bool BuildUI(HarnessScene &hs, int integrator, int nu, int nv, const std::string &file3d = "") {
    const float sigma = 60.0f;
    auto dragonMat = Matte(0.2f, 0.8f, 0.2f, sigma);
    auto white = Matte(0.91f, 0.91f, 0.91f, sigma);
    auto red = Matte(0.9f, 0.1f, 0.17f, sigma);
    auto blue = Matte(0.14f, 0.21f, 0.87f, sigma);
    if (!file3d.empty()) {
        FILE *fp = fopen(file3d.c_str(), "rb");
        if (!fp) { hs.error = "missing mesh " + file3d; return false; }
        fclose(fp);
        plyInfo plyi(file3d);
        gnxsk::Mesh m;
        for (int i = 0; i < plyi.nVertices; ++i) for (int c = 0; c < 3; ++c) m.P.push_back(plyi.vertexArray[i][c]);
        m.idx.assign(plyi.vertexIndices, plyi.vertexIndices + 3 * plyi.nTriangles);
        plyi.Release();
        AddMesh(hs, m, Translate(Vector3f(0.f, -2.9f, 0.f)), dragonMat, nullptr);
    } else {
        gnxsk::Mesh knot = gnxsk::torus_knot(nu, nv);
        for (float &x : knot.P) x *= 20;
        AddMesh(hs, knot, Translate(Vector3f(0.f, -2.9f, 0.f)), dragonMat, nullptr);
    }
    gnxsk::Mesh walls = gnxsk::cornell_walls(5.0f);
    Transform box2world = Translate(Vector3f(-2.5f, -2.5f, -2.5f));
    for (int i = 0; i < 10; ++i) {
        gnxsk::Mesh one;
        for (int v = 0; v < 3; ++v) {
            one.P.insert(one.P.end(), {walls.P[9 * i + 3 * v], walls.P[9 * i + 3 * v + 1], walls.P[9 * i + 3 * v + 2]});
            one.idx.push_back(v);
        }
        AddMesh(hs, one, box2world, (i == 6 || i == 7) ? red : (i == 8 || i == 9) ? blue : white, nullptr);
    }
    Spectrum Le(5.0f);
    AddMesh(hs, gnxsk::area_light_quad(1.4f), Translate(Vector3f(0.0f, 2.45f, 0.0f)), dragonMat, &Le);
    // ui/ModelList.cpp:163-170: SkyBoxLight(Transform(), (0,0,0), 10, "1", 1)
    hs.lights.push_back(std::make_shared<SkyBoxLight>(Transform(), Point3f(0.f, 0.f, 0.f), 10.0f, "1", 1));
    stbi_set_flip_vertically_on_load(0);
    SetupCamera(hs, Point3f(0.f, 0.f, 5.0f), Point3f(0.f, 0.f, 0.0f));
    hs.integrator = (integrator == 2 || integrator == 3) ? integrator : 0;
    Finish(hs);
    return true;
}

}  // namespace

extern "C" {

// name: "cornell" (p0 = variant, p1 = sphere subdivision, -1 = no spheres)
//       "dragon"  (p0 = variant, p1 = nu, p2 = nv)
//       "nano"    (p0 = variant, p1 = nu, p2 = nv)
//       "smoke"   (p0 = variant)   -> VolPathIntegrator
//       "ui"      (p0 = gnx_integrator: 0 Path, 2 Whitted, 3 DirectLighting; p1 = nu, p2 = nv)  the UI's live scene
//       "whitted" / "direct" (p0 = light mask, p1 = sphere subdivision) -> WhittedIntegrator / DirectLightingIntegrator
void *gnxh_scene_create(const char *name, int width, int height, int spp, int p0, int p1, int p2) {
    auto *hs = new HarnessScene;
    hs->name = name;
    hs->width = width; hs->height = height; hs->spp = spp;
    if (hs->name == "cornell") BuildCornell(*hs, p0, p1);
    else if (hs->name == "dragon") BuildDragon(*hs, p0, p1, p2, "MonValley1000.hdr");
    else if (hs->name.rfind("dragon3d:", 0) == 0) BuildDragon(*hs, p0, 0, 0, "MonValley1000.hdr", hs->name.substr(9));
    else if (hs->name.rfind("obj:", 0) == 0) BuildObj(*hs, p0, hs->name.substr(4));
    else if (hs->name == "nano") BuildNano(*hs, p0, p1, p2);
    else if (hs->name == "smoke") BuildSmoke(*hs, p0);
    else if (hs->name == "ui") BuildUI(*hs, p0, p1 > 0 ? p1 : 2048, p2 > 0 ? p2 : 213);
    else if (hs->name.rfind("ui3d:", 0) == 0) BuildUI(*hs, p0, 0, 0, hs->name.substr(5));
    else if (hs->name == "lights_path") {
        // the same room under the wavefront PathIntegrator: delta lights + SkyBoxLight through EstimateDirect
        BuildLightsRoom(*hs, p0 > 0 ? p0 : 31, p1 > 0 ? p1 : 2);
        hs->integrator = 0;
    }
    else if (hs->name == "whitted_tex" || hs->name == "direct_tex") {
        // p2: 1 = EWA, 2 = trilinear filtering of the wall texture
        BuildLightsRoom(*hs, p0 > 0 ? p0 : 31, p1 > 0 ? p1 : 2, p2 == 2 ? 2 : 1);
        hs->integrator = hs->name == "whitted_tex" ? 2 : 3;
    }
    else if (hs->name == "whitted" || hs->name == "direct" || hs->name == "lights") {
        // "lights": the scene kit's name for the same room, p2 = gnx_integrator (2 Whitted, 3 DirectLighting)
        BuildLightsRoom(*hs, p0 > 0 ? p0 : 31, p1 > 0 ? p1 : 2);
        hs->integrator = hs->name == "whitted" ? 2 : hs->name == "direct" ? (p2 == 4 ? 4 : 3) : (p2 == 3 || p2 == 4 ? p2 : 2);
    }
    else hs->error = "unknown scene";
    return hs;
}
const char *gnxh_scene_error(void *h) { return ((HarnessScene *)h)->error.c_str(); }
void gnxh_scene_destroy(void *h) { delete (HarnessScene *)h; }
int gnxh_scene_num_prims(void *h) { return (int)((HarnessScene *)h)->prims.size(); }
// lightSampleStrategy of both integrators: 0 "uniform", 1 "spatial" (default), 2 "power"
void gnxh_scene_set_light_strategy(void *h, int strategy) {
    auto *hs = (HarnessScene *)h;
    hs->strategy = strategy == 0 ? "uniform" : strategy == 2 ? "power" : "spatial";
    hs->DropExt();
}
// Image reconstruction of the drop-in class: GaussianFilter(radius, alpha); radius <= 0 = the reference's box average.
void gnxh_scene_set_gaussian_filter(void *h, float radius, float alpha) {
    auto *hs = (HarnessScene *)h;
    hs->filterRadius = radius; hs->filterAlpha = alpha;
    hs->DropExt();
}
double gnxh_scene_bvh_seconds(void *h) { return ((HarnessScene *)h)->bvhSeconds; }
// The sampler handed to both integrators: 0 = HaltonSampler (the UI's, ui/RenderThread.cpp:159), 2 = the Sobol'
// GlobalSampler of gnxraytracer_b200/bridge/SobolSampler.h.  (The PCG stream sampler belongs to the grid-medium scene.)
void gnxh_scene_set_sampler(void *h, int kind) {
    auto *hs = (HarnessScene *)h;
    Bounds2i bounds(Point2i(0, 0), Point2i(hs->width, hs->height));
    if (kind == 2) hs->sampler = std::make_shared<gnx::SobolSampler>(hs->spp, bounds);
    else hs->sampler = std::make_shared<HaltonSampler>(hs->spp, bounds, false);
    hs->DropExt();
}

// Every integrator gets a fresh copy of the scene's sampler: DirectLightingIntegrator::Preprocess(UniformSampleAll) appends
// sample-array requests to the sampler it is given, and a second Preprocess on the same object would append them again.
static SamplerIntegrator *MakeReferenceIntegrator(HarnessScene *hs, int maxDepth, std::shared_ptr<Sampler> *samplerOut = nullptr) {
    Bounds2i bounds(Point2i(0, 0), Point2i(hs->width, hs->height));
    std::shared_ptr<Sampler> fresh(hs->sampler->Clone(0).release());
    if (samplerOut) *samplerOut = fresh;
    struct Swap { std::shared_ptr<Sampler> &a, b; Swap(std::shared_ptr<Sampler> &x, std::shared_ptr<Sampler> y) : a(x), b(x) { a = y; } ~Swap() { a = b; } } swap(hs->sampler, fresh);
    switch (hs->integrator) {
    case 4: return new DirectLightingIntegrator(LightStrategy::UniformSampleAll, maxDepth, hs->camera, hs->sampler, bounds, hs->fb.get());
    case 1: return new VolPathIntegrator(maxDepth, hs->camera, hs->sampler, bounds, 1.f, hs->strategy, hs->fb.get());
    case 2: return new WhittedIntegrator(maxDepth, hs->camera, hs->sampler, bounds, hs->fb.get());
    case 3: return new DirectLightingIntegrator(LightStrategy::UniformSampleOne, maxDepth, hs->camera, hs->sampler, bounds, hs->fb.get());
    default: return new PathIntegrator(maxDepth, hs->camera, hs->sampler, bounds, hs->fb.get(), 1.f, hs->strategy);
    }
}

// The reference's own render: PathIntegrator::Render with `threads` OpenMP threads (0 = default).
// rgba_out receives FrameBuffer's float buffer; *seconds the reference's own timeConsume.
// n_passes Render() calls on a cleared FrameBuffer, as the UI's loop makes them (ui/RenderThread.cpp:169-175): rgba_out /
// u8_out receive the float and the 8-bit buffer (running mean + tonemap of ui/FrameBuffer.h:127-149), *seconds the
// timeConsume of the last pass.
int gnxh_render_reference_passes(void *h, int maxDepth, int threads, int n_passes, float *rgba_out, unsigned char *u8_out, double *seconds) {
    auto *hs = (HarnessScene *)h;
    if (!hs->scene) return -1;
    if (threads > 0) omp_set_num_threads(threads);
    hs->fb->InitBuffer(hs->width, hs->height, 4);
    hs->fb->renderCountClear();
    std::unique_ptr<SamplerIntegrator> integ(MakeReferenceIntegrator(hs, maxDepth));
    double t = 0;
    for (int k = 0; k < n_passes; ++k) integ->Render(*hs->scene, t);
    if (seconds) *seconds = t;
    const size_t n = (size_t)4 * hs->width * hs->height;
    if (rgba_out) memcpy(rgba_out, hs->fb->getFbuffer(), sizeof(float) * n);
    if (u8_out) memcpy(u8_out, hs->fb->getUCbuffer(), n);
    return 0;
}
int gnxh_render_reference(void *h, int maxDepth, int threads, float *rgba_out, double *seconds) {
    return gnxh_render_reference_passes(h, maxDepth, threads, 1, rgba_out, nullptr, seconds);
}

// Li of individual camera samples (pixel, sample number) straight from PathIntegrator::Li, and the
// index (in the scene's original primitive order) of each sample's primary hit (-1 = miss).
int gnxh_reference_samples(void *h, int maxDepth, int n, const int *px, const int *py, const int *sample, float *rgb_out,
                           int *prim_out) {
    auto *hs = (HarnessScene *)h;
    if (!hs->scene) return -1;
    Bounds2i bounds(Point2i(0, 0), Point2i(hs->width, hs->height));
    std::shared_ptr<Sampler> smp;
    std::unique_ptr<SamplerIntegrator> integp(MakeReferenceIntegrator(hs, maxDepth, &smp));
    SamplerIntegrator &integ = *integp;
    integ.Preprocess(*hs->scene, *smp);
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < n; ++i) {
        MemoryArena arena;
        std::unique_ptr<Sampler> s = smp->Clone(hs->width * py[i] + px[i]);
        Point2i pixel(px[i], py[i]);
        s->StartPixel(pixel);
        s->SetSampleNumber(sample[i]);
        CameraSample cs = s->GetCameraSample(pixel);
        RayDifferential ray;
        hs->camera->GenerateRayDifferential(cs, &ray);
        ray.ScaleDifferentials(1 / std::sqrt((Float)s->samplesPerPixel));
        if (prim_out) {
            SurfaceInteraction isect;
            Ray r(ray);
            bool hit = hs->scene->Intersect(r, &isect);
            int idx = -1;
            if (hit) {
                auto it = hs->originalIndex.find(isect.primitive);
                idx = it == hs->originalIndex.end() ? -2 : it->second;
            }
            prim_out[i] = idx;
        }
        if (rgb_out) {
            Spectrum L = integ.Li(ray, *hs->scene, *s, arena, 0);
            rgb_out[3 * i] = L[0]; rgb_out[3 * i + 1] = L[1]; rgb_out[3 * i + 2] = L[2];
        }
    }
    return 0;
}

// The reference's GaussianFilter::Evaluate (filters/GaussianFilter.cpp:9-12) at n points.
int gnxh_reference_gaussian_eval(float radius, float alpha, int n, const float *x, const float *y, float *out) {
    GaussianFilter filt(Vector2f(radius, radius), alpha);
    for (int i = 0; i < n; ++i) out[i] = filt.Evaluate(Point2f(x[i], y[i]));
    return 0;
}

// Gaussian-filtered image.  The reference has the filter class but no Film that uses it (its Render() box-averages,
// core/Integrator.cpp:274-293), so this is the splat of the renderer it descends from (pbrt-v3 Film::AddSample /
// WriteImage) around the reference's OWN pieces: Sampler::GetCameraSample for pFilm, the integrator's Li for the
// radiance, GaussianFilter::Evaluate for the weight (evaluated exactly, no 16 x 16 table).  sums_out (optional)
// receives (sum L f, sum f) per pixel.
int gnxh_reference_gaussian_film(void *h, int maxDepth, float radius, float alpha, float *rgba_out, float *sums_out) {
    auto *hs = (HarnessScene *)h;
    if (!hs->scene) return -1;
    const int W = hs->width, H = hs->height, spp = hs->spp;
    std::shared_ptr<Sampler> smp;
    std::unique_ptr<SamplerIntegrator> integp(MakeReferenceIntegrator(hs, maxDepth, &smp));
    SamplerIntegrator &integ = *integp;
    integ.Preprocess(*hs->scene, *smp);
    std::vector<float> L((size_t)W * H * spp * 3), pf((size_t)W * H * spp * 2);
#pragma omp parallel for schedule(dynamic, 16)
    for (int pix = 0; pix < W * H; ++pix) {
        MemoryArena arena;
        Point2i pixel(pix % W, pix / W);
        std::unique_ptr<Sampler> s = smp->Clone(pix);
        s->StartPixel(pixel);
        for (int k = 0; k < spp; ++k) {
            s->SetSampleNumber(k);
            CameraSample cs = s->GetCameraSample(pixel);
            RayDifferential ray;
            hs->camera->GenerateRayDifferential(cs, &ray);
            ray.ScaleDifferentials(1 / std::sqrt((Float)s->samplesPerPixel));
            Spectrum Li = integ.Li(ray, *hs->scene, *s, arena, 0);
            size_t i = (size_t)pix * spp + k;
            L[3 * i] = Li[0]; L[3 * i + 1] = Li[1]; L[3 * i + 2] = Li[2];
            pf[2 * i] = cs.pFilm.x; pf[2 * i + 1] = cs.pFilm.y;
            arena.Reset();
        }
    }
    GaussianFilter filt(Vector2f(radius, radius), alpha);
    std::vector<float> sums((size_t)W * H * 4, 0.f);
    for (size_t i = 0; i < (size_t)W * H * spp; ++i) {  // Film::AddSample
        Point2f pFilmDiscrete(pf[2 * i] - 0.5f, pf[2 * i + 1] - 0.5f);
        int x0 = std::max((int)std::ceil(pFilmDiscrete.x - radius), 0), x1 = std::min((int)std::floor(pFilmDiscrete.x + radius) + 1, W);
        int y0 = std::max((int)std::ceil(pFilmDiscrete.y - radius), 0), y1 = std::min((int)std::floor(pFilmDiscrete.y + radius) + 1, H);
        for (int y = y0; y < y1; ++y)
            for (int x = x0; x < x1; ++x) {
                Float w = filt.Evaluate(Point2f(x - pFilmDiscrete.x, y - pFilmDiscrete.y));
                float *px = &sums[4 * ((size_t)y * W + x)];
                px[0] += L[3 * i] * w; px[1] += L[3 * i + 1] * w; px[2] += L[3 * i + 2] * w; px[3] += w;
            }
    }
    for (size_t q = 0; q < (size_t)W * H; ++q) {  // Film::WriteImage
        const float *px = &sums[4 * q];
        float inv = px[3] != 0 ? 1 / px[3] : 0.f;
        for (int c = 0; c < 3; ++c) rgba_out[4 * q + c] = std::max(0.f, px[c] * inv);
        rgba_out[4 * q + 3] = 1.f;
    }
    if (sums_out) memcpy(sums_out, sums.data(), sums.size() * sizeof(float));
    return 0;
}

int gnxh_reference_sample_dims(void *h, int n, const int64_t *index, const int *dim, float *out) {
    auto *hs = (HarnessScene *)h;
    auto *gs = dynamic_cast<GlobalSampler *>(hs->sampler.get());
    if (!gs) return -1;
    std::unique_ptr<Sampler> own = gs->Clone(0);  // (SampleDimension of the Sobol' sampler looks at the current pixel: (0, 0) here)
    auto *g = dynamic_cast<GlobalSampler *>(own.get());
    g->StartPixel(Point2i(0, 0));
    for (int i = 0; i < n; ++i) out[i] = g->SampleDimension(index[i], dim[i]);
    return 0;
}

// HaltonSampler::GetIndexForSample(sample) for a pixel.
int64_t gnxh_reference_sample_index(void *h, int px, int py, int sample) {
    auto *hs = (HarnessScene *)h;
    std::unique_ptr<Sampler> s = hs->sampler->Clone(0);
    s->StartPixel(Point2i(px, py));
    auto *g = dynamic_cast<GlobalSampler *>(s.get());
    return g ? g->GetIndexForSample(sample) : -1;
}

int gnxh_max_threads(void) { return omp_get_max_threads(); }

}  // extern "C"
