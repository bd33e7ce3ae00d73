// ref_harness.h — TEST INFRASTRUCTURE ONLY (oracle side): the scene record shared by the two harness libraries.
//   oracle/_ref/libgnxref.so     the UNMODIFIED reference + ref_harness.cpp: builds the scenes through the reference's class
//                                API and renders them with the reference's own integrators.  Links nothing of the product.
//   oracle/_ref/libgnxbridge.so  the product's drop-in class (gnxraytracer_b200/bridge) + bridge_harness.cpp: instantiates
//                                gnx::CUDAPathIntegrator on the very same pbr::Scene.  Links libgnxref.so and libgnxrt.so.
#pragma once
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "core/Integrator.h"
#include "core/Scene.h"
#include "core/Transform.h"

struct HarnessScene {
    int width = 0, height = 0, spp = 0;
    std::string name;
    std::vector<std::unique_ptr<pbr::Transform>> transforms;  // Triangle keeps raw pointers to these
    std::vector<std::shared_ptr<pbr::Primitive>> prims;       // original (pre-BVH) order
    std::vector<std::shared_ptr<pbr::Light>> lights;
    std::vector<std::shared_ptr<pbr::Medium>> media;
    std::unique_ptr<pbr::Transform> cam2world;
    std::unique_ptr<pbr::AnimatedTransform> animated;
    std::shared_ptr<const pbr::Camera> camera;
    std::shared_ptr<pbr::Sampler> sampler;
    int integrator = 0;  // gnx_integrator: 0 Path, 1 VolPath, 2 Whitted, 3 DirectLighting (UniformSampleOne)
    std::unique_ptr<pbr::Scene> scene;
    std::unique_ptr<FrameBuffer> fb;
    float filterRadius = 0, filterAlpha = 0;  // CUDAPathIntegrator::SetGaussianFilter
    // state of the product side (libgnxbridge.so), opaque here so that this library links nothing of the product
    void *ext = nullptr;
    void (*ext_free)(void *) = nullptr;
    ~HarnessScene() { if (ext && ext_free) ext_free(ext); }
    void DropExt() { if (ext && ext_free) ext_free(ext); ext = nullptr; }
    std::string strategy = "spatial";  // lightSampleStrategy handed to both integrators
    std::unordered_map<const pbr::Primitive *, int> originalIndex; // position in `prims` (the caller's order)
    std::string error;
    double bvhSeconds = 0;

    const pbr::Transform *keep(const pbr::Transform &t) {
        transforms.emplace_back(new pbr::Transform(t));
        return transforms.back().get();
    }
};

