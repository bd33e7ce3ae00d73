// CUDAPathIntegrator — the reference-side half of the drop-in.
//
// A pbr::Integrator (core/Integrator.h:17-23) with the constructor signature of
// PathIntegrator (integrators/PathIntegrator.h:17-20).  Swapping
//     std::make_shared<PathIntegrator>(5, camera, sampler, bounds, fb)
// for
//     std::make_shared<gnx::CUDAPathIntegrator>(5, camera, sampler, bounds, fb)
// in ui/RenderThread.cpp:163-164 is the whole integration.  Render() flattens
// the existing pbr::Scene into the plain buffers of include/gnxrt.h, hands them to libgnxrt.so
// and receives the FrameBuffer's two images — the running mean over Render() calls and its 8-bit
// tonemapped copy, both computed on the device with update_f_u_c's formulas (ui/FrameBuffer.h:127-149) —
// straight into FrameBuffer::fbuffer / ubuffer (gnx_render_framebuffer).  With `devices` (or the
// environment variable GNX_DEVICES=all | 0,1,..) one Render() call uses several GPUs of the box.
//
// This translation unit is the only one compiled against the reference's headers; it contains no
// rendering code.
#ifndef GNX_CUDA_PATH_INTEGRATOR_H
#define GNX_CUDA_PATH_INTEGRATOR_H

#include <memory>
#include <string>
#include <vector>

#include "core/Integrator.h"
#include "gnxrt.h"

namespace gnx {

// Owns the host copies that a gnx_scene_desc points into.
struct FlatScene {
    gnx_scene_desc desc;
    std::vector<gnx_bvh_node> nodes;
    std::vector<float> prim_p, prim_uv, prim_n;
    std::vector<uint8_t> prim_has_n, prim_is_transition, prim_flags;
    std::vector<int32_t> prim_material, prim_light, prim_medium_in, prim_medium_out, prim_id;
    std::vector<gnx_material> materials;
    std::vector<gnx_texture> textures;
    std::vector<std::vector<float>> texture_texels;
    std::vector<gnx_light> lights;
    std::vector<float> light_power;
    std::vector<int32_t> light_n_samples;
    std::vector<float> env_texels, env_cond_func, env_cond_cdf, env_cond_int, env_marg_func, env_marg_cdf;
    std::vector<gnx_medium> media;
    std::vector<std::vector<float>> media_density;
    std::vector<uint16_t> perms;
    std::vector<const void *> prim_ptr;  // ordered pbr::Primitive* (parity hook: pointer -> ordered index)
    std::string error;                   // non-empty when the scene uses something outside the hot path
};

// Reads the private state of Scene/BVHAccel/Triangle/Material/Light/Camera/Sampler
// into `out`.  Returns false (and sets out->error) for unsupported content.
bool FlattenScene(const pbr::Scene &scene, const pbr::Camera &camera, const pbr::Sampler &sampler,
                  FlatScene *out);

class CUDAPathIntegrator : public pbr::Integrator {
  public:
    CUDAPathIntegrator(int maxDepth, std::shared_ptr<const pbr::Camera> camera,
                       std::shared_ptr<pbr::Sampler> sampler, const pbr::Bounds2i &pixelBounds,
                       FrameBuffer *pFrameBuffer, pbr::Float rrThreshold = 1,
                       const std::string &lightSampleStrategy = "spatial", bool volumetric = false,
                       int device = 0, const std::vector<int> &devices = {});
    ~CUDAPathIntegrator() override;

    // Integrator interface.  Synchronous like the reference: the FrameBuffer is complete on return.
    void Render(const pbr::Scene &scene, double &timeConsume) override;

    // Extras for tests / measurement (not part of pbr::Integrator).
    bool ok() const { return ctx_ != nullptr && error_.empty(); }
    const std::string &error() const { return error_; }
    const gnx_stats &lastStats() const { return stats_; }
    bool PrimaryHits(const pbr::Scene &scene, int sample, std::vector<int32_t> *orderedPrimIndex);
    const FlatScene *flat() const { return flat_.get(); }
    gnx_render_params MakeParams() const;
    // GNX_INTEGRATOR_WHITTED / GNX_INTEGRATOR_DIRECT / GNX_INTEGRATOR_DIRECT_ALL (LightStrategy::UniformSampleAll): stand in for pbr::WhittedIntegrator /
    // pbr::DirectLightingIntegrator(LightStrategy::UniformSampleOne) instead of PathIntegrator / VolPathIntegrator
    void SetIntegrator(int gnxIntegrator) { integrator_ = gnxIntegrator; }
    // Reconstruct the image with the reference's GaussianFilter(Vector2f(radius, radius), alpha)
    // (filters/GaussianFilter.h:12-33) instead of Render()'s box average; radius <= 0 switches back to the box.
    void SetGaussianFilter(pbr::Float radius, pbr::Float alpha) { filterRadius_ = radius; filterAlpha_ = alpha; }
    // gnx_partition of a multi-GPU Render(): GNX_PARTITION_SAMPLES (default) or GNX_PARTITION_TILES
    void SetPartition(int gnxPartition) { partition_ = gnxPartition; }
    // The reference's UI calls Render() in a loop and averages the passes, every pass drawing the SAME samples
    // (ui/RenderThread.cpp:169-175, core/Integrator.cpp:262-291).  Progressive: pass k renders samples
    // [k spp, (k+1) spp) of every pixel, so the running mean converges instead of repeating itself.  Off by default.
    void SetProgressive(bool on) { progressive_ = on; }
    // The flattened copy is keyed on the Scene's address and a fingerprint of what hangs off it (aggregate, lights,
    // sampler, camera); call this after editing materials / transforms in place to force a new upload.
    void Invalidate() { uploaded_ = nullptr; }
    int NumDevices() const { return ctx_ ? gnx_num_devices(ctx_) : 0; }

  private:
    bool EnsureUploaded(const pbr::Scene &scene);

    const int maxDepth_;
    std::shared_ptr<const pbr::Camera> camera_;
    std::shared_ptr<pbr::Sampler> sampler_;
    const pbr::Bounds2i pixelBounds_;
    FrameBuffer *fb_;
    const pbr::Float rrThreshold_;
    const std::string lightSampleStrategy_;
    int integrator_;  // gnx_integrator
    pbr::Float filterRadius_ = 0, filterAlpha_ = 0;
    gnx_ctx *ctx_ = nullptr;
    int partition_ = 0;
    bool progressive_ = false;
    const pbr::Scene *uploaded_ = nullptr;
    size_t fingerprint_ = 0;
    std::unique_ptr<FlatScene> flat_;
    gnx_stats stats_{};
    std::string error_;
};

}  // namespace gnx
#endif
