// bridge_harness.cpp — TEST INFRASTRUCTURE ONLY: the product side of the parity harness.
//
// Instantiates the product's drop-in class gnx::CUDAPathIntegrator (gnxraytracer_b200/bridge) on the pbr::Scene objects
// that oracle/ref_harness.cpp builds through the reference's class API, which is how the drop-in boundary is tested:
// same Scene, same Camera, same Sampler, same FrameBuffer class, the reference's integrator on one side and the CUDA one
// on the other.  Linked with the bridge into oracle/_ref/libgnxbridge.so (needs libgnxref.so and libgnxrt.so); the
// reference-only library libgnxref.so — what bench.py's reference arm loads — contains none of this.
#include <chrono>
#include <cstring>
#include <memory>
#include <unordered_map>
#include <vector>

#include "ref_harness.h"
#include "gnxraytracer_b200/bridge/CUDAPathIntegrator.h"

using namespace pbr;

namespace {

struct BridgeState {
    std::unique_ptr<gnx::CUDAPathIntegrator> cuda;
    int maxDepth = -1;
    std::vector<int> devices;  // empty: one GPU (device 0)
    int partition = 0;
    bool progressive = false;
    std::unique_ptr<gnx::FlatScene> flatOnly;                  // gnxh_flatten without a GPU
    std::unordered_map<const Primitive *, int> orderedIndex;  // BVH-ordered index of each primitive
};

BridgeState *StateOf(HarnessScene *hs) {
    if (!hs->ext) {
        hs->ext = new BridgeState;
        hs->ext_free = [](void *p) { delete (BridgeState *)p; };
    }
    return (BridgeState *)hs->ext;
}

void IndexPrims(BridgeState *bs, const gnx::FlatScene *flat) {
    if (!flat) return;
    bs->orderedIndex.clear();
    for (size_t k = 0; k < flat->prim_ptr.size(); ++k) bs->orderedIndex[(const Primitive *)flat->prim_ptr[k]] = (int)k;
}

gnx::CUDAPathIntegrator *EnsureCuda(HarnessScene *hs, int maxDepth) {
    BridgeState *bs = StateOf(hs);
    if (!bs->cuda || bs->maxDepth != maxDepth) {
        Bounds2i bounds(Point2i(0, 0), Point2i(hs->width, hs->height));
        bs->cuda.reset(new gnx::CUDAPathIntegrator(maxDepth, hs->camera, hs->sampler, bounds, hs->fb.get(), 1.f, hs->strategy,
                                                   hs->integrator == 1, bs->devices.empty() ? 0 : bs->devices[0], bs->devices));
        bs->cuda->SetIntegrator(hs->integrator);
        bs->cuda->SetGaussianFilter(hs->filterRadius, hs->filterAlpha);
        bs->cuda->SetPartition(bs->partition);
        bs->cuda->SetProgressive(bs->progressive);
        bs->maxDepth = maxDepth;
    }
    return bs->cuda.get();
}

}  // namespace

extern "C" {

// Flatten only (no GPU needed): returns the gnx_scene_desc the bridge would upload, or NULL.
const gnx_scene_desc *gnxh_flatten(void *h) {
    auto *hs = (HarnessScene *)h;
    BridgeState *bs = StateOf(hs);
    std::unique_ptr<gnx::FlatScene> flat(new gnx::FlatScene);
    if (!gnx::FlattenScene(*hs->scene, *hs->camera, *hs->sampler, flat.get())) {
        hs->error = flat->error;
        return nullptr;
    }
    IndexPrims(bs, flat.get());
    bs->flatOnly = std::move(flat);
    return &bs->flatOnly->desc;
}

// The GPUs the drop-in class drives (gnx_create_multi): n = 0 back to device 0 alone; partition = gnx_partition.
void gnxh_cuda_set_devices(void *h, int n, const int *ids, int partition) {
    auto *hs = (HarnessScene *)h;
    BridgeState *bs = StateOf(hs);
    bs->devices.assign(ids, ids + (n > 0 ? n : 0));
    bs->partition = partition;
    bs->cuda.reset();
}
// Successive Render() calls continue the sample sequence instead of repeating it (CUDAPathIntegrator::SetProgressive).
void gnxh_cuda_set_progressive(void *h, int on) {
    auto *hs = (HarnessScene *)h;
    StateOf(hs)->progressive = on != 0;
    if (StateOf(hs)->cuda) StateOf(hs)->cuda->SetProgressive(on != 0);
}

// n_passes calls of CUDAPathIntegrator::Render on a cleared FrameBuffer (the UI's loop, ui/RenderThread.cpp:169-175);
// rgba_out / u8_out receive the FrameBuffer's float and 8-bit buffers, *seconds the timeConsume of the last pass and
// *wall_seconds (optional) the wall-clock time of that Render() call as the caller sees it.
int gnxh_render_cuda_passes(void *h, int maxDepth, int n_passes, float *rgba_out, unsigned char *u8_out, double *seconds,
                            double *wall_seconds, gnx_stats *stats) {
    auto *hs = (HarnessScene *)h;
    gnx::CUDAPathIntegrator *c = EnsureCuda(hs, maxDepth);
    hs->fb->InitBuffer(hs->width, hs->height, 4);
    hs->fb->renderCountClear();
    double t = 0, wall = 0;
    for (int k = 0; k < n_passes; ++k) {
        auto t0 = std::chrono::steady_clock::now();
        c->Render(*hs->scene, t);
        wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (!c->error().empty()) { hs->error = c->error(); return -1; }
    }
    IndexPrims(StateOf(hs), c->flat());
    if (seconds) *seconds = t;
    if (wall_seconds) *wall_seconds = wall;
    if (stats) *stats = c->lastStats();
    const size_t n = (size_t)4 * hs->width * hs->height;
    if (rgba_out) memcpy(rgba_out, hs->fb->getFbuffer(), sizeof(float) * n);
    if (u8_out) memcpy(u8_out, hs->fb->getUCbuffer(), n);
    return 0;
}

int gnxh_render_cuda(void *h, int maxDepth, float *rgba_out, double *seconds, gnx_stats *stats) {
    return gnxh_render_cuda_passes(h, maxDepth, 1, rgba_out, nullptr, seconds, nullptr, stats);
}

// Re-times Render() on the already uploaded scene without touching the harness buffers' contents: wall seconds per call.
double gnxh_time_cuda_render(void *h, int maxDepth, int n_calls) {
    auto *hs = (HarnessScene *)h;
    gnx::CUDAPathIntegrator *c = EnsureCuda(hs, maxDepth);
    double t = 0;
    c->Render(*hs->scene, t);  // upload + warm-up
    if (!c->error().empty()) { hs->error = c->error(); return -1; }
    auto t0 = std::chrono::steady_clock::now();
    for (int k = 0; k < n_calls; ++k) c->Render(*hs->scene, t);
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / (n_calls > 0 ? n_calls : 1);
}

int gnxh_cuda_primary_hits(void *h, int sample, int *ordered_out) {
    auto *hs = (HarnessScene *)h;
    BridgeState *bs = StateOf(hs);
    gnx::CUDAPathIntegrator *c = EnsureCuda(hs, bs->maxDepth < 0 ? 5 : bs->maxDepth);
    std::vector<int32_t> v;
    if (!c->PrimaryHits(*hs->scene, sample, &v)) { hs->error = c->error(); return -1; }
    IndexPrims(bs, c->flat());
    memcpy(ordered_out, v.data(), v.size() * sizeof(int));
    return 0;
}

// Maps BVH-ordered primitive indices (gnx_geometry::prim_id of a bridge-flattened scene) to the
// scene's original primitive order; -1 stays -1.  Requires gnxh_flatten or a CUDA call before.
int gnxh_ordered_to_original(void *h, int n, const int *ordered, int *original) {
    auto *hs = (HarnessScene *)h;
    BridgeState *bs = StateOf(hs);
    std::vector<int> map(bs->orderedIndex.size(), -2);
    for (auto &kv : bs->orderedIndex) {
        auto it = hs->originalIndex.find(kv.first);
        if (kv.second >= 0 && kv.second < (int)map.size() && it != hs->originalIndex.end()) map[kv.second] = it->second;
    }
    for (int i = 0; i < n; ++i) original[i] = ordered[i] < 0 ? -1 : (ordered[i] < (int)map.size() ? map[ordered[i]] : -2);
    return 0;
}

}  // extern "C"
