// gnx_render.cu — the C ABI of include/gnxrt.h: context, scene upload (host SoA -> HBM layout),
// the wavefront render loop and the parity hooks.  All compute is in gnx_kernels.cuh; there is no
// CPU code path for any of it.
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#define GNX_TU_MAIN
#include "gnx_kernels.cuh"
#include "gnx_lbvh.cuh"
#include "gnx_pack.h"

using namespace gnx;

// Every translation unit has its own copy of the __constant__ Sobol' matrices (gnx_sampler.cuh) and registers an
// uploader for it here; gnx_upload_scene calls them all.
static std::vector<void (*)(const uint32_t *)> &sobol_uploaders() {
    static std::vector<void (*)(const uint32_t *)> v;
    return v;
}
void gnx_register_sobol_uploader(void (*fn)(const uint32_t *)) { sobol_uploaders().push_back(fn); }

namespace {
thread_local std::string g_create_error;

// ---- NCCL, loaded at run time: the single-GPU product has no dependency on it, an N-device job uses it for the one
// collective the path has (the framebuffer sum-reduce) and falls back to peer-to-peer loads without it.
// The few types are restated (nccl.h: ncclUniqueId is 128 opaque bytes, ncclFloat32 = 7, ncclSum = 0, ncclSuccess = 0).
typedef void *NcclComm;
struct NcclId { char internal[GNX_COMM_ID_BYTES]; };
struct NcclApi {
    void *h = nullptr;
    int (*GetUniqueId)(NcclId *) = nullptr;
    int (*CommInitRank)(NcclComm *, int, NcclId, int) = nullptr;
    int (*CommInitAll)(NcclComm *, int, const int *) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*Reduce)(const void *, void *, size_t, int, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
constexpr int kNcclFloat32 = 7, kNcclSum = 0;
NcclApi *nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {getenv("GNX_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *n : names)
            if (n && n[0] && (api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
        if (!api.h) return;
        bool ok = true;
        auto sym = [&](const char *name) { void *p = dlsym(api.h, name); ok = ok && p; return p; };
        api.GetUniqueId = (int (*)(NcclId *))sym("ncclGetUniqueId");
        api.CommInitRank = (int (*)(NcclComm *, int, NcclId, int))sym("ncclCommInitRank");
        api.CommInitAll = (int (*)(NcclComm *, int, const int *))sym("ncclCommInitAll");
        api.CommDestroy = (int (*)(NcclComm))sym("ncclCommDestroy");
        api.Reduce = (int (*)(const void *, void *, size_t, int, int, int, NcclComm, cudaStream_t))sym("ncclReduce");
        api.GroupStart = (int (*)())sym("ncclGroupStart");
        api.GroupEnd = (int (*)())sym("ncclGroupEnd");
        api.GetErrorString = (const char *(*)(int))sym("ncclGetErrorString");
        if (!ok) { dlclose(api.h); api.h = nullptr; }
    });
    return api.h ? &api : nullptr;
}
}  // namespace

// Share of an N-device job that one device renders (gnx_render_params::partition).
struct Share {
    int n = 1, idx = 0, partition = GNX_PARTITION_SAMPLES;
};

struct gnx_ctx {
    int device = 0;
    std::string err;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int sm_count = 148;
    void *geom_base = nullptr;
    size_t geom_bytes = 0;
    // GNX_L2_ENV=1: the environment light's tables (texels, Distribution2D, guides: one slab) get a persisting L2 window —
    // every shaded vertex samples them at a random row, while the path state streams through L2 in between
    int l2_env = 0;
    void *env_base = nullptr;
    long long l2_persist_max = -1, l2_window_max = 0;
    size_t l2_set_aside = 0;
    size_t env_bytes = 0;
    int l2_persist = 0;  // measured on B200/C2: 52.1 ms with the window vs 46.1 ms without (set-aside starves the rest)
    int grid_trace = 148 * 8, grid_shade = 148 * 4, grid_shade8 = 148 * 4, grid_volpath = 148 * 2, grid_recursive = 148 * 2;  // SM count x resident blocks (occupancy query at create)
    // scene
    bool has_scene = false;
    bool has_next_lights = false;  // point / spot / distant / skybox lights present
    bool tex_needs_pyramid = false; // an image texture came as one level that is not a power of two: no MIPMap pyramid to filter with
    cudaStream_t stream_any = nullptr;  // the any-hit launches of a bounce run here, next to the extension launch on `stream`
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool anyhit_overlap = true;    // GNX_ANYHIT_OVERLAP=0: any-hit launch on the render stream, after the extension launch
    int grid_anyhit8 = 148 * 8, grid_trace8 = 148 * 8, grid_trace8p = 148 * 8;
    bool closest8 = false;         // GNX_CLOSEST_BVH8=1: extension / camera rays through the 8-wide tree too (k_trace<., true> + retrace of
                                   // flagged rays).  Bit-equal, but measured SLOWER than the two-child tree (C2 extend 16.9 -> 18.3 ms, U1p
                                   // 81.4 -> 84.8 ms, C3 46.9 -> 52.9 ms): the kernel is bound by instruction issue, and decoding eight
                                   // quantised boxes costs about as many instructions as the three two-child nodes it replaces
    bool anyhit8 = true;           // GNX_ANYHIT_BVH8=0: any-hit rays walk the two-child tree like the closest-hit rays
    bool merge_extend = true;      // GNX_MERGE_EXTEND=0: the any-hit rays of a bounce get their own launch(es)
    int film_chunk = 0;            // GNX_FILM_CHUNK: samples per pixel staged at a time by the tiled Gaussian gather (0 = auto)
    bool film_simple = false;      // GNX_FILM_SIMPLE=1: Gaussian film with the per-pixel gather instead of the tiled one
    bool merge_shadow = true;      // GNX_MERGE_SHADOW=0: shadow A and B rays in two launches
    float bvh_build_ms = 0;        // device time of the last device-side BVH build (0: the caller supplied the nodes)
    DeviceScene sc{};
    std::vector<void *> scene_allocs;
    unsigned shade_type_mask = 0;  // which k_shade variants the scene needs
    bool spatial_built = false;
    // the two scene-wide light tables (uniform / power); choose_light reads whichever sc.ld.uni_* points at
    const float *one_func = nullptr, *one_cdf = nullptr, *pow_func = nullptr, *pow_cdf = nullptr;
    float one_int = 0, pow_int = 0;
    std::string pow_error;         // why there is no power table
    int n_lights_host = 0, n_textures_host = 0;
    float wb[6]{};
    // wavefront buffers
    int capacity = 0;
    PathState ps{};
    Queues q{};
    VolWave vw{};                  // VolPath wavefront state (allocated on the first VolPath render)
    int vw_capacity = 0;
    std::vector<void *> vw_allocs;
    // shade queues put back into slot order before shading (k_qs_*): bitmap over the path slots + per-block counts
    unsigned *qs_bits = nullptr;
    int *qs_blocks = nullptr;
    size_t qs_words = 0;
    std::vector<void *> qs_allocs;
    bool rebin = false;            // GNX_REBIN=1 (experiment): extension queue sorted by origin Morton code + direction octant
    uint32_t *rb_keys[2] = {nullptr, nullptr};
    int *rb_vals[2] = {nullptr, nullptr};
    void *rb_tmp = nullptr;
    size_t rb_tmp_bytes = 0, rb_slots = 0;
    std::vector<void *> rb_allocs;
    bool sort_queues = true;       // GNX_SORT_QUEUES=0: shade in the completion order of the traversal kernel
    // Sorting pays on dense queues only (kQsMinDensity).  The kernels check that themselves, but in a scene where few camera
    // rays hit anything even their empty launches cost ~1 %: the previous render call's hit count of the first batch (copied
    // to pinned memory without a synchronisation) lets the host skip the launches altogether.  Results do not depend on it.
    int *qs_probe_dev = nullptr;
    volatile int *qs_probe_host = nullptr;  // [0] hits, [1] slots of that batch; slots 0 = nothing known yet
    // WhittedIntegrator with its first vertex staged (k_whitted_vertex): one shadow item and one contribution plane per light and path
    ShadowItem *ww_items = nullptr;
    float4 *ww_planes = nullptr;
    size_t ww_slots = 0;           // capacity x lights the two buffers hold
    std::vector<void *> ww_allocs;
    bool whitted_staged = true;    // GNX_WHITTED_STAGED=0: every sample through the per-lane recursion (k_recursive<0>)
    int grid_whitted_vertex = 148 * 4;
    bool vol_megakernel = false;   // GNX_VOLPATH_MEGAKERNEL=1: the per-lane kernel k_volpath instead of the staged wavefront
    int grid_vp_logic = 148 * 4, grid_vp_track = 148 * 8;
    int grid_vp_kernel[4] = {148 * 4, 148 * 4, 148 * 4, 148 * 4};  // per logic kernel (VolKernel): each is sized from its own occupancy
    std::vector<void *> wave_allocs;
    float4 *accum = nullptr, *rgba = nullptr;
    int film_pixels = 0;
    DevStats *d_stats = nullptr;
    bool stage_timers = false;
    // per-stage CUDA-event pairs (recorded on the render stream, read back after the render)
    std::vector<cudaEvent_t> ev_pool;
    std::vector<int> ev_stage;  // stage id of pair i (events 2i, 2i+1)
    size_t ev_used = 0;
    // ---- N-device jobs.  gnx_create_multi: this context is the root (device_ids[0]) and owns one full single-device
    // context per further GPU; gnx_comm_attach: this context is rank `rank` of an n_ranks-process job.
    std::vector<gnx_ctx *> peers;
    std::vector<void *> comms;     // single process: NCCL communicator of the root, then of every peer; attached: [0] = this rank's
    int n_ranks = 1, rank = 0;
    int reduce_mode = 0;           // GNX_REDUCE: 0 = NCCL when it can be loaded, else peer-to-peer; 1 = nccl; 2 = p2p
    bool p2p_direct = false;       // the root can load from every peer's memory (cudaDeviceEnablePeerAccess)
    cudaEvent_t ev_done = nullptr; // this device's share is queued up to here
    float4 *stage = nullptr;       // root, without direct peer access: copies of the peers' frames
    size_t stage_pixels = 0;
    double ms_reduce = 0;          // device time of the last reduce (root)
    // ---- gnx_render_framebuffer: the FrameBuffer's running mean on the device, the 8-bit image, pinned staging
    float4 *fb_state = nullptr;
    uchar4 *fb_u8 = nullptr;
    size_t fb_pixels = 0;          // allocation
    int fb_w = 0, fb_h = 0;        // image the state belongs to (0: no state)
    void *fb_pinned = nullptr;     // pinned host staging: float4[pixels] then uchar4[pixels]
    size_t fb_pinned_pixels = 0;
    cudaEvent_t fb_ev[8] = {};
};

enum { ST_RAYGEN = 0, ST_EXTEND, ST_SHADE, ST_SHADOW, ST_FILM, ST_VP_EXTEND, ST_VP_VERTEX, ST_VP_SHADOW, ST_VP_MIS, ST_VP_TRACK, ST_COUNT };

// RAII-free helper: brackets the launches issued between begin() and end() with two events.
struct StageTimer {
    gnx_ctx *ctx; cudaStream_t st; bool on;
    void begin(int stage) {
        if (!on) return;
        if (ctx->ev_used + 2 > ctx->ev_pool.size()) {
            for (int i = 0; i < 2; ++i) { cudaEvent_t e; cudaEventCreate(&e); ctx->ev_pool.push_back(e); }
        }
        ctx->ev_stage.push_back(stage);
        cudaEventRecord(ctx->ev_pool[ctx->ev_used], st);
    }
    void end() {
        if (!on) return;
        cudaEventRecord(ctx->ev_pool[ctx->ev_used + 1], st);
        ctx->ev_used += 2;
    }
};

#define GNX_CUDA(ctx, call)                                                                         \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                        \
            return GNX_ERR_CUDA;                                                                    \
        }                                                                                           \
    } while (0)

static int fail(gnx_ctx *ctx, int code, const std::string &msg) {
    ctx->err = msg;
    return code;
}

// ---- experiment (GNX_REBIN=1, off by default): the extension queue re-binned by the Morton code of the ray origins and the
// direction octant between bounces ("measure ray re-binning by hit-point Morton code").  The sort is
// cub's radix sort over the whole batch (queue lengths are not known on the host); see DESIGN.md section 11 for the numbers.
__device__ __forceinline__ uint32_t rebin_spread3(uint32_t v) {  // 9 bits -> every third bit
    v &= 0x1ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__global__ void k_rebin_keys(const DeviceScene sc, PathState ps, const int *list, const int *count, uint32_t *keys, int *vals, int nPad) {
    const int n = *count;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nPad; i += gridDim.x * blockDim.x) {
        uint32_t key = 0xffffffffu;
        int slot = 0;
        if (i < n) {
            slot = list[i];
            const float4 o = ps.ray_o[slot], d = ps.ray_d[slot];
            const float p[3] = {o.x, o.y, o.z};
            uint32_t q[3];
            for (int a = 0; a < 3; ++a) {
                const float ext = sc.wb_max[a] - sc.wb_min[a];
                float t = ext > 0.f ? (p[a] - sc.wb_min[a]) / ext : 0.f;
                t = fminf(fmaxf(t, 0.f), 1.f);
                q[a] = (uint32_t)(t * 511.f);
            }
            const uint32_t morton = rebin_spread3(q[0]) | (rebin_spread3(q[1]) << 1) | (rebin_spread3(q[2]) << 2);
            const uint32_t oct = (d.x < 0.f ? 1u : 0u) | (d.y < 0.f ? 2u : 0u) | (d.z < 0.f ? 4u : 0u);
            key = (morton << 3) | oct;
        }
        keys[i] = key;
        vals[i] = slot;
    }
}

template <typename T>
static int dupload(gnx_ctx *ctx, std::vector<void *> &pool, const T *host, size_t n, T **out) {
    *out = nullptr;
    if (n == 0) return GNX_OK;
    void *p = nullptr;
    GNX_CUDA(ctx, cudaMalloc(&p, n * sizeof(T)));
    pool.push_back(p);
    if (host) GNX_CUDA(ctx, cudaMemcpy(p, host, n * sizeof(T), cudaMemcpyHostToDevice));
    else GNX_CUDA(ctx, cudaMemset(p, 0, n * sizeof(T)));
    *out = (T *)p;
    return GNX_OK;
}

static void free_pool(std::vector<void *> &pool) {
    for (void *p : pool) cudaFree(p);
    pool.clear();
}

extern "C" {

int gnx_abi_version(void) { return GNX_ABI_VERSION; }

int gnx_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char *gnx_last_error(const gnx_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int gnx_create(gnx_ctx **out, int device) {
    if (!out) { g_create_error = "gnx_create: out is NULL"; return GNX_ERR_INVALID; }
    *out = nullptr;
    int n = gnx_device_count();
    if (n <= 0) {
        g_create_error = "no usable CUDA device (libgnxrt has no CPU fallback)";
        return GNX_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n) { g_create_error = "device index out of range"; return GNX_ERR_INVALID; }
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return GNX_ERR_CUDA; }
    gnx_ctx *ctx = new gnx_ctx;
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    if ((e = cudaStreamCreate(&ctx->stream)) != cudaSuccess ||
        (e = cudaEventCreate(&ctx->ev0)) != cudaSuccess || (e = cudaEventCreate(&ctx->ev1)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&ctx->ev_done, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&ctx->stream_any, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaMalloc((void **)&ctx->d_stats, sizeof(DevStats))) != cudaSuccess) {
        g_create_error = cudaGetErrorString(e);
        delete ctx;
        return GNX_ERR_CUDA;
    }
    {
        int b = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_trace<0>, kBlock, 0) == cudaSuccess && b > 0) ctx->grid_trace = ctx->sm_count * b;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_anyhit8<0>, kBlock, 0) == cudaSuccess && b > 0) ctx->grid_anyhit8 = ctx->sm_count * b;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_trace<0, true>, kBlock, 0) == cudaSuccess && b > 0) ctx->grid_trace8 = ctx->sm_count * b;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_trace<3, true>, kBlock, 0) == cudaSuccess && b > 0) ctx->grid_trace8p = ctx->sm_count * b;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_shade<2>, kShadeBlock, 0) == cudaSuccess && b > 0) ctx->grid_shade = ctx->sm_count * b;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_shade<8>, kShadeBlock, 0) == cudaSuccess && b > 0) ctx->grid_shade8 = ctx->sm_count * b;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_volpath<false>, kBlock, 0) == cudaSuccess && b > 0) ctx->grid_volpath = ctx->sm_count * b;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_recursive<0, false>, kBlock, 0) == cudaSuccess && b > 0) ctx->grid_recursive = ctx->sm_count * b;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_whitted_vertex<false>, kBlock, 0) == cudaSuccess && b > 0) ctx->grid_whitted_vertex = ctx->sm_count * b;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_vp_logic<VK_MIS, false>, kBlock, 0) == cudaSuccess && b > 0) ctx->grid_vp_logic = ctx->sm_count * b;
        for (int k = 0; k < 4; ++k) ctx->grid_vp_kernel[k] = ctx->grid_vp_logic;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_vp_logic<VK_EXTEND, false>, kBlock, 0) == cudaSuccess && b > 0) ctx->grid_vp_kernel[VK_EXTEND] = ctx->sm_count * b;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_vp_logic<VK_VERTEX, false>, kBlock, 0) == cudaSuccess && b > 0) ctx->grid_vp_kernel[VK_VERTEX] = ctx->sm_count * b;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_vp_logic<VK_SHADOW, false>, kBlock, 0) == cudaSuccess && b > 0) ctx->grid_vp_kernel[VK_SHADOW] = ctx->sm_count * b;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_vp_logic<VK_MIS, false>, kBlock, 0) == cudaSuccess && b > 0) ctx->grid_vp_kernel[VK_MIS] = ctx->sm_count * b;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_vp_track, kBlock, 0) == cudaSuccess && b > 0) ctx->grid_vp_track = ctx->sm_count * b;
    }
    // the tiled Gaussian-film gather stages up to 72 KB per block (above the 48 KB a kernel gets without asking)
    cudaFuncSetAttribute(k_accumulate_gauss_tiled<2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
    cudaFuncSetAttribute(k_accumulate_gauss_tiled<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
    cudaFuncSetAttribute(k_accumulate_gauss_tiled<1, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
    cudaFuncSetAttribute(k_accumulate_gauss_tiled<3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
    cudaFuncSetAttribute(k_accumulate_gauss_tiled<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
    cudaGetLastError();
    if (const char *l2 = getenv("GNX_L2_PERSIST")) ctx->l2_persist = l2[0] != '0';
    if (const char *l2 = getenv("GNX_L2_ENV")) ctx->l2_env = l2[0] != '0';
    if (const char *ms = getenv("GNX_MERGE_SHADOW")) ctx->merge_shadow = ms[0] != '0';
    if (const char *fs = getenv("GNX_FILM_SIMPLE")) ctx->film_simple = fs[0] == '1';
    if (const char *fc = getenv("GNX_FILM_CHUNK")) ctx->film_chunk = atoi(fc);
    if (const char *me = getenv("GNX_MERGE_EXTEND")) ctx->merge_extend = me[0] != '0';
    if (const char *a8 = getenv("GNX_ANYHIT_BVH8")) ctx->anyhit8 = a8[0] != '0';
    if (const char *a8 = getenv("GNX_ANYHIT_OVERLAP")) ctx->anyhit_overlap = a8[0] != '0';
    if (const char *c8 = getenv("GNX_CLOSEST_BVH8")) ctx->closest8 = c8[0] == '1';
    if (const char *vm = getenv("GNX_VOLPATH_MEGAKERNEL")) ctx->vol_megakernel = vm[0] == '1';
    if (const char *ws = getenv("GNX_WHITTED_STAGED")) ctx->whitted_staged = ws[0] != '0';
    if (const char *sq = getenv("GNX_SORT_QUEUES")) ctx->sort_queues = sq[0] != '0';
    if (const char *rb = getenv("GNX_REBIN")) ctx->rebin = rb[0] == '1';
    if (const char *rm = getenv("GNX_REDUCE")) ctx->reduce_mode = !strcmp(rm, "nccl") ? 1 : (!strcmp(rm, "p2p") ? 2 : 0);
    const char *t = getenv("GNX_STAGE_TIMERS");
    ctx->stage_timers = !(t && t[0] == '0');
    *out = ctx;
    return GNX_OK;
}

void gnx_destroy(gnx_ctx *ctx) {
    if (!ctx) return;
    if (NcclApi *nc = ctx->comms.empty() ? nullptr : nccl_api())
        for (void *c : ctx->comms) if (c) nc->CommDestroy(c);
    ctx->comms.clear();
    for (gnx_ctx *p : ctx->peers) gnx_destroy(p);
    ctx->peers.clear();
    cudaSetDevice(ctx->device);
    if (ctx->stage) cudaFree(ctx->stage);
    if (ctx->ev_done) cudaEventDestroy(ctx->ev_done);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->stream_any) cudaStreamDestroy(ctx->stream_any);
    if (ctx->fb_state) cudaFree(ctx->fb_state);
    if (ctx->fb_u8) cudaFree(ctx->fb_u8);
    if (ctx->fb_pinned) cudaFreeHost(ctx->fb_pinned);
    if (ctx->qs_probe_host) cudaFreeHost((void *)ctx->qs_probe_host);
    if (ctx->qs_probe_dev) cudaFree(ctx->qs_probe_dev);
    for (cudaEvent_t e : ctx->fb_ev) if (e) cudaEventDestroy(e);
    free_pool(ctx->scene_allocs);
    free_pool(ctx->wave_allocs);
    free_pool(ctx->vw_allocs);
    free_pool(ctx->ww_allocs);
    free_pool(ctx->qs_allocs);
    free_pool(ctx->rb_allocs);
    if (ctx->accum) cudaFree(ctx->accum);
    if (ctx->rgba) cudaFree(ctx->rgba);
    if (ctx->d_stats) cudaFree(ctx->d_stats);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

// A copy of the per-primitive arrays of a scene description in another primitive order (device-built BVH).
struct ReorderedGeometry {
    gnx_scene_desc desc;
    std::vector<float> p, uv, n;
    std::vector<uint8_t> has_n, flags, transition;
    std::vector<int32_t> material, light, id, med_in, med_out;
    std::vector<gnx_light> lights;
    void build(const gnx_scene_desc &src, const std::vector<int> &order) {
        desc = src;
        const gnx_geometry &g = src.geom;
        const int np = g.n_prims;
        std::vector<int> newIndex(np);
        for (int k = 0; k < np; ++k) newIndex[order[k]] = k;
        auto gather = [&](auto &dst, const auto *srcArr, int stride) {
            if (!srcArr) return;
            dst.resize((size_t)np * stride);
            for (int k = 0; k < np; ++k)
                for (int c = 0; c < stride; ++c) dst[(size_t)k * stride + c] = srcArr[(size_t)order[k] * stride + c];
        };
        gather(p, g.prim_p, 9); gather(uv, g.prim_uv, 6); gather(n, g.prim_n, 9); gather(has_n, g.prim_has_n, 1);
        gather(flags, g.prim_flags, 1); gather(transition, g.prim_is_transition, 1); gather(material, g.prim_material, 1);
        gather(light, g.prim_light, 1); gather(med_in, g.prim_medium_in, 1); gather(med_out, g.prim_medium_out, 1);
        id.resize(np);
        for (int k = 0; k < np; ++k) id[k] = g.prim_id ? g.prim_id[order[k]] : order[k];
        gnx_geometry &o = desc.geom;
        o.prim_p = p.data(); o.prim_id = id.data();
        if (g.prim_uv) o.prim_uv = uv.data();
        if (g.prim_n) o.prim_n = n.data();
        if (g.prim_has_n) o.prim_has_n = has_n.data();
        if (g.prim_flags) o.prim_flags = flags.data();
        if (g.prim_is_transition) o.prim_is_transition = transition.data();
        if (g.prim_material) o.prim_material = material.data();
        if (g.prim_light) o.prim_light = light.data();
        if (g.prim_medium_in) o.prim_medium_in = med_in.data();
        if (g.prim_medium_out) o.prim_medium_out = med_out.data();
        lights.assign(src.lights, src.lights + src.n_lights);
        for (gnx_light &l : lights)
            if (l.type == GNX_LIGHT_AREA_TRI && l.prim >= 0 && l.prim < np) l.prim = newIndex[l.prim];
        desc.lights = lights.data();
    }
};

static int upload_one(gnx_ctx *ctx, const gnx_scene_desc *d_in);

// The scene is replicated on every device of a multi-device context: one host thread per GPU.
int gnx_upload_scene(gnx_ctx *ctx, const gnx_scene_desc *d_in) {
    if (!ctx || !d_in) return GNX_ERR_INVALID;
    if (ctx->peers.empty()) return upload_one(ctx, d_in);
    std::vector<int> rcs(ctx->peers.size(), GNX_OK);
    std::vector<std::thread> th;
    for (size_t g = 0; g < ctx->peers.size(); ++g)
        th.emplace_back([&, g] { rcs[g] = upload_one(ctx->peers[g], d_in); });
    int rc = upload_one(ctx, d_in);
    for (std::thread &t : th) t.join();
    for (size_t g = 0; g < rcs.size(); ++g)
        if (rc == GNX_OK && rcs[g] != GNX_OK) { rc = rcs[g]; ctx->err = "device " + std::to_string(ctx->peers[g]->device) + ": " + ctx->peers[g]->err; }
    if (rc != GNX_OK) { ctx->has_scene = false; for (gnx_ctx *p : ctx->peers) p->has_scene = false; }
    return rc;
}

static int upload_one(gnx_ctx *ctx, const gnx_scene_desc *d_in) {
    if (d_in->abi_version != GNX_ABI_VERSION) return fail(ctx, GNX_ERR_INVALID, "abi_version mismatch");
    GNX_CUDA(ctx, cudaSetDevice(ctx->device));
    free_pool(ctx->scene_allocs);
    ctx->env_base = nullptr;
    ctx->env_bytes = 0;
    ctx->has_scene = false;
    ctx->spatial_built = false;
    ctx->has_next_lights = false;
    ctx->tex_needs_pyramid = false;
    ctx->bvh_build_ms = 0;
    DeviceScene sc{};
    std::vector<void *> &pool = ctx->scene_allocs;
    if (d_in->geom.n_prims < 0 || d_in->geom.n_nodes < 0 || (d_in->geom.n_prims > 0 && (!d_in->geom.prim_p || !d_in->geom.prim_material)) ||
        (d_in->geom.n_nodes > 0 && !d_in->geom.nodes))
        return fail(ctx, GNX_ERR_INVALID, "geometry arrays missing");
    // No node array: the BVH is built here, on the device (gnx_lbvh.cuh), and the per-primitive arrays are
    // re-ordered to match it.  prim_id keeps pointing at the caller's order.
    const bool deviceBuilt = d_in->geom.n_prims > 0 && d_in->geom.n_nodes == 0;
    float4 *lbvhNodes = nullptr;
    int lbvhCount = 0;
    ReorderedGeometry reordered;
    const gnx_scene_desc *d = d_in;
    if (deviceBuilt) {
        std::vector<int> order;
        cudaError_t be = lbvh_build(d_in->geom.prim_p, d_in->geom.n_prims, ctx->stream, &lbvhNodes, &lbvhCount, order, &ctx->bvh_build_ms);
        if (be != cudaSuccess) { ctx->err = std::string("device BVH build: ") + cudaGetErrorString(be); return GNX_ERR_CUDA; }
        for (int i = 0; i < d_in->n_lights; ++i)
            if (d_in->lights[i].type == GNX_LIGHT_AREA_TRI && (d_in->lights[i].prim < 0 || d_in->lights[i].prim >= d_in->geom.n_prims)) {
                cudaFree(lbvhNodes);
                return fail(ctx, GNX_ERR_INVALID, "area light primitive out of range");
            }
        reordered.build(*d_in, order);
        d = &reordered.desc;
    }
    const gnx_geometry &g = d->geom;
    {
        std::string verr;
        if (!validate_scene_desc(*d, &verr)) { if (lbvhNodes) cudaFree(lbvhNodes); return fail(ctx, GNX_ERR_INVALID, verr); }
    }
    if (d->n_materials > (1 << 20) - 2) return fail(ctx, GNX_ERR_UNSUPPORTED, "too many materials");
    for (int i = 0; i < d->n_materials; ++i)
        if (d->materials[i].type < 0 || d->materials[i].type > GNX_MAT_DISNEY)
            return fail(ctx, GNX_ERR_INVALID, "unknown material type");

    // ---- nodes: the reference's 32-byte LinearBVHNode array re-packed into 64-byte two-child records
    static_assert(sizeof(gnx_bvh_node) == 32, "node size");
    int rc;
    // ---- BVH nodes (64-byte two-child records re-packed from the reference's LinearBVHNode array) and
    // 48-byte triangle records, in ONE allocation so that a single L2 access-policy window can keep the
    // traversal working set resident while path state and queues stream through (see set_l2_window).
    unsigned typeMask = 0;
    {
        std::vector<float4> n2, tris;
        std::string perr;
        int nNodes = 0;
        float4 *devNodes = nullptr;  // nodes built on the device (no node array supplied)
        if (deviceBuilt) {
            devNodes = lbvhNodes;
            nNodes = lbvhCount;
        } else if (!build_nodes(g.nodes, g.n_nodes, n2, &nNodes, &perr)) return fail(ctx, GNX_ERR_INVALID, perr);
        if (!pack_triangles(*d, tris, &typeMask, &perr)) { if (devNodes) cudaFree(devNodes); return fail(ctx, GNX_ERR_INVALID, perr); }
        const size_t nodeF4 = deviceBuilt ? (size_t)nNodes * 4 : n2.size();
        float4 *dn;
        if ((rc = dupload<float4>(ctx, pool, nullptr, nodeF4 + tris.size(), &dn))) { if (devNodes) cudaFree(devNodes); return rc; }
        cudaError_t ce;
        if (deviceBuilt) {
            ce = cudaMemcpy(dn, devNodes, nodeF4 * sizeof(float4), cudaMemcpyDeviceToDevice);
            cudaFree(devNodes);
        } else ce = cudaMemcpy(dn, n2.data(), nodeF4 * sizeof(float4), cudaMemcpyHostToDevice);
        if (ce == cudaSuccess) ce = cudaMemcpy(dn + nodeF4, tris.data(), tris.size() * sizeof(float4), cudaMemcpyHostToDevice);
        GNX_CUDA(ctx, ce);
        sc.nodes2 = dn;
        sc.n_nodes2 = nNodes;
        sc.tris = dn + nodeF4;
        ctx->geom_base = dn;
        ctx->geom_bytes = (nodeF4 + tris.size()) * sizeof(float4);
#if GNX_BVH_WIDTH == 2
        // ---- the any-hit tree: the two-child records collapsed into compressed 8-wide nodes (gnx_bvh8.cuh).  A tree built
        // on the device is read back for the collapse (64 bytes per interior node).
        if ((ctx->anyhit8 || ctx->closest8) && nNodes > 0) {
            std::vector<uint4> n8;
            if (deviceBuilt) {
                n2.resize(nodeF4);
                GNX_CUDA(ctx, cudaMemcpy(n2.data(), dn, nodeF4 * sizeof(float4), cudaMemcpyDeviceToHost));
            }
            if (build_node8(n2.data(), nNodes, n8) && !n8.empty()) {
                uint4 *d8;
                if ((rc = dupload(ctx, pool, n8.data(), n8.size(), &d8))) return rc;
                sc.nodes8 = d8;
                sc.n_nodes8 = (int)(n8.size() / kNode8Words);
                sc.wide_any = ctx->anyhit8;
                sc.wide_closest = ctx->closest8;
            }
        }
#endif
    }
    sc.n_nodes = g.n_nodes;
    sc.n_prims = g.n_prims;
    ctx->shade_type_mask = typeMask;
    float *df;
    uint8_t *du8;
    if (g.prim_uv) { if ((rc = dupload(ctx, pool, g.prim_uv, (size_t)g.n_prims * 6, &df))) return rc; sc.tri_uv = df; }
    if (g.prim_n && g.prim_has_n) {
        if ((rc = dupload(ctx, pool, g.prim_n, (size_t)g.n_prims * 9, &df))) return rc;
        sc.tri_n = df;
        if ((rc = dupload(ctx, pool, g.prim_has_n, (size_t)g.n_prims, &du8))) return rc;
        sc.tri_has_n = du8;
    }
    if (g.prim_medium_in && g.prim_medium_out) {
        std::vector<int2> med((size_t)g.n_prims);
        for (int k = 0; k < g.n_prims; ++k) med[k] = make_int2(g.prim_medium_in[k], g.prim_medium_out[k]);
        int2 *dm;
        if ((rc = dupload(ctx, pool, med.data(), med.size(), &dm))) return rc;
        sc.tri_media = dm;
        if (g.prim_is_transition) {
            if ((rc = dupload(ctx, pool, g.prim_is_transition, (size_t)g.n_prims, &du8))) return rc;
            sc.tri_transition = du8;
        }
    }
    for (int c = 0; c < 3; ++c) { sc.wb_min[c] = g.world_bound[c]; sc.wb_max[c] = g.world_bound[3 + c]; }
    memcpy(ctx->wb, g.world_bound, sizeof(ctx->wb));

    // ---- materials, textures
    gnx_material *dm;
    if ((rc = dupload(ctx, pool, d->materials, (size_t)d->n_materials, &dm))) return rc;
    sc.materials = dm;
    sc.n_materials = d->n_materials;
    ctx->n_textures_host = d->n_textures;
    {
        std::vector<DevTexture> tex((size_t)d->n_textures);
        for (int i = 0; i < d->n_textures; ++i) {
            const gnx_texture &t = d->textures[i];
            if (t.width <= 0 || t.height <= 0 || (t.n_channels != 1 && t.n_channels != 3) || !t.texels)
                return fail(ctx, GNX_ERR_INVALID, "bad texture descriptor");
            if (t.n_levels > kMaxMipLevels) return fail(ctx, GNX_ERR_INVALID, "texture with more than 16 MIPMap levels");
            // the MIPMap pyramid: the caller's own levels, or built here from a power-of-two level 0
            std::vector<float> pyr;
            std::vector<int> offs;
            int nLevels = 1;
            build_mip_pyramid(t, pyr, offs, &nLevels);
            if (nLevels == 1 && (t.width > 1 || t.height > 1)) ctx->tex_needs_pyramid = true;  // not a power of two, no levels given
            float *dtex;
            if ((rc = dupload(ctx, pool, pyr.data(), pyr.size(), &dtex))) return rc;
            fill_dev_texture(t, dtex, offs, nLevels, tex[i]);
        }
        {
            float lut[128];
            make_ewa_lut(lut);
            float *dlut;
            if ((rc = dupload(ctx, pool, lut, (size_t)128, &dlut))) return rc;
            sc.ewa_lut = dlut;
        }
        DevTexture *dt;
        if ((rc = dupload(ctx, pool, tex.data(), tex.size(), &dt))) return rc;
        sc.textures = dt;
        for (int i = 0; i < d->n_materials; ++i) {
            for (int s = 0; s < GNX_MAT_MAX_RGB; ++s)
                if (d->materials[i].rgb_tex[s] >= d->n_textures) return fail(ctx, GNX_ERR_INVALID, "texture index out of range");
            for (int s = 0; s < GNX_MAT_MAX_F; ++s)
                if (d->materials[i].f_tex[s] >= d->n_textures) return fail(ctx, GNX_ERR_INVALID, "texture index out of range");
        }
    }

    // ---- lights
    for (int i = 0; i < d->n_lights; ++i) {
        const gnx_light &l = d->lights[i];
        if (l.type == GNX_LIGHT_AREA_TRI) {
            if (l.prim < 0 || l.prim >= g.n_prims) return fail(ctx, GNX_ERR_INVALID, "area light primitive out of range");
        } else if (l.type == GNX_LIGHT_INFINITE) {
            if (!d->env.present || d->env.light_index != i) return fail(ctx, GNX_ERR_INVALID, "infinite light without envmap record");
        } else if (l.type == GNX_LIGHT_POINT || l.type == GNX_LIGHT_SPOT || l.type == GNX_LIGHT_DISTANT) {
            ctx->has_next_lights = true;  // rendered by the Whitted / DirectLighting integrators only
        } else if (l.type == GNX_LIGHT_SKYBOX) {
            if (!d->skybox.present || d->skybox.light_index != i) return fail(ctx, GNX_ERR_INVALID, "skybox light without skybox record");
            ctx->has_next_lights = true;
        } else {
            return fail(ctx, GNX_ERR_INVALID, "unknown light type");
        }
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
            // getLightValue indexes (w + h * W) * C + 0..2 with w = u * W, h = v * H, u and v up to 1: one padding row
            std::vector<float> img((size_t)sb.width * (sb.height + 1) * sb.channels + 4, 0.f);
            memcpy(img.data(), sb.data, sizeof(float) * (size_t)sb.width * sb.height * sb.channels);
            float *dimg;
            if ((rc = dupload(ctx, pool, img.data(), img.size(), &dimg))) return rc;
            sc.skybox.data = dimg;
        }
    }
    gnx_light *dl;
    if ((rc = dupload(ctx, pool, d->lights, (size_t)d->n_lights, &dl))) return rc;
    sc.lights = dl;
    sc.n_lights = d->n_lights;
    ctx->n_lights_host = d->n_lights;
    sc.light_nsamples = nullptr;
    if (d->light_n_samples && d->n_lights > 0) {
        for (int i = 0; i < d->n_lights; ++i)
            if (d->light_n_samples[i] < 1 || d->light_n_samples[i] > 4096) return fail(ctx, GNX_ERR_INVALID, "light_n_samples out of range");
        int *dn;
        if ((rc = dupload(ctx, pool, d->light_n_samples, (size_t)d->n_lights, &dn))) return rc;
        sc.light_nsamples = dn;
    }
    if (d->env.present) {
        const gnx_envmap &e = d->env;
        if (!e.texels || !e.cond_func || !e.cond_cdf || !e.cond_int || !e.marg_func || !e.marg_cdf)
            return fail(ctx, GNX_ERR_INVALID, "envmap arrays missing");
        DevEnv &de = sc.env;
        de.present = 1; de.light_index = e.light_index;
        de.w = e.width; de.h = e.height; de.dw = e.dist_w; de.dh = e.dist_h;
        // all of the light's tables in ONE allocation (a single L2 access-policy window can cover them, set_l2_window)
        std::vector<float4> tex;
        pack_env_texels(e.texels, (size_t)e.width * e.height, tex);
        de.marg_int = e.marg_int;
        de.cond_guide = de.marg_guide = nullptr;
        de.cond_g = de.marg_g = 0;
        std::vector<uint16_t> cg, mg;
        if (e.dist_w < 65536 && e.dist_h < 65536 && !getenv("GNX_NO_CDF_GUIDE")) {
            // guide tables for the two inverse-cdf searches of InfiniteAreaLight::Sample_Li (same results, ~half the
            // dependent loads)
            de.cond_g = guide_buckets(e.dist_w);
            de.marg_g = guide_buckets(e.dist_h);
            cg.resize((size_t)e.dist_h * (de.cond_g + 1));
            mg.resize((size_t)de.marg_g + 1);
            for (int v = 0; v < e.dist_h; ++v)
                make_cdf_guide(e.cond_cdf + (size_t)v * (e.dist_w + 1), e.dist_w + 1, de.cond_g, cg.data() + (size_t)v * (de.cond_g + 1));
            make_cdf_guide(e.marg_cdf, e.dist_h + 1, de.marg_g, mg.data());
        }
        struct Piece { const void *src; size_t bytes; size_t off; };
        Piece pieces[] = {{tex.data(), tex.size() * sizeof(float4), 0},
                          {e.cond_func, (size_t)e.dist_w * e.dist_h * sizeof(float), 0},
                          {e.cond_cdf, (size_t)(e.dist_w + 1) * e.dist_h * sizeof(float), 0},
                          {e.cond_int, (size_t)e.dist_h * sizeof(float), 0},
                          {e.marg_func, (size_t)e.dist_h * sizeof(float), 0},
                          {e.marg_cdf, ((size_t)e.dist_h + 1) * sizeof(float), 0},
                          {cg.data(), cg.size() * sizeof(uint16_t), 0},
                          {mg.data(), mg.size() * sizeof(uint16_t), 0}};
        size_t total = 0;
        for (Piece &pc : pieces) { pc.off = total; total += (pc.bytes + 255) & ~(size_t)255; }
        char *slab = nullptr;
        if ((rc = dupload<char>(ctx, pool, nullptr, total, &slab))) return rc;
        for (const Piece &pc : pieces)
            if (pc.bytes) GNX_CUDA(ctx, cudaMemcpy(slab + pc.off, pc.src, pc.bytes, cudaMemcpyHostToDevice));
        de.texels = (const float4 *)(slab + pieces[0].off);
        de.cond_func = (const float *)(slab + pieces[1].off);
        de.cond_cdf = (const float *)(slab + pieces[2].off);
        de.cond_int = (const float *)(slab + pieces[3].off);
        de.marg_func = (const float *)(slab + pieces[4].off);
        de.marg_cdf = (const float *)(slab + pieces[5].off);
        if (!cg.empty()) {
            de.cond_guide = (const uint16_t *)(slab + pieces[6].off);
            de.marg_guide = (const uint16_t *)(slab + pieces[7].off);
        }
        ctx->env_base = slab;
        ctx->env_bytes = total;
        memcpy(de.l2w.m, e.light_to_world, 64);
        memcpy(de.w2l.m, e.world_to_light, 64);
        de.world_radius = e.world_radius;
    }
    // uniform light distribution: Distribution1D over n ones (core/LightDistribution.cpp:35-38)
    if (d->n_lights > 0) {
        std::vector<float> func, cdf;
        float funcInt = uniform_light_distribution(d->n_lights, func, cdf);
        if ((rc = dupload(ctx, pool, func.data(), func.size(), &df))) return rc; sc.ld.uni_func = df;
        if ((rc = dupload(ctx, pool, cdf.data(), cdf.size(), &df))) return rc; sc.ld.uni_cdf = df;
        sc.ld.uni_int = funcInt;
        ctx->one_func = sc.ld.uni_func; ctx->one_cdf = sc.ld.uni_cdf; ctx->one_int = funcInt;
        // power light distribution (PowerLightDistribution, core/LightDistribution.cpp:44-50)
        std::vector<float> power((size_t)d->n_lights);
        ctx->pow_func = ctx->pow_cdf = nullptr; ctx->pow_error.clear();
        for (int i = 0; i < d->n_lights; ++i) {
            if (d->light_power) power[i] = d->light_power[i];
            else if (!derive_light_power(d->lights[i], &power[i])) {
                ctx->pow_error = "power light distribution: the scene description has no light_power table and light " +
                                 std::to_string(i) + " is an environment / skybox light";
                break;
            }
        }
        if (ctx->pow_error.empty()) {
            ctx->pow_int = power_light_distribution(d->n_lights, power.data(), func, cdf);
            if ((rc = dupload(ctx, pool, func.data(), func.size(), &df))) return rc; ctx->pow_func = df;
            if ((rc = dupload(ctx, pool, cdf.data(), cdf.size(), &df))) return rc; ctx->pow_cdf = df;
        }
    }
    sc.ld.mode = GNX_LIGHTS_UNIFORM;

    // ---- media (records only; the volumetric kernels validate them at render time)
    if (d->n_media > 0) {
        std::vector<DevMedium> med((size_t)d->n_media);
        for (int i = 0; i < d->n_media; ++i) {
            const gnx_medium &m = d->media[i];
            DevMedium &dmv = med[i];
            const float *ddens = nullptr;
            if (m.type == GNX_MEDIUM_GRID) {
                if (!m.density || m.nx <= 0 || m.ny <= 0 || m.nz <= 0) return fail(ctx, GNX_ERR_INVALID, "grid medium without density");
                if ((rc = dupload(ctx, pool, m.density, (size_t)m.nx * m.ny * m.nz, &df))) return rc;
                ddens = df;
            }
            fill_dev_medium(m, ddens, dmv);
        }
        DevMedium *dmed;
        if ((rc = dupload(ctx, pool, med.data(), med.size(), &dmed))) return rc;
        sc.media = dmed;
        sc.n_media = d->n_media;
    }

    // ---- camera
    memcpy(sc.cam.r2c.m, d->camera.raster_to_camera, 64);
    memcpy(sc.cam.c2w.m, d->camera.camera_to_world, 64);
    sc.cam.lens_radius = d->camera.lens_radius;
    sc.cam.focal_distance = d->camera.focal_distance;
    sc.cam.medium = d->camera.medium;
    sc.cam.dx_camera = V3(d->camera.dx_camera[0], d->camera.dx_camera[1], d->camera.dx_camera[2]);
    sc.cam.dy_camera = V3(d->camera.dy_camera[0], d->camera.dy_camera[1], d->camera.dy_camera[2]);

    // ---- sampler
    const gnx_sampler &s = d->sampler;
    if (s.type != GNX_SAMPLER_HALTON && s.type != GNX_SAMPLER_PCG32 && s.type != GNX_SAMPLER_SOBOL) return fail(ctx, GNX_ERR_INVALID, "unknown sampler type");
    sc.smp.sobol32 = nullptr; sc.smp.sobol_vdc = sc.smp.sobol_vdc_inv = nullptr;
    sc.smp.sobol_dims = sc.smp.sobol_log2res = sc.smp.sobol_res = 0;
    if (s.type == GNX_SAMPLER_SOBOL) {
        if (!s.sobol_matrices32 || s.n_sobol_dimensions < 5 || s.sobol_log2_resolution < 0 || s.sobol_log2_resolution > 15 ||
            s.sobol_resolution != (1 << s.sobol_log2_resolution) || (s.sobol_log2_resolution > 0 && (!s.sobol_vdc || !s.sobol_vdc_inv)))
            return fail(ctx, GNX_ERR_INVALID, "bad Sobol parameters");
        uint32_t *dm; uint64_t *dv;
        if ((rc = dupload(ctx, pool, s.sobol_matrices32, (size_t)s.n_sobol_dimensions * kSobolMatrixSize, &dm))) return rc;
        sc.smp.sobol32 = dm;
        std::vector<uint64_t> zero(kSobolMatrixSize, 0);
        if ((rc = dupload(ctx, pool, s.sobol_vdc ? s.sobol_vdc : zero.data(), (size_t)kSobolMatrixSize, &dv))) return rc;
        sc.smp.sobol_vdc = dv;
        if ((rc = dupload(ctx, pool, s.sobol_vdc_inv ? s.sobol_vdc_inv : zero.data(), (size_t)kSobolMatrixSize, &dv))) return rc;
        sc.smp.sobol_vdc_inv = dv;
        sc.smp.sobol_dims = s.n_sobol_dimensions; sc.smp.sobol_log2res = s.sobol_log2_resolution; sc.smp.sobol_res = s.sobol_resolution;
        // the first dimensions' matrices in __constant__ memory (one broadcast per read: the lanes of a warp walk the index bits in step)
        std::vector<uint32_t> head((size_t)kSobolConstDims * kSobolMatrixSize, 0u);
        memcpy(head.data(), s.sobol_matrices32, sizeof(uint32_t) * (size_t)std::min(kSobolConstDims, s.n_sobol_dimensions) * kSobolMatrixSize);
        for (auto fn : sobol_uploaders()) fn(head.data());
        GNX_CUDA(ctx, cudaGetLastError());
    }
    if (s.type == GNX_SAMPLER_HALTON && (s.base_scales[0] <= 0 || s.base_scales[1] <= 0 || s.sample_stride <= 0))
        return fail(ctx, GNX_ERR_INVALID, "bad Halton parameters");
    sc.smp.type = s.type;
    sc.smp.base_scale0 = s.base_scales[0]; sc.smp.base_scale1 = s.base_scales[1];
    sc.smp.base_exp0 = s.base_exponents[0]; sc.smp.base_exp1 = s.base_exponents[1];
    sc.smp.stride = s.sample_stride;
    sc.smp.mult_inv0 = s.mult_inverse[0]; sc.smp.mult_inv1 = s.mult_inverse[1];
    sc.smp.at_center = s.sample_at_pixel_center;
    sc.smp.spp = s.samples_per_pixel > 0 ? s.samples_per_pixel : 1;
    sc.smp.stride_over_scale0 = s.type == GNX_SAMPLER_HALTON ? s.sample_stride / s.base_scales[0] : 0;
    sc.smp.stride_over_scale1 = s.type == GNX_SAMPLER_HALTON ? s.sample_stride / s.base_scales[1] : 0;
    {
        std::vector<int> primes, sums;
        make_primes(primes, sums);
        std::vector<uint16_t> perms;
        const uint16_t *hp = s.perms;
        size_t np = (size_t)s.n_perm_entries;
        if (!hp) { make_permutations(primes, perms); hp = perms.data(); np = perms.size(); }
        size_t need = (size_t)sums.back() + primes.back();
        if (np < need) return fail(ctx, GNX_ERR_INVALID, "Halton permutation table too short");
        std::vector<uint4> dims;
        make_dim_table(primes, sums, dims);
        uint16_t *dp; int *di; uint4 *dd;
        if ((rc = dupload(ctx, pool, hp, np, &dp))) return rc; sc.smp.perms = dp;
        if ((rc = dupload(ctx, pool, primes.data(), primes.size(), &di))) return rc; sc.smp.primes = di;
        if ((rc = dupload(ctx, pool, dims.data(), dims.size(), &dd))) return rc; sc.smp.dims = dd;
        sc.smp.n_primes = (int)primes.size();
    }
    ctx->sc = sc;
    ctx->has_scene = true;
    if (ctx->qs_probe_host) ctx->qs_probe_host[1] = 0;  // another scene: its hit density is not known yet
    return GNX_OK;
}

}  // extern "C"

// SpatialLightDistribution ctor: voxel resolution (core/LightDistribution.cpp:70-87) + dense tables.
static int ensure_light_distribution(gnx_ctx *ctx, int strategy) {
    DeviceScene &sc = ctx->sc;
    // CreateLightSampleDistribution: uniform when asked for or when exactly one light exists
    sc.ld.uni_func = ctx->one_func; sc.ld.uni_cdf = ctx->one_cdf; sc.ld.uni_int = ctx->one_int;
    if (strategy == GNX_LIGHTS_UNIFORM || sc.n_lights <= 1) { sc.ld.mode = GNX_LIGHTS_UNIFORM; return GNX_OK; }
    if (strategy == GNX_LIGHTS_POWER) {
        if (!ctx->pow_func) return fail(ctx, GNX_ERR_UNSUPPORTED, ctx->pow_error.c_str());
        sc.ld.uni_func = ctx->pow_func; sc.ld.uni_cdf = ctx->pow_cdf; sc.ld.uni_int = ctx->pow_int;
        sc.ld.mode = GNX_LIGHTS_POWER;
        return GNX_OK;
    }
    if (!ctx->spatial_built) {
        size_t nv = spatial_voxel_resolution(ctx->wb, sc.ld.nvox);
        if (nv * (size_t)sc.n_lights > ((size_t)1 << 28))
            return fail(ctx, GNX_ERR_UNSUPPORTED, "too many lights for dense spatial light tables");
        float *f, *c, *fi;
        int rc;
        if ((rc = dupload<float>(ctx, ctx->scene_allocs, nullptr, nv * sc.n_lights, &f))) return rc;
        if ((rc = dupload<float>(ctx, ctx->scene_allocs, nullptr, nv * (sc.n_lights + 1), &c))) return rc;
        if ((rc = dupload<float>(ctx, ctx->scene_allocs, nullptr, nv, &fi))) return rc;
        sc.ld.sp_func = f; sc.ld.sp_cdf = c; sc.ld.sp_int = fi;
        sc.ld.mode = GNX_LIGHTS_SPATIAL;
        int blocks = (int)std::min<size_t>((nv + 127) / 128, (size_t)ctx->sm_count * 16);
        k_build_spatial<<<blocks, 128, 0, ctx->stream>>>(sc, f, c, fi);
        GNX_CUDA(ctx, cudaGetLastError());
        GNX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->spatial_built = true;
    }
    sc.ld.mode = GNX_LIGHTS_SPATIAL;
    return GNX_OK;
}

static int ensure_wavefront(gnx_ctx *ctx, int capacity, int npix, int npixFrame) {
    if (capacity > ctx->capacity) {
        free_pool(ctx->wave_allocs);
        ctx->capacity = 0;
        std::vector<void *> &pool = ctx->wave_allocs;
        int rc;
        size_t n = (size_t)capacity;
        char *records = nullptr;  // one 96-byte record per path slot (gnx_scene.cuh), the accumulators as arrays next to it
        if ((rc = dupload<char>(ctx, pool, nullptr, n * kPathRecordBytes, &records))) return rc;
        ctx->ps.bind(records);
        if ((rc = dupload<float4>(ctx, pool, nullptr, n, &ctx->ps.L))) return rc;
        if ((rc = dupload<float4>(ctx, pool, nullptr, n, &ctx->ps.Lb))) return rc;
        if ((rc = dupload<float4>(ctx, pool, nullptr, n, &ctx->ps.La))) return rc;
        ctx->ps.film_off = ctx->ps.Lb;  // dead by the time the film kernels run (k_film_prepare folds Lb into L first)
        if ((rc = dupload<int>(ctx, pool, nullptr, n, &ctx->q.extend_q[0]))) return rc;
        if ((rc = dupload<int>(ctx, pool, nullptr, n, &ctx->q.extend_q[1]))) return rc;
        if ((rc = dupload<int>(ctx, pool, nullptr, n * kNumShadeTypes, &ctx->q.shade_q))) return rc;
        if ((rc = dupload<ShadowItem>(ctx, pool, nullptr, n * 2, &ctx->q.shadow_q))) return rc;
        if ((rc = dupload<ProbeItem>(ctx, pool, nullptr, n, &ctx->q.probe_q))) return rc;
        if ((rc = dupload<int>(ctx, pool, nullptr, kNumCounters, &ctx->q.counts))) return rc;
        ctx->q.miss_q = nullptr;
        ctx->q.capacity = capacity;
        ctx->capacity = capacity;
        GNX_CUDA(ctx, cudaDeviceSynchronize());
    }
    npix = std::max(npix, npixFrame);
    if (npix > ctx->film_pixels) {
        if (ctx->accum) cudaFree(ctx->accum);
        if (ctx->rgba) cudaFree(ctx->rgba);
        ctx->accum = ctx->rgba = nullptr;
        ctx->film_pixels = 0;
        GNX_CUDA(ctx, cudaMalloc((void **)&ctx->accum, (size_t)npix * sizeof(float4)));
        GNX_CUDA(ctx, cudaMalloc((void **)&ctx->rgba, (size_t)npix * sizeof(float4)));
        ctx->film_pixels = npix;
    }
    return GNX_OK;
}

static int ensure_volwave(gnx_ctx *ctx, int capacity) {
    if (capacity <= ctx->vw_capacity) return GNX_OK;
    free_pool(ctx->vw_allocs);
    ctx->vw_capacity = 0;
    std::vector<void *> &pool = ctx->vw_allocs;
    const size_t n = (size_t)capacity;
    int rc;
    char *records = nullptr;  // one 160-byte record per path (gnx_volwave.cuh)
    if ((rc = dupload<char>(ctx, pool, nullptr, n * kVolRecordBytes, &records))) return rc;
    ctx->vw.bind(records);
    ctx->vw_capacity = capacity;
    return GNX_OK;
}

static int ensure_rebin(gnx_ctx *ctx, int capacity) {
    if ((size_t)capacity <= ctx->rb_slots) return GNX_OK;
    free_pool(ctx->rb_allocs);
    ctx->rb_slots = 0;
    int rc;
    for (int k = 0; k < 2; ++k) {
        if ((rc = dupload<uint32_t>(ctx, ctx->rb_allocs, nullptr, (size_t)capacity, &ctx->rb_keys[k]))) return rc;
        if ((rc = dupload<int>(ctx, ctx->rb_allocs, nullptr, (size_t)capacity, &ctx->rb_vals[k]))) return rc;
    }
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, ctx->rb_keys[0], ctx->rb_keys[1], ctx->rb_vals[0], ctx->rb_vals[1], capacity, 0, 31);
    char *tmp = nullptr;
    if ((rc = dupload<char>(ctx, ctx->rb_allocs, nullptr, bytes, &tmp))) return rc;
    ctx->rb_tmp = tmp;
    ctx->rb_tmp_bytes = bytes;
    ctx->rb_slots = (size_t)capacity;
    return GNX_OK;
}

static int ensure_queue_sort(gnx_ctx *ctx, int capacity) {
    const size_t words = ((size_t)capacity + 31) / 32;
    if (words <= ctx->qs_words) return GNX_OK;
    free_pool(ctx->qs_allocs);
    ctx->qs_words = 0;
    int rc;
    if ((rc = dupload<unsigned>(ctx, ctx->qs_allocs, nullptr, words, &ctx->qs_bits))) return rc;   // zeroed; k_qs_emit leaves it zeroed
    if ((rc = dupload<int>(ctx, ctx->qs_allocs, nullptr, (words + kQsBlock - 1) / kQsBlock, &ctx->qs_blocks))) return rc;
    if (!ctx->qs_probe_dev) {
        GNX_CUDA(ctx, cudaMalloc((void **)&ctx->qs_probe_dev, sizeof(int)));
        GNX_CUDA(ctx, cudaMallocHost((void **)&ctx->qs_probe_host, 2 * sizeof(int)));
        ctx->qs_probe_host[0] = ctx->qs_probe_host[1] = 0;
    }
    ctx->qs_words = words;
    return GNX_OK;
}

static int ensure_whitted(gnx_ctx *ctx, int capacity, int nLights) {
    const size_t need = (size_t)capacity * (size_t)nLights;
    if (need <= ctx->ww_slots) return GNX_OK;
    free_pool(ctx->ww_allocs);
    ctx->ww_slots = 0;
    int rc;
    if ((rc = dupload<ShadowItem>(ctx, ctx->ww_allocs, nullptr, need, &ctx->ww_items))) return rc;
    if ((rc = dupload<float4>(ctx, ctx->ww_allocs, nullptr, need, &ctx->ww_planes))) return rc;
    ctx->ww_slots = need;
    return GNX_OK;
}

static int validate_params(gnx_ctx *ctx, const gnx_render_params *p) {
    if (!ctx->has_scene) return fail(ctx, GNX_ERR_NO_SCENE, "no scene uploaded");
    if (!p || p->width <= 0 || p->height <= 0 || p->spp <= 0 || p->first_sample < 0 || p->max_depth < 0 || p->max_depth > 250)
        return fail(ctx, GNX_ERR_INVALID, "bad render parameters");
    if ((long long)p->width * p->height > (1ll << 28)) return fail(ctx, GNX_ERR_INVALID, "image too large");
    if (p->integrator < GNX_INTEGRATOR_PATH || p->integrator > GNX_INTEGRATOR_DIRECT_ALL) return fail(ctx, GNX_ERR_INVALID, "unknown integrator");
    const bool recursive = p->integrator >= GNX_INTEGRATOR_WHITTED;
    if (p->integrator == GNX_INTEGRATOR_DIRECT_ALL && 5 + 4 * (long long)p->max_depth * ctx->sc.n_lights + 64 > 1000)
        return fail(ctx, GNX_ERR_UNSUPPORTED, "UniformSampleAll: the sample arrays of maxDepth x lights need more Halton dimensions than the sampler has (1000)");
    if (p->integrator == GNX_INTEGRATOR_VOLPATH && ctx->has_next_lights)
        return fail(ctx, GNX_ERR_UNSUPPORTED, "point / spot / distant / skybox lights are rendered by the Path, Whitted and DirectLighting integrators, not by VolPath");
    // the recursive integrators keep their depth-first frames in a fixed array (gnx_whitted.cuh, kMaxRecDepth)
    if (recursive && p->max_depth > kMaxRecDepth)
        return fail(ctx, GNX_ERR_UNSUPPORTED, "Whitted / DirectLighting: max_depth above 16 is not supported (fixed recursion frame stack)");
    if (recursive && ctx->sc.smp.type == GNX_SAMPLER_PCG32) return fail(ctx, GNX_ERR_UNSUPPORTED, "Whitted / DirectLighting use a GlobalSampler (Halton or Sobol)");
    if (recursive && ctx->sc.n_media > 0) return fail(ctx, GNX_ERR_UNSUPPORTED, "Whitted / DirectLighting ignore participating media");
    if (p->integrator != GNX_INTEGRATOR_PATH && ctx->tex_needs_pyramid)
        return fail(ctx, GNX_ERR_UNSUPPORTED, "an image texture has a single level that is not a power of two: pass the MIPMap's own levels (gnx_texture::n_levels) for the integrators that filter with ray differentials");
    if (p->integrator == GNX_INTEGRATOR_PATH && ctx->sc.smp.type == GNX_SAMPLER_PCG32)
        return fail(ctx, GNX_ERR_UNSUPPORTED, "the wavefront PathIntegrator keeps a GlobalSampler (index, dimension) per path (Halton or Sobol); the PCG32 stream sampler is for VolPath");
    if (p->integrator == GNX_INTEGRATOR_PATH && ctx->sc.n_media > 0)
        return fail(ctx, GNX_ERR_UNSUPPORTED, "scene has participating media: use GNX_INTEGRATOR_VOLPATH (PathIntegrator ignores media)");
    if (p->film != GNX_FILM_BOX && p->film != GNX_FILM_GAUSSIAN && p->film != GNX_FILM_GAUSSIAN_SUMS)
        return fail(ctx, GNX_ERR_INVALID, "unknown film");
    if (p->film != GNX_FILM_BOX && !(p->filter_radius > 0.f && p->filter_radius <= 16.f && p->filter_alpha >= 0.f))
        return fail(ctx, GNX_ERR_INVALID, "Gaussian film: filter_radius must be in (0, 16] and filter_alpha >= 0");
    if (p->partition != GNX_PARTITION_SAMPLES && p->partition != GNX_PARTITION_TILES) return fail(ctx, GNX_ERR_INVALID, "unknown partition");
    const int nShare = ctx->peers.empty() ? ctx->n_ranks : 1 + (int)ctx->peers.size();
    if (nShare > 1 && p->partition == GNX_PARTITION_TILES && p->film != GNX_FILM_BOX)
        return fail(ctx, GNX_ERR_UNSUPPORTED, "the tile partition finishes every pixel on one device: box film only (use GNX_PARTITION_SAMPLES with a Gaussian film)");
    (void)nShare;
    // the Halton index must fit the 32-bit path state
    const unsigned long long lastSample = (unsigned long long)p->first_sample + (unsigned long long)p->spp;
    unsigned long long maxIdx = lastSample * (unsigned long long)ctx->sc.smp.stride;
    if (ctx->sc.smp.type == GNX_SAMPLER_HALTON && maxIdx >= (1ull << 32)) return fail(ctx, GNX_ERR_UNSUPPORTED, "sample index exceeds 32 bits");
    // Sobol: index = sample << 2 log2(resolution) | pixel bits
    if (ctx->sc.smp.type == GNX_SAMPLER_SOBOL && (lastSample << (2 * ctx->sc.smp.sobol_log2res)) >= (1ull << 32))
        return fail(ctx, GNX_ERR_UNSUPPORTED, "Sobol sample index exceeds 32 bits (samples x resolution^2 must stay below 2^32)");
    if (ctx->sc.smp.type == GNX_SAMPLER_PCG32 && lastSample >= (1ull << 20)) return fail(ctx, GNX_ERR_UNSUPPORTED, "PCG32 stream ids hold 20 bits of sample number");
    return GNX_OK;
}

// Marks the node + triangle arrays as persisting in L2 for work launched on `st`: the random gathers of
// the traversal then compete less with the streaming path-state / queue traffic (C2: 98 MB of geometry
// against ~1 GB of state per bounce, 126 MB of L2).
static void set_l2_window(gnx_ctx *ctx, cudaStream_t st) {
    if (ctx->l2_env && ctx->env_base && ctx->sc.env.present) {
        if (ctx->l2_persist_max < 0) {  // (the property query takes milliseconds: once per context)
            int a = 0, b = 0;
            cudaDeviceGetAttribute(&a, cudaDevAttrMaxPersistingL2CacheSize, ctx->device);
            cudaDeviceGetAttribute(&b, cudaDevAttrMaxAccessPolicyWindowSize, ctx->device);
            ctx->l2_persist_max = a;
            ctx->l2_window_max = b;
            cudaGetLastError();
        }
        if (ctx->l2_persist_max <= 0) return;
        const size_t setAside = std::min<size_t>((size_t)ctx->l2_persist_max, ctx->env_bytes);
        if (ctx->l2_set_aside != setAside) { cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, setAside); ctx->l2_set_aside = setAside; }
        cudaStreamAttrValue attr{};
        const size_t win = std::min<size_t>(ctx->env_bytes, (size_t)ctx->l2_window_max);
        attr.accessPolicyWindow.base_ptr = ctx->env_base;
        attr.accessPolicyWindow.num_bytes = win;
        attr.accessPolicyWindow.hitRatio = win > 0 ? std::min(1.0f, (float)setAside / (float)win) : 0.f;
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
        cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr);
        if (ctx->stream_any) cudaStreamSetAttribute(ctx->stream_any, cudaStreamAttributeAccessPolicyWindow, &attr);
        cudaGetLastError();
        return;
    }
    if (!ctx->l2_persist || !ctx->geom_base) return;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, ctx->device) != cudaSuccess || prop.persistingL2CacheMaxSize <= 0) return;
    size_t setAside = std::min<size_t>((size_t)prop.persistingL2CacheMaxSize, ctx->geom_bytes);
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, setAside);
    cudaStreamAttrValue attr{};
    size_t win = std::min<size_t>(ctx->geom_bytes, (size_t)prop.accessPolicyMaxWindowSize);
    attr.accessPolicyWindow.base_ptr = ctx->geom_base;
    attr.accessPolicyWindow.num_bytes = win;
    attr.accessPolicyWindow.hitRatio = win > 0 ? std::min(1.0f, (float)setAside / (float)win) : 0.f;
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr);
    cudaGetLastError();
}

// Queues one device's share of a render on `userStream`.  share.n == 1: the whole job.  Otherwise p_in describes the
// WHOLE job and this device renders its sample range (its frame then holds sum / total spp, alpha 1 on share 0 and 0
// elsewhere) or its interleaved tiles (scattered into a zeroed full frame); the caller sums the frames.
static int render_impl(gnx_ctx *ctx, const gnx_render_params *p_in, float *rgba_dev_out, cudaStream_t userStream, gnx_stats *stats,
                       bool callerSyncs = false, Share share = Share()) {
    int rc = validate_params(ctx, p_in);
    if (rc) return rc;
    GNX_CUDA(ctx, cudaSetDevice(ctx->device));
    gnx_render_params pv = *p_in;
    const gnx_render_params *p = &pv;
    const int npixFrame = p->width * p->height;
    const bool tiles = share.n > 1 && share.partition == GNX_PARTITION_TILES;
    int tilesX = 0, tilesY = 0;
    if (share.n > 1 && !tiles) {
        // contiguous sample range of share idx (the first spp % n shares take one sample more)
        const int base = p_in->spp / share.n, rem = p_in->spp % share.n;
        pv.spp = base + (share.idx < rem ? 1 : 0);
        pv.first_sample = p_in->first_sample + share.idx * base + std::min(share.idx, rem);
        pv.spp_normalize = p_in->spp_normalize > 0 ? p_in->spp_normalize : p_in->spp;
        if (pv.spp == 0) {  // fewer samples than devices: an empty share contributes zeros
            float *o = rgba_dev_out;
            if (!o) { if ((rc = ensure_wavefront(ctx, 0, 0, npixFrame))) return rc; o = (float *)ctx->rgba; }
            GNX_CUDA(ctx, cudaMemsetAsync(o, 0, (size_t)npixFrame * sizeof(float4), userStream));
            if (stats) memset(stats, 0, sizeof(*stats));
            return GNX_OK;
        }
    }
    if (tiles) { tilesX = (p->width + kTile - 1) / kTile; tilesY = (p->height + kTile - 1) / kTile; }
    const bool recursiveInteg = p->integrator >= GNX_INTEGRATOR_WHITTED;
    if (!recursiveInteg && (rc = ensure_light_distribution(ctx, p->light_strategy))) return rc;
    // pixels this device renders: the image, or its tiles (edge tiles padded to kTile x kTile)
    const int npix = tiles ? local_tile_count(tilesX, tilesY, share.n, share.idx) * kTile * kTile : npixFrame;
    // Paths in flight per wavefront batch.  Late bounces carry few rays and every launch has a tail, so
    // the batch is made as large as memory comfortably allows (profiles/README.md: 4 M -> 64 M slots took
    // C2 from 126 ms to 71 ms): up to 64 M slots, and never more than a quarter of the free HBM.
    // WhittedIntegrator with the first vertex staged: needs per-light buffers, so only with a handful of lights (the reference
    // UI's scene has 3, the W1 workload 6); a plane index light * capacity + slot must also fit the item's int
    const bool whittedStaged = p->integrator == GNX_INTEGRATOR_WHITTED && ctx->whitted_staged &&
                               ctx->sc.n_lights >= 1 && ctx->sc.n_lights <= 8;
    int batch_spp = p->batch_spp;
    if (batch_spp <= 0) {
        size_t freeB = 0, totalB = 0;
        long long slots = 64ll << 20;
        // (cudaMemGetInfo costs ~0.5 ms: only asked when the buffers of an earlier call do not already cover the batch)
        const long long want = std::min(slots, (long long)npix * p->spp);
        if ((long long)ctx->capacity < want && cudaMemGetInfo(&freeB, &totalB) == cudaSuccess) {
            // path record 96 B + accumulators 48 B + queues 180 B, rounded up; the VolPath wavefront keeps 160 B more per path
            long long bytesPerSlot = (p->integrator == GNX_INTEGRATOR_VOLPATH && !ctx->vol_megakernel) ? 480 : 320;
            if (whittedStaged) bytesPerSlot += 64ll * ctx->sc.n_lights;  // a shadow item and a contribution per light
            long long avail = (long long)((freeB + (size_t)ctx->capacity * bytesPerSlot) / 4 / bytesPerSlot);
            slots = std::max(1ll << 20, std::min(slots, avail));
        }
        batch_spp = (int)std::max(1ll, slots / std::max(1, npix));
    }
    batch_spp = std::min(batch_spp, p->spp);
    if ((long long)npix * batch_spp > (1ll << 30)) batch_spp = std::max(1, (int)((1ll << 30) / std::max(1, npix)));
    if ((rc = ensure_wavefront(ctx, npix * batch_spp, npix, npixFrame))) return rc;
    const bool volWave = p->integrator == GNX_INTEGRATOR_VOLPATH && !ctx->vol_megakernel;
    const bool tex = ctx->n_textures_host > 0;  // image textures: integrators with a RayDifferential filter them through the MIPMap
    if (volWave && (rc = ensure_volwave(ctx, ctx->capacity))) return rc;
    if (whittedStaged && (rc = ensure_whitted(ctx, ctx->capacity, ctx->sc.n_lights))) return rc;
    if (ctx->sort_queues && p->integrator == GNX_INTEGRATOR_PATH && (rc = ensure_queue_sort(ctx, ctx->capacity))) return rc;
    if (ctx->rebin && p->integrator == GNX_INTEGRATOR_PATH && (rc = ensure_rebin(ctx, ctx->capacity))) return rc;

    // escaped rays of scenes with a SkyBoxLight are queued for k_escape
    if (ctx->sc.skybox.present && (p->integrator == GNX_INTEGRATOR_PATH || whittedStaged) && !ctx->q.miss_q && ctx->capacity > 0)
        if ((rc = dupload<int>(ctx, ctx->wave_allocs, nullptr, (size_t)ctx->capacity, &ctx->q.miss_q))) return rc;
    Queues qv = ctx->q;
    if (!ctx->sc.skybox.present) qv.miss_q = nullptr;
    cudaStream_t st = userStream;  // gnx_render_device maps a NULL stream to the legacy default stream (see gnxrt.h)
    set_l2_window(ctx, st);
    const DeviceScene &sc = ctx->sc;
    // the second accumulator is only needed when shadow A and B rays share a launch
    PathState psv = ctx->ps;
    const bool wantMixed = ctx->merge_extend && ctx->merge_shadow && p->integrator == GNX_INTEGRATOR_PATH && sc.n_lights > 0 &&
                           !((ctx->shade_type_mask >> (kNumShadeTypes - 1)) & 1u);
    if (!wantMixed) psv.La = nullptr;
    if (!wantMixed && !(sc.env.present && ctx->merge_shadow)) psv.Lb = nullptr;
    const int gridTrace = ctx->grid_trace, gridShade = ctx->grid_shade, gridWide = ctx->sm_count * 16;
    const bool hasNull = (ctx->shade_type_mask >> (kNumShadeTypes - 1)) & 1u;
    unsigned long long launches = 0;

    GNX_CUDA(ctx, cudaMemsetAsync(ctx->d_stats, 0, sizeof(DevStats), st));
    // per-stage timers cost two event records per kernel group (~1 us each); on whenever the caller
    // asks for stats, unless GNX_STAGE_TIMERS=0
    StageTimer tm{ctx, st, stats != nullptr && ctx->stage_timers};
    ctx->ev_used = 0;
    ctx->ev_stage.clear();
    unsigned long long extendLaunches = 0, vpRounds = 0;
    GNX_CUDA(ctx, cudaEventRecord(ctx->ev0, st));
    GNX_CUDA(ctx, cudaMemsetAsync(ctx->accum, 0, (size_t)npix * sizeof(float4), st));

    // film: the reference's box average, or the Gaussian reconstruction filter (gnx_film.cuh)
    const bool gaussian = p->film != GNX_FILM_BOX;
    FilmFilter filt{};
    if (gaussian) {
        filt.radius = p->filter_radius; filt.alpha = p->filter_alpha;
        filt.expv = std::exp(-filt.alpha * filt.radius * filt.radius);
        filt.reach = (int)std::floor(filt.radius + 0.5f);
    }
    auto accumulate = [&](const PathState &psb, const RenderConsts &rcb) -> int {
        if (!gaussian) { k_accumulate<<<gridWide, 256, 0, st>>>(psb, ctx->accum, rcb); return 1; }
        k_film_prepare<<<gridWide, 256, 0, st>>>(sc, psb, rcb);
        // tiled gather when a stage of at least one sample per pixel fits 72 KB of shared memory (radius < 4.5), else per-pixel
        const int nw = 2 * filt.reach + 1, srcPx = (kFilmTW + 2 * filt.reach) * (kFilmTH + 2 * filt.reach);
        const size_t perSample = (size_t)(3 + 2 * nw) * srcPx * sizeof(float);
        int chunk = (int)std::min<size_t>((size_t)72 * 1024 / perSample, 5);
        if (chunk % 2 == 0 && chunk > 0) --chunk;  // the kernel's sample stride is chunk | 1: an odd chunk wastes no padding
        if (ctx->film_chunk > 0 && ctx->film_chunk <= chunk) chunk = ctx->film_chunk;  // GNX_FILM_CHUNK (tuning)
        if (chunk >= 1 && !ctx->film_simple) {
            const int tilesX = (rcb.width + kFilmTW - 1) / kFilmTW, tilesY = (rcb.height + kFilmTH - 1) / kFilmTH;
            const int grid = std::min(tilesX * tilesY, ctx->sm_count * 8);
            const size_t smem = perSample * (size_t)(chunk | 1);
            auto launch = [&](auto kern) {
                kern<<<grid, kFilmTW * kFilmTH, smem, st>>>(psb, ctx->accum, rcb, filt, chunk, tilesX, tilesY);
            };
            if (filt.reach == 2 && chunk == 3) launch(k_accumulate_gauss_tiled<2, 3>);
            else if (filt.reach == 2 && chunk == 1) launch(k_accumulate_gauss_tiled<2, 1>);
            else if (filt.reach == 1 && chunk == 5) launch(k_accumulate_gauss_tiled<1, 5>);
            else if (filt.reach == 3 && chunk == 1) launch(k_accumulate_gauss_tiled<3, 1>);
            else launch(k_accumulate_gauss_tiled<0, 0>);
        } else k_accumulate_gauss<<<gridWide, 128, 0, st>>>(psb, ctx->accum, rcb, filt);
        return 2;
    };

    // queue `t` of the shade queues back into slot order (k_qs_*; the kernels leave a sparse queue alone).  PathIntegrator only:
    // measured without effect on the VolPath phase queues (C4 152.2 vs 151.8 ms) and on the staged Whitted vertex (camera rays
    // arrive almost in order: U1w 34.2 vs 33.9 ms)
    auto sortQueue = [&](const Queues &qq, int t, int nSlots) {
        const int nWords = (nSlots + 31) / 32, nBlocks = (nWords + kQsBlock - 1) / kQsBlock;
        int *list = qq.shade_q + (size_t)t * qq.capacity;
        const int *cnt = qq.counts + kCntShade0 + t;
        k_qs_mark<<<gridWide, 256, 0, st>>>(list, cnt, ctx->qs_bits, nSlots);
        k_qs_count<<<nBlocks, kQsBlock, 0, st>>>(ctx->qs_bits, nWords, ctx->qs_blocks, cnt, nSlots);
        k_qs_scan<<<1, 1024, 0, st>>>(ctx->qs_blocks, nBlocks, cnt, nSlots);
        k_qs_emit<<<nBlocks, kQsBlock, 0, st>>>(ctx->qs_bits, nWords, ctx->qs_blocks, list, cnt, nSlots);
        launches += 4;
    };
    for (int done = 0; done < p->spp; done += batch_spp) {
        RenderConsts rcn{};
        rcn.width = p->width; rcn.height = p->height; rcn.npix = npix;
        rcn.max_depth = p->max_depth;
        rcn.rr_threshold = p->rr_threshold;
        rcn.batch_spp = std::min(batch_spp, p->spp - done);
        rcn.first_sample = p->first_sample + done;
        rcn.capacity = ctx->capacity;
        if (tiles) { rcn.tile_n = share.n; rcn.tile_dev = share.idx; rcn.tiles_x = tilesX; rcn.tiles_y = tilesY; }
        if (npix == 0) break;  // more devices than tiles
        if (volWave) {
            // VolPath as a staged wavefront (gnx_volwave.cuh): the start launch carries every path to its first tracking
            // walk (or its end); then rounds of { k_vp_track over the queued walks, k_vp_logic per resume phase } until no
            // walk is queued any more.  The queue length is read back every few rounds (the loop has no fixed depth:
            // medium boundaries do not count as bounces).
            const int gl = ctx->grid_vp_logic, gt = ctx->grid_vp_track;
            k_vp_reset<<<1, 32, 0, st>>>(qv.counts, -1);
            tm.begin(ST_VP_EXTEND);
            if (tex) k_vp_logic<VK_EXTEND, true><<<gl, kBlock, 0, st>>>(sc, psv, ctx->vw, qv, rcn, -1, ctx->d_stats);  // (sets the camera-differential flag)
            else k_vp_logic<VK_EXTEND><<<gl, kBlock, 0, st>>>(sc, psv, ctx->vw, qv, rcn, -1, ctx->d_stats);
            tm.end();
            launches += 2;
            // no DisneyMaterial in the scene: no BSDF has more than two lobes (the k_shade<2> / <8> split of the PathIntegrator)
            const bool fewLobes = !((ctx->shade_type_mask >> GNX_MAT_DISNEY) & 1u) && !getenv("GNX_VOL_MAXL8");
            auto logic = [&](int queue) {
                tm.begin(ST_VP_EXTEND + vol_queue_kernel(queue));
                k_vp_reset<<<1, 32, 0, st>>>(qv.counts, -2);  // (the cursor only)
                switch (vol_queue_kernel(queue)) {
                case VK_EXTEND: k_vp_logic<VK_EXTEND><<<ctx->grid_vp_kernel[VK_EXTEND], kBlock, 0, st>>>(sc, psv, ctx->vw, qv, rcn, queue, ctx->d_stats); break;
                case VK_VERTEX:
                    if (fewLobes) {
                        if (tex) k_vp_logic<VK_VERTEX, true, 2><<<ctx->grid_vp_kernel[VK_VERTEX], kBlock, 0, st>>>(sc, psv, ctx->vw, qv, rcn, queue, ctx->d_stats);
                        else k_vp_logic<VK_VERTEX, false, 2><<<ctx->grid_vp_kernel[VK_VERTEX], kBlock, 0, st>>>(sc, psv, ctx->vw, qv, rcn, queue, ctx->d_stats);
                    } else if (tex) k_vp_logic<VK_VERTEX, true><<<ctx->grid_vp_kernel[VK_VERTEX], kBlock, 0, st>>>(sc, psv, ctx->vw, qv, rcn, queue, ctx->d_stats);
                    else k_vp_logic<VK_VERTEX><<<ctx->grid_vp_kernel[VK_VERTEX], kBlock, 0, st>>>(sc, psv, ctx->vw, qv, rcn, queue, ctx->d_stats);
                    break;
                case VK_SHADOW: k_vp_logic<VK_SHADOW><<<ctx->grid_vp_kernel[VK_SHADOW], kBlock, 0, st>>>(sc, psv, ctx->vw, qv, rcn, queue, ctx->d_stats); break;
                default:
                    if (fewLobes) {
                        if (tex) k_vp_logic<VK_MIS, true, 2><<<ctx->grid_vp_kernel[VK_MIS], kBlock, 0, st>>>(sc, psv, ctx->vw, qv, rcn, queue, ctx->d_stats);
                        else k_vp_logic<VK_MIS, false, 2><<<ctx->grid_vp_kernel[VK_MIS], kBlock, 0, st>>>(sc, psv, ctx->vw, qv, rcn, queue, ctx->d_stats);
                    } else if (tex) k_vp_logic<VK_MIS, true><<<ctx->grid_vp_kernel[VK_MIS], kBlock, 0, st>>>(sc, psv, ctx->vw, qv, rcn, queue, ctx->d_stats);
                    else k_vp_logic<VK_MIS><<<ctx->grid_vp_kernel[VK_MIS], kBlock, 0, st>>>(sc, psv, ctx->vw, qv, rcn, queue, ctx->d_stats);
                    break;
                }
                k_vp_reset<<<1, 32, 0, st>>>(qv.counts, kCntShade0 + queue);
                tm.end();
                launches += 3;
            };
            for (int round = 0; round < 100000; ++round) {
                // one round: the queued tracking walks, then every logic kernel once, in the order a path flows through them
                tm.begin(ST_VP_TRACK);
                k_vp_reset<<<1, 32, 0, st>>>(qv.counts, -2);
                k_vp_track<<<gt, kBlock, 0, st>>>(sc, psv, ctx->vw, qv, rcn, ctx->d_stats);
                k_vp_reset<<<1, 32, 0, st>>>(qv.counts, kCntExtend0);
                tm.end();
                logic(VQ_VERTEX);
                logic(VQ_SHADOW_RESUME);
                logic(VQ_SHADOW);
                logic(VQ_MIS_RESUME);
                logic(VQ_MIS);
                logic(VQ_EXTEND);
                launches += 3;
                ++extendLaunches;
                ++vpRounds;
                if (round % 4 == 3 || round >= 16) {
                    int hc[kNumCounters];
                    GNX_CUDA(ctx, cudaMemcpyAsync(hc, qv.counts, sizeof(hc), cudaMemcpyDeviceToHost, st));
                    GNX_CUDA(ctx, cudaStreamSynchronize(st));
                    int remaining = hc[kCntExtend0];
                    for (int k = 0; k < VQ_COUNT; ++k) remaining += hc[kCntShade0 + k];
                    if (remaining == 0) break;
                }
            }
            tm.begin(ST_FILM);
            launches += accumulate(psv, rcn);
            tm.end();
            continue;
        }
        if (whittedStaged) {
            // camera rays -> first vertices (emitted light + one shadow item per light) -> any-hit -> sum in light order;
            // the samples whose vertex has specular lobes go through the recursion (gnx_whitted.cuh)
            const int nL = sc.n_lights, nSlots = rcn.npix * rcn.batch_spp;
            GNX_CUDA(ctx, cudaMemsetAsync(qv.counts, 0, sizeof(int) * kNumCounters, st));
            GNX_CUDA(ctx, cudaMemsetAsync(ctx->ww_planes, 0, (size_t)nL * ctx->capacity * sizeof(float4), st));
            PathState psw = psv;
            psw.La = psw.Lb = nullptr;
            Queues qw = qv;
            qw.shadow_q = ctx->ww_items;
            tm.begin(ST_EXTEND);
            k_trace<3><<<gridTrace, kBlock, 0, st>>>(sc, psw, qw, rcn, -1, ctx->d_stats);
            tm.end();
            if (qw.miss_q) k_escape<<<gridWide, 256, 0, st>>>(sc, psw, qw);
            tm.begin(ST_SHADE);
            for (int t = 0; t < kNumShadeTypes; ++t) {
                if (!((ctx->shade_type_mask >> t) & 1u)) continue;
                if (tex) k_whitted_vertex<true><<<ctx->grid_whitted_vertex, kBlock, 0, st>>>(sc, psw, qw, rcn, t);
                else k_whitted_vertex<false><<<ctx->grid_whitted_vertex, kBlock, 0, st>>>(sc, psw, qw, rcn, t);
                ++launches;
            }
            tm.end();
            tm.begin(ST_SHADOW);
            PathState psl = psw;
            psl.La = ctx->ww_planes;  // shadow_finish: planes[light * capacity + slot] = contribution when nothing is hit
            if (sc.nodes8 && sc.wide_any) k_anyhit8<0><<<ctx->grid_anyhit8, kBlock, 0, st>>>(sc, psl, qw, rcn, ctx->d_stats);
            else k_trace<1><<<gridTrace, kBlock, 0, st>>>(sc, psl, qw, rcn, 0, ctx->d_stats);
            k_whitted_sum<<<gridWide, 256, 0, st>>>(psw.L, ctx->ww_planes, nL, ctx->capacity, nSlots);
            tm.end();
            tm.begin(ST_EXTEND);
            k_zero_counter<<<1, 32, 0, st>>>(qw.counts, kCntFetch);
            RenderConsts rcl = rcn;
            rcl.rec_list = 1;  // extend queue 0
            if (tex) k_recursive<0, true><<<ctx->grid_recursive, kBlock, 0, st>>>(sc, psw, qw, rcl, ctx->d_stats);
            else k_recursive<0, false><<<ctx->grid_recursive, kBlock, 0, st>>>(sc, psw, qw, rcl, ctx->d_stats);
            tm.end();
            tm.begin(ST_FILM);
            launches += accumulate(psw, rcn);
            tm.end();
            launches += 8;
            ++extendLaunches;
            continue;
        }
        if (p->integrator != GNX_INTEGRATOR_PATH) {
            k_reset_counts<<<1, 32, 0, st>>>(qv.counts, 1);
            tm.begin(ST_EXTEND);
            // (scenes with image textures: the variants that carry ray differentials for MIPMap::Lookup)
            const int gr = ctx->grid_recursive;
            if (p->integrator == GNX_INTEGRATOR_VOLPATH) {
                if (tex) k_volpath<true><<<ctx->grid_volpath, kBlock, 0, st>>>(sc, psv, qv, rcn, ctx->d_stats);
                else k_volpath<false><<<ctx->grid_volpath, kBlock, 0, st>>>(sc, psv, qv, rcn, ctx->d_stats);
            } else if (p->integrator == GNX_INTEGRATOR_WHITTED) {
                if (tex) k_recursive<0, true><<<gr, kBlock, 0, st>>>(sc, psv, qv, rcn, ctx->d_stats);
                else k_recursive<0, false><<<gr, kBlock, 0, st>>>(sc, psv, qv, rcn, ctx->d_stats);
            } else if (p->integrator == GNX_INTEGRATOR_DIRECT) {
                if (tex) k_recursive<1, true><<<gr, kBlock, 0, st>>>(sc, psv, qv, rcn, ctx->d_stats);
                else k_recursive<1, false><<<gr, kBlock, 0, st>>>(sc, psv, qv, rcn, ctx->d_stats);
            } else {
                if (tex) k_recursive<2, true><<<gr, kBlock, 0, st>>>(sc, psv, qv, rcn, ctx->d_stats);
                else k_recursive<2, false><<<gr, kBlock, 0, st>>>(sc, psv, qv, rcn, ctx->d_stats);
            }
            tm.end();
            tm.begin(ST_FILM);
            launches += accumulate(psv, rcn);
            tm.end();
            launches += 2;
            ++extendLaunches;
            continue;
        }
        int in = 0;
        // Mixed launches: from the second bounce on, the extension rays of bounce d+1 and the shadow / environment-MIS
        // rays of bounce d go through ONE traversal launch (k_trace<4>): every launch of the persistent kernel ends in
        // a tail of a few long rays, and the late bounces are mostly tail.  Not with material-less surfaces (the loop
        // below polls the queue from the host) and not when the caller turned it off.
        const bool mixed = ctx->merge_extend && psv.La && psv.Lb && !hasNull && sc.n_lights > 0;
        // any-hit rays through the compressed 8-wide tree (both queue halves in one launch: needs the second accumulator)
        const bool close8 = sc.wide_closest && sc.nodes8 != nullptr;
        // (decided once per call, from what the previous call saw)
        const bool sortQueues = ctx->sort_queues && ctx->qs_bits &&
                                (ctx->qs_probe_host[1] == 0 || (long long)ctx->qs_probe_host[0] * kQsMinDensity >= ctx->qs_probe_host[1]);
        const bool any8 = sc.wide_any && sc.nodes8 != nullptr && ctx->merge_shadow && (psv.Lb || !sc.env.present);
        // bounces 0..maxDepth; surfaces without a material do not count as bounces, so scenes that
        // have them keep iterating until the queue drains.
        for (int iter = 0;; ++iter) {
            const int out = 1 - in;
            if (mixed && iter > 0) k_reset_counts_keep_rays<<<1, 32, 0, st>>>(qv.counts, out);
            else k_reset_counts<<<1, 32, 0, st>>>(qv.counts, out);
            if (ctx->rebin && iter > 0 && ctx->rb_tmp) {
                const int nPad = rcn.npix * rcn.batch_spp;
                k_rebin_keys<<<gridWide, 256, 0, st>>>(sc, psv, qv.extend_q[in], qv.counts + in, ctx->rb_keys[0], ctx->rb_vals[0], nPad);
                size_t bytes = ctx->rb_tmp_bytes;
                cub::DeviceRadixSort::SortPairs(ctx->rb_tmp, bytes, ctx->rb_keys[0], ctx->rb_keys[1], ctx->rb_vals[0], ctx->rb_vals[1], nPad, 0, 31, st);
                GNX_CUDA(ctx, cudaMemcpyAsync(qv.extend_q[in], ctx->rb_vals[1], (size_t)nPad * sizeof(int), cudaMemcpyDeviceToDevice, st));
            }
            tm.begin(ST_EXTEND);
            // Closest hits through the compressed 8-wide tree (k_trace<., true>); the rays it flags (two candidates within
            // the tie band: the reference's answer depends on its visiting order) are traced again in reference order by a
            // launch over the list the wide launch left in the free extend queue.
            auto extendRays = [&](int inq) {
                if (close8) {
                    k_trace<0, true><<<ctx->grid_trace8, kBlock, 0, st>>>(sc, psv, qv, rcn, inq, ctx->d_stats);
                    k_trace<0><<<gridTrace, kBlock, 0, st>>>(sc, psv, qv, rcn, (1 - inq) | 2, ctx->d_stats);
                    k_retrace_reset<<<1, 32, 0, st>>>(qv.counts, 1 - inq);
                    launches += 2;
                } else k_trace<0><<<gridTrace, kBlock, 0, st>>>(sc, psv, qv, rcn, inq, ctx->d_stats);
            };
            if (iter == 0) {  // ray-gen fused in
                if (close8) {
                    k_trace<3, true><<<ctx->grid_trace8p, kBlock, 0, st>>>(sc, psv, qv, rcn, -1, ctx->d_stats);
                    k_trace<3><<<gridTrace, kBlock, 0, st>>>(sc, psv, qv, rcn, 1, ctx->d_stats);
                    k_retrace_reset<<<1, 32, 0, st>>>(qv.counts, 1);
                    launches += 2;
                } else k_trace<3><<<gridTrace, kBlock, 0, st>>>(sc, psv, qv, rcn, -1, ctx->d_stats);
            } else if (mixed && any8) {
                // the any-hit rays of bounce d walk the compressed 8-wide tree in their own kernel, launched on a second
                // stream behind the extension launch of bounce d+1: its blocks move in as the extension kernel's blocks
                // drain (both are persistent grids), which fills the tail like the mixed launch did
                if (ctx->anyhit_overlap) {
                    GNX_CUDA(ctx, cudaEventRecord(ctx->ev_fork, st));
                    extendRays(in);
                    GNX_CUDA(ctx, cudaStreamWaitEvent(ctx->stream_any, ctx->ev_fork, 0));
                    k_anyhit8<1><<<ctx->grid_anyhit8, kBlock, 0, ctx->stream_any>>>(sc, psv, qv, rcn, ctx->d_stats);
                    GNX_CUDA(ctx, cudaEventRecord(ctx->ev_join, ctx->stream_any));
                    GNX_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_join, 0));
                } else {
                    extendRays(in);
                    k_anyhit8<1><<<ctx->grid_anyhit8, kBlock, 0, st>>>(sc, psv, qv, rcn, ctx->d_stats);
                }
                ++launches;
            } else if (mixed) k_trace<4><<<gridTrace, kBlock, 0, st>>>(sc, psv, qv, rcn, in, ctx->d_stats);
            else extendRays(in);
            tm.end();
            launches += 2;
            ++extendLaunches;
            if (qv.miss_q) { k_escape<<<gridWide, 256, 0, st>>>(sc, psv, qv); ++launches; }
            if (mixed && iter > 0) { k_reset_ray_counts<<<1, 32, 0, st>>>(qv.counts); ++launches; }
            if (ctx->sort_queues && iter == 0 && done == 0) {  // for the NEXT call's decision (no synchronisation)
                k_qs_probe<<<1, 32, 0, st>>>(qv.counts, ctx->qs_probe_dev);
                GNX_CUDA(ctx, cudaMemcpyAsync((void *)ctx->qs_probe_host, ctx->qs_probe_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
                ctx->qs_probe_host[1] = rcn.npix * rcn.batch_spp;
            }
            tm.begin(ST_SHADE);
            if (sortQueues) {
                // the shade queues back into slot order (the traversal kernel filled them in completion order)
                for (int t = 0; t < kNumShadeTypes; ++t)
                    if ((ctx->shade_type_mask >> t) & 1u) sortQueue(qv, t, rcn.npix * rcn.batch_spp);
            }
            for (int t = 0; t < kNumShadeTypes - 1; ++t) {
                if (!((ctx->shade_type_mask >> t) & 1u)) continue;
                if (ctx->has_next_lights) {  // point / spot / distant / skybox records: the variant with every Light::Sample_Li
                    if (t == GNX_MAT_DISNEY) k_shade<8, true><<<ctx->grid_shade8, kShadeBlock, 0, st>>>(sc, psv, qv, rcn, t, out);
                    else k_shade<2, true><<<gridShade, kShadeBlock, 0, st>>>(sc, psv, qv, rcn, t, out);
                } else if (t == GNX_MAT_DISNEY) k_shade<8><<<ctx->grid_shade8, kShadeBlock, 0, st>>>(sc, psv, qv, rcn, t, out);
                else k_shade<2><<<gridShade, kShadeBlock, 0, st>>>(sc, psv, qv, rcn, t, out);
                ++launches;
            }
            if (hasNull) { k_shade_null<<<gridShade, kBlock, 0, st>>>(sc, psv, qv, rcn, out); ++launches; }
            tm.end();
            tm.begin(ST_SHADOW);
            const bool last = iter >= p->max_depth && !hasNull;
            if (sc.n_lights > 0) {
                // shadow rays (A) and, with an environment light, its MIS rays (B) in one launch: they add to separate
                // accumulators (ps.L / ps.Lb), so the two rays of a path do not race.  In mixed mode they wait for the
                // next bounce's extension launch, except after the last bounce.
                if (!mixed || last) {
                    if (any8) k_anyhit8<0><<<ctx->grid_anyhit8, kBlock, 0, st>>>(sc, psv, qv, rcn, ctx->d_stats);
                    else k_trace<1><<<gridTrace, kBlock, 0, st>>>(sc, psv, qv, rcn, (psv.Lb && ctx->merge_shadow) ? 2 : 0, ctx->d_stats);
                    ++launches;
                    if (!any8 && sc.env.present && !(psv.Lb && ctx->merge_shadow)) { k_trace<1><<<gridTrace, kBlock, 0, st>>>(sc, psv, qv, rcn, 1, ctx->d_stats); ++launches; }
                }
                if (sc.n_lights > (sc.env.present ? 1 : 0)) { k_trace<2><<<gridTrace, kBlock, 0, st>>>(sc, psv, qv, rcn, 0, ctx->d_stats); ++launches; }
            }
            tm.end();
            in = out;
            if (iter >= p->max_depth) {
                if (!hasNull) break;
                int remaining = 0;
                GNX_CUDA(ctx, cudaMemcpyAsync(&remaining, qv.counts + in, sizeof(int), cudaMemcpyDeviceToHost, st));
                GNX_CUDA(ctx, cudaStreamSynchronize(st));
                if (remaining == 0 || iter > p->max_depth + 4096) break;
            }
        }
        tm.begin(ST_FILM);
        launches += accumulate(psv, rcn);
        tm.end();
    }
    const float norm = (float)(p->spp_normalize > 0 ? p->spp_normalize : p->spp);
    float4 *out = rgba_dev_out ? (float4 *)rgba_dev_out : ctx->rgba;
    if (gaussian) k_film_gauss<<<gridWide, 256, 0, st>>>(ctx->accum, out, npix, p->film == GNX_FILM_GAUSSIAN);
    else if (tiles) {
        RenderConsts rct{};
        rct.width = p->width; rct.height = p->height; rct.npix = npix;
        rct.tile_n = share.n; rct.tile_dev = share.idx; rct.tiles_x = tilesX; rct.tiles_y = tilesY;
        GNX_CUDA(ctx, cudaMemsetAsync(out, 0, (size_t)npixFrame * sizeof(float4), st));
        if (npix > 0) k_film_tiles<<<gridWide, 256, 0, st>>>(ctx->accum, out, rct, norm);
    } else k_film<<<gridWide, 256, 0, st>>>(ctx->accum, out, npix, norm, share.idx == 0 ? 1.f : 0.f);
    ++launches;
    GNX_CUDA(ctx, cudaEventRecord(ctx->ev1, st));
    GNX_CUDA(ctx, cudaGetLastError());
    if (stats || (!rgba_dev_out && !callerSyncs)) {
        GNX_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
        GNX_CUDA(ctx, cudaGetLastError());
    }
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        DevStats hs;
        GNX_CUDA(ctx, cudaMemcpy(&hs, ctx->d_stats, sizeof(hs), cudaMemcpyDeviceToHost));
        float ms = 0;
        GNX_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        stats->paths = hs.paths;
        stats->rays_extend = hs.rays[0];
        stats->rays_shadow = hs.rays[1];
        stats->rays_mis = hs.rays[2];
        stats->nodes_visited = hs.nodes[0] + hs.nodes[1] + hs.nodes[2];
        stats->tris_tested = hs.tris[0] + hs.tris[1] + hs.tris[2];
        stats->device_ms = ms;
        stats->kernel_launches = launches;
        const unsigned long long rays = hs.rays[0] + hs.rays[1] + hs.rays[2];
        // 32 B per node, 48 B per triangle, 32 B ray read + 16 B hit write per ray
        stats->bytes_algorithmic = 32ull * stats->nodes_visited + 48ull * stats->tris_tested + 48ull * rays;
        stats->extend_nodes = hs.nodes[0];
        stats->extend_tris = hs.tris[0];
        stats->extend_launches = extendLaunches;
        // (mixed launches trace the previous bounce's any-hit rays as well: their nodes / triangles are in nodes[0] / tris[0])
        // (likewise the k_anyhit8 launches that run next to an extension launch, inside the extend stage's timers)
        stats->extend_bytes = 32ull * (hs.nodes[0] + hs.any_nodes_in_extend) + 48ull * (hs.tris[0] + hs.any_tris_in_extend) +
                              48ull * (hs.rays[0] + hs.shadow_rays_in_extend_launches);
        double acc[ST_COUNT] = {};
        for (size_t i = 0; i * 2 + 1 < ctx->ev_used + 0 && i < ctx->ev_stage.size(); ++i) {
            float t = 0;
            if (cudaEventElapsedTime(&t, ctx->ev_pool[2 * i], ctx->ev_pool[2 * i + 1]) == cudaSuccess) acc[ctx->ev_stage[i]] += t;
        }
        // the VolPath wavefront's stages are also folded into the classic ones: extend, shade = vertex + MIS, shadow = shadow walk + tracking
        stats->ms_raygen = acc[ST_RAYGEN]; stats->ms_extend = acc[ST_EXTEND] + acc[ST_VP_EXTEND];
        stats->ms_shade = acc[ST_SHADE] + acc[ST_VP_VERTEX] + acc[ST_VP_MIS];
        stats->ms_shadow = acc[ST_SHADOW] + acc[ST_VP_SHADOW] + acc[ST_VP_TRACK]; stats->ms_film = acc[ST_FILM];
        for (int k = 0; k < 5; ++k) { stats->vp_ms[k] = acc[ST_VP_EXTEND + k]; stats->vp_items[k] = hs.vp_items[k]; }
        stats->vp_track_steps = hs.track_steps;
        stats->vp_rounds = vpRounds;
    }
    return GNX_OK;
}

static void add_stats(gnx_stats *a, const gnx_stats &b) {
    a->paths += b.paths; a->rays_extend += b.rays_extend; a->rays_shadow += b.rays_shadow; a->rays_mis += b.rays_mis;
    a->nodes_visited += b.nodes_visited; a->tris_tested += b.tris_tested; a->kernel_launches += b.kernel_launches;
    a->bytes_algorithmic += b.bytes_algorithmic; a->extend_nodes += b.extend_nodes; a->extend_tris += b.extend_tris;
    a->extend_launches = std::max(a->extend_launches, b.extend_launches); a->extend_bytes += b.extend_bytes;
    // times: the slowest device bounds the job
    a->device_ms = std::max(a->device_ms, b.device_ms); a->ms_raygen = std::max(a->ms_raygen, b.ms_raygen);
    a->ms_extend = std::max(a->ms_extend, b.ms_extend); a->ms_shade = std::max(a->ms_shade, b.ms_shade);
    a->ms_shadow = std::max(a->ms_shadow, b.ms_shadow); a->ms_film = std::max(a->ms_film, b.ms_film);
    for (int k = 0; k < 5; ++k) { a->vp_ms[k] = std::max(a->vp_ms[k], b.vp_ms[k]); a->vp_items[k] += b.vp_items[k]; }
    a->vp_track_steps += b.vp_track_steps; a->vp_rounds = std::max(a->vp_rounds, b.vp_rounds);
}

// One render on every GPU of a multi-device context (single process): the shares are queued from one host thread per
// device, the partial frames are summed onto the root — ncclReduce queued on every device's stream right behind its
// last film kernel, or one kernel on the root that loads the peers' frames over NVLink — and the root alone copies to
// the host.  Exactly one of host_out / dev_out (memory of the root device) is given.
static int render_multi(gnx_ctx *root, const gnx_render_params *p, float *host_out, float *dev_out, gnx_stats *stats, bool keepOnRoot = false) {
    if (!host_out && !dev_out && !keepOnRoot) return GNX_ERR_INVALID;
    int rc = validate_params(root, p);
    if (rc) return rc;
    const int G = 1 + (int)root->peers.size();
    const size_t npix = (size_t)p->width * p->height;
    // Gaussian film: the shares produce the unresolved sums (sum L f, sum f), the root divides after the reduce
    gnx_render_params pShare = *p;
    const bool resolveAfter = p->film == GNX_FILM_GAUSSIAN;
    if (resolveAfter) pShare.film = GNX_FILM_GAUSSIAN_SUMS;
    const gnx_render_params *pj = &pShare;
    std::vector<int> rcs(G, GNX_OK);
    std::vector<gnx_stats> sts(stats ? G : 0);
    auto ctxOf = [&](int g) { return g == 0 ? root : root->peers[g - 1]; };
    auto run = [&](int g) {
        gnx_ctx *c = ctxOf(g);
        Share sh;
        sh.n = G; sh.idx = g; sh.partition = p->partition;
        rcs[g] = render_impl(c, pj, g == 0 ? dev_out : nullptr, c->stream, stats ? &sts[g] : nullptr, true, sh);
        if (rcs[g] == GNX_OK && cudaEventRecord(c->ev_done, c->stream) != cudaSuccess) { c->err = "cudaEventRecord failed"; rcs[g] = GNX_ERR_CUDA; }
    };
    std::vector<std::thread> th;
    for (int g = 1; g < G; ++g) th.emplace_back(run, g);
    run(0);
    for (std::thread &t : th) t.join();
    for (int g = 0; g < G; ++g)
        if (rcs[g] != GNX_OK) {
            if (g > 0) root->err = "device " + std::to_string(ctxOf(g)->device) + ": " + ctxOf(g)->err;
            for (int k = 0; k < G; ++k) { cudaSetDevice(ctxOf(k)->device); cudaStreamSynchronize(ctxOf(k)->stream); }
            return rcs[g];
        }
    GNX_CUDA(root, cudaSetDevice(root->device));
    float4 *frame0 = dev_out ? (float4 *)dev_out : root->rgba;
    GNX_CUDA(root, cudaEventRecord(root->ev0, root->stream));
    NcclApi *nc = root->comms.empty() ? nullptr : nccl_api();
    if (nc) {
        int r = nc->GroupStart();
        for (int g = 0; g < G && r == 0; ++g) {
            gnx_ctx *c = ctxOf(g);
            const float4 *send = g == 0 ? frame0 : c->rgba;
            r = nc->Reduce(send, g == 0 ? (void *)frame0 : nullptr, npix * 4, kNcclFloat32, kNcclSum, 0, root->comms[g], c->stream);
        }
        int r2 = nc->GroupEnd();
        if (r == 0) r = r2;
        if (r != 0) return fail(root, GNX_ERR_CUDA, std::string("ncclReduce: ") + nc->GetErrorString(r));
        GNX_CUDA(root, cudaSetDevice(root->device));
    } else {
        PeerFrames pf{};
        pf.n = G - 1;
        if (!root->p2p_direct && root->stage_pixels < npix * (size_t)(G - 1)) {
            if (root->stage) cudaFree(root->stage);
            root->stage = nullptr; root->stage_pixels = 0;
            GNX_CUDA(root, cudaMalloc((void **)&root->stage, npix * (size_t)(G - 1) * sizeof(float4)));
            root->stage_pixels = npix * (size_t)(G - 1);
        }
        for (int g = 1; g < G; ++g) {
            gnx_ctx *c = ctxOf(g);
            GNX_CUDA(root, cudaStreamWaitEvent(root->stream, c->ev_done, 0));
            if (root->p2p_direct) pf.part[g - 1] = c->rgba;
            else {
                float4 *dst = root->stage + npix * (size_t)(g - 1);
                GNX_CUDA(root, cudaMemcpyPeerAsync(dst, root->device, c->rgba, c->device, npix * sizeof(float4), root->stream));
                pf.part[g - 1] = dst;
            }
        }
        k_reduce_peers<<<root->sm_count * 8, 256, 0, root->stream>>>(frame0, pf, (int)npix);
        GNX_CUDA(root, cudaGetLastError());
    }
    if (resolveAfter) {
        k_film_gauss<<<root->sm_count * 8, 256, 0, root->stream>>>(frame0, frame0, (int)npix, 1);
        GNX_CUDA(root, cudaGetLastError());
    }
    GNX_CUDA(root, cudaEventRecord(root->ev1, root->stream));
    if (host_out) GNX_CUDA(root, cudaMemcpyAsync(host_out, frame0, npix * sizeof(float4), cudaMemcpyDeviceToHost, root->stream));
    if (host_out || stats) GNX_CUDA(root, cudaStreamSynchronize(root->stream));
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        for (int g = 0; g < G; ++g) add_stats(stats, sts[g]);
        stats->paths = (uint64_t)npix * (uint64_t)p->spp;  // (tile shares count the padding of their edge tiles)
        float ms = 0;
        if (cudaEventElapsedTime(&ms, root->ev0, root->ev1) == cudaSuccess) { root->ms_reduce = ms; stats->device_ms += ms; stats->ms_film += ms; }
        stats->kernel_launches += 1;
    }
    return GNX_OK;
}

// The same job across processes (gnx_comm_attach): this rank renders its share on `st`, ncclReduce to rank 0 behind it.
static int render_attached(gnx_ctx *ctx, const gnx_render_params *p, float *host_out, float *dev_out, cudaStream_t st, gnx_stats *stats) {
    NcclApi *nc = nccl_api();
    if (!nc || ctx->comms.empty()) return fail(ctx, GNX_ERR_INVALID, "the context is not attached to a communicator");
    if (ctx->rank == 0 && !host_out && !dev_out) return GNX_ERR_INVALID;
    if (!p) return GNX_ERR_INVALID;
    Share sh;
    sh.n = ctx->n_ranks; sh.idx = ctx->rank; sh.partition = p->partition;
    gnx_render_params pShare = *p;
    const bool resolveAfter = p->film == GNX_FILM_GAUSSIAN;
    if (resolveAfter) pShare.film = GNX_FILM_GAUSSIAN_SUMS;
    int rc = render_impl(ctx, &pShare, dev_out, st, stats, true, sh);
    if (rc) return rc;
    const size_t npix = (size_t)p->width * p->height;
    float4 *frame = dev_out ? (float4 *)dev_out : ctx->rgba;
    int r = nc->Reduce(frame, ctx->rank == 0 ? (void *)frame : nullptr, npix * 4, kNcclFloat32, kNcclSum, 0, ctx->comms[0], st);
    if (r != 0) return fail(ctx, GNX_ERR_CUDA, std::string("ncclReduce: ") + nc->GetErrorString(r));
    if (resolveAfter && ctx->rank == 0) {
        k_film_gauss<<<ctx->sm_count * 8, 256, 0, st>>>(frame, frame, (int)npix, 1);
        GNX_CUDA(ctx, cudaGetLastError());
    }
    if (host_out && ctx->rank == 0) GNX_CUDA(ctx, cudaMemcpyAsync(host_out, frame, npix * sizeof(float4), cudaMemcpyDeviceToHost, st));
    if (host_out || stats) GNX_CUDA(ctx, cudaStreamSynchronize(st));
    return GNX_OK;
}

extern "C" {

double gnx_bvh_build_ms(const gnx_ctx *ctx) { return ctx ? (double)ctx->bvh_build_ms : 0.0; }

int gnx_render(gnx_ctx *ctx, const gnx_render_params *params, float *rgba_out, gnx_stats *stats) {
    if (!ctx) return GNX_ERR_INVALID;
    if (!ctx->peers.empty()) return render_multi(ctx, params, rgba_out, nullptr, stats);
    if (ctx->n_ranks > 1) return render_attached(ctx, params, rgba_out, nullptr, ctx->stream, stats);
    if (!rgba_out) return GNX_ERR_INVALID;
    // the device-to-host copy is queued right behind the film kernel; one synchronisation at the end
    int rc = render_impl(ctx, params, nullptr, ctx->stream, stats, true);
    if (rc) return rc;
    const size_t bytes = (size_t)params->width * params->height * sizeof(float4);
    GNX_CUDA(ctx, cudaMemcpyAsync(rgba_out, ctx->rgba, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    GNX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GNX_OK;
}

int gnx_render_device(gnx_ctx *ctx, const gnx_render_params *params, float *rgba_dev, void *stream, gnx_stats *stats) {
    if (!ctx) return GNX_ERR_INVALID;
    // stream 0 is the CALLER's default stream (cudaStreamLegacy), not a private one: work queued by the caller before and
    // after this call on that stream is ordered with the render, whatever the caller's stream flags are
    cudaStream_t st = stream ? (cudaStream_t)stream : cudaStreamLegacy;
    if (!ctx->peers.empty()) return rgba_dev ? render_multi(ctx, params, nullptr, rgba_dev, stats) : GNX_ERR_INVALID;
    if (ctx->n_ranks > 1) return (rgba_dev || ctx->rank != 0) ? render_attached(ctx, params, nullptr, rgba_dev, st, stats) : GNX_ERR_INVALID;
    if (!rgba_dev) return GNX_ERR_INVALID;
    return render_impl(ctx, params, rgba_dev, st, stats);
}

int gnx_render_framebuffer(gnx_ctx *ctx, const gnx_render_params *params, int32_t pass_count, float *fbuffer, uint8_t *ubuffer,
                           gnx_stats *stats) {
    if (!ctx || pass_count < 1) return GNX_ERR_INVALID;
    if (ctx->n_ranks > 1) return fail(ctx, GNX_ERR_UNSUPPORTED, "gnx_render_framebuffer on a context attached to a multi-process job");
    int rc = validate_params(ctx, params);
    if (rc) return rc;
    GNX_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t npix = (size_t)params->width * params->height;
    if (npix > ctx->fb_pixels) {
        if (ctx->fb_state) cudaFree(ctx->fb_state);
        if (ctx->fb_u8) cudaFree(ctx->fb_u8);
        ctx->fb_state = nullptr; ctx->fb_u8 = nullptr; ctx->fb_pixels = 0; ctx->fb_w = ctx->fb_h = 0;
        GNX_CUDA(ctx, cudaMalloc((void **)&ctx->fb_state, npix * sizeof(float4)));
        GNX_CUDA(ctx, cudaMalloc((void **)&ctx->fb_u8, npix * sizeof(uchar4)));
        ctx->fb_pixels = npix;
    }
    if (npix > ctx->fb_pinned_pixels) {
        if (ctx->fb_pinned) cudaFreeHost(ctx->fb_pinned);
        ctx->fb_pinned = nullptr; ctx->fb_pinned_pixels = 0;
        GNX_CUDA(ctx, cudaHostAlloc(&ctx->fb_pinned, npix * (sizeof(float4) + sizeof(uchar4)), cudaHostAllocDefault));
        ctx->fb_pinned_pixels = npix;
    }
    for (cudaEvent_t &e : ctx->fb_ev) if (!e) GNX_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    float4 *pinF = (float4 *)ctx->fb_pinned;
    uchar4 *pinU = (uchar4 *)((char *)ctx->fb_pinned + ctx->fb_pinned_pixels * sizeof(float4));
    cudaStream_t st = ctx->stream;
    const bool haveState = ctx->fb_w == params->width && ctx->fb_h == params->height;
    if (pass_count == 1 || (!haveState && !fbuffer)) {
        GNX_CUDA(ctx, cudaMemsetAsync(ctx->fb_state, 0, npix * sizeof(float4), st));  // (1 - 1/1) * f: the first pass forgets f
    } else if (!haveState) {
        // the running mean of passes this context has not seen: taken from the caller's FrameBuffer
        memcpy(pinF, fbuffer, npix * sizeof(float4));
        GNX_CUDA(ctx, cudaMemcpyAsync(ctx->fb_state, pinF, npix * sizeof(float4), cudaMemcpyHostToDevice, st));
    }
    ctx->fb_w = params->width; ctx->fb_h = params->height;
    // the render itself, result in ctx->rgba on this (the root) device
    if (!ctx->peers.empty()) rc = render_multi(ctx, params, nullptr, (float *)nullptr, stats, true);
    else rc = render_impl(ctx, params, nullptr, st, stats, true);
    if (rc) { ctx->fb_w = ctx->fb_h = 0; return rc; }
    k_framebuffer_update<<<ctx->sm_count * 8, 256, 0, st>>>(ctx->rgba, ctx->fb_state, ctx->fb_u8, (int)npix, 1.0f / (float)pass_count);
    GNX_CUDA(ctx, cudaGetLastError());
    // device -> pinned staging in row chunks, each chunk unpacked into the caller's (pageable) buffers while the next
    // one is in flight: colour channels only for the float image, all four bytes for the 8-bit one
    const int nChunks = 8;
    const size_t per = (npix + nChunks - 1) / nChunks;
    for (int c = 0; c < nChunks; ++c) {
        const size_t a = std::min(npix, per * c), b = std::min(npix, per * (c + 1));
        if (b > a) {
            if (fbuffer) GNX_CUDA(ctx, cudaMemcpyAsync(pinF + a, ctx->fb_state + a, (b - a) * sizeof(float4), cudaMemcpyDeviceToHost, st));
            if (ubuffer) GNX_CUDA(ctx, cudaMemcpyAsync(pinU + a, ctx->fb_u8 + a, (b - a) * sizeof(uchar4), cudaMemcpyDeviceToHost, st));
        }
        GNX_CUDA(ctx, cudaEventRecord(ctx->fb_ev[c], st));
    }
    // (four host threads share the chunks: one thread unpacks ~8 GB/s, the copies arrive at ~25 GB/s)
    const int nWorkers = npix >= (1u << 18) ? 4 : 1;
    std::vector<cudaError_t> werr(nWorkers, cudaSuccess);
    auto unpack = [&](int wk) {
        for (int c = wk; c < nChunks; c += nWorkers) {
            cudaError_t e = cudaEventSynchronize(ctx->fb_ev[c]);
            if (e != cudaSuccess) { werr[wk] = e; return; }
            const size_t a = std::min(npix, per * c), b = std::min(npix, per * (c + 1));
            if (fbuffer) {
                const float4 *src = pinF + a;
                float *dst = fbuffer + 4 * a;
                for (size_t i = 0; i < b - a; ++i) { dst[4 * i] = src[i].x; dst[4 * i + 1] = src[i].y; dst[4 * i + 2] = src[i].z; }
            }
            if (ubuffer && b > a) memcpy(ubuffer + 4 * a, pinU + a, (b - a) * sizeof(uchar4));
        }
    };
    std::vector<std::thread> workers;
    for (int wk = 1; wk < nWorkers; ++wk) workers.emplace_back(unpack, wk);
    unpack(0);
    for (std::thread &t : workers) t.join();
    for (cudaError_t e : werr) GNX_CUDA(ctx, e);
    return GNX_OK;
}

int gnx_num_devices(const gnx_ctx *ctx) { return ctx ? (ctx->peers.empty() ? ctx->n_ranks : 1 + (int)ctx->peers.size()) : 0; }

int gnx_create_multi(gnx_ctx **out, const int *device_ids, int n_devices) {
    if (!out) { g_create_error = "gnx_create_multi: out is NULL"; return GNX_ERR_INVALID; }
    *out = nullptr;
    if (n_devices < 1 || n_devices > kMaxDevices) { g_create_error = "gnx_create_multi: n_devices must be in [1, 16]"; return GNX_ERR_INVALID; }
    // A device may be listed more than once: its shares then time-share that GPU (tests of the partition logic on a
    // one-GPU box, oversubscription); NCCL cannot put one GPU into a communicator twice, so the reduce is the kernel then.
    std::vector<int> ids(n_devices);
    bool dup = false;
    for (int g = 0; g < n_devices; ++g) {
        ids[g] = device_ids ? device_ids[g] : g;
        for (int k = 0; k < g; ++k) dup = dup || ids[k] == ids[g];
    }
    gnx_ctx *root = nullptr;
    int rc = gnx_create(&root, ids[0]);
    if (rc) return rc;
    for (int g = 1; g < n_devices; ++g) {
        gnx_ctx *p = nullptr;
        if ((rc = gnx_create(&p, ids[g]))) { gnx_destroy(root); return rc; }
        root->peers.push_back(p);
    }
    if (n_devices > 1) {
        // peer-to-peer loads from every peer (NVLink / NVSwitch on a B200 box), for the one-kernel reduce
        cudaSetDevice(ids[0]);
        root->p2p_direct = true;
        for (int g = 1; g < n_devices; ++g) {
            int can = 0;
            if (ids[g] == ids[0]) continue;  // the root's own memory
            if (cudaDeviceCanAccessPeer(&can, ids[0], ids[g]) != cudaSuccess || !can) { root->p2p_direct = false; continue; }
            cudaError_t e = cudaDeviceEnablePeerAccess(ids[g], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) root->p2p_direct = false;
        }
        cudaGetLastError();
        if (dup && root->reduce_mode == 1) { g_create_error = "GNX_REDUCE=nccl with a device listed twice"; gnx_destroy(root); return GNX_ERR_INVALID; }
        NcclApi *nc = (root->reduce_mode == 2 || dup) ? nullptr : nccl_api();
        if (nc) {
            root->comms.assign(n_devices, nullptr);
            int r = nc->CommInitAll(root->comms.data(), n_devices, ids.data());
            if (r != 0) {
                std::string why = std::string("ncclCommInitAll: ") + nc->GetErrorString(r);
                root->comms.clear();
                if (root->reduce_mode == 1) { g_create_error = why; gnx_destroy(root); return GNX_ERR_CUDA; }
            }
            cudaSetDevice(ids[0]);
        } else if (root->reduce_mode == 1) {
            g_create_error = "GNX_REDUCE=nccl, but libnccl could not be loaded";
            gnx_destroy(root);
            return GNX_ERR_UNSUPPORTED;
        }
    }
    *out = root;
    return GNX_OK;
}

int gnx_comm_unique_id(void *id_out) {
    if (!id_out) return GNX_ERR_INVALID;
    NcclApi *nc = nccl_api();
    if (!nc) { g_create_error = "libnccl could not be loaded"; return GNX_ERR_UNSUPPORTED; }
    NcclId id;
    int r = nc->GetUniqueId(&id);
    if (r != 0) { g_create_error = std::string("ncclGetUniqueId: ") + nc->GetErrorString(r); return GNX_ERR_CUDA; }
    memcpy(id_out, &id, sizeof(id));
    return GNX_OK;
}

int gnx_comm_attach(gnx_ctx *ctx, int n_ranks, int rank, const void *id) {
    if (!ctx || !id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return GNX_ERR_INVALID;
    if (!ctx->peers.empty() || ctx->n_ranks > 1) return fail(ctx, GNX_ERR_INVALID, "the context already belongs to an N-device job");
    if (n_ranks == 1) return GNX_OK;
    NcclApi *nc = nccl_api();
    if (!nc) return fail(ctx, GNX_ERR_UNSUPPORTED, "libnccl could not be loaded");
    GNX_CUDA(ctx, cudaSetDevice(ctx->device));
    NcclId nid;
    memcpy(&nid, id, sizeof(nid));
    NcclComm comm = nullptr;
    int r = nc->CommInitRank(&comm, n_ranks, nid, rank);
    if (r != 0) return fail(ctx, GNX_ERR_CUDA, std::string("ncclCommInitRank: ") + nc->GetErrorString(r));
    ctx->comms.assign(1, comm);
    ctx->n_ranks = n_ranks;
    ctx->rank = rank;
    return GNX_OK;
}

int gnx_primary_hits(gnx_ctx *ctx, const gnx_render_params *p, int32_t sample, int32_t *prim_id_out) {
    if (!ctx || !prim_id_out) return GNX_ERR_INVALID;
    if (!ctx->has_scene) return fail(ctx, GNX_ERR_NO_SCENE, "no scene uploaded");
    if (!p || p->width <= 0 || p->height <= 0 || sample < 0) return fail(ctx, GNX_ERR_INVALID, "bad parameters");
    GNX_CUDA(ctx, cudaSetDevice(ctx->device));
    const int npix = p->width * p->height;
    int *d = nullptr;
    GNX_CUDA(ctx, cudaMalloc((void **)&d, (size_t)npix * sizeof(int)));
    k_primary_hits<<<ctx->sm_count * 4, kBlock, 0, ctx->stream>>>(ctx->sc, p->width, p->height, sample, d);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(prim_id_out, d, (size_t)npix * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail(ctx, GNX_ERR_CUDA, cudaGetErrorString(e));
    return GNX_OK;
}

int gnx_sample_dimensions(gnx_ctx *ctx, int32_t n, const int64_t *index, const int32_t *dim, float *out) {
    if (!ctx || n < 0 || !index || !dim || !out) return GNX_ERR_INVALID;
    if (!ctx->has_scene) return fail(ctx, GNX_ERR_NO_SCENE, "no scene uploaded");
    if (n == 0) return GNX_OK;
    GNX_CUDA(ctx, cudaSetDevice(ctx->device));
    long long *di = nullptr; int *dd = nullptr; float *dout = nullptr;
    cudaError_t e = cudaMalloc((void **)&di, (size_t)n * 8);
    if (e == cudaSuccess) e = cudaMalloc((void **)&dd, (size_t)n * 4);
    if (e == cudaSuccess) e = cudaMalloc((void **)&dout, (size_t)n * 4);
    if (e == cudaSuccess) e = cudaMemcpy(di, index, (size_t)n * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(dd, dim, (size_t)n * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        k_sample_dims<<<std::min((n + 255) / 256, ctx->sm_count * 8), 256, 0, ctx->stream>>>(ctx->sc, n, di, dd, dout);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpy(out, dout, (size_t)n * 4, cudaMemcpyDeviceToHost);
    cudaFree(di); cudaFree(dd); cudaFree(dout);
    if (e != cudaSuccess) return fail(ctx, GNX_ERR_CUDA, cudaGetErrorString(e));
    return GNX_OK;
}

int gnx_tonemap_rgba8(gnx_ctx *ctx, const float *rgba, int32_t n_pixels, uint8_t *rgba8_out) {
    if (!ctx || !rgba || !rgba8_out || n_pixels < 0) return GNX_ERR_INVALID;
    if (n_pixels == 0) return GNX_OK;
    GNX_CUDA(ctx, cudaSetDevice(ctx->device));
    float4 *din = nullptr; uchar4 *dout = nullptr;
    cudaError_t e = cudaMalloc((void **)&din, (size_t)n_pixels * 16);
    if (e == cudaSuccess) e = cudaMalloc((void **)&dout, (size_t)n_pixels * 4);
    if (e == cudaSuccess) e = cudaMemcpy(din, rgba, (size_t)n_pixels * 16, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        k_tonemap<<<std::min((n_pixels + 255) / 256, ctx->sm_count * 8), 256, 0, ctx->stream>>>(din, dout, n_pixels);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpy(rgba8_out, dout, (size_t)n_pixels * 4, cudaMemcpyDeviceToHost);
    cudaFree(din); cudaFree(dout);
    if (e != cudaSuccess) return fail(ctx, GNX_ERR_CUDA, cudaGetErrorString(e));
    return GNX_OK;
}

}  // extern "C"
