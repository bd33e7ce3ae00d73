// gnx_scene.cuh — the flattened scene as the kernels see it (device pointers, passed by value as a
// kernel parameter), plus the wavefront path state.  Layout in HBM is described in DESIGN.md §3.
#pragma once
#include "gnx_math.cuh"
#include "gnxrt.h"

namespace gnx {

constexpr int kMaxMipLevels = 16;
struct DevTexture {
    int w, h, nch, wrap;
    float su, sv, du, dv;
    const float *texels;  // level 0, row-major; the further MIPMap levels follow back to back (core/MIPMap.h:86-199)
    int n_levels, do_trilinear;
    float max_aniso;
    int level_off[kMaxMipLevels];  // offset of level l in texels, in TEXELS
};

struct DevEnv {
    int present, light_index;
    int w, h;                 // Lmap level 0
    const float4 *texels;     // [h][w] rgb0
    int dw, dh;               // Distribution2D resolution
    const float *cond_func, *cond_cdf, *cond_int, *marg_func, *marg_cdf;
    const uint16_t *cond_guide, *marg_guide;  // find_interval_guided tables: [dh][cond_g + 1], [marg_g + 1] (null: plain search)
    int cond_g, marg_g;
    float marg_int;
    M44 l2w, w2l;
    float world_radius;
};

struct DevSkybox {
    int present, light_index;
    int w, h, nc;
    const float *data;
    V3 center;
    float radius;
};

struct DevLightDistrib {
    int mode;                 // gnx_light_strategy actually in force
    const float *uni_func, *uni_cdf;   // [nL], [nL+1]  (uniform / power: one table for the whole scene)
    float uni_int;
    int nvox[3];              // spatial
    const float *sp_func, *sp_cdf, *sp_int;  // [nvoxels][nL], [nvoxels][nL+1], [nvoxels]
};

struct DevCamera {
    M44 r2c, c2w;
    float lens_radius, focal_distance;
    int medium;
    V3 dx_camera, dy_camera;  // PerspectiveCamera::dxCamera / dyCamera (camera/Perspective.cpp:26-32): ray differentials
};

struct DevSampler {
    int type;
    int base_scale0, base_scale1, base_exp0, base_exp1;
    int stride, mult_inv0, mult_inv1, at_center;
    int stride_over_scale0, stride_over_scale1;
    const uint16_t *perms;
    const int *primes;
    const uint4 *dims;        // per dimension {prime, PrimeSums, ceil(2^38 / prime) lo, hi}
    int n_primes;
    int spp;                  // Sampler::samplesPerPixel: ScaleDifferentials(1 / sqrt(spp)), core/Integrator.cpp:277
    // Sobol' (GNX_SAMPLER_SOBOL): generator matrices [dims][52] (the first kSobolConstDims also in __constant__ memory),
    // the two van der Corput rows of SobolIntervalToIndex for this resolution
    const uint32_t *sobol32;
    const uint64_t *sobol_vdc, *sobol_vdc_inv;
    int sobol_dims, sobol_log2res, sobol_res;
};

struct DevMedium {
    int type;
    float sigma_a[3], sigma_s[3], sigma_t[3];
    float g;
    int nx, ny, nz;
    const float *density;
    M44 w2m;
    float inv_max_density, sigma_t_scalar;
};

struct DeviceScene {
    // geometry: four float4 per interior node (gnx_bvh.cuh "Node2"); three float4 per ordered primitive:
    //   a = (p0.x p0.y p0.z p1.x)  b = (p1.y p1.z p2.x p2.y)  c = (p2.z, bits(material | type<<20 | flags<<24), bits(light), bits(prim_id))
    const float4 *nodes2;
    int n_nodes2;
    const uint4 *nodes8;          // compressed 8-wide tree over the same ordered primitives, any-hit queries (gnx_bvh8.cuh); null: none
    int n_nodes8;
    int wide_any, wide_closest;   // which queries go through nodes8 (the rest, and flagged closest-hit rays, walk nodes2)
    const float4 *tris;
    const float *tri_uv;          // [n][6] or null
    const float *tri_n;           // [n][9] or null
    const uint8_t *tri_has_n;     // [n] or null
    const int2 *tri_media;        // [n] (inside, outside) or null
    const uint8_t *tri_transition;
    int n_nodes, n_prims;
    float wb_min[3], wb_max[3];
    const gnx_material *materials;
    int n_materials;
    const DevTexture *textures;
    const float *ewa_lut;         // MIPMap::weightLut, 128 entries (core/MIPMap.h:189-196)
    const gnx_light *lights;
    int n_lights;
    const int *light_nsamples;    // [n_lights] Light::nSamples or null (= 1), UniformSampleAllLights
    DevEnv env;
    DevSkybox skybox;
    DevLightDistrib ld;
    const DevMedium *media;
    int n_media;
    DevCamera cam;
    DevSampler smp;
};

// ---- wavefront state over the path slots of one batch ------------------------------------------------------------
// One field of a per-slot record.  The queues name path slots in no particular order, so with one array per field every
// access of a lane touches its own half-used 32-byte sector (a float4) or an eighth-used one (a 4-byte field); with the
// fields of a slot side by side the sectors a kernel fetches are used in full.  A plain pointer converts to a field with
// the element's own stride (the host emulation works on one record at a time).
template <class T>
struct SlotField {
    char *base;
    int stride;
    SlotField() = default;
    GNX_HD SlotField(T *p) : base((char *)p), stride((int)sizeof(T)) {}
    GNX_HD SlotField(void *recordBase, int offset, int recordBytes) : base((char *)recordBase + offset), stride(recordBytes) {}
    GNX_HD T &operator[](size_t slot) const { return *(T *)(base + slot * (size_t)stride); }
    GNX_HD SlotField &operator-=(long n) { base -= n * (long)stride; return *this; }
};
// Device layout of a path record (96 bytes = three sectors): ray_o 0, ray_d 16, beta 32, hit 48, hidx 64, meta 68,
// medium 72.  The radiance accumulators stay arrays of their own: the film kernels read them densely, in slot order.
constexpr int kPathRecordBytes = 96;
struct PathState {
    SlotField<float4> ray_o;   // xyz origin, w = tMax
    SlotField<float4> ray_d;   // xyz direction, w = etaScale
    SlotField<float4> beta;    // xyz throughput, w unused
    float4 *L;         // xyz radiance accumulated by this path; w != 0: Lb[slot] is live as well
    SlotField<float4> hit;     // b0 b1 b2 bits(prim) written by extend
    SlotField<uint32_t> hidx;  // Halton sample index (low 32 bits; the reference's int64 never exceeds 2^32 at the configs)
    SlotField<uint32_t> meta;  // dimension (16 bits) | bounces (8) | flags (8)
    SlotField<int32_t> medium; // current ray medium (VolPath), -1 none
    float4 *Lb;        // contributions of the environment-MIS (shadow B) rays, kept apart from L so that the shadow A
                       // and shadow B rays of a path can be traced in the same launch without racing on one float4
                       // (null: they go to L, sequential callers)
    float4 *La;        // likewise for the shadow (A) rays, needed when they share a launch with the next bounce's
                       // extension rays, whose escape adds the environment radiance to L (null: they go to L)
    float4 *film_off;  // Gaussian film: the samples' film offsets, dense in slot order (k_film_prepare)
    void bind(void *records) {
        ray_o = SlotField<float4>(records, 0, kPathRecordBytes);
        ray_d = SlotField<float4>(records, 16, kPathRecordBytes);
        beta = SlotField<float4>(records, 32, kPathRecordBytes);
        hit = SlotField<float4>(records, 48, kPathRecordBytes);
        hidx = SlotField<uint32_t>(records, 64, kPathRecordBytes);
        meta = SlotField<uint32_t>(records, 68, kPathRecordBytes);
        medium = SlotField<int32_t>(records, 72, kPathRecordBytes);
    }
};
constexpr uint32_t kFlagSpecular = 1u;
constexpr uint32_t kFlagCameraDiff = 2u;  // VolPath: the path segment is still the camera's RayDifferential

struct ShadowItem {      // 48 bytes: any-hit query "add contrib to path if nothing is hit"
    float4 o_tmax;       // origin, tMax
    float4 d_path;       // direction, bits(path slot)
    float4 contrib;      // rgb, w unused
};
struct ProbeItem {       // closest-hit query "add contrib if the closest hit is primitive `expect`"
    float4 o_tmax;
    float4 d_path;
    float4 contrib_expect;  // rgb, bits(expected ordered primitive)
};

struct Queues {
    int *extend_q[2];        // ping-pong lists of path slots that need a closest-hit query
    int *shade_q;            // [n_mat_types][capacity] lists per material type
    ShadowItem *shadow_q;
    ProbeItem *probe_q;
    int *miss_q;             // escaped rays of scenes with a SkyBoxLight (k_escape); null otherwise
    // counters: [0..1] extend ping/pong, [2..2+NT) shade per type, then shadow, probe
    int *counts;
    int capacity;
};
constexpr int kNumShadeTypes = 7;  // 6 gnx_material_type + "no material" (medium boundary)
// queue counters: extend ping/pong, one per shade type, shadow A (light samples), shadow B (environment MIS
// probes), probe (area-light MIS probes); then the dynamic-fetch cursors of the four traversal launches
constexpr int kCntExtend0 = 0, kCntExtend1 = 1, kCntShade0 = 2, kCntShadow = kCntShade0 + kNumShadeTypes,
              kCntProbe = kCntShadow + 2, kCntFetch = kCntProbe + 1, kCntMiss = kCntFetch + 4, kNumCounters = kCntMiss + 1;

struct DevStats {  // index 0 = extension rays, 1 = shadow rays, 2 = MIS probe rays
    unsigned long long rays[3], nodes[3], tris[3], paths;
    unsigned long long shadow_rays_in_extend_launches;  // any-hit rays traced by the mixed launches (booked under nodes[0] / tris[0])
    unsigned long long any_nodes_in_extend, any_tris_in_extend;  // k_anyhit8 launched next to an extension launch: its share of nodes[1] / tris[1]
    unsigned long long track_steps;                     // VolPath: tracking steps taken by k_vp_track
    unsigned long long vp_items[5];                     // VolPath: items per stage (extend, vertex, shadow, MIS, track)
};

}  // namespace gnx
