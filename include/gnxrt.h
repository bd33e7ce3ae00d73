/*
 * gnxrt.h — C ABI of the B200-native path-tracing core (libgnxrt.so).
 *
 * Drop-in boundary: the library replaces the body of
 *     pbr::Integrator::Render(const Scene&, double& timeConsume)      core/Integrator.h:17-23
 * as implemented by
 *     SamplerIntegrator::Render                                        core/Integrator.cpp:225-319
 *     PathIntegrator::Li                                               integrators/PathIntegrator.cpp:62-208
 *     VolPathIntegrator::Li                                            integrators/VolPathIntegrator.cpp:24-159
 * of zhouxuguang/GNXRayTracer.  A reference-side `CUDAPathIntegrator : pbr::Integrator`
 * (gnxraytracer_b200/bridge/CUDAPathIntegrator.cpp) flattens the existing
 * pbr::Scene / Camera / Sampler into the plain structure-of-arrays buffers described here and
 * calls these entry points.  No C++ or torch types cross this boundary; no exception crosses it.
 *
 * Conventions
 *   - every function returns 0 on success and a negative gnx_status otherwise;
 *     gnx_last_error() gives a human-readable message for the last failure on that context
 *     (the reference has no error convention at all: void Render, core/Integrator.h:22).
 *   - all pointers in the descriptors are HOST pointers, copied during the call; the caller
 *     keeps ownership and may free them when the call returns.
 *   - there is NO CPU fallback: every compute entry point fails with GNX_ERR_NO_DEVICE when no
 *     CUDA device is usable.
 *   - Float == float and Spectrum == RGB (3 floats), as in the reference build
 *     (core/GNXRayTracer.h:82-86,113-117).
 */
#ifndef GNXRT_H
#define GNXRT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GNX_ABI_VERSION 4

typedef enum gnx_status {
    GNX_OK = 0,
    GNX_ERR_INVALID = -1,     /* bad argument / inconsistent descriptor            */
    GNX_ERR_NO_DEVICE = -2,   /* no usable CUDA device (no CPU fallback exists)     */
    GNX_ERR_CUDA = -3,        /* CUDA runtime error, message in gnx_last_error      */
    GNX_ERR_UNSUPPORTED = -4, /* scene feature outside the hot path */
    GNX_ERR_NO_SCENE = -5     /* render called before a successful upload           */
} gnx_status;

/* ------------------------------------------------------------------------------------------
 * Geometry.  Replaces BVHAccel::{nodes,primitives} (accelerator/BVHAccel.h:58-59).
 * ---------------------------------------------------------------------------------------- */

/* Bit-identical to LinearBVHNode (accelerator/BVHAccel.cpp:54-65): 32 bytes. */
typedef struct gnx_bvh_node {
    float bmin[3];
    float bmax[3];
    int32_t offset;    /* leaf: first ordered primitive; interior: index of second child */
    uint16_t n_prims;  /* 0 -> interior (first child is this index + 1)                  */
    uint8_t axis;      /* interior: split axis, picks the near child from the ray sign   */
    uint8_t pad;
} gnx_bvh_node;

/* Per-primitive flags (prim_flags) */
#define GNX_PRIM_FLIP_N 1u       /* reverseOrientation ^ transformSwapsHandedness (shape/Triangle.cpp:223-226) */
#define GNX_PRIM_REVERSE_ORI 2u  /* reverseOrientation alone (shape/Triangle.cpp:295)                         */

/* Triangles in BVH order (ordered primitive k = BVHAccel::primitives[k]); vertices are world
 * space as in TriangleMesh (shape/Triangle.cpp:24-29).  Optional arrays may be NULL. */
typedef struct gnx_geometry {
    int32_t n_nodes;
    const gnx_bvh_node *nodes;
    int32_t n_prims;
    const float *prim_p;            /* [n_prims][3 vertices][3]                                   */
    const float *prim_uv;           /* [n_prims][3][2] or NULL -> (0,0),(1,0),(1,1) (Triangle.h:60-74) */
    const float *prim_n;            /* [n_prims][3][3] shading normals, or NULL                   */
    const uint8_t *prim_has_n;      /* [n_prims] 1 when the primitive's mesh has normals, or NULL */
    const int32_t *prim_material;   /* [n_prims] material index, -1 = no material (medium boundary) */
    const int32_t *prim_light;      /* [n_prims] index into lights[] of its DiffuseAreaLight, -1 = none */
    const int32_t *prim_medium_in;  /* [n_prims] medium index inside, -1 = none; NULL = no media  */
    const int32_t *prim_medium_out; /* [n_prims] medium index outside                             */
    const uint8_t *prim_is_transition; /* [n_prims] MediumInterface::IsMediumTransition, or NULL  */
    const uint8_t *prim_flags;      /* [n_prims] GNX_PRIM_* or NULL                               */
    const int32_t *prim_id;         /* [n_prims] caller's own id of the primitive (parity hook)   */
    float world_bound[6];           /* Scene::WorldBound(): min xyz, max xyz (core/Scene.h:31)    */
} gnx_geometry;

/* ------------------------------------------------------------------------------------------
 * Materials and textures (materials/*.cpp, textures/*.h).
 * ---------------------------------------------------------------------------------------- */
typedef enum gnx_material_type {
    GNX_MAT_MATTE = 0,   /* rgb[0]=Kd  f[0]=sigma                        materials/MatteMaterial.cpp:14-32  */
    GNX_MAT_MIRROR = 1,  /* rgb[0]=Kr                                    materials/MirrorMaterial.cpp:13-23 */
    GNX_MAT_PLASTIC = 2, /* rgb[0]=Kd rgb[1]=Ks f[0]=roughness           materials/PlasticMaterial.cpp:15-41 */
    GNX_MAT_METAL = 3,   /* rgb[0]=eta rgb[1]=k f[0]=uRough f[1]=vRough  materials/MetalMaterial.cpp:28-49  */
    GNX_MAT_GLASS = 4,   /* rgb[0]=Kr rgb[1]=Kt f[0]=uRough f[1]=vRough f[2]=index  materials/GlassMaterial.cpp:14-59 */
    GNX_MAT_DISNEY = 5   /* rgb[0]=color rgb[1]=scatterDistance f[0..11]= metallic, eta, roughness,
                            specularTint, anisotropic, sheen, sheenTint, clearcoat, clearcoatGloss,
                            specTrans, flatness, diffTrans       materials/DisneyMaterial.cpp:467-581 */
} gnx_material_type;

#define GNX_MATF_REMAP_ROUGHNESS 1u /* remapRoughness                                             */
#define GNX_MATF_BUMP_IDENTITY 2u   /* a bump map is attached but constant: Material::Bump re-derives
                                       shading.n from dpdu x dpdv (core/Material.cpp:16-52)       */
#define GNX_MATF_THIN 4u            /* Disney thin surface                                        */

#define GNX_MAT_MAX_RGB 2
#define GNX_MAT_MAX_F 12

typedef struct gnx_material {
    int32_t type;
    uint32_t flags;
    float rgb[GNX_MAT_MAX_RGB][3];
    int32_t rgb_tex[GNX_MAT_MAX_RGB]; /* texture index or -1 = use the constant */
    float f[GNX_MAT_MAX_F];
    int32_t f_tex[GNX_MAT_MAX_F];
} gnx_material;

typedef enum gnx_wrap { GNX_WRAP_REPEAT = 0, GNX_WRAP_BLACK = 1, GNX_WRAP_CLAMP = 2 } gnx_wrap;

/* ImageTexture + UVMapping2D + MIPMap (textures/ImageTexture.h:42-87, core/Texture.cpp:165-175,
 * core/MIPMap.h).  Texels are the MIPMap's own pyramid levels, row-major (not BlockedArray order),
 * level 0 first.  PathIntegrator only ever reads level 0 bilinearly. */
typedef struct gnx_texture {
    int32_t width, height;   /* level 0 */
    int32_t n_channels;      /* 1 or 3  */
    int32_t n_levels;        /* >= 1    */
    int32_t wrap;            /* gnx_wrap */
    int32_t do_trilinear;
    float max_aniso;
    float su, sv, du, dv;    /* UVMapping2D */
    const float *texels;     /* all levels, back to back */
} gnx_texture;

/* ------------------------------------------------------------------------------------------
 * Lights (lights/*.cpp) and the light-sampling distribution (core/LightDistribution.cpp).
 * ---------------------------------------------------------------------------------------- */
typedef enum gnx_light_type {
    GNX_LIGHT_AREA_TRI = 0, /* DiffuseAreaLight on one Triangle   lights/DiffuseAreaLight.cpp:37-58 */
    GNX_LIGHT_INFINITE = 1, /* InfiniteAreaLight                  lights/InfiniteAreaLight.cpp:12-132 */
    GNX_LIGHT_POINT = 2,    /* PointLight    lights/PointLight.cpp:13-22      (WHITTED / DIRECT integrators) */
    GNX_LIGHT_SPOT = 3,     /* SpotLight     lights/SpotLight.cpp:22-43                                       */
    GNX_LIGHT_DISTANT = 4,  /* DistantLight  lights/DistantLight.cpp:16-27                                    */
    GNX_LIGHT_SKYBOX = 5    /* SkyBoxLight   lights/SkyBoxLight.cpp:45-86, image in gnx_scene_desc.skybox     */
} gnx_light_type;

typedef struct gnx_light {
    int32_t type;
    int32_t prim;        /* AREA_TRI: ordered primitive index of its triangle        */
    int32_t two_sided;   /* AREA_TRI: DiffuseAreaLight::twoSided                     */
    int32_t medium;      /* medium index the light sits in, -1 = none                */
    float L[3];          /* AREA_TRI: Lemit; POINT/SPOT: I; DISTANT: L               */
    float area;          /* AREA_TRI: Shape::Area(); DISTANT: worldRadius (Preprocess) */
    float p[3];          /* POINT/SPOT position; DISTANT direction wLight (world)    */
    float cos_total, cos_falloff; /* SPOT */
    float world_to_light[16];     /* SPOT: WorldToLight; SKYBOX: LightToWorld (row-major)   */
} gnx_light;

/* SkyBoxLight private state (lights/SkyBoxLight.h:44-48). */
typedef struct gnx_skybox {
    int32_t present;
    int32_t light_index;       /* its position in lights[]                                  */
    int32_t width, height, channels; /* stbi_loadf image, rows flipped as loaded; 0 = none  */
    const float *data;         /* [height][width][channels] or NULL (procedural colours)    */
    float center[3];           /* worldCenter                                               */
    float radius;              /* worldRadius                                               */
} gnx_skybox;

/* InfiniteAreaLight private state (lights/InfiniteAreaLight.h:39-42) flattened by value. */
typedef struct gnx_envmap {
    int32_t present;
    int32_t light_index;       /* its position in lights[]                                  */
    int32_t width, height;     /* Lmap level 0 (already resampled to a power of two)        */
    const float *texels;       /* [height][width][3]                                        */
    int32_t dist_w, dist_h;    /* Distribution2D resolution (2x the map)                    */
    const float *cond_func;    /* [dist_h][dist_w]      Distribution1D::func per row        */
    const float *cond_cdf;     /* [dist_h][dist_w + 1]  Distribution1D::cdf per row         */
    const float *cond_int;     /* [dist_h]              funcInt per row                     */
    const float *marg_func;    /* [dist_h]                                                  */
    const float *marg_cdf;     /* [dist_h + 1]                                              */
    float marg_int;
    float light_to_world[16];  /* row-major Matrix4x4 (core/Transform.h)                    */
    float world_to_light[16];
    float world_center[3];
    float world_radius;        /* InfiniteAreaLight::Preprocess (lights/InfiniteAreaLight.h:23-26) */
} gnx_envmap;

typedef enum gnx_light_strategy {
    GNX_LIGHTS_UNIFORM = 0, /* UniformLightDistribution, also used when exactly one light exists
                               (core/LightDistribution.cpp:15-33)                               */
    GNX_LIGHTS_SPATIAL = 1, /* SpatialLightDistribution: voxel grid, 128 Halton points per voxel */
    GNX_LIGHTS_POWER = 2    /* PowerLightDistribution: one Distribution1D over Light::Power().y()
                               (core/LightDistribution.cpp:44-50)                                */
} gnx_light_strategy;

/* ------------------------------------------------------------------------------------------
 * Media (media/*.cpp).
 * ---------------------------------------------------------------------------------------- */
typedef enum gnx_medium_type { GNX_MEDIUM_HOMOGENEOUS = 0, GNX_MEDIUM_GRID = 1 } gnx_medium_type;

typedef struct gnx_medium {
    int32_t type;
    float sigma_a[3], sigma_s[3];
    float g;
    int32_t nx, ny, nz;          /* GRID */
    const float *density;        /* GRID: [nz][ny][nx]                          */
    float world_to_medium[16];   /* GRID: row-major                             */
    float inv_max_density;       /* GRID: media/GridDensityMedium.h:19-42       */
} gnx_medium;

/* ------------------------------------------------------------------------------------------
 * Camera (camera/Perspective.cpp:35-112, core/Camera.h:54-75) and sampler
 * (samplers/HaltonSampler.cpp:33-94).
 * ---------------------------------------------------------------------------------------- */
typedef struct gnx_camera {
    float raster_to_camera[16]; /* row-major, applied with the homogeneous divide           */
    float camera_to_world[16];
    float lens_radius, focal_distance;
    float shutter_open, shutter_close;
    float dx_camera[3], dy_camera[3]; /* ray differentials (VolPath first vertex only)      */
    int32_t medium;                   /* camera medium index, -1 = none                     */
} gnx_camera;

typedef enum gnx_sampler_type {
    GNX_SAMPLER_HALTON = 0, /* HaltonSampler + GlobalSampler dimension bookkeeping            */
    GNX_SAMPLER_PCG32 = 1,  /* per-pixel PCG32 stream (core/RNG.h:30-110), for unbounded-dimension paths */
    GNX_SAMPLER_SOBOL = 2   /* (0,2)-sequence-in-the-first-two-dimensions Sobol' GlobalSampler built from the reference's
                               SobolIntervalToIndex / SobolSample helpers and generator matrices
                               (samplers/LowDiscrepancy.h:194-252, samplers/SobolMatrices.h:12-17); the reference ships
                               no sampler class for them, gnxraytracer_b200/bridge/SobolSampler.h is that class */
} gnx_sampler_type;

typedef struct gnx_sampler {
    int32_t type;
    int32_t samples_per_pixel;
    /* Halton state, exactly the HaltonSampler members (samplers/HaltonSampler.h:25-34) */
    int32_t base_scales[2], base_exponents[2];
    int32_t sample_stride;
    int32_t mult_inverse[2];
    int32_t sample_at_pixel_center;
    int32_t n_perm_entries;      /* length of perms                                             */
    const uint16_t *perms;       /* radicalInversePermutations, or NULL: the library generates the
                                    table itself from a default-seeded PCG32 like the reference
                                    (samplers/HaltonSampler.cpp:36-39)                          */
    /* Sobol' state (GNX_SAMPLER_SOBOL).  The generator matrices cannot be regenerated and are not copied into this
     * repository: the caller hands over the reference's own tables by pointer (the bridge has them linked in). */
    int32_t sobol_resolution;        /* RoundUpPow2(max(image width, height))                      */
    int32_t sobol_log2_resolution;
    int32_t n_sobol_dimensions;      /* NumSobolDimensions (1024)                                  */
    const uint32_t *sobol_matrices32; /* SobolMatrices32[n_sobol_dimensions][52]                   */
    const uint64_t *sobol_vdc;       /* VdCSobolMatrices[log2_resolution - 1][52]                  */
    const uint64_t *sobol_vdc_inv;   /* VdCSobolMatricesInv[log2_resolution - 1][52]               */
} gnx_sampler;

/* ------------------------------------------------------------------------------------------
 * Whole scene.
 * ---------------------------------------------------------------------------------------- */
typedef struct gnx_scene_desc {
    uint32_t abi_version; /* GNX_ABI_VERSION */
    gnx_geometry geom;
    int32_t n_materials;
    const gnx_material *materials;
    int32_t n_textures;
    const gnx_texture *textures;
    int32_t n_lights;
    const gnx_light *lights;
    gnx_envmap env;
    int32_t n_media;
    const gnx_medium *media;
    gnx_camera camera;
    gnx_sampler sampler;
    gnx_skybox skybox;
    const float *light_power; /* [n_lights] Light::Power().y() of every light, the input of
                                 ComputeLightPowerDistribution (core/Integrator.cpp:212-220).  Only read for
                                 GNX_LIGHTS_POWER; NULL = the library derives it for area / point / spot /
                                 distant lights and refuses power sampling of an environment / skybox light */
    const int32_t *light_n_samples; /* [n_lights] Light::nSamples (core/Light.h), read by GNX_INTEGRATOR_DIRECT_ALL only;
                                       NULL = 1 sample per light */
} gnx_scene_desc;

typedef enum gnx_integrator {
    GNX_INTEGRATOR_PATH = 0,     /* integrators/PathIntegrator.cpp                                          */
    GNX_INTEGRATOR_VOLPATH = 1,  /* integrators/VolPathIntegrator.cpp                                       */
    GNX_INTEGRATOR_WHITTED = 2,  /* integrators/WhittedIntegrator.cpp (the UI's default, ui/RenderThread.cpp:163) */
    GNX_INTEGRATOR_DIRECT = 3,   /* integrators/DirectLightingIntegrator.cpp, LightStrategy::UniformSampleOne */
    GNX_INTEGRATOR_DIRECT_ALL = 4 /* the same with LightStrategy::UniformSampleAll: every light at every vertex, with
                                    light_n_samples[j] samples each out of the sampler's 2-D sample arrays
                                    (DirectLightingIntegrator.cpp:13-27, core/Integrator.cpp:25-55)              */
} gnx_integrator;
typedef enum gnx_film {
    GNX_FILM_BOX = 0,           /* the reference: mean of the pixel's own samples (core/Integrator.cpp:274-293)      */
    GNX_FILM_GAUSSIAN = 1,      /* GaussianFilter (filters/GaussianFilter.h:12-33) with Film::AddSample semantics:
                                   rgb = max(0, sum(L f) / sum(f)) over the samples within the filter radius, alpha 1 */
    GNX_FILM_GAUSSIAN_SUMS = 2  /* the same sums unresolved: rgba = (sum(L f), sum(f)); an N-GPU job adds the ranks'
                                   buffers and divides afterwards (spp_normalize is not used by either Gaussian film) */
} gnx_film;

typedef struct gnx_render_params {
    int32_t width, height;      /* pixelBounds.pMax (core/Integrator.cpp:257-259)                */
    int32_t spp;                /* samples taken per pixel by THIS call                           */
    int32_t first_sample;       /* index of the first of them (Sampler::SetSampleNumber); a rank
                                   of an N-GPU job renders [first_sample, first_sample + spp)     */
    int32_t spp_normalize;      /* divisor of the pixel sum; 0 -> spp.  N-GPU jobs pass the total */
    int32_t max_depth;          /* PathIntegrator::maxDepth                                       */
    float rr_threshold;         /* PathIntegrator::rrThreshold (default 1)                        */
    int32_t integrator;         /* gnx_integrator                                                 */
    int32_t light_strategy;     /* gnx_light_strategy                                             */
    int32_t film;               /* gnx_film; BOX == the reference (core/Integrator.cpp:293)       */
    float filter_radius, filter_alpha; /* GAUSSIAN films: radius in pixels (both axes), falloff alpha  */
    int32_t batch_spp;          /* samples per pixel in flight per wavefront batch; 0 = auto      */
    int32_t partition;          /* gnx_partition: how an N-device job (gnx_create_multi / gnx_comm_attach) is dealt
                                   out; ignored by a single-device context                         */
} gnx_render_params;

/* Partition of ONE render over the GPUs of a multi-device context.  params describe the WHOLE job
 * (spp = all samples of a pixel); the library deals out the work and sums the partial framebuffers onto the root. */
typedef enum gnx_partition {
    GNX_PARTITION_SAMPLES = 0, /* device g renders a contiguous range of every pixel's samples (Sampler::SetSampleNumber
                                  semantics, core/Sampler.cpp:155-160); perfect balance; float sum-reduce (differs from the
                                  single-device sum by the rounding of the regrouped additions)                            */
    GNX_PARTITION_TILES = 1    /* interleaved 32 x 32 pixel tiles, every pixel finished on one device; the reduce only adds
                                  zeros: bit-equal to the single-device render.  Box film only                             */
} gnx_partition;

typedef struct gnx_stats {
    uint64_t paths;             /* camera samples started                                         */
    uint64_t rays_extend;       /* closest-hit queries for path continuation (incl. primary)      */
    uint64_t rays_shadow;       /* any-hit queries (VisibilityTester::Unoccluded)                 */
    uint64_t rays_mis;          /* closest-hit queries of the BSDF-sampling half of EstimateDirect */
    uint64_t nodes_visited;     /* BVH nodes popped, all ray kinds                                */
    uint64_t tris_tested;       /* triangles tested, all ray kinds                                */
    double device_ms;           /* CUDA-event time of the whole render (ray-gen .. film)          */
    double ms_raygen, ms_extend, ms_shade, ms_shadow, ms_film; /* per-stage CUDA-event sums       */
    uint64_t kernel_launches;   /* kernels launched by this call                                  */
    uint64_t bytes_algorithmic; /* 32*nodes + 48*tris + 48*rays over all traversal kernels */
    /* the dominant kernel (k_extend) on its own, for the roofline line of bench.py */
    uint64_t extend_nodes, extend_tris; /* BVH nodes popped / triangles tested by closest-hit extension rays */
    uint64_t extend_launches;   /* number of k_extend launches; ms_extend is their CUDA-event sum   */
    uint64_t extend_bytes;      /* 32*extend_nodes + 48*extend_tris + 48*rays_extend                */
    /* VolPath wavefront (GNX_INTEGRATOR_VOLPATH), per stage: [0] extend kernel, [1] vertex kernel, [2] shadow-walk kernel,
     * [3] MIS-walk + continuation kernel, [4] tracking kernel.  items = paths advanced (walks for [4]); vp_track_steps =
     * tracking steps (8 grid reads each); vp_rounds = rounds of { track, logic kernels } until every path had ended */
    double vp_ms[5];
    uint64_t vp_items[5];
    uint64_t vp_track_steps;
    uint64_t vp_rounds;
} gnx_stats;

typedef struct gnx_ctx gnx_ctx;

/* Lifecycle -------------------------------------------------------------------------------- */
int gnx_abi_version(void);
int gnx_device_count(void);                              /* 0 when no CUDA device is usable   */
int gnx_create(gnx_ctx **out, int device);               /* one context == one GPU            */
/* One context driving n_devices GPUs of this node from ONE process (what a reference app calling
 * integrator->Render(scene, t) needs to use a whole 8-GPU box, ui/RenderThread.cpp:169-175): the scene is replicated by
 * gnx_upload_scene, gnx_render splits the job (gnx_render_params::partition) over one stream per device, the partial
 * framebuffers are summed onto device_ids[0] — ncclReduce over NVLink queued behind each device's last film kernel
 * (libnccl is loaded at run time), or one peer-to-peer kernel on the root (GNX_REDUCE=p2p, and when NCCL is not
 * available) — and only the root copies to the host.  device_ids == NULL: devices 0 .. n_devices-1. */
int gnx_create_multi(gnx_ctx **out, const int *device_ids, int n_devices);
int gnx_num_devices(const gnx_ctx *ctx);
/* The same job partition across PROCESSES (one rank per GPU, e.g. under torchrun): rank 0 obtains an id, the host
 * program distributes its 128 bytes to every rank (any transport), each rank attaches its single-device context.
 * From then on gnx_render / gnx_render_device on these contexts are collective: params describe the whole job on every
 * rank, each rank renders its share, ncclReduce leaves the image on rank 0 (rgba_out may be NULL elsewhere). */
#define GNX_COMM_ID_BYTES 128
int gnx_comm_unique_id(void *id_out);
int gnx_comm_attach(gnx_ctx *ctx, int n_ranks, int rank, const void *id);
void gnx_destroy(gnx_ctx *ctx);
const char *gnx_last_error(const gnx_ctx *ctx);          /* ctx may be NULL: last create error */

/* Scene ------------------------------------------------------------------------------------ */
/* scene->geom.n_nodes == 0 (nodes NULL): the library builds the BVH itself, on the GPU (linear BVH, SURVEY 8f
 * rank 2; replaces BVHAccel's host build, accelerator/BVHAccel.cpp:147-189), and re-orders the per-primitive
 * arrays to match; prim_id keeps the caller's numbering.  Otherwise the node array is the reference's own
 * LinearBVHNode layout and the reference's traversal order is reproduced exactly. */
int gnx_upload_scene(gnx_ctx *ctx, const gnx_scene_desc *scene);
double gnx_bvh_build_ms(const gnx_ctx *ctx); /* device time of that build; 0 when the caller supplied nodes */

/* Render: replaces SamplerIntegrator::Render.  rgba_out is a HOST buffer of width*height*4
 * floats receiving the per-pixel mean radiance in (x + y*width)*4 + c order with alpha 1, i.e.
 * what update_f_u_c stores on the first pass (ui/FrameBuffer.h:136-139).  stats may be NULL. */
int gnx_render(gnx_ctx *ctx, const gnx_render_params *params, float *rgba_out, gnx_stats *stats);

/* Same, but the result stays in device memory owned by the caller (a CUDA device pointer to
 * width*height*4 floats, e.g. a torch tensor) so a multi-GPU host can reduce it with NCCL
 * without a host round trip.  stream is the cudaStream_t the render is queued on; 0 means the caller's legacy default
 * stream (cudaStreamLegacy).  The call returns once the work is QUEUED (unless stats are requested, which waits for it):
 * consumers must be ordered behind it on the same stream, or synchronise that stream. */
int gnx_render_device(gnx_ctx *ctx, const gnx_render_params *params, float *rgba_dev,
                      void *stream, gnx_stats *stats);

/* Render + the reference's output sink in one call: what SamplerIntegrator::Render does to the FrameBuffer it was
 * given (core/Integrator.cpp:230,293-310, ui/FrameBuffer.h:127-149), with the running mean over Render() calls and the
 * exposure tonemap computed on the device behind the film kernel:
 *     f   = (1 / pass_count) * mean + (1 - 1 / pass_count) * f          per colour channel (update_f_u_c)
 *     u8  = (1 - exp(-f / (1 - 0.75))) * 255, alpha byte 255              (update_f_u_c, set_uc)
 * fbuffer is the HOST float[width*height*4] of the FrameBuffer (only its three colour channels are written, the float
 * alpha stays the caller's), ubuffer the HOST uint8[width*height*4]; either may be NULL.  pass_count is
 * FrameBuffer::curRenderCount after renderCountIncrease(), i.e. 1 on the first pass.  The running mean lives on the
 * device between calls (the library owns pinned staging for the copies); a call with pass_count > 1 on a context that
 * has no state of that size yet reads it from fbuffer first. */
int gnx_render_framebuffer(gnx_ctx *ctx, const gnx_render_params *params, int32_t pass_count, float *fbuffer,
                           uint8_t *ubuffer, gnx_stats *stats);

/* Parity hook: id (gnx_geometry::prim_id) of the primitive hit by the camera ray of sample
 * `sample` of every pixel, -1 for a miss; prim_id_out is HOST int32[width*height]. */
int gnx_primary_hits(gnx_ctx *ctx, const gnx_render_params *params, int32_t sample,
                     int32_t *prim_id_out);

/* Parity hook: the sampler value SampleDimension(index, dim) for n (index, dim) pairs. */
int gnx_sample_dimensions(gnx_ctx *ctx, int32_t n, const int64_t *index, const int32_t *dim,
                          float *out);

/* Tonemap + 8-bit pack identical to FrameBuffer::update_f_u_c (ui/FrameBuffer.h:141-147):
 * u8 = (1 - exp(-x / (1 - 0.75))) * 255, alpha 255.  Host buffers (a utility for images that are already on the host;
 * the render path tonemaps on the device, gnx_render_framebuffer). */
int gnx_tonemap_rgba8(gnx_ctx *ctx, const float *rgba, int32_t n_pixels, uint8_t *rgba8_out);

#ifdef __cplusplus
}
#endif
#endif /* GNXRT_H */
