/*
 * gnx_scenekit.h — C ABI of the host-side scene kit (libgnxscenekit.so).
 *
 * The reference has no scene file format: its scenes are built in code by the Qt UI
 * (ui/RenderThread.cpp:46-164, ui/ModelList.cpp:20-178, ui/MaterialList.cpp:31-92) and its BVH by
 * BVHAccel's constructor (accelerator/BVHAccel.cpp:147-189).  When the host application is the
 * reference itself, the bridge (gnxraytracer_b200/bridge) flattens that live scene.  The scene kit is
 * the stand-alone counterpart: it builds the same BASELINE.json configs — geometry from
 * host/scenekit_mesh.h, its own binned-SAH BVH, the InfiniteAreaLight importance tables
 * (lights/InfiniteAreaLight.cpp:12-82), Halton parameters (samplers/HaltonSampler.cpp:33-61) and the
 * perspective camera (camera/Perspective.cpp:114-135) — directly as a gnx_scene_desc, so that
 * bench.py and smoke() need neither the reference nor the oracle.
 */
#ifndef GNX_SCENEKIT_H
#define GNX_SCENEKIT_H

#include "gnxrt.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gnxsk_scene gnxsk_scene;

/* name: "cornell" (p0: 0 = Lambert walls, 1 = Oren-Nayar sigma 60; p1: icosphere subdivision, -1 = none)
 *       "dragon"  (p0: 0 = Plastic, 1 = Metal; p1 x p2: torus-knot quads, 0 = 2048 x 213)
 *       "dragon3d:<path>"  the same scene with the mesh read from a .3d file — the reference's mesh format
 *                 (shape/plyRead.h:19-48) — and placed as ui/ModelList.cpp:49-69 places dragon.3d (x20, y - 2.9)
 *       "obj:<path>"       the same scene with a Wavefront OBJ mesh (uv / normals kept when every corner has them),
 *                 fitted into a sphere of radius 2.5 around (0, -0.4, 0)
 *       "nano"    (p0: 0 = Disney, 1 = thin Disney; p1 x p2: knot quads, 0 = 320 x 64; textured, smooth-shaded)
 *       "smoke"   (p0: 0 = grid density medium in fog, PCG32 stream sampler; 1 = fog only, Halton) -> render with
 *                 GNX_INTEGRATOR_VOLPATH
 *       "lights"  (p0: light mask, bit 0 area, 1 point, 2 spot, 3 distant, 4 skybox, 0 = all; p1: sphere subdivision)
 *                 -> render with GNX_INTEGRATOR_WHITTED or GNX_INTEGRATOR_DIRECT
 * resources: directory holding MonValley1000.hdr (only read by "dragon").
 * Never returns NULL; check gnxsk_error(). */
gnxsk_scene *gnxsk_create(const char *name, int width, int height, int spp, int p0, int p1, int p2,
                          const char *resources);
void gnxsk_destroy(gnxsk_scene *s);
const char *gnxsk_error(const gnxsk_scene *s);         /* "" when the scene is usable */
const gnx_scene_desc *gnxsk_desc(const gnxsk_scene *s); /* valid until gnxsk_destroy   */
int gnxsk_num_prims(const gnxsk_scene *s);
double gnxsk_build_seconds(const gnxsk_scene *s);      /* BVH build wall time          */
/* Drops the kit's host-built BVH from the description (geom.n_nodes = 0): gnx_upload_scene then builds the
 * hierarchy on the GPU.  The primitive arrays stay as they are (any order is valid without nodes). */
void gnxsk_strip_bvh(gnxsk_scene *s);
/* Damages the description in one specific way so that tests can check that gnx_upload_scene refuses it with
 * GNX_ERR_INVALID instead of reading out of range on the device: kind 1 = a BVH leaf whose primitive range leaves the
 * array, 2 = prim_light past the light list, 3 = a texture index below -1, 4 = materials NULL with a non-zero count,
 * 5 = prim_light pointing at a light that is not an area light.  Returns 0 when applied. */
int gnxsk_corrupt(gnxsk_scene *s, int kind);

/* Mesh files on their own.  Both return 0, or -1 with the reason copied into err (NUL-terminated, err_len bytes).
 * gnxsk_mesh_info parses a .3d file (or a Wavefront OBJ when the path ends in ".obj") and reports its counts;
 * gnxsk_write_knot_3d writes the kit's nu x nv torus knot in the .3d layout (tests, tools). */
int gnxsk_mesh_info(const char *path, int *n_vertices, int *n_triangles, int *has_uv, int *has_normals, char *err, int err_len);
int gnxsk_write_knot_3d(const char *path, int nu, int nv, char *err, int err_len);

#ifdef __cplusplus
}
#endif
#endif
