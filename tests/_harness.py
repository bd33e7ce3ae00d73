"""ctypes wrappers used by the tests only: the oracle (compiled reference + harness) and the CPU
emulation of the product's device functions."""
import ctypes
import os

import numpy as np

from gnxraytracer_b200.api import RenderParams, Stats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libgnxref.so")        # the unmodified reference + scene harness
BRIDGE_LIB = os.path.join(ROOT, "oracle", "_ref", "libgnxbridge.so")  # the product's drop-in class on those scenes
EMUL_LIB = os.path.join(ROOT, "tests", "emul", "_build", "libgnxemul.so")
vp, ci = ctypes.c_void_p, ctypes.c_int

# (scene name, p0, p1, p2) presets small enough for CPU tests
SCENES = {
    "cornell": ("cornell", 0, 2, 0),        # Lambert walls, mirror + glass icospheres (320 tris each)
    "cornell_on": ("cornell", 1, 2, 0),     # the UI's Oren-Nayar sigma = 60 walls
    "cornell_2l": ("cornell", 2, 2, 0),     # + a second, smaller emitter of another colour (unequal light powers)
    "cornell_full": ("cornell", 0, 3, 0),   # BASELINE config 1 geometry (2 x 1280-triangle spheres)
    "dragon": ("dragon", 0, 256, 32),       # Plastic knot, MonValley environment
    "dragon_metal": ("dragon", 1, 256, 32),
    "dragon_full": ("dragon", 0, 2048, 213),  # BASELINE config 2 geometry (872 448 triangles)
    "nano": ("nano", 0, 96, 24),            # config 3 stand-in: Disney + ImageTexture + shading normals, small
    "nano_thin": ("nano", 1, 96, 24),       # thin Disney surface: transmission lobes
    "nano_full": ("nano", 0, 320, 64),      # ~ 90 k triangles
    "smoke": ("smoke", 0, 0, 0),            # config 4: VolPath, grid density medium in homogeneous fog, PCG stream sampler
    "fog": ("smoke", 1, 0, 0),              # VolPath, homogeneous fog only, Halton sampler
    # Whitted / DirectLighting with area + point + spot + distant + skybox lights (p0 = light mask)
    "whitted": ("whitted", 31, 2, 0),
    "whitted_img": ("whitted", 1 | 2 | 32, 2, 0),   # skybox with the awesomeface.jpg image
    "direct": ("direct", 31, 2, 0),
    "direct_area": ("direct", 1, 2, 0),             # area light only: both MIS halves of EstimateDirect
    # LightStrategy::UniformSampleAll: every light at every vertex, samples out of the sampler's 2-D arrays
    "direct_all": ("direct", 31, 2, 4),
    "direct_all_area": ("direct", 1, 2, 4),         # two area-light triangles with nSamples = 5 each
    # the reference UI's live scene (ui/RenderThread.cpp:60-164): mesh inside the Cornell box, area light + SkyBoxLight;
    # p0 = gnx_integrator (0: the commented-in PathIntegrator line :164, 2: the default WhittedIntegrator :163)
    # image textures under the integrators that carry ray differentials: EWA (p2 = 1) / trilinear (p2 = 2) MIPMap::Lookup
    "whitted_tex": ("whitted_tex", 31, 2, 1),
    "whitted_tri": ("whitted_tex", 31, 2, 2),
    "direct_tex": ("direct_tex", 1 | 2, 2, 1),
    "fog_tex": ("smoke", 2, 0, 0),                  # VolPath: textured floor seen directly by the camera, EWA
    "fog_tri": ("smoke", 3, 0, 0),                  # ... trilinear
    "lights_path": ("lights_path", 31, 2, 0),       # PathIntegrator with area + point + spot + distant + skybox lights
    "lights_path_img": ("lights_path", 1 | 2 | 32, 2, 0),
    "ui_path": ("ui", 0, 256, 32),
    "ui_whitted": ("ui", 2, 256, 32),
    "ui_path_full": ("ui", 0, 2048, 213),
    "ui_whitted_full": ("ui", 2, 2048, 213),
}
INTEGRATOR_OF = {"whitted": 2, "direct": 3, "smoke": 1, "whitted_tex": 2, "direct_tex": 3}


def integrator_of(preset):
    """gnx_integrator of a preset (p2 = 4 selects DirectLightingIntegrator with UniformSampleAll)."""
    name, p0, _, p2 = SCENES[preset]
    if name == "ui":
        return p0
    return 4 if name == "direct" and p2 == 4 else INTEGRATOR_OF.get(name, 0)



def grid(width, height):
    ys, xs = np.mgrid[0:height, 0:width]
    return xs.ravel().astype(np.int32), ys.ravel().astype(np.int32)


_bridge = None


def bridge_lib():
    """oracle/_ref/libgnxbridge.so: gnx::CUDAPathIntegrator + harness hooks (pulls in libgnxref.so and libgnxrt.so).
    Loaded on first use, so that callers of the reference alone (bench.py --impl reference) map nothing of the product."""
    global _bridge
    if _bridge is None:
        ctypes.CDLL(REF_LIB)
        b = ctypes.CDLL(BRIDGE_LIB)
        b.gnxh_flatten.restype = vp
        b.gnxh_flatten.argtypes = [vp]
        b.gnxh_render_cuda.argtypes = [vp, ci, vp, vp, vp]
        b.gnxh_render_cuda_passes.argtypes = [vp, ci, ci, vp, vp, vp, vp, vp]
        b.gnxh_time_cuda_render.argtypes = [vp, ci, ci]
        b.gnxh_time_cuda_render.restype = ctypes.c_double
        b.gnxh_cuda_primary_hits.argtypes = [vp, ci, vp]
        b.gnxh_ordered_to_original.argtypes = [vp, ci, vp, vp]
        b.gnxh_cuda_set_devices.argtypes = [vp, ci, vp, ci]
        b.gnxh_cuda_set_progressive.argtypes = [vp, ci]
        _bridge = b
    return _bridge


class RefScene:
    def __init__(self, lib, h, width, height, spp):
        self.lib, self.h, self.width, self.height, self.spp = lib, h, width, height, spp
        self._desc = None

    @property
    def blib(self):
        return bridge_lib()

    @property
    def desc(self):
        """gnx_scene_desc* produced by the bridge's FlattenScene from the live pbr::Scene."""
        if self._desc is None:
            self._desc = self.blib.gnxh_flatten(self.h)
            if not self._desc:
                raise RuntimeError(self.lib.gnxh_scene_error(self.h).decode())
        return self._desc

    def set_light_strategy(self, strategy):
        """lightSampleStrategy of the reference integrators and of the drop-in class: LIGHTS_UNIFORM / SPATIAL / POWER."""
        self.lib.gnxh_scene_set_light_strategy(self.h, int(strategy))
        self._desc = None  # the product-side state (and the flattened description it owns) is dropped with the setting

    def set_sampler(self, kind):
        """0 = HaltonSampler, 2 = the Sobol' GlobalSampler (gnxraytracer_b200/bridge/SobolSampler.h) — for both integrators."""
        self.lib.gnxh_scene_set_sampler(self.h, int(kind))
        self._desc = None

    def set_gaussian_filter(self, radius, alpha):
        """Film of the drop-in class: GaussianFilter(radius, alpha); radius <= 0 = the reference's box average."""
        self.lib.gnxh_scene_set_gaussian_filter(self.h, float(radius), float(alpha))
        self._desc = None

    def reference_gaussian_film(self, radius, alpha, max_depth=5):
        """Film::AddSample splat of the reference's own samples / Li / GaussianFilter::Evaluate: (image, sums)."""
        out = np.zeros((self.height, self.width, 4), np.float32)
        sums = np.zeros((self.height, self.width, 4), np.float32)
        rc = self.lib.gnxh_reference_gaussian_film(self.h, max_depth, float(radius), float(alpha), out.ctypes.data, sums.ctypes.data)
        assert rc == 0
        return out, sums

    def render_reference(self, max_depth=5, threads=0):
        out = np.zeros((self.height, self.width, 4), np.float32)
        sec = ctypes.c_double()
        rc = self.lib.gnxh_render_reference(self.h, max_depth, threads, out.ctypes.data, ctypes.byref(sec))
        assert rc == 0
        return out, sec.value

    def reference_samples(self, px, py, sample, max_depth=5, want_rgb=True, want_prim=True):
        self.desc  # builds the pointer -> ordered-index map
        n = px.size
        rgb = np.zeros((n, 3), np.float32)
        prim = np.zeros(n, np.int32)
        rc = self.lib.gnxh_reference_samples(self.h, max_depth, n, px.ctypes.data, py.ctypes.data, sample.ctypes.data,
                                             rgb.ctypes.data if want_rgb else None, prim.ctypes.data if want_prim else None)
        assert rc == 0
        return rgb, prim

    def sample_dims(self, index, dim):
        out = np.zeros(index.size, np.float32)
        self.lib.gnxh_reference_sample_dims(self.h, index.size, index.ctypes.data, dim.ctypes.data, out.ctypes.data)
        return out

    def sample_index(self, x, y, s):
        return self.lib.gnxh_reference_sample_index(self.h, int(x), int(y), int(s))

    def render_cuda(self, max_depth=5):
        """Through the drop-in class gnx::CUDAPathIntegrator::Render (needs a GPU)."""
        out = np.zeros((self.height, self.width, 4), np.float32)
        sec = ctypes.c_double()
        st = Stats()
        rc = self.blib.gnxh_render_cuda(self.h, max_depth, out.ctypes.data, ctypes.byref(sec), ctypes.byref(st))
        if rc != 0:
            raise RuntimeError(self.lib.gnxh_scene_error(self.h).decode())
        return out, sec.value, st

    def render_cuda_passes(self, n_passes, max_depth=5):
        """n_passes calls of CUDAPathIntegrator::Render on a cleared FrameBuffer: (float buffer, 8-bit buffer, wall seconds
        of the last call)."""
        out = np.zeros((self.height, self.width, 4), np.float32)
        u8 = np.zeros((self.height, self.width, 4), np.uint8)
        sec, wall = ctypes.c_double(), ctypes.c_double()
        rc = self.blib.gnxh_render_cuda_passes(self.h, max_depth, n_passes, out.ctypes.data, u8.ctypes.data, ctypes.byref(sec),
                                               ctypes.byref(wall), None)
        if rc != 0:
            raise RuntimeError(self.lib.gnxh_scene_error(self.h).decode())
        return out, u8, wall.value

    def render_reference_passes(self, n_passes, max_depth=5, threads=0):
        """n_passes calls of the reference's Render on a cleared FrameBuffer: (float buffer, 8-bit buffer)."""
        out = np.zeros((self.height, self.width, 4), np.float32)
        u8 = np.zeros((self.height, self.width, 4), np.uint8)
        rc = self.lib.gnxh_render_reference_passes(self.h, max_depth, threads, n_passes, out.ctypes.data, u8.ctypes.data, None)
        assert rc == 0
        return out, u8

    def time_cuda_render(self, n_calls, max_depth=5):
        """Wall seconds per CUDAPathIntegrator::Render call on the uploaded scene (after one warm-up call)."""
        t = self.blib.gnxh_time_cuda_render(self.h, max_depth, n_calls)
        if t < 0:
            raise RuntimeError(self.lib.gnxh_scene_error(self.h).decode())
        return t

    def set_devices(self, devices, partition=0):
        """GPUs the drop-in class drives ([] = device 0 alone) and the gnx_partition of its renders."""
        ids = (ctypes.c_int * max(1, len(devices)))(*devices)
        self.blib.gnxh_cuda_set_devices(self.h, len(devices), ids, partition)

    def set_progressive(self, on):
        self.blib.gnxh_cuda_set_progressive(self.h, 1 if on else 0)

    def cuda_primary_hits(self, sample=0):
        out = np.zeros(self.width * self.height, np.int32)
        rc = self.blib.gnxh_cuda_primary_hits(self.h, sample, out.ctypes.data)
        if rc != 0:
            raise RuntimeError(self.lib.gnxh_scene_error(self.h).decode())
        return out

    def to_original(self, ordered):
        """BVH-ordered primitive indices (a bridge-flattened scene's prim_id) -> original scene order."""
        ordered = np.ascontiguousarray(ordered, np.int32).ravel()
        out = np.zeros(ordered.size, np.int32)
        self.blib.gnxh_ordered_to_original(self.h, ordered.size, ordered.ctypes.data, out.ctypes.data)
        return out

    def close(self):
        if self.h:
            self.lib.gnxh_scene_destroy(self.h)
            self.h = None


class Ref:
    def __init__(self):
        l = ctypes.CDLL(REF_LIB)
        l.gnxh_scene_create.restype = vp
        l.gnxh_scene_create.argtypes = [ctypes.c_char_p] + [ci] * 6
        l.gnxh_scene_error.restype = ctypes.c_char_p
        l.gnxh_scene_error.argtypes = [vp]
        l.gnxh_scene_destroy.argtypes = [vp]
        l.gnxh_scene_num_prims.argtypes = [vp]
        l.gnxh_scene_bvh_seconds.argtypes = [vp]
        l.gnxh_scene_set_light_strategy.argtypes = [vp, ci]
        l.gnxh_scene_set_sampler.argtypes = [vp, ci]
        l.gnxh_scene_set_gaussian_filter.argtypes = [vp, ctypes.c_float, ctypes.c_float]
        l.gnxh_reference_gaussian_film.argtypes = [vp, ci, ctypes.c_float, ctypes.c_float, vp, vp]
        l.gnxh_reference_gaussian_eval.argtypes = [ctypes.c_float, ctypes.c_float, ci, vp, vp, vp]
        l.gnxh_scene_bvh_seconds.restype = ctypes.c_double
        l.gnxh_render_reference.argtypes = [vp, ci, ci, vp, vp]
        l.gnxh_render_reference_passes.argtypes = [vp, ci, ci, ci, vp, vp, vp]
        l.gnxh_reference_samples.argtypes = [vp, ci, ci, vp, vp, vp, vp, vp]
        l.gnxh_reference_sample_dims.argtypes = [vp, ci, vp, vp, vp]
        l.gnxh_reference_sample_index.restype = ctypes.c_int64
        l.gnxh_reference_sample_index.argtypes = [vp, ci, ci, ci]
        self.lib = l

    def scene(self, preset, width, height, spp):
        name, p0, p1, p2 = SCENES[preset]
        h = self.lib.gnxh_scene_create(name.encode(), width, height, spp, p0, p1, p2)
        err = self.lib.gnxh_scene_error(h).decode()
        if err:
            raise RuntimeError(err)
        return RefScene(self.lib, h, width, height, spp)

    def scene_named(self, name, width, height, spp, p0=0, p1=0, p2=0):
        """A harness scene by its raw name, e.g. "dragon3d:<path>" (mesh read by the reference's plyInfo)."""
        h = self.lib.gnxh_scene_create(name.encode(), width, height, spp, p0, p1, p2)
        err = self.lib.gnxh_scene_error(h).decode()
        if err:
            raise RuntimeError(err)
        return RefScene(self.lib, h, width, height, spp)

    def max_threads(self):
        return self.lib.gnxh_max_threads()


class EmulScene:
    def __init__(self, lib, desc):
        self.lib = lib
        self.h = lib.gnxe_create(desc)

    def render(self, params):
        out = np.zeros((params.height, params.width, 4), np.float32)
        st = Stats()
        self.lib.gnxe_render(self.h, ctypes.byref(params), out.ctypes.data, ctypes.byref(st))
        return out, st

    def render_share(self, params, n_shares, share):
        """The frame share `share` of an n_shares-device job hands to the reduce (the library's partition arithmetic)."""
        out = np.zeros((params.height, params.width, 4), np.float32)
        self.lib.gnxe_render_share(self.h, ctypes.byref(params), n_shares, share, out.ctypes.data)
        return out

    def samples(self, params, px, py, sample):
        rgb = np.zeros((px.size, 3), np.float32)
        self.lib.gnxe_samples(self.h, ctypes.byref(params), px.size, px.ctypes.data, py.ctypes.data, sample.ctypes.data, rgb.ctypes.data)
        return rgb

    def primary_hits(self, width, height, sample=0):
        out = np.zeros(width * height, np.int32)
        self.lib.gnxe_primary_hits(self.h, width, height, sample, out.ctypes.data)
        return out

    def sample_dims(self, index, dim):
        out = np.zeros(index.size, np.float32)
        self.lib.gnxe_sample_dims(self.h, index.size, index.ctypes.data, dim.ctypes.data, out.ctypes.data)
        return out

    def sample_index(self, x, y, s):
        return self.lib.gnxe_sample_index(self.h, int(x), int(y), int(s))

    def close(self):
        if self.h:
            self.lib.gnxe_destroy(self.h)
            self.h = None


class Emul:
    def __init__(self):
        l = ctypes.CDLL(EMUL_LIB)
        l.gnxe_create.restype = vp
        l.gnxe_create.argtypes = [vp]
        l.gnxe_destroy.argtypes = [vp]
        l.gnxe_render.argtypes = [vp, ctypes.POINTER(RenderParams), vp, ctypes.POINTER(Stats)]
        l.gnxe_samples.argtypes = [vp, ctypes.POINTER(RenderParams), ci, vp, vp, vp, vp]
        l.gnxe_render_share.argtypes = [vp, ctypes.POINTER(RenderParams), ci, ci, vp]
        l.gnxe_primary_hits.argtypes = [vp, ci, ci, ci, vp]
        l.gnxe_gaussian_eval.argtypes = [ctypes.c_float, ctypes.c_float, ci, vp, vp, vp]
        l.gnxe_sample_dims.argtypes = [vp, ci, vp, vp, vp]
        l.gnxe_sample_index.restype = ctypes.c_int64
        l.gnxe_sample_index.argtypes = [vp, ci, ci, ci]
        self.lib = l

    def scene(self, desc):
        return EmulScene(self.lib, desc)


def rel_mse(img, ref_img):
    """Relative MSE as used for renderer comparisons: mean((a-b)^2 / (b^2 + eps)) over RGB."""
    a, b = img[..., :3].astype(np.float64), ref_img[..., :3].astype(np.float64)
    return float(np.mean((a - b) ** 2 / (b ** 2 + 1e-2)))


# ---- oracle/restate: independent plain-C++ restatement of the path (oracle/_build/libgnxrestate.so) ----
RESTATE_LIB = os.path.join(ROOT, "oracle", "_build", "libgnxrestate.so")


class RestateScene:
    def __init__(self, lib, desc):
        self.lib, self.h = lib, lib.gnxr_create(desc)

    def render(self, params):
        out = np.zeros((params.height, params.width, 4), np.float32)
        counts = np.zeros(5, np.uint64)
        rc = self.lib.gnxr_render(self.h, ctypes.byref(params), out.ctypes.data, counts.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"restatement does not cover this scene (rc={rc})")
        return out, dict(zip(["rays_extend", "rays_shadow", "rays_mis", "nodes_visited", "tris_tested"], (int(c) for c in counts)))

    def samples(self, params, px, py, sample):
        rgb = np.zeros((px.size, 3), np.float32)
        rc = self.lib.gnxr_samples(self.h, ctypes.byref(params), px.size, px.ctypes.data, py.ctypes.data, sample.ctypes.data, rgb.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"restatement does not cover this scene (rc={rc})")
        return rgb

    def primary_hits(self, width, height, sample=0):
        out = np.zeros(width * height, np.int32)
        self.lib.gnxr_primary_hits(self.h, width, height, sample, out.ctypes.data)
        return out

    def sample_dims(self, index, dim):
        out = np.zeros(index.size, np.float32)
        self.lib.gnxr_sample_dims(self.h, index.size, index.ctypes.data, dim.ctypes.data, out.ctypes.data)
        return out

    def sample_index(self, x, y, s):
        return self.lib.gnxr_sample_index(self.h, int(x), int(y), int(s))

    def close(self):
        if self.h:
            self.lib.gnxr_destroy(self.h)
            self.h = None


class Restate:
    def __init__(self):
        l = ctypes.CDLL(RESTATE_LIB)
        l.gnxr_create.restype = vp
        l.gnxr_create.argtypes = [vp]
        l.gnxr_destroy.argtypes = [vp]
        l.gnxr_render.argtypes = [vp, ctypes.POINTER(RenderParams), vp, vp]
        l.gnxr_samples.argtypes = [vp, ctypes.POINTER(RenderParams), ci, vp, vp, vp, vp]
        l.gnxr_primary_hits.argtypes = [vp, ci, ci, ci, vp]
        l.gnxr_sample_dims.argtypes = [vp, ci, vp, vp, vp]
        l.gnxr_sample_index.restype = ctypes.c_int64
        l.gnxr_sample_index.argtypes = [vp, ci, ci, ci]
        self.lib = l

    def scene(self, desc):
        return RestateScene(self.lib, desc)
