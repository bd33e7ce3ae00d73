"""ctypes bindings of include/gnxrt.h and include/gnx_scenekit.h.

Plumbing only: every call lands in libgnxrt.so (CUDA).  If the library is missing, or no CUDA device
is usable, the calls raise — there is no Python or CPU fallback.
"""
import ctypes
import os

import numpy as np

from .build import repo_root

c_int32, c_float, c_u64, c_double, c_void_p = ctypes.c_int32, ctypes.c_float, ctypes.c_uint64, ctypes.c_double, ctypes.c_void_p

LIGHTS_UNIFORM, LIGHTS_SPATIAL, LIGHTS_POWER = 0, 1, 2
INTEGRATOR_PATH, INTEGRATOR_VOLPATH, INTEGRATOR_WHITTED, INTEGRATOR_DIRECT, INTEGRATOR_DIRECT_ALL = 0, 1, 2, 3, 4
PARTITION_SAMPLES, PARTITION_TILES = 0, 1
FILM_BOX, FILM_GAUSSIAN, FILM_GAUSSIAN_SUMS = 0, 1, 2

STATUS = {0: "GNX_OK", -1: "GNX_ERR_INVALID", -2: "GNX_ERR_NO_DEVICE", -3: "GNX_ERR_CUDA", -4: "GNX_ERR_UNSUPPORTED",
          -5: "GNX_ERR_NO_SCENE"}


class GnxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{STATUS.get(code, code)}: {msg}")
        self.code = code


class RenderParams(ctypes.Structure):
    """gnx_render_params (include/gnxrt.h)."""
    _fields_ = [
        ("width", c_int32), ("height", c_int32), ("spp", c_int32), ("first_sample", c_int32),
        ("spp_normalize", c_int32), ("max_depth", c_int32), ("rr_threshold", c_float), ("integrator", c_int32),
        ("light_strategy", c_int32), ("film", c_int32), ("filter_radius", c_float), ("filter_alpha", c_float),
        ("batch_spp", c_int32), ("partition", c_int32),
    ]

    @classmethod
    def make(cls, width, height, spp, max_depth=5, first_sample=0, spp_normalize=0, rr_threshold=1.0,
             light_strategy=LIGHTS_SPATIAL, integrator=INTEGRATOR_PATH, batch_spp=0, film=FILM_BOX, filter_radius=0.0,
             filter_alpha=0.0, partition=0):
        return cls(width, height, spp, first_sample, spp_normalize, max_depth, rr_threshold, integrator,
                   light_strategy, film, filter_radius, filter_alpha, batch_spp, partition)


class Stats(ctypes.Structure):
    """gnx_stats (include/gnxrt.h)."""
    _fields_ = [
        ("paths", c_u64), ("rays_extend", c_u64), ("rays_shadow", c_u64), ("rays_mis", c_u64),
        ("nodes_visited", c_u64), ("tris_tested", c_u64), ("device_ms", c_double), ("ms_raygen", c_double),
        ("ms_extend", c_double), ("ms_shade", c_double), ("ms_shadow", c_double), ("ms_film", c_double),
        ("kernel_launches", c_u64), ("bytes_algorithmic", c_u64),
        ("extend_nodes", c_u64), ("extend_tris", c_u64), ("extend_launches", c_u64), ("extend_bytes", c_u64),
        ("vp_ms", c_double * 5), ("vp_items", c_u64 * 5), ("vp_track_steps", c_u64), ("vp_rounds", c_u64),
    ]

    @property
    def rays(self):
        return self.rays_extend + self.rays_shadow + self.rays_mis

    def as_dict(self):
        d = {k: (list(getattr(self, k)) if k in ("vp_ms", "vp_items") else getattr(self, k)) for k, _ in self._fields_}
        d["rays"] = self.rays
        return d


EXPORTS = ["gnx_abi_version", "gnx_device_count", "gnx_create", "gnx_destroy", "gnx_last_error", "gnx_upload_scene", "gnx_bvh_build_ms",
           "gnx_render", "gnx_render_device", "gnx_primary_hits", "gnx_sample_dimensions", "gnx_tonemap_rgba8",
           "gnx_create_multi", "gnx_num_devices", "gnx_comm_unique_id", "gnx_comm_attach", "gnx_render_framebuffer"]

_lib = None


def library_path():
    # GNX_LIB: another build of the same CUDA library (tuning experiments, tools/profile_step.py)
    return os.environ.get("GNX_LIB") or os.path.join(repo_root(), "gnxraytracer_b200", "lib", "libgnxrt.so")


def load_library():
    """Loads libgnxrt.so; raises if it has not been built (run `python -c 'import __graft_entry__ as g; g.build()'`)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: the CUDA library must be built first (there is no fallback)")
    lib = ctypes.CDLL(path)
    lib.gnx_create.argtypes = [ctypes.POINTER(c_void_p), ctypes.c_int]
    lib.gnx_destroy.argtypes = [c_void_p]
    lib.gnx_destroy.restype = None
    lib.gnx_last_error.argtypes = [c_void_p]
    lib.gnx_last_error.restype = ctypes.c_char_p
    lib.gnx_upload_scene.argtypes = [c_void_p, c_void_p]
    lib.gnx_bvh_build_ms.argtypes = [c_void_p]
    lib.gnx_bvh_build_ms.restype = c_double
    lib.gnx_render.argtypes = [c_void_p, ctypes.POINTER(RenderParams), c_void_p, ctypes.POINTER(Stats)]
    lib.gnx_render_device.argtypes = [c_void_p, ctypes.POINTER(RenderParams), c_void_p, c_void_p, ctypes.POINTER(Stats)]
    lib.gnx_primary_hits.argtypes = [c_void_p, ctypes.POINTER(RenderParams), c_int32, c_void_p]
    lib.gnx_sample_dimensions.argtypes = [c_void_p, c_int32, c_void_p, c_void_p, c_void_p]
    lib.gnx_tonemap_rgba8.argtypes = [c_void_p, c_void_p, c_int32, c_void_p]
    lib.gnx_render_framebuffer.argtypes = [c_void_p, ctypes.POINTER(RenderParams), c_int32, c_void_p, c_void_p, ctypes.POINTER(Stats)]
    lib.gnx_create_multi.argtypes = [ctypes.POINTER(c_void_p), ctypes.POINTER(ctypes.c_int), ctypes.c_int]
    lib.gnx_num_devices.argtypes = [c_void_p]
    lib.gnx_comm_unique_id.argtypes = [c_void_p]
    lib.gnx_comm_attach.argtypes = [c_void_p, ctypes.c_int, ctypes.c_int, c_void_p]
    _lib = lib
    return lib


class Context:
    """One gnx_ctx: one GPU (device = int), or several GPUs of this node driven from this process (devices = [ids],
    gnx_create_multi: the library replicates the scene, splits every render and reduces onto devices[0])."""

    def __init__(self, device=0, devices=None):
        self.lib = load_library()
        self.h = c_void_p()
        if devices is not None:
            ids = (ctypes.c_int * len(devices))(*devices)
            rc = self.lib.gnx_create_multi(ctypes.byref(self.h), ids, len(devices))
            device = devices[0]
        else:
            rc = self.lib.gnx_create(ctypes.byref(self.h), device)
        if rc != 0:
            raise GnxError(rc, self.lib.gnx_last_error(None).decode())
        self.device = device

    @property
    def num_devices(self):
        return self.lib.gnx_num_devices(self.h)

    def comm_unique_id(self):
        """128-byte id for gnx_comm_attach (rank 0 creates it, the host program hands it to every rank)."""
        buf = ctypes.create_string_buffer(128)
        rc = self.lib.gnx_comm_unique_id(buf)
        if rc != 0:
            raise GnxError(rc, self.lib.gnx_last_error(None).decode())
        return bytes(buf.raw)

    def comm_attach(self, n_ranks, rank, comm_id):
        """Joins this single-device context to an n_ranks-process job: renders become collective (see gnxrt.h)."""
        self._check(self.lib.gnx_comm_attach(self.h, n_ranks, rank, ctypes.c_char_p(comm_id)))

    def _check(self, rc):
        if rc != 0:
            raise GnxError(rc, self.lib.gnx_last_error(self.h).decode())

    def upload(self, desc_ptr):
        """desc_ptr: address of a gnx_scene_desc (from SceneKit or from the bridge's FlattenScene)."""
        self._check(self.lib.gnx_upload_scene(self.h, c_void_p(int(desc_ptr))))

    @property
    def bvh_build_ms(self):
        """Device time of the BVH build done by the last upload (0 when the description carried its own nodes)."""
        return self.lib.gnx_bvh_build_ms(self.h)

    def render(self, params, out=None, want_stats=True):
        if out is None:
            out = np.empty((params.height, params.width, 4), np.float32)
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.size == params.width * params.height * 4
        st = Stats()
        self._check(self.lib.gnx_render(self.h, ctypes.byref(params), out.ctypes.data, ctypes.byref(st) if want_stats else None))
        return out, st

    def render_framebuffer(self, params, pass_count, fbuffer, ubuffer, want_stats=False):
        """gnx_render_framebuffer: the reference FrameBuffer's running mean + 8-bit tonemapped image, computed on the
        device; fbuffer float32[h, w, 4] (in/out, colour channels only), ubuffer uint8[h, w, 4] (out)."""
        st = Stats()
        fp = fbuffer.ctypes.data if hasattr(fbuffer, "ctypes") else fbuffer
        up = ubuffer.ctypes.data if hasattr(ubuffer, "ctypes") else ubuffer
        self._check(self.lib.gnx_render_framebuffer(self.h, ctypes.byref(params), int(pass_count), c_void_p(int(fp)) if fp else None,
                                                    c_void_p(int(up)) if up else None, ctypes.byref(st) if want_stats else None))
        return st

    def render_host_ptr(self, params, host_ptr, want_stats=True):
        """Like render(), into caller-owned (e.g. pinned) host memory."""
        st = Stats()
        self._check(self.lib.gnx_render(self.h, ctypes.byref(params), c_void_p(int(host_ptr)), ctypes.byref(st) if want_stats else None))
        return st

    def render_device(self, params, dev_ptr, stream=0, want_stats=True):
        st = Stats()
        self._check(self.lib.gnx_render_device(self.h, ctypes.byref(params), c_void_p(int(dev_ptr)), c_void_p(int(stream)),
                                               ctypes.byref(st) if want_stats else None))
        return st

    def primary_hits(self, params, sample=0):
        out = np.empty((params.height, params.width), np.int32)
        self._check(self.lib.gnx_primary_hits(self.h, ctypes.byref(params), sample, out.ctypes.data))
        return out

    def sample_dimensions(self, index, dim):
        index = np.ascontiguousarray(index, np.int64)
        dim = np.ascontiguousarray(dim, np.int32)
        out = np.empty(index.size, np.float32)
        self._check(self.lib.gnx_sample_dimensions(self.h, index.size, index.ctypes.data, dim.ctypes.data, out.ctypes.data))
        return out

    def tonemap(self, rgba):
        rgba = np.ascontiguousarray(rgba, np.float32)
        out = np.empty(rgba.shape[:-1] + (4,), np.uint8)
        self._check(self.lib.gnx_tonemap_rgba8(self.h, rgba.ctypes.data, rgba.size // 4, out.ctypes.data))
        return out

    def close(self):
        if self.h:
            self.lib.gnx_destroy(self.h)
            self.h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- host-side scene kit (include/gnx_scenekit.h): builds gnx_scene_desc for the BASELINE configs ----
SCENEKIT_EXPORTS = ["gnxsk_create", "gnxsk_destroy", "gnxsk_desc", "gnxsk_error", "gnxsk_num_prims", "gnxsk_build_seconds", "gnxsk_strip_bvh",
                    "gnxsk_corrupt"]
_sk = None


def scenekit_path():
    return os.path.join(repo_root(), "gnxraytracer_b200", "lib", "libgnxscenekit.so")


def load_scenekit():
    global _sk
    if _sk is not None:
        return _sk
    path = scenekit_path()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: build it first")
    sk = ctypes.CDLL(path)
    sk.gnxsk_create.restype = c_void_p
    sk.gnxsk_create.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                ctypes.c_int, ctypes.c_char_p]
    sk.gnxsk_destroy.argtypes = [c_void_p]
    sk.gnxsk_destroy.restype = None
    sk.gnxsk_desc.argtypes = [c_void_p]
    sk.gnxsk_desc.restype = c_void_p
    sk.gnxsk_error.argtypes = [c_void_p]
    sk.gnxsk_error.restype = ctypes.c_char_p
    sk.gnxsk_num_prims.argtypes = [c_void_p]
    sk.gnxsk_build_seconds.argtypes = [c_void_p]
    sk.gnxsk_build_seconds.restype = c_double
    sk.gnxsk_strip_bvh.argtypes = [c_void_p]
    sk.gnxsk_strip_bvh.restype = None
    sk.gnxsk_corrupt.argtypes = [c_void_p, ctypes.c_int]
    ip = ctypes.POINTER(ctypes.c_int)
    sk.gnxsk_mesh_info.argtypes = [ctypes.c_char_p, ip, ip, ip, ip, ctypes.c_char_p, ctypes.c_int]
    sk.gnxsk_write_knot_3d.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.c_int]
    _sk = sk
    return sk


def resources_dir():
    """Where the HDR environment maps and the volume live: GNX_RESOURCES, else gnxraytracer_b200/resources (staged from the
    reference's Resources/ by the build, gnxraytracer_b200.build.stage_resources)."""
    env = os.environ.get("GNX_RESOURCES")
    if env:
        return env
    return os.path.join(repo_root(), "gnxraytracer_b200", "resources")


def mesh_info(path):
    """Parses a .3d file (the reference's mesh format, shape/plyRead.h) or a Wavefront .obj with the scene kit's readers:
    {"vertices", "triangles", "has_uv", "has_normals"}; raises RuntimeError with the reader's message."""
    sk = load_scenekit()
    v = [ctypes.c_int() for _ in range(4)]
    err = ctypes.create_string_buffer(512)
    if sk.gnxsk_mesh_info(str(path).encode(), *[ctypes.byref(x) for x in v], err, 512) != 0:
        raise RuntimeError(err.value.decode())
    return {"vertices": v[0].value, "triangles": v[1].value, "has_uv": bool(v[2].value), "has_normals": bool(v[3].value)}


def write_knot_3d(path, nu, nv):
    """Writes the kit's nu x nv torus knot as a .3d file (tests, tools)."""
    sk = load_scenekit()
    err = ctypes.create_string_buffer(512)
    if sk.gnxsk_write_knot_3d(str(path).encode(), int(nu), int(nv), err, 512) != 0:
        raise RuntimeError(err.value.decode())


class SceneKit:
    """A scene built natively by the product's host-side kit (own BVH build, own env-map tables)."""

    def __init__(self, name, width, height, spp, p0=0, p1=0, p2=0, resources=None):
        self.sk = load_scenekit()
        res = (resources or resources_dir()).encode()
        self.h = self.sk.gnxsk_create(name.encode(), width, height, spp, p0, p1, p2, res)
        err = self.sk.gnxsk_error(self.h).decode()
        if err:
            raise RuntimeError(f"scenekit({name}): {err}")
        self.width, self.height, self.spp = width, height, spp

    @property
    def desc(self):
        return self.sk.gnxsk_desc(self.h)

    @property
    def num_prims(self):
        return self.sk.gnxsk_num_prims(self.h)

    @property
    def build_seconds(self):
        return self.sk.gnxsk_build_seconds(self.h)

    def strip_bvh(self):
        """Drop the host-built BVH: the library then builds one on the GPU at upload."""
        self.sk.gnxsk_strip_bvh(self.h)

    def corrupt(self, kind):
        """Damage the description (see gnx_scenekit.h) — for tests of the upload validation."""
        return self.sk.gnxsk_corrupt(self.h, int(kind))

    def close(self):
        if self.h:
            self.sk.gnxsk_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
