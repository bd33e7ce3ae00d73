"""gnxraytracer_b200 — B200-native path-tracing core behind GNXRayTracer's Integrator::Render.

The product is the C-ABI shared library ``lib/libgnxrt.so`` (CUDA kernels for sm_100a, declared in
``include/gnxrt.h``) plus the reference-side bridge ``bridge/CUDAPathIntegrator.cpp``.  This Python
package is plumbing for tests and ``bench.py``: ctypes bindings, the build recipe, and the
``torch.distributed`` sharding of a render over several GPUs.  Nothing here computes an image.
"""
from .build import build_all, build_product, repo_root  # noqa: F401
from .api import (  # noqa: F401
    Context,
    GnxError,
    RenderParams,
    Stats,
    SceneKit,
    load_library,
    LIGHTS_UNIFORM,
    LIGHTS_SPATIAL,
)
