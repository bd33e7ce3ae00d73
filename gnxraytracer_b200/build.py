"""Build recipes: nvcc for the product library (sm_100a only), g++ for the host-side scene kit,
make for the oracle (reference objects + harness) and the CPU emulation test tool."""
import os
import shutil
import subprocess
import sys


def repo_root() -> str:
    return os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
    # no FMA contraction: discrete decisions (hit / lobe / light choice) then agree with the x86 -O2
    # reference on all but grazing cases
    "-fmad=false",
    # kernel templates are instantiated in one translation unit and launched from another (csrc/tu_*.cu): keep their host
    # stubs external, which is CUDA 12's default, stated explicitly because nvcc announces a change of that default
    "-static-global-template-stub=false", "-diag-suppress", "20281",
    "-Xcompiler", "-fPIC",
]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources if os.path.exists(s))


def _run(cmd, **kw):
    print("+", " ".join(cmd), file=sys.stderr, flush=True)
    subprocess.run(cmd, check=True, **kw)


def _host_cxx():
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def build_product(force: bool = False) -> str:
    """nvcc -> gnxraytracer_b200/lib/libgnxrt.so, g++ -> lib/libgnxscenekit.so."""
    root = repo_root()
    csrc = os.path.join(root, "gnxraytracer_b200", "csrc")
    host = os.path.join(root, "gnxraytracer_b200", "host")
    lib = os.path.join(root, "gnxraytracer_b200", "lib")
    os.makedirs(lib, exist_ok=True)
    out = os.path.join(lib, "libgnxrt.so")
    srcs = [os.path.join(csrc, f) for f in os.listdir(csrc)] + [os.path.join(root, "include", "gnxrt.h")]
    if force or _newer(out, srcs):
        # one object per translation unit (gnx_render.cu: host code + the light kernels; tu_*.cu: explicit instantiations
        # of the heavy kernel templates), compiled in parallel, linked into one shared library
        from concurrent.futures import ThreadPoolExecutor
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        objdir = os.path.join(root, "build", "obj")
        os.makedirs(objdir, exist_ok=True)
        units = sorted(f for f in os.listdir(csrc) if f.endswith(".cu"))
        headers = [os.path.join(csrc, f) for f in os.listdir(csrc) if not f.endswith(".cu")] + [os.path.join(root, "include", "gnxrt.h")]

        def compile_unit(u):
            obj = os.path.join(objdir, u[:-3] + ".o")
            if force or _newer(obj, [os.path.join(csrc, u)] + headers):
                _run([nvcc, *NVCC_FLAGS, "-I", os.path.join(root, "include"), "-c", os.path.join(csrc, u), "-o", obj])
            return obj
        with ThreadPoolExecutor(max_workers=min(len(units), os.cpu_count() or 4)) as ex:
            objs = list(ex.map(compile_unit, units))
        _run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out, *objs])
    sk = os.path.join(lib, "libgnxscenekit.so")
    sk_src = os.path.join(host, "scenekit.cpp")
    if os.path.exists(sk_src):
        deps = [os.path.join(host, f) for f in os.listdir(host)] + [os.path.join(root, "include", f) for f in os.listdir(os.path.join(root, "include"))]
        if force or _newer(sk, deps):
            _run([_host_cxx(), "-std=c++17", "-O2", "-fopenmp", "-fPIC", "-shared", "-I", os.path.join(root, "include"),
                  "-I", host, "-I", csrc, "-I", "/usr/local/cuda/include", "-x", "c++", sk_src, "-o", sk])
    return out


RESOURCE_FILES = ("MonValley1000.hdr", "TropicalRuins1000.hdr", "awesomeface.jpg", "density_render.70.volume")


def stage_resources() -> str:
    """The environment maps and the volume of the BASELINE configs are DATA of the reference (Resources/), not sources:
    where the reference is present they are staged under gnxraytracer_b200/resources/ (git-ignored, travels to the GPU box
    like the built libraries), so that the product (scene kit, bench.py, smoke()) needs nothing under oracle/."""
    root = repo_root()
    dst = os.path.join(root, "gnxraytracer_b200", "resources")
    src = os.path.join(os.environ.get("GNX_REFERENCE", "/root/reference"), "Resources")
    if os.path.isdir(src):
        os.makedirs(dst, exist_ok=True)
        for f in RESOURCE_FILES:
            if os.path.exists(os.path.join(src, f)) and _newer(os.path.join(dst, f), [os.path.join(src, f)]):
                shutil.copyfile(os.path.join(src, f), os.path.join(dst, f))
    return dst


def build_emul(force: bool = False) -> str:
    """g++ -> tests/emul/_build/libgnxemul.so (TEST TOOL: device functions compiled for the host)."""
    root = repo_root()
    src = os.path.join(root, "tests", "emul", "host_emul.cpp")
    out = os.path.join(root, "tests", "emul", "_build", "libgnxemul.so")
    csrc = os.path.join(root, "gnxraytracer_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc)]
    if force or _newer(out, deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        _run([_host_cxx(), "-std=c++17", "-O2", "-fopenmp", "-fPIC", "-shared", "-x", "c++", "-I", os.path.join(root, "include"),
              "-I", csrc, "-I", "/usr/local/cuda/include", src, "-o", out])
    return out


def build_oracle() -> None:
    """make -C oracle ref (only where /root/reference exists) and make restate."""
    root = repo_root()
    ref = os.environ.get("GNX_REFERENCE", "/root/reference")
    if os.path.isdir(ref):
        _run(["make", "-C", os.path.join(root, "oracle"), "-j8", "ref", "REF=" + ref])
    if os.listdir(os.path.join(root, "oracle", "restate")):
        _run(["make", "-C", os.path.join(root, "oracle"), "restate"])


def build_all(force: bool = False) -> None:
    build_product(force)
    stage_resources()
    build_emul(force)
    build_oracle()
