"""The C-ABI libraries load on a CPU-only box and export every symbol the headers declare; compute
entry points fail loudly without a GPU (there is no CPU fallback in the product)."""
import ctypes
import os
import re

import pytest

from gnxraytracer_b200 import api
from gnxraytracer_b200.build import build_product, repo_root


def _declared(header, prefix):
    text = open(os.path.join(repo_root(), "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(" + prefix + r"_\w+)\s*\(", text)))


def test_libgnxrt_exports_every_declared_symbol():
    build_product()
    lib = ctypes.CDLL(api.library_path())
    names = _declared("gnxrt.h", "gnx")
    assert set(api.EXPORTS) <= set(names)
    assert len(names) >= 11
    for n in names:
        assert hasattr(lib, n), f"libgnxrt.so does not export {n}"
    assert lib.gnx_abi_version() == 4


def test_scenekit_exports_every_declared_symbol():
    build_product()
    lib = ctypes.CDLL(api.scenekit_path())
    names = _declared("gnx_scenekit.h", "gnxsk")
    assert set(api.SCENEKIT_EXPORTS) <= set(names)
    for n in names:
        assert hasattr(lib, n), f"libgnxscenekit.so does not export {n}"


def test_no_cpu_fallback_without_a_device():
    lib = api.load_library()
    if lib.gnx_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(api.GnxError) as e:
        api.Context(0)
    assert e.value.code == -2  # GNX_ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def test_product_library_does_not_link_the_oracle():
    import subprocess
    out = subprocess.run(["ldd", api.library_path()], capture_output=True, text=True).stdout
    out += subprocess.run(["ldd", api.scenekit_path()], capture_output=True, text=True).stdout
    assert "gnxref" not in out and "gnxemul" not in out and "gnxrestate" not in out
