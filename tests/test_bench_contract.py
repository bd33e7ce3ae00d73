"""The JSON line bench.py prints: exactly one line on stdout, with the keys the driver reads.  The reference arm runs
on the CPU (oracle/_ref); the product arm needs a GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e"}


def _run(args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, f"stdout must carry exactly one line, got {len(lines)}"
    return json.loads(lines[0])


def test_reference_arm_prints_one_json_line():
    from _harness import REF_LIB
    if not os.path.exists(REF_LIB):
        pytest.skip("oracle/_ref not built")
    d = _run(["--impl", "reference", "--workload", "c1", "--steps", "1", "--warmup", "0", "--ref-spp", "1"])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "Mpaths/s" and d["unit"] == "Mpaths/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


@pytest.mark.gpu
def test_product_arm_prints_one_json_line():
    d = _run(["--workload", "c1", "--steps", "2", "--warmup", "3", "--no-cpu-baseline"])
    assert BASE_KEYS | {"roofline", "gpu_launches", "clocks"} <= set(d)
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["value"] > 0 and d["gpu_launches"] > 0
    r = d["roofline"]
    assert r["bound_contract"] in ("hbm", "tensor") and r["unit"] == "GB/s" and 0 < r["frac"] < 2 and r["peak"] > 0
    assert 0.99 <= d["config"]["hit_coverage"] <= 1.0  # the Cornell box is closed towards the camera
    e = d["e2e"]
    assert e["value"] > 0 and e["d2h_bytes_per_step"] == 512 * 512 * 20 and e["h2d_bytes_per_step"] > 0
    assert e["value"] <= d["value"] * 1.05
    assert d["clocks"] is None or "sm_mhz" in d["clocks"]


def test_reference_arm_under_torchrun_prints_one_line():
    """N > 1: the driver launches the reference arm like the product arm; rank 0 alone runs and prints, the other ranks
    exit 0 without work, and nothing else (launcher or library banners) reaches stdout."""
    from _harness import REF_LIB
    if not os.path.exists(REF_LIB):
        pytest.skip("oracle/_ref not built")
    port = 29600 + (os.getpid() % 300)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--workload", "c1",
                        "--steps", "1", "--warmup", "0", "--ref-spp", "1"], capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0
