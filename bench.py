#!/usr/bin/env python
"""bench.py — headline benchmark of the path-tracing hot path.

Workload (BASELINE.json configs[1], "C2"): dragon-class mesh (872 448 triangles; torus-knot stand-in
for the stripped Resources/dragon.3d) with the UI's purple Plastic material under
InfiniteAreaLight(MonValley1000.hdr), 1024 x 1024, 64 spp, maxDepth 5, PathIntegrator + Halton.
One step = one full render = W*H*spp camera paths.

  python bench.py [--gpus N] [--steps K] [--warmup W]           our arm   (CUDA, libgnxrt.so)
  python bench.py --impl reference [...]                         reference arm (the UNMODIFIED reference's
                                                                  OpenMP Render on the host cores)

N > 1 (torchrun, one rank per GPU): every rank attaches its context to ONE job through the library's own
communicator (gnx_comm_attach; torch.distributed only carries the 128-byte id and the timing reduction).  The job is
the per-GPU workload times N (weak scaling: N * 64 spp), dealt out as sample ranges by the library, summed onto rank 0
by ncclReduce queued behind each rank's last film kernel, inside the timed region.  The same line carries a second
record, "strong_scaling": BASELINE config 5 (3840 x 2160 x 1024 spp, ONE fixed job) split over the N ranks.
Prints ONE JSON line on rank 0.
"""
import argparse
import csv
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line: NCCL writes its "NCCL version ..." banner (any NCCL_DEBUG level from VERSION up,
# WARN included) and its warnings to stdout unless told otherwise
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
# ... and NCCL 2.28 prints the banner with a plain printf whatever NCCL_DEBUG_FILE says, so file descriptor 1 itself is pointed at
# stderr for the whole run and the one JSON line is written to a duplicate of the original stdout (emit()).
sys.stdout.flush()
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(obj):
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


METRIC = "Mpaths/s"
# gnx_integrator values (include/gnxrt.h)
PATH, VOLPATH, WHITTED, DIRECT, DIRECT_ALL = 0, 1, 2, 3, 4
WORKLOADS = {
    "c2": dict(scene="dragon", args=(0, 0, 0), w=1024, h=1024, spp=64, depth=5, integ=PATH,
               desc="C2: dragon-class mesh 872448 tris (torus-knot stand-in for dragon.3d), Plastic, InfiniteAreaLight MonValley1000.hdr, 1024x1024, 64 spp, maxDepth 5, PathIntegrator+Halton"),
    "c3": dict(scene="nano", args=(0, 0, 0), w=1920, h=1080, spp=128, depth=5, integ=PATH,
               desc="C3: UV-mapped smooth-shaded mesh (~90k tris, stand-in for nanosuit), DisneyMaterial + ImageTexture, InfiniteAreaLight TropicalRuins1000.hdr, 1920x1080, 128 spp, maxDepth 5"),
    "c4": dict(scene="smoke", args=(0, 0, 0), w=1024, h=1024, spp=64, depth=5, integ=VOLPATH,
               desc="C4: VolPathIntegrator, GridDensityMedium density_render.70.volume inside HomogeneousMedium fog, Matte ground, MonValley env, 1024x1024, 64 spp, maxDepth 5, PCG32 stream sampler"),
    "c1": dict(scene="cornell", args=(0, 3, 0), w=512, h=512, spp=16, depth=5, integ=PATH,
               desc="C1: Cornell box + 2 icospheres (Mirror, Glass), DiffuseAreaLight, 512x512, 16 spp, maxDepth 5"),
    # BASELINE config 5: ONE fixed job split across the ranks by sample range (strong scaling, "scaling": "strong")
    "c5": dict(scene="dragon", args=(0, 0, 0), w=3840, h=2160, spp=1024, depth=5, integ=PATH, strong=True,
               desc="C5: dragon-class mesh 872448 tris, Plastic, MonValley env, 3840x2160, 1024 spp in total, sample ranges dealt to the ranks, maxDepth 5"),
    # The reference UI's live scene (ui/RenderThread.cpp:60-164): the mesh INSIDE the Cornell box, area light + SkyBoxLight,
    # every camera ray hits a surface.  u1p = the PathIntegrator line (:164, maxDepth 15), u1w = the default Whitted (:163).
    "u1p": dict(scene="ui", args=(PATH, 0, 0), w=1024, h=1024, spp=32, depth=15, integ=PATH,
                desc="U1p: the UI's live scene (872448-tri mesh Matte sigma 60 inside the Oren-Nayar Cornell box, area light + SkyBoxLight; ui/RenderThread.cpp:60-164), PathIntegrator maxDepth 15 (line :164), HaltonSampler 32 spp, 1024x1024"),
    "u1w": dict(scene="ui", args=(WHITTED, 0, 0), w=1024, h=1024, spp=32, depth=5, integ=WHITTED,
                desc="U1w: the UI's live scene, WhittedIntegrator maxDepth 5 (the UI's default, ui/RenderThread.cpp:163), HaltonSampler 32 spp, 1024x1024"),
    # not BASELINE configs, measured to the same bar
    "w1": dict(scene="lights", args=(31, 4, WHITTED), w=1024, h=1024, spp=16, depth=5, integ=WHITTED,
               desc="W1: WhittedIntegrator, Cornell room + Mirror / Glass / Plastic spheres (11 532 tris), area + Point + Spot + Distant + SkyBox lights, 1024x1024, 16 spp, maxDepth 5"),
    "d1": dict(scene="lights", args=(31, 4, DIRECT), w=1024, h=1024, spp=16, depth=5, integ=DIRECT,
               desc="D1: DirectLightingIntegrator (UniformSampleOne), same scene as W1, 1024x1024, 16 spp, maxDepth 5"),
    "da1": dict(scene="lights", args=(31, 4, DIRECT_ALL), w=1024, h=1024, spp=16, depth=5, integ=DIRECT_ALL,
                desc="DA1: DirectLightingIntegrator (UniformSampleAll: every light at every vertex, 5 samples per area-light triangle), same scene as W1, 1024x1024, 16 spp, maxDepth 5"),
}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("hbm_gbs"), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_evidence(workload):
    """Figures that only a profiler can give, read from the committed ncu summaries of THIS command (never measured under
    the timer): per-launch DRAM bytes of the traversal launches and their lane / issue statistics."""
    out = {"traffic": None, "traffic_source": None}
    for name in ("r02_dram_trace_%s.csv" % workload, "r01_dram_trace.csv" if workload == "c2" else None):
        path = os.path.join(ROOT, "profiles", name) if name else None
        if not path or not os.path.exists(path):
            continue
        per_launch = {}
        try:
            with open(path) as f:
                rows = [r for r in csv.reader(l for l in f if l.startswith('"'))]
            hdr = rows[0]
            ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
            for r in rows[1:]:
                kn = r[ik].replace(", 0>", ">")  # k_trace<KIND, WIDE = false>
                if "k_trace<3>" in kn or "k_trace<4>" in kn or "k_trace<0>" in kn or "k_anyhit8<1>" in kn or "k_vol" in kn or "k_recursive" in kn:
                    if r[im] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                        per_launch[r[iid]] = per_launch.get(r[iid], 0.0) + float(r[iv].replace(",", ""))
            if per_launch:
                out["traffic"] = sum(per_launch.values()) / len(per_launch)
                out["traffic_source"] = "profiles/" + name + f" ({len(per_launch)} launches)"
                break
        except Exception:
            continue
    path = os.path.join(ROOT, "profiles", "r02_trace_metrics.json")
    if os.path.exists(path):
        try:
            out["issue"] = json.load(open(path)).get(workload)
        except Exception:
            pass
    return out


def host_cpu():
    model, phys = None, set()
    try:
        pid = None
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name") and model is None:
                model = line.split(":", 1)[1].strip()
            elif line.startswith("physical id"):
                pid = line.split(":", 1)[1].strip()
            elif line.startswith("core id"):
                phys.add((pid, line.split(":", 1)[1].strip()))
    except Exception:
        pass
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    return {"model": model, "physical_cores": len(phys) or None, "threads": threads}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def scene_args(wl):
    """(scene name, p0, p1, p2) for the scene kit and for the oracle harness (same names on both sides)."""
    p0, p1, p2 = wl["args"]
    if wl["scene"] in ("dragon",) or wl["scene"].startswith("dragon3d:"):
        p1, p2 = p1 or 2048, p2 or 213
    if wl["scene"] == "nano":   # the kit's default tessellation (the harness takes it literally)
        p1, p2 = p1 or 320, p2 or 64
    return wl["scene"], p0, p1, p2


def run_reference(args, wl):
    """Reference arm: the unmodified reference's own Render (oracle/_ref/libgnxref.so, built from /root/reference by
    oracle/Makefile; that library links nothing of this repo's product) on all host threads, printf no-op'ed
    (core/Integrator.cpp:143).  Each step is a bounded sample of the workload: the same scene and resolution at a reduced
    spp (per-sample cost is independent of spp, core/Integrator.cpp:274-291).  value = best of the timed steps
    (best of 3), OMP_PROC_BIND=spread OMP_PLACES=cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # before libgomp initialises (it is loaded with libgnxref.so)
    os.environ.setdefault("OMP_PROC_BIND", "spread")
    os.environ.setdefault("OMP_PLACES", "cores")
    cpu = host_cpu()
    os.environ["OMP_NUM_THREADS"] = str(cpu["threads"])  # all the host threads the process may use (torchrun sets it to 1)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _harness
    W, H, spp, depth, desc = wl["w"], wl["h"], wl["spp"], wl["depth"], wl["desc"]
    if not os.path.exists(_harness.REF_LIB):
        emit({"impl": "reference", "unavailable": "oracle/_ref/libgnxref.so was not built (no /root/reference at build time)"})
        return 0
    ref = _harness.Ref()
    cores = cpu["threads"]
    sample_spp = min(args.ref_spp, spp)
    lib = ref.lib
    scene, p0, p1, p2 = scene_args(wl)
    h = lib.gnxh_scene_create(scene.encode(), W, H, sample_spp, p0, p1, p2)
    err = lib.gnxh_scene_error(h).decode()
    if err:
        emit({"impl": "reference", "unavailable": err})
        return 0
    rs = _harness.RefScene(lib, h, W, H, sample_spp)
    times = []
    steps = max(args.steps, 1)
    for i in range(args.warmup + steps):
        _, sec = rs.render_reference(max_depth=depth, threads=cores)
        if i >= args.warmup:
            times.append(sec)
    t = min(times)
    paths = W * H * sample_spp
    value = paths / t / 1e6
    sample = (f"{W}x{H} x {sample_spp} spp of the {spp}-spp workload per step, best of {len(times)} steps (mean {sum(times) / len(times) * 1e3:.0f} ms), "
              f"{cores} OpenMP threads on {cpu['physical_cores']} physical cores ({cpu['model']}), OMP_PROC_BIND={os.environ['OMP_PROC_BIND']} "
              f"OMP_PLACES={os.environ['OMP_PLACES']}, reference timeConsume, -O2 -fopenmp, printf interposed")
    line = {"metric": METRIC, "value": value, "unit": "Mpaths/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": desc, "sample": sample, "bvh_build_s": lib.gnxh_scene_bvh_seconds(h), "host_cpu": cpu},
            "cpu_baseline": {"value": value, "unit": "Mpaths/s", "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)
    return 0


def cpu_baseline(wl, budget_s=15.0):
    """The reference on this box's host cores, bounded to ~budget_s seconds (rank 0, N = 1 only)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _harness
    if not os.path.exists(_harness.REF_LIB):
        return None, None
    W, H, spp, depth = wl["w"], wl["h"], wl["spp"], wl["depth"]
    ref = _harness.Ref()
    lib = ref.lib
    scene, p0, p1, p2 = scene_args(wl)
    t0 = time.time()
    h = lib.gnxh_scene_create(scene.encode(), W, H, 1, p0, p1, p2)
    if lib.gnxh_scene_error(h):
        return None, None
    rs = _harness.RefScene(lib, h, W, H, 1)
    cpu = host_cpu()
    cores = cpu["threads"]
    _, sec1 = rs.render_reference(max_depth=depth, threads=cores)           # 1 spp probe (also warms the light tables)
    n = int(max(1, min(spp, budget_s / max(sec1, 1e-3))))
    rs.close()
    h = lib.gnxh_scene_create(scene.encode(), W, H, n, p0, p1, p2)
    rs = _harness.RefScene(lib, h, W, H, n)
    img, sec = rs.render_reference(max_depth=depth, threads=cores)
    val = W * H * n / sec / 1e6
    info = {"value": val, "unit": "Mpaths/s", "cores": cores, "kind": "reference", "host_cpu": cpu,
            "sample": f"{W}x{H} x {n} spp of the {spp}-spp workload, one render, reference timeConsume {sec:.2f} s, "
                      f"-O2 -fopenmp, printf interposed, scene+BVH build {time.time() - t0 - sec - sec1:.1f} s untimed"}
    return info, (rs, img, n)


def bridge_e2e(wl, steps):
    """Wall time of gnx::CUDAPathIntegrator::Render itself — the call a reference app makes (ui/RenderThread.cpp:169-175) —
    on the reference's own pbr::Scene, FrameBuffer::fbuffer / ubuffer complete on return.  Needs the reference's classes,
    i.e. oracle/_ref/libgnxbridge.so; reported next to the C-ABI e2e, which needs nothing outside the product."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _harness
    if not (os.path.exists(_harness.REF_LIB) and os.path.exists(_harness.BRIDGE_LIB)):
        return None
    scene, p0, p1, p2 = scene_args(wl)
    ref = _harness.Ref()
    h = ref.lib.gnxh_scene_create(scene.encode(), wl["w"], wl["h"], wl["spp"], p0, p1, p2)
    if ref.lib.gnxh_scene_error(h):
        return None
    rs = _harness.RefScene(ref.lib, h, wl["w"], wl["h"], wl["spp"])
    try:
        sec = rs.time_cuda_render(steps, max_depth=wl["depth"])
    except Exception as e:  # noqa: BLE001
        return {"error": str(e)}
    finally:
        pass
    paths = wl["w"] * wl["h"] * wl["spp"]
    rs.close()
    return {"value": paths / sec / 1e6, "unit": "Mpaths/s", "ms_per_call": sec * 1e3,
            "call": "gnx::CUDAPathIntegrator::Render(const pbr::Scene&, double&) on the reference's own Scene / Camera / Sampler / FrameBuffer objects (oracle/_ref/libgnxbridge.so)"}


def run_ours(args, wl, name):
    import numpy as np
    import torch
    import torch.distributed as dist
    from gnxraytracer_b200.api import FILM_BOX, FILM_GAUSSIAN, Context, RenderParams, SceneKit

    W, H, spp, depth, desc, integ = wl["w"], wl["h"], wl["spp"], wl["depth"], wl["desc"], wl["integ"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    ctx = Context(local)
    if world > 1:
        # ONE job over all ranks, inside the library: rank 0's communicator id travels over torch.distributed (plumbing),
        # every rank attaches its context; from here on the renders are collective and the library deals out the samples
        box = [ctx.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ctx.comm_attach(world, rank, box[0])

    def load(wl_):
        t0 = time.time()
        sc, p0, p1, p2 = scene_args(wl_)
        sk_ = SceneKit(sc, wl_["w"], wl_["h"], wl_["spp"], p0, p1, p2)
        tb = time.time() - t0
        t0 = time.time()
        ctx.upload(sk_.desc)
        return sk_, tb, time.time() - t0

    sk, t_build, t_upload = load(wl)
    strong = bool(wl.get("strong"))
    gauss = args.film == "gaussian"
    # the whole job, identical on every rank: weak scaling = N times the per-GPU samples of every pixel
    job_spp = spp if strong else spp * world
    params = RenderParams.make(W, H, job_spp, max_depth=depth, integrator=integ, film=FILM_GAUSSIAN if gauss else FILM_BOX,
                               filter_radius=2.0 if gauss else 0.0, filter_alpha=2.0 if gauss else 0.0,
                               partition=1 if args.partition == "tiles" else 0)
    fb = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    host = torch.empty((H, W, 4), dtype=torch.float32).pin_memory()
    host_u8 = np.zeros((H, W, 4), np.uint8)
    host_f = np.zeros((H, W, 4), np.float32)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_device(p, steps, warmup):
        for _ in range(warmup):
            ctx.render_device(p, fb.data_ptr(), stream, want_stats=False)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            ctx.render_device(p, fb.data_ptr(), stream, want_stats=False)
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / steps

    # ---- device-resident throughput ("value") -----------------------------------------------------------
    clocks = ClockSampler(local)
    for _ in range(args.warmup):
        ctx.render_device(params, fb.data_ptr(), stream, want_stats=False)
    barrier()
    if rank == 0:
        clocks.start()
    ms_step = timed_device(params, args.steps, 0)
    clk = clocks.stop() if rank == 0 else None

    # ---- one instrumented step: per-stage CUDA events, ray / byte counters (this rank's share) ---------------------
    st = ctx.render_device(params, fb.data_ptr(), stream, want_stats=True)
    barrier()
    # ---- end to end: host buffers in and out, copies inside the timed region ---------------------------------------
    #   N = 1: gnx_render_framebuffer, the call the drop-in class makes — float running mean + tonemapped 8-bit image into
    #          (pageable) host arrays of the FrameBuffer's shape;  N > 1: gnx_render on the attached contexts, the image
    #          reduced onto rank 0 and copied to pinned host memory there (no host -> device bounce anywhere)
    def e2e_call():
        if world == 1:
            ctx.render_framebuffer(params, 1, host_f, host_u8)
        else:
            ctx.render_host_ptr(params, host.data_ptr() if rank == 0 else 0, want_stats=False)

    for _ in range(2):
        e2e_call()
    barrier()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_call()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - w0) * 1e3 / args.steps)
    # the plain float-RGBA call as well (round 1's e2e), for continuity
    e2e_plain_ms = None
    if world == 1:
        ctx.render_host_ptr(params, host.data_ptr(), want_stats=False)
        w0 = time.perf_counter()
        for _ in range(args.steps):
            ctx.render_host_ptr(params, host.data_ptr(), want_stats=False)
        e2e_plain_ms = (time.perf_counter() - w0) * 1e3 / args.steps

    hit_coverage = None
    if rank == 0:
        one = Context(local)  # (primary hits are a single-device parity hook)
        one.upload(sk.desc)
        hit_coverage = float(np.mean(one.primary_hits(RenderParams.make(W, H, 1, max_depth=depth, integrator=integ), 0) >= 0))
        one.close()

    # ---- second record: BASELINE config 5 as ONE fixed job over the N ranks (strong scaling) ------------------------
    strong_rec = None
    if args.strong_record and not strong and name == "c2":
        wl5 = WORKLOADS["c5"]
        sk5, _, _ = load(wl5)
        p5 = RenderParams.make(wl5["w"], wl5["h"], wl5["spp"], max_depth=wl5["depth"], integrator=wl5["integ"])
        fb5 = torch.zeros((wl5["h"], wl5["w"], 4), dtype=torch.float32, device="cuda")

        def timed5(steps, warmup):
            for _ in range(warmup):
                ctx.render_device(p5, fb5.data_ptr(), stream, want_stats=False)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                ctx.render_device(p5, fb5.data_ptr(), stream, want_stats=False)
            e1.record()
            barrier()
            return max_over_ranks(e0.elapsed_time(e1)) / steps
        ms5 = timed5(2, 1)
        paths5 = wl5["w"] * wl5["h"] * wl5["spp"]
        strong_rec = {"workload": wl5["desc"], "scaling": "strong", "n_gpus": world, "value": paths5 / ms5 / 1e3, "unit": "Mpaths/s",
                      "ms_per_step": ms5, "steps": 2, "warmup": 1,
                      "note": "the 1024 samples of every pixel are dealt to the ranks as contiguous ranges by the library (gnx_comm_attach), 133 MB ncclReduce to rank 0 inside the timed region; speed-up = this value / the N=1 line's"}
        del fb5
        sk5.close()

    if rank == 0:
        peak, peak_src = measured_peaks()
        paths_rank = int(st.paths)                # camera paths of this rank's share in the instrumented step
        total_paths = W * H * job_spp             # all ranks together
        value = total_paths / ms_step / 1e3       # Mpaths/s
        e2e_val = total_paths / e2e_ms / 1e3
        ext_ms = st.ms_extend / max(1, st.extend_launches)
        ext_bytes = st.extend_bytes / max(1, st.extend_launches)
        achieved = (ext_bytes / (ext_ms * 1e-3)) / 1e9 if ext_ms > 0 else None
        ev = ncu_evidence(name)
        vp_stages = None
        if integ == VOLPATH and sum(st.vp_items) > 0:
            # VolPath wavefront: one roofline line per stage.  ALGORITHMIC bytes: path state 100 B per load
            # or store, walk state 144 B per load or store, 48 B triangle + 136 B material + 132 B environment tables at a vertex,
            # 32 B per BVH node and 48 B per triangle of the stage's rays, 32 B of grid reads per tracking step
            rays = max(1, st.rays)
            trav = (32.0 * st.nodes_visited + 48.0 * st.tris_tested) / rays
            it, ms = list(st.vp_items), list(st.vp_ms)
            by = [it[0] * 200 + st.rays_extend * trav,
                  it[1] * (200 + 144 + 48 + 136 + 132),
                  it[2] * (200 + 288) + st.rays_shadow * trav,
                  it[3] * (200 + 288 + 48 + 136) + st.rays_mis * trav,
                  st.vp_track_steps * 32 + it[4] * 56]
            names = ["k_vp_logic<extend>", "k_vp_logic<vertex>", "k_vp_logic<shadow walk>", "k_vp_logic<MIS walk + continuation>", "k_vp_track"]
            vp_stages = [{"kernel": names[k], "ms": ms[k], "items": it[k], "algorithmic_bytes": by[k],
                          "achieved": (by[k] / (ms[k] * 1e-3) / 1e9) if ms[k] > 0 else None,
                          "frac": (by[k] / (ms[k] * 1e-3) / 1e9 / peak) if ms[k] > 0 else None} for k in range(5)]
            dom = max(range(5), key=lambda k: ms[k])
            achieved, ext_ms, ext_bytes = vp_stages[dom]["achieved"], ms[dom] / max(1, st.vp_rounds), by[dom] / max(1, st.vp_rounds)
        whitted_any = None
        if integ == WHITTED and st.ms_shadow > st.ms_extend:
            # staged first vertex (csrc/gnx_whitted.cuh): the any-hit launch over the vertices' shadow items is the largest stage
            any_bytes = 32.0 * (st.nodes_visited - st.extend_nodes) + 48.0 * (st.tris_tested - st.extend_tris) + 48.0 * (st.rays_shadow + st.rays_mis)
            ext_ms, ext_bytes = st.ms_shadow, any_bytes
            achieved = (any_bytes / (st.ms_shadow * 1e-3)) / 1e9
            whitted_any = "k_anyhit8<0> (+ k_whitted_sum): the shadow rays of the staged first vertices, one per light and vertex, on the compressed 8-wide tree (the step's largest stage)"
        kernel = {PATH: "k_trace<3|0> + k_anyhit8<1> (the extend stage: closest-hit extension launches on the two-child tree and, on a second stream next to them, the previous bounce's any-hit rays on the compressed 8-wide tree)",
                  VOLPATH: (vp_stages[dom]["kernel"] + " (the stage with the largest share of the step; every stage under `stages`)") if vp_stages else "k_volpath",
                  WHITTED: whitted_any or "k_trace<3> + k_recursive<0> (camera rays, then the per-lane recursion over the samples whose first vertex has specular lobes; the other first vertices are staged: k_whitted_vertex -> k_anyhit8 -> k_whitted_sum)",
                  }.get(integ, "k_recursive (DirectLighting, one launch per batch)")
        line = {
            "metric": METRIC, "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": desc + (", Gaussian film (radius 2, alpha 2)" if gauss else ""),
                       "hit_coverage": hit_coverage,
                       "hit_coverage_note": "fraction of the pixels whose camera ray (sample 0) hits a surface; the others read the environment and end",
                       "paths_per_step_per_gpu": paths_rank,
                       "l2_policy": f"inputs larger than L2 (126 MB): every step rewrites {paths_rank * 16 / 1e9:.2f} GB of per-sample radiance plus the path state and queues of the paths that hit something (buffers sized {paths_rank * 288 / 1e9:.1f} GB for the wavefront integrators) between launches",
                       "job": (f"ONE job of {job_spp} spp per pixel over {world} ranks, dealt out by the library ({args.partition}); ncclReduce to rank 0 queued behind each rank's film kernel, inside the timed region"
                               if world > 1 else f"samples [0, {spp})"),
                       "scene_build_s": round(t_build, 3), "bvh_build_s": round(sk.build_seconds, 3), "scene_upload_s": round(t_upload, 3),
                       "num_prims": sk.num_prims},
            "mrays_per_s": st.rays * (total_paths / max(1, st.paths)) / ms_step / 1e3,
            "rays_per_path": st.rays / max(1, st.paths),
            "stage_ms": {"raygen": st.ms_raygen, "extend": st.ms_extend, "shade": st.ms_shade, "shadow": st.ms_shadow, "film": st.ms_film,
                         "device_total": st.device_ms,
                         "note": "rank 0's share; extend = camera rays (k_trace<3>), then per bounce the extension rays (k_trace<0>) with the previous bounce's shadow / environment-MIS rays next to them (k_anyhit8<1> on a second stream; scenes without the second accumulator: any-hit rays in the shadow stage); shadow = the any-hit launch after the last bounce and the area-light MIS probes"},
            "roofline": {"bound": "latency+issue (dependent L1/L2 gathers); NOT hbm: see traffic", "bound_contract": "hbm",
                         "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": ev["traffic"],
                         "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum), " + str(ev["traffic_source"]), "peak_source": peak_src,
                         "launches_per_step": st.extend_launches, "avg_launch_ms": ext_ms,
                         "algorithmic_bytes_per_launch": ext_bytes,
                         "issue": ev.get("issue"), "stages": vp_stages,
                         "vp_rounds": int(st.vp_rounds) if vp_stages else None, "vp_track_steps": int(st.vp_track_steps) if vp_stages else None,
                         "note": "achieved = ALGORITHMIC bytes / launch time: 32 B x BVH nodes popped + 48 B x triangles tested + 48 B x rays (ray read + hit write), counted by the kernel itself.  The scene (97 MB of nodes + triangles for C2) is L2-resident, so real DRAM traffic (`traffic`) is an order of magnitude below the algorithmic bytes: frac is NOT a fraction of HBM bandwidth in use; the kernel is bound by instruction issue and L1 wavefronts of the per-lane node gathers (`issue`: active lanes per instruction, issue-slot use from the ncu capture under profiles/)"},
            "e2e": {"value": e2e_val, "unit": "Mpaths/s", "h2d_bytes_per_step": ctypes.sizeof(params),
                    "d2h_bytes_per_step": W * H * (16 + 4) if world == 1 else W * H * 16,
                    "ms_per_step": e2e_ms,
                    "call": ("gnx_render_framebuffer(): params in; the FrameBuffer's float running mean (16 B/pixel) and its tonemapped 8-bit image (4 B/pixel) out, unpacked into pageable host arrays of the FrameBuffer's layout — the call gnx::CUDAPathIntegrator::Render makes"
                             if world == 1 else "gnx_render() on contexts attached to one job: every rank renders its sample range, ncclReduce to rank 0, float RGBA copied to pinned host memory on rank 0 only"),
                    "note": "scene uploaded once (like the reference, which excludes scene build from timeConsume)"},
            "gpu_launches": int(st.kernel_launches) * args.steps,
            "clocks": clk,
        }
        if e2e_plain_ms:
            line["e2e"]["gnx_render_float_rgba_ms"] = e2e_plain_ms
        if strong_rec:
            line["strong_scaling"] = strong_rec
        if world == 1 and not args.no_cpu_baseline:
            info, extra = cpu_baseline(wl)
            if info:
                line["cpu_baseline"] = info
                rs, img_ref, n = extra
                # parity at the baseline's spp: same scene through the bridge-free scene kit
                one = Context(local)
                one.upload(sk.desc)
                img, _ = one.render(RenderParams.make(W, H, n, max_depth=depth, integrator=integ))
                one.close()
                sys.path.insert(0, os.path.join(ROOT, "tests"))
                import _harness
                if wl["scene"] == "nano":
                    # the kit paints a procedural colour map (no JPEG decoder), the reference side loads awesomeface.jpg: same
                    # geometry, materials and light, another texture — the image comparison for this config is the GPU parity
                    # test through the bridge (tests/test_gpu_fullsize.py), not this line
                    line["rel_mse_vs_cpu_ref"] = None
                    line["rel_mse_note"] = "not comparable: procedural texture (kit) vs awesomeface.jpg (reference side); see tests/test_gpu_fullsize.py[nano_full]"
                else:
                    line["rel_mse_vs_cpu_ref"] = _harness.rel_mse(img, img_ref)
                line["rel_mse_spp"] = n
                rs.close()
            if not args.no_bridge:
                br = bridge_e2e(wl, max(2, args.steps))
                if br:
                    line["e2e"]["bridge_render"] = br
                    if br.get("ms_per_call"):
                        line["e2e"]["bridge_over_c_abi"] = br["ms_per_call"] / e2e_ms
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--ref-spp", type=int, default=16, help="spp of one reference-arm step (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-bridge", action="store_true", help="skip the CUDAPathIntegrator::Render timing (needs oracle/_ref)")
    ap.add_argument("--no-strong-record", dest="strong_record", action="store_false", help="skip the config-5 strong-scaling record")
    ap.add_argument("--partition", default="samples", choices=["samples", "tiles"], help="how an N-rank job is dealt out (gnx_partition)")
    ap.add_argument("--mesh", default=None, help="a .3d mesh file (the reference's format) to render in place of C2's stand-in mesh")
    ap.add_argument("--film", default="box", choices=["box", "gaussian"], help="box = the reference's film (the headline); gaussian = GaussianFilter(2, 2)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    wl = dict(WORKLOADS[args.workload])
    if args.mesh:
        # the real asset when the user has it: C2 / U1 with the mesh read from a .3d file (shape/plyRead.h layout) by the scene kit's
        # reader (our arm) and by the reference's plyInfo (reference arm), placed as ui/ModelList.cpp:49-69 places dragon.3d
        if args.workload not in ("c2", "u1p", "u1w"):
            ap.error("--mesh replaces the mesh of workloads c2 / u1p / u1w")
        wl["scene"] = ("dragon3d:" if args.workload == "c2" else "ui3d:") + os.path.abspath(args.mesh)
        wl["desc"] = wl["desc"].replace("872448 tris (torus-knot stand-in for dragon.3d)", "from " + os.path.basename(args.mesh))
    if args.impl == "reference":
        return run_reference(args, wl)
    return run_ours(args, wl, args.workload)


if __name__ == "__main__":
    sys.exit(main())
