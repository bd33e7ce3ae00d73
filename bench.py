#!/usr/bin/env python
"""bench.py — headline benchmark of the path-tracing hot path.

Workload (BASELINE.json configs[1], "C2"): dragon-class mesh (872 448 triangles; torus-knot stand-in
for the stripped Resources/dragon.3d) with the UI's purple Plastic material under
InfiniteAreaLight(MonValley1000.hdr), 1024 x 1024, 64 spp, maxDepth 5, PathIntegrator + Halton.
One step = one full render = W*H*spp camera paths.

  python bench.py [--gpus N] [--steps K] [--warmup W]           our arm   (CUDA, libgnxrt.so)
  python bench.py --impl reference [...]                         reference arm (the UNMODIFIED reference's
                                                                  OpenMP PathIntegrator::Render on the host cores)

Under torchrun (N > 1) every rank renders the full per-GPU workload on its own slice of the Halton
sequence (weak scaling: N * 64 spp in total) and the partial framebuffers are sum-reduced to rank 0
with NCCL inside the timed region.  Prints ONE JSON line on rank 0.
"""
import argparse
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
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel (the six k_trace<3|4> launches of one
# C2 step: 1676 + 1491 + 941 + 519 + 251 + 82 MB), from the ncu pass of this command kept in profiles/r01_dram_trace.csv
NCU_TRAFFIC_BYTES_PER_LAUNCH = {"c2": 4959e6 / 6}
WORKLOADS = {
    # name: (scene, p0, p1, p2, width, height, spp, max_depth, description)
    "c2": ("dragon", 0, 0, 0, 1024, 1024, 64, 5,
           "C2: dragon-class mesh 872448 tris (torus-knot stand-in for dragon.3d), Plastic, InfiniteAreaLight MonValley1000.hdr, 1024x1024, 64 spp, maxDepth 5, PathIntegrator+Halton"),
    "c3": ("nano", 0, 0, 0, 1920, 1080, 128, 5,
           "C3: UV-mapped smooth-shaded mesh (~90k tris, stand-in for nanosuit), DisneyMaterial + ImageTexture, InfiniteAreaLight TropicalRuins1000.hdr, 1920x1080, 128 spp, maxDepth 5"),
    "c4": ("smoke", 0, 0, 0, 1024, 1024, 64, 5,
           "C4: VolPathIntegrator, GridDensityMedium density_render.70.volume inside HomogeneousMedium fog, Matte ground, MonValley env, 1024x1024, 64 spp, maxDepth 5, PCG32 stream sampler"),
    "c1": ("cornell", 0, 3, 0, 512, 512, 16, 5,
           "C1: Cornell box + 2 icospheres (Mirror, Glass), DiffuseAreaLight, 512x512, 16 spp, maxDepth 5"),
    # BASELINE config 5: ONE fixed job split across the ranks by sample range (strong scaling, "scaling": "strong")
    "c5": ("dragon", 0, 0, 0, 3840, 2160, 1024, 5,
           "C5: dragon-class mesh 872448 tris, Plastic, MonValley env, 3840x2160, 1024 spp in total, sample ranges dealt to the ranks, maxDepth 5"),
    # SURVEY 8f rank 1 (the UI's default integrator): not a BASELINE config, measured to the same bar
    "w1": ("lights", 31, 4, 2, 1024, 1024, 16, 5,
           "W1: WhittedIntegrator, Cornell room + Mirror / Glass / Plastic spheres (11 532 tris), area + Point + Spot + Distant + SkyBox lights, 1024x1024, 16 spp, maxDepth 5"),
    "d1": ("lights", 31, 4, 3, 1024, 1024, 16, 5,
           "D1: DirectLightingIntegrator (UniformSampleOne), same scene as W1, 1024x1024, 16 spp, maxDepth 5"),
    "da1": ("lights", 31, 4, 4, 1024, 1024, 16, 5,
            "DA1: DirectLightingIntegrator (UniformSampleAll: every light at every vertex, 5 samples per area-light triangle), same scene as W1, 1024x1024, 16 spp, maxDepth 5"),
}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("hbm_gbs"), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


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


def run_reference(args, wl):
    """Reference arm: the unmodified reference's PathIntegrator::Render (oracle/_ref, built from
    /root/reference by oracle/Makefile) on all host threads, printf no-op'ed (core/Integrator.cpp:143).
    Each step is a bounded sample of the workload: the same scene and resolution at a reduced spp
    (per-sample cost is independent of spp, core/Integrator.cpp:274-291)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _harness
    scene, p0, p1, p2, W, H, spp, depth, desc = wl
    if not os.path.exists(_harness.REF_LIB):
        emit({"impl": "reference", "unavailable": "oracle/_ref/libgnxref.so was not built (no /root/reference at build time)"})
        return 0
    ref = _harness.Ref()
    # all the host threads the process may use, whatever OMP_NUM_THREADS says (torchrun sets it to 1)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    sample_spp = args.ref_spp
    lib = ref.lib
    h = lib.gnxh_scene_create(scene.encode(), W, H, sample_spp, p0, p1 or (2048 if scene == "dragon" else 0), p2 or (213 if scene == "dragon" else 0))
    err = lib.gnxh_scene_error(h).decode()
    if err:
        emit({"impl": "reference", "unavailable": err})
        return 0
    rs = _harness.RefScene(lib, h, W, H, sample_spp)
    times = []
    for i in range(args.warmup + args.steps):
        _, sec = rs.render_reference(max_depth=depth, threads=cores)
        if i >= args.warmup:
            times.append(sec)
    t = sum(times) / len(times)
    paths = W * H * sample_spp
    value = paths / t / 1e6
    sample = f"{W}x{H} x {sample_spp} spp of the {spp}-spp workload per step, best-effort all {cores} threads, reference timeConsume, printf interposed"
    line = {"metric": METRIC, "value": value, "unit": "Mpaths/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": desc, "sample": sample, "bvh_build_s": lib.gnxh_scene_bvh_seconds(h)},
            "cpu_baseline": {"value": value, "unit": "Mpaths/s", "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)
    return 0


def cpu_baseline(wl, budget_s=15.0):
    """The reference on this box's host cores, bounded to ~budget_s seconds (rank 0, N = 1 only)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _harness
    import numpy as np
    if not os.path.exists(_harness.REF_LIB):
        return None, None
    scene, p0, p1, p2, W, H, spp, depth, _ = wl
    ref = _harness.Ref()
    lib = ref.lib
    full = (2048, 213) if scene == "dragon" else (p1, p2)
    t0 = time.time()
    h = lib.gnxh_scene_create(scene.encode(), W, H, 1, p0, p1 or full[0], p2 or full[1])
    if lib.gnxh_scene_error(h):
        return None, None
    rs = _harness.RefScene(lib, h, W, H, 1)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    _, sec1 = rs.render_reference(max_depth=depth, threads=cores)           # 1 spp probe (also warms the light tables)
    n = int(max(1, min(spp, budget_s / max(sec1, 1e-3))))
    rs.close()
    h = lib.gnxh_scene_create(scene.encode(), W, H, n, p0, p1 or full[0], p2 or full[1])
    rs = _harness.RefScene(lib, h, W, H, n)
    img, sec = rs.render_reference(max_depth=depth, threads=cores)
    val = W * H * n / sec / 1e6
    info = {"value": val, "unit": "Mpaths/s", "cores": cores, "kind": "reference",
            "sample": f"{W}x{H} x {n} spp of the {spp}-spp workload, one render, reference timeConsume {sec:.2f} s, "
                      f"-O2 -fopenmp, printf interposed, scene+BVH build {time.time() - t0 - sec - sec1:.1f} s untimed"}
    return info, (rs, img, n)


def run_ours(args, wl):
    import numpy as np
    import torch
    import torch.distributed as dist
    from gnxraytracer_b200.api import FILM_BOX, FILM_GAUSSIAN, FILM_GAUSSIAN_SUMS, Context, RenderParams, SceneKit
    from gnxraytracer_b200.dist import reduce_filtered_sums, reduce_framebuffer, sample_range, weak_sample_range

    scene, p0, p1, p2, W, H, spp, depth, desc = wl
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    ctx = Context(local)
    t0 = time.time()
    sk = SceneKit(scene, W, H, spp, p0, p1, p2)
    t_build = time.time() - t0
    t0 = time.time()
    ctx.upload(sk.desc)
    t_upload = time.time() - t0

    strong = args.workload == "c5"
    first, count = sample_range(spp, rank, world) if strong else weak_sample_range(spp, rank)
    # VolPathIntegrator for the participating-media config; "lights" carries its gnx_integrator in p2
    integ = 1 if scene == "smoke" else (p2 if scene == "lights" else 0)
    # --film gaussian: GaussianFilter(radius 2, alpha 2) reconstruction instead of the reference's box average; N > 1 ranks
    # exchange the unresolved sums (gnxraytracer_b200.dist.reduce_filtered_sums)
    gauss = args.film == "gaussian"
    film = (FILM_GAUSSIAN_SUMS if world > 1 else FILM_GAUSSIAN) if gauss else FILM_BOX
    params = RenderParams.make(W, H, count, max_depth=depth, first_sample=first, spp_normalize=spp if strong else spp * world, integrator=integ,
                               film=film, filter_radius=2.0 if gauss else 0.0, filter_alpha=2.0 if gauss else 0.0)
    fb = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    host = torch.empty((H, W, 4), dtype=torch.float32).pin_memory()

    def step_device(want_stats=False):
        st = ctx.render_device(params, fb.data_ptr(), stream, want_stats=want_stats)
        if world > 1:
            (reduce_filtered_sums if gauss else reduce_framebuffer)(fb, 0)
        return st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value") -----------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    clk = clocks.stop() if rank == 0 else None

    # ---- one instrumented step: per-stage CUDA events, ray / byte counters ----------------------------------
    st = step_device(want_stats=True)
    barrier()
    paths_step = W * H * count
    # ---- end to end: the call a user makes (gnx_render into HOST memory), D2H inside the timed region ------
    for _ in range(2):
        ctx.render_host_ptr(params, host.data_ptr(), want_stats=False)
    barrier()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.render_host_ptr(params, host.data_ptr(), want_stats=False)
        if world > 1:
            hb = host.cuda(non_blocking=True)
            (reduce_filtered_sums if gauss else reduce_framebuffer)(hb, 0)
    barrier()
    e2e_ms = (time.perf_counter() - w0) * 1e3 / args.steps
    t = torch.tensor([e2e_ms], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())

    if rank == 0:
        peak, peak_src = measured_peaks()
        total_paths = W * H * spp if strong else paths_step * world  # all ranks together
        value = total_paths / ms_step / 1e3  # Mpaths/s
        e2e_val = total_paths / e2e_ms / 1e3
        ext_ms = st.ms_extend / max(1, st.extend_launches)
        ext_bytes = st.extend_bytes / max(1, st.extend_launches)
        achieved = (ext_bytes / (ext_ms * 1e-3)) / 1e9 if ext_ms > 0 else None
        line = {
            "metric": METRIC, "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": desc + (", Gaussian film (radius 2, alpha 2)" if gauss else ""), "paths_per_step_per_gpu": paths_step, "l2_policy": f"inputs larger than L2 (126 MB): every step rewrites {paths_step * 16 / 1e9:.2f} GB of per-sample radiance plus the path state and queues of the paths that hit something (buffers sized {paths_step * 288 / 1e9:.1f} GB for the wavefront integrators) between launches",
                       "sample_range": (f"the {spp} samples of every pixel are dealt out as {world} contiguous ranges; NCCL sum-reduce to rank 0 inside the timed region" if strong else f"rank r renders Halton samples [{spp}r, {spp}r+{spp}) of every pixel; NCCL sum-reduce to rank 0 inside the timed region") if world > 1 else f"samples [0, {spp})",
                       "scene_build_s": round(t_build, 3), "bvh_build_s": round(sk.build_seconds, 3), "scene_upload_s": round(t_upload, 3),
                       "num_prims": sk.num_prims},
            "mrays_per_s": st.rays * (total_paths / max(1, st.paths)) / ms_step / 1e3,
            "rays_per_path": st.rays / st.paths,
            "stage_ms": {"raygen": st.ms_raygen, "extend": st.ms_extend, "shade": st.ms_shade, "shadow": st.ms_shadow, "film": st.ms_film,
                         "device_total": st.device_ms,
                         "note": "extend = the k_trace<3|4|0> launches: camera rays, then per bounce the extension rays TOGETHER WITH the previous bounce's shadow / environment-MIS rays (one mixed launch); shadow = the any-hit launch after the last bounce and the area-light MIS probes"},
            "roofline": {"bound": "hbm", "kernel": "k_trace<3|4|0> (closest-hit extension launches, incl. the any-hit rays they carry)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH.get(args.workload),
                         "traffic_unit": "bytes per launch (ncu dram read + write, profiles/r01_dram_trace.csv)", "peak_source": peak_src,
                         "launches_per_step": st.extend_launches, "avg_launch_ms": ext_ms,
                         "algorithmic_bytes_per_launch": ext_bytes,
                         "note": "algorithmic bytes = 32 B x BVH nodes popped + 48 B x triangles tested + 48 B x rays (ray read + hit write), counted by the kernel itself; the scene (97 MB of nodes+triangles) is L2-resident, so DRAM traffic is ~17x below the algorithmic bytes (no re-reads from HBM) and the kernel is bound by instruction issue and L1 wavefronts (profiles/README.md), not by HBM"},
            "e2e": {"value": e2e_val, "unit": "Mpaths/s", "h2d_bytes_per_step": ctypes.sizeof(params), "d2h_bytes_per_step": W * H * 16,
                    "ms_per_step": e2e_ms, "note": "gnx_render(): params in, float RGBA framebuffer out to pinned host memory; scene uploaded once (like the reference, which excludes scene build from timeConsume)"},
            "gpu_launches": int(st.kernel_launches) * args.steps,
            "clocks": clk,
        }
        if world == 1 and not args.no_cpu_baseline:
            info, extra = cpu_baseline(wl)
            if info:
                line["cpu_baseline"] = info
                rs, img_ref, n = extra
                # parity at the baseline's spp: same scene through the bridge-free scene kit
                img, _ = ctx.render(RenderParams.make(W, H, n, max_depth=depth, integrator=integ))
                sys.path.insert(0, os.path.join(ROOT, "tests"))
                import _harness
                line["rel_mse_vs_cpu_ref"] = _harness.rel_mse(img, img_ref)
                line["rel_mse_spp"] = n
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
    ap.add_argument("--mesh", default=None, help="a .3d mesh file (the reference's format) to render in place of C2's stand-in mesh")
    ap.add_argument("--film", default="box", choices=["box", "gaussian"], help="box = the reference's film (the headline); gaussian = GaussianFilter(2, 2)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    wl = WORKLOADS[args.workload]
    if args.mesh:
        # the real asset when the user has it: C2 with the mesh read from a .3d file (shape/plyRead.h layout) by the scene kit's
        # reader (our arm) and by the reference's plyInfo (reference arm), placed as ui/ModelList.cpp:49-69 places dragon.3d
        if args.workload != "c2":
            ap.error("--mesh replaces the mesh of workload c2")
        wl = ("dragon3d:" + os.path.abspath(args.mesh),) + wl[1:8] + (wl[8].replace("872448 tris (torus-knot stand-in for dragon.3d)", "from " + os.path.basename(args.mesh)),)
    if args.impl == "reference":
        return run_reference(args, wl)
    return run_ours(args, wl)


if __name__ == "__main__":
    sys.exit(main())
