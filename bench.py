#!/usr/bin/env python3
"""bench.py -- Mrays/s and ms per frame of the raytrace hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload config5] [--others config3,config4,config2,config1]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        # the CPU arm: the oracle port of the reference

Headline workload at every N: configs[4] of BASELINE.json (8K frame, 10 M triangles, row bands over the GPUs) -- the
configuration the multi-GPU numbers are quoted on; the 4K configurations the metric's "ms per 4K frame" names
(configs[2], configs[3]) and configs[1], configs[0] are measured in the same run and reported under `others`, each
with its own ms per frame, counters and roofline.  A step = one frame through softray_render_device with the scene
resident in HBM (`value`), and through softray_render with a pinned host framebuffer (`e2e`).  At N > 1 the SAME
frame is split into interleaved row bands (strong scaling); the bands land in rank 0's framebuffer through
peer-mapped stores issued by the kernels themselves (--gather peer) or an NCCL gather (--gather nccl), and rank 0
also renders the whole frame alone to check the gathered frame against it (`matches_single_gpu`).
rays = primary + shadow + secondary as the reference counts them; `rays_traced` leaves out the shadow rays that a
cone test answered together (`rays_bundled`).
"""
import argparse
import copy
import json
import os
import shutil
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Mrays/s (primary+shadow+secondary)"
UNIT = "Mrays/s"
HEADLINE = "config5"
OTHERS = "config3,config4,config2,config1"

# Algorithmic work model of SURVEY.md section 8(d): flops per unit (FMA = 2, everything else 1)
FLOPS = dict(primary=22, primary_aa=26, tri_test=41, tri_filter=11, sphere_test=34, node=40, phong=112, lambert=80,
             shadow_setup=6, reflect_setup=14, texture=12)
BYTES = dict(node=64, tri=128, tri_filter=64, sphere=48, pixel=4)   # what ONE visit / test / pixel has to move
COUNTERS = ("rays_primary", "rays_shadow", "rays_secondary", "node_visits", "prim_tests", "sphere_tests", "hits_primary",
            "shaded_hits", "filter_tests", "filter_unsure", "rays_bundled", "rays_fallback", "rays_short_listed")

DESCRIPTIONS = {
    "config1": "configs[0]: Raytracer/obj.3DS (152 triangles) via the native 3DS loader, 512x512, 1 spp, primary rays + Lambert",
    "config2": "configs[1]: procedural 1000-sphere scene in a 12-triangle room, 1920x1080, 1 spp, Phong + 100 soft-shadow rays per hit",
    "config2-hard": "configs[1] variant: 1000 spheres, 1920x1080, Phong + 1 shadow ray per hit",
    "config3": "configs[2]: 1M-triangle height field, 3840x2160, 100 shadow rays per hit + 2-bounce reflection + Texture3D",
    "config4": "configs[3]: 100k-triangle mesh x 100 instances (10M), 3840x2160, 16 spp",
    "config5": "configs[4]: 10M triangles flattened, 7680x4320, Phong + 1 shadow ray per hit, row bands over the GPUs",
}
SIZES = {"config1": (512, 512), "config2": (1920, 1080), "config2-hard": (1920, 1080), "config3": (3840, 2160),
         "config4": (3840, 2160), "config5": (7680, 4320)}


def workload(name, scale=1.0):
    from softray_b200 import synth

    if name not in SIZES:
        raise SystemExit(f"unknown workload {name}")
    w, h = int(SIZES[name][0] * scale), int(SIZES[name][1] * scale)
    if name == "config1":
        import numpy as np

        from softray_b200 import lib
        from softray_b200.scene import FrameParams

        fx = np.load(os.path.join(ROOT, "tests", "golden", "reference_fixtures.npz"))
        m = [lib.load_3ds(fx["model/obj.3ds"].tobytes())]
        s = None
        f = FrameParams(width=w, height=h, instances=[synth.camera(1.0)], background=synth.BACKGROUND,
                        shading=True, specular_lighting=False, shadows=False)
    elif name == "config2":
        m, s, f = synth.config2(width=w, height=h)
    elif name == "config2-hard":
        m, s, f = synth.config2(width=w, height=h, shadow_samples=1)
    elif name == "config3":
        m, s, f = synth.config3(width=w, height=h)
    elif name == "config4":
        m, s, f = synth.config4(width=w, height=h)
    else:
        m, s, f = synth.config5(width=w, height=h)
    return m, s, f, DESCRIPTIONS[name]


def config_of(args, world):
    """The `config` object of the JSON line: what was asked for, nothing measured -- identical in both arms."""
    from softray_b200 import multi_gpu

    W, H = int(SIZES[args.workload][0] * args.scale), int(SIZES[args.workload][1] * args.scale)
    bh = args.band_height or multi_gpu.default_band_height(H, world)
    return {"workload": args.workload, "description": DESCRIPTIONS[args.workload], "width": W, "height": H, "scale": args.scale,
            "others": [o for o in args.others.split(",") if o and o != args.workload],
            "l2": "flushed between timed steps (256 MB fill)",
            "partition": (f"{world} ranks, interleaved bands of {bh} rows, gather={args.gather}" if world > 1 else "single GPU")}


def algorithmic_work(st, frame):
    """(flops32, flops64, bytes) of one step from the kernels' counters.
    FP32: BVH node visits and the filtered triangle tests (charged their back-face-reject cost: most end
    there).  FP64: the reference-arithmetic primitive tests, ray generation, shading, ray setup."""
    tri_tests = st["prim_tests"] - st["sphere_tests"]
    shade = (FLOPS["phong"] if frame.specular_lighting else FLOPS["lambert"]) if frame.shading else 0
    traced_shadow = st["rays_shadow"] - st.get("rays_bundled", 0)
    flops32 = st["node_visits"] * FLOPS["node"] + st.get("filter_tests", 0) * FLOPS["tri_filter"]
    flops64 = (st["rays_primary"] * (FLOPS["primary_aa"] if frame.sub_pixel_res > 1 else FLOPS["primary"])
               + tri_tests * FLOPS["tri_test"] + st["sphere_tests"] * FLOPS["sphere_test"]
               + st["shaded_hits"] * (shade + (FLOPS["texture"] if frame.texture3d_id else 0))
               + traced_shadow * FLOPS["shadow_setup"] + st["rays_secondary"] * FLOPS["reflect_setup"])
    nbytes = (st["node_visits"] * BYTES["node"] + tri_tests * BYTES["tri"] + st.get("filter_tests", 0) * BYTES["tri_filter"]
              + st["sphere_tests"] * BYTES["sphere"] + frame.width * frame.height * BYTES["pixel"])
    return float(flops32), float(flops64), float(nbytes)


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU during the timed region (B200_PROFILING.md)."""

    def __init__(self, index, period=0.0005):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def finish(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s), "window": "warm-up + timed steps of the headline workload"}


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference, bounded sample of the same frame
# --------------------------------------------------------------------------------------------------
def csharp_runtime():
    """BASELINE.md section 2 prefers the real C# reference under Mono/.NET.  Probed at run time; even where a
    runtime exists the reference SOURCES do not (/root/reference is not on the GPU box), so the CPU arm is the
    oracle port either way and says which runtime it saw."""
    for exe in ("dotnet", "mono", "mcs", "csc"):
        p = shutil.which(exe)
        if p:
            return p
    return None


class CpuArm:
    def __init__(self, meshes, spheres, frame):
        import oracle

        self.oracle = oracle
        self.frame = frame
        self.ncores = os.cpu_count() or 1
        self.meshes, self.spheres = meshes, spheres
        self.scenes = {}
        t = time.perf_counter()
        self.scene_for(self.ncores)
        self.build_s = time.perf_counter() - t

    def scene_for(self, threads):
        if threads not in self.scenes:
            self.scenes[threads] = self.oracle.Scene(self.meshes, self.spheres, options=self.oracle.default_options(n_threads=threads))
        return self.scenes[threads]

    def sample(self, seconds_target, threads):
        """Times the oracle on evenly spaced rows of the frame (the reference's own rayTraceStartRow/EndRow idea,
        Renderer.cs:134-136, generalised to spaced rows).  Rows are independent, so the frame time is the sample's
        time scaled by height / rows."""
        f = copy.copy(self.frame)
        H = f.height
        sc = self.scene_for(threads)

        def run(n_rows):
            n_rows = max(1, min(H, n_rows))
            f.band_height, f.band_count, f.band_index = 1, max(1, H // n_rows), 0
            t = time.perf_counter()
            out = sc.render(f)
            dt = time.perf_counter() - t
            return out["stats"].rays, dt, len(range(0, H, f.band_count))

        rays, dt, rows = run(max(2, threads))                      # calibration: one row per thread
        n_rows = int(max(rows, min(H, rows * seconds_target / max(dt, 1e-3))))
        if n_rows > rows:
            rays, dt, rows = run(n_rows)
        return {"value": rays / dt / 1e6, "rays": rays, "seconds": dt, "rows": rows, "frame_ms": 1e3 * dt * H / rows,
                "sample": f"{rows} evenly spaced rows of {H} ({f.width} px wide), {rays} rays, {dt:.1f} s on {threads} threads"}


def cpu_baseline_entry(arm, seconds, threads_all):
    """All host threads (the headline of the CPU arm) and the reference's default rayTraceConcurrency = 4
    (Renderer.cs:82)."""
    a = arm.sample(seconds, threads_all)
    b = arm.sample(max(2.0, seconds / 3), 4) if threads_all != 4 else a
    return {"value": a["value"], "unit": UNIT, "cores": threads_all, "kind": "port", "sample": a["sample"],
            "ms_per_frame_extrapolated": a["frame_ms"],
            "threads_4": {"value": b["value"], "unit": UNIT, "cores": 4, "sample": b["sample"], "ms_per_frame_extrapolated": b["frame_ms"],
                          "note": "the reference's default rayTraceConcurrency (Renderer.cs:82)"},
            "csharp_runtime_on_box": csharp_runtime(),
            "note": "oracle/softray_oracle.c: the C restatement of the reference's algorithm (kd-style SpatialSubdivision tree depth <= 15, "
                    "linear sphere list, FP64), not the C# runtime; tree built in %.1f s" % arm.build_s}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The C# sources cannot be built here
    (no mono/dotnet in the image, and /root/reference does not travel), so this is the oracle port (kind "port")."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    meshes, spheres, frame, _ = workload(args.workload, args.scale)
    arm = CpuArm(meshes, spheres, frame)
    threads = args.cpu_threads or arm.ncores
    per_step = max(1.0, min(30.0, 120.0 / (args.steps + args.warmup)))
    runs = []
    for i in range(args.warmup + args.steps):
        r = arm.sample(per_step, threads)
        if i >= args.warmup:
            runs.append(r)
    value = sum(r["value"] for r in runs) / len(runs)
    frame_ms = sum(r["frame_ms"] for r in runs) / len(runs)
    four = arm.sample(max(2.0, per_step / 3), 4) if threads != 4 else runs[-1]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": frame_ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(args, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": runs[-1]["sample"],
                         "ms_per_frame_extrapolated": frame_ms,
                         "threads_4": {"value": four["value"], "unit": UNIT, "cores": 4, "sample": four["sample"],
                                       "ms_per_frame_extrapolated": four["frame_ms"]},
                         "csharp_runtime_on_box": csharp_runtime()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "ms_per_step = one whole frame, extrapolated from the timed row sample (rows are independent); "
                "oracle tree built in %.1f s (not timed)" % arm.build_s,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
class Gpu:
    """Everything one rank keeps across workloads."""

    def __init__(self, args):
        import torch

        from softray_b200 import lib

        self.torch = torch
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist

            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist
        self.L = lib.load()
        self.ctx = lib.Context(self.local_rank)
        # a real (non-NULL) stream: the C ABI reads stream == NULL as "the context's own stream", and the CUDA events
        # below must be recorded on the stream the kernels are launched on
        self.stream = torch.cuda.Stream()
        torch.cuda.set_stream(self.stream)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2
        self.peak32 = self.ctx.measure_fma_peak(False)
        self.peak64 = self.ctx.measure_fma_peak(True)
        try:
            self.peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            self.peaks = {}
        try:
            self.traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        except Exception:
            self.traffic = {}

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def all_max(self, x):
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device="cuda")
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def all_sum_counters(self, st):
        c = {k: int(getattr(st, k)) for k in COUNTERS}
        if self.dist is not None:
            t = self.torch.tensor([c[k] for k in COUNTERS], dtype=self.torch.int64, device="cuda")
            self.dist.all_reduce(t)
            c = dict(zip(COUNTERS, [int(v) for v in t.tolist()]))
        return c


def roofline_of(g, counters, frame, kernel_ms, name, stages=None, cam_counters=None):
    """FP-issue roofline of the frame's kernels, and of the dominant stage kernel when the frame ran the stage
    pipeline.  Mixed-precision rule of SURVEY 8d: FP32 work against the measured FFMA peak, FP64 work against the
    measured DFMA peak; frac = minimum possible issue time / measured time."""
    world = g.world
    flops32, flops64, nbytes = algorithmic_work(counters, frame)
    t_s = kernel_ms * 1e-3
    t_min = (flops32 / (g.peak32 * 1e12) + flops64 / (g.peak64 * 1e12)) / world if g.peak32 > 0 and g.peak64 > 0 else 0.0
    hbm_peak = g.peaks.get("hbm_gbs", 6650.0)
    traffic = g.traffic.get(name)
    out = {
        "bound": "fp-issue", "achieved": (flops32 + flops64) / t_s / 1e12, "peak": ((flops32 + flops64) / t_min / 1e12) if t_min > 0 else None,
        "unit": "TFLOP/s", "frac": (t_min / t_s) if t_s > 0 else None,
        "traffic": traffic,
        "traffic_source": ("dram__bytes_read + dram__bytes_write of one frame from the committed ncu captures under profiles/ "
                           "(profiles/traffic.json; not measured in this run)") if traffic else None,
        "kernel": "all kernels of one frame (sr::render_kernel, or the sr_wave.cu stage kernels: see `stages_ms`)", "kernel_ms": kernel_ms,
        "flops_fp32": flops32, "flops_fp64": flops64,
        "fp32_fma_peak_tflops": g.peak32 * world, "fp64_fma_peak_tflops": g.peak64 * world,
        "peak_source": "measured in this run by softray_measure_fma_peak (FFMA / DFMA chains, FMA = 2 flops); MEASURED_PEAKS.json has no "
                       "vector-FP figure.  peak = the flop-weighted mix of the two (flops / minimum issue time), frac = minimum issue time / kernel time",
        "note": "algorithmic flops = SURVEY 8d constants x the kernels' own counters; a BVH walk is mostly min/max/compare/load issue "
                "slots, which this model does not credit (ncu issue-slot utilisation is in profiles/)",
        "l1l2_served_bytes": {"achieved": nbytes / t_s / 1e9, "unit": "GB/s",
                              "note": "algorithmic bytes of all node / primitive fetches per second; they are served by L1/L2, NOT an HBM figure "
                                      f"(measured HBM copy peak {hbm_peak * world:.0f} GB/s; DRAM traffic is `traffic`)"},
    }
    if stages:
        out["stages_ms"] = stages
        dom = max(stages, key=stages.get)
        out["dominant_kernel"] = {"stage": dom, "ms": stages[dom], "share_of_frame": stages[dom] / max(sum(stages.values()), 1e-9),
                                  "note": "per-stage CUDA-event times of one extra frame rendered with softray_frame.profile_stages "
                                          "(chunks serialised), max over ranks"}
        if cam_counters is not None and dom in ("search", "shadow"):
            # the camera search's own counters come from a stats frame with shadows and reflection switched off;
            # the shadow stage's are the rest
            if dom == "search":
                c32 = cam_counters["node_visits"] * FLOPS["node"] + cam_counters["filter_tests"] * FLOPS["tri_filter"]
                c64 = cam_counters["rays_primary"] * (FLOPS["primary_aa"] if frame.sub_pixel_res > 1 else FLOPS["primary"])
            else:
                c32 = ((counters["node_visits"] - cam_counters["node_visits"]) * FLOPS["node"]
                       + (counters["filter_tests"] - cam_counters["filter_tests"]) * FLOPS["tri_filter"])
                c64 = (counters["rays_shadow"] - counters["rays_bundled"]) * FLOPS["shadow_setup"]
            tk = stages[dom] * 1e-3
            tmin = (c32 / (g.peak32 * 1e12) + c64 / (g.peak64 * 1e12)) / world
            out["dominant_kernel"].update({"flops_fp32": float(c32), "flops_fp64": float(c64), "achieved": (c32 + c64) / tk / 1e12,
                                           "unit": "TFLOP/s", "frac": tmin / tk})
    return out


def measure(g, name, steps, warmup, want_e2e=True, sample_clocks=False, check_single=False):
    """One workload on this rank's share of the frame.  Returns a dict (complete on rank 0)."""
    import numpy as np

    from softray_b200 import abi, lib, multi_gpu

    torch, args, world, rank = g.torch, g.args, g.world, g.rank
    meshes, spheres, frame, desc = workload(name, args.scale)
    W, H = frame.width, frame.height
    t0 = time.perf_counter()
    accel = {"bvh": abi.ACCEL_BVH, "lbvh": abi.ACCEL_LBVH}[os.environ.get("SOFTRAY_ACCEL", "bvh")]   # experiment knob
    scene = lib.Scene(g.ctx, meshes, spheres, accel=accel)
    scene_ms = (time.perf_counter() - t0) * 1e3
    bh = args.band_height or multi_gpu.default_band_height(H, world)
    my_rows = multi_gpu.apply_partition(frame, rank, world, bh)
    c_frame = frame.to_c(g.L.softray_instance_init)
    stream = g.stream

    local_fb = torch.zeros((H, W), dtype=torch.int32, device="cuda")
    peer = None
    if world > 1 and args.gather == "peer":
        peer = multi_gpu.PeerFramebuffer(g.ctx, W, H)
    target_ptr = peer.ptr if peer is not None else local_fb.data_ptr()

    def step_device():
        scene.render_device(frame, target_ptr, stream=stream.cuda_stream, c_frame=c_frame)
        if world > 1 and args.gather == "nccl":
            return multi_gpu.gather_frame(local_fb, my_rows, H, world, bh)
        return None

    # counters of this rank's share (one stats-enabled frame, also a first warm-up)
    st = scene.render_device(frame, target_ptr, stream=stream.cuda_stream, want_stats=True, c_frame=c_frame)
    counters = g.all_sum_counters(st)
    launches_per_frame = int(st.launches)
    rays = counters["rays_primary"] + counters["rays_shadow"] + counters["rays_secondary"]

    sampler = ClockSampler(g.local_rank) if sample_clocks else None
    if sampler:
        sampler.start()
    for _ in range(warmup):
        step_device()
    g.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    wall0 = time.perf_counter()
    for a, b in ev:
        g.flush.fill_(rank + 1)          # L2 flush between timed iterations (outside the events)
        if g.dist is not None:
            g.dist.barrier()             # a frame starts on all ranks together
        a.record(stream)
        step_device()
        b.record(stream)
    g.barrier()
    wall_ms = (time.perf_counter() - wall0) * 1e3
    clocks = sampler.finish() if sampler else None
    step_ms = torch.tensor([a.elapsed_time(b) for a, b in ev], dtype=torch.float64, device="cuda")
    if g.dist is not None:
        g.dist.all_reduce(step_ms, op=g.dist.ReduceOp.MAX)    # a frame is done when its slowest band is
    step_ms = step_ms.tolist()
    ms_per_step = sum(step_ms) / len(step_ms)

    # kernel-only time of the frame's kernels, from the library's own events on the stream
    kst = scene.render_device(frame, target_ptr, stream=stream.cuda_stream, want_stats=True, c_frame=c_frame)
    kernel_ms = g.all_max(kst.ms_kernel)
    # per-stage times (stage-kernel pipeline only) and the camera stage's own counters
    stages, cam_counters = None, None
    if launches_per_frame > 1:
        pf = copy.copy(frame)
        pf.profile_stages = True
        pst = scene.render_device(pf, target_ptr, stream=stream.cuda_stream, want_stats=True)
        stages = {k: g.all_max(pst.ms_stage[i]) for i, k in enumerate(abi.STAGE_NAMES)}
        stages = {k: v for k, v in stages.items() if v > 0.0}
        cf = copy.copy(frame)
        cf.shadows, cf.reflection_depth = False, 0
        cam_counters = g.all_sum_counters(scene.render_device(cf, target_ptr, stream=stream.cuda_stream, want_stats=True))
        scene.render_device(frame, target_ptr, stream=stream.cuda_stream, c_frame=c_frame)     # the real frame back in the buffer
        torch.cuda.synchronize()

    # ---- N > 1: the gathered frame against the same frame rendered by rank 0 alone
    matches_single = None
    if world > 1 and check_single:
        g.barrier()
        gathered = step_device()
        g.barrier()
        if rank == 0:
            solo = copy.copy(frame)
            solo.band_height, solo.band_count, solo.band_index = 0, 1, 0
            solo_fb = torch.zeros((H, W), dtype=torch.int32, device="cuda")
            scene.render_device(solo, solo_fb.data_ptr(), stream=stream.cuda_stream)
            torch.cuda.synchronize()
            if peer is not None:
                gathered = _as_tensor(peer.ptr, H, W)
            matches_single = bool(torch.equal(gathered, solo_fb)) if gathered is not None else None
            del solo_fb
        g.barrier()

    # ---- e2e: the C-ABI host-buffer call, pinned host framebuffer, copies inside the timed region
    e2e = None
    if want_e2e:
        e2e_steps = max(1, steps)
        e2e_ok = None
        if world == 1:
            host_px = torch.empty((H, W), dtype=torch.int32).pin_memory()
            hp = host_px.numpy().view(np.uint32)
            for _ in range(2):
                scene.render(frame, pixels=hp, want_stats=False)
            t = time.perf_counter()
            for _ in range(e2e_steps):
                scene.render(frame, pixels=hp, want_stats=False)
            e2e_ms = (time.perf_counter() - t) * 1e3 / e2e_steps
            e2e_ok = bool(np.array_equal(hp, local_fb.cpu().numpy().view(np.uint32)))
            del host_px
        else:
            # the caller's surface is one shared-memory section every rank process maps and page-locks: each rank's
            # softray_render stores its bands straight into it over its own GPU's PCIe link (multi_gpu "host" variant)
            shared = multi_gpu.SharedHostFramebuffer(W, H, ctx=g.ctx)

            def e2e_step():
                scene.render(frame, pixels=shared.pixels, want_stats=False)     # returns when this rank's bands are in host memory
                shared.barrier()                                                # ... and now everybody's are (softray_host_barrier)

            for _ in range(2):
                e2e_step()
            g.barrier()
            t = time.perf_counter()
            for _ in range(e2e_steps):
                e2e_step()
            e2e_ms = g.all_max((time.perf_counter() - t) * 1e3 / e2e_steps)
            if rank == 0 and peer is not None:      # the assembled host surface equals the device-side gather
                ref = _as_tensor(peer.ptr, H, W).cpu().numpy().view(np.uint32)
                e2e_ok = bool(np.array_equal(ref, shared.pixels))
            shared.close()
        h2d = 320 * len(frame.instances) + (24 * frame.shadow_samples if frame.shadows else 0)
        e2e = {"value": rays / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": int(h2d) * world, "d2h_bytes_per_step": W * H * 4,
               "matches_device_frame": e2e_ok,
               "note": "softray_render (C ABI) with a pinned host framebuffer, which the kernels write directly over PCIe (zero-copy stores: the "
                       "D2H bytes leave the GPU while tracing continues); the scene is resident (uploaded once by softray_scene_create, like the "
                       "reference caches its geometry)"
                       + ("; N > 1: the host surface is a shared-memory section every rank maps and page-locks, each GPU writes its own row "
                          "bands into it over its own PCIe link, a frame ends with a barrier" if world > 1 else "")}

    res = {
        "workload": name, "description": desc, "width": W, "height": H,
        "ms_per_step": ms_per_step, "value": rays / (ms_per_step * 1e-3) / 1e6, "unit": UNIT,
        "rays_per_step": rays, "rays_traced": rays - counters["rays_bundled"],
        "value_traced": (rays - counters["rays_bundled"]) / (ms_per_step * 1e-3) / 1e6,
        "counters": counters, "launches_per_frame": launches_per_frame,
        "pipeline": "stage kernels (sr_wave.cu)" if launches_per_frame > 1 else "fused kernel (sr_render.cu)",
        "scene_create_ms": scene_ms, "wall_ms_timed_region": wall_ms, "e2e": e2e, "matches_single_gpu": matches_single,
        "roofline": roofline_of(g, counters, frame, kernel_ms, name, stages, cam_counters),
        "clocks": clocks, "meshes": meshes, "spheres": spheres,
    }
    if peer is not None:
        g.barrier()
        peer.close()
    scene.close()
    del local_fb
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=HEADLINE)
    ap.add_argument("--others", default=OTHERS, help="comma-separated workloads also measured and reported under `others` ('' = none)")
    ap.add_argument("--scale", type=float, default=1.0, help="resolution scale (debug only; 1.0 = the named config)")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"])
    ap.add_argument("--band-height", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
        return

    g = Gpu(args)
    head = measure(g, args.workload, args.steps, args.warmup, want_e2e=not args.no_e2e, sample_clocks=True, check_single=True)
    others = {}
    for name in [o for o in args.others.split(",") if o and o != args.workload]:
        r = measure(g, name, max(3, min(args.steps, 5)), 3, want_e2e=not args.no_e2e, check_single=g.world > 1)
        for k in ("meshes", "spheres", "clocks", "wall_ms_timed_region"):
            r.pop(k, None)
        others[name] = r

    cpu_baseline = None
    if g.rank == 0 and g.world == 1 and not args.no_cpu:
        frame_cpu = workload(args.workload, args.scale)[2]
        arm = CpuArm(head["meshes"], head["spheres"], frame_cpu)
        cpu_baseline = cpu_baseline_entry(arm, args.cpu_seconds, args.cpu_threads or arm.ncores)

    if g.rank == 0:
        line = {
            "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": g.world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": config_of(args, g.world),
            "measured": {k: head[k] for k in ("rays_per_step", "rays_traced", "value_traced", "counters", "launches_per_frame", "pipeline",
                                              "scene_create_ms", "wall_ms_timed_region", "matches_single_gpu")},
            "clocks": head["clocks"], "e2e": head["e2e"],
            "gpu_launches": args.steps * head["launches_per_frame"] * g.world,
            "roofline": head["roofline"], "cpu_baseline": cpu_baseline,
            "others": others,
            "ms_per_4k_frame": {k: others[k]["ms_per_step"] for k in ("config3", "config4") if k in others},
        }
        print(json.dumps(line), flush=True)
    g.ctx.close()
    if g.dist is not None:
        g.dist.destroy_process_group()


def _as_tensor(ptr, H, W):
    """View a raw device pointer (this rank's own allocation) as an [H, W] int32 torch tensor."""
    import torch

    class _Holder:
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (H, W), "typestr": "<i4", "data": (int(ptr), False), "version": 3}
    return torch.as_tensor(h, device="cuda")


if __name__ == "__main__":
    main()
