#!/usr/bin/env python3
"""bench.py -- Mrays/s of the raytrace hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload config2|config3|config4|config5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        # the CPU arm: the oracle port of the reference

A step = one frame of the workload through softray_render_device with the scene resident in HBM
(`value`), and through softray_render with host buffers (`e2e`).  At N > 1 the SAME frame is split
into interleaved row bands (strong scaling); the bands land in rank 0's framebuffer through
peer-mapped stores issued by the render kernel itself (--gather peer) or an NCCL gather
(--gather nccl).  rays = primary + shadow + secondary, counted by the kernel.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Mrays/s (primary+shadow+secondary)"
UNIT = "Mrays/s"

# Algorithmic work model of SURVEY.md section 8(d): flops per unit (FMA = 2, everything else 1)
FLOPS = dict(primary=22, primary_aa=26, tri_test=41, tri_filter=11, sphere_test=34, node=40, phong=112, lambert=80,
             shadow_setup=6, reflect_setup=14, texture=12)
BYTES = dict(node=64, tri=128, tri_filter=64, sphere=48, pixel=4)   # what ONE visit / test / pixel has to move


def workload(name, scale=1.0):
    from softray_b200 import synth

    if name == "config1":
        # configs[0]: the reference's own test model through the native loader, 512x512, Lambert
        import numpy as np

        from softray_b200 import lib
        from softray_b200.scene import FrameParams

        fx = np.load(os.path.join(ROOT, "tests", "golden", "reference_fixtures.npz"))
        m = [lib.load_3ds(fx["model/obj.3ds"].tobytes())]
        s = None
        f = FrameParams(width=int(512 * scale), height=int(512 * scale), instances=[synth.camera(1.0)], background=synth.BACKGROUND,
                        shading=True, specular_lighting=False, shadows=False)
        desc = "configs[0]: Raytracer/obj.3DS (152 triangles) via the native 3DS loader, 512x512, 1 spp, primary rays + Lambert"
    elif name == "config2":
        m, s, f = synth.config2(width=int(1920 * scale), height=int(1080 * scale))
        desc = "configs[1]: procedural 1000-sphere scene in a 12-triangle room, 1920x1080, 1 spp, Phong + 100 soft-shadow rays per hit"
    elif name == "config2-hard":
        m, s, f = synth.config2(width=int(1920 * scale), height=int(1080 * scale), shadow_samples=1)
        desc = "configs[1] variant: 1000 spheres, 1920x1080, Phong + 1 shadow ray per hit"
    elif name == "config3":
        m, s, f = synth.config3(width=int(3840 * scale), height=int(2160 * scale))
        desc = "configs[2]: 1M-triangle height field, 3840x2160, 100 shadow rays + 2-bounce reflection + Texture3D"
    elif name == "config4":
        m, s, f = synth.config4(width=int(3840 * scale), height=int(2160 * scale))
        desc = "configs[3]: 100k-triangle mesh x 100 instances (10M), 3840x2160, 16 spp"
    elif name == "config5":
        m, s, f = synth.config5(width=int(7680 * scale), height=int(4320 * scale))
        desc = "configs[4]: 10M triangles flattened, 7680x4320, shading + 1 shadow ray per hit"
    else:
        raise SystemExit(f"unknown workload {name}")
    return m, s, f, desc


def algorithmic_work(st, frame):
    """(flops32, flops64, bytes) of one step from the kernel's counters.
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
                "samples": len(s), "window": "warm-up + timed steps"}


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference, all host threads, bounded sample of the same frame
# --------------------------------------------------------------------------------------------------
def cpu_sample(meshes, spheres, frame, seconds_target, threads=0):
    """Times the oracle on evenly spaced rows of the workload's frame (the reference's own
    rayTraceStartRow/EndRow idea, Renderer.cs:134-136, generalised to spaced rows).  Returns
    (Mrays/s, cores, sample description, seconds)."""
    import copy

    import oracle

    ncores = os.cpu_count() or 1
    opt = oracle.default_options(n_threads=threads or ncores)
    sc = oracle.Scene(meshes, spheres, options=opt)
    f = copy.copy(frame)
    H = frame.height

    def run(n_rows):
        n_rows = max(1, min(H, n_rows))
        f.band_height, f.band_count, f.band_index = 1, max(1, H // n_rows), 0
        t = time.perf_counter()
        out = sc.render(f)
        dt = time.perf_counter() - t
        return out["stats"].rays, dt, len(range(0, H, f.band_count))

    rays, dt, rows = run(max(2, ncores))                      # calibration: one row per thread
    n_rows = int(max(rows, min(H, rows * seconds_target / max(dt, 1e-3))))
    if n_rows > rows:
        rays, dt, rows = run(n_rows)
    sample = f"{rows} evenly spaced rows of {H} ({frame.width} px wide), {rays} rays, {dt:.1f} s"
    return rays / dt / 1e6, (threads or ncores), sample, dt


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The C# sources cannot be
    built here (no mono/dotnet), so this is the oracle port (kind "port"): same algorithm -- linear
    GeometryCollection scan over the spheres, SpatialSubdivision tree for the mesh, FP64."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    meshes, spheres, frame, desc = workload(args.workload, args.scale)
    per_step = max(1.0, min(30.0, 150.0 / (args.steps + args.warmup)))
    vals, secs, sample, cores = [], [], "", 0
    for i in range(args.warmup + args.steps):
        v, cores, sample, dt = cpu_sample(meshes, spheres, frame, per_step, args.cpu_threads)
        if i >= args.warmup:
            vals.append(v)
            secs.append(dt)
    value = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "description": desc, "width": frame.width, "height": frame.height},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config2")
    ap.add_argument("--scale", type=float, default=1.0, help="resolution scale (debug only; 1.0 = the named config)")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"])
    ap.add_argument("--band-height", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch

    from softray_b200 import abi, lib, multi_gpu

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    L = lib.load()
    ctx = lib.Context(local_rank)
    meshes, spheres, frame, desc = workload(args.workload, args.scale)
    W, H = frame.width, frame.height
    t0 = time.perf_counter()
    accel = {"bvh": abi.ACCEL_BVH, "lbvh": abi.ACCEL_LBVH}[os.environ.get("SOFTRAY_ACCEL", "bvh")]   # experiment knob
    scene = lib.Scene(ctx, meshes, spheres, accel=accel)
    scene_ms = (time.perf_counter() - t0) * 1e3

    bh = args.band_height or multi_gpu.default_band_height(H, world)
    my_rows = multi_gpu.apply_partition(frame, rank, world, bh)
    c_frame = frame.to_c(L.softray_instance_init)

    # a real (non-NULL) stream: the C ABI reads stream == NULL as "the context's own stream", and
    # the CUDA events below must be recorded on the stream the kernel is launched on
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    local_fb = torch.zeros((H, W), dtype=torch.int32, device="cuda")
    peer = None
    if world > 1 and args.gather == "peer":
        peer = multi_gpu.PeerFramebuffer(ctx, W, H)
    target_ptr = peer.ptr if peer is not None else local_fb.data_ptr()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def step_device():
        scene.render_device(frame, target_ptr, stream=stream.cuda_stream, c_frame=c_frame)
        if world > 1 and args.gather == "nccl":
            return multi_gpu.gather_frame(local_fb, my_rows, H, world, bh)
        return None

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # counters of this rank's share (one stats-enabled frame, also a first warm-up)
    st = scene.render_device(frame, target_ptr, stream=stream.cuda_stream, want_stats=True, c_frame=c_frame)
    counters = {k: getattr(st, k) for k in ("rays_primary", "rays_shadow", "rays_secondary", "node_visits", "prim_tests",
                                            "sphere_tests", "hits_primary", "shaded_hits", "filter_tests", "filter_unsure",
                                            "rays_bundled", "rays_fallback")}
    if dist is not None:
        t = torch.tensor([counters[k] for k in sorted(counters)], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        counters = dict(zip(sorted(counters), [int(v) for v in t.tolist()]))
    rays = counters["rays_primary"] + counters["rays_shadow"] + counters["rays_secondary"]

    # clocks are sampled from the first warm-up step to the end of the timed region: a config-2 frame takes
    # under a millisecond, so the timed region alone is shorter than a few NVML polls
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step_device()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    wall0 = time.perf_counter()
    for a, b in ev:
        flush.fill_(rank + 1)            # L2 flush between timed iterations (outside the events)
        if dist is not None:
            dist.barrier()               # a frame starts on all ranks together
        a.record(stream)
        step_device()
        b.record(stream)
    barrier()
    wall_ms = (time.perf_counter() - wall0) * 1e3
    clocks = sampler.finish()
    step_ms = torch.tensor([a.elapsed_time(b) for a, b in ev], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(step_ms, op=dist.ReduceOp.MAX)    # a frame is done when its slowest band is
    step_ms = step_ms.tolist()
    ms_per_step = sum(step_ms) / len(step_ms)
    value = rays / (ms_per_step * 1e-3) / 1e6

    # kernel-only time of the dominant (only) kernel, from the library's own events on the stream
    kst = scene.render_device(frame, target_ptr, stream=stream.cuda_stream, want_stats=True, c_frame=c_frame)
    kernel_ms = torch.tensor([kst.ms_kernel], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(kernel_ms, op=dist.ReduceOp.MAX)
    kernel_ms = float(kernel_ms.item())

    # ---- e2e: the C-ABI host-buffer call, pinned host framebuffer, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        host_px = torch.empty((H, W), dtype=torch.int32).pin_memory()
        hp = host_px.numpy().view(np.uint32)
        e2e_steps = max(1, args.steps)
        if world == 1:
            for _ in range(2):
                scene.render(frame, pixels=hp, want_stats=False)
            t = time.perf_counter()
            for _ in range(e2e_steps):
                scene.render(frame, pixels=hp, want_stats=False)
            e2e_ms = (time.perf_counter() - t) * 1e3 / e2e_steps
        else:
            # the caller's surface is one shared-memory section every rank process maps and page-locks: each rank's
            # softray_render stores its bands straight into it over its own GPU's PCIe link (multi_gpu "host" variant)
            shared = multi_gpu.SharedHostFramebuffer(W, H, ctx=ctx)

            def e2e_step():
                scene.render(frame, pixels=shared.pixels, want_stats=False)     # returns when this rank's bands are in host memory
                shared.barrier()                                                # ... and now everybody's are (softray_host_barrier)

            for _ in range(2):
                e2e_step()
            barrier()
            t = time.perf_counter()
            for _ in range(e2e_steps):
                e2e_step()
            e2e_ms = (time.perf_counter() - t) * 1e3 / e2e_steps
            tt = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e2e_ms = float(tt.item())
            e2e_ok = None
            if rank == 0:      # the assembled host surface equals the device-side gather of the timed region
                if peer is not None:
                    ref = _as_tensor(peer.ptr, H, W).cpu().numpy().view(np.uint32)
                    e2e_ok = bool(np.array_equal(ref, shared.pixels))
            shared.close()
        # frame constants uploaded per call: DevInstance records + the area-light offsets
        h2d = 288 * len(frame.instances) + (24 * frame.shadow_samples if frame.shadows else 0)
        e2e = {"value": rays / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": int(h2d) * world, "d2h_bytes_per_step": W * H * 4,
               "note": "softray_render (C ABI) with a pinned host framebuffer, which the kernel writes directly over PCIe (zero-copy stores: the D2H bytes leave the GPU while tracing continues); the scene is resident "
                       "(uploaded once by softray_scene_create, like the reference caches its geometry)"
                       + ("; N > 1: the host surface is a shared-memory section every rank maps and page-locks, each GPU writes its own "
                          "row bands into it over its own PCIe link, a frame ends with a barrier" if world > 1 else "")}
        if world > 1 and rank == 0:
            e2e["matches_device_gather"] = e2e_ok

    # ---- roofline of the render kernel: FP issue (branchy FP32 search + FP64 reference arithmetic; not
    # HBM-bound, not tensor work).  Mixed-precision rule of SURVEY 8d: FP32 work against the measured
    # FFMA peak, FP64 work against the measured DFMA peak; frac = share of the minimum possible issue time.
    flops32, flops64, nbytes = algorithmic_work(counters, frame)
    peak64 = ctx.measure_fma_peak(True)
    peak32 = ctx.measure_fma_peak(False)
    t_s = kernel_ms * 1e-3
    achieved = (flops32 + flops64) / t_s / 1e12
    t_min = (flops32 / (peak32 * 1e12) + flops64 / (peak64 * 1e12)) / world if peak32 > 0 and peak64 > 0 else 0.0
    eff_peak = (flops32 + flops64) / t_min / 1e12 if t_min > 0 else None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    traffic = None
    try:   # dram bytes of one launch of this workload from the committed ncu capture, if there is one
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload)
    except Exception:
        pass
    roofline = {
        "bound": "fp-issue", "achieved": achieved, "peak": eff_peak, "unit": "TFLOP/s",
        "frac": (t_min / t_s) if t_s > 0 else None, "traffic": traffic,
        "kernel": "sr::render_kernel", "kernel_ms": kernel_ms,
        "flops_fp32": flops32, "flops_fp64": flops64,
        "fp32_fma_peak_tflops": peak32 * world, "fp64_fma_peak_tflops": peak64 * world,
        "peak_source": "measured in this run by softray_measure_fma_peak (FFMA / DFMA chains, FMA = 2 flops); "
                       "MEASURED_PEAKS.json has no vector-FP figure.  peak = the flop-weighted mix of the two "
                       "(flops / minimum issue time), frac = minimum issue time / kernel time",
        "note": "algorithmic flops = SURVEY 8d constants x the kernel's own counters; a BVH walk is mostly "
                "min/max/compare/load issue slots, which this model does not credit (ncu issue-slot "
                "utilisation is in profiles/)",
        "hbm": {"achieved": nbytes / t_s / 1e9, "peak": hbm_peak * world, "unit": "GB/s",
                "frac": nbytes / t_s / 1e9 / (hbm_peak * world),
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback",
                "note": "algorithmic bytes of all node / primitive fetches; they are served by L1/L2 (the "
                        "scene is cache resident at this size), DRAM traffic is `traffic`"},
    }

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        frame_cpu = workload(args.workload, args.scale)[2]
        v, cores, sample, _ = cpu_sample(meshes, spheres, frame_cpu, args.cpu_seconds, args.cpu_threads)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "description": desc, "width": W, "height": H,
                       "rays_per_step": rays, "counters": counters, "l2": "flushed between timed steps (256 MB fill)",
                       "partition": (f"{world} ranks, interleaved bands of {bh} rows, gather={args.gather}" if world > 1
                                     else "single GPU"),
                       "scene_create_ms": scene_ms, "accel": os.environ.get("SOFTRAY_ACCEL", "bvh"), "wall_ms_timed_region": wall_ms},
            "clocks": clocks, "e2e": e2e, "gpu_launches": args.steps * world,
            "roofline": roofline, "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line), flush=True)
    if peer is not None:
        barrier()
        peer.close()
    scene.close()
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


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
