#!/usr/bin/env python
"""bench.py — BASELINE.json's metric (Mrays/s, plus Mpix·spp/s) on BASELINE configs[1]:
procedural Sponza-scale atrium, 262 144 triangles, 25 PBR materials, 1920x1080, max_bounces 8,
directional light + sky.

A "step" is one reference frame = one sample per pixel over the whole image (the reference renders
1 spp per frame and accumulates, path_tracing.cpp:231-246,463-480); 256 steps = the 256-spp config.

  python bench.py --gpus N --steps K --warmup W            our CUDA path (one process per GPU)
  python bench.py --impl reference ...                     the CPU oracle port of the reference path

Contract keys: value (device-resident whole-job throughput), e2e (through the public pass API with
host buffers), roofline (dominant kernel vs measured HBM peak), cpu_baseline, clocks, gpu_launches.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WIDTH, HEIGHT, BOUNCES, RAY_LENGTH = 1920, 1080, 8, 100.0
WORKLOAD = "atrium_262144tri_25mat_1920x1080_depth8_dirlight_sky"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--width", type=int, default=WIDTH)
    ap.add_argument("--height", type=int, default=HEIGHT)
    ap.add_argument("--bounces", type=int, default=BOUNCES)
    ap.add_argument("--accel", default="merged", choices=["merged", "two_level"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--device-only", action="store_true", help="only the device-resident timed loop (for ncu captures)")
    ap.add_argument("--cpu-tile-stride", type=int, default=0, help="oracle sample: every n-th 16x16 tile (0 = auto)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_sample(scene, accel_mode, width, height, bounces, tile_stride, frames, threads=0):
    """Times the CPU oracle (the port of the reference's path) on a bounded, spatially uniform sample of
    the same workload. Returns (rays/s, stats, description, cores)."""
    from bisemutum_engine_b200 import capi
    from oracle import oracle_py
    ctx = oracle_py.OracleContext(width, height, threads)
    ctx.upload_scene(scene, accel_mode)
    cam = oracle_py.camera_matrices(scene.camera, width, height)
    st = capi.Settings(ray_length=RAY_LENGTH, max_bounces=bounces)
    ctx.set_tile_sample(tile_stride, 0)
    ctx.render(cam, 1000, 1, st)          # warm-up (page in, spawn threads)
    ctx.reset_counters()
    t0 = time.perf_counter()
    ctx.render(cam, 0, frames, st)
    dt = time.perf_counter() - t0
    c, s = ctx.counters(), ctx.stats()
    rays = c.extend_rays + c.shadow_rays
    desc = f"every {tile_stride}th 16x16 tile of the {width}x{height} frame, {frames} spp, depth {bounces} ({c.samples} pixel-samples, {rays} rays)"
    out = dict(rays_per_s=rays / dt, seconds=dt, rays=rays, pixel_samples=c.samples, cores=ctx.threads, desc=desc,
               ext_nodes=s.extend_nodes / max(1, s.extend_rays), ext_tris=s.extend_tris / max(1, s.extend_rays),
               shd_nodes=s.shadow_nodes / max(1, s.shadow_rays), shd_tris=s.shadow_tris / max(1, s.shadow_rays),
               ext_inst=s.extend_instances / max(1, s.extend_rays), ext_wide_nodes=0.0, ext_leaf_boxes=0.0, ext_wide_ray_share=0.0)
    # Untimed second pass for the byte accounting: the same rays through the trees the CUDA kernels actually walk
    # (merged mode: binary tree at bounce 1, 4-wide quantised tree + exact leaf boxes from bounce WIDE_FROM on).
    wide_from = wide_from_bounce()
    if accel_mode == capi.ACCEL_MERGED and wide_from:
        ctx.set_wide_from_bounce(wide_from)
        ctx.reset_counters()
        ctx.render(cam, 0, max(1, min(frames, 4)), st)
        w = ctx.stats()
        n = max(1, w.extend_rays)
        out.update(ext_nodes=w.extend_nodes / n, ext_tris=w.extend_tris / n, ext_wide_nodes=w.extend_wide_nodes / n,
                   ext_leaf_boxes=w.extend_leaf_boxes / n, ext_wide_ray_share=w.extend_wide_rays / n)
    ctx.close()
    return out


def wide_from_bounce():
    """Mirror of use_wide() in csrc/render.cu: the first bounce whose rays walk the 4-wide tree (0 = never)."""
    if os.environ.get("BPT_WIDE", "1") == "0":
        return 0
    return int(os.environ.get("BPT_WIDE_FROM_BOUNCE", "2"))


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path cannot be built here
    (HLSL + Vulkan RT), so the arm times the oracle port with all host threads."""
    if rank != 0:
        return
    from bisemutum_engine_b200 import capi, scenes
    scene = scenes.atrium()
    mode = capi.ACCEL_MERGED if args.accel == "merged" else capi.ACCEL_TWO_LEVEL
    stride = args.cpu_tile_stride or 16
    from oracle import oracle_py
    ctx = oracle_py.OracleContext(args.width, args.height)
    ctx.upload_scene(scene, mode)
    cam = oracle_py.camera_matrices(scene.camera, args.width, args.height)
    st = capi.Settings(ray_length=RAY_LENGTH, max_bounces=args.bounces)
    ctx.set_tile_sample(stride, 0)
    for w in range(args.warmup):
        ctx.render(cam, 100000 + w, 1, st)
    ctx.reset_counters()
    t0 = time.perf_counter()
    for k in range(args.steps):
        ctx.render(cam, k, 1, st)
    dt = time.perf_counter() - t0
    c = ctx.counters()
    rays = c.extend_rays + c.shadow_rays
    value = rays / dt / 1e6
    sample = f"each step = 1 spp over every {stride}th 16x16 tile of the {args.width}x{args.height} frame ({c.samples // max(1, args.steps)} pixel-samples/step)"
    print(json.dumps({
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / max(1, args.steps) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": WORKLOAD, "accel": args.accel, "note": "CPU oracle port of the reference path; reference itself is not buildable (HLSL/DXC + Vulkan RT + window)"},
        "mpix_spp_per_s": c.samples / dt / 1e6,
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": ctx.threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import bisemutum_engine_b200 as pkg
    from bisemutum_engine_b200 import capi, engine, scenes, sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    W, H, B = args.width, args.height, args.bounces
    mode = capi.ACCEL_MERGED if args.accel == "merged" else capi.ACCEL_TWO_LEVEL
    scene = scenes.atrium()
    lib = pkg.load_library()

    # ---------------- device-resident arm: value -------------------------------------------------
    ctx = capi.Context(lib, W, H, device=local_rank)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    t0 = time.perf_counter()
    ctx.upload_scene(scene, mode)
    ctx.sync()
    build_ms = (time.perf_counter() - t0) * 1e3
    cam = engine.camera_matrices(scene.camera, W, H)
    st = capi.Settings(ray_length=RAY_LENGTH, max_bounces=B)
    if world > 1:
        sharding.comm_init(ctx, torch.device("cuda", local_rank))     # the library's own NCCL communicator (bpt_comm_init)
        ctx.reduce(0); ctx.sync()                                       # first collective sets the rings up, outside the timed region

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # sample s of the job ↔ frame_index = s; rank r renders the block [r*steps, (r+1)*steps)  
    ctx.render(cam, 1_000_000 + rank * args.warmup, args.warmup, st)
    ctx.sync()
    ctx.clear_accum()
    ctx.reset_counters()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    ctx.render(cam, sharding.first_frame(args.steps, rank, world), args.steps, st)     # K steps = K frames of 1 spp
    if world > 1:   # the one exchange step: FP32 sum buffers → rank 0 (one NCCL reduce per batch of frames, issued by the library)
        ctx.reduce(0)
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    c = ctx.counters()
    launches = c.kernel_launches
    stats = torch.tensor([ms, float(c.extend_rays), float(c.shadow_rays), float(c.samples)], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms = float(mx[0]); ext, shd, samples = float(sm[1]), float(sm[2]), float(sm[3])
    else:
        ext, shd, samples = float(c.extend_rays), float(c.shadow_rays), float(c.samples)
    value = (ext + shd) / (ms * 1e-3) / 1e6
    mpix = samples / (ms * 1e-3) / 1e6

    if args.device_only:
        if rank == 0:
            print(json.dumps({"metric": "Mrays/s", "value": value, "ms_per_step": ms / max(1, args.steps), "device_only": True}))
        return

    # ---------------- per-kernel timing of the dominant kernel (CUDA events inside libbpt) --------
    prof_steps = min(16, max(1, args.steps))
    ctx.profile_enable(True)
    ctx.reset_counters()
    ctx.render(cam, sharding.first_frame(args.steps, rank, world), prof_steps, st)
    kt = ctx.profile_read()
    ctx.profile_enable(False)
    pc = ctx.counters()

    # ---------------- e2e arm: through PathTracingPass with host buffers ---------------------------
    e2e = None
    if rank == 0 or world > 1:
        r = engine.Renderer(W, H, device=local_rank)
        r.ctx.set_stream(stream.cuda_stream)
        r.set_scene(scene, mode)
        host_img = torch.empty(H, W, 4, dtype=torch.float32).pin_memory()
        dev_img = torch.empty(H, W, 4, dtype=torch.float32, device="cuda")
        lights_bytes = scene.dir_lights.nbytes + scene.point_lights.nbytes + scene.rect_lights.nbytes
        h2d = lights_bytes + 3 * 64 + 32            # lights (update_params) + camera matrices + settings
        d2h = W * H * 16
        e2e_steps = max(8, min(args.steps, 128))
        prefetch = 32                                   # samples per wave of the pass (the library clamps it to its wave size)
        r.set_prefetch(prefetch)
        for i in range(3):
            r.ctx.upload_lights(scene); r.frame(RAY_LENGTH, B, True)
        r.ctx.sync(); r.ctx.reset_counters()
        barrier()
        # ring of pinned host buffers deep enough for one sample wave + slack: the pass prefetches a wave of samples,
        # so results arrive in bursts; the host must be able to queue the next wave while the previous read-backs drain
        ring = prefetch + 8
        host_imgs = [host_img] + [torch.empty(H, W, 4, dtype=torch.float32).pin_memory() for _ in range(ring - 1)]
        dev_imgs = [dev_img] + [torch.empty(H, W, 4, dtype=torch.float32, device="cuda") for _ in range(ring - 1)]
        copy_stream = torch.cuda.Stream()
        done = [None] * ring
        t0 = time.perf_counter()
        n = 0
        for i in range(e2e_steps):
            b = i % ring
            if done[b] is not None:
                done[b].synchronize()                       # the host buffer of frame i-ring has landed
            r.ctx.upload_lights(scene)                      # PathTracingPass::update_params: per-frame H2D of the light arrays
            n = r.frame(RAY_LENGTH, B, True)                # Camera::update_shader_params + PathTracingPass::render + RenderGraph::execute
            r.ctx.resolve_device(n, dev_imgs[b].data_ptr()) # the frame's result ...
            ready = torch.cuda.Event(); ready.record(stream)
            with torch.cuda.stream(copy_stream):            # ... read back to pinned host memory while the next frames render
                copy_stream.wait_event(ready)
                host_imgs[b].copy_(dev_imgs[b], non_blocking=True)
                done[b] = torch.cuda.Event(); done[b].record(copy_stream)
        for ev in done:
            if ev is not None:
                ev.synchronize()
        stream.synchronize()
        barrier()
        dt = time.perf_counter() - t0
        ec = r.ctx.counters()
        erays = torch.tensor([float(ec.extend_rays + ec.shadow_rays), dt], dtype=torch.float64, device="cuda")
        if world > 1:
            tot = erays.clone(); dist.all_reduce(tot, op=dist.ReduceOp.SUM)
            mxx = erays.clone(); dist.all_reduce(mxx, op=dist.ReduceOp.MAX)
            e_rays, e_dt = float(tot[0]), float(mxx[1])
        else:
            e_rays, e_dt = float(erays[0]), dt
        e2e = {"value": e_rays / e_dt / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "steps": e2e_steps, "ms_per_step": e_dt / e2e_steps * 1e3}
        r.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- CPU baseline + algorithmic bytes (oracle counters on the bit-identical BVH) --
    cpu = None
    peak, peak_src = peaks()
    roofline = None
    ora = None
    if not args.no_cpu_baseline:
        stride = args.cpu_tile_stride or 8
        ora = oracle_sample(scene, mode, W, H, B, stride, 1)
        if ora["seconds"] < 5.0:        # aim for ~10-30 s of CPU work
            frames = int(min(1024, max(2, 15.0 / max(ora["seconds"], 1e-3))))
            ora = oracle_sample(scene, mode, W, H, B, stride, frames)
        cpu = {"value": ora["rays_per_s"] / 1e6, "unit": "Mrays/s", "cores": ora["cores"], "kind": "port",
               "sample": ora["desc"], "seconds": ora["seconds"]}
    if kt.extend_launches:
        # extend ray = 32 B ray in + 16 B hit out + 64 B x nodes + 48 B x tris (oracle counters on the same rays);
        # rays that walk the 4-wide tree: 64 B per wide node + 32 B per exact leaf box instead of the binary nodes
        nodes, tris = (ora["ext_nodes"], ora["ext_tris"]) if ora else (None, None)
        if nodes is not None:
            per_ray = 32 + 16 + 64 * nodes + 64 * ora["ext_wide_nodes"] + 32 * ora["ext_leaf_boxes"] + 48 * tris
            rays_per_launch = pc.extend_rays / kt.extend_launches
            avg_s = kt.extend_ms * 1e-3 / kt.extend_launches
            achieved = per_ray * rays_per_launch / avg_s / 1e9
            total_k = kt.raygen_ms + kt.extend_ms + kt.shade_ms + kt.connect_ms + kt.other_ms
            traffic, traffic_src, traffic_detail = None, None, None
            tpath = os.path.join(ROOT, "profiles", "r1_extend_traffic.json")
            if os.path.exists(tpath) and (W, H, B, args.accel) == (WIDTH, HEIGHT, BOUNCES, "merged"):
                with open(tpath) as f:
                    tj = json.load(f)
                traffic, traffic_src = tj["traffic_bytes_per_launch"] / 1e9, tj["source"]     # GB per launch (dram read + write, ncu --set full)
                traffic_detail = {k: tj[k] for k in ("avg_launch_ms_under_ncu", "dram_gb_per_s", "capture") if k in tj}
                if "dram_gb_per_s" in traffic_detail:
                    traffic_detail["dram_frac_of_peak"] = traffic_detail["dram_gb_per_s"] / peak
            roofline = {"bound": "hbm", "kernel": "k_trace_spec<false,false> (extend)", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": traffic, "traffic_unit": "GB per launch", "traffic_source": traffic_src, "traffic_detail": traffic_detail,
                        "algorithmic_gb_per_launch": per_ray * rays_per_launch / 1e9,
                        "note": "the 29 MB scene+BVH is L2-resident (traffic << algorithmic bytes), so HBM does not bind this kernel: bounce 1 (binary tree) is issue-bound, later bounces (4-wide quantised tree) are L1/issue-bound; see profiles/r1_final_kernels.md (and r1_v8_kernels.md for the readings)",
                        "peak_source": peak_src, "bytes_per_ray": per_ray, "nodes_per_ray": nodes, "tris_per_ray": tris,
                        "wide_nodes_per_ray": ora["ext_wide_nodes"], "leaf_boxes_per_ray": ora["ext_leaf_boxes"], "wide_ray_share": ora["ext_wide_ray_share"],
                        "rays_per_launch": rays_per_launch, "avg_launch_ms": avg_s * 1e3,
                        "kernel_time_share": {"raygen": kt.raygen_ms / total_k, "extend": kt.extend_ms / total_k, "shade": kt.shade_ms / total_k,
                                              "connect": kt.connect_ms / total_k, "other": kt.other_ms / total_k}}
    kernel_ms = {"steps": prof_steps, "raygen": kt.raygen_ms / prof_steps, "extend": kt.extend_ms / prof_steps, "shade": kt.shade_ms / prof_steps,
                 "connect": kt.connect_ms / prof_steps, "other": kt.other_ms / prof_steps}
    out = {
        "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "width": W, "height": H, "max_bounces": B, "accel": args.accel, "triangles": scene.num_triangles,
                   "step": "1 spp over the full frame per GPU (rank r renders frames [r*steps, (r+1)*steps)); samples_per_wave = min(steps, 2^26 / pixels)",
                   "l2_policy": "inputs larger than L2: ~%d MB of per-path wavefront state streams through HBM every step; the %.0f MB scene+BVH is the steady-state L2-resident working set"
                   % (W * H * (6 * 16 + 20 + 16) // 1000000, (scene.num_triangles * (48 + 64)) / 1e6),
                   "multi_gpu": "sample-index sharding, scene+BVH replicated, one NCCL reduce of the FP32 sum buffer per batch" if world > 1 else "single GPU"},
        "mpix_spp_per_s": mpix, "extend_rays": ext, "shadow_rays": shd, "accel_build_ms": build_ms,
        "clocks": clocks, "gpu_launches": int(launches), "kernel_ms_per_step": kernel_ms, "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu,
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    # stdout carries exactly ONE line, the JSON: anything a library prints there while the bench runs (NCCL announces
    # "NCCL version ..." on stdout at communicator creation) is sent to stderr; the real stdout comes back for the print()s.
    sys.stdout.flush()
    _real_stdout = os.dup(1)
    os.dup2(2, 1)
    _print = print

    def print(*a, **k):                                   # noqa: A001 — the module's own prints go to the real stdout
        sys.stdout.flush()
        os.dup2(_real_stdout, 1)
        try:
            _print(*a, **k)
            sys.stdout.flush()
        finally:
            os.dup2(2, 1)
    main()
