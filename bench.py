#!/usr/bin/env python
"""bench.py — BASELINE.json's metric (Mrays/s, plus Mpix·spp/s).

Workloads (`--config`):
  atrium     BASELINE configs[1], the headline: procedural Sponza-scale atrium, 262 144 triangles, 25 PBR materials,
             1920x1080, max_bounces 8, directional light + sky, merged acceleration structure          (default)
  instanced  BASELINE configs[3]: 2 M-triangle mesh x 512 instances (two-level BVH), 3840x2160, max_bounces 8 — the
             DRAM-sized configuration (224 MB BLAS), where HBM can bind

A "step" is one reference frame = one sample per pixel over the whole image (the reference renders 1 spp per engine frame and
accumulates, path_tracing.cpp:231-246,463-480).

  python bench.py --gpus N --steps K --warmup W            our CUDA path (one process per GPU; torchrun for N > 1)
  python bench.py --impl reference ...                     the CPU oracle port of the reference path (rank 0 only)
  --scaling weak    every rank renders K frames (default)          --scaling strong   the K-frame job is split over the ranks

JSON keys: value (device-resident whole-job throughput), e2e (through the PathTracingPass host mirror with HOST buffers: per
frame the light arrays H2D from pinned memory and the frame's image D2H in the reference's rgba16_sfloat), roofline (dominant
kernel against the ceiling that binds it, both ceilings shown), cpu_baseline, clocks (NVML, sampled in-process during the
timed region), gpu_launches, reduce_check (N > 1).
"""
import argparse
import glob
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

CONFIGS = {
    "atrium": dict(width=1920, height=1080, bounces=8, ray_length=100.0, accel="merged",
                   workload="atrium_262144tri_25mat_1920x1080_depth8_dirlight_sky"),
    "instanced": dict(width=3840, height=2160, bounces=8, ray_length=1000.0, accel="two_level",
                      workload="instanced_2Mtri_x512_two_level_3840x2160_depth8_dirlight_sky"),
}
SM_COUNT, SCHEDULERS_PER_SM = 148, 4


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=128)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="atrium", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--bounces", type=int, default=0)
    ap.add_argument("--accel", default="", choices=["", "merged", "two_level"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-waves", default="", help="e2e arm: explicit wave schedule, e.g. 14,6 (must add up to the steps of a rank); default: tapered")
    ap.add_argument("--device-only", action="store_true", help="only the device-resident timed loop (for ncu captures)")
    ap.add_argument("--cpu-tile-stride", type=int, default=0, help="oracle sample: every n-th 16x16 tile (0 = auto)")
    a = ap.parse_args()
    cfg = dict(CONFIGS[a.config])
    a.width = a.width or cfg["width"]; a.height = a.height or cfg["height"]; a.bounces = a.bounces or cfg["bounces"]
    a.accel = a.accel or cfg["accel"]; a.ray_length = cfg["ray_length"]; a.workload = cfg["workload"]
    a.standard = (a.width, a.height, a.bounces, a.accel) == (cfg["width"], cfg["height"], cfg["bounces"], cfg["accel"])
    return a


def make_scene(args):
    from bisemutum_engine_b200 import scenes
    return scenes.atrium() if args.config == "atrium" else scenes.instanced()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            j = json.load(f)
        return float(j["hbm_gbs"]), float(j.get("sm_max_mhz", 1965.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1965.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and clock-event reasons DURING the timed region, polled in-process through NVML every ~2 ms (a 20-step
    timed region lasts ~35 ms: `nvidia-smi -lms 100` never saw it)."""
    PERIOD_S = 0.002

    def __init__(self, torch_device_index):
        self.samples, self.reasons, self.err = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            import torch
            self.nv = pynvml
            pynvml.nvmlInit()
            try:
                uuid = "GPU-" + str(torch.cuda.get_device_properties(torch_device_index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(torch_device_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:      # noqa: BLE001 — a bench without NVML still runs; the line says why there are no samples
            self.err = f"NVML unavailable: {e}"

    def _loop(self):
        nv = self.nv
        flags = ((nv.nvmlClocksEventReasonHwSlowdown, "hw_slowdown"), (nv.nvmlClocksEventReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                 (nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_thermal_slowdown"), (nv.nvmlClocksEventReasonSwPowerCap, "sw_power_cap"),
                 (nv.nvmlClocksEventReasonHwPowerBrakeSlowdown, "hw_power_brake_slowdown"))
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in flags:
                    if r & bit:
                        self.reasons.add(name)
            except Exception as e:      # noqa: BLE001
                self.err = str(e)
                return
            time.sleep(self.PERIOD_S)

    def start(self):
        if self.err is None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()

    def stop(self):
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=1.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": getattr(self, "max_mhz", None), "reasons": sorted(self.reasons), "samples": 0, "note": self.err or "no samples"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_min_mhz": float(min(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples), "period_ms": self.PERIOD_S * 1e3, "source": "NVML in-process"}


def wide_from_bounce():
    """Mirror of use_wide() in csrc/render.cu: the first bounce whose rays walk the 4-wide tree (0 = never)."""
    if os.environ.get("BPT_WIDE", "1") == "0":
        return 0
    return int(os.environ.get("BPT_WIDE_FROM_BOUNCE", "2"))


def oracle_sample(scene, accel_mode, args, tile_stride, frames, threads=0, count_bytes=True):
    """Times the CPU oracle (the port of the reference's path) on a bounded, spatially uniform sample of the same workload."""
    from bisemutum_engine_b200 import capi
    from oracle import oracle_py
    W, H, B = args.width, args.height, args.bounces
    ctx = oracle_py.OracleContext(W, H, threads)
    ctx.upload_scene(scene, accel_mode)
    cam = oracle_py.camera_matrices(scene.camera, W, H)
    st = capi.Settings(ray_length=args.ray_length, max_bounces=B)
    ctx.set_tile_sample(tile_stride, 0)
    ctx.render(cam, 1000, 1, st)          # warm-up (page in, spawn threads)
    ctx.reset_counters()
    t0 = time.perf_counter()
    ctx.render(cam, 0, frames, st)
    dt = time.perf_counter() - t0
    c, s = ctx.counters(), ctx.stats()
    rays = c.extend_rays + c.shadow_rays
    desc = f"every {tile_stride}th 16x16 tile of the {W}x{H} frame, {frames} spp, depth {B} ({c.samples} pixel-samples, {rays} rays)"
    out = dict(rays_per_s=rays / dt, seconds=dt, rays=rays, pixel_samples=c.samples, cores=ctx.threads, desc=desc,
               ext_nodes=s.extend_nodes / max(1, s.extend_rays), ext_tris=s.extend_tris / max(1, s.extend_rays),
               shd_nodes=s.shadow_nodes / max(1, s.shadow_rays), shd_tris=s.shadow_tris / max(1, s.shadow_rays),
               ext_inst=s.extend_instances / max(1, s.extend_rays), ext_wide_nodes=0.0, ext_leaf_boxes=0.0, ext_wide_ray_share=0.0)
    # Untimed second pass for the byte accounting: the same rays through the trees the CUDA kernels actually walk
    # (merged mode: binary tree at bounce 1, 4-wide quantised tree + exact leaf boxes from bounce WIDE_FROM on).
    wide_from = wide_from_bounce()
    if count_bytes and accel_mode == capi.ACCEL_MERGED and wide_from:
        ctx.set_wide_from_bounce(wide_from)
        ctx.reset_counters()
        ctx.render(cam, 0, max(1, min(frames, 4)), st)
        w = ctx.stats()
        n = max(1, w.extend_rays)
        out.update(ext_nodes=w.extend_nodes / n, ext_tris=w.extend_tris / n, ext_wide_nodes=w.extend_wide_nodes / n,
                   ext_leaf_boxes=w.extend_leaf_boxes / n, ext_wide_ray_share=w.extend_wide_rays / n)
    ctx.close()
    return out


def vulkan_probe():
    """BASELINE.md §4 / SURVEY §8d: "the reference's own Vulkan path-tracing pass is also shown if it runs headless on the box".
    Looks for what that would need and says why it is not runnable."""
    import ctypes.util
    import shutil
    loader = ctypes.util.find_library("vulkan")
    icds = sorted(sum((glob.glob(os.path.join(d, "*.json")) for d in ("/usr/share/vulkan/icd.d", "/etc/vulkan/icd.d", os.path.expanduser("~/.local/share/vulkan/icd.d"))), []))
    tools = {t: bool(shutil.which(t)) for t in ("xmake", "dxc", "vulkaninfo", "glslangValidator")}
    ref_binary = any(os.path.exists(p) for p in ("/root/reference/build", os.path.join(ROOT, "baseline", "_ref", "bisemutum")))
    reasons = []
    if not ref_binary:
        reasons.append("no reference binary: the engine builds with xmake + ~20 un-vendored packages and compiles its HLSL with DXC (tools present: "
                       + ", ".join(f"{k}={'yes' if v else 'no'}" for k, v in tools.items()) + ")")
    if not loader:
        reasons.append("no Vulkan loader (libvulkan.so.1)")
    if not icds:
        reasons.append("no Vulkan ICD manifest under /usr/share/vulkan/icd.d or /etc/vulkan/icd.d")
    reasons.append("the engine opens a GLFW window unconditionally (no headless mode) and needs VK_KHR_ray_tracing_pipeline; B200 exposes no RT cores or graphics queue")
    return {"attempted": True, "runnable": False, "vulkan_loader": loader, "icd_manifests": icds, "reason": "not runnable: " + "; ".join(reasons)}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path cannot be built here (HLSL + Vulkan RT),
    so the arm times the oracle port with all host threads, each step a bounded sample of the frame. Rank 0 only."""
    if rank != 0:
        return
    from bisemutum_engine_b200 import capi
    scene = make_scene(args)
    mode = capi.ACCEL_MERGED if args.accel == "merged" else capi.ACCEL_TWO_LEVEL
    stride = args.cpu_tile_stride or 16
    from oracle import oracle_py
    ctx = oracle_py.OracleContext(args.width, args.height)
    ctx.upload_scene(scene, mode)
    cam = oracle_py.camera_matrices(scene.camera, args.width, args.height)
    st = capi.Settings(ray_length=args.ray_length, max_bounces=args.bounces)
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
        "ms_per_step": dt / max(1, args.steps) * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": args.workload, "width": args.width, "height": args.height, "max_bounces": args.bounces, "accel": args.accel,
                   "triangles": scene.num_triangles,
                   "note": "CPU oracle port of the reference path (oracle/, all host threads); the reference itself is not buildable here (HLSL/DXC + Vulkan RT + window)"},
        "mpix_spp_per_s": c.samples / dt / 1e6,
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": ctx.threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "reference_vulkan_pt": vulkan_probe(),
    }))


def pinned_like(arr):
    """A pinned-host copy of a numpy structured array (the per-frame H2D source of the e2e arm)."""
    import torch
    raw = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
    t = torch.empty(max(raw.size, 1), dtype=torch.uint8).pin_memory()
    v = t.numpy()[: raw.size]
    v[:] = raw
    return t, v.view(arr.dtype).reshape(arr.shape)


def extend_counters(args):
    """Per-launch hardware counters of the extend kernel from the committed ncu capture of this round (profiles/), or None."""
    path = os.path.join(ROOT, "profiles", f"r2_extend_counters_{args.config}.json")
    if not (os.path.exists(path) and args.standard):
        return None
    with open(path) as f:
        return json.load(f)


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
    from bisemutum_engine_b200 import capi, engine, sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    W, H, B = args.width, args.height, args.bounces
    npx = W * H
    mode = capi.ACCEL_MERGED if args.accel == "merged" else capi.ACCEL_TWO_LEVEL
    scene = make_scene(args)
    lib = pkg.load_library()
    # frames per rank: weak = K each; strong = the K-frame job split (the first K % N ranks take one more)
    if args.scaling == "strong":
        my_first, my_steps = sharding.split_frames(args.steps, rank, world)
        if args.steps < world:
            raise SystemExit("bench.py: --scaling strong needs --steps >= --gpus")
    else:
        my_steps, my_first = args.steps, sharding.first_frame(args.steps, rank, world)
    total_steps = args.steps * (world if args.scaling == "weak" else 1)

    # ---------------- device-resident arm: value -------------------------------------------------
    ctx = capi.Context(lib, W, H, device=local_rank)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    t0 = time.perf_counter()
    ctx.upload_scene(scene, mode)
    ctx.sync()
    build_ms = (time.perf_counter() - t0) * 1e3
    cam = engine.camera_matrices(scene.camera, W, H)
    st = capi.Settings(ray_length=args.ray_length, max_bounces=B)
    if world > 1:
        sharding.comm_init(ctx, torch.device("cuda", local_rank))     # the library's own NCCL communicator (bpt_comm_init)
        ctx.reduce(0); ctx.sync()                                       # first collective sets the rings up, outside the timed region

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ctx.render(cam, 1_000_000 + rank * max(args.warmup, 3), max(args.warmup, 3), st)     # W >= 3 untimed steps
    ctx.sync()
    ctx.clear_accum()
    ctx.reset_counters()
    sampler = ClockSampler(local_rank)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.start()
    ev0.record(stream)
    ctx.render(cam, my_first, my_steps, st)         # the rank's frames of 1 spp each
    ev_mid = torch.cuda.Event(enable_timing=True)
    ev_mid.record(stream)
    if world > 1:   # the one exchange step: FP32 sum buffers → rank 0 (one NCCL reduce per batch of frames, issued by the library)
        ctx.reduce(0)
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    render_ms, exchange_ms = ev0.elapsed_time(ev_mid), ev_mid.elapsed_time(ev1)      # this rank's own split of the timed region
    c = ctx.counters()
    launches = c.kernel_launches
    stats = torch.tensor([ms, float(c.extend_rays), float(c.shadow_rays), float(c.samples), render_ms, exchange_ms], dtype=torch.float64, device="cuda")
    timed_split = None
    if world > 1:
        mx = stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        mn = stats.clone(); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        sm = stats.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms = float(mx[0]); ext, shd, samples = float(sm[1]), float(sm[2]), float(sm[3])
        # where the timed region went: the slowest and fastest rank's render time, and the exchange as each rank saw it (a rank that finishes
        # rendering early waits inside the reduce for the slowest one, so max exchange = skew + the collective itself)
        timed_split = {"render_ms_max": float(mx[4]), "render_ms_min": float(mn[4]), "exchange_ms_max": float(mx[5]), "exchange_ms_min": float(mn[5])}
    else:
        ext, shd, samples = float(c.extend_rays), float(c.shadow_rays), float(c.samples)
    assert samples == float(total_steps) * npx, (samples, total_steps, npx)          # every counted sample was traced inside the timed region
    value = (ext + shd) / (ms * 1e-3) / 1e6
    mpix = samples / (ms * 1e-3) / 1e6

    if args.device_only:
        if rank == 0:
            print(json.dumps({"metric": "Mrays/s", "value": value, "ms_per_step": ms / max(1, args.steps), "device_only": True, "steps": args.steps,
                              "extend_rays": ext, "shadow_rays": shd, "extend_rays_per_bounce": [int(x) for x in c.extend_rays_per_bounce],
                              "shadow_rays_per_bounce": [int(x) for x in c.shadow_rays_per_bounce]}))
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- N > 1: is the reduced buffer the sum of the ranks' buffers? (untimed) ---------------------------------
    reduce_check = None
    if world > 1:
        chk = min(4, my_steps)
        ctx.clear_accum()
        ctx.render(cam, my_first, chk, st)
        mine = torch.empty(H, W, 4, dtype=torch.float32, device="cuda")
        ctx.resolve_device(1, mine.data_ptr())                          # total = 1: the raw sums (alpha forced to 1)
        expect = mine[..., :3].double().contiguous()
        dist.all_reduce(expect, op=dist.ReduceOp.SUM)                   # torch's own NCCL all-reduce in FP64: the independent answer
        ctx.reduce(0)
        got = torch.empty(H, W, 4, dtype=torch.float32, device="cuda")
        ctx.resolve_device(1, got.data_ptr())
        torch.cuda.synchronize()
        if rank == 0:
            g = got[..., :3].double()
            scale = float(expect.abs().max().clamp_min(1e-30))
            sum_rel = abs(float(g.sum()) - float(expect.sum())) / max(abs(float(expect.sum())), 1e-30)
            pix_rel = float((g - expect).abs().max()) / scale
            reduce_check = {"ok": bool(sum_rel <= 1e-6 and pix_rel <= 1e-6), "sum_rel_err": sum_rel, "max_pixel_err_rel_to_max": pix_rel,
                            "frames_per_rank": chk, "what": "rank-0 buffer after bpt_reduce vs FP64 all-reduce of the per-rank sum buffers"}
            assert reduce_check["ok"], reduce_check
        ctx.clear_accum()

    # ---------------- per-kernel timing of the dominant kernel (CUDA events inside libbpt, on its launch stream) --------
    prof_steps = min(32, max(1, my_steps))
    ctx.profile_enable(True)
    ctx.reset_counters()
    ctx.render(cam, my_first, prof_steps, st)
    kt = ctx.profile_read()
    ctx.profile_enable(False)
    pc = ctx.counters()
    ctx.clear_accum()

    # ---------------- e2e arm: through PathTracingPass (host mirror) with HOST buffers -------------------------------------------
    # Per step (= engine frame): the light arrays go H2D from pinned memory (PathTracingPass::update_params), the camera matrices are
    # recomputed on the host (Camera::update_shader_params) and travel as kernel parameters, the pass renders (it traces a wave of
    # `prefetch` samples when it has none left and folds one into the history per frame), and the frame's image is resolved to the
    # format of the reference's OutputData.color (rgba16_sfloat, path_tracing.cpp:248-252) and read back to pinned host memory —
    # on rank 0 only when N > 1: the other ranks' partial sums are not a result; the job's result is the ONE reduce at the end,
    # read back in FP32 on rank 0. Everything the timed frames consume is traced inside the timed region: the history is reset after
    # the warm-up frames (which drops their prefetched samples) and the waves of the timed frames add up to the step count exactly.
    def run_e2e(n_steps):
        r = engine.Renderer(W, H, device=local_rank)
        r.ctx.set_stream(stream.cuda_stream)
        r.set_scene(scene, mode)
        if world > 1:
            sharding.comm_init(r.ctx, torch.device("cuda", local_rank))
            r.ctx.reduce(0); r.ctx.sync()
        wave = max(1, min(32, (1 << 26) // npx))        # what the library traces per wave at this resolution (render.cu: wave_slots)
        # Wave schedule of a FINITE job. The frames of a wave become available together when its last bounce ends, and their images then
        # leave over PCIe one after the other (0.30 ms each at 1080p) while the NEXT wave is traced; only the last wave's read-back has
        # nothing to hide behind. So the application (which knows the job's length; the pass does not) asks for waves that shrink towards
        # the end — each 4/5 of what is left, at most the library's wave — instead of one full wave whose read-back is all tail.
        # Every sample is still traced inside the timed region and consumed by exactly one frame (asserted below).
        schedule = sharding.wave_schedule(n_steps, wave)
        if args.e2e_waves and sum(int(x) for x in args.e2e_waves.split(",")) == n_steps:
            schedule = [int(x) for x in args.e2e_waves.split(",")]
        e2e_steps = n_steps
        prefetch = max(schedule)
        r.set_prefetch(schedule[0])
        job_steps = torch.tensor([float(e2e_steps)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(job_steps, op=dist.ReduceOp.SUM)
        job_steps = int(job_steps.item())                # frames in the reduced image (all ranks)
        pin = [pinned_like(a) for a in (scene.dir_lights, scene.point_lights, scene.rect_lights)]
        import copy
        lights = copy.copy(scene)
        lights.dir_lights, lights.point_lights, lights.rect_lights = (p[1] for p in pin)
        lights_bytes = sum(p[1].nbytes for p in pin)
        for i in range(max(args.warmup, 3)):
            r.ctx.upload_lights(lights); r.frame(args.ray_length, B, True)
        r.reset_history()                               # the timed frames start a new accumulation: nothing traced ahead survives
        r.set_frame(my_first)                           # this rank's block of frame indices
        r.ctx.sync(); r.ctx.reset_counters()
        reads = rank == 0
        ring = prefetch + 4
        host_imgs = [torch.empty(H, W, 4, dtype=torch.float16).pin_memory() for _ in range(ring)] if reads else []
        dev_imgs = [torch.empty(H, W, 4, dtype=torch.float16, device="cuda") for _ in range(ring)] if reads else []
        final_host = torch.empty(H, W, 4, dtype=torch.float32).pin_memory() if reads else None
        final_dev = torch.empty(H, W, 4, dtype=torch.float32, device="cuda") if reads else None
        copy_stream = torch.cuda.Stream()
        done = [None] * ring
        if reads and os.environ.get("BENCH_E2E_PRETOUCH", "1") == "1":
            # an application reuses its result buffers from job to job: map them for DMA once, outside the timed region (the first copy
            # into a freshly pinned 33 MB buffer was measured at 8-30 GB/s instead of 55)
            final_host.copy_(final_dev, non_blocking=True)
            for hb, db in zip(host_imgs, dev_imgs):
                hb.copy_(db, non_blocking=True)
            torch.cuda.synchronize()
        trace = os.environ.get("BENCH_E2E_TRACE") == "1" and reads     # diagnostic: where the timed region goes (GPU events + host clock per frame)
        tr_ready, tr_done, tr_host = [], [], []
        barrier()
        t0 = time.perf_counter()
        if trace:
            tr_start = torch.cuda.Event(enable_timing=True); tr_start.record(stream)
        n = 0
        wave_i, wave_left = 0, 0
        for i in range(e2e_steps):
            if wave_left == 0:                              # this frame finds no sample traced ahead: the pass traces the next wave now
                r.set_prefetch(schedule[wave_i])
                wave_left = schedule[wave_i]; wave_i += 1
            wave_left -= 1
            r.ctx.upload_lights(lights)                     # PathTracingPass::update_params: per-frame H2D of the light arrays
            if reads:
                b = i % ring
                if done[b] is not None:
                    done[b].synchronize()                   # the host buffer of frame i - ring has landed
                r.set_color_target(dev_imgs[b].data_ptr())  # the device memory behind this frame's OutputData.color
            n = r.frame(args.ray_length, B, True)           # Camera::update_shader_params + PathTracingPass::render + RenderGraph::execute
            if reads:                                       # OutputData.color of this frame (written by the pass) ...
                ready = torch.cuda.Event(enable_timing=trace); ready.record(stream)
                with torch.cuda.stream(copy_stream):        # ... read back to pinned host memory while the next frames render
                    copy_stream.wait_event(ready)
                    host_imgs[b].copy_(dev_imgs[b], non_blocking=True)
                    done[b] = torch.cuda.Event(enable_timing=trace); done[b].record(copy_stream)
                if trace:
                    tr_ready.append(ready); tr_done.append(done[b]); tr_host.append((time.perf_counter() - t0) * 1e3)
        assert n == e2e_steps, (n, e2e_steps)
        if world > 1:
            r.ctx.reduce(0)                                 # the job's one exchange: every rank's sums -> rank 0
        if reads:
            r.ctx.resolve_device(job_steps, final_dev.data_ptr())
            if trace:
                tr_res = torch.cuda.Event(enable_timing=True); tr_res.record(stream)
            if os.environ.get("BENCH_E2E_FINAL_ON_COPY_STREAM") == "1":
                fin = torch.cuda.Event(); fin.record(stream)
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(fin)
                    final_host.copy_(final_dev, non_blocking=True)
                    fdone = torch.cuda.Event(); fdone.record(copy_stream)
                stream.wait_event(fdone)
            else:
                final_host.copy_(final_dev, non_blocking=True)  # the job's result, FP32, on the host
        if trace:
            tr_end = torch.cuda.Event(enable_timing=True); tr_end.record(stream); th0 = (time.perf_counter() - t0) * 1e3
        for ev in done:
            if ev is not None:
                ev.synchronize()
        if trace:
            th1 = (time.perf_counter() - t0) * 1e3
        stream.synchronize()
        if trace:
            th2 = (time.perf_counter() - t0) * 1e3
        barrier()
        dt = time.perf_counter() - t0
        if trace:
            sys.stderr.write("e2e tail: host after issue %.3f, after copy events %.3f, after stream sync %.3f, after barrier %.3f; GPU end of the FP32 resolve %.3f, of the FP32 copy %.3f\n"
                             % (th0, th1, th2, dt * 1e3, tr_start.elapsed_time(tr_res), tr_start.elapsed_time(tr_end)))
        if trace:
            sys.stderr.write("e2e trace (ms since start): total %.3f\n" % (dt * 1e3))
            for i in range(len(tr_ready)):
                sys.stderr.write("  frame %3d host-issued %.3f resolved %.3f copied %.3f\n" % (i, tr_host[i], tr_start.elapsed_time(tr_ready[i]), tr_start.elapsed_time(tr_done[i])))
        ec = r.ctx.counters()
        assert ec.samples == e2e_steps * npx, (ec.samples, e2e_steps, npx)       # traced inside the timed region == consumed
        erays = torch.tensor([float(ec.extend_rays + ec.shadow_rays), dt, float(ec.samples)], dtype=torch.float64, device="cuda")
        if world > 1:
            tot = erays.clone(); dist.all_reduce(tot, op=dist.ReduceOp.SUM)
            mxx = erays.clone(); dist.all_reduce(mxx, op=dist.ReduceOp.MAX)
            e_rays, e_dt, e_samples = float(tot[0]), float(mxx[1]), float(tot[2])
        else:
            e_rays, e_dt, e_samples = float(erays[0]), dt, float(erays[2])
        if reads:
            img = final_host.numpy()
            assert np.isfinite(img).all() and float(img[..., :3].mean()) > 0.0, "e2e: the image read back is empty"
            last = host_imgs[(e2e_steps - 1) % ring].float().numpy()
            if world == 1:      # the last per-frame image IS the final image, in half precision
                assert np.allclose(last[..., :3], img[..., :3], rtol=2e-3, atol=1e-4), "e2e: rgba16f frame and FP32 result disagree"
        h2d = lights_bytes + 3 * 64 + 32                    # light arrays (pinned) + camera matrices + settings (kernel parameters)
        d2h = npx * 8 + npx * 16 / e2e_steps                # per frame rgba16f + the FP32 result once per job
        e2e = {"value": e_rays / e_dt / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "steps": e2e_steps, "job_steps": job_steps, "ms_per_step": e_dt / e2e_steps * 1e3, "mpix_spp_per_s": e_samples / e_dt / 1e6, "samples_per_wave": prefetch, "wave_schedule": schedule,
               "result_format": "per frame: rgba16_sfloat (the reference's OutputData.color) read back on rank 0; once per job: the FP32 image"
                                + (" after the NCCL reduce" if world > 1 else ""),
               "d2h_note": "rank 0's bytes; the other ranks copy nothing to the host" if world > 1 else "every frame's image"}
        assert e2e["value"] > 0.0
        r.close()
        return e2e


    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(my_steps)
        # the same measurement over 128 frames per rank when the job is shorter: a K-frame job ends with the read-back of its last wave
        # (samples_per_wave frames x the PCIe time of one image), which a long-running pass amortises
        if my_steps < 128 and args.scaling == "weak":
            e2e["steady_state_128_steps"] = {k: v for k, v in run_e2e(128).items() if k in ("value", "ms_per_step", "steps", "samples_per_wave", "wave_schedule", "mpix_spp_per_s")}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- CPU baseline + algorithmic bytes (oracle counters on the bit-identical BVH) --
    cpu = None
    hbm_peak, sm_max_mhz, peak_src = peaks()
    roofline = None
    ora = None
    if not args.no_cpu_baseline:
        stride = args.cpu_tile_stride or (8 if args.config == "atrium" else 16)
        ora = oracle_sample(scene, mode, args, stride, 1)
        if ora["seconds"] < 5.0:        # aim for ~10-30 s of CPU work
            frames = int(min(1024, max(2, 15.0 / max(ora["seconds"], 1e-3))))
            ora = oracle_sample(scene, mode, args, stride, frames)
        cpu = {"value": ora["rays_per_s"] / 1e6, "unit": "Mrays/s", "cores": ora["cores"], "kind": "port",
               "sample": ora["desc"], "seconds": ora["seconds"]}
    total_k = kt.raygen_ms + kt.extend_ms + kt.shade_ms + kt.connect_ms + kt.other_ms
    if kt.extend_launches:
        rays_per_launch = pc.extend_rays / kt.extend_launches
        avg_s = kt.extend_ms * 1e-3 / kt.extend_launches
        clk_mhz = clocks.get("sm_mhz") or sm_max_mhz
        ceilings = {}
        algorithmic = None
        if ora is not None:
            # extend ray = 32 B ray in + 16 B hit out + 64 B x nodes + 48 B x tris (oracle counters on the same rays);
            # rays that walk the 4-wide tree: 64 B per wide node + 32 B per exact leaf box instead of the binary nodes
            per_ray = 32 + 16 + 64 * ora["ext_nodes"] + 64 * ora["ext_wide_nodes"] + 32 * ora["ext_leaf_boxes"] + 48 * ora["ext_tris"]
            algorithmic = {"bytes_per_ray": per_ray, "gb_per_launch": per_ray * rays_per_launch / 1e9, "gb_per_s": per_ray * rays_per_launch / avg_s / 1e9,
                           "nodes_per_ray": ora["ext_nodes"], "tris_per_ray": ora["ext_tris"], "wide_nodes_per_ray": ora["ext_wide_nodes"],
                           "leaf_boxes_per_ray": ora["ext_leaf_boxes"], "wide_ray_share": ora["ext_wide_ray_share"], "instances_per_ray": ora["ext_inst"],
                           "note": "bytes the algorithm asks for (SURVEY §8d), served mostly from L1/L2: compare with `traffic`"}
        hw = extend_counters(args)
        traffic = None
        if hw is not None:
            # issue ceiling: every SM sub-partition issues at most one warp instruction per clock
            inst_per_ray = hw["warp_inst_per_ray"]
            issue_peak = SM_COUNT * SCHEDULERS_PER_SM * clk_mhz * 1e6 / 1e9             # G warp-instructions / s
            issue_ach = inst_per_ray * rays_per_launch / avg_s / 1e9
            ceilings["issue"] = {"achieved": issue_ach, "peak": issue_peak, "unit": "Gwarp-inst/s", "frac": issue_ach / issue_peak,
                                 "warp_inst_per_ray": inst_per_ray, "sm_clock_mhz": clk_mhz,
                                 "active_threads_per_warp": hw["active_threads_per_warp"], "simt_efficiency": hw["active_threads_per_warp"] / 32.0,
                                 "useful_lane_frac": issue_ach / issue_peak * hw["active_threads_per_warp"] / 32.0}
            traffic = hw["dram_bytes_per_ray"] * rays_per_launch / 1e9                  # GB per launch (dram read + write, ncu --set full)
            hbm_ach = traffic / avg_s
            ceilings["hbm"] = {"achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak, "dram_bytes_per_ray": hw["dram_bytes_per_ray"]}
        if ceilings:
            bound = max(ceilings, key=lambda k: ceilings[k]["frac"])
            top = ceilings[bound]
            roofline = {"bound": bound, "kernel": hw["kernel"], "achieved": top["achieved"], "peak": top["peak"], "unit": top["unit"], "frac": top["frac"],
                        "traffic": traffic, "traffic_unit": "GB per launch (dram__bytes_read.sum + dram__bytes_write.sum)",
                        "ceilings": ceilings, "algorithmic": algorithmic, "counters_source": hw["source"], "peak_source": peak_src,
                        "rays_per_launch": rays_per_launch, "avg_launch_ms": avg_s * 1e3,
                        "how": "launch duration: CUDA events around every extend launch on the library's stream, live in this run; warp instructions and "
                               "DRAM bytes PER RAY: the committed ncu capture of the same kernel on the same workload (profiles/), scaled by this run's rays per launch; "
                               "bound = the ceiling with the larger fraction"}
        elif algorithmic is not None:
            roofline = {"bound": "hbm", "kernel": "k_trace_spec (extend)", "achieved": algorithmic["gb_per_s"], "peak": hbm_peak, "unit": "GB/s",
                        "frac": algorithmic["gb_per_s"] / hbm_peak, "traffic": None, "algorithmic": algorithmic, "peak_source": peak_src,
                        "rays_per_launch": rays_per_launch, "avg_launch_ms": avg_s * 1e3,
                        "how": "ALGORITHMIC bytes only (no ncu counters committed for this configuration): an upper bound on DRAM traffic, not a measurement of it"}
    kernel_ms = {"steps": prof_steps, "raygen": kt.raygen_ms / prof_steps, "extend": kt.extend_ms / prof_steps, "shade": kt.shade_ms / prof_steps,
                 "connect": kt.connect_ms / prof_steps, "other": kt.other_ms / prof_steps}
    if total_k > 0:
        kernel_ms["share"] = {"raygen": kt.raygen_ms / total_k, "extend": kt.extend_ms / total_k, "shade": kt.shade_ms / total_k,
                              "connect": kt.connect_ms / total_k, "other": kt.other_ms / total_k}
    state_mb = npx * (6 * 16 + 20 + 16) // 1000000
    out = {
        "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / max(1, args.steps), "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "width": W, "height": H, "max_bounces": B, "accel": args.accel, "triangles": scene.num_triangles,
                   "step": ("1 spp over the full frame per GPU (rank r renders frames [r*steps, (r+1)*steps))" if args.scaling == "weak" else
                            "1 spp over the full frame; the steps-frame job is split over the GPUs in contiguous blocks") + "; samples per wave = min(frames of the rank, 2^26 / pixels)",
                   "l2_policy": "inputs larger than L2: ~%d MB of per-path wavefront state per sample streams through HBM every step" % state_mb,
                   "multi_gpu": "sample-index sharding, scene+BVH replicated, one NCCL reduce of the FP32 sum buffer per batch" if world > 1 else "single GPU"},
        "mpix_spp_per_s": mpix, "extend_rays": ext, "shadow_rays": shd, "accel_build_ms": build_ms,
        "clocks": clocks, "gpu_launches": int(launches), "kernel_ms_per_step": kernel_ms, "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu,
    }
    if reduce_check is not None:
        out["reduce_check"] = reduce_check
    if timed_split is not None:
        out["timed_region_split"] = timed_split
    if e2e is not None and args.scaling == "weak":
        # an end-to-end step contains the device step: it cannot be faster (2 % allowance for run-to-run noise); reported, not asserted,
        # so that a noisy box still yields a line
        e2e["not_faster_than_device_step"] = bool(e2e["ms_per_step"] >= 0.98 * ms / max(1, args.steps))
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
