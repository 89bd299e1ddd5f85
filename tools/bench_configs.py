"""Throughput of every BASELINE.json config on one GPU (device-resident, CUDA events), written as JSON lines.
Not the contract bench (that is bench.py on configs[1]); this records the other configs for profiles/."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bisemutum_engine_b200 as pkg
from bisemutum_engine_b200 import capi, engine, scenes

lib = pkg.load_library()
ONLY = [a for a in sys.argv[1:] if not a.startswith("-")]       # optional filter: config-name prefixes ("1_atrium_1080p_64spp", "3_", ...)
GOLDEN = os.path.join(pkg.REPO_ROOT, "tests", "golden")


def want(name):
    return not ONLY or any(name.startswith(o) for o in ONLY)


def run(name, scene, W, H, bounces, spp, mode, ray_length=100.0):
    if not want(name):
        return None
    if callable(scene):
        scene = scene()
    ctx = capi.Context(lib, W, H)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    t0 = time.perf_counter(); ctx.upload_scene(scene, mode); ctx.sync(); build_ms = (time.perf_counter() - t0) * 1e3
    cam = engine.camera_matrices(scene.camera, W, H)
    st = capi.Settings(ray_length=ray_length, max_bounces=bounces)
    ctx.render(cam, 10_000, min(spp, 8), st); ctx.sync(); ctx.reset_counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); ctx.render(cam, 0, spp, st); e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1); c = ctx.counters()
    out = dict(config=name, width=W, height=H, spp=spp, max_bounces=bounces, accel="merged" if mode else "two_level", triangles=scene.num_triangles,
               ms_total=ms, ms_per_spp=ms / spp, mrays_per_s=(c.extend_rays + c.shadow_rays) / ms / 1e3, mpix_spp_per_s=W * H * spp / ms / 1e3,
               extend_rays=c.extend_rays, shadow_rays=c.shadow_rays, accel_build_ms=build_ms)
    if "--classes" in sys.argv:          # per-class kernel milliseconds per sample (CUDA events inside the library; serialises nothing)
        ctx.profile_enable(True); ctx.render(cam, 0, min(spp, 8), st); kt = ctx.profile_read(); ctx.profile_enable(False)
        for f, _ in kt._fields_:
            v = getattr(kt, f)
            if isinstance(v, float):
                out[f] = v / min(spp, 8)
    print(json.dumps(out), flush=True)
    ctx.close()
    return out


luts = scenes.load_ltc_luts(os.path.join(GOLDEN, "ltc_luts.npz"))
res = []
res.append(run("0_cornell_512x512_16spp_depth5", scenes.cornell_box, 512, 512, 5, 16, capi.ACCEL_MERGED))
atr = scenes.atrium()
res.append(run("1_atrium_1080p_256spp_depth8", atr, 1920, 1080, 8, 256, capi.ACCEL_MERGED))
res.append(run("1_atrium_1080p_64spp_depth8_two_level", atr, 1920, 1080, 8, 64, capi.ACCEL_TWO_LEVEL))
res.append(run("2_mixed_lights_1080p_128spp_depth3", lambda: scenes.mixed_lights(luts), 1920, 1080, 3, 128, capi.ACCEL_MERGED))
res.append(run("3_instanced_2Mx512_4k_8spp_depth8", scenes.instanced, 3840, 2160, 8, 8, capi.ACCEL_TWO_LEVEL, ray_length=1000.0))
# config 4: probes
if want("4_ddgi"):
    ctx = capi.Context(lib, 1920, 1080); ctx.upload_scene(atr, capi.ACCEL_MERGED)
    vol = scenes.probe_volume(atr, (32, 32, 16), 256); tab = scenes.ddgi_sample_randoms()
    n_probes = 32 * 32 * 16
    dev = torch.empty((n_probes * 256, 4), dtype=torch.float32, device="cuda")          # per-ray results stay on the device (bpt.h: host or device pointer)
    ctx.trace_probes_range_into(vol, tab, 100, 2, 0, n_probes, dev.data_ptr()); ctx.reset_counters()
    torch.cuda.synchronize()
    t0 = time.perf_counter(); ctx.trace_probes_range_into(vol, tab, 0, 2, 0, n_probes, dev.data_ptr()); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    c = ctx.counters()
    t1 = time.perf_counter(); ctx.trace_probes(vol, tab, 0, 2); dth = time.perf_counter() - t1
    res.append(dict(config="4_ddgi_probes_32x32x16x256_2bounces", ms_total=dt * 1e3, mrays_per_s=(c.extend_rays + c.shadow_rays) / dt / 1e6,
                    extend_rays=c.extend_rays, shadow_rays=c.shadow_rays, ms_total_with_host_readback=dth * 1e3,
                    note="wall clock (synchronised), per-ray results device-resident; the host variant adds a 67 MB read-back"))
    print(json.dumps(res[-1]))
res = [r for r in res if r]
os.makedirs(os.path.join(pkg.REPO_ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(pkg.REPO_ROOT, "gpurun_out", "configs.json"), "w"), indent=1)
