"""Kernel-tuning loop: configs[1] (atrium, 1080p, depth 8) device-resident, one JSON line with the total and the per-class
kernel milliseconds per 1-spp frame (CUDA events inside libbpt). Not the contract bench (bench.py)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bisemutum_engine_b200 as pkg
from bisemutum_engine_b200 import capi, engine, scenes

ap = argparse.ArgumentParser()
ap.add_argument("--spp", type=int, default=64)
ap.add_argument("--accel", default="merged")
ap.add_argument("--tag", default="")
ap.add_argument("--lib", default="", help="experimental libbpt variant (tools/build_variant.sh)")
ap.add_argument("--config", default="atrium", choices=["atrium", "instanced", "mixed"])
args = ap.parse_args()

lib = capi.Library(os.path.abspath(args.lib), "bpt_", capi.BPT_ONLY_API) if args.lib else pkg.load_library()
W, H, B = 1920, 1080, 8
RAY_LENGTH = 100.0
if args.config == "instanced":
    W, H, RAY_LENGTH = 3840, 2160, 1000.0
    scene = scenes.instanced(); args.accel = "two_level"
elif args.config == "mixed":
    B = 3
    scene = scenes.mixed_lights(scenes.load_ltc_luts(os.path.join(pkg.REPO_ROOT, "tests", "golden", "ltc_luts.npz")))
else:
    scene = scenes.atrium()
mode = capi.ACCEL_MERGED if args.accel == "merged" else capi.ACCEL_TWO_LEVEL
ctx = capi.Context(lib, W, H)
stream = torch.cuda.current_stream()
ctx.set_stream(stream.cuda_stream)
ctx.upload_scene(scene, mode)
cam = engine.camera_matrices(scene.camera, W, H)
st = capi.Settings(ray_length=RAY_LENGTH, max_bounces=B)
ctx.render(cam, 10_000, 16 if args.config == "atrium" else 8, st); ctx.sync(); ctx.reset_counters()
best = 1e30
for rep in range(3):
    ctx.reset_counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); ctx.render(cam, 0, args.spp, st); e1.record(stream); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
c = ctx.counters()
PS = 16 if args.config == "atrium" else 8
ctx.profile_enable(True)
ctx.render(cam, 0, PS, st)
kt = ctx.profile_read()
ctx.profile_enable(False)
out = {"tag": args.tag, "config": args.config, "ms_per_spp": best / args.spp, "mrays_per_s": (c.extend_rays + c.shadow_rays) / best / 1e3}
for f, _ in kt._fields_:
    v = getattr(kt, f)
    out[f] = (v / PS) if isinstance(v, float) else v
print(json.dumps(out))
