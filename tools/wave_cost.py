"""Cost of one sample wave as a function of its size (configs[1], device-resident): what the e2e arm's wave schedule trades against the
read-back tail. One JSON line: {n: ms per wave}."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bisemutum_engine_b200 as pkg
from bisemutum_engine_b200 import capi, engine, scenes

lib = pkg.load_library()
W, H, B = 1920, 1080, 8
scene = scenes.atrium()
ctx = capi.Context(lib, W, H)
stream = torch.cuda.current_stream()
ctx.set_stream(stream.cuda_stream)
ctx.upload_scene(scene, capi.ACCEL_MERGED)
cam = engine.camera_matrices(scene.camera, W, H)
st = capi.Settings(ray_length=100.0, max_bounces=B)
ctx.render(cam, 10_000, 16, st); ctx.sync()
out = {}
for n in (1, 2, 3, 4, 5, 6, 8, 10, 12, 14, 15, 16, 17, 18, 20, 24, 32):
    best = 1e30
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); ctx.render(cam, 0, n, st); e1.record(stream); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out[n] = best
print(json.dumps(out))
