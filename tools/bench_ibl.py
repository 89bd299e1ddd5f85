"""Sky IBL precompute (csrc/ibl.cu = SkyboxPrecomputePass) at the reference's sizes (skybox.cpp:11-29: diffuse 256^2 cube, specular 256^2 cube
with 5 mips, 128^2 BRDF LUT) from a 256^2 procedural sky: wall clock of bpt_precompute_sky_ibl (synchronised). Not the contract bench."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bisemutum_engine_b200 as pkg
from bisemutum_engine_b200 import capi, scenes

lib = pkg.load_library()
scene = scenes.small_test_scene()
scene.sky_faces = scenes.procedural_sky(256, (0.4, 1.0, 0.6))
ctx = capi.Context(lib, 64, 64); ctx.upload_scene(scene, capi.ACCEL_MERGED)
d = capi.SkyIblDesc()
ts = []
for _ in range(3):
    t0 = time.perf_counter(); ctx._call("precompute_sky_ibl", capi.C.byref(d)); ctx.sync(); ts.append((time.perf_counter() - t0) * 1e3)
samples = 6 * 256 * 256 * 4096 + sum(6 * (256 >> l) ** 2 for l in range(5)) * 1024 + 128 * 128 * 1024
print(json.dumps(dict(case="sky_ibl_256_256x5_128", ms_min=min(ts), ms_all=ts, sky_samples=samples, gsamples_per_s=samples / min(ts) / 1e6)))
