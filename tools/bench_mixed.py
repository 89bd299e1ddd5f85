import os, sys, json
sys.path.insert(0, '/root/repo')
import torch
import bisemutum_engine_b200 as pkg
from bisemutum_engine_b200 import capi, engine, scenes
lib = pkg.load_library()
luts = scenes.load_ltc_luts(os.path.join(pkg.REPO_ROOT, "tests", "golden", "ltc_luts.npz"))
scene = scenes.mixed_lights(luts)
W, H = 1920, 1080
ctx = capi.Context(lib, W, H); stream = torch.cuda.current_stream(); ctx.set_stream(stream.cuda_stream)
ctx.upload_scene(scene, capi.ACCEL_MERGED)
cam = engine.camera_matrices(scene.camera, W, H); st = capi.Settings(max_bounces=3)
ctx.render(cam, 1000, 2, st); ctx.sync(); ctx.reset_counters()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream); ctx.render(cam, 0, 16, st); e1.record(stream); torch.cuda.synchronize()
c = ctx.counters(); ms = e0.elapsed_time(e1)
out = {"tag": os.environ.get("BPT_WIDE_CONNECT_FROM_BOUNCE", "default"), "ms_per_spp": ms / 16, "mrays_per_s": (c.extend_rays + c.shadow_rays) / ms / 1e3,
       "extend_rays_per_spp": c.extend_rays / 16, "shadow_rays_per_spp": c.shadow_rays / 16}
ctx.profile_enable(True); ctx.render(cam, 0, 16, st); kt = ctx.profile_read(); ctx.profile_enable(False)
for f, _ in kt._fields_:
    v = getattr(kt, f)
    out[f] = (v / 16) if isinstance(v, float) else v
print(json.dumps(out))
