"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck): both accel modes, lights of every
kind, any-hit materials, probes + blending, prefetch path, TLAS update. Kept small: the tools slow kernels ~50x."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch   # first: its bundled NCCL must be the one in the process before libbpt dlopen()s "libnccl.so.2" (bpt_comm_init below)

import bisemutum_engine_b200 as pkg
from bisemutum_engine_b200 import capi, engine, scenes

lib = pkg.load_library()
luts = scenes.load_ltc_luts(os.path.join(pkg.REPO_ROOT, "tests", "golden", "ltc_luts.npz"))
scene = scenes.add_mixed_lights(scenes.small_test_scene(), 3, 2, luts, keep_dir_lights=True, light_range=12.0)
basic = scenes.scene_basic(os.path.join(pkg.REPO_ROOT, "tests", "golden", "scene_basic.npz"))
W, H = 40, 24
for sc in (scene, basic):
    for mode in (capi.ACCEL_TWO_LEVEL, capi.ACCEL_MERGED):
        ctx = capi.Context(lib, W, H)
        ctx.upload_scene(sc, mode)
        cam = engine.camera_matrices(sc.camera, W, H)
        ctx.render(cam, 0, 3, capi.Settings(max_bounces=5, rect_shadow=1, russian_roulette=1, pixel_jitter=1))
        img = ctx.resolve(3)
        assert np.isfinite(img).all()
        rays = np.zeros(64, capi.RAY); rays["origin"] = (0, 2, 5); rays["direction"] = (0, -0.3, -1); rays["tmin"] = 0.001; rays["tmax"] = 50
        ctx.trace_rays(rays); ctx.trace_shadow_rays(rays)
        if mode == capi.ACCEL_TWO_LEVEL:
            ctx.upload_instances(sc.instances); ctx.update_tlas()
        vol = scenes.probe_volume(sc, (2, 2, 2), 16, ray_length=50.0); tab = scenes.ddgi_sample_randoms()
        r = ctx.trace_probes(vol, tab, 0, 2)
        irr, vis = ctx.blend_probes(vol, tab, 0, r)
        ctx.blend_probes(vol, tab, 1, r, irr, vis)
        # added after the first sanitizer run: DDGI consumer + feedback, fp16 state mode, primary outputs, RTAO (the merged
        # mode runs all of these through the 4-wide tree kernels from bounce 2 on)
        ctx.set_ddgi_volume(vol, irr, vis)
        ctx.trace_probes(vol, tab, 1, 2)
        pts = np.float32(np.random.default_rng(0).uniform(-3, 3, (64, 3))); up = np.float32(np.tile([0, 1, 0], (64, 1)))
        ctx.ddgi_lighting(pts, up, up)
        ctx.set_ddgi_volume(None)
        ctx.clear_accum()
        ctx.render(cam, 0, 2, capi.Settings(max_bounces=4, state_precision=capi.STATE_REFERENCE_FP16))
        assert np.isfinite(ctx.resolve(2)).all()
        ctx.clear_accum()
        depth, g = ctx.render_primary(cam, 0, capi.Settings(max_bounces=4))
        ctx.trace_ao(cam, 1, depth, g["normal_roughness"], 0.5, 0.5, True)
        ctx.trace_ao(cam, 2, depth, g["normal_roughness"], 0.5, 0.5, False)
        # added in v9: ray-traced reflections (two-level mode = the wide TLAS + wide BLAS kernels), post-process (bloom level kernels
        # with two shared-memory tiles), rebuilds out of the scratch arena (single-block LBVH for the TLAS / small BLASes, the
        # multi-kernel radix-sort build for the rest)
        ctx.trace_reflection(cam, 1, depth, g, capi.ReflectionSettings(16.0, 1.0, 1.0, 0.5, True))
        ctx.trace_reflection(cam, 2, depth, g, capi.ReflectionSettings(16.0, 1.0, 0.3, 0.1, False))
        refl, hitp = ctx.trace_reflection(cam, 1, depth, g, capi.ReflectionSettings(16.0, 1.0, 1.0, 0.5, True))
        ctx.upscale_half_res(cam, 1, depth, g["normal_roughness"], refl)
        ctx.precompute_sky_ibl(capi.SkyIblDesc(4, 8, 4, 8))
        ctx.trace_reflection(cam, 3, depth, g, capi.ReflectionSettings(16.0, 1.0, 1.0, 0.5, False, ibl=True))
        ctx.trace_probes_range(vol, tab, 2, 2, 3, 4)
        ctx.render(cam, 0, 2, capi.Settings(max_bounces=3))
        assert np.isfinite(ctx.post_process(capi.PostSettings(True, 0.3, 0.5), 2)).all()
        ctx.post_process(capi.PostSettings(False), 2)
        ctx.build_accel(mode); ctx.build_accel(mode)
        ctx.clear_accum(); ctx.render(cam, 0, 1, capi.Settings(max_bounces=3))
        if mode == capi.ACCEL_MERGED:
            ctx.read_wide()
            ctx.comm_init(lib.comm_unique_id(), 0, 1); ctx.render(cam, 0, 1, capi.Settings(max_bounces=3)); ctx.reduce(0); ctx.sync()
        ctx.close()
r = engine.Renderer(W, H); r.set_scene(scene, capi.ACCEL_MERGED)
for _ in range(3):
    n = r.frame(max_bounces=4)
target = torch.zeros(H, W, 4, dtype=torch.float16, device="cuda")      # OutputData.color written by the pass (bpt_accumulate_ahead_rgba16f)
r.set_color_target(target.data_ptr())
for _ in range(2):
    n = r.frame(max_bounces=4)
r.ctx.sync(); assert torch.isfinite(target).all()
r.image(n); r.post_process(True, 0.3, 0.5); r.close()
pc = capi.Context(lib, W, H); pc.upload_scene(scenes.small_test_scene(), capi.ACCEL_MERGED)
engine.run_renderer(pc, scenes.small_test_scene(), W, H, 2, max_bounces=3, bloom=True, bloom_threshold=0.3); pc.close()
# a tree above the single-block limit (4096 leaves) through the multi-kernel build, both modes, and an odd-sized post-process
big = scenes.cornell_box(tess=40)
for mode in (capi.ACCEL_TWO_LEVEL, capi.ACCEL_MERGED):
    ctx = capi.Context(lib, 37, 23)
    ctx.upload_scene(big, mode)
    cam = engine.camera_matrices(big.camera, 37, 23)
    ctx.render(cam, 0, 2, capi.Settings(max_bounces=3))
    assert np.isfinite(ctx.post_process(capi.PostSettings(True, 0.2, 1.0), 2)).all()
    ctx.close()
# round 2: textured rect lights (level-0 decode + mip kernels, three formats), the specialised kernels' counterparts (a scene with and one without
# any-hit instances / rect lights ran above), vertex colours, the rgba16f resolve, a small wave budget, ReBLUR at full and half resolution
tl = scenes.add_mixed_lights(scenes.small_test_scene(), 1, 3, luts, keep_dir_lights=True, light_range=12.0)
scenes.texture_rect_lights(tl, [scenes.light_texture(37, 22, capi.TEXTURE_RGBA8_SRGB), scenes.light_texture(16, 16, capi.TEXTURE_RGBA8_UNORM, mip_linear=0),
                                scenes.light_texture(9, 5, capi.TEXTURE_RGBA32_FLOAT, levels=3, linear=0)])
tl.colors = np.random.default_rng(1).uniform(0, 1, tl.positions.size).astype(np.float32)
tl.drawables["color_offset"] = tl.drawables["position_offset"]; tl.drawable_va = tl.drawable_va | capi.VA_COLOR
tl.materials["flags"][0] = (tl.materials["flags"][0] & ~np.uint32(0xff00)) | np.uint32(capi.MATERIAL_KIND_VERTEX_COLOR << 8)
for mode in (capi.ACCEL_TWO_LEVEL, capi.ACCEL_MERGED):
    ctx = capi.Context(lib, W, H); ctx.upload_scene(tl, mode)
    ctx.set_wave_budget(W * H * 2)
    cam = engine.camera_matrices(tl.camera, W, H)
    ctx.render(cam, 0, 5, capi.Settings(max_bounces=4))
    half = torch.empty(H, W, 4, dtype=torch.float16, device="cuda")
    ctx.resolve_device_rgba16f(5, half.data_ptr()); ctx.sync()
    assert torch.isfinite(half).all()
    for k in range(3):
        ctx.read_light_texture(k)
    ctx.upload_light_textures([])
    ctx.close()
sb = scenes.scene_basic(os.path.join(pkg.REPO_ROOT, "tests", "golden", "scene_basic.npz"))
ctx = capi.Context(lib, W, H); ctx.upload_scene(sb, capi.ACCEL_MERGED)
cam = engine.camera_matrices(sb.camera, W, H)
for half_res in (False, True):
    ctx.reblur_reset()
    for k in range(3):
        depth, g = ctx.render_primary(cam, k, capi.Settings(max_bounces=4))
        refl, hitp = ctx.trace_reflection(cam, k, depth, g, capi.ReflectionSettings(16.0, 1.0, 1.0, 0.6, half_res))
        vel = np.full((H, W, 2), 0.003, np.float32) if k == 2 else None
        mask = (np.random.default_rng(k).uniform(size=refl.shape[:2]) < 0.1).astype(np.uint8) if k == 2 else None
        out = ctx.denoise_reblur(cam, k, refl, hitp, depth, np.ascontiguousarray(g["normal_roughness"]), vel, mask)
        assert np.isfinite(out).all()
ctx.close()
print("sanitize smoke ok")
