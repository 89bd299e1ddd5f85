"""Generates tests/golden/golden_v1.npz with the CPU oracle (the reference ships no golden vectors for
this path — SURVEY §4/§8c "parity unpinned" — so these pin the ORACLE against drift and give the CUDA
path a fixed, committed target in addition to the live oracle comparison).

    python tests/golden/make_golden.py          (build container; needs oracle/_build/liboracle.so)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from bisemutum_engine_b200 import capi, scenes  # noqa: E402
from oracle import oracle_py  # noqa: E402

OUT = os.path.join(HERE, "golden_v1.npz")
W, H = 48, 32


def golden_rays(scene, n=512, seed=21):
    rng = np.random.default_rng(seed)
    lo, hi = scene.bounds
    rays = np.zeros(n, capi.RAY)
    rays["origin"] = rng.uniform(lo - 0.5, hi + 0.5, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    rays["direction"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays["tmin"], rays["tmax"] = 0.001, 100.0
    return rays


def cases():
    luts = scenes.load_ltc_luts(os.path.join(HERE, "ltc_luts.npz"))
    yield "small", scenes.small_test_scene(), 6
    yield "cornell", scenes.cornell_box(tess=8), 5
    yield "mixed", scenes.add_mixed_lights(scenes.small_test_scene(), 5, 3, luts, keep_dir_lights=True, light_range=12.0), 3
    # the reference's own example project: its meshes, its five materials (textures, alpha test, translucency)
    yield "scene_basic", scenes.scene_basic(os.path.join(HERE, "scene_basic.npz")), 3


def compute(name, scene, bounces, make_ctx):
    out = {}
    for mode, mname in ((capi.ACCEL_TWO_LEVEL, "two_level"), (capi.ACCEL_MERGED, "merged")):
        ctx = make_ctx(W, H)
        ctx.upload_scene(scene, mode)
        k = f"{name}.{mname}"
        b = ctx.read_bvh(0)
        out[f"{k}.bvh0.morton"] = b["morton"]; out[f"{k}.bvh0.prims"] = b["prims"]
        out[f"{k}.bvh0.child0"] = b["nodes"]["child0"]; out[f"{k}.bvh0.child1"] = b["nodes"]["child1"]
        if mode == capi.ACCEL_TWO_LEVEL:
            t = ctx.read_bvh(capi.BVH_TLAS)
            out[f"{k}.tlas.prims"] = t["prims"]; out[f"{k}.tlas.child0"] = t["nodes"]["child0"]
        hits = ctx.trace_rays(golden_rays(scene), 2)
        for f in ("t", "u", "v", "instance", "primitive"):
            out[f"{k}.hits.{f}"] = hits[f]
        cam = oracle_py.camera_matrices(scene.camera, W, H)
        st = capi.Settings(max_bounces=bounces)
        for frame in (0, 9):
            ctx.clear_accum()
            ctx.render(cam, frame, 1, st)
            out[f"{k}.image.f{frame}"] = ctx.resolve(1)[..., :3].copy()
        c = ctx.counters()
        out[f"{k}.extend_per_bounce"] = np.array(list(c.extend_rays_per_bounce), np.uint64)
        out[f"{k}.shadow_per_bounce"] = np.array(list(c.shadow_rays_per_bounce), np.uint64)
        ctx.close()
    return out


# ---- golden_v2.npz: the passes either side of the path (post-process, sky IBL, ray-traced reflections) ----
OUT2 = os.path.join(HERE, "golden_v2.npz")
IBL_DESC = dict(diffuse_size=4, specular_size=8, specular_levels=4, brdf_lut_size=8, diffuse_strength=0.9, specular_strength=0.7)


def compute_v2(make_ctx, post_process):
    """make_ctx(w, h) -> context of the implementation under test; post_process(ctx, sums, spp, settings) -> image."""
    out = {}
    scene = scenes.scene_basic(os.path.join(HERE, "scene_basic.npz"))
    ctx = make_ctx(W, H)
    ctx.upload_scene(scene, capi.ACCEL_TWO_LEVEL)
    cam = oracle_py.camera_matrices(scene.camera, W, H)
    ctx.render(cam, 0, 2, capi.Settings(max_bounces=3))
    sums = ctx.resolve(1)
    out["post.bloom"] = post_process(ctx, sums, 2, capi.PostSettings(True, 0.5, 0.5))[..., :3].copy()
    d, s, b = ctx.precompute_sky_ibl(capi.SkyIblDesc(**IBL_DESC))
    out["ibl.diffuse"] = d[..., :3].copy(); out["ibl.brdf"] = b
    for l, lv in enumerate(s):
        out[f"ibl.specular{l}"] = lv[..., :3].copy()
    depth, g = ctx.render_primary(cam, 0, capi.Settings(max_bounces=3))
    for half, ibl in ((True, False), (False, True)):
        refl, hit = ctx.trace_reflection(cam, 3, depth, g, capi.ReflectionSettings(16.0, 1.0, 1.0, 0.5, half, ibl))
        out[f"rtr.half{int(half)}.ibl{int(ibl)}.color"] = refl[..., :3].copy(); out[f"rtr.half{int(half)}.ibl{int(ibl)}.hit"] = hit
    ctx.close()
    return out


def main():
    oracle_py.build()
    out = {}
    for name, scene, bounces in cases():
        out.update(compute(name, scene, bounces, oracle_py.OracleContext))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(out), "arrays")
    out2 = compute_v2(oracle_py.OracleContext, lambda ctx, sums, spp, st: oracle_py.post_process_image(sums * (np.float32(1.0) / np.float32(spp)), st))
    np.savez_compressed(OUT2, **out2)
    print("wrote", OUT2, os.path.getsize(OUT2), "bytes,", len(out2), "arrays")


if __name__ == "__main__":
    main()
