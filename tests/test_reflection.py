"""Ray-traced reflections: ReflectionPass::render_raytraced
(bisemutum/src/renderer/pass/reflection.cpp:317-450; specular_sample.hlsl, rt_gbuffer.hlsl, deferred_lighting_secondary.hlsl).

CPU: the CUDA source's per-thread functions (csrc/bpt_aov.cuh: rtr_pixel_ray + the shared trace / shade functions, host build)
== oracle bit for bit, plus the properties the shaders imply. GPU: bpt_trace_reflection through the C ABI == oracle.
"""
import os

import numpy as np
import pytest

import _hostcheck as HC
import bisemutum_engine_b200 as pkg
from bisemutum_engine_b200 import capi, scenes

GOLDEN = os.path.join(pkg.REPO_ROOT, "tests", "golden")
W, H = 64, 40


def _glossy_scene():
    """The reference's example scene with its materials made glossy enough for the default max_roughness = 0.3."""
    scene = scenes.scene_basic(os.path.join(GOLDEN, "scene_basic.npz"))
    return scene


def _inputs(oracle, scene, mode):
    ctx = oracle.OracleContext(W, H); ctx.upload_scene(scene, mode)
    cam = oracle.camera_matrices(scene.camera, W, H)
    depth, g = ctx.render_primary(cam, 0, capi.Settings(max_bounces=4))
    return ctx, cam, depth, g


SETTINGS = [capi.ReflectionSettings(16.0, 1.0, 0.3, 0.1, True), capi.ReflectionSettings(16.0, 1.0, 0.3, 0.1, False),
            capi.ReflectionSettings(4.0, 2.5, 1.0, 0.6, True), capi.ReflectionSettings(100.0, 1.0, 0.6, 0.9, False)]


@pytest.mark.parametrize("mode", [capi.ACCEL_TWO_LEVEL, capi.ACCEL_MERGED])
def test_cuda_source_bit_exact_on_host(oracle, mode):
    scene = _glossy_scene()
    ctx, cam, depth, g = _inputs(oracle, scene, mode)
    hs = HC.HostScene(scene, ctx, mode)
    seen_rays = 0
    for k, rs in enumerate(SETTINGS):
        for frame in ((0, 1, 2, 3) if rs.half_resolution and k == 0 else (5,)):
            refl, hit = ctx.trace_reflection(cam, frame, depth, g, rs)
            hrefl, hhit = hs.trace_reflection(cam, W, H, frame, depth, g, rs)
            np.testing.assert_array_equal(refl.view(np.uint32), hrefl.view(np.uint32))
            np.testing.assert_array_equal(hit.view(np.uint32), hhit.view(np.uint32))
            rh, rw = refl.shape[:2]
            assert (rh, rw) == ((H // 2, W // 2) if rs.half_resolution else (H, W))
            assert np.isfinite(refl).all() and (refl[..., 3] == 1).all() and (refl[..., :3] >= 0).all()
            has_ray = (hit[..., :3] != 0).any(axis=2)
            seen_rays += int(has_ray.sum())
            # no ray: background pixels and pixels rougher than max_roughness -> colour 0, hit position (0, 0, 0, -1)
            assert (refl[~has_ray][:, :3] == 0).all() and (hit[~has_ray][:, 3] == -1).all()
            # a hit lies within the range; a miss stores the unit ray direction
            hits = has_ray & (hit[..., 3] >= 0)
            assert (hit[hits][:, 3] <= rs.range + 1e-3).all()
            miss = has_ray & (hit[..., 3] < 0)
            if miss.any():
                np.testing.assert_allclose(np.linalg.norm(hit[miss][:, :3], axis=1), 1.0, atol=1e-5)
            if not rs.half_resolution:
                tex_rough = g["normal_roughness"][..., 3]
                np.testing.assert_array_equal(has_ray, (depth > 0) & (tex_rough <= np.float32(rs.max_roughness)))
    assert seen_rays > 200
    # strength scales the result linearly (the weight is multiplied before lighting)
    a, _ = ctx.trace_reflection(cam, 5, depth, g, capi.ReflectionSettings(16.0, 1.0, 1.0, 0.6, False))
    b, _ = ctx.trace_reflection(cam, 5, depth, g, capi.ReflectionSettings(16.0, 2.0, 1.0, 0.6, False))
    np.testing.assert_allclose(b[..., :3], 2 * a[..., :3], rtol=1e-5, atol=1e-7)
    assert a[..., :3].max() > 0


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [capi.ACCEL_TWO_LEVEL, capi.ACCEL_MERGED])
def test_gpu_reflection_matches_oracle(oracle, mode):
    scene = _glossy_scene()
    ctx, cam, depth, g = _inputs(oracle, scene, mode)
    gpu = capi.Context(pkg.load_library(), W, H)
    gpu.upload_scene(scene, mode)
    gdepth, gg = gpu.render_primary(cam, 0, capi.Settings(max_bounces=4))
    np.testing.assert_array_equal(gdepth, depth)
    for rs, frame in zip(SETTINGS, (0, 3, 6, 9)):
        ctx.reset_counters(); gpu.reset_counters()
        refl, hit = ctx.trace_reflection(cam, frame, depth, g, rs)
        grefl, ghit = gpu.trace_reflection(cam, frame, gdepth, gg, rs)
        np.testing.assert_array_equal(ghit.view(np.uint32), hit.view(np.uint32))               # ray set, hits and hit positions: bit-exact
        scale = max(float(refl[..., :3].max()), 1e-6)
        assert np.abs(grefl - refl).max() <= 1e-4 * scale                                       # colour: 1e-4 (unordered shadow-ray atomics)
        c, gc = ctx.counters(), gpu.counters()
        assert (c.extend_rays, c.shadow_rays) == (gc.extend_rays, gc.shadow_rays) and c.extend_rays > 100
    gpu.close(); ctx.close()
