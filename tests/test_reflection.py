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


# ---- sky IBL: SkyboxPrecomputePass (skybox_precompute.cpp:66-162) + the IBL block of the lighting shader --------------------
IBL = capi.SkyIblDesc(diffuse_size=8, specular_size=16, specular_levels=5, brdf_lut_size=16, diffuse_strength=0.8, specular_strength=0.6)


def _ggx_brdf_lut_reference(n, res=16, samples=4096):
    """Independent float64 Monte-Carlo of the split-sum integrals the LUT stores (isotropic GGX alpha = the LUT's `roughness`
    coordinate, height-correlated-free separable G1 weights as the shader: E[(1 - F) G1(wo)], E[F G1(wo)] under VNDF sampling),
    evaluated by plain numerical quadrature of D * G1(wi) * G1(wo) / (4 cos_i) over the hemisphere."""
    out = np.zeros((res, res, 2))
    th = (np.arange(n) + 0.5) / n * (np.pi / 2); ph = (np.arange(4 * n) + 0.5) / (4 * n) * 2 * np.pi
    T, P = np.meshgrid(th, ph, indexing="ij")
    wo = np.stack([np.sin(T) * np.cos(P), np.sin(T) * np.sin(P), np.cos(T)], -1)
    dw = (np.pi / 2 / n) * (2 * np.pi / (4 * n)) * np.sin(T)
    for y in range(res):
        a = (y + 0.5) / res
        for x in range(res):
            c = (x + 0.5) / res
            wi = np.array([np.sqrt(1 - c * c), 0.0, c])
            h = wo + wi; h /= np.linalg.norm(h, axis=-1, keepdims=True)
            D = 1.0 / (np.pi * a * a * ((h[..., 0] / a) ** 2 + (h[..., 1] / a) ** 2 + h[..., 2] ** 2) ** 2)
            g1 = lambda v: 2.0 / (1.0 + np.sqrt(1.0 + (a * a * (v[..., 0] ** 2 + v[..., 1] ** 2)) / np.maximum(v[..., 2] ** 2, 1e-4)))
            hv = np.maximum((h * wi).sum(-1), 0.0)
            f = (1 - hv) ** 5
            w = D * g1(wi[None, None]) * g1(wo) / (4.0 * c) * dw      # pdf_vndf(wo) * G1(wo) dw
            out[y, x] = ((1 - f) * w).sum(), (f * w).sum()
    return out


def test_sky_ibl_precompute_bit_exact_and_plausible(oracle):
    scene = scenes.small_test_scene()                                      # 16^2 procedural sky (gradient + sun lobe)
    ctx = oracle.OracleContext(8, 8); ctx.upload_scene(scene, capi.ACCEL_MERGED)
    diffuse, spec, brdf = ctx.precompute_sky_ibl(IBL)
    hd, hs, hb = HC.precompute_sky_ibl(scene, IBL)
    np.testing.assert_array_equal(diffuse.view(np.uint32), hd.view(np.uint32))
    np.testing.assert_array_equal(brdf.view(np.uint32), hb.view(np.uint32))
    for a, b in zip(spec, hs):
        np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32))
    HC.precompute_sky_ibl(scene, None)
    # formats: rgba16_sfloat cubes, rg8_unorm LUT
    np.testing.assert_array_equal(diffuse, diffuse.astype(np.float16).astype(np.float32))
    np.testing.assert_allclose(brdf * 255, np.round(brdf * 255), atol=1e-4)
    assert [s.shape for s in spec] == [(6, 16 >> l, 16 >> l, 4) for l in range(5)]
    # level 0 (roughness 0) is the sky itself at the texel's direction (luminance-clamped), rougher levels are smoother
    sky = np.asarray(scene.sky_faces, np.float32)
    lum = sky[..., :3] @ np.array([0.212671, 0.715160, 0.072169], np.float32)
    clamped = sky[..., :3] * (12.0 / np.maximum(lum, 12.0))[..., None]
    np.testing.assert_allclose(spec[0][..., :3], clamped, rtol=2e-3, atol=1e-3)
    assert spec[4][..., :3].std() < spec[1][..., :3].std() < spec[0][..., :3].std()
    # diffuse irradiance: up-facing texels (more sky) are brighter than down-facing ones; value scale = pi * mean radiance-ish
    assert diffuse[2, ..., :3].mean() > diffuse[3, ..., :3].mean() > 0
    # BRDF LUT against an independent float64 quadrature of the same integrals (16 x 16 LUT; 8-bit storage + 1024 samples)
    want = _ggx_brdf_lut_reference(96)
    sel = np.s_[3:, 2:]                                                    # (very low roughness / grazing cells need finer quadrature than this test affords)
    np.testing.assert_allclose(brdf[sel], want[sel], atol=0.03)
    assert brdf[-1, -1, 0] < brdf[1, -1, 0] and brdf[8, 0, 1] > brdf[8, -1, 1]      # rougher -> less energy; grazing -> more f90 weight


@pytest.mark.parametrize("mode", [capi.ACCEL_MERGED])
def test_reflection_with_ibl_bit_exact_on_host(oracle, mode):
    scene = _glossy_scene()
    ctx, cam, depth, g = _inputs(oracle, scene, mode)
    ctx.precompute_sky_ibl(IBL)
    HC.precompute_sky_ibl(scene, IBL)
    hs = HC.HostScene(scene, ctx, mode)
    for half in (True, False):
        rs = capi.ReflectionSettings(16.0, 1.0, 1.0, 0.6, half, ibl=True)
        refl, hit = ctx.trace_reflection(cam, 2, depth, g, rs)
        hrefl, hhit = hs.trace_reflection(cam, W, H, 2, depth, g, rs)
        np.testing.assert_array_equal(refl.view(np.uint32), hrefl.view(np.uint32))
        np.testing.assert_array_equal(hit.view(np.uint32), hhit.view(np.uint32))
        base, bhit = ctx.trace_reflection(cam, 2, depth, g, capi.ReflectionSettings(16.0, 1.0, 1.0, 0.6, half, ibl=False))
        np.testing.assert_array_equal(bhit, hit)                           # IBL only adds light at hits
        hits = hit[..., 3] >= 0
        assert (refl[..., :3] >= base[..., :3]).all() and (refl[hits][:, :3] > base[hits][:, :3]).any()
        np.testing.assert_array_equal(refl[~hits], base[~hits])
    HC.precompute_sky_ibl(scene, None)
    ctx.upload_sky(scene)                                                  # a new sky invalidates the derived textures
    with pytest.raises(capi.BptError):
        ctx.trace_reflection(cam, 2, depth, g, capi.ReflectionSettings(ibl=True))


@pytest.mark.gpu
def test_gpu_sky_ibl_and_reflection_with_ibl(oracle):
    scene = _glossy_scene()
    ctx, cam, depth, g = _inputs(oracle, scene, capi.ACCEL_MERGED)
    gpu = capi.Context(pkg.load_library(), W, H); gpu.upload_scene(scene, capi.ACCEL_MERGED)
    with pytest.raises(capi.BptError):
        gpu.trace_reflection(cam, 2, depth, g, capi.ReflectionSettings(ibl=True))     # not precomputed yet
    d0, s0, b0 = ctx.precompute_sky_ibl(IBL)
    d1, s1, b1 = gpu.precompute_sky_ibl(IBL)
    np.testing.assert_array_equal(d1.view(np.uint32), d0.view(np.uint32))
    np.testing.assert_array_equal(b1.view(np.uint32), b0.view(np.uint32))
    for a, b in zip(s1, s0):
        np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32))
    rs = capi.ReflectionSettings(16.0, 1.0, 1.0, 0.6, True, ibl=True)
    refl, hit = ctx.trace_reflection(cam, 2, depth, g, rs)
    grefl, ghit = gpu.trace_reflection(cam, 2, depth, g, rs)
    np.testing.assert_array_equal(ghit.view(np.uint32), hit.view(np.uint32))
    assert np.abs(grefl - refl).max() <= 1e-4 * max(float(refl[..., :3].max()), 1e-6)
    # the reference's sizes (skybox.cpp:11-29) run too; spot-check the LUT against the oracle's small one is not possible (different
    # resolution), so check the GPU at full size against the properties only
    dd, ss, bb = gpu.precompute_sky_ibl(capi.SkyIblDesc())
    assert dd.shape == (6, 256, 256, 4) and len(ss) == 5 and bb.shape == (128, 128, 2) and np.isfinite(dd).all() and np.isfinite(ss[4]).all()
    gpu.close(); ctx.close()


# ---- "RTR Upscale Hit / Color": simple_upscale.hlsl (reflection.cpp:452-530) ------------------------------------------------
def _upscale_case(oracle, size=(W, H)):
    w, h = size
    scene = _glossy_scene()
    ctx = oracle.OracleContext(w, h); ctx.upload_scene(scene, capi.ACCEL_MERGED)
    cam = oracle.camera_matrices(scene.camera, w, h)
    depth, g = ctx.render_primary(cam, 0, capi.Settings(max_bounces=4))
    return ctx, cam, depth, g


def test_upscale_half_res_bit_exact_on_host(oracle):
    ctx, cam, depth, g = _upscale_case(oracle)
    nr = g["normal_roughness"]
    for frame in range(4):
        refl, hit = ctx.trace_reflection(cam, frame, depth, g, capi.ReflectionSettings(16.0, 1.0, 1.0, 0.6, True))
        for half in (refl, hit):
            full = ctx.upscale_half_res(cam, frame, depth, nr, half)
            hfull = HC.upscale_half_res(cam, W, H, frame, depth, nr, half)
            np.testing.assert_array_equal(full.view(np.uint32), hfull.view(np.uint32))
            assert full.shape == (H, W, 4) and np.isfinite(full).all()
    # a constant input stays that constant wherever any tap has weight (the filter is a normalised average)
    const = np.full((H // 2, W // 2, 4), 0.625, np.float32)
    full = ctx.upscale_half_res(cam, 1, depth, nr, const)
    inner = full[:-2, :-2]                                    # (the last row / column also average the out-of-range half-res texel, which Loads as 0)
    nz = inner[..., 0] != 0
    assert nz.mean() > 0.5
    np.testing.assert_allclose(inner[nz], 0.625, rtol=1e-6)
    # the pixel whose sub-pixel equals the traced one takes (almost) only its own half-res texel: gaussian(0) = 1, the others <= exp(-12.5)
    refl, _ = ctx.trace_reflection(cam, 3, depth, g, capi.ReflectionSettings(16.0, 1.0, 1.0, 0.6, True))
    full = ctx.upscale_half_res(cam, 3, depth, nr, refl)
    own = full[1::2, 1::2]                                   # frame 3 traced sub-pixel (0.75, 0.75) = odd x, odd y
    sel = (depth[1::2, 1::2] > 0) & (refl[..., :3].max(axis=2) > 0)
    np.testing.assert_allclose(own[sel][:, :3], refl[sel][:, :3], rtol=2e-3, atol=1e-5)


def test_fixed_order_exp(oracle):
    """exp_neg of the numeric contract (oracle_math.hpp = csrc/bpt_math.cuh) against libm, through the upscale weights: a one-texel
    impulse spreads to its neighbours with gaussian(sqrt(|offset|), 0.2) = exp(-25 |offset|)."""
    ctx, cam, depth, g = _upscale_case(oracle)
    nr = np.zeros((H, W, 4), np.float32)                      # oct (0, 0) -> normal (0, 0, 1) everywhere: normal weight 1
    flat = np.full((H, W), 0.5, np.float32)                  # constant depth: depth weight 1
    imp = np.zeros((H // 2, W // 2, 4), np.float32); imp[8, 10] = 1.0
    full = ctx.upscale_half_res(cam, 0, flat, nr, imp)       # frame 0: traced sub-pixel (0.25, 0.25) = even x, even y
    # full-res pixel (20, 16) has the impulse as its centre tap with offset 0; pixel (21, 16) sees it at offset (-0.5, 0)
    offs = lambda x, y: [((dx + 0.25) - (0.75 if x & 1 else 0.25), (dy + 0.25) - (0.75 if y & 1 else 0.25)) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
    for (x, y) in ((20, 16), (21, 16), (21, 17)):
        ws = np.array([np.exp(-25.0 * np.hypot(ox, oy)) for ox, oy in offs(x, y)])
        want = ws[4] / ws.sum()                               # the impulse is the centre tap (x / 2, y / 2) = (10, 8)
        np.testing.assert_allclose(full[y, x, 0], want, rtol=2e-5)


@pytest.mark.gpu
def test_gpu_upscale_half_res_bit_exact(oracle):
    ctx, cam, depth, g = _upscale_case(oracle)
    gpu = capi.Context(pkg.load_library(), W, H)
    nr = g["normal_roughness"]
    for frame in (0, 3):
        refl, hit = ctx.trace_reflection(cam, frame, depth, g, capi.ReflectionSettings(16.0, 1.0, 1.0, 0.6, True))
        for half in (refl, hit):
            np.testing.assert_array_equal(gpu.upscale_half_res(cam, frame, depth, nr, half).view(np.uint32),
                                          ctx.upscale_half_res(cam, frame, depth, nr, half).view(np.uint32))
    gpu.close(); ctx.close()


@pytest.mark.gpu
def test_gpu_bulk_pointers_may_be_device_pointers(oracle):
    """include/bpt/bpt.h: images, G-buffers and results of the pass-level entry points may live on the device (no PCIe round trip):
    reflection, upscale and RTAO with torch device tensors as inputs and outputs == the host-pointer calls, bit for bit."""
    import ctypes as C
    import torch
    ctx, cam, depth, g = _inputs(oracle, _glossy_scene(), capi.ACCEL_MERGED)
    ctx.close()
    gpu = capi.Context(pkg.load_library(), W, H); gpu.upload_scene(_glossy_scene(), capi.ACCEL_MERGED)
    rs = capi.ReflectionSettings(16.0, 1.0, 1.0, 0.6, True)
    refl, hit = gpu.trace_reflection(cam, 2, depth, g, rs)
    full = gpu.upscale_half_res(cam, 2, depth, g["normal_roughness"], refl)
    ao = gpu.trace_ao(cam, 2, depth, g["normal_roughness"], 0.5, 0.5, True)
    d_depth = torch.from_numpy(np.ascontiguousarray(depth)).cuda()
    d_g = torch.from_numpy(np.ascontiguousarray(g).view(np.float32).reshape(H, W, 16).copy()).cuda()
    d_nr = torch.from_numpy(np.ascontiguousarray(g["normal_roughness"])).cuda()
    d_refl = torch.empty((H // 2, W // 2, 4), device="cuda"); d_hit = torch.empty_like(d_refl)
    vp = lambda t: C.c_void_p(t.data_ptr())
    gpu._call("trace_reflection", C.byref(cam), 2, C.byref(rs), vp(d_depth), vp(d_g), vp(d_refl), vp(d_hit))
    np.testing.assert_array_equal(d_hit.cpu().numpy().view(np.uint32), hit.view(np.uint32))
    assert np.abs(d_refl.cpu().numpy() - refl).max() <= 1e-4 * max(float(refl[..., :3].max()), 1e-6)
    d_full = torch.empty((H, W, 4), device="cuda")
    gpu._call("upscale_half_res", C.byref(cam), 2, vp(d_depth), vp(d_nr), vp(torch.from_numpy(refl).cuda()), vp(d_full))
    np.testing.assert_array_equal(d_full.cpu().numpy().view(np.uint32), full.view(np.uint32))
    d_ao = torch.empty((H // 2, W // 2, 2), device="cuda")
    aos = capi.AoSettings(0.5, 0.5, 1)
    gpu._call("trace_ao", C.byref(cam), 2, C.byref(aos), vp(d_depth), vp(d_nr), vp(d_ao))
    np.testing.assert_array_equal(d_ao.cpu().numpy().view(np.uint32), ao.view(np.uint32))
    gpu.close()
