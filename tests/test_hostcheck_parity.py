"""CPU parity of the CUDA SOURCE against the oracle: tests/hostcheck compiles the kernels' per-thread
device functions (csrc/*.cuh) for the host and runs the same wavefront logic path by path. Bit-exact
agreement here means the two independently written implementations follow one numeric contract;
the GPU tests then only have to confirm that the device executes that source the same way."""
import ctypes as C
import os

import numpy as np
import pytest

import _hostcheck as HC
from bisemutum_engine_b200 import capi, scenes

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODES = [capi.ACCEL_TWO_LEVEL, capi.ACCEL_MERGED]


def test_math_functions_bit_exact(oracle):
    H, L = HC.lib(), oracle.library().lib
    rng = np.random.default_rng(0)
    for a, b in rng.integers(0, 2 ** 32, (200, 2)):
        assert H.hc_rng_tea(int(a), int(b)) == L.obpt_rng_tea(int(a), int(b))
    s1, c1, s2, c2 = C.c_float(), C.c_float(), C.c_float(), C.c_float()
    for u in rng.uniform(0, 1, 2000).astype(np.float32):
        H.hc_sincos_2pi(float(u), C.byref(s1), C.byref(c1)); L.obpt_sincos_2pi(float(u), C.byref(s2), C.byref(c2))
        assert s1.value == s2.value and c1.value == c2.value
    for y, x in rng.normal(size=(2000, 2)).astype(np.float32):
        assert H.hc_atan2(float(y), float(x)) == L.obpt_atan2(float(y), float(x))
        assert H.hc_acos(float(np.clip(x, -1, 1))) == L.obpt_acos(float(np.clip(x, -1, 1)))
    o1, o2 = np.zeros(3, np.float32), np.zeros(3, np.float32)
    for _ in range(500):
        v = rng.normal(size=3).astype(np.float32); v /= np.linalg.norm(v)
        rx, ry, u1, u2 = (float(np.float32(x)) for x in rng.uniform(0.001, 1, 4))
        H.hc_ggx_vndf_sample(v.ctypes.data_as(C.c_void_p), rx, ry, u1, u2, o1.ctypes.data_as(C.c_void_p))
        L.obpt_ggx_vndf_sample(v.ctypes.data_as(C.c_void_p), rx, ry, u1, u2, o2.ctypes.data_as(C.c_void_p))
        np.testing.assert_array_equal(o1, o2)
        N = np.array([0, 0, 1], np.float32); T = np.array([1, 0, 0], np.float32)
        Ld = rng.normal(size=3).astype(np.float32); Ld /= np.linalg.norm(Ld)
        cols = [rng.uniform(0, 1, 3).astype(np.float32) for _ in range(3)]
        args = [a.ctypes.data_as(C.c_void_p) for a in (N, T, v, Ld, *cols)]
        H.hc_surface_eval_lit(*args, rx, ry, o1.ctypes.data_as(C.c_void_p))
        L.obpt_surface_eval_lit(*args, rx, ry, o2.ctypes.data_as(C.c_void_p))
        np.testing.assert_array_equal(o1, o2)


def _scene(name):
    if name == "cornell":
        return scenes.cornell_box(tess=8)
    if name == "small":
        return scenes.small_test_scene()
    if name == "mixed":
        luts = scenes.load_ltc_luts(os.path.join(GOLDEN, "ltc_luts.npz"))
        return scenes.add_mixed_lights(scenes.small_test_scene(), 5, 3, luts, keep_dir_lights=True, light_range=12.0)
    raise KeyError(name)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("name,bounces", [("cornell", 5), ("small", 6), ("mixed", 3)])
def test_render_bit_exact(oracle, name, bounces, mode):
    scene = _scene(name)
    W, H = 40, 28
    ctx = oracle.OracleContext(W, H)
    ctx.upload_scene(scene, mode)
    cam = oracle.camera_matrices(scene.camera, W, H)
    st = capi.Settings(max_bounces=bounces)
    ctx.render(cam, 5, 2, st)
    ref = ctx.resolve(1)
    got = HC.HostScene(scene, ctx, mode).render(cam, W, H, 5, 2, st)
    assert np.isfinite(ref).all() and ref[..., :3].max() > 0
    np.testing.assert_array_equal(got[..., :3], ref[..., :3])


@pytest.mark.parametrize("mode", MODES)
def test_nee_none_bit_exact(oracle, mode):
    scene = _scene("small")
    W, H = 32, 24
    ctx = oracle.OracleContext(W, H); ctx.upload_scene(scene, mode)
    cam = oracle.camera_matrices(scene.camera, W, H)
    st = capi.Settings(max_bounces=4, nee_mode=capi.NEE_NONE)
    ctx.render(cam, 0, 1, st)
    np.testing.assert_array_equal(HC.HostScene(scene, ctx, mode).render(cam, W, H, 0, 1, st)[..., :3], ctx.resolve(1)[..., :3])


@pytest.mark.parametrize("mode", MODES)
def test_trace_bit_exact(oracle, mode):
    scene = _scene("small")
    ctx = oracle.OracleContext(8, 8); ctx.upload_scene(scene, mode)
    rng = np.random.default_rng(9)
    n = 5000
    rays = np.zeros(n, capi.RAY)
    rays["origin"] = rng.uniform(-4.5, 4.5, (n, 3)).astype(np.float32) * np.float32([1, 0.4, 1]) + np.float32([0, 1.5, 0])
    d = rng.normal(size=(n, 3)); rays["direction"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays["tmin"] = 0.001; rays["tmax"] = rng.choice([100.0, 2.0], n).astype(np.float32)
    rays["direction"][:50, 0] = 0.0          # axis-parallel components exercise the 2^-80 clamp of 1/d
    hits, vis = HC.HostScene(scene, ctx, mode).trace(rays, 3)
    ref = ctx.trace_rays(rays, 3)
    for f in ("t", "u", "v", "instance", "primitive"):
        np.testing.assert_array_equal(hits[f], ref[f])
    np.testing.assert_array_equal(vis, ctx.trace_shadow_rays(rays, 3))
    assert 0.2 < (ref["t"] >= 0).mean() < 0.98


@pytest.mark.parametrize("mode", MODES)
def test_probe_tracing_bit_exact(oracle, mode):
    """DDGI-style probe rays (ddgi/trace_gbuffer.hlsl + ddgi/deferred_lighting.hlsl) through the same code."""
    scene = _scene("mixed")
    table = scenes.ddgi_sample_randoms()
    assert table.shape == (8192, 2) and table.min() >= 0 and table.max() < 1 and tuple(table[0]) == (0.5, 0.5)
    vol = scenes.probe_volume(scene, (3, 2, 3), 16, ray_length=50.0)
    ctx = oracle.OracleContext(8, 8); ctx.upload_scene(scene, mode)
    for bounces in (1, 3):
        ref = ctx.trace_probes(vol, table, 4, bounces)
        got = HC.HostScene(scene, ctx, mode).trace_probes(vol, table, 4, bounces)
        np.testing.assert_array_equal(got, ref)
    assert (ref[:, 3] > 0).any() and (ref[:, 3] < 0).any() and np.isfinite(ref).all()


@pytest.mark.parametrize("switch", ["russian_roulette", "pixel_jitter", "rect_shadow"])
def test_mode_switches_bit_exact(oracle, switch):
    """The NEW switches of SURVEY §0 (default off) are implemented identically on both sides."""
    scene = _scene("mixed" if switch == "rect_shadow" else "small")
    W, H = 36, 24
    ctx = oracle.OracleContext(W, H); ctx.upload_scene(scene, capi.ACCEL_MERGED)
    cam = oracle.camera_matrices(scene.camera, W, H)
    st = capi.Settings(max_bounces=8, **{switch: 1})
    base = oracle.OracleContext(W, H); base.upload_scene(scene, capi.ACCEL_MERGED)
    base.render(cam, 2, 2, capi.Settings(max_bounces=8))
    ctx.render(cam, 2, 2, st)
    ref = ctx.resolve(1)
    np.testing.assert_array_equal(HC.HostScene(scene, ctx, capi.ACCEL_MERGED).render(cam, W, H, 2, 2, st)[..., :3], ref[..., :3])
    assert (ref != base.resolve(1)).any()                                       # the switch does something
    if switch == "russian_roulette":
        assert ctx.counters().extend_rays < base.counters().extend_rays          # fewer rays on dark paths
    with pytest.raises(capi.BptError):
        ctx.render(cam, 0, 1, capi.Settings(state_precision=1))                 # rejected, not silently ignored


def test_probe_blending_bit_exact(oracle):
    """DDGI probe blending: irradiance / visibility octahedral gathers, gamma-5 temporal blend,
    border + corner copies — CUDA source (host build) vs oracle, plus structural properties of the atlases."""
    scene = _scene("small")
    table = scenes.ddgi_sample_randoms()
    vol = scenes.probe_volume(scene, (3, 2, 2), 64, ray_length=100.0)
    ctx = oracle.OracleContext(8, 8); ctx.upload_scene(scene, capi.ACCEL_MERGED)
    rays0, rays1 = ctx.trace_probes(vol, table, 0, 2), ctx.trace_probes(vol, table, 1, 2)
    irr0, vis0 = ctx.blend_probes(vol, table, 0, rays0)
    assert irr0.shape == (2 * 8, 3 * 2 * 8, 4) and vis0.shape == (2 * 16, 3 * 2 * 16, 2)
    h_irr, h_vis = HC.blend_probes(vol, table, 0, rays0, np.zeros_like(irr0), np.zeros_like(vis0))
    np.testing.assert_array_equal(h_irr, irr0); np.testing.assert_array_equal(h_vis, vis0)
    irr1, vis1 = ctx.blend_probes(vol, table, 1, rays1, irr0, vis0)                      # temporal blend, alpha 0.97
    h_irr, h_vis = HC.blend_probes(vol, table, 1, rays1, irr0.copy(), vis0.copy(), history_valid=1)
    np.testing.assert_array_equal(h_irr, irr1); np.testing.assert_array_equal(h_vis, vis1)
    assert np.isfinite(irr1).all() and np.isfinite(vis1).all()
    assert 0 < np.abs(irr1 - irr0).max() < 0.2 * irr0.max()                               # history dominates (alpha = 0.97)
    p = irr0[0:8, 0:8, :3]                                                                # first probe: 6x6 interior + border
    np.testing.assert_array_equal(p[0, 0], p[6, 6]); np.testing.assert_array_equal(p[7, 7], p[1, 1])   # corners: diagonally opposite interior
    np.testing.assert_array_equal(p[0, 2], p[1, 5])                                       # top border row mirrors the first interior row
    # a uniform radiance field gives that radiance back in every texel (weights normalise)
    flat = rays0.copy(); flat[:, :3] = (0.25, 0.5, 1.0)
    irr_flat, _ = ctx.blend_probes(vol, table, 0, flat)
    np.testing.assert_allclose(irr_flat[..., :3], np.broadcast_to(np.float32([0.25, 0.5, 1.0]), irr_flat[..., :3].shape), rtol=2e-6)


@pytest.mark.parametrize("mode", MODES)
def test_reference_example_scene_bit_exact(oracle, mode):
    """examples/scene_basic of the reference (its meshes, its five materials incl. sRGB / normal-map textures,
    alpha-tested cage, translucent cube, world-space checkerboard): CUDA source (host build) == oracle."""
    scene = scenes.scene_basic(os.path.join(GOLDEN, "scene_basic.npz"))
    assert scene.num_triangles == 2 + 3 * 12 + 960 and len(scene.textures) == 3
    W, H = 80, 40
    ctx = oracle.OracleContext(W, H); ctx.upload_scene(scene, mode)
    cam = oracle.camera_matrices(scene.camera, W, H)
    st = capi.Settings(max_bounces=4)
    ctx.render(cam, 0, 2, st)
    ref = ctx.resolve(1)
    np.testing.assert_array_equal(HC.HostScene(scene, ctx, mode).render(cam, W, H, 0, 2, st)[..., :3], ref[..., :3])
    assert ref[..., :3].mean() > 0.05
