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
    if name == "lighttex":                       # rect lights with textures: sRGB / unorm / float chains, linear and nearest mip filters
        luts = scenes.load_ltc_luts(os.path.join(GOLDEN, "ltc_luts.npz"))
        sc = scenes.add_mixed_lights(scenes.small_test_scene(), 2, 4, luts, keep_dir_lights=True, light_range=12.0)
        sc.rect_lights["two_sided"][::2] = 1
        texs = [scenes.light_texture(37, 22, capi.TEXTURE_RGBA8_SRGB, mip_linear=1), scenes.light_texture(16, 16, capi.TEXTURE_RGBA8_UNORM, mip_linear=0, seed=8),
                scenes.light_texture(9, 5, capi.TEXTURE_RGBA32_FLOAT, levels=3, mip_linear=1, linear=0, seed=9)]
        return scenes.texture_rect_lights(sc, texs)
    if name == "vcolor":                         # the vertex-colour stream (VA_TYPE_COLOR) through a `base_color = vertex.color` material
        sc = scenes.small_test_scene()
        rng = np.random.default_rng(3)
        sc.colors = rng.uniform(0.05, 1.0, sc.positions.size).astype(np.float32)
        sc.drawables["color_offset"] = sc.drawables["position_offset"]
        sc.drawable_va = sc.drawable_va | capi.VA_COLOR
        kinds = (sc.materials["flags"] >> 8) & 0xff
        pick = np.nonzero(kinds != capi.MATERIAL_KIND_TRANSPARENT)[0][:2]
        sc.materials["flags"][pick] = (sc.materials["flags"][pick] & ~np.uint32(0xff00)) | np.uint32(capi.MATERIAL_KIND_VERTEX_COLOR << 8)
        return sc
    raise KeyError(name)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("name,bounces", [("cornell", 5), ("small", 6), ("mixed", 3), ("lighttex", 3), ("vcolor", 4)])
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
        ctx.render(cam, 0, 1, capi.Settings(state_precision=1))                 # a change of accumulation rule needs clear_accum
    with pytest.raises(capi.BptError):
        ctx.render(cam, 0, 1, capi.Settings(state_precision=7))                 # unknown value: rejected, not silently ignored


def test_fp16_storage_formats_bit_exact(oracle):
    """state_precision = reference_fp16 building blocks: the half store is IEEE round-to-nearest-even (pinned against
    numpy.float16, an independent implementation), and the G-buffer pack -> texture formats -> unpack round trip
    (gbuffer.hlsl:18-45, pack.hlsl) is bit-identical between the CUDA source (host build) and the oracle."""
    H, L = HC.lib(), oracle.library().lib
    rng = np.random.default_rng(5)
    vals = np.concatenate([
        rng.normal(size=4000).astype(np.float32) * np.float32(10.0) ** rng.integers(-9, 6, 4000).astype(np.float32),
        np.float32([0.0, -0.0, 65504.0, 65519.99, 65520.0, 1e9, -1e9, 2.0 ** -24, 2.0 ** -25, 2.0 ** -25 * 1.0001, 2.0 ** -14, 6.0e-8, 3.0e-8,
                    1.0 + 2.0 ** -11, 1.0 + 3 * 2.0 ** -11, 0.1, 1 / 3, np.inf, -np.inf]),
        (np.arange(2048, dtype=np.float32) + 0.5) * np.float32(2.0 ** -24),          # every denormal-half tie
    ])
    with np.errstate(over="ignore"):
        want = vals.astype(np.float16).astype(np.float32)
    got_h = np.float32([H.hc_q_half(float(v)) for v in vals]); got_o = np.float32([L.obpt_store_half(float(v)) for v in vals])
    np.testing.assert_array_equal(got_h.view(np.uint32), want.view(np.uint32))
    np.testing.assert_array_equal(got_o.view(np.uint32), want.view(np.uint32))
    assert np.isnan(H.hc_q_half(float("nan"))) and np.isnan(L.obpt_store_half(float("nan")))
    o1, o2, m1, m2 = np.zeros(18, np.float32), np.zeros(18, np.float32), C.c_uint32(), C.c_uint32()
    worst_n = worst_f0 = 0.0
    for k in range(1500):
        N = rng.normal(size=3).astype(np.float32); N /= np.linalg.norm(N)
        if k < 6: N = np.float32(np.eye(3)[k % 3] * (1 if k < 3 else -1))               # axis-aligned normals (oct-map corners / poles)
        T = np.cross(N, rng.normal(size=3)).astype(np.float32); T /= np.linalg.norm(T)
        surf = np.concatenate([rng.uniform(0, 1, 3), rng.uniform(0.02, 1, 6), rng.uniform(0.02, 1, 1), rng.uniform(0, 1, 1), rng.uniform(1.0, 2.5, 1)]).astype(np.float32)
        args = (N.ctypes.data_as(C.c_void_p), T.ctypes.data_as(C.c_void_p), surf.ctypes.data_as(C.c_void_p), 1)
        H.hc_surface_through_gbuffer(*args, o1.ctypes.data_as(C.c_void_p), C.byref(m1))
        L.obpt_gbuffer_roundtrip(*args, o2.ctypes.data_as(C.c_void_p), C.byref(m2))
        np.testing.assert_array_equal(o1.view(np.uint32), o2.view(np.uint32))
        assert m1.value == m2.value == 1                                                # MATERIAL_SURFACE_MODEL_LIT survives /256 -> unorm8 -> *255.5
        worst_n = max(worst_n, float(np.abs(o1[0:3] - N).max())); worst_f0 = max(worst_f0, float(np.abs(o1[9:12] - surf[3:6]).max()))
        assert abs(np.linalg.norm(o1[0:3]) - 1) < 1e-6 and abs(np.linalg.norm(o1[3:6]) - 1) < 1e-6 and abs(np.dot(o1[0:3], o1[3:6])) < 1e-6
        assert np.abs(o1[6:9] - surf[0:3]).max() <= 2.0 ** -11 and abs(o1[15] - surf[9]) <= 2.0 ** -11     # half: 11 significant bits
        assert abs(o1[16] - surf[10]) <= 0.5 / 255 + 1e-7 and abs(1 / o1[17] - 1 / surf[11]) <= 0.5 / 255 + 1e-7
    assert 0 < worst_n < 2e-3 and 0 < worst_f0 < 34 / 2047         # quantised; green can lose 32/2047: the unorm16 quirk below takes 1 off the HIGH half-word when blue >= 0.5
    # Reference quirk reproduced literally: pack_u32_to_unorm16x2 stores k/65536 in a unorm16 texel, which rounds to k-1
    # for k > 32768, and unpack_u32_from_unorm16x2 only tolerates errors upwards — so a zero 11-bit red field BORROWS:
    # f0 = (0, .5, .5) comes back with red = 2047/2047.
    N, T = np.float32([0, 0, 1]), np.float32([1, 0, 0])
    surf = np.float32([0.5, 0.5, 0.5, 0.0, 0.5, 0.5, 1, 1, 1, 0.5, 0, 1.5])
    L.obpt_gbuffer_roundtrip(N.ctypes.data_as(C.c_void_p), T.ctypes.data_as(C.c_void_p), surf.ctypes.data_as(C.c_void_p), 1, o2.ctypes.data_as(C.c_void_p), C.byref(m2))
    assert o2[9] == 1.0 and abs(o2[10] - 0.5) < 34 / 2047 and abs(o2[11] - 0.5) < 2e-3


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("name,bounces", [("cornell", 5), ("small", 6)])
def test_render_reference_fp16_bit_exact(oracle, name, bounces, mode):
    """state_precision = reference_fp16 (half state, packed G-buffer, half additive blit, running half lerp): CUDA source
    (host build) == oracle bit for bit; every stored value is a half; and the mode differs from fp32 by ~1e-3, not more."""
    scene = _scene(name)
    W, H = 40, 28
    ctx = oracle.OracleContext(W, H); ctx.upload_scene(scene, mode)
    cam = oracle.camera_matrices(scene.camera, W, H)
    st = capi.Settings(max_bounces=bounces, state_precision=capi.STATE_REFERENCE_FP16)
    ctx.render(cam, 3, 2, st); ctx.render(cam, 5, 1, st)                                  # 3 samples, in two calls (count carries over)
    ref = ctx.resolve(3)
    got = HC.HostScene(scene, ctx, mode).render(cam, W, H, 3, 3, st)
    np.testing.assert_array_equal(got[..., :3].view(np.uint32), ref[..., :3].view(np.uint32))
    np.testing.assert_array_equal(ref[..., :3], ref[..., :3].astype(np.float16).astype(np.float32))      # an rgba16_sfloat image
    full = oracle.OracleContext(W, H); full.upload_scene(scene, mode)
    full.render(cam, 3, 3, capi.Settings(max_bounces=bounces))
    f32 = full.resolve(3)[..., :3]
    assert (ref[..., :3] != f32).any()
    # the fp16 state decorrelates individual paths (a 1e-3 change of a direction is a different sample after a bounce),
    # but the image mean must agree closely
    assert abs(float(ref[..., :3].mean()) - float(f32.mean())) < 0.03 * float(f32.mean())
    one = oracle.OracleContext(W, H); one.upload_scene(scene, mode)
    one.render(cam, 3, 1, st)
    first = one.resolve(1)[..., :3]
    assert first.max() > 0 and np.isfinite(first).all()


@pytest.mark.parametrize("mode", MODES)
def test_primary_outputs_and_rtao_bit_exact(oracle, mode):
    """OutputData{depth, gbuffer} of the pass (pt_depth.hlsl, rt_gbuffer_hit.hlsl + gbuffer.hlsl packing into the texture
    formats) and the RTAO pass that consumes depth + normal G-buffer (ambient_occlusion_rt.hlsl): CUDA source (host build)
    == oracle, plus the properties the formats and the geometry imply."""
    scene = scenes.scene_basic(os.path.join(GOLDEN, "scene_basic.npz"))       # has an alpha-tested and a translucent drawable
    W, H = 64, 40
    ctx = oracle.OracleContext(W, H); ctx.upload_scene(scene, mode)
    cam = oracle.camera_matrices(scene.camera, W, H)
    st = capi.Settings(max_bounces=4)
    depth, g = ctx.render_primary(cam, 0, st)
    hs = HC.HostScene(scene, ctx, mode)
    hdepth, hg = hs.render_primary(cam, W, H, 0, st)
    np.testing.assert_array_equal(depth.view(np.uint32), hdepth.view(np.uint32))
    for f in capi.GBUFFER_TEXEL.names:
        np.testing.assert_array_equal(g[f].view(np.uint32), hg[f].view(np.uint32), err_msg=f)
    hit = depth > 0
    assert 0.3 < hit.mean() < 1.0 and (depth[hit] < 1.0).all()                                   # reverse-Z: (0, 1), 0 = background
    assert (g["base_color"][~hit] == 0).all() and (g["base_color"][hit][:, 3] == 1).all()
    for f in ("base_color", "normal_roughness"):                                                 # rgba16_sfloat texels
        np.testing.assert_array_equal(g[f], g[f].astype(np.float16).astype(np.float32))
    np.testing.assert_array_equal(g["fresnel"] * 65535, np.round(g["fresnel"] * 65535))          # rgba16_unorm
    np.testing.assert_allclose(g["material_0"] * 255, np.round(g["material_0"] * 255), atol=1e-4)   # rgba8_unorm
    assert (np.round(g["material_0"][hit][:, 3] * 255) == 1).all()                               # MATERIAL_SURFACE_MODEL_LIT / 256 -> texel 1
    # depth -> position: the hit point reprojected with inv(proj_view) lies on the camera ray at the traced distance
    # (checked through the AO pass below, whose origin is that reconstruction)
    for half, frame in ((False, 3), (True, 0), (True, 1), (True, 2), (True, 3)):
        ao = ctx.trace_ao(cam, frame, depth, g["normal_roughness"], 0.5, 0.5, half)
        hao = hs.trace_ao(cam, W, H, frame, depth, g["normal_roughness"], 0.5, 0.5, half)
        np.testing.assert_array_equal(ao.view(np.uint32), hao.view(np.uint32))
        valid = ao[..., 1] == 1
        assert set(np.unique(ao[..., 1])) <= {0.0, 1.0} and valid.any()
        assert set(np.unique(ao[valid][:, 0])) <= {np.float32(np.float16(1 - k * 0.5 / 4)) for k in range(5)}   # 4 rays, strength 0.5, half store
        assert (ao[~valid][:, 0] == 1).all()
        if not half:
            np.testing.assert_array_equal(valid, hit)
            assert 0.02 < (ao[valid][:, 0] < 1).mean() < 0.9                                     # contact shadows exist, open floor is unoccluded
    far = ctx.trace_ao(cam, 3, depth, g["normal_roughness"], 50.0, 1.0, False)                   # longer rays can only occlude more
    near = ctx.trace_ao(cam, 3, depth, g["normal_roughness"], 0.5, 1.0, False)
    assert (far[..., 0] <= near[..., 0]).all() and (far[..., 0] < near[..., 0]).any()
    # fp16 state: the camera ray direction is a half, so the primary outputs move slightly
    d16, _ = ctx.render_primary(cam, 0, capi.Settings(max_bounces=4, state_precision=capi.STATE_REFERENCE_FP16))
    h16, _ = hs.render_primary(cam, W, H, 0, capi.Settings(max_bounces=4, state_precision=capi.STATE_REFERENCE_FP16))
    np.testing.assert_array_equal(d16.view(np.uint32), h16.view(np.uint32))
    assert (d16 != depth).any() and np.abs(d16 - depth)[hit & (d16 > 0)].max() < 0.05


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


def _random_rays(scene, n, seed):
    rng = np.random.default_rng(seed)
    lo, hi = scene.bounds
    rays = np.zeros(n, capi.RAY)
    rays["origin"] = rng.uniform(lo - 0.5, hi + 0.5, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d[: n // 8] = np.round(d[: n // 8])                                      # some axis-aligned / diagonal directions (zeros included)
    d[np.abs(d).sum(axis=1) == 0] = (0, 0, 1)
    rays["direction"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays["tmin"] = 0.001
    rays["tmax"] = rng.choice([100.0, 3.0, 0.5], n).astype(np.float32)
    return rays


@pytest.mark.parametrize("name", ["cornell", "small", "scene_basic"])
def test_wide_bvh_bit_exact(oracle, name):
    """The 4-wide quantised tree of merged mode (csrc/bpt_wide.cuh): its nodes equal the oracle's definition bit for bit, every
    decoded child box contains the exact one, and traversing it (with the exact leaf-box test) returns exactly the hits and
    visibilities of the binary traversal — on axis-aligned geometry (Cornell), any-hit materials (the reference's example
    scene) and grazing / axis-parallel rays."""
    scene = scenes.scene_basic(os.path.join(GOLDEN, "scene_basic.npz")) if name == "scene_basic" else _scene(name)
    ctx = oracle.OracleContext(8, 8); ctx.upload_scene(scene, capi.ACCEL_MERGED)
    hs = HC.HostScene(scene, ctx, capi.ACCEL_MERGED)
    wide, leafbox = hs.read_wide()
    owide, oleafbox = ctx.read_wide()
    np.testing.assert_array_equal(wide, owide); np.testing.assert_array_equal(leafbox.view(np.uint32), oleafbox.view(np.uint32))
    bv = ctx.read_bvh(0)
    n = bv["n"]
    assert wide.shape == (n - 1, 16)
    # structure: reachable wide nodes partition the leaves; decoded boxes are supersets of the binary child boxes
    origin = wide[:, 0:3].view(np.float32); ebits = wide[:, 3]
    scale = np.stack([np.ldexp(np.float32(1), ((ebits >> (8 * a)) & 0xff).astype(np.int32) - 127) for a in range(3)], axis=1).astype(np.float32)
    child = wide[:, 4:8].view(np.int32)
    qlo = np.stack([(wide[:, 8 + a][:, None] >> (8 * np.arange(4))) & 0xff for a in range(3)], axis=2).astype(np.float32)          # (node, child, axis)
    qhi = np.stack([(wide[:, [11, 12, 13][a]][:, None] >> (8 * np.arange(4))) & 0xff for a in range(3)], axis=2).astype(np.float32)
    dlo = origin[:, None, :] + qlo * scale[:, None, :]; dhi = origin[:, None, :] + qhi * scale[:, None, :]
    nodes = bv["nodes"]
    blo = np.stack([np.stack([nodes[f"c{c}_lo_{a}"] for a in "xyz"], axis=1) for c in (0, 1)], axis=1)                            # (node, 2, 3): child boxes held by the parent
    bhi = np.stack([np.stack([nodes[f"c{c}_hi_{a}"] for a in "xyz"], axis=1) for c in (0, 1)], axis=1)
    parent_of = {}                                                                                                                 # binary node / leaf ref -> (its exact box)
    for i in range(n - 1):
        for c, ref in enumerate((int(nodes["child0"][i]), int(nodes["child1"][i]))):
            parent_of[ref] = (blo[i, c], bhi[i, c])
    seen_leaves, stack, visited = [], [int(bv["root"])], 0
    while stack:
        i = stack.pop(); visited += 1
        kids = [int(c) for c in child[i] if c != 0x7ffffffe]
        assert 2 <= len(kids) <= 4
        for k, ref in enumerate(kids):
            elo, ehi = parent_of[ref]
            assert (dlo[i, k] <= elo).all() and (dhi[i, k] >= ehi).all(), (i, k)
            assert ((elo - dlo[i, k]) <= scale[i] * 1.0001).all() and ((dhi[i, k] - ehi) <= scale[i] * 1.0001).all()                # and by less than one grid step
            (stack if ref >= 0 else seen_leaves).append(ref if ref >= 0 else ~ref)
    assert sorted(seen_leaves) == list(range(n)) and visited < 0.62 * (n - 1)                                                      # every leaf once; about half the binary nodes
    # traversal: wide == binary, hits and visibility, bit for bit
    rays = _random_rays(scene, 6000, 21)
    cam = oracle.camera_matrices(scene.camera, 40, 30)                                                                             # + camera rays (coherent, hit the walls head-on)
    want_h, want_v = ctx.trace_rays(rays, 7), ctx.trace_shadow_rays(rays, 7)
    got_h, got_v = hs.trace_wide(rays, 7)
    for f in ("t", "u", "v", "instance", "primitive"):
        np.testing.assert_array_equal(got_h[f], want_h[f], err_msg=f)
    np.testing.assert_array_equal(got_v, want_v)
    assert (want_h["t"] >= 0).mean() > 0.1
    t, prim, cnt = ctx.wide_stats(rays, 4, True, 0)                                                                                # the oracle's own wide traversal agrees too
    if name == "cornell":                                                                                                          # (it ignores any-hit materials, which the other two scenes have)
        np.testing.assert_array_equal(t, want_h["t"])
    ctx.reset_counters(); ctx.trace_rays(rays, 7); st = ctx.stats()
    assert cnt["nodes"] < 0.7 * st.extend_nodes and cnt["leaf_boxes"] >= cnt["tris"]                                               # fewer steps; the leaf test filters
    del cam


@pytest.mark.parametrize("name", ["cornell", "small", "scene_basic"])
def test_wide_two_level_traversal_bit_exact(oracle, name):
    """Two-level mode through the 4-wide trees (wide TLAS whose proposed instances must pass their exact world box, then the
    instance's wide BLAS with the object-space ray — csrc/bpt_wide.cuh: trace_ray_wide_two_level, the steps k_trace_spec<.,true,true>
    interleaves): hits and visibilities equal the oracle's binary two-level traversal bit for bit."""
    scene = scenes.scene_basic(os.path.join(GOLDEN, "scene_basic.npz")) if name == "scene_basic" else _scene(name)
    ctx = oracle.OracleContext(8, 8); ctx.upload_scene(scene, capi.ACCEL_TWO_LEVEL)
    hs = HC.HostScene(scene, ctx, capi.ACCEL_TWO_LEVEL)
    rays = _random_rays(scene, 6000, 22)
    want_h, want_v = ctx.trace_rays(rays, 9), ctx.trace_shadow_rays(rays, 9)
    got_h, got_v = hs.trace_wide(rays, 9)
    for f in ("t", "u", "v", "instance", "primitive"):
        np.testing.assert_array_equal(got_h[f], want_h[f], err_msg=f)
    np.testing.assert_array_equal(got_v, want_v)
    assert (want_h["t"] >= 0).mean() > 0.1


def test_ddgi_volume_lighting_and_feedback_bit_exact(oracle):
    """The consumer of the probe atlases, calc_ddgi_volume_lighting (ddgi_lighting.hlsl:7-83), and the previous-update
    feedback of the probe lighting pass (ddgi/deferred_lighting.hlsl:102-115): CUDA source (host build) == oracle, and the
    properties the formula implies (constant atlas -> that constant; outside the volume -> 0; feedback only adds light)."""
    scene = _scene("small")
    table = scenes.ddgi_sample_randoms()
    vol = scenes.probe_volume(scene, (4, 3, 3), 64, ray_length=100.0)
    ctx = oracle.OracleContext(8, 8); ctx.upload_scene(scene, capi.ACCEL_MERGED)
    rays0 = ctx.trace_probes(vol, table, 0, 1)                                             # the reference's one bounce, no history yet
    irr0, vis0 = ctx.blend_probes(vol, table, 0, rays0)
    ctx.set_ddgi_volume(vol, irr0, vis0); HC.set_ddgi(vol, irr0, vis0)
    rng = np.random.default_rng(3)
    lo = np.float32(vol.base_position[:]); ext = np.float32(vol.extent[:])
    n = 600
    pos = (lo + rng.uniform(-0.15, 1.15, (n, 3)) * ext).astype(np.float32)                 # some outside
    nrm = rng.normal(size=(n, 3)); nrm = (nrm / np.linalg.norm(nrm, axis=1, keepdims=True)).astype(np.float32)
    view = rng.normal(size=(n, 3)); view = (view / np.linalg.norm(view, axis=1, keepdims=True)).astype(np.float32)
    a, b = ctx.ddgi_lighting(pos, nrm, view), HC.ddgi_lighting(pos, nrm, view)
    np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32))
    moved = pos + nrm * np.float32(0.2) + view * np.float32(0.8)
    inside = ((moved >= lo) & (moved <= lo + ext)).all(axis=1)
    margin = (np.abs(moved - lo) > 1e-4).all(axis=1) & (np.abs(moved - lo - ext) > 1e-4).all(axis=1)
    assert (a[~inside & margin] == 0).all() and (a[inside & margin][:, 3] <= 1).all() and inside.sum() > 50
    assert np.isfinite(a).all() and (a[..., :3] >= 0).all() and a[inside & margin][:, :3].max() > 0
    # constant atlases: every probe says (0.25, 0.5, 1.0) and visibility (mean 50, mean-square 2600) -> that colour back
    irr_c = np.zeros_like(irr0); irr_c[..., :3] = (0.25, 0.5, 1.0); irr_c[..., 3] = 1
    vis_c = np.zeros_like(vis0); vis_c[..., 0] = 50.0; vis_c[..., 1] = 2600.0
    ctx.set_ddgi_volume(vol, irr_c, vis_c)
    c = ctx.ddgi_lighting(pos, nrm, view)
    ok = inside & margin & (c[:, 3] == 1)
    np.testing.assert_allclose(c[ok][:, :3], np.broadcast_to(np.float32([0.25, 0.5, 1.0]), c[ok][:, :3].shape), rtol=3e-6)
    # feedback: probe paths with the previous update bound == host build, and never darker than without it
    ctx.set_ddgi_volume(vol, irr0, vis0)
    for bounces in (1, 2):
        fb = ctx.trace_probes(vol, table, 1, bounces)
        np.testing.assert_array_equal(fb.view(np.uint32), HC.HostScene(scene, ctx, capi.ACCEL_MERGED).trace_probes(vol, table, 1, bounces).view(np.uint32))
        ctx.set_ddgi_volume(None); HC.set_ddgi(None)
        plain = ctx.trace_probes(vol, table, 1, bounces)
        np.testing.assert_array_equal(plain.view(np.uint32), HC.HostScene(scene, ctx, capi.ACCEL_MERGED).trace_probes(vol, table, 1, bounces).view(np.uint32))
        assert (fb[:, :3] >= plain[:, :3]).all() and (fb[:, :3] > plain[:, :3]).any()
        np.testing.assert_array_equal(fb[:, 3], plain[:, 3])                                # hit distances unchanged
        ctx.set_ddgi_volume(vol, irr0, vis0); HC.set_ddgi(vol, irr0, vis0)
    HC.set_ddgi(None)
    with pytest.raises(capi.BptError):
        oracle.OracleContext(8, 8).ddgi_lighting(pos, nrm, view)                             # nothing bound


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


def test_exact_shortcuts_of_the_texture_fetch():
    """csrc/bpt_scene.cuh: unorm8_to_float(k) == (float)k / 255 for every k (two FMAs instead of an IEEE division); wrap_tc2 == two wrap_tc
    calls for every address mode, size (power of two or not) and coordinate, including far negative ones."""
    import ctypes as C
    H = HC.lib()
    H.hc_unorm8_to_float.argtypes, H.hc_unorm8_to_float.restype = [C.c_uint32], C.c_float
    H.hc_wrap_tc.argtypes, H.hc_wrap_tc.restype = [C.c_int, C.c_int, C.c_uint32], C.c_int
    H.hc_wrap_tc2.argtypes = [C.c_int, C.c_int, C.c_uint32, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    for k in range(256):
        assert np.float32(H.hc_unorm8_to_float(k)) == np.float32(k) / np.float32(255.0), k
    a, b = C.c_int(), C.c_int()
    for mode in (capi.ADDRESS_REPEAT, capi.ADDRESS_CLAMP):
        for n in (1, 2, 3, 5, 8, 16, 37, 64, 100, 256):
            for c in list(range(-3 * n - 2, 3 * n + 3)) + [-100000, 99999, -(1 << 30), (1 << 30) - 1]:
                H.hc_wrap_tc2(c, n, mode, C.byref(a), C.byref(b))
                assert (a.value, b.value) == (H.hc_wrap_tc(c, n, mode), H.hc_wrap_tc(c + 1, n, mode)), (mode, n, c)
