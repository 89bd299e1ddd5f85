"""CPU tests of the oracle itself: known-answer values, properties, and an independent numpy
brute-force check of its BVH + traversal. (`-m "not gpu"`)"""
import ctypes as C

import numpy as np
import pytest

from bisemutum_engine_b200 import capi, scenes

# SURVEY.md Appendix D — computed from shaders/core/utils/random.hlsl:3-26 (the only external pin)
RNG_KATS = [
    ((0, 0), 0x741C187D, (0xD2FFB8, 0x013FB7)),
    ((1, 0), 0x8DA6B311, (0xE7D13C, 0x6E7B6B)),
    ((0, 1), 0x70D3AEF1, (0x6BDB9C, 0xA5424B)),
    ((1037220, 1), 0x10109D9D, (0x688258, 0x44A1D7)),
    ((2073599, 276), 0x79330A80, (0xB57BDF, 0xC817B2)),
    ((123456, 42), 0xBC8A51F3, (0xD6EEB6, 0x3B969D)),
]


def test_rng_known_answers(oracle):
    L = oracle.library().lib
    for (a, b), tea, lcgs in RNG_KATS:
        assert L.obpt_rng_tea(a, b) == tea
        st = C.c_uint32(tea)
        assert L.obpt_rng_lcg(C.byref(st)) == lcgs[0]
        assert L.obpt_rng_lcg(C.byref(st)) == lcgs[1]


def test_rng_tea_matches_python_restatement(oracle):
    def tea(v0, v1):
        s0 = 0
        for _ in range(16):
            s0 = (s0 + 0x9E3779B9) & 0xFFFFFFFF
            v0 = (v0 + ((((v1 << 4) + 0xA341316C) ^ (v1 + s0) ^ ((v1 >> 5) + 0xC8013EA4)) & 0xFFFFFFFF)) & 0xFFFFFFFF
            v1 = (v1 + ((((v0 << 4) + 0xAD90777D) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7E95761E)) & 0xFFFFFFFF)) & 0xFFFFFFFF
        return v0
    rng = np.random.default_rng(1)
    L = oracle.library().lib
    for a, b in rng.integers(0, 2 ** 32, (200, 2)):
        assert L.obpt_rng_tea(int(a), int(b)) == tea(int(a), int(b))


def test_trig_polynomials_accuracy(oracle):
    L = oracle.library().lib
    s, c = C.c_float(), C.c_float()
    for u in np.linspace(0, 1, 4001).astype(np.float32):
        L.obpt_sincos_2pi(float(u), C.byref(s), C.byref(c))
        assert abs(s.value - np.sin(2 * np.pi * float(u))) < 3e-7 and abs(c.value - np.cos(2 * np.pi * float(u))) < 3e-7
    for x in np.linspace(-1, 1, 2001).astype(np.float32):
        assert abs(L.obpt_acos(float(x)) - np.arccos(float(x))) < 1e-6
    rng = np.random.default_rng(0)
    for y, x in rng.normal(size=(2000, 2)).astype(np.float32):
        assert abs(L.obpt_atan2(float(y), float(x)) - np.arctan2(float(y), float(x))) < 1e-6


def test_vndf_sample_properties(oracle):
    L = oracle.library().lib
    rng = np.random.default_rng(2)
    out = np.zeros(3, np.float32)
    for _ in range(500):
        v = rng.normal(size=3); v[2] = abs(v[2]) + 0.05; v = (v / np.linalg.norm(v)).astype(np.float32)
        rx, ry = rng.uniform(0.01, 1.0, 2)
        L.obpt_ggx_vndf_sample(v.ctypes.data_as(C.c_void_p), rx, ry, rng.uniform(), rng.uniform(), out.ctypes.data_as(C.c_void_p))
        assert abs(np.linalg.norm(out) - 1) < 1e-5 and out[2] >= 0
        assert np.dot(out, v) > -1e-5          # visible normals face the viewer


def test_bsdf_matches_float64_formula(oracle):
    """surface_eval_lit against a float64 numpy restatement of lit.hlsl:5-35."""
    L = oracle.library().lib
    rng = np.random.default_rng(3)

    def ref(N, T, V, Ld, base, f0, f90, rough, aniso):
        B = np.cross(N, T); H = (V + Ld) / np.linalg.norm(V + Ld)
        lh, lv, ll = (np.array([x @ T, x @ B, x @ N]) for x in (H, V, Ld))
        if lv[2] <= 0 or ll[2] <= 0:
            return np.zeros(3)
        fr = f0 + (f90 - f0) * (1 - max(V @ H, 0)) ** 5
        diff = (1 - fr) * base / np.pi * ll[2]
        a = np.sqrt(1 - 0.9 * aniso); r2 = rough * rough
        rx, ry = max(r2 / a, 1e-3), max(r2 * a, 1e-3)
        ndf = 1 / (np.pi * rx * ry * ((lh[0] / rx) ** 2 + (lh[1] / ry) ** 2 + lh[2] ** 2) ** 2)
        vv = ll[2] * np.sqrt((rx * lv[0]) ** 2 + (ry * lv[1]) ** 2 + lv[2] ** 2)
        lll = lv[2] * np.sqrt((rx * ll[0]) ** 2 + (ry * ll[1]) ** 2 + ll[2] ** 2)
        return diff + fr * ndf * (0.5 / max(vv + lll, 1e-4)) * ll[2]
    out = np.zeros(3, np.float32)
    for _ in range(300):
        N = np.array([0, 0, 1.0]); T = np.array([1.0, 0, 0])
        V = rng.normal(size=3); V[2] = abs(V[2]) + 0.1; V /= np.linalg.norm(V)
        Ld = rng.normal(size=3); Ld[2] = abs(Ld[2]) + 0.1; Ld /= np.linalg.norm(Ld)
        base, f0 = rng.uniform(0, 1, 3), rng.uniform(0, 1, 3); f90 = np.ones(3)
        rough, aniso = rng.uniform(0.2, 1), rng.uniform(0, 0.8)
        args = [np.asarray(a, np.float32) for a in (N, T, V, Ld, base, f0, f90)]
        L.obpt_surface_eval_lit(*[a.ctypes.data_as(C.c_void_p) for a in args], rough, aniso, out.ctypes.data_as(C.c_void_p))
        np.testing.assert_allclose(out, ref(*[a.astype(np.float64) for a in args], rough, aniso), rtol=2e-4, atol=1e-6)


def check_bvh_structure(b, prim_lo, prim_hi):
    n = b["n"]
    assert sorted(b["prims"].tolist()) == list(range(n))                 # a permutation
    assert (np.diff(b["morton"].astype(np.uint64)) >= 0).all() if n > 1 else True
    same = b["morton"][1:] == b["morton"][:-1]
    assert (b["prims"][1:][same] > b["prims"][:-1][same]).all()          # stable: ties keep ascending ids
    if n < 2:
        assert b["root"] == -1                                           # ~0: the single leaf
        return
    nodes = b["nodes"]
    seen_leaf, seen_node = np.zeros(n, bool), np.zeros(n - 1, bool)
    lo = np.zeros((n - 1, 3)); hi = np.zeros((n - 1, 3))

    def visit(i, parent):
        assert not seen_node[i]; seen_node[i] = True
        assert nodes["parent"][i] == parent
        boxes = []
        for k, ch in enumerate((nodes["child0"][i], nodes["child1"][i])):
            clo = np.array([nodes[f"c{k}_lo_{a}"][i] for a in "xyz"]); chi = np.array([nodes[f"c{k}_hi_{a}"][i] for a in "xyz"])
            if ch < 0:
                j = ~int(ch); assert not seen_leaf[j]; seen_leaf[j] = True
                p = b["prims"][j]
                np.testing.assert_array_equal(clo, prim_lo[p]); np.testing.assert_array_equal(chi, prim_hi[p])
            else:
                visit(int(ch), i)
                np.testing.assert_array_equal(clo, lo[ch]); np.testing.assert_array_equal(chi, hi[ch])
            boxes.append((clo, chi))
        lo[i] = np.minimum(boxes[0][0], boxes[1][0]); hi[i] = np.maximum(boxes[0][1], boxes[1][1])
    import sys
    sys.setrecursionlimit(10000)
    visit(0, -1)
    assert seen_leaf.all() and seen_node.all()


def world_triangles(scene):
    tris = []
    for inst in scene.instances:
        b = scene.blas[int(inst["blas"])]
        idx = scene.indices[int(b["index_offset"]): int(b["index_offset"]) + 3 * int(b["num_triangles"])]
        p = scene.positions[int(b["position_offset"]):].reshape(-1, 3)[idx].astype(np.float64)
        m = inst["transform"].astype(np.float64)
        tris.append((p @ m[:, :3].T + m[:, 3]).reshape(-1, 3, 3))
    return np.concatenate(tris)


@pytest.mark.parametrize("mode", [capi.ACCEL_TWO_LEVEL, capi.ACCEL_MERGED])
def test_lbvh_structure_and_bruteforce_traversal(oracle, mode):
    scene = scenes.small_test_scene(with_translucent=False)
    ctx = oracle.OracleContext(8, 8)
    ctx.upload_scene(scene, mode)
    if mode == capi.ACCEL_MERGED:
        b = ctx.read_bvh(0)
        # float32 world-space boxes as the oracle builds them are value-identical to a float32 numpy transform only
        # up to rounding order, so the structure check uses the boxes recorded in the leaves themselves
        assert b["n"] == scene.num_triangles
    else:
        for bi, bd in enumerate(scene.blas):
            b = ctx.read_bvh(bi)
            idx = scene.indices[int(bd["index_offset"]): int(bd["index_offset"]) + 3 * int(bd["num_triangles"])]
            p = scene.positions[int(bd["position_offset"]):].reshape(-1, 3)[idx].reshape(-1, 3, 3)
            check_bvh_structure(b, p.min(1), p.max(1))
        assert ctx.read_bvh(capi.BVH_TLAS)["n"] == len(scene.instances)
    # brute force: float64 Moller-Trumbore over every world triangle
    tris = world_triangles(scene)
    rng = np.random.default_rng(5)
    n = 400
    rays = np.zeros(n, capi.RAY)
    rays["origin"] = rng.uniform(-4, 4, (n, 3)) * [1, 0.3, 1] + [0, 2.0, 0]
    d = rng.normal(size=(n, 3)); rays["direction"] = d / np.linalg.norm(d, axis=1, keepdims=True)
    rays["tmin"], rays["tmax"] = 0.001, 100.0
    hits = ctx.trace_rays(rays)
    v0, e1, e2 = tris[:, 0], tris[:, 1] - tris[:, 0], tris[:, 2] - tris[:, 0]
    for i in range(n):
        O, D = rays["origin"][i].astype(np.float64), rays["direction"][i].astype(np.float64)
        pvec = np.cross(D, e2); det = np.einsum("ij,ij->i", e1, pvec)
        ok = np.abs(det) > 1e-12
        inv = np.where(ok, 1 / np.where(ok, det, 1), 0)
        tvec = O - v0; u = np.einsum("ij,ij->i", tvec, pvec) * inv
        qvec = np.cross(tvec, e1); v = (qvec @ D) * inv; t = np.einsum("ij,ij->i", e2, qvec) * inv
        m = ok & (u >= 0) & (v >= 0) & (u + v <= 1) & (t > 0.001) & (t <= 100)
        tb = t[m].min() if m.any() else -1.0
        if tb >= 0:   # (a float64 miss may be a float32 grazing hit on an edge: not asserted)
            assert hits["t"][i] > 0 and abs(hits["t"][i] - tb) <= 1e-3 * max(1.0, tb)


def test_morton_is_monotone_per_axis(oracle):
    L = oracle.library().lib
    lo, hi = np.zeros(3, np.float32), np.ones(3, np.float32)

    def code(c):
        c = np.asarray(c, np.float32)
        return L.obpt_morton63(c.ctypes.data_as(C.c_void_p), lo.ctypes.data_as(C.c_void_p), hi.ctypes.data_as(C.c_void_p))
    assert code([0, 0, 0]) == 0 and code([1, 1, 1]) == (1 << 63) - 1
    assert code([1, 0, 0]) > code([0, 1, 0]) > code([0, 0, 1])                 # x is the most significant axis
    xs = np.linspace(0, 1, 50)
    for axis in range(3):
        vals = [code(np.eye(3)[axis] * x) for x in xs]
        assert all(a <= b for a, b in zip(vals, vals[1:]))
    # degenerate extent → 0 on that axis
    hi0 = np.array([1, 0, 1], np.float32)
    c = np.array([0.5, 0.0, 0.5], np.float32)
    assert L.obpt_morton63(c.ctypes.data_as(C.c_void_p), lo.ctypes.data_as(C.c_void_p), hi0.ctypes.data_as(C.c_void_p)) & 0x2492492492492492 == 0


def test_render_invariants(oracle):
    """Size-independent properties: determinism, sample-sum linearity, NEE on/off consistency, counters."""
    scene = scenes.small_test_scene()
    W, H = 40, 30
    ctx = oracle.OracleContext(W, H)
    ctx.upload_scene(scene, capi.ACCEL_MERGED)
    cam = oracle.camera_matrices(scene.camera, W, H)
    st = capi.Settings(max_bounces=5)
    ctx.render(cam, 0, 2, st); a = ctx.resolve(1).copy()
    ctx.clear_accum(); ctx.render(cam, 0, 2, st)
    np.testing.assert_array_equal(a, ctx.resolve(1))                            # deterministic (threads do not matter)
    one = oracle.OracleContext(W, H, threads=1); one.upload_scene(scene, capi.ACCEL_MERGED)
    one.render(cam, 0, 2, st)
    np.testing.assert_array_equal(a, one.resolve(1))
    c = ctx.counters()
    assert c.extend_rays_per_bounce[1] == 2 * 2 * W * H and c.samples == 2 * 2 * W * H
    assert all(c.extend_rays_per_bounce[i] >= c.extend_rays_per_bounce[i + 1] for i in range(1, 5))
    assert c.extend_rays_per_bounce[5] == 0                                    # max_bounces 5 → 4 extend passes
    # unshadowed direct light is never darker than shadowed
    none = oracle.OracleContext(W, H); none.upload_scene(scene, capi.ACCEL_MERGED)
    none.render(cam, 0, 2, capi.Settings(max_bounces=2, nee_mode=capi.NEE_NONE))
    shad = oracle.OracleContext(W, H); shad.upload_scene(scene, capi.ACCEL_MERGED)
    shad.render(cam, 0, 2, capi.Settings(max_bounces=2))
    assert (none.resolve(2)[..., :3] >= shad.resolve(2)[..., :3] - 1e-6).all()
    assert none.counters().shadow_rays == 0 and shad.counters().shadow_rays > 0


def test_empty_and_edge_inputs(oracle):
    ctx = oracle.OracleContext(4, 4)
    with pytest.raises(capi.BptError):
        ctx.build_accel(capi.ACCEL_MERGED)                                      # nothing uploaded
    with pytest.raises(capi.BptError):
        ctx.render(capi.Camera(), 0, 1, capi.Settings())                        # render before build
    # single triangle, single instance: root is a leaf
    b = scenes.SceneBuilder("one_tri")
    m = b.add_material((0.8, 0.8, 0.8))
    P = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    N = np.tile(np.array([[0, 0, 1]], np.float32), (3, 1)); T = np.tile(np.array([[1, 0, 0, 1]], np.float32), (3, 1))
    b.add_drawable(b.add_mesh((P, N, T, P[:, :2].copy(), np.array([[0, 1, 2]], np.uint32))), m)
    sc = b.finish(dir_lights=scenes.dir_light((0, 0, 1)), camera=dict(position=(0.3, 0.3, 3), front_dir=(0, 0, -1), up_dir=(0, 1, 0), yfov=30, near_z=0.01, far_z=100))
    for mode in (capi.ACCEL_TWO_LEVEL, capi.ACCEL_MERGED):
        ctx.upload_scene(sc, mode)
        assert ctx.read_bvh(0)["root"] == -1 and ctx.read_bvh(0)["n"] == 1
        rays = np.zeros(3, capi.RAY)
        rays["origin"] = [(0.2, 0.2, 1), (2, 2, 1), (0.2, 0.2, 1)]; rays["direction"] = [(0, 0, -1)] * 3
        rays["tmin"] = 0.001; rays["tmax"] = [10, 10, 0.5]
        h = ctx.trace_rays(rays)
        assert h["t"][0] == 1.0 and h["t"][1] < 0 and h["t"][2] < 0             # hit, miss, beyond tmax
        assert h["primitive"][0] == 0 and h["instance"][0] == 0
    # max_bounces is clamped to [2, 16] (path_tracing.cpp:290)
    ctx.render(oracle.camera_matrices(sc.camera, 4, 4), 0, 1, capi.Settings(max_bounces=0))
    assert ctx.counters().extend_rays_per_bounce[1] == 16 and ctx.counters().extend_rays_per_bounce[2] == 0


def test_wide_tree_render_is_identical(oracle):
    """The oracle can walk the 4-wide quantised tree from any bounce on (what the CUDA kernels do from bounce 2): image, ray counts
    and queue contents are unchanged — only the work counters differ (fewer node steps, an exact leaf-box test per proposed leaf)."""
    from bisemutum_engine_b200 import capi, scenes
    scene = scenes.small_test_scene()
    W, H = 48, 32
    cam = oracle.camera_matrices(scene.camera, W, H)
    st = capi.Settings(max_bounces=6)
    a = oracle.OracleContext(W, H); a.upload_scene(scene, capi.ACCEL_MERGED)
    b = oracle.OracleContext(W, H); b.upload_scene(scene, capi.ACCEL_MERGED)
    b.set_wide_from_bounce(2)
    a.render(cam, 0, 3, st); b.render(cam, 0, 3, st)
    np.testing.assert_array_equal(a.resolve(3), b.resolve(3))
    ca, cb = a.counters(), b.counters()
    assert ca.extend_rays == cb.extend_rays and ca.shadow_rays == cb.shadow_rays
    sa, sb = a.stats(), b.stats()
    assert sa.extend_wide_nodes == 0 and sb.extend_wide_nodes > 0 and sb.extend_wide_rays == cb.extend_rays - W * H * 3
    # (the visiting ORDER differs, so the cull distance at the time a leaf is proposed does too: the number of triangle tests moves by a fraction of a percent)
    assert abs(sa.extend_tris - sb.extend_tris) < 0.02 * sa.extend_tris and sb.extend_leaf_boxes > 0
    per_ray_bin = (sa.extend_nodes - sb.extend_nodes) / sb.extend_wide_rays                       # binary steps those rays took before
    per_ray_wide = sb.extend_wide_nodes / sb.extend_wide_rays
    assert per_ray_wide < 0.7 * per_ray_bin
    b.set_wide_from_bounce(1); b.clear_accum(); b.render(cam, 0, 3, st)
    np.testing.assert_array_equal(a.resolve(3), b.resolve(3))
    two = oracle.OracleContext(W, H); two.upload_scene(scene, capi.ACCEL_TWO_LEVEL)
    with pytest.raises(capi.BptError):
        two.set_wide_from_bounce(2)                                                               # merged mode only


# ---- rect-light textures (lights.hlsl:425-447, shaders/core/mipmap.hlsl) against numpy float64 ----------------------------------------
def _mip_chain_f64(img, fmt, levels):
    """shaders/core/mipmap.hlsl:46-93 as generate_mipmaps_2d dispatches it, restated independently in float64 numpy."""
    def eotf(v):
        return np.where(v <= 0.04045, v / 12.92, ((v + 0.055) / 1.055) ** 2.4)

    def oetf(x):
        return np.where(x <= 0.0031308, x * 12.92, 1.055 * np.maximum(x, 0) ** (1 / 2.4) - 0.055)
    if fmt == capi.TEXTURE_RGBA32_FLOAT:
        cur = img.astype(np.float64)
    else:
        cur = img.astype(np.float64) / 255.0
        if fmt == capi.TEXTURE_RGBA8_SRGB:
            cur[..., :3] = eotf(cur[..., :3])
    out = [cur]
    full = int(np.floor(np.log2(max(img.shape[0], img.shape[1])))) + 1
    for _ in range(1, min(levels, full)):
        sh, sw = cur.shape[:2]
        dh, dw = max(sh // 2, 1), max(sw // 2, 1)
        pad = np.zeros((2 * dh + 3, 2 * dw + 3, 4)); pad[:sh, :sw] = cur            # reads past the level return 0
        ys, xs = 2 * np.arange(dh)[:, None], 2 * np.arange(dw)[None, :]
        acc = pad[ys, xs] + pad[ys + 1, xs] + pad[ys, xs + 1] + pad[ys + 1, xs + 1]
        num = 4
        if dw & 1:
            acc = acc + pad[ys, xs + 2] + pad[ys + 1, xs + 2]; num += 2
        if dh & 1:
            acc = acc + pad[ys + 2, xs] + pad[ys + 2, xs + 1]; num += 2
        if (dw & 1) and (dh & 1):
            acc = acc + pad[ys + 2, xs + 2]; num += 1
        nxt = acc / num
        if fmt != capi.TEXTURE_RGBA32_FLOAT:
            q = np.clip(nxt, 0, 1)
            if fmt == capi.TEXTURE_RGBA8_SRGB:
                q[..., :3] = eotf(np.rint(oetf(q[..., :3]) * 255) / 255)
                q[..., 3] = np.rint(q[..., 3] * 255) / 255
            else:
                q = np.rint(q * 255) / 255
            nxt = q
        out.append(nxt); cur = nxt
    return out


@pytest.mark.parametrize("fmt", [capi.TEXTURE_RGBA8_UNORM, capi.TEXTURE_RGBA8_SRGB, capi.TEXTURE_RGBA32_FLOAT])
@pytest.mark.parametrize("size", [(37, 22), (16, 16), (9, 5), (1, 7)])
def test_light_texture_mip_chain_matches_float64(oracle, fmt, size):
    t = scenes.light_texture(size[0], size[1], fmt, levels=16)
    ctx = oracle.OracleContext(8, 8)
    ctx.upload_light_textures([t])
    flat = ctx.read_light_texture(0)
    ref = _mip_chain_f64(t["texels"], fmt, 16)
    assert sum(r.shape[0] * r.shape[1] for r in ref) == len(flat)
    at = 0
    for lv, r in enumerate(ref):
        n = r.shape[0] * r.shape[1]
        got = flat[at: at + n].reshape(r.shape); at += n
        # 8-bit targets: equal up to ONE code where the float32 mean falls on the other side of a rounding boundary
        # (sRGB: one code is up to 2.4 / 255 in the linear domain)
        tol = 1e-6 if fmt == capi.TEXTURE_RGBA32_FLOAT else (1e-7 if lv == 0 else (2.4 / 255 if fmt == capi.TEXTURE_RGBA8_SRGB else 1.0 / 255))
        err = np.abs(got - r)
        assert err.max() <= tol * 1.0001, (lv, err.max())
        # such boundary cases are the exact ties (a mean of four 8-bit codes ends in .5 for one sum in four; float32 and float64 land on
        # different sides of it) — a minority, never a systematic offset
        assert (err > 1e-6).mean() < 0.12 and abs((got - r).mean()) < 0.1 * tol + 1e-7, (lv, (err > 1e-6).mean(), (got - r).mean())
    ctx.close()


def test_rect_light_texture_lookup_matches_float64(oracle):
    """The textured rect light against a float64 restatement of rect_light_sample_texture + trilinear SampleLevel: one lit point under a
    large textured light, diffuse lobe only (base colour white, f0 = 0 ... the specular term is removed by differencing two roughnesses)."""
    import os
    L = oracle.library().lib
    L.obpt_unit_rect_light.argtypes = [C.c_void_p, C.c_void_p] + [C.c_void_p] * 8 + [C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    luts = scenes.load_ltc_luts(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ltc_luts.npz"))
    sc = scenes.add_mixed_lights(scenes.small_test_scene(), 0, 1, luts, keep_dir_lights=True)
    tex = scenes.light_texture(64, 32, capi.TEXTURE_RGBA32_FLOAT, levels=7, mip_linear=1)
    rl = sc.rect_lights
    c = np.float32([0.0, 3.0, 0.0]); w, h = 2.0, 1.0
    rl[0]["center_position"] = c; rl[0]["normal"] = (0, -1, 0); rl[0]["emission"] = (1, 1, 1); rl[0]["two_sided"] = 0
    # facing down (-Y): local +X -> world +X, local +Y -> world +Z  (normal = X x Z' ... chosen so that the quad winds towards -Y)
    ex, ey = np.float32([1, 0, 0]), np.float32([0, 0, -1])
    rl[0]["position0"] = c + w * ex + h * ey; rl[0]["position1"] = c - w * ex + h * ey
    rl[0]["position2"] = c - w * ex - h * ey; rl[0]["position3"] = c + w * ex - h * ey
    rl[0]["inv_width_sqr"] = 1.0 / (2 * w) ** 2; rl[0]["inv_height_sqr"] = 1.0 / (2 * h) ** 2
    scenes.texture_rect_lights(sc, [tex], 2 * w, 2 * h)
    ctx = oracle.OracleContext(8, 8)
    ctx.upload_scene(sc, capi.ACCEL_MERGED)
    chain = _mip_chain_f64(tex["texels"], capi.TEXTURE_RGBA32_FLOAT, 7)

    def bilinear(img, u, v):
        hh, ww = img.shape[:2]
        x, y = u * ww - 0.5, v * hh - 0.5
        x0, y0 = int(np.floor(x)), int(np.floor(y)); fx, fy = x - x0, y - y0
        cl = lambda a, n: min(max(a, 0), n - 1)
        t = img[cl(y0, hh), cl(x0, ww)] * (1 - fx) + img[cl(y0, hh), cl(x0 + 1, ww)] * fx
        b = img[cl(y0 + 1, hh), cl(x0, ww)] * (1 - fx) + img[cl(y0 + 1, hh), cl(x0 + 1, ww)] * fx
        return (t * (1 - fy) + b * fy)[:3]

    def lookup(direction, P, roughness):
        n = rl[0]["normal"].astype(np.float64); p1, p2, p3 = (rl[0][k].astype(np.float64) for k in ("position1", "position2", "position3"))
        step = abs(direction @ n)
        t = abs((P - p2) @ n) / step
        rp = P + direction * t - p2
        u = np.clip(rp @ (p3 - p2) * float(rl[0]["inv_width_sqr"]), 0, 1); v = np.clip(rp @ (p1 - p2) * float(rl[0]["inv_height_sqr"]), 0, 1)
        level = np.log2(max(1.0, t * roughness * float(rl[0]["inv_texel_size"])))
        level = min(level, len(chain) - 1)
        l0 = int(np.floor(level)); l1 = min(l0 + 1, len(chain) - 1)
        return bilinear(chain[l0], u, v) * (1 - (level - l0)) + bilinear(chain[l1], u, v) * (level - l0)

    def form_factor_vector(P, quad):            # sum over edges of theta * normalised cross (float64, exact acos): light fully above the horizon here
        v = [(q - P) / np.linalg.norm(q - P) for q in quad]
        s = np.zeros(3)
        for i in range(4):
            a, b = v[i], v[(i + 1) % 4]
            cr = np.cross(a, b)
            s += np.arccos(np.clip(a @ b, -1, 1)) * cr / np.linalg.norm(cr)
        return s
    rng = np.random.default_rng(11)
    out = np.zeros(3, np.float32); mrp = np.zeros(3, np.float32)
    N, T, B = np.float32([0, 1, 0]), np.float32([1, 0, 0]), np.float32([0, 0, -1])      # B = N x T
    for _ in range(40):
        P = np.float32([rng.uniform(-1.5, 1.5), rng.uniform(0.2, 2.0), rng.uniform(-0.8, 0.8)])
        V = np.float32([0.1, 0.9, 0.2]); V /= np.linalg.norm(V)
        base, zero = np.ones(3, np.float32), np.zeros(3, np.float32)
        args = [a.ctypes.data_as(C.c_void_p) for a in (P, N, T, B, V, base, zero, zero)]
        L.obpt_unit_rect_light(ctx._h, rl[0:1].ctypes.data_as(C.c_void_p), *args, 0.7, 0.0, out.ctypes.data_as(C.c_void_p), mrp.ctypes.data_as(C.c_void_p))
        quad = [rl[0][k].astype(np.float64) for k in ("position3", "position2", "position1", "position0")]
        S = form_factor_vector(P.astype(np.float64), quad)           # in world space; local z = N.S
        integral = max(S @ N.astype(np.float64), 0.0)
        d = S / np.linalg.norm(S)
        np.testing.assert_allclose(mrp, d, atol=2e-3)                # the rational theta/sin(theta) fit bends the direction by < 0.1 degree
        expect = (1.0 / np.pi) * integral * lookup(d, P.astype(np.float64), 1.0)      # (1 - F) base / pi with f0 = f90 = 0
        np.testing.assert_allclose(out, expect, rtol=6e-3, atol=2e-4)
    ctx.close()


# ---- independent float64 pins of the largest transliterations (VERDICT r1 "what's weak" 1) ---------------------------------------------
def _clip_polygon_above_horizon(poly):
    """Sutherland-Hodgman against z >= 0 (an independent statement of what lights.hlsl:275-365's 16-case table does)."""
    out = []
    for i in range(len(poly)):
        a, b = poly[i], poly[(i + 1) % len(poly)]
        ia, ib = a[2] > 0, b[2] > 0
        if ia:
            out.append(a)
        if ia != ib:
            t = a[2] / (a[2] - b[2])
            out.append(a + t * (b - a))
    return out


def _polygon_integral_f64(P, N, T, B, quad, two_sided):
    """sum over the edges of the horizon-clipped polygon of theta_i * (normalised cross)_z: the analytic polygon form factor (x 2 pi)."""
    frame = np.stack([T, B, N])
    loc = [frame @ (q - P) for q in quad]
    poly = _clip_polygon_above_horizon(loc)
    if len(poly) < 3:
        return 0.0
    v = [p / np.linalg.norm(p) for p in poly]
    s = 0.0
    for i in range(len(v)):
        a, b = v[i], v[(i + 1) % len(v)]
        cr = np.cross(a, b); n = np.linalg.norm(cr)
        if n > 1e-14:
            s += np.arccos(np.clip(a @ b, -1, 1)) * cr[2] / n
    return abs(s) if two_sided else max(s, 0.0)


def test_ltc_diffuse_integral_is_the_polygon_form_factor(oracle):
    """ltc_integrate with the identity matrix (the diffuse lobe, lights.hlsl:383-423 + the 16-case clip table :275-365 + the rational
    theta/sin(theta) fit :366-382) against the closed-form polygon integral with an independent clipper — every clip configuration."""
    L = oracle.library().lib
    L.obpt_unit_ltc_integrate.argtypes = [C.c_void_p] * 4 + [C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_float), C.c_void_p]
    rng = np.random.default_rng(5)
    seen = set()
    integral = C.c_float(); mrp = np.zeros(3, np.float32)
    worst = 0.0
    for trial in range(4000):
        N = rng.normal(size=3); N /= np.linalg.norm(N)
        T = np.cross(N, rng.normal(size=3)); T /= np.linalg.norm(T)
        B = np.cross(N, T)
        P = rng.uniform(-1, 1, 3)
        c = P + rng.normal(size=3) * rng.uniform(0.3, 3.0)
        ex = rng.normal(size=3); ex /= np.linalg.norm(ex)
        ey = np.cross(ex, rng.normal(size=3)); ey /= np.linalg.norm(ey)
        w, h = rng.uniform(0.2, 2.5, 2)
        quad = [c + w * ex - h * ey, c - w * ex - h * ey, c - w * ex + h * ey, c + w * ex + h * ey]      # position3, 2, 1, 0 (lights.hlsl:462-467)
        zs = [float(N @ (q - P)) for q in quad]
        if min(abs(z) for z in zs) < 0.02:
            continue                                   # a corner on the horizon: the two clippers may legitimately classify it differently
        config = sum(1 << k for k in range(4) if zs[k] > 0)
        # an edge seen under nearly 180 degrees: the reference's fit evaluates pi * rsqrt(1 - x^2) - a/b at x -> -1, which cancels
        # catastrophically in FP32 (a property of lights.hlsl:374-376 itself, reproduced by the oracle; seen once in 4000 trials) — not compared
        frame = np.stack([T, B, N])
        pv = [p / np.linalg.norm(p) for p in _clip_polygon_above_horizon([frame @ (q - P) for q in quad])]
        if pv and min(pv[i] @ pv[(i + 1) % len(pv)] for i in range(len(pv))) < -0.9:
            continue
        seen.add(config)
        two_sided = trial & 1
        a32 = [np.float32(x) for x in (P, N, T, B)]
        Lq = np.float32(np.stack(quad))
        L.obpt_unit_ltc_integrate(*[a.ctypes.data_as(C.c_void_p) for a in a32], None, Lq.ctypes.data_as(C.c_void_p), two_sided, C.byref(integral), mrp.ctypes.data_as(C.c_void_p))
        expect = _polygon_integral_f64(P, N, T, B, quad, two_sided)
        worst = max(worst, abs(integral.value - expect))
        assert abs(integral.value - expect) <= 4e-3 + 2e-3 * abs(expect), (config, integral.value, expect)      # the fit's own error (lights.hlsl:366-376)
    assert seen >= {0, 1, 2, 3, 4, 6, 7, 8, 9, 11, 12, 13, 14, 15}, seen                                         # (5 and 10 cannot occur for a planar quad)
    assert worst > 0.0


def test_ltc_specular_symmetries_and_ggx_integral(oracle):
    """The specular LTC term (LUT addressing, the four quadrant flips + windings, the roughness swap, the matrix inverse;
    lights.hlsl:203-273,449-513), pinned two ways without restating it:
    (a) symmetry — an anisotropic GGX lobe is mirror-symmetric in the tangent and in the bitangent axis, so the view directions
        phi, 180 - phi, 180 + phi, 360 - phi (one per quadrant branch of get_ltc_matrix_and_brdf) with the light moved along must give the
        same radiance; swapping (rx, ry) and rotating everything by 90 degrees about N (the flip_roughness branch) as well. A wrong flip
        matrix, a missing winding or a transposed inverse breaks these by tens of percent.
    (b) magnitude — against a float64 quadrature of D * Vis * cos over the light (the BSDF of lit.hlsl): the ratio is about 2 pi
        (ltc_integrate returns sum.w, lights.hlsl:416 has no 1 / 2 pi), loosely: the reference's tables are not an exact fit of its BSDF."""
    import os
    L = oracle.library().lib
    L.obpt_unit_rect_light.argtypes = [C.c_void_p, C.c_void_p] + [C.c_void_p] * 8 + [C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    luts = scenes.load_ltc_luts(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ltc_luts.npz"))
    sc = scenes.add_mixed_lights(scenes.small_test_scene(), 0, 1, luts, keep_dir_lights=True)
    ctx = oracle.OracleContext(8, 8)
    ctx.upload_scene(sc, capi.ACCEL_MERGED)
    rl = sc.rect_lights[0:1].copy()
    N, T = np.array([0, 0, 1.0]), np.array([1.0, 0, 0]); B = np.cross(N, T)

    def alphas(rough, aniso):
        a = np.sqrt(1 - 0.9 * aniso); r2 = rough * rough
        return max(r2 / a, 1e-3), max(r2 * a, 1e-3)

    def ggx_spec(V, Ld, rough, aniso):
        rx, ry = alphas(rough, aniso)
        H = V + Ld; H = H / np.linalg.norm(H, axis=-1, keepdims=True)
        ndf = 1 / (np.pi * rx * ry * ((H[..., 0] / rx) ** 2 + (H[..., 1] / ry) ** 2 + H[..., 2] ** 2) ** 2)
        vv = Ld[..., 2] * np.sqrt((rx * V[0]) ** 2 + (ry * V[1]) ** 2 + V[2] ** 2)
        ll = V[2] * np.sqrt((rx * Ld[..., 0]) ** 2 + (ry * Ld[..., 1]) ** 2 + Ld[..., 2] ** 2)
        return ndf * 0.5 / np.maximum(vv + ll, 1e-4)

    def evaluate(phi_deg, rough, aniso, theta_deg=35.0, quadrature=False):
        th, ph = np.radians(theta_deg), np.radians(phi_deg)
        V = np.array([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)])
        R = 2 * (V @ N) * N - V                        # the light sits around the mirror direction, facing the shaded point
        c = 3.0 * R; ex = np.cross(R, N); ex /= np.linalg.norm(ex); ey = np.cross(R, ex)
        w, h = 0.9, 0.5
        for k, v in {"position0": c + w * ex + h * ey, "position1": c - w * ex + h * ey, "position2": c - w * ex - h * ey, "position3": c + w * ex - h * ey}.items():
            rl[0][k] = v
        rl[0]["normal"] = -R; rl[0]["emission"] = (1, 1, 1); rl[0]["two_sided"] = 1; rl[0]["texture_index"] = -1; rl[0]["center_position"] = c
        keep = [np.float32(x) for x in (np.zeros(3), N, T, B, V)]
        zero, one, out = np.zeros(3, np.float32), np.ones(3, np.float32), np.zeros(3, np.float32)
        # base colour 0 removes the diffuse term; f0 = f90 = 1 makes the specular factor brdf.x + brdf.y, the lobe's albedo
        L.obpt_unit_rect_light(ctx._h, rl.ctypes.data_as(C.c_void_p), *[k.ctypes.data_as(C.c_void_p) for k in keep], zero.ctypes.data_as(C.c_void_p),
                               one.ctypes.data_as(C.c_void_p), one.ctypes.data_as(C.c_void_p), rough, aniso, out.ctypes.data_as(C.c_void_p), None)
        if not quadrature:
            return float(out[0])
        n = 200
        s = (np.arange(n) + 0.5) / n * 2 - 1
        pts = c + (s[:, None, None] * w) * ex + (s[None, :, None] * h) * ey
        r2_ = (pts ** 2).sum(-1); Ld = pts / np.sqrt(r2_)[..., None]
        dw = (2 * w / n) * (2 * h / n) * np.abs(Ld @ (-R)) / r2_
        return float(out[0]), float((ggx_spec(V, Ld, rough, aniso) * Ld[..., 2] * dw * (Ld[..., 2] > 0)).sum())
    for rough, aniso in ((0.5, 0.0), (0.7, 0.6), (0.45, 0.8), (0.9, 0.3)):
        for phi in (20.0, 65.0):
            vals = [evaluate(p, rough, aniso) for p in (phi, 180.0 - phi, 180.0 + phi, 360.0 - phi)]
            assert min(vals) > 0 and max(vals) / min(vals) < 1.01, (rough, aniso, phi, vals)
    # flip_roughness (ry > rx happens for anisotropy < 0): equal to the swapped lobe seen from 90 degrees further round.
    # alphas(r, a) swapped = alphas(r', a') with r'^2 = r^2 and sqrt(1 - 0.9 a') = 1 / sqrt(1 - 0.9 a)
    for rough, aniso in ((0.6, 0.5), (0.8, 0.7)):
        aniso_swapped = (1 - 1 / (1 - 0.9 * aniso)) / 0.9
        assert np.allclose(alphas(rough, aniso_swapped), alphas(rough, aniso)[::-1])
        for phi in (25.0, 130.0, 250.0, 320.0):
            a, b = evaluate(phi, rough, aniso), evaluate(phi + 90.0, rough, aniso_swapped)
            assert abs(a / b - 1.0) < 0.01, (rough, aniso, phi, a, b)
    for rough, aniso, theta in ((0.7, 0.0, 35.0), (0.9, 0.0, 20.0), (0.7, 0.5, 35.0), (0.5, 0.0, 35.0)):
        got, quad = evaluate(30.0, rough, aniso, theta, quadrature=True)
        # the tables are addressed with alpha = roughness^2 / a (lights.hlsl:213) and track the engine's own BSDF only loosely:
        # measured 0.79 (roughness 0.5), 1.00 (0.7), 1.47 (0.9) x 2 pi — a property of the reference's data, the bounds only catch gross errors
        assert 0.6 < got / (2 * np.pi * quad) < 1.7, (rough, aniso, theta, got / quad)
    ctx.close()


def test_point_and_spot_attenuation_formula(oracle):
    """point_light_eval (lights.hlsl:14-25) against float64: windowed inverse square, spot cone ramp."""
    L = oracle.library().lib
    L.obpt_unit_point_light.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
    rng = np.random.default_rng(4)
    rad, d = np.zeros(3, np.float32), np.zeros(3, np.float32); dist = C.c_float()
    for k in range(500):
        pl = np.zeros(1, capi.POINT_LIGHT)
        pl["emission"] = rng.uniform(0.1, 5, 3); pl["position"] = rng.uniform(-2, 2, 3)
        rng_ = rng.uniform(1.0, 8.0); pl["range_sqr_inv"] = 1.0 / rng_ ** 2
        ax = rng.normal(size=3); ax /= np.linalg.norm(ax); pl["direction"] = ax
        if k % 2:
            outer = rng.uniform(20, 80); inner = rng.uniform(5, outer - 1)
            pl["cos_outer"], pl["cos_inner"] = np.cos(np.radians(outer)), np.cos(np.radians(inner))
        P = np.float32(rng.uniform(-3, 3, 3))
        L.obpt_unit_point_light(pl.ctypes.data_as(C.c_void_p), P.ctypes.data_as(C.c_void_p), rad.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p), C.byref(dist))
        lv = pl["position"][0].astype(np.float64) - P.astype(np.float64)
        d2 = lv @ lv; ld = lv / np.sqrt(d2)
        att = np.clip(1 - (d2 * float(pl["range_sqr_inv"][0])) ** 2, 0, 1) / max(d2, 0.001)
        ci, co = float(pl["cos_inner"][0]), float(pl["cos_outer"][0])
        if ci > co:
            ct = np.clip(ld @ pl["direction"][0].astype(np.float64), co, ci)
            att *= (ct - co) / max(ci - co, 0.001)
        np.testing.assert_allclose(rad, pl["emission"][0].astype(np.float64) * att, rtol=3e-5, atol=1e-7)
        np.testing.assert_allclose(d, ld, atol=2e-6); assert abs(dist.value - np.sqrt(d2)) < 1e-5


def test_cube_lookup_follows_cubemap_hlsl_and_is_seamless(oracle):
    """skybox.SampleLevel(dir): (a) a direction through the centre of texel (face, i, j) — built with cubemap_direction_from_layered_uv
    (core/utils/cubemap.hlsl:3-21), the function the reference FILLS its cube maps with — returns exactly that texel; (b) the lookup is
    continuous across cube edges and corners (seamless filtering, as Vulkan samples a TextureCube), which a per-face clamp is not."""
    L = oracle.library().lib
    L.obpt_unit_sample_sky.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    n = 8
    rng = np.random.default_rng(6)
    faces = rng.uniform(0, 1, (6, n, n, 4)).astype(np.float32)
    sc = scenes.small_test_scene(); sc.sky_faces = faces
    ctx = oracle.OracleContext(8, 8); ctx.upload_scene(sc, capi.ACCEL_MERGED)

    def direction(layer, u, v):                              # cubemap.hlsl:3-21
        u, v = 2 * u - 1, 2 * v - 1
        return np.array([[1, -v, -u], [-1, -v, u], [u, 1, v], [u, -1, -v], [u, -v, 1], [-u, -v, -1]][layer], np.float64)

    def sample(d):
        out = np.zeros(3, np.float32); d32 = np.float32(d)
        L.obpt_unit_sample_sky(ctx._h, d32.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        return out.astype(np.float64)
    for layer in range(6):
        for j in range(n):
            for i in range(n):
                np.testing.assert_allclose(sample(direction(layer, (i + 0.5) / n, (j + 0.5) / n) * 3.7), faces[layer, j, i, :3], atol=2e-6)
    # continuity: step across every edge at random places along it; the jump must be of the order of the step, not of the texel contrast
    eps = 1e-4
    worst = 0.0
    for layer in range(6):
        for edge in range(4):
            for _ in range(12):
                s = rng.uniform(0.02, 0.98)
                u, v = [(0.0, s), (1.0, s), (s, 0.0), (s, 1.0)][edge]
                du, dv = [(eps, 0), (-eps, 0), (0, eps), (0, -eps)][edge]
                inside = direction(layer, u + du, v + dv); outside = direction(layer, u - du, v - dv)       # outside: beyond the face -> re-projected by the lookup
                worst = max(worst, np.abs(sample(inside) - sample(outside)).max())
    assert worst < 0.02, worst                                # (a clamped per-face lookup jumps by the neighbouring texels' difference: ~0.3 here)
    for corner in ((1, 1, 1), (-1, 1, 1), (1, -1, -1), (-1, -1, 1)):
        c = np.array(corner, np.float64)
        vals = [sample(c + eps * np.array(o)) for o in ((3, -1, -1), (-1, 3, -1), (-1, -1, 3))]
        assert max(np.abs(vals[0] - vals[1]).max(), np.abs(vals[0] - vals[2]).max()) < 0.03
    ctx.close()


def test_hit_vertex_normal_is_the_inverse_transpose(oracle):
    """fetch_vertex_attributes (core/raytracing/hit.hlsl:27-164): world normal = normalize(transpose(inverse(M)) n) for a NON-uniformly
    scaled, sheared instance; tangent = normalize(M t); bitangent = normalize(cross(N, T)) * w; position = M p; barycentric attributes."""
    L = oracle.library().lib
    L.obpt_unit_hit_vertex.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, C.c_float, C.c_void_p]
    sc = scenes.small_test_scene()
    rng = np.random.default_rng(12)
    M = np.eye(3) + rng.normal(size=(3, 3)) * 0.4; M[0] *= 2.5; M[2] *= 0.4
    assert abs(np.linalg.det(M)) > 0.05
    slot = 1
    sc.instances["transform"][slot][:, :3] = M.astype(np.float32); sc.instances["transform"][slot][:, 3] = (0.3, -0.2, 0.9)
    sc.colors = rng.uniform(0, 1, sc.positions.size).astype(np.float32)          # one colour per vertex, same offsets as the positions
    sc.drawables["color_offset"] = sc.drawables["position_offset"]
    sc.drawable_va = sc.drawable_va | capi.VA_COLOR
    ctx = oracle.OracleContext(8, 8); ctx.upload_scene(sc, capi.ACCEL_TWO_LEVEL)
    inst = sc.instances[slot]; dr = sc.drawables[int(inst["instance_id_and_mask"]) & 0xffffff]
    M32 = inst["transform"].astype(np.float64)
    out = np.zeros(17, np.float32)
    ntri = int(sc.blas[int(inst["blas"])]["num_triangles"])
    for _ in range(60):
        prim = int(rng.integers(0, ntri)); u = float(rng.uniform(0, 1)); v = float(rng.uniform(0, 1 - u))
        assert L.obpt_unit_hit_vertex(ctx._h, slot, prim, u, v, out.ctypes.data_as(C.c_void_p)) == 0
        idx = sc.indices[int(dr["index_offset"]) + 3 * prim: int(dr["index_offset"]) + 3 * prim + 3].astype(int)
        bary = np.array([1 - u - v, u, v])
        nrm = (sc.normals[int(dr["normal_offset"]):].reshape(-1, 3)[idx].astype(np.float64) * bary[:, None]).sum(0)
        tan4 = (sc.tangents[int(dr["tangent_offset"]):].reshape(-1, 4)[idx].astype(np.float64) * bary[:, None]).sum(0)
        pos = (sc.positions[int(dr["position_offset"]):].reshape(-1, 3)[idx].astype(np.float64) * bary[:, None]).sum(0)
        uv = (sc.texcoords[int(dr["texcoord_offset"]):].reshape(-1, 2)[idx].astype(np.float64) * bary[:, None]).sum(0)
        A = M32[:, :3]
        Nw = np.linalg.inv(A).T @ nrm; Nw /= np.linalg.norm(Nw)
        Tw = A @ tan4[:3]; Tw /= np.linalg.norm(Tw)
        Bw = np.cross(Nw, Tw); Bw = Bw / np.linalg.norm(Bw) * tan4[3]
        np.testing.assert_allclose(out[0:3], Nw, atol=3e-5); np.testing.assert_allclose(out[3:6], Tw, atol=3e-5)
        np.testing.assert_allclose(out[6:9], Bw, atol=3e-5); np.testing.assert_allclose(out[9:12], A @ pos + M32[:, 3], atol=3e-5)
        np.testing.assert_allclose(out[12:14], uv, atol=1e-5)
        # vertex colour with the upstream quirk (hit.hlsl:108-112 reads the third corner at index.x): c0 + (c1 - c0) u + (c0 - c0) v
        col = sc.colors[int(dr["color_offset"]):].reshape(-1, 3)[idx].astype(np.float64)
        np.testing.assert_allclose(out[14:17], col[0] + (col[1] - col[0]) * u, atol=1e-6)
        assert np.abs(out[14:17] - (col * bary[:, None]).sum(0)).max() > 1e-4 or v < 0.02          # NOT the true barycentric colour
    ctx.close()


def test_log2_polynomial_accuracy(oracle):
    L = oracle.library().lib
    L.obpt_unit_log2.argtypes, L.obpt_unit_log2.restype = [C.c_float], C.c_float
    xs = np.concatenate([np.linspace(1, 2, 2001), np.exp(np.linspace(0, np.log(2.0 ** 24), 4001)), [1.0, 2.0, 4.0, 1.41421356, 1.41421357]]).astype(np.float32)
    err = max(abs(L.obpt_unit_log2(float(x)) - np.log2(float(x))) for x in xs)
    assert err < 3e-6, err
    assert L.obpt_unit_log2(1.0) == 0.0 and L.obpt_unit_log2(8.0) == 3.0
